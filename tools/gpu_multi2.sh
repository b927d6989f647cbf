#!/bin/bash
OUT=gpurun_out/${1:-m2}; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_fused_allreduce.py > $OUT/check.log 2>&1
grep -E "world|loss sums|rror|assert|File \"/tmp/code" $OUT/check.log | head -20
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "fine_pass" > $OUT/fine.log 2>&1; grep -E "Error|assert|Mismatch|Max|passed|failed" $OUT/fine.log | head -20

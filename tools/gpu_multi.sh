#!/bin/bash
# multi-GPU visit: fused all-reduce correctness (uneven shards) + bench at N ranks.  Usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh tag N'
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_fused_allreduce.py > $OUT/check.log 2>&1; echo "check exit $?"
grep -E "world|loss sums|Error|error" $OUT/check.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench exit $?"
python -c "
import json
d=json.load(open('$OUT/bench_${N}gpu.json'))
print('N=$N', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['device_ray_table']['value']), d['config']['collective'], d['config']['replicas_bit_identical'], 'render', d['render'] and round(d['render']['ms_per_frame'],2), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"
tail -3 $OUT/bench_${N}gpu.err

#!/bin/bash
# parity tests, then a full ncu capture of the three tensor-core kernels of one training step, then a short bench.
# Usage: gpurun -- 'bash tools/gpu_prof.sh tag'
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
python tools/profile_step.py 1024 500 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_(forward|bwd)_kernel|composite_loss' -s 3 -c 3 \
    -o $OUT/prof -f python tools/profile_step.py 1024 500 2 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err

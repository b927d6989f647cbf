#!/bin/bash
# quick GPU visit: parity tests + short bench.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag'
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err

#!/bin/bash
# One GPU-box visit: every GPU parity test (no -x: see all failures), then a short bench.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_visit.sh tag [pytest -k expr]'
TAG=${1:-visit}
KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -s -rfE --tb=short -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
else
  timeout 1200 python -m pytest tests -m gpu -q -s -rfE --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
fi
grep -E "passed|failed|error|FAILED|ERROR|exit" $OUT/pytest.log | tail -40
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err

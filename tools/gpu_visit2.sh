#!/bin/bash
TAG=${1:-visit}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rfE --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|error|FAILED|ERROR|exit" $OUT/pytest.log | tail -20
for lt in 256 128; do
NERFCA_LOSS_THREADS=$lt timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/bench_lt$lt.json 2> $OUT/bench_lt$lt.err; echo "bench exit $?"
python -c "
import json
d=json.load(open('$OUT/bench_lt$lt.json'))
print('loss_threads $lt', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'e2e', round(d['e2e']['value']), round(d['e2e']['device_ray_table']['value']))"
done

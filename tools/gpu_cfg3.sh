#!/bin/bash
OUT=gpurun_out/${1:-cfg3}; mkdir -p $OUT
timeout 900 python bench.py --config 3 --steps 60 --warmup 5 --no-cpu-baseline --no-eager-baseline > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "exit $?"
python -c "
import json
d=json.load(open('$OUT/bench_cfg3.json'))
print('cfg3', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['device_ray_table']['value']), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, d['render'] and d['render']['ms_per_frame'])"
tail -3 $OUT/bench_cfg3.err
timeout 600 python -m pytest tests -m gpu -q --tb=short -k "fine_pass or repack" 2>&1 | tail -3

#!/bin/bash
OUT=gpurun_out/r6f; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "config5 or layerwise" 2>&1 | tail -4
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b.json 2> $OUT/b.err
python -c "
import json
d=json.load(open('$OUT/b.json'))
print('cfg5', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'loss', d['config']['loss_last_step'])" || tail -5 $OUT/b.err
NERFCA_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file $OUT/launches_cfg5.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/ncu_bench.log 2>&1

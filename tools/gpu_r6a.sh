#!/bin/bash
OUT=gpurun_out/r6a; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "config5 or layerwise" 2>&1 | tail -5
for v in 0 1; do
NERFCA_WIDE_TMA=$v timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b$v.json 2> $OUT/b$v.err
python -c "
import json
d=json.load(open('$OUT/b$v.json'))
print('tma=$v', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'loss', d['config']['loss_last_step'])" || tail -5 $OUT/b$v.err
done

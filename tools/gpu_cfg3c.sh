#!/bin/bash
# 30-phase (config 3) check: parity tests that cover the latent fallback, then the backward's role split at config 3 and the config 2 line
OUT=gpurun_out/${1:-cfg3c}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --tb=short -k "30_phases or full_size_step_vs_oracle or trajectory" 2>&1 | tail -3
for sp in "29,45" "25,49" "22,52" "20,54"; do
NERFCA_BWD_SPLIT=$sp timeout 600 python bench.py --config 3 --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b.json 2> $OUT/b.err
python -c "
import json
d=json.load(open('$OUT/b.json'))
print('cfg3 split $sp', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"
done
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b2.json 2> $OUT/b2.err
python -c "
import json
d=json.load(open('$OUT/b2.json'))
print('cfg2', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"

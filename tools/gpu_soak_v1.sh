#!/bin/bash
# soak of the first-generation one-launch backward (NERFCA_BWD_V1=1) under graph replay + its bench line.  Usage: gpurun -- 'bash tools/gpu_soak_v1.sh tag [steps]'
TAG=${1:-soakv1}; OUT=gpurun_out/$TAG; mkdir -p $OUT; N=${2:-40000}
for rep in 1 2; do NERFCA_BWD_V1=1 timeout 400 python tools/soak.py $N 2>&1 | tail -2 | tee -a $OUT/soak_v1.log; done
NERFCA_BWD_V1=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b_v1.json 2> $OUT/b_v1.err
python -c "
import json
d=json.load(open('$OUT/b_v1.json'))
print('v1', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"

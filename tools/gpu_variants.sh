#!/bin/bash
# A/B of experiment builds (make -C nerf-ca_b200/csrc VARIANT=name EXTRA=-D...): bash tools/gpu_variants.sh tag "name1 name2 ..." [bench args]
OUT=gpurun_out/${1:-var}; mkdir -p $OUT
shift; VARS=$1; shift
for v in base $VARS; do
if [ $v = base ]; then L=libnerfca_b200.so; else L=libnerfca_b200_$v.so; fi
NERFCA_LIB=$L timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin "$@" > $OUT/b_$v.json 2> $OUT/b_$v.err
python -c "
import json
d=json.load(open('$OUT/b_$v.json'))
print('$v', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'loss', d['config'].get('loss_last_step'))" || tail -3 $OUT/b_$v.err
done

#!/bin/bash
# A/B of wait/poll variants (libnerfca_{b200,B,C,D}.so) x backward generation
OUT=gpurun_out/${1:-ab3}; mkdir -p $OUT
for lib in libnerfca_b200.so libnerfca_B.so libnerfca_C.so libnerfca_D.so; do
 for v1 in 0 1; do
  for sp in "30,44"; do
  NERFCA_LIB=$lib NERFCA_BWD_V1=$v1 NERFCA_BWD_SPLIT=$([ $v1 = 1 ] && echo "31,43" || echo $sp) timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/b.json 2> $OUT/b.err
  python -c "
import json
d=json.load(open('$OUT/b.json'))
print('$lib v1=$v1', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"
  done
 done
done

#!/bin/bash
# A/B of library VARIANT builds (make -C nerf-ca_b200/csrc VARIANT=name EXTRA=-D...) and runtime switches: short bench per item.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_var.sh tag "base fd rel:NERFCA_BWD_SPLIT=33,41 ..." [pytest -k expr] [extra bench args]'
# item = variant[:VAR=value[:VAR=value...]]
TAG=${1:-var}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for item in $2; do
  v=${item%%:*}; envs=""; [ "$item" != "$v" ] && envs=$(echo "${item#*:}" | tr ':' ' ')
  if [ "$v" = base ]; then lib=libnerfca_b200.so; else lib=libnerfca_b200_$v.so; fi
  name=$(echo $item | tr ':=,' '___')
  if [ -n "$3" ]; then env $envs NERFCA_LIB=$lib timeout 600 python -m pytest tests -m gpu -x -q -k "$3" 2>&1 | tail -2; fi
  for rep in 1 2; do
  env $envs NERFCA_LIB=$lib timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin $4 > $OUT/b_$name.json 2> $OUT/b_$name.err
  python -c "
import json
d=json.load(open('$OUT/b_$name.json'))
print('$item', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'loss', d['config']['loss_last_step'])" || tail -3 $OUT/b_$name.err
  done
done

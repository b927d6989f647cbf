#!/bin/bash
# timing-only A/B (no parity tests): bash tools/gpu_ab2.sh tag "VAR=v,VAR=v ..."   (use ':' for ',' inside a value)
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for c in $2; do
  echo "== $c"
  ( for kv in ${c//,/ }; do export "${kv//:/,}"; done
    timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('rays/s %.0f  ms/step %.4f  fwd %.1f us  bwd %.1f us  loss %.1f us' % (d['value'], d['ms_per_step'], k['field_forward']['ms_per_step']*1e3, k['field_backward']['ms_per_step']*1e3, k['integral_loss']['ms_per_step']*1e3))" )
done 2>&1 | tee $OUT/ab.txt

#!/bin/bash
# A/B of runtime switches of the tensor-core kernels: for each combination (comma-separated VAR=value list) the full-size property /
# composite-step parity tests and the step time.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_ab.sh tag "NERFCA_BWD_MERGED=1 NERFCA_BWD_MERGED=0 ..."'
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for c in ${2:-NERFCA_BWD_MERGED=1 NERFCA_BWD_MERGED=0}; do
  echo "== $c"
  ( export ${c//,/ }
    timeout 300 python -m pytest tests -m gpu -x -q -k "full_size or composite_step" 2>&1 | tail -2
    timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('rays/s %.0f  ms/step %.4f  fwd %.1f us  bwd %.1f us  loss %.1f us' % (d['value'], d['ms_per_step'], k['field_forward']['ms_per_step']*1e3, k['field_backward']['ms_per_step']*1e3, k['integral_loss']['ms_per_step']*1e3))" )
done 2>&1 | tee $OUT/ab.txt

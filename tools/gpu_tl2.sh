#!/bin/bash
# bwd2 timeline of one top-role CTA (developer build): bash tools/gpu_tl2.sh tag [cta]
OUT=gpurun_out/${1:-tl2}; mkdir -p $OUT
NERFCA_LIB=libnerfca_b200_tl.so NERFCA_TIMELINE=${3:-top} NERFCA_TIMELINE_CTA=${2:-0} timeout 120 python tools/profile_step.py 1024 500 2 > $OUT/tl.log 2>&1
grep -c "^TL" $OUT/tl.log; grep -v "^TL" $OUT/tl.log | tail -5

#!/bin/bash
# bwd2 timeline of one CTA (developer build, make TL=1) at n_phases: bash tools/gpu_tl3.sh tag role cta n_phases [split]
OUT=gpurun_out/${1:-tl3}; mkdir -p $OUT
[ -n "$5" ] && export NERFCA_BWD_SPLIT=$5
NERFCA_LIB=libnerfca_b200_tl.so NERFCA_TIMELINE=${2:-bot} NERFCA_TIMELINE_CTA=${3:-100} timeout 120 python tools/profile_step.py 1024 500 2 ${4:-10} > $OUT/tl_$2_$4.log 2>&1
grep -c "^TL" $OUT/tl_$2_$4.log; grep -v "^TL" $OUT/tl_$2_$4.log | tail -3

#!/bin/bash
# kernel event timeline of one CTA (needs a library built with EXTRA=-DNERFCA_TIMELINE_BUILD): bash tools/gpu_tl.sh tag fwd|bot [cta]
OUT=gpurun_out/${1:-tl}; mkdir -p $OUT
NERFCA_TIMELINE=${2:-fwd} NERFCA_TIMELINE_CTA=${3:-0} python tools/profile_step.py 1024 500 2 > $OUT/tl.log 2>&1
grep -c "^TL" $OUT/tl.log

#!/bin/bash
OUT=gpurun_out/${1:-tl}; mkdir -p $OUT
NERFCA_TIMELINE=1 python tools/profile_step.py 1024 500 2 > $OUT/tl.log 2>&1
grep -c "^TL" $OUT/tl.log

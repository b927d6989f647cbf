#!/bin/bash
# top- and bottom-role timelines of a developer build at several role splits: bash tools/gpu_tl2v.sh tag lib "31,43 33,41"
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT; lib=$2
for sp in $3; do
n=$(echo $sp | tr ',' '_')
NERFCA_BWD_SPLIT=$sp NERFCA_LIB=$lib NERFCA_TIMELINE=bot NERFCA_TIMELINE_CTA=80 python tools/profile_step.py 1024 500 2 > $OUT/tl_bot_$n.log 2>&1
NERFCA_BWD_SPLIT=$sp NERFCA_LIB=$lib NERFCA_TIMELINE=top NERFCA_TIMELINE_CTA=0 python tools/profile_step.py 1024 500 2 > $OUT/tl_top_$n.log 2>&1
grep -c "^TL" $OUT/tl_bot_$n.log $OUT/tl_top_$n.log
done

#!/bin/bash
# full ncu capture of selected kernels of one training step: bash tools/gpu_ncu_fwd.sh tag regex
OUT=gpurun_out/${1:-ncu}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${2:-tc_forward}" -s 2 -c ${3:-1} \
    -o $OUT/prof -f python tools/profile_step.py 1024 500 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log

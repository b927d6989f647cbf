#!/bin/bash
OUT=gpurun_out/${1:-ncu_bwd2}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_bwd2_kernel|tc_forward_kernel|composite_loss' -s 3 -c 3 -o $OUT/prof -f python tools/profile_step.py 1024 500 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log; ls -la $OUT

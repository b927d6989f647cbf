#!/bin/bash
# run full-size steps with the developer (timeline) build, which prints which bounded wait gave up: bash tools/gpu_dbgrun.sh [n_steps]
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
timeout 200 python tools/profile_step.py 1024 500 ${1:-3} 2>&1 | grep -v "^TL" | sort | uniq -c | sort -rn | head -${2:-12}
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

#!/bin/bash
# run the full-size step once with the developer (timeline) build, which prints which bounded wait gave up
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
NERFCA_BWD_MERGED=${1:-0} timeout 120 python tools/profile_step.py 1024 500 1 2>&1 | grep -v "^TL" | sort | uniq -c | sort -rn | head -30
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

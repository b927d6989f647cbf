#!/bin/bash
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
timeout 300 python tools/soak.py ${1:-20000} 2>&1 | grep -v "^TL" | sort | uniq -c | sort -rn | head -${2:-14} | cut -c1-220
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

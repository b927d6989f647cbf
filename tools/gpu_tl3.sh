#!/bin/bash
TAG=${1:-tl3}
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
NERFCA_BWD_MERGED=0 bash tools/gpu_tl.sh $TAG/sep_top top 0
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

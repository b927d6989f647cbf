#!/bin/bash
# event timelines of one forward / top-role / bottom-role CTA with the developer build (make -C nerf-ca_b200/csrc TL=1)
# Usage: gpurun -- 'bash tools/gpu_tlall.sh tag'
TAG=${1:-tl}
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
bash tools/gpu_tl.sh $TAG/tl_top top 0; bash tools/gpu_tl.sh $TAG/tl_bot bot 74; bash tools/gpu_tl.sh $TAG/tl_fwd fwd 0
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

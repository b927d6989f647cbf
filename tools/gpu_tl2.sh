#!/bin/bash
# timelines with alternative libraries: bash tools/gpu_tl2.sh tag
TAG=${1:-tl2}
cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so
if [ -f nerf-ca_b200/libnerfca_b200_old_tl.so ]; then
  cp nerf-ca_b200/libnerfca_b200_old_tl.so nerf-ca_b200/libnerfca_b200.so
  bash tools/gpu_tl.sh $TAG/old_top top 0; bash tools/gpu_tl.sh $TAG/old_bot bot 0
fi
cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
NERFCA_BWD_MERGED=0 bash tools/gpu_tl.sh $TAG/sep_top top 0; NERFCA_BWD_MERGED=0 bash tools/gpu_tl.sh $TAG/sep_bot bot 0
bash tools/gpu_tl.sh $TAG/mrg_top top 0; bash tools/gpu_tl.sh $TAG/mrg_bot bot 74
cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so

#!/bin/bash
# parity tests, the training-step time for a sweep of top/bottom role splits of the backward kernel, a short bench, and (when the
# timeline build libnerfca_b200_tl.so is present) the event timelines of one forward / top-role / bottom-role CTA.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_split.sh tag "37,37 34,40 ..."'
TAG=${1:-split}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
for s in ${2:-37,37}; do
  echo "== split $s"; NERFCA_BWD_SPLIT=$s timeout 120 python tools/time_fields.py 1024 500 bf16 2>&1 | grep "train step"
done | tee $OUT/split.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-render > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ -f nerf-ca_b200/libnerfca_b200_tl.so ]; then
  cp nerf-ca_b200/libnerfca_b200.so /tmp/lib_keep.so; cp nerf-ca_b200/libnerfca_b200_tl.so nerf-ca_b200/libnerfca_b200.so
  bash tools/gpu_tl.sh $TAG/tl_top top 0; bash tools/gpu_tl.sh $TAG/tl_bot bot 74; bash tools/gpu_tl.sh $TAG/tl_fwd fwd 0
  cp /tmp/lib_keep.so nerf-ca_b200/libnerfca_b200.so
fi

#!/bin/bash
OUT=gpurun_out/${1:-dbg4}; mkdir -p $OUT
python - > $OUT/fp32.log 2>&1 <<'PY'
import sys; sys.path[:0]=['.','nerf-ca_b200','nerf-ca_b200/train','tests']
import parity
for k in range(4):
    try:
        print(parity.run_composite_step_parity(n_rays=96, n_depth=77, precision='fp32', seed=5, fused=True))
    except AssertionError as e:
        print("FAIL", e)
for seed in (1,2,3):
    try:
        print(seed, parity.run_composite_step_parity(n_rays=96, n_depth=77, precision='fp32', seed=seed, fused=True))
    except AssertionError as e:
        print("FAIL", seed, e)
PY
cat $OUT/fp32.log | cut -c1-400

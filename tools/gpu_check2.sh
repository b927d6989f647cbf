#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list of the bench command, one full ncu capture of the
# tensor-core kernels + the fused loss kernel.
# Usage (from the repo root): gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
NERFCA_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render > $OUT/ncu_bench.log 2>&1
python tools/launch_shares.py $OUT/launches.csv > $OUT/launch_shares.txt 2>&1; cat $OUT/launch_shares.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_(forward|bwd2?)_kernel|composite_loss|adam' -s 6 -c 3 \
    -o $OUT/prof -f python tools/profile_step.py 1024 500 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT

#!/bin/bash
# One full GPU-box visit = what a profiles/<tag>/ directory holds: every GPU parity test, the bench line (both arms) of config 2 and
# the config 3 / config 5 lines, a soak run, the ncu launch list of the bench command, one full ncu capture of the step's kernels, the
# same two for the layer-wise path, and a compute-sanitizer memcheck run.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_visit.sh tag [soak_steps]'
# Afterwards here: python tools/ncu_summary.py / tools/make_traffic.py / tools/sass_summary.py, copy into profiles/<tag>/.
TAG=${1:-run}; SOAK=${2:-70000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 200 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
for cfg in 3 5; do
  timeout 900 python bench.py --config $cfg --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline > $OUT/bench_cfg$cfg.json 2> $OUT/bench_cfg$cfg.err; echo "bench cfg$cfg exit $?"
  python -c "
import json
d=json.load(open('$OUT/bench_cfg$cfg.json'))
print('cfg$cfg', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'frac', round(d['roofline']['frac'],3))"
done
for rep in 1 2; do timeout 600 python tools/soak.py $SOAK 2>&1 | tail -1 | tee -a $OUT/soak.log; done
NERFCA_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/ncu_bench.log 2>&1
python tools/launch_shares.py $OUT/launches.csv > $OUT/launch_shares.txt 2>&1; cat $OUT/launch_shares.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_(forward|bwd2?)_kernel|composite_loss' -s 3 -c 3 \
    -o $OUT/prof -f python tools/profile_step.py 1024 500 3 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
# layer-wise path (config 5): launch list + full capture of the GEMM kernels, then compute-sanitizer memcheck of both paths
NERFCA_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv \
    --log-file $OUT/launches_cfg5.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/ncu_bench_cfg5.log 2>&1
python tools/launch_list.py $OUT/launches_cfg5.csv > $OUT/launch_shares_cfg5.txt 2>&1; head -6 $OUT/launch_shares_cfg5.txt
timeout 800 ncu --set full --clock-control none --import-source on -k regex:wide_gemm2 -s 8 -c 5 -o $OUT/wide -f \
    python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/ncu_wide.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q \
    -k "config5_widened_stress_shapes and bf16 or layerwise_tensor_core_path_other_shapes or composite_step_vs_oracle or trainer_step_from_ids or composite_step_30_phases or fused_render" > $OUT/memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck.log
ls -la $OUT

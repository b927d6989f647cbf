#!/bin/bash
# One full GPU-box visit = what a profiles/<tag>/ directory holds: every GPU parity test, the bench line (both arms) of config 2 and
# the config 3 / config 5 lines, a soak run, the ncu launch list of the bench command and one full ncu capture of the step's kernels.
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
ls -la $OUT

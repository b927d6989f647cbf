#!/bin/bash
# backward v2 bring-up: the parity tests that exercise the backward, then benches of v2 and v1 side by side
TAG=${1:-bwd2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s -rfE --tb=short -x -k "composite_step or static_step or full_size or trajectory or graph_replay or 30_phases or soak" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|error|FAILED|ERROR|exit|Error|timeout" $OUT/pytest.log | tail -30
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/bench_v2.json 2> $OUT/bench_v2.err; echo "bench v2 exit $?"
python -c "
import json,sys
d=json.load(open('$OUT/bench_v2.json'))
print('v2', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'e2e', d['e2e']['value'])"
tail -3 $OUT/bench_v2.err
NERFCA_GRAPH=0 NERFCA_BWD_V1=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/bench_v1.json 2> $OUT/bench_v1.err; echo "bench v1 exit $?"
python -c "
import json,sys
d=json.load(open('$OUT/bench_v1.json'))
print('v1', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, 'e2e', d['e2e']['value'])"
for sp in "25,49" "27,47" "29,45" "31,43" "33,41"; do
NERFCA_BWD_SPLIT=$sp timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-render --no-dropin > $OUT/bench_s.json 2> $OUT/bench_s.err
python -c "
import json,sys
d=json.load(open('$OUT/bench_s.json'))
print('split $sp', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})"
done

#!/bin/bash
# Scaling visit on N GPUs of one box: config 3 weak + strong (8192 global rays) and config 2 weak, each with the sharded render leg.
# Usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh TAG N [check]'
TAG=${1:-scale}; N=${2:-2}; CHECK=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ $N -gt 1 ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540"; else LAUNCH=python; fi
if [ -n "$CHECK" ] && [ $N -gt 1 ]; then
  timeout 600 $LAUNCH tools/check_fused_allreduce.py > $OUT/check_${N}gpu.log 2>&1; echo "check exit $?"
  grep -E "world|loss sums" $OUT/check_${N}gpu.log | tail -4
fi
run() {
  name=$1; shift
  timeout 900 $LAUNCH bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-dropin "$@" > $OUT/${name}_${N}gpu.json 2> $OUT/${name}_${N}gpu.err; echo "$name exit $?"
  python -c "
import json
d=json.load(open('$OUT/${name}_${N}gpu.json'))
print('$name N=$N', d['scaling'], round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'ids', round(d['e2e']['device_ray_table']['value']), d['config']['collective'], d['config']['replicas_bit_identical'], 'render', d['render'] and round(d['render']['ms_per_frame'],2), {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()})" || tail -3 $OUT/${name}_${N}gpu.err
}
run cfg3_weak --config 3 --no-render
run cfg3_strong8192 --config 3 --strong 8192 --no-render
run cfg2_weak

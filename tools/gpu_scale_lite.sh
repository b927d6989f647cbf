#!/bin/bash
# weak-scaling lines only (config 2 with the sharded render leg, config 3): gpurun --gpus N -- 'bash tools/gpu_scale_lite.sh TAG N'
TAG=${1:-scale}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540"
run() {
  name=$1; shift
  timeout 900 $LAUNCH bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-dropin "$@" > $OUT/${name}_${N}gpu.json 2> $OUT/${name}_${N}gpu.err; echo "$name exit $?"
  python -c "
import json
d=json.load(open('$OUT/${name}_${N}gpu.json'))
print('$name N=$N', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'ids', round(d['e2e']['device_ray_table']['value']), d['config']['collective'], d['config']['replicas_bit_identical'], 'render', d['render'] and round(d['render']['ms_per_frame'],2))" || tail -3 $OUT/${name}_${N}gpu.err
}
run cfg2_weak
run cfg3_weak --config 3 --no-render

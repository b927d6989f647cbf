"""Ad-hoc timing of the field kernels (CUDA events, current stream).  Usage: python tools/time_fields.py [n_rays] [n_depth]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import parity  # noqa: E402
from oracle import nerfca_oracle as orc  # noqa: E402
from nerfca import ops  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_depth = int(sys.argv[2]) if len(sys.argv) > 2 else 500
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "bf16"]
dev = "cuda:0"
sd_s = orc.init_field_state(75, 128, 4, seed=1)
sd_d = orc.init_field_state(83, 128, 4, 10, 8, seed=2)
mask, _ = orc.freq_mask(12, 75000, 150000, 1)
rays, phases, z = parity.synthetic_batch(n_rays, n_depth, seed=9)
rays, phases, z = rays.to(dev), phases.to(dev), z.to(dev)
i0 = torch.full((n_rays,), parity.I0, device=dev)
w = orc.schedule_weights(50000, orc.COMPOSITE_HP)
lc = ops.LossConfig(w["favor_s"], w["dyn_entro"], w["occl"], w["l1"], 1e-4, 0.03, True, n_rays)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for prec in modes:
    s, t = parity.build_models(sd_s, sd_d, dev, prec, mask=mask)
    smp = ops.Samples.from_rays(rays[:, 0, :], rays[:, 1, :], z, phases)

    def fwd():
        with torch.no_grad():
            s.forward_rays(smp); t.forward_rays(smp)
    ms = timeit(fwd)
    print(f"{prec}: forward both fields {ms:.3f} ms  -> {n_rays / ms * 1e3:.3e} rays/s (fwd only), "
          f"{n_rays * n_depth * 303104 / ms / 1e9:.1f} TFLOP/s", flush=True)
    try:
        def step():
            for p in list(s.parameters()) + list(t.parameters()):
                p.grad = None
            ops.train_step_composite(s, t, rays, phases, i0, z, "softplus", lc)
        ms = timeit(step)
        print(f"{prec}: fused train step {ms:.3f} ms -> {n_rays / ms * 1e3:.3e} rays/s, {n_rays * n_depth * 870912 / ms / 1e9:.1f} TFLOP/s",
              flush=True)
    except NotImplementedError as e:
        print(prec, "step unavailable:", e)

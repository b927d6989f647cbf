"""Fused composite training steps at the bench size (for ncu / the in-kernel timeline): python tools/profile_step.py [n_rays] [n_depth] [n_steps] [n_phases]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for p in (ROOT, os.path.join(ROOT, "nerf-ca_b200"), os.path.join(ROOT, "nerf-ca_b200", "train"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity  # noqa: E402
from oracle import nerfca_oracle as orc  # noqa: E402
from nerfca import ops  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_depth = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n_phases = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = "cuda:0"
sd_s = orc.init_field_state(75, 128, 4, seed=1)
sd_d = orc.init_field_state(83, 128, 4, n_phases, 8, seed=2)
mask, _ = orc.freq_mask(12, 75000, 150000, 1)
rays, phases, z = parity.synthetic_batch(n_rays, n_depth, seed=9, n_phases=n_phases)
rays, phases, z = rays.to(dev), phases.to(dev), z.to(dev)
i0 = torch.full((n_rays,), parity.I0, device=dev)
w = orc.schedule_weights(50000, orc.COMPOSITE_HP)
lc = ops.LossConfig(w["favor_s"], w["dyn_entro"], w["occl"], w["l1"], 1e-4, 0.03, True, n_rays)
s, t = parity.build_models(sd_s, sd_d, dev, "bf16", mask=mask)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(n_steps):
    e0.record()
    ops.train_step_composite(s, t, rays, phases, i0, z, "softplus", lc)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {k}: {e0.elapsed_time(e1):.3f} ms", flush=True)

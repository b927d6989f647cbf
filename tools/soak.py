"""Soak test of the training step: python tools/soak.py [n_steps] [n_phases] -- many steps on rotating batches, reports the first failure.
(n_phases > 12 exercises the second-generation backward with the latent fallback, i.e. config 3.)"""
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for p in (ROOT, os.path.join(ROOT, "nerf-ca_b200"), os.path.join(ROOT, "nerf-ca_b200", "train"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity  # noqa: E402
from nerfca import trainer as tr  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_phases = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
torch.manual_seed(0)
t = tr.CompositeTrainer.from_config(device=dev, precision="bf16", n_depth=500, n_phases=n_phases)
t.set_iteration(50000)
batches = []
for k in range(8):
    rays, phases, z = parity.synthetic_batch(1024, 500, seed=200 + k, n_phases=n_phases)
    batches.append((rays.to(dev), phases.to(dev).int(), z.to(dev)))
t0 = time.time()
done = 0
try:
    for k in range(n_steps):
        t.step_device(*batches[k % 8])
        if (k + 1) % 2000 == 0:
            torch.cuda.synchronize()
            done = k + 1
    torch.cuda.synchronize()
    print(f"soak ok ({n_phases} phases): {n_steps} steps in {time.time() - t0:.1f} s, loss terms finite {bool(torch.isfinite(t.last_terms).all())}", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"soak FAILED after >= {done} steps: {type(e).__name__}: {str(e)[:200]}", flush=True)
    sys.exit(1)

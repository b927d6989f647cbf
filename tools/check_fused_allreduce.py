"""N-GPU check of nerfca_allreduce_adam_step (gradient sum fused with the optimizer step over NVLink peer memory) against the
NCCL all-reduce + nerfca_adam_step path.  Run:  torchrun --nproc-per-node N tools/check_fused_allreduce.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for p in (ROOT, os.path.join(ROOT, "nerf-ca_b200"), os.path.join(ROOT, "nerf-ca_b200", "train"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity  # noqa: E402
from nerfca import trainer as tr  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N_DEPTH, STEPS = 500, 6
N_RAYS = 256 if rank == 0 else 192                 # UNEVEN shards: the 1/B of every mean must be the global batch size
N_GLOBAL = 256 + 192 * (world - 1)
batches = []
for k in range(STEPS):
    rays, phases, z = parity.synthetic_batch(N_RAYS, N_DEPTH, seed=100 + 10 * k + rank)      # a different shard on every rank
    _, _, z = parity.synthetic_batch(1, N_DEPTH, seed=100 + 10 * k)                            # the same depth draw on every rank
    batches.append((rays.to(dev), phases.to(dev).int(), z.to(dev), N_GLOBAL))


def run(fused: bool):
    os.environ["NERFCA_FUSED_ALLREDUCE"] = "1" if fused else "0"
    torch.manual_seed(0)
    t = tr.CompositeTrainer.from_config(device=dev, precision="bf16", n_depth=N_DEPTH, world_size=world)
    assert (t.peer_grads is not None) == fused, "peer-mapped gradients could not be set up"
    t.set_iteration(50000)
    for b in batches:
        t.step_device(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for k in range(30):
        t.step_device(*batches[k % STEPS])
    e1.record()
    torch.cuda.synchronize()
    return t, e0.elapsed_time(e1) / 30


def params_after(fused):
    os.environ["NERFCA_FUSED_ALLREDUCE"] = "1" if fused else "0"
    torch.manual_seed(0)
    t = tr.CompositeTrainer.from_config(device=dev, precision="bf16", n_depth=N_DEPTH, world_size=world)
    t.set_iteration(50000)
    for b in batches:
        t.step_device(*b)
    torch.cuda.synchronize()
    return t.flat_p.clone(), t.schedule.t, float(t.flat_g.abs().max()), t.last_terms.clone(), (t.peer_grads is not None)


pf, step_f, gmax_f, terms_f, was_fused = params_after(True)
pn, step_n, gmax_n, terms_n, was_nccl_fused = params_after(False)
assert was_fused and not was_nccl_fused
gathered = [torch.zeros_like(pf) for _ in range(world)]
dist.all_gather(gathered, pf)
identical = all(torch.equal(gathered[0], g) for g in gathered[1:])
rel = parity.rel_l2(pf.cpu().numpy(), pn.cpu().numpy())
terms_rel = parity.rel_l2(terms_f.cpu().numpy(), terms_n.cpu().numpy())      # the loss sums that rode in the gradient exchange
_, ms_f = run(True)
_, ms_n = run(False)
if rank == 0:
    print(f"world {world}: replicas bit-identical {identical}; fused vs NCCL params rel-L2 {rel:.3e}; steps {step_f}/{step_n}; "
          f"grad buffers cleared {gmax_f == 0.0}/{gmax_n == 0.0}; ms/step fused {ms_f:.4f}  nccl {ms_n:.4f}", flush=True)
    print(f"loss sums fused vs NCCL rel-L2 {terms_rel:.3e}", flush=True)
    # (the two runs are separate trainings: the backward's red.global.add flush order differs run to run, so parameters and loss sums
    # agree to a few 1e-6 / 1e-8 after six steps rather than bit for bit; the replicas INSIDE one run are bit-identical)
    assert identical and rel <= 1e-4 and step_f == step_n == STEPS and gmax_f == 0.0 and terms_rel <= 1e-6
dist.destroy_process_group()

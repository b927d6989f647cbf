"""N > 1 host logic on CPU: two `gloo` ranks shard one global ray batch, each back-propagates the loss of its shard composed by the
product's `ops.loss_from_terms` with the GLOBAL 1/B, the flat gradient buffers are summed by the product's all-reduce helper, and the
result must equal the single-process gradient of the whole batch (rays are independent, one all-reduce per step); then
the same optimizer step on every rank must leave bit-identical replicas.  The arithmetic here is the CPU oracle (test infrastructure);
the CUDA counterpart of the additivity claim is tests/test_gpu_parity.py::test_full_size_step_properties."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402
from oracle import nerfca_oracle as orc  # noqa: E402
from nerfca import _lib as L, ops, trainer as tr  # noqa: E402

N_RAYS, N_DEPTH, IT = 10, 12, 50000       # 10 rays over 2 ranks: shards of 5; over 3 ranks: 4 + 3 + 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup():
    sd_s = orc.init_field_state(75, 32, 2, seed=1)
    sd_d = orc.init_field_state(83, 32, 2, 10, 8, seed=2)
    mask, _ = orc.freq_mask(12, IT, 150000, 1)
    cfg = {"n_freq": 12, "n_hidden": 2, "pos_enc": "free_windowed", "window": mask}
    rays, phases, z = parity.synthetic_batch(N_RAYS, N_DEPTH, seed=3)
    return sd_s, sd_d, cfg, rays, phases, z


def _shard_loss(sd_s, sd_d, cfg, rays, phases, z, n_rays_global):
    """Loss of one shard with the global batch size in every mean: built from the shard's SUMS exactly as the CUDA path does."""
    b = rays.shape[0]
    i0 = torch.full((b,), parity.I0, dtype=torch.float32)
    _, out = orc.composite_step_loss(sd_s, sd_d, cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z, rays[:, 2, 0], rays[:, 3, 0],
                                     orc.COMPOSITE_HP, IT)
    terms = [torch.zeros((), dtype=torch.float64) for _ in range(L.N_LOSS_TERMS)]
    terms[L.T_PIXEL_SUM] = out["pixel"].double() * b
    terms[L.T_FAVOR_SUM] = out["favor_s"].double() * b * N_DEPTH
    terms[L.T_D_ENT_SUM] = out["d_entropy"].double() * b
    terms[L.T_OCCL_SUM] = out["d_occl"].double() * b
    terms[L.T_L1_SUM] = out["s_l1"].double()
    terms[L.T_L2_SUM] = out["s_l2"].double()
    w = orc.schedule_weights(IT, orc.COMPOSITE_HP)
    lc = ops.LossConfig(w["favor_s"], w["dyn_entro"], w["occl"], w["l1"], 1e-4, 0.03, True, n_rays_global)
    return ops.loss_from_terms(torch.stack(terms), lc, n_rays_global, N_DEPTH)


class _Holder(torch.nn.Module):
    def __init__(self, sd):
        super().__init__()
        self.p = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(v.clone()) for k, v in sd.items()})
        self.keys = list(sd.keys())

    def state(self):
        return {k: self.p[k.replace(".", "_")] for k in self.keys}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        sd_s, sd_d, cfg, rays, phases, z = _setup()
        hs, hd = _Holder(sd_s), _Holder(sd_d)
        flat_p, flat_g, _ = tr.flatten_parameters([hs, hd], "cpu")       # the product's flat parameter / gradient buffers
        sl = tr.shard_slice(N_RAYS, rank, world)
        loss = _shard_loss(hs.state(), hd.state(), cfg, rays[sl], phases[sl], z, N_RAYS)
        params = list(hs.parameters()) + list(hd.parameters())
        grads = torch.autograd.grad(loss, params)
        for p, g in zip(params, grads):
            p.grad.copy_(g)                                              # .grad are views of flat_g
        tr.allreduce_sum_(flat_g)                                        # the step's only collective
        # identical replicas: the same Adam step from the same reduced gradient on every rank
        opt = torch.optim.Adam([flat_p.requires_grad_(False)], lr=1e-3)
        flat_p.grad = flat_g.clone()
        opt.step()
        gathered = [torch.zeros_like(flat_p) for _ in range(world)]
        dist.all_gather(gathered, flat_p)
        if rank == 0:
            ret["flat_g"] = flat_g.clone().numpy()
            ret["replicas_equal"] = all(torch.equal(gathered[0], g) for g in gathered[1:])
            ret["shard_sizes"] = [tr.shard_slice(N_RAYS, r, world).stop - tr.shard_slice(N_RAYS, r, world).start for r in range(world)]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gradients_sum_to_the_full_batch(world):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        got, equal, sizes = np.array(ret["flat_g"]), ret["replicas_equal"], list(ret["shard_sizes"])
    assert sum(sizes) == N_RAYS and max(sizes) - min(sizes) <= 1
    assert equal
    # single process, whole batch, the reference's own means (oracle composite_step_loss)
    sd_s, sd_d, cfg, rays, phases, z = _setup()
    hs, hd = _Holder(sd_s), _Holder(sd_d)
    i0 = torch.full((N_RAYS,), parity.I0, dtype=torch.float32)
    loss, _ = orc.composite_step_loss(hs.state(), hd.state(), cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z, rays[:, 2, 0],
                                      rays[:, 3, 0], orc.COMPOSITE_HP, IT)
    params = list(hs.parameters()) + list(hd.parameters())
    want = torch.autograd.grad(loss, params)
    off = 0
    for p, g in zip(params, want):
        n = p.numel()
        seg = got[off:off + n]
        assert parity.rel_l2(seg, g.numpy()) <= 1e-5, (off, n)
        off += (n + 3) // 4 * 4


def test_shard_slices_cover_the_batch():
    for n, w in [(1024, 8), (10, 3), (7, 8), (8192, 8)]:
        seen = []
        for r in range(w):
            s = tr.shard_slice(n, r, w)
            seen.extend(range(s.start, s.stop))
        assert seen == list(range(n))


def test_adam_schedule_scalars_equal_torch():
    """The host-side scalars handed to nerfca_adam_step (learning rate of LinearLR's recursive form, bias corrections) are the ones
    torch.optim.Adam + LinearLR(1 -> 0.01 over 150000) use, bit for bit, over the first 3000 updates and around the end of the decay."""
    p = torch.nn.Parameter(torch.zeros(4))
    opt = torch.optim.Adam([p], lr=1e-3)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.01, total_iters=40)
    mine = tr.AdamSchedule(1e-3, (0.9, 0.999), 1e-8, 0.01, 40)
    for t in range(1, 60):
        cfg = mine.next()
        assert cfg.lr == opt.param_groups[0]["lr"], t
        assert cfg.bias_correction1 == 1 - 0.9 ** t and cfg.bias_correction2_sqrt == (1 - 0.999 ** t) ** 0.5
        p.grad = torch.ones(4)
        opt.step()
        sched.step()
    assert mine.lr == opt.param_groups[0]["lr"] == pytest.approx(1e-5)
    opt = torch.optim.Adam([p], lr=1e-3)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.01, total_iters=150000)
    mine = tr.AdamSchedule(1e-3, (0.9, 0.999), 1e-8, 0.01, 150000)
    for t in range(1, 3001):
        assert mine.next().lr == opt.param_groups[0]["lr"], t
        opt.step()
        sched.step()


def test_global_batch_size_of_uneven_shards():
    """10 rays over 3 ranks are sharded 4 / 3 / 3: every rank must divide by the GLOBAL 10, not by local * world (12 / 9 / 9)."""
    class T:                      # the method only needs world_size
        world_size = 3
    sizes = [tr.shard_slice(10, r, 3).stop - tr.shard_slice(10, r, 3).start for r in range(3)]
    assert sizes == [4, 3, 3]
    for b in sizes:
        assert tr.CompositeTrainer._global_rays(T, b, 10) == 10
    assert tr.CompositeTrainer._global_rays(T, 4, None) == 12          # default: even shards

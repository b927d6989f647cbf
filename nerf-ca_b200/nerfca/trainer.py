"""CompositeTrainer: the composite training iteration of train/run_composite.py:227-308 as one host object.

It owns the two drop-in modules (model.CPPN.CPPN, model.Temporal.Temporal -- same state_dict keys as upstream), re-homes
their parameters / gradients into ONE flat fp32 buffer each (the nn.Parameters become views, so state_dict(), .save() and
any torch optimizer keep working), and runs a step as

    [H2D of the batch rows]  ->  fields forward (tcgen05)  ->  line integral + 11 loss terms + closed-form dL/d_raw
    ->  fields backward (tcgen05 dgrad / wgrad / latent scatter)  ->  [all-reduce of the flat gradient over NCCL]
    ->  fused Adam + LinearLR (also clears the gradient buffer)  ->  [D2H of the 16 loss sums]

Every arithmetic kernel is a C-ABI call into libnerfca_b200.so; torch supplies device memory, the stream and NCCL.
Rays are independent, so N GPUs each take B rays of the global batch and the only exchange is the gradient all-reduce
(SURVEY 8(e)); the 1/B of every mean uses the GLOBAL batch size.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops

# train/composite.txt:45-66
COMPOSITE_HP = {
    "entro_mask_thre": 1e-4, "entro_use_weighting": True, "entro_weighted_thresh": 0.03,
    "favor_s_weight_start": 1e-12, "favor_s_weight_end": 1e-10, "favor_s_weight_delay_steps": 40000,
    "dynamic_entro_weight_start": 1e-10, "dynamic_entro_weight_end": 1e-8,
    "occl_weight_start": 1e-8, "occl_weight_end": 1e-4,
    "l1_weight_start": 1e-8, "l1_weight_end": 1e-15,
    "hyperparam_decay_steps": 100000,
    "static_window_decay_steps": 150000, "temp_window_decay_steps": 150000,
}


def linear_param_decay(curr_iter, start_weight, end_weight, steps, delay_steps=0):
    """train/model_helpers.py:264-269."""
    if curr_iter < delay_steps:
        return 0
    alpha = min((curr_iter - delay_steps) / steps, 1.0)
    return (1.0 - alpha) * start_weight + alpha * end_weight


def flatten_parameters(modules, device, grad_alloc=None):
    """Move every parameter of `modules` into one flat fp32 buffer (16-byte aligned segments) and give each a .grad view
    into a second flat buffer (allocated by `grad_alloc(n)` if given, e.g. peer-mapped memory).  Returns (flat_params, flat_grads)."""
    params = [p for m in modules for p in m.parameters()]
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    flat_p = torch.zeros(total, dtype=torch.float32, device=device)
    flat_g = torch.zeros(total, dtype=torch.float32, device=device) if grad_alloc is None else grad_alloc(total)
    for p, o in zip(params, offs):
        seg = flat_p[o:o + p.numel()].view(p.shape)
        seg.copy_(p.data.to(device=device, dtype=torch.float32))
        p.data = seg
        p.grad = flat_g[o:o + p.numel()].view(p.shape)
    return flat_p, flat_g


def shard_slice(n_global: int, rank: int, world_size: int) -> slice:
    """Contiguous slice of a global batch of n_global rays owned by `rank` (every rank draws the same global ray ids
    from the same host RNG stream and keeps its slice; the remainder goes to the first ranks)."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """The step's only collective: in-place sum of one flat buffer (gradients, or the loss sums) over the ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class PeerGradients:
    """Peer mapping of the flat gradient buffer for nerfca_allreduce_adam_step: the buffer and a signal pad are allocated in
    torch's symmetric memory (used for the NVLink mapping and the pointer exchange only; the reduction + optimizer kernel is
    ours).  Raises if symmetric memory is unavailable -- the caller then stays on the NCCL all-reduce."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.symm, self.device = symm, device
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.handle = None
        self.epoch = 0

    def alloc(self, n: int) -> torch.Tensor:
        self.buf = self.symm.empty(n, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = self.symm.rendezvous(self.buf, self.group)
        h = self.handle
        assert h.world_size == self.world and h.rank == self.rank and h.signal_pad_size >= 1024
        self.peers = L.PeersStruct(self.rank, self.world, int(h.buffer_ptrs_dev), int(h.signal_pad_ptrs_dev), int(h.signal_pad_ptrs[self.rank]))
        torch.cuda.synchronize(self.device)
        h.barrier()                                    # every rank's buffer is zeroed and mapped before anybody's first step
        return self.buf

    def step(self, flat_p, flat_g, exp_avg, exp_avg_sq, step_dev, adam_cfg):
        assert flat_g.data_ptr() == self.buf.data_ptr()
        self.epoch += 1
        L.check(L.load().nerfca_allreduce_adam_step(C.byref(self.peers), self.epoch, L.ptr(flat_p), L.ptr(flat_g), L.ptr(exp_avg),
                                                    L.ptr(exp_avg_sq), flat_p.numel(), L.ptr(step_dev), C.byref(adam_cfg), L.stream_ptr()),
                "nerfca_allreduce_adam_step")


class PendingLoss:
    """Loss sums of one enqueued step: a pinned host slot + the event recorded behind its D2H copy (a ring of 8 per trainer, so a
    handle must be read before 8 further steps are enqueued)."""

    def __init__(self, trainer):
        self.trainer = trainer
        self.host = torch.zeros(L.N_LOSS_TERMS, dtype=torch.float64).pin_memory()
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()     # set by the batch gather for a ray id outside the table
        self.event = torch.cuda.Event()
        self.n_rays_global = 0
        self.cfg = None

    def loss(self) -> float:
        self.event.synchronize()
        if int(self.err_host[0]) != 0:
            self.err_host.zero_()
            self.trainer._gather_err.zero_()
            raise ValueError("ray id outside the ray table (nerfca_gather_batch)")
        return float(ops.loss_from_terms(self.host, self.cfg, self.n_rays_global, self.trainer.n_depth))


class CompositeTrainer:
    def __init__(self, static_model, temp_model, device, lr=1e-3, lr_end_factor=0.01, lr_decay_steps=150000, betas=(0.9, 0.999),
                 eps=1e-8, i0=float(np.log(8.670397)), near=3.2, far=8.8, n_depth=500, output_activation="softplus", hp=None,
                 world_size=1, process_group=None):
        self.static, self.temp, self.device = static_model, temp_model, torch.device(device)
        self.hp = dict(COMPOSITE_HP if hp is None else hp)
        self.output_activation = output_activation
        self.world_size, self.dist = int(world_size), process_group
        # N > 1: gradient sum fused with the optimizer step over NVLink peer memory (NERFCA_FUSED_ALLREDUCE=0: NCCL all-reduce + Adam)
        self.peer_grads = None
        if self.world_size > 1 and self.device.type == "cuda" and os.environ.get("NERFCA_FUSED_ALLREDUCE", "1") != "0":
            try:
                self.peer_grads = PeerGradients(self.device)
                self.flat_p, self.flat_g = flatten_parameters([static_model, temp_model], self.device, self.peer_grads.alloc)
            except Exception as e:                          # symmetric memory unavailable on this box: keep the NCCL path
                print(f"nerfca: peer-mapped gradients unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce", file=sys.stderr)
                self.peer_grads = None
        if self.peer_grads is None:
            self.flat_p, self.flat_g = flatten_parameters([static_model, temp_model], self.device)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.adam = L.AdamCfgStruct(float(lr), float(betas[0]), float(betas[1]), float(eps), float(lr_end_factor), int(lr_decay_steps))
        self.i0_value = float(i0)
        self.n_depth = int(n_depth)
        t = torch.linspace(0., 1., self.n_depth)
        self.depth_uniform = (near * (1. - t) + far * t).to(self.device)          # train/data_helpers.py:167-171
        self.terms = torch.zeros(L.N_LOSS_TERMS, dtype=torch.float64, device=self.device)
        self.last_terms = self.terms
        self._terms_host = torch.zeros(L.N_LOSS_TERMS, dtype=torch.float64).pin_memory() if self.device.type == "cuda" else None
        self._i0_cache = {}
        self._pending_slots = [PendingLoss(self) for _ in range(8)] if self.device.type == "cuda" else []
        self._pending_next = 0
        self.rays_table = self.phases_table = None       # device-resident ray table (attach_ray_table)
        self._gather_err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.iteration = 0
        self.loss_cfg = ops.LossConfig()
        self.set_iteration(0)

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, device, precision="bf16", n_freq=12, hidden=128, n_early=4, n_latent=8, n_phases=10, window_start=1,
                    **kw):
        """Build both fields with the default torch init (train/run_composite.py:147-207) and wrap them."""
        from model.CPPN import CPPN
        from model.Temporal import Temporal
        base = {"num_early_layers": n_early, "num_late_layers": 0, "num_filters": hidden, "num_input_channels": 3,
                "num_output_channels": 1, "use_bias": True, "pos_enc": "free_windowed", "pos_enc_window_start": window_start,
                "pos_enc_basis": n_freq, "fourier_sigma": 0.0, "fourier_gaussian": None, "act_func": "relu", "device": device,
                "precision": precision}
        temp_def = dict(base, num_input_times=1, use_time_latents=True, num_time_dim=n_latent)
        temp = Temporal(temp_def)
        if n_phases != temp.time_latents.shape[0]:     # config 3 needs 30 rows; upstream hard-codes 10 (Temporal.py:25-26)
            temp.time_latents = torch.nn.Parameter(torch.rand((n_phases, n_latent)))
        static = CPPN(dict(base))
        static.to(device)
        temp.to(device)
        return cls(static, temp, device, **kw)

    # ---- per-iteration host scalars (run_composite.py:238-247, 276-279) ---------------------------------------------
    def set_iteration(self, n_iter: int):
        hp = self.hp
        self.iteration = int(n_iter)
        self.static.update_freq_mask_alpha(n_iter, hp["static_window_decay_steps"])
        self.temp.update_freq_mask_alpha(n_iter, hp["temp_window_decay_steps"])
        d = hp["hyperparam_decay_steps"]
        self.loss_cfg = ops.LossConfig(
            favor_s_weight=linear_param_decay(n_iter, hp["favor_s_weight_start"], hp["favor_s_weight_end"], d, hp["favor_s_weight_delay_steps"]),
            dyn_entropy_weight=linear_param_decay(n_iter, hp["dynamic_entro_weight_start"], hp["dynamic_entro_weight_end"], d),
            occl_weight=linear_param_decay(n_iter, hp["occl_weight_start"], hp["occl_weight_end"], d, hp["favor_s_weight_delay_steps"]),
            l1_weight=linear_param_decay(n_iter, hp["l1_weight_start"], hp["l1_weight_end"], d),
            entro_mask_thre=hp["entro_mask_thre"], entro_weighted_thresh=hp["entro_weighted_thresh"],
            entro_use_weighting=hp["entro_use_weighting"], n_rays_global=0)

    def jitter(self, t_rand: torch.Tensor) -> torch.Tensor:
        """Stratified depth jitter (train/model_helpers.py:3-12) from a caller-supplied uniform draw."""
        return ops.jitter_depth(self.depth_uniform, t_rand.to(self.device, non_blocking=True))

    def _i0(self, n):
        if n not in self._i0_cache:
            self._i0_cache[n] = torch.full((n,), self.i0_value, dtype=torch.float32, device=self.device)
        return self._i0_cache[n]

    # ---- the step -------------------------------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(L.load().nerfca_launch_count())

    def step_device(self, rays: torch.Tensor, phases: torch.Tensor, depth: torch.Tensor):
        """One optimisation step on device-resident inputs: rays [B,4,3] f64, phases [B], depth [N] (already jittered).
        Loss sums of this rank are left in self.last_terms (device, float64)."""
        B = rays.shape[0]
        cfg = self.loss_cfg
        cfg.n_rays_global = B * self.world_size
        self.terms.zero_()
        ops.train_step_composite(self.static, self.temp, rays, phases, self._i0(B), depth, self.output_activation, cfg, self.terms)
        if self.peer_grads is not None:
            self.peer_grads.step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_dev, self.adam)
        else:
            if self.world_size > 1:
                allreduce_sum_(self.flat_g)
            L.check(L.load().nerfca_adam_step(L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                              self.flat_p.numel(), L.ptr(self.step_dev), C.byref(self.adam), 1.0, 1, L.stream_ptr()),
                    "nerfca_adam_step")
        self.last_terms = self.terms
        return self.terms

    d2h_bytes_per_step = L.N_LOSS_TERMS * 8

    def step_host(self, rays_host: torch.Tensor, phases_host: torch.Tensor, t_rand_host: torch.Tensor) -> float:
        """The call a user of the drop-in makes per iteration with HOST batch rows (run_composite.py:262-308):
        H2D of rays / phases / the jitter draw, the step, D2H of the loss sums; returns the total loss (python float)."""
        return self.step_host_async(rays_host, phases_host, t_rand_host).loss()

    def step_host_async(self, rays_host: torch.Tensor, phases_host: torch.Tensor, t_rand_host: torch.Tensor) -> "PendingLoss":
        """step_host without the host-side wait: everything (H2D copies, the step, the D2H copy of the loss sums into a pinned
        slot) is enqueued on the current stream and a handle is returned; `.loss()` waits for that step only.  A driver that
        reads the loss one iteration late (logging, early-stop checks) keeps the GPU busy back to back."""
        rays = rays_host.to(self.device, non_blocking=True)
        phases = phases_host.to(self.device, non_blocking=True)
        depth = self.jitter(t_rand_host)
        terms = self.step_device(rays, phases, depth)
        if self.world_size > 1:
            terms = allreduce_sum_(terms.clone())    # sums; the two maxima are per-rank diagnostics
        slot = self._pending_slots[self._pending_next % len(self._pending_slots)]
        self._pending_next += 1
        slot.host.copy_(terms, non_blocking=True)
        slot.event.record(torch.cuda.current_stream())
        slot.n_rays_global = rays.shape[0] * self.world_size
        slot.cfg = self.loss_cfg
        return slot

    # ---- N1: device-resident ray table, batches assembled on the device --------------------------------------------------
    def attach_ray_table(self, rays_train, phases_train):
        """Keep `rays_train [R,4,3] float64` / `phases_train [R] int64` (train/data_helpers.py:157-163; numpy or torch) in HBM.
        96 B + 8 B per ray: config 2 = 166 MB, config 3 = 6.5 GB of the 180 GB."""
        self.rays_table = torch.as_tensor(rays_train, dtype=torch.float64).to(self.device).contiguous()
        self.phases_table = torch.as_tensor(phases_train).to(device=self.device, dtype=torch.int64).contiguous()
        assert tuple(self.rays_table.shape[1:]) == (4, 3) and self.phases_table.numel() == self.rays_table.shape[0]

    def step_ids_async(self, ids_host: torch.Tensor, t_rand_host: torch.Tensor) -> "PendingLoss":
        """One iteration of run_composite.py:250-308 from the step's ray ids (int64 [B], host, drawn by the caller's RNG as
        upstream does): H2D of 8 B per ray instead of 96, the batch rows are gathered from the resident table by
        nerfca_gather_batch, then the step; returns the loss handle like step_host_async."""
        assert self.rays_table is not None, "attach_ray_table() first"
        ids = ids_host.to(self.device, non_blocking=True)
        rays, phases = ops.gather_batch(self.rays_table, self.phases_table, ids, self._gather_err)
        depth = self.jitter(t_rand_host)
        terms = self.step_device(rays, phases, depth)
        if self.world_size > 1:
            terms = allreduce_sum_(terms.clone())
        slot = self._pending_slots[self._pending_next % len(self._pending_slots)]
        self._pending_next += 1
        slot.host.copy_(terms, non_blocking=True)
        slot.err_host.copy_(self._gather_err, non_blocking=True)
        slot.event.record(torch.cuda.current_stream())
        slot.n_rays_global = rays.shape[0] * self.world_size
        slot.cfg = self.loss_cfg
        return slot

    def loss_from(self, terms: torch.Tensor, n_rays_global: Optional[int] = None) -> torch.Tensor:
        n = n_rays_global or (self.loss_cfg.n_rays_global or 1)
        return ops.loss_from_terms(terms, self.loss_cfg, n, self.n_depth)

    # ---- per-kernel device timing ------------------------------------------------------------------------------------
    def kernel_times(self, step_fn: Callable[[int], None], n_steps: int) -> Dict[str, dict]:
        """Runs n_steps steps with the library's event recording on and returns, per kernel family,
        {ms_per_step, ms_per_launch, launches_per_step}."""
        lib = L.load()
        torch.cuda.synchronize()
        lib.nerfca_profile_enable(1)
        for k in range(n_steps):
            step_fn(k)
        torch.cuda.synchronize()
        lib.nerfca_profile_enable(0)
        out = {}
        for kind, name in L.KERNEL_FAMILY_NAMES.items():
            ms, n = C.c_double(0), C.c_int64(0)
            L.check(lib.nerfca_profile_read(kind, C.byref(ms), C.byref(n)), "nerfca_profile_read")
            if n.value == 0:
                continue
            out[name] = {"ms_per_step": ms.value / n_steps, "ms_per_launch": ms.value / n.value,
                         "launches_per_step": n.value / n_steps}
        return out

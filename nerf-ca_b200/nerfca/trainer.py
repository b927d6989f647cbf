"""CompositeTrainer: the composite training iteration of train/run_composite.py:227-308 as one host object.

It owns the two drop-in modules (model.CPPN.CPPN, model.Temporal.Temporal -- same state_dict keys as upstream), re-homes
their parameters / gradients into ONE flat fp32 buffer each (the nn.Parameters become views, so state_dict(), .save() and
any torch optimizer keep working), and runs a step as ONE CUDA-graph launch of four kernels

    fields forward (tcgen05; also clears the loss sums)  ->  line integral + 11 loss terms + closed-form dL/d_raw
    ->  fields backward (tcgen05 dgrad / wgrad / latent gradient)
    ->  Adam + LinearLR + gradient clearing + bf16 re-pack of the updated weights   (N > 1: fused with the gradient AND
        loss-sum exchange over NVLink peer memory)

preceded by the H2D copies of the step's inputs (batch rows, or only the ray ids when the ray table is resident) and the
tiny depth-jitter / batch-gather kernels, followed by the D2H copy of the 16 loss sums.  Every arithmetic kernel is a
C-ABI call into libnerfca_b200.so; torch supplies device memory, streams and the process group.
Rays are independent, so N GPUs each take their shard of the global batch and the only exchange is the gradient sum
(SURVEY 8(e)); the 1/B of every mean uses the GLOBAL batch size.
"""
from __future__ import annotations

import ctypes as C
import math
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


def flatten_parameters(modules, device, alloc=None, tail: int = 0):
    """Move every parameter of `modules` into one flat fp32 buffer (16-byte aligned segments) and give each a .grad view
    into a second flat buffer (allocated by `alloc(n + tail)` if given, e.g. peer-mapped memory; `tail` extra floats follow
    the gradients).  Returns (flat_params [n], flat_grads [n], grad_buffer [n + tail])."""
    params = [p for m in modules for p in m.parameters()]
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    flat_p = torch.zeros(total, dtype=torch.float32, device=device)
    gbuf = torch.zeros(total + tail, dtype=torch.float32, device=device) if alloc is None else alloc(total + tail)
    flat_g = gbuf[:total]
    for p, o in zip(params, offs):
        seg = flat_p[o:o + p.numel()].view(p.shape)
        seg.copy_(p.data.to(device=device, dtype=torch.float32))
        p.data = seg
        p.grad = flat_g[o:o + p.numel()].view(p.shape)
    return flat_p, flat_g, gbuf


def shard_slice(n_global: int, rank: int, world_size: int) -> slice:
    """Contiguous slice of a global batch of n_global rays owned by `rank` (every rank draws the same global ray ids
    from the same host RNG stream and keeps its slice; the remainder goes to the first ranks)."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of one flat buffer over the ranks (NCCL; the fallback when peer-mapped memory is unavailable)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class AdamSchedule:
    """The host-side scalars of torch.optim.Adam + LinearLR(start_factor=1, end_factor, total_iters) for update t = 1, 2, ...,
    computed in python float arithmetic exactly as torch does (torch/optim/adam.py::_multi_tensor_adam,
    torch/optim/lr_scheduler.py::LinearLR.get_lr -- the RECURSIVE form, whose roundings differ from the closed form)."""

    def __init__(self, lr, betas, eps, end_factor, total_iters):
        self.base_lr, self.beta1, self.beta2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.start_factor, self.end_factor, self.total_iters = 1.0, float(end_factor), int(total_iters)
        self.t = 0                      # completed updates == LinearLR.last_epoch
        self.lr = self.base_lr * self.start_factor

    def next(self) -> L.AdamStepStruct:
        """Scalars of the next update; advances the step count and the scheduler (optimizer.step(); lr_scheduler.step())."""
        self.t += 1
        t = self.t
        bc1 = 1 - self.beta1 ** t
        bc2 = 1 - self.beta2 ** t
        cfg = L.AdamStepStruct(self.lr, self.beta1, self.beta2, self.eps, bc1, bc2 ** 0.5)
        # lr_scheduler.step(): last_epoch becomes t
        if self.total_iters > 0 and t <= self.total_iters:
            sf, ef = self.start_factor, self.end_factor
            self.lr = self.lr * (1.0 + (ef - sf) / (self.total_iters * sf + (t - 1) * (ef - sf)))
        return cfg


class PeerGradients:
    """Peer mapping of the flat gradient buffer (+ the 16 loss sums behind it) for nerfca_allreduce_adam_step: the buffer and a
    signal pad are allocated in torch's symmetric memory (used for the NVLink mapping and the pointer exchange only; the
    reduction + optimizer kernel is ours).  Raises if symmetric memory is unavailable -- the caller then stays on NCCL."""

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

    def step(self, flat_p, flat_g, exp_avg, exp_avg_sq, adam_cfg, repack, terms_offset, terms_out):
        assert flat_g.data_ptr() == self.buf.data_ptr()
        self.epoch += 1
        L.check(L.load().nerfca_allreduce_adam_step(C.byref(self.peers), self.epoch, L.ptr(flat_p), L.ptr(flat_g), L.ptr(exp_avg),
                                                    L.ptr(exp_avg_sq), flat_p.numel(), C.byref(adam_cfg),
                                                    C.byref(repack) if repack is not None else None, terms_offset, L.ptr(terms_out),
                                                    L.stream_ptr()),
                "nerfca_allreduce_adam_step")


class _InputSlot:
    """Device copies of one step's host inputs (batch rows or ray ids, phases, the jitter draw).  A ring of four per batch size: the
    H2D copies of step k run on the trainer's copy stream while the kernels of step k - 1 are still busy on the compute stream
    (`ready`: copies landed, the compute stream waits for it; `done`: the step that read the slot has been enqueued behind it, the
    copy stream waits for it before the slot is overwritten four steps later)."""

    def __init__(self, B, n_depth, device):
        self.rays = torch.empty((B, 4, 3), dtype=torch.float64, device=device)
        self.phases = torch.empty((B,), dtype=torch.int32, device=device)
        self.ids = torch.empty((B,), dtype=torch.int64, device=device)
        self.t_rand = torch.empty((n_depth,), dtype=torch.float32, device=device)
        self.ready, self.done = torch.cuda.Event(), torch.cuda.Event()
        self.used = False


class PendingLoss:
    """Loss sums of one enqueued step: a pinned host slot + the event recorded behind its D2H copy (a ring of 8 per trainer, so a
    handle must be read before 8 further steps are enqueued)."""

    def __init__(self, trainer):
        self.trainer = trainer
        self.host = torch.zeros(L.N_LOSS_TERMS, dtype=torch.float64).pin_memory()
        self.event = torch.cuda.Event()
        self.n_rays_global = 0
        self.cfg = None

    def loss(self) -> float:
        self.event.synchronize()
        return float(ops.loss_from_terms(self.host, self.cfg, self.n_rays_global, self.trainer.n_depth))


class CompositeTrainer:
    def __init__(self, static_model, temp_model, device, lr=1e-3, lr_end_factor=0.01, lr_decay_steps=150000, betas=(0.9, 0.999),
                 eps=1e-8, i0=float(np.log(8.670397)), near=3.2, far=8.8, n_depth=500, output_activation="softplus", hp=None,
                 world_size=1, process_group=None, use_graph=None):
        self.static, self.temp, self.device = static_model, temp_model, torch.device(device)
        self.hp = dict(COMPOSITE_HP if hp is None else hp)
        self.output_activation = output_activation
        self.world_size, self.group = int(world_size), process_group
        if self.world_size > 1:
            import torch.distributed as dist
            if dist.get_world_size(self.group) != self.world_size:
                raise ValueError(f"world_size {self.world_size} != size of the process group {dist.get_world_size(self.group)}")
        # N > 1: gradient sum fused with the optimizer step over NVLink peer memory (NERFCA_FUSED_ALLREDUCE=0: NCCL all-reduce + Adam)
        self.peer_grads = None
        tail = 2 * L.N_LOSS_TERMS                           # the loss sums (float64) ride behind the gradients
        if self.world_size > 1 and self.device.type == "cuda":
            import torch.distributed as dist
            ok = 0
            if os.environ.get("NERFCA_FUSED_ALLREDUCE", "1") != "0":
                try:
                    self.peer_grads = PeerGradients(self.device, self.group)
                    self.flat_p, self.flat_g, self.gbuf = flatten_parameters([static_model, temp_model], self.device, self.peer_grads.alloc, tail)
                    ok = 1
                except Exception as e:                          # symmetric memory unavailable on this box
                    print(f"nerfca: peer-mapped gradients unavailable ({type(e).__name__}: {e})", file=sys.stderr)
                    self.peer_grads = None
            # every rank must take the same route (a rank on NCCL would wait forever for peers spinning on flags)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0 and self.peer_grads is not None:
                self.peer_grads = None                          # (its buffer stays this rank's ordinary gradient buffer)
        if self.peer_grads is None:
            self.flat_p, self.flat_g, self.gbuf = flatten_parameters([static_model, temp_model], self.device, None, tail)
        n = self.flat_p.numel()
        self.terms_offset = n
        self.terms = self.gbuf[n:n + tail].view(torch.float64)              # this rank's loss sums of the step in flight
        self.terms_global = torch.zeros(L.N_LOSS_TERMS, dtype=torch.float64, device=self.device) if self.world_size > 1 else self.terms
        self.last_terms = self.terms_global
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.schedule = AdamSchedule(lr, betas, eps, lr_end_factor, lr_decay_steps)
        self.i0_value = float(i0)
        self.n_depth = int(n_depth)
        t = torch.linspace(0., 1., self.n_depth)
        self.depth_uniform = (near * (1. - t) + far * t).to(self.device)          # train/data_helpers.py:167-171
        self._depth_buf = torch.empty_like(self.depth_uniform)
        self._i0_cache = {}
        self._plans = {}
        self._param_version = 0                          # bumped by every parameter update; a plan's packed bf16 blocks carry the
                                                         # version they were written for
        self._pending_slots = [PendingLoss(self) for _ in range(8)] if self.device.type == "cuda" else []
        self._pending_next = 0
        self.rays_table = self.phases_table = None       # device-resident ray table (attach_ray_table)
        self._gather_err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.iteration = 0
        self.loss_cfg = ops.LossConfig()
        # one CUDA-graph launch per step (NERFCA_GRAPH=0: the same calls launched one by one); capture needs a non-default stream
        if use_graph is None:
            use_graph = os.environ.get("NERFCA_GRAPH", "1") != "0"
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._input_rings, self._input_next = {}, 0
        self._graph = C.c_void_p()
        if self.use_graph:
            L.check(L.load().nerfca_graph_create(C.byref(self._graph)), "nerfca_graph_create")
        self.set_iteration(0)

    def __del__(self):
        try:
            if getattr(self, "_graph", None) is not None and self._graph.value:
                L.load().nerfca_graph_destroy(self._graph)
        except Exception:
            pass

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
        for plan in getattr(self, "_plans", {}).values():
            plan.refresh_bands()

    def parameters_changed(self):
        """Call after writing the parameters from outside the trainer (load_state_dict, a torch optimizer): the bf16 operand copies
        are rebuilt by the next step."""
        self._param_version += 1

    def jitter(self, t_rand: torch.Tensor) -> torch.Tensor:
        """Stratified depth jitter (train/model_helpers.py:3-12) from a caller-supplied uniform draw."""
        return ops.jitter_depth(self.depth_uniform, t_rand.to(self.device, non_blocking=True))

    def _i0(self, n):
        if n not in self._i0_cache:
            self._i0_cache[n] = torch.full((n,), self.i0_value, dtype=torch.float32, device=self.device)
        return self._i0_cache[n]

    def _plan(self, n_rays: int) -> ops.StepPlan:
        plan = self._plans.get(n_rays)
        if plan is None:
            plan = ops.StepPlan(self.static, self.temp, n_rays, self.n_depth, self.device, self.output_activation, shared_scratch=False)
            plan.packed_version = -1                     # a new workspace: its packed blocks have never been written
            self._plans[n_rays] = plan
        return plan

    # ---- the step -------------------------------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(L.load().nerfca_launch_count())

    def graph_stats(self) -> dict:
        if not self.use_graph:
            return {"enabled": False}
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        L.check(L.load().nerfca_graph_stats(self._graph, C.byref(a), C.byref(b), C.byref(c)), "nerfca_graph_stats")
        return {"enabled": True, "launches": a.value, "updates": b.value, "instantiations": c.value}

    def _enqueue_step(self, rays, phases, depth, n_rays_global, pre=None):
        """The step's kernels on the current stream (== self.stream when graphs are on): optional `pre` launches (batch gather,
        depth jitter), nerfca_train_step, the optimizer.  Allocation-free, so it can sit inside a capture bracket."""
        B = rays.shape[0]
        plan = self._plan(B)
        cfg = self.loss_cfg
        cfg.n_rays_global = int(n_rays_global)
        adam_cfg = self.schedule.next()
        lib = L.load()
        sp = L.stream_ptr()
        capturing = False
        if self.use_graph and not self._profiling and not (self.world_size > 1 and self.peer_grads is None):
            L.check(lib.nerfca_graph_begin(self._graph, sp), "nerfca_graph_begin")
            capturing = True
        try:
            if pre is not None:
                pre()
            bf16 = plan.prec == L.PREC_BF16
            flags = L.STEP_ZERO_TERMS | (L.STEP_PACKED if (bf16 and plan.packed_version == self._param_version) else 0)
            if B > 0:
                plan.run(rays, phases, self._i0(B), depth, cfg, self.terms, flags)
            else:
                self.terms.zero_()                       # an empty shard still joins the exchange below
            repack = plan.repack_struct() if bf16 and B > 0 else None
            if self.peer_grads is not None:
                self.peer_grads.step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, adam_cfg, repack, self.terms_offset,
                                     self.terms_global)
            else:
                if self.world_size > 1:
                    allreduce_sum_(self.flat_g, self.group)
                    self.terms_global.copy_(self.terms)
                    mx = self.terms_global[2:4].clone()
                    allreduce_sum_(self.terms_global, self.group)
                    import torch.distributed as dist
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
                    self.terms_global[2:4] = mx
                L.check(lib.nerfca_adam_step(L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                             self.flat_p.numel(), C.byref(adam_cfg), 1.0, 1, C.byref(repack) if repack is not None else None,
                                             sp), "nerfca_adam_step")
            self._param_version += 1
            if repack is not None:
                plan.packed_version = self._param_version
        except Exception:
            if capturing:
                lib.nerfca_graph_abort(self._graph, sp)
            raise
        if capturing:
            L.check(lib.nerfca_graph_end_launch(self._graph, sp), "nerfca_graph_end_launch")
        self.last_terms = self.terms_global
        return self.terms_global

    _profiling = False

    def _on_stream(self):
        """Context: run on the trainer's stream, ordered after the caller's current stream (and the caller's stream after it)."""
        return _StreamScope(self.stream)

    def _global_rays(self, n_local: int, n_rays_global: Optional[int]) -> int:
        return int(n_rays_global) if n_rays_global else n_local * self.world_size

    def step_device(self, rays: torch.Tensor, phases: torch.Tensor, depth: torch.Tensor, n_rays_global: Optional[int] = None):
        """One optimisation step on device-resident inputs: rays [B,4,3] f64, phases [B] int32, depth [N] f32 (already jittered).
        n_rays_global: rays of the WHOLE job's batch (default B * world_size; pass it when the shards are uneven).
        The loss sums (of the global batch when N > 1) are left in self.last_terms (device, float64)."""
        if phases.dtype != torch.int32:
            phases = phases.to(torch.int32)
        with self._on_stream():
            return self._enqueue_step(rays, phases, depth, self._global_rays(rays.shape[0], n_rays_global))

    d2h_bytes_per_step = L.N_LOSS_TERMS * 8

    def _finish_async(self, n_rays_global) -> "PendingLoss":
        slot = self._pending_slots[self._pending_next % len(self._pending_slots)]
        self._pending_next += 1
        slot.host.copy_(self.terms_global, non_blocking=True)
        slot.event.record(torch.cuda.current_stream())
        slot.n_rays_global = int(n_rays_global)
        slot.cfg = self.loss_cfg
        return slot

    def step_host(self, rays_host: torch.Tensor, phases_host: torch.Tensor, t_rand_host: torch.Tensor, n_rays_global=None) -> float:
        """The call a user of the drop-in makes per iteration with HOST batch rows (run_composite.py:262-308):
        H2D of rays / phases / the jitter draw, the step, D2H of the loss sums; returns the total loss (python float)."""
        return self.step_host_async(rays_host, phases_host, t_rand_host, n_rays_global).loss()

    def step_host_async(self, rays_host: torch.Tensor, phases_host: torch.Tensor, t_rand_host: torch.Tensor,
                        n_rays_global=None) -> "PendingLoss":
        """step_host without the host-side wait: everything (H2D copies, the step, the D2H copy of the loss sums into a pinned
        slot) is enqueued and a handle is returned; `.loss()` waits for that step only.  A driver that reads the loss one
        iteration late (logging, early-stop checks) keeps the GPU busy back to back."""
        if phases_host.dtype != torch.int32:
            phases_host = phases_host.to(torch.int32)          # run_composite.py:265 `.int()`, on the host: 4 B per ray cross PCIe
        n_glob = self._global_rays(rays_host.shape[0], n_rays_global)
        slot = self._stage(rays_host.shape[0], rays=rays_host, phases=phases_host, t_rand=t_rand_host)
        with self._on_stream():
            torch.cuda.current_stream().wait_event(slot.ready)
            self._enqueue_step(slot.rays, slot.phases, self._depth_buf, n_glob, pre=lambda: self._jitter_into(slot.t_rand))
            slot.done.record(torch.cuda.current_stream())
            return self._finish_async(n_glob)

    def _stage(self, B: int, **host) -> _InputSlot:
        """H2D copies of one step's host inputs into the next input slot, on the copy stream (they overlap the previous step's kernels)."""
        ring = self._input_rings.get(B)
        if ring is None:
            ring = self._input_rings[B] = [_InputSlot(B, self.n_depth, self.device) for _ in range(4)]
        slot = ring[self._input_next % len(ring)]
        self._input_next += 1
        cs = self._copy_stream
        if slot.used:
            cs.wait_event(slot.done)
        slot.used = True
        with torch.cuda.stream(cs):
            for name, src in host.items():
                getattr(slot, name).copy_(src, non_blocking=True)
            slot.ready.record(cs)
        return slot

    def _jitter_into(self, t_rand_dev):
        L.check(L.load().nerfca_jitter_depth(L.ptr(self.depth_uniform), L.ptr(t_rand_dev), self.n_depth, L.ptr(self._depth_buf),
                                             L.stream_ptr()), "nerfca_jitter_depth")

    # ---- N1: device-resident ray table, batches assembled on the device --------------------------------------------------
    def attach_ray_table(self, rays_train, phases_train):
        """Keep `rays_train [R,4,3] float64` / `phases_train [R] int64` (train/data_helpers.py:157-163; numpy or torch) in HBM.
        96 B + 8 B per ray: config 2 = 166 MB, config 3 = 6.5 GB of the 180 GB."""
        self.rays_table = torch.as_tensor(rays_train, dtype=torch.float64).to(self.device).contiguous()
        self.phases_table = torch.as_tensor(phases_train).to(device=self.device, dtype=torch.int64).contiguous()
        assert tuple(self.rays_table.shape[1:]) == (4, 3) and self.phases_table.numel() == self.rays_table.shape[0]

    def step_ids_async(self, ids_host: torch.Tensor, t_rand_host: torch.Tensor, n_rays_global=None) -> "PendingLoss":
        """One iteration of run_composite.py:250-308 from the step's ray ids (int64 [B], host, drawn by the caller's RNG as
        upstream does): H2D of 8 B per ray instead of 96, the batch rows are gathered from the resident table by
        nerfca_gather_batch inside the step's graph; returns the loss handle like step_host_async.  An id outside the table
        raises IndexError before anything is enqueued (upstream: numpy fancy indexing raises before the step)."""
        assert self.rays_table is not None, "attach_ray_table() first"
        B = int(ids_host.numel())
        R = self.rays_table.shape[0]
        if B > 0 and (int(ids_host.min()) < 0 or int(ids_host.max()) >= R):
            raise IndexError(f"ray id outside the ray table of {R} rays")
        n_glob = self._global_rays(B, n_rays_global)
        slot = self._stage(B, ids=ids_host, t_rand=t_rand_host)
        rays, phases, ids = slot.rays, slot.phases, slot.ids          # the gathered rows land in the slot's own row buffers
        lib = L.load()

        def pre():
            if B > 0:
                L.check(lib.nerfca_gather_batch(L.ptr(self.rays_table), L.ptr(self.phases_table), R, L.ptr(ids), B, L.ptr(rays), L.ptr(phases),
                                                L.ptr(self._gather_err), L.stream_ptr()), "nerfca_gather_batch")
            self._jitter_into(slot.t_rand)
        with self._on_stream():
            torch.cuda.current_stream().wait_event(slot.ready)
            self._enqueue_step(rays, phases, self._depth_buf, n_glob, pre=pre)
            slot.done.record(torch.cuda.current_stream())
            return self._finish_async(n_glob)

    def loss_from(self, terms: torch.Tensor, n_rays_global: Optional[int] = None) -> torch.Tensor:
        n = n_rays_global or (self.loss_cfg.n_rays_global or 1)
        return ops.loss_from_terms(terms, self.loss_cfg, n, self.n_depth)

    # ---- per-kernel device timing ------------------------------------------------------------------------------------
    def kernel_times(self, step_fn: Callable[[int], None], n_steps: int) -> Dict[str, dict]:
        """Runs n_steps steps with the library's event recording on (kernels launched one by one instead of as a graph) and
        returns, per kernel family, {ms_per_step, ms_per_launch, launches_per_step}."""
        lib = L.load()
        torch.cuda.synchronize()
        lib.nerfca_profile_enable(1)
        self._profiling = True
        try:
            for k in range(n_steps):
                step_fn(k)
            torch.cuda.synchronize()
        finally:
            self._profiling = False
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


class _StreamScope:
    """`with` block on `stream`: waits for the caller's current stream on entry, makes the caller's stream wait on exit.  A no-op
    when the caller already runs on `stream` (the benchmark loop does) or on CPU."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        if self.stream is None:
            return self
        self.outer = torch.cuda.current_stream(self.stream.device)
        self.same = self.outer == self.stream
        if not self.same:
            self.stream.wait_stream(self.outer)
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.stream is None or self.same:
            return False
        self.ctx.__exit__(*exc)
        self.outer.wait_stream(self.stream)
        return False

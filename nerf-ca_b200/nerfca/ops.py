"""Host-side operators over the C ABI: sample-set descriptors, autograd bridges for the two fields
and the line integral, and the fused training step (forward + loss + backward in five kernels'
worth of C calls, no autograd graph).

Reference surface mirrored here: model/CPPN.py:88-135, model/Temporal.py:113-177,
train/model_helpers.py:3-160, 189-262, train/run_composite.py:262-305.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Sequence

import torch

from . import _lib as L

ACTIVATIONS = {"softplus": L.ACT_SOFTPLUS, "clamp": L.ACT_CLAMP}  # anything else -> Sigmoid (model_helpers.py:63-70)
PRECISIONS = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}


def default_precision() -> str:
    return os.environ.get("NERFCA_PRECISION", "bf16")


def activation_code(name: str) -> int:
    return ACTIVATIONS.get(name, L.ACT_SIGMOID)


def _byte_buffer(nbytes: int, device) -> Optional[torch.Tensor]:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device) if nbytes > 0 else None


# ------------------------------------------------------------------------------------------------
# sample sets
# ------------------------------------------------------------------------------------------------


@dataclass
class Samples:
    """Where the P sample positions come from (nerfca_samples_t)."""
    n_points: int
    points: Optional[torch.Tensor] = None        # [P,3] f32
    origins: Optional[torch.Tensor] = None       # [B,3] f32/f64, rows may be strided (views of rays_train[B,4,3])
    dirs: Optional[torch.Tensor] = None
    depth: Optional[torch.Tensor] = None         # [N] f32
    phase_point: Optional[torch.Tensor] = None   # [P] int32
    phase_ray: Optional[torch.Tensor] = None     # [B] int32
    _struct: Optional[L.SamplesStruct] = dc_field(default=None, repr=False)

    @staticmethod
    def from_points(x: torch.Tensor, phases: Optional[torch.Tensor] = None) -> "Samples":
        if x.dim() != 2 or x.shape[-1] != 3:
            raise ValueError(f"expected points of shape [P,3], got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        ph = None
        if phases is not None:
            ph = phases.detach().flatten().to(device=x.device).to(torch.int64).to(torch.int32).contiguous()  # ts.flatten().long(), Temporal.py:144
            if ph.numel() != x.shape[0]:
                raise ValueError("one phase per sample point expected")
        return Samples(n_points=x.shape[0], points=x, phase_point=ph)

    @staticmethod
    def from_rays(origins: torch.Tensor, dirs: torch.Tensor, depth: torch.Tensor,
                  phase_ray: Optional[torch.Tensor] = None) -> "Samples":
        if origins.shape != dirs.shape or origins.dim() != 2 or origins.shape[-1] != 3:
            raise ValueError("origins / directions must both be [B,3]")
        if origins.dtype != dirs.dtype or origins.dtype not in (torch.float32, torch.float64):
            origins, dirs = origins.to(torch.float32), dirs.to(torch.float32)

        def rows(t):
            t = t.detach()
            if t.stride(1) != 1 or t.stride(0) < 3:
                t = t.contiguous()
            return t
        o, d = rows(origins), rows(dirs)
        if o.stride(0) != d.stride(0):
            o, d = o.contiguous(), d.contiguous()
        z = depth.detach().to(device=o.device, dtype=torch.float32).contiguous()
        ph = None
        if phase_ray is not None:
            ph = phase_ray.detach().flatten().to(device=o.device).to(torch.int64).to(torch.int32).contiguous()
            if ph.numel() != o.shape[0]:
                raise ValueError("one phase per ray expected")
        return Samples(n_points=o.shape[0] * z.shape[0], origins=o, dirs=d, depth=z, phase_ray=ph)

    @property
    def device(self):
        return self.points.device if self.points is not None else self.origins.device

    def struct(self) -> L.SamplesStruct:
        if self._struct is None:
            s = L.SamplesStruct()
            s.n_points = self.n_points
            s.points = L.ptr(self.points)
            if self.points is None:
                s.n_rays, s.n_depth = self.origins.shape[0], self.depth.shape[0]
                s.origins, s.dirs, s.depth = L.ptr(self.origins), L.ptr(self.dirs), L.ptr(self.depth)
                s.ray_dtype = L.F64 if self.origins.dtype == torch.float64 else L.F32
                s.ray_stride = self.origins.stride(0) if self.origins.shape[0] > 1 else 3
            s.phase_point = L.ptr(self.phase_point)
            s.phase_ray = L.ptr(self.phase_ray)
            self._struct = s
        return self._struct


def sample_points(samples: Samples) -> torch.Tensor:
    """A4: materialise [P,3] float32 sample positions (bit-identical to the reference's expression)."""
    out = torch.empty((samples.n_points, 3), dtype=torch.float32, device=samples.device)
    L.check(L.load().nerfca_sample_points(C.byref(samples.struct()), L.ptr(out), L.stream_ptr()), "nerfca_sample_points")
    return out


def jitter_depth(z: torch.Tensor, t_rand: torch.Tensor) -> torch.Tensor:
    """A3 on the device; the uniform draw is supplied by the caller (CPU generator in the reference)."""
    z = z.to(torch.float32).contiguous()
    t = t_rand.to(device=z.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(z)
    L.check(L.load().nerfca_jitter_depth(L.ptr(z), L.ptr(t), z.numel(), L.ptr(out), L.stream_ptr()), "nerfca_jitter_depth")
    return out


def gather_batch(rays_table: torch.Tensor, phases_table: Optional[torch.Tensor], ids: torch.Tensor,
                 err_flag: Optional[torch.Tensor] = None):
    """N1 (train/run_composite.py:250-273): rays_table[ids] / phases_table[ids].int() assembled on the device from the
    device-resident ray table [R,4,3] float64 (+ phases [R] int64) and int64 ray ids [B].  Bit copy of the rows.
    `err_flag` (device int32 [1]) is set to 1 by an id outside [0, R); the caller decides when to read it."""
    assert rays_table.dtype == torch.float64 and rays_table.is_contiguous() and tuple(rays_table.shape[1:]) == (4, 3)
    ids = ids.to(device=rays_table.device, dtype=torch.int64).contiguous()
    B = ids.numel()
    rays = torch.empty((B, 4, 3), dtype=torch.float64, device=rays_table.device)
    phases = torch.empty((B,), dtype=torch.int32, device=rays_table.device)
    if phases_table is not None:
        assert phases_table.dtype == torch.int64 and phases_table.is_contiguous() and phases_table.numel() == rays_table.shape[0]
    L.check(L.load().nerfca_gather_batch(L.ptr(rays_table), L.ptr(phases_table) if phases_table is not None else None,
                                         rays_table.shape[0], L.ptr(ids), B, L.ptr(rays), L.ptr(phases),
                                         L.ptr(err_flag) if err_flag is not None else None, L.stream_ptr()), "nerfca_gather_batch")
    return rays, phases


# ------------------------------------------------------------------------------------------------
# fields
# ------------------------------------------------------------------------------------------------


@dataclass
class FieldSpec:
    """Static description of one field + the tensors its descriptor points at."""
    enc_mode: int
    n_freq: int
    n_latent: int
    n_phases: int
    hidden: int
    n_hidden: int
    use_bias: bool
    band_weight: Optional[torch.Tensor] = None
    fourier_coeff: Optional[torch.Tensor] = None

    @property
    def enc_dim(self) -> int:
        if self.enc_mode == L.ENC_NONE or self.n_freq <= 0:
            return 3
        return 6 * self.n_freq if self.enc_mode == L.ENC_FOURIER else 3 + 6 * self.n_freq

    @property
    def in_dim(self) -> int:
        return self.enc_dim + self.n_latent

    def n_params(self) -> int:
        return (1 if self.n_latent else 0) + (self.n_hidden + 2) * (2 if self.use_bias else 1)

    def split(self, params: Sequence[torch.Tensor]):
        """params in canonical order -> (latents, weights[], biases[])."""
        params = list(params)
        lat = params.pop(0) if self.n_latent else None
        if self.use_bias:
            return lat, params[0::2], params[1::2]
        return lat, params, [None] * len(params)

    def struct(self, params: Sequence[torch.Tensor]) -> L.FieldStruct:
        lat, ws, bs = self.split(params)
        f = L.FieldStruct()
        f.enc_mode, f.n_freq, f.n_latent, f.n_phases = self.enc_mode, self.n_freq, self.n_latent, self.n_phases
        f.hidden, f.n_hidden = self.hidden, self.n_hidden
        f.band_weight = L.ptr(self.band_weight)
        f.fourier_coeff = L.ptr(self.fourier_coeff)
        f.latents = L.ptr(lat)
        for k, (w, b) in enumerate(zip(ws, bs)):
            f.weight[k] = L.ptr(w)
            f.bias[k] = L.ptr(b)
        return f


def _check_params(spec: FieldSpec, params: Sequence[torch.Tensor]):
    if len(params) != spec.n_params():
        raise ValueError(f"expected {spec.n_params()} parameter tensors, got {len(params)}")
    out = []
    for p in params:
        p = p.detach()
        if p.dtype != torch.float32 or not p.is_contiguous():
            p = p.to(torch.float32).contiguous()
        out.append(p)
    return out


def field_forward_raw(spec: FieldSpec, samples: Samples, precision: int, params: Sequence[torch.Tensor], keep_stash: bool):
    """-> (raw[P] float32, stash or None)."""
    lib = L.load()
    dev = samples.device
    ps = _check_params(spec, params)
    fs = spec.struct(ps)
    P = samples.n_points
    raw = torch.empty((P,), dtype=torch.float32, device=dev)
    if P == 0:
        return raw, None
    stash = _byte_buffer(lib.nerfca_field_stash_bytes(C.byref(fs), P, precision), dev) if keep_stash else None
    ws = _byte_buffer(lib.nerfca_field_workspace_bytes(C.byref(fs), P, precision, 0), dev)
    L.check(lib.nerfca_field_forward(C.byref(fs), C.byref(samples.struct()), precision, L.ptr(raw), L.ptr(stash), L.ptr(ws),
                                     L.stream_ptr()), "nerfca_field_forward")
    return raw, stash


def field_backward_raw(spec: FieldSpec, samples: Samples, precision: int, params: Sequence[torch.Tensor], d_raw: torch.Tensor,
                       stash: torch.Tensor, grads: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """Accumulates into `grads` (allocated zero-filled when None); returns them in canonical parameter order."""
    lib = L.load()
    ps = _check_params(spec, params)
    fs = spec.struct(ps)
    P = samples.n_points
    if grads is None:
        grads = [torch.zeros_like(p) for p in ps]
    if P == 0:
        return grads
    glat, gws, gbs = spec.split(grads)
    g = L.FieldGradsStruct()
    g.latents = L.ptr(glat)
    for k, (w, b) in enumerate(zip(gws, gbs)):
        g.weight[k] = L.ptr(w)
        g.bias[k] = L.ptr(b)
    d_raw = d_raw.detach().reshape(-1).to(torch.float32).contiguous()
    ws = _byte_buffer(lib.nerfca_field_workspace_bytes(C.byref(fs), P, precision, 1), samples.device)
    L.check(lib.nerfca_field_backward(C.byref(fs), C.byref(samples.struct()), precision, L.ptr(d_raw), L.ptr(stash), L.ptr(ws),
                                      C.byref(g), L.stream_ptr()), "nerfca_field_backward")
    return grads


class FieldFunction(torch.autograd.Function):
    """raw[P,1] = field(samples); differentiable w.r.t. the field's parameters only (as in the reference, where
    the sample positions never require grad)."""

    @staticmethod
    def forward(ctx, spec: FieldSpec, samples: Samples, precision: int, *params):
        need = any(ctx.needs_input_grad[3:])
        raw, stash = field_forward_raw(spec, samples, precision, params, keep_stash=need)
        ctx.spec, ctx.samples, ctx.precision, ctx.stash = spec, samples, precision, stash
        ctx.save_for_backward(*params)
        return raw.unsqueeze(-1)

    @staticmethod
    def backward(ctx, d_raw):
        grads = field_backward_raw(ctx.spec, ctx.samples, ctx.precision, ctx.saved_tensors, d_raw, ctx.stash)
        ctx.stash = None
        return (None, None, None, *grads)


def encode(spec: FieldSpec, samples: Samples, params: Sequence[torch.Tensor]) -> torch.Tensor:
    """A5 debug / parity entry: the first-layer input [P, in_dim] float32."""
    fs = spec.struct(_check_params(spec, params))
    out = torch.empty((samples.n_points, spec.in_dim), dtype=torch.float32, device=samples.device)
    L.check(L.load().nerfca_encode(C.byref(fs), C.byref(samples.struct()), L.ptr(out), L.stream_ptr()), "nerfca_encode")
    return out


# ------------------------------------------------------------------------------------------------
# line integral
# ------------------------------------------------------------------------------------------------


class IntegrateFunction(torch.autograd.Function):
    """train/model_helpers.py:72-97.  raw_d None -> single-field form.  acc64: float64 ray sums (training)."""

    @staticmethod
    def forward(ctx, raw_s, raw_d, i0, depth, act: int, acc64: bool):
        lib = L.load()
        B, N = raw_s.shape[0], raw_s.shape[1]
        dev = raw_s.device
        rs = raw_s.detach().reshape(B, N).to(torch.float32).contiguous()
        rd = raw_d.detach().reshape(B, N).to(torch.float32).contiguous() if raw_d is not None else None
        z = depth.detach().to(device=dev, dtype=torch.float32).contiguous()
        i0f = i0.detach().to(device=dev, dtype=torch.float32).contiguous()
        acc = torch.float64 if acc64 else torch.float32
        pix = torch.empty((B,), dtype=acc, device=dev)
        dists = torch.empty((N,), dtype=acc, device=dev)
        ss = torch.empty((B, N), dtype=torch.float32, device=dev)
        sd = torch.empty((B, N), dtype=torch.float32, device=dev) if rd is not None else None
        L.check(lib.nerfca_integrate(L.ptr(rs), L.ptr(rd), L.ptr(z), L.ptr(i0f), B, N, act, L.F64 if acc64 else L.F32,
                                     L.ptr(pix), L.ptr(ss), L.ptr(sd), L.ptr(dists), L.stream_ptr()), "nerfca_integrate")
        ctx.save_for_backward(rs, rd, z)
        ctx.act, ctx.acc64, ctx.shape_s = act, acc64, raw_s.shape
        ctx.mark_non_differentiable(dists)
        if rd is None:
            return pix, ss, dists
        return pix, ss, sd, dists

    @staticmethod
    def backward(ctx, *gouts):
        rs, rd, z = ctx.saved_tensors
        B, N = rs.shape
        acc = torch.float64 if ctx.acc64 else torch.float32
        d_pix = gouts[0]
        d_ss = gouts[1]
        d_sd = gouts[2] if rd is not None else None
        d_pix = d_pix.to(acc).contiguous() if d_pix is not None else None
        d_ss = d_ss.to(torch.float32).contiguous() if d_ss is not None else None
        d_sd = d_sd.to(torch.float32).contiguous() if d_sd is not None else None
        g_s = torch.empty_like(rs)
        g_d = torch.empty_like(rd) if rd is not None else None
        L.check(L.load().nerfca_integrate_backward(L.ptr(rs), L.ptr(rd), L.ptr(z), B, N, ctx.act, L.F64 if ctx.acc64 else L.F32,
                                                   L.ptr(d_pix), L.ptr(d_ss), L.ptr(d_sd), L.ptr(g_s), L.ptr(g_d), L.stream_ptr()),
                "nerfca_integrate_backward")
        g_s = g_s.reshape(ctx.shape_s)
        g_d = g_d.reshape(ctx.shape_s) if g_d is not None else None
        return g_s, g_d, None, None, None, None


# ------------------------------------------------------------------------------------------------
# fused training step (no autograd graph): fields forward -> integral + losses + dL/d_raw -> fields backward
# ------------------------------------------------------------------------------------------------


@dataclass
class LossConfig:
    """Scheduled weights + thresholds of one step (train/run_composite.py:276-292, composite.txt:45-66)."""
    favor_s_weight: float = 0.0
    dyn_entropy_weight: float = 0.0
    occl_weight: float = 0.0
    l1_weight: float = 0.0
    entro_mask_thre: float = 1e-4
    entro_weighted_thresh: float = 0.03
    entro_use_weighting: bool = True
    n_rays_global: int = 0

    def struct(self, n_rays: int) -> L.LossCfgStruct:
        c = L.LossCfgStruct()
        c.favor_s_weight, c.dyn_entropy_weight = float(self.favor_s_weight), float(self.dyn_entropy_weight)
        c.occl_weight, c.l1_weight = float(self.occl_weight), float(self.l1_weight)
        c.entro_mask_thre, c.entro_weighted_thresh = float(self.entro_mask_thre), float(self.entro_weighted_thresh)
        c.entro_use_weighting = int(bool(self.entro_use_weighting))
        c.n_rays_global = int(self.n_rays_global or n_rays)
        return c


def composite_loss(raw_s: torch.Tensor, raw_d: Optional[torch.Tensor], depth: torch.Tensor, i0: torch.Tensor,
                   gt: torch.Tensor, wpix: torch.Tensor, act: int, cfg: LossConfig, want_grad: bool = True,
                   terms: Optional[torch.Tensor] = None):
    """nerfca_composite_loss: -> (pix[B] f64, terms[16] f64 raw sums, d_raw_s, d_raw_d)."""
    B, N = gt.shape[0], depth.shape[0]
    dev = raw_s.device
    if gt.dtype != torch.float64 or wpix.dtype != torch.float64 or gt.stride(0) != wpix.stride(0):
        gt, wpix = gt.to(torch.float64).contiguous(), wpix.to(torch.float64).contiguous()
    stride = gt.stride(0) if B > 1 else 1
    pix = torch.empty((B,), dtype=torch.float64, device=dev)
    if terms is None:
        terms = torch.zeros((L.N_LOSS_TERMS,), dtype=torch.float64, device=dev)
    g_s = torch.empty_like(raw_s) if want_grad else None
    g_d = torch.empty_like(raw_d) if (want_grad and raw_d is not None) else None
    c = cfg.struct(B)
    L.check(L.load().nerfca_composite_loss(L.ptr(raw_s), L.ptr(raw_d), L.ptr(depth), L.ptr(i0), L.ptr(gt), L.ptr(wpix), stride, B, N,
                                           act, C.byref(c), L.ptr(pix), L.ptr(terms), L.ptr(g_s), L.ptr(g_d), L.stream_ptr()),
            "nerfca_composite_loss")
    return pix, terms, g_s, g_d


def loss_from_terms(terms: torch.Tensor, cfg: LossConfig, n_rays_global: int, n_depth: int, static_only: bool = False):
    """Total loss of run_composite.py:292 (or run_nerf.py:230) from the device-side sums; stays on the device."""
    B = float(n_rays_global)
    pixel = terms[L.T_PIXEL_SUM] / B
    if static_only:
        return pixel + cfg.occl_weight * terms[L.T_OCCL_SUM] / B
    return (pixel + cfg.favor_s_weight * terms[L.T_FAVOR_SUM] / (B * n_depth) + cfg.dyn_entropy_weight * terms[L.T_D_ENT_SUM] / B
            + cfg.occl_weight * terms[L.T_OCCL_SUM] / B + cfg.l1_weight * terms[L.T_L2_SUM] + cfg.l1_weight * terms[L.T_L1_SUM])


def _grad_buffers(params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """The .grad tensors of leaf parameters (created zero-filled when absent) so the kernels accumulate in place."""
    out = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        out.append(p.grad)
    return out


class _Scratch:
    """Per-(device, size) scratch tensors of the fused step, reused across steps (pointer-stable for CUDA graphs)."""
    _cache = {}

    @classmethod
    def get(cls, key, nbytes_or_shape, dtype, device):
        k = (key, str(device))
        t = cls._cache.get(k)
        shape = (int(nbytes_or_shape),) if not isinstance(nbytes_or_shape, tuple) else nbytes_or_shape
        if t is None or t.dtype != dtype or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, device=device)
            cls._cache[k] = t
        return t

    @classmethod
    def clear(cls):
        cls._cache.clear()


def _grad_struct(spec: FieldSpec, grads: Sequence[torch.Tensor]) -> L.FieldGradsStruct:
    glat, gws, gbs = spec.split(grads)
    g = L.FieldGradsStruct()
    g.latents = L.ptr(glat)
    for k, (w, b) in enumerate(zip(gws, gbs)):
        g.weight[k] = L.ptr(w)
        g.bias[k] = L.ptr(b)
    return g


class StepPlan:
    """Everything one fused training step needs that does not change from step to step: the field / gradient descriptors over
    the models' parameter tensors, scratch (raw outputs, their gradients, activation stash, workspace) and the ray-sum output
    buffer, for a fixed (n_rays, n_depth).  `run` only patches the per-step pointers into the prepared C structs and makes ONE
    C call, so it is allocation-free (usable inside a nerfca_graph_* capture bracket) and cheap on the host."""

    def __init__(self, static_model, temp_model, n_rays: int, n_depth: int, device, output_activation: str, shared_scratch=True):
        lib = L.load()
        self.models = [static_model] + ([temp_model] if temp_model is not None else [])
        self.n_rays, self.n_depth, self.device = int(n_rays), int(n_depth), torch.device(device)
        self.prec = self.models[0]._precision_code()
        if any(m._precision_code() != self.prec for m in self.models):
            raise ValueError("both fields must use the same precision in a fused step")
        self.specs = [m._spec() for m in self.models]
        self.params = [_check_params(sp, m._param_list()) for sp, m in zip(self.specs, self.models)]
        self.grads = [_grad_buffers(m._param_list()) for m in self.models]
        self.fstructs = [sp.struct(p) for sp, p in zip(self.specs, self.params)]
        self.gstructs = [_grad_struct(sp, g) for sp, g in zip(self.specs, self.grads)]
        B, P, n = self.n_rays, self.n_rays * self.n_depth, len(self.models)
        dev = self.device
        self.pix = torch.empty((B,), dtype=torch.float64, device=dev)
        self.samples = L.SamplesStruct()
        self.samples.n_points, self.samples.n_rays, self.samples.n_depth = P, B, self.n_depth
        self.loss = L.LossCfgStruct()
        st = self.step = L.StepStruct()
        st.static_field, st.static_grads = C.pointer(self.fstructs[0]), C.pointer(self.gstructs[0])
        if n > 1:
            st.dynamic_field, st.dynamic_grads = C.pointer(self.fstructs[1]), C.pointer(self.gstructs[1])
        st.samples, st.loss = C.pointer(self.samples), C.pointer(self.loss)
        st.precision, st.activation = self.prec, activation_code(output_activation)
        get = _Scratch.get if shared_scratch else (lambda key, shape, dtype, d: torch.empty(shape if isinstance(shape, tuple) else (int(shape),), dtype=dtype, device=d))
        self.raw = get(("raw", n, P), (4 if n > 1 else 2, P), torch.float32, dev)
        st.raw_s, st.d_raw_s = L.ptr(self.raw[0]), L.ptr(self.raw[1])
        if n > 1:
            st.raw_d, st.d_raw_d = L.ptr(self.raw[2]), L.ptr(self.raw[3])
        if P > 0:
            self.stash = get(("stash", n, P, self.prec), lib.nerfca_step_stash_bytes(C.byref(st)), torch.uint8, dev)
            self.ws = get(("ws", n, P, self.prec), lib.nerfca_step_workspace_bytes(C.byref(st)), torch.uint8, dev)
            st.stash, st.workspace = L.ptr(self.stash), L.ptr(self.ws)
        st.pix_out = L.ptr(self.pix)
        self.refresh_bands()

    def refresh_bands(self):
        """Re-read the per-band window tensors of the models (they change with the iteration; an H2D copy when they do -- call it
        outside a capture bracket)."""
        for m, sp, fs in zip(self.models, self.specs, self.fstructs):
            if sp.enc_mode == L.ENC_BANDS:
                sp.band_weight = m._band_weight(sp.n_freq)
                fs.band_weight = L.ptr(sp.band_weight)

    def repack_struct(self) -> L.RepackStruct:
        r = L.RepackStruct()
        r.static_field = C.pointer(self.fstructs[0])
        if len(self.models) > 1:
            r.dynamic_field = C.pointer(self.fstructs[1])
        r.workspace = L.ptr(self.ws)
        return r

    def run(self, rays: torch.Tensor, phases, i0: torch.Tensor, depth: torch.Tensor, cfg: "LossConfig", terms: torch.Tensor,
            flags: int = 0):
        """rays [B,4,3] f64 (contiguous rows), phases [B] int32 or None, i0 [B] f32, depth [N] f32 -- all on the device."""
        B = self.n_rays
        if tuple(rays.shape) != (B, 4, 3) or rays.dtype != torch.float64 or not rays.is_contiguous():
            raise ValueError(f"rays must be a contiguous float64 [{B},4,3] tensor")
        if depth.dtype != torch.float32 or depth.numel() != self.n_depth or not depth.is_contiguous():
            raise ValueError("depth must be a contiguous float32 [n_depth] tensor")
        if B == 0:
            return terms
        s = self.samples
        base = rays.data_ptr()
        s.origins, s.dirs, s.depth = base, base + 3 * 8, L.ptr(depth)
        s.ray_dtype, s.ray_stride = L.F64, 12
        if len(self.models) > 1:
            if phases is None or phases.dtype != torch.int32 or phases.numel() != B:
                raise ValueError("phases must be an int32 [B] tensor")
            s.phase_ray = L.ptr(phases)
        st = self.step
        st.i0, st.gt, st.wpix, st.gw_stride = L.ptr(i0), base + 6 * 8, base + 9 * 8, 12
        c = self.loss
        c.favor_s_weight, c.dyn_entropy_weight = float(cfg.favor_s_weight), float(cfg.dyn_entropy_weight)
        c.occl_weight, c.l1_weight = float(cfg.occl_weight), float(cfg.l1_weight)
        c.entro_mask_thre, c.entro_weighted_thresh = float(cfg.entro_mask_thre), float(cfg.entro_weighted_thresh)
        c.entro_use_weighting = int(bool(cfg.entro_use_weighting))
        c.n_rays_global = int(cfg.n_rays_global or B)
        st.terms_out, st.flags = L.ptr(terms), int(flags)
        L.check(L.load().nerfca_train_step(C.byref(st), L.stream_ptr()), "nerfca_train_step")
        return terms


def _fused_step(static_model, temp_model, rays: torch.Tensor, phases, i0: torch.Tensor, depth: torch.Tensor, output_activation: str,
                cfg: LossConfig, terms: Optional[torch.Tensor]):
    """nerfca_train_step: forward of both fields (one launch), line integral + losses + dL/d_raw, both backward passes."""
    if rays.dim() != 3 or rays.shape[1:] != (4, 3):
        raise ValueError("rays must be [B,4,3] (origin, direction, pixel, weight rows)")
    dev = rays.device
    rays = rays.detach().to(torch.float64).contiguous()
    z = depth.detach().to(device=dev, dtype=torch.float32).contiguous()
    plan = StepPlan(static_model, temp_model, rays.shape[0], z.shape[0], dev, output_activation)
    ph = None
    if temp_model is not None:
        ph = phases.detach().flatten().to(device=dev).to(torch.int64).to(torch.int32).contiguous()
    if terms is None:
        terms = torch.zeros((L.N_LOSS_TERMS,), dtype=torch.float64, device=dev)
    i0f = i0.detach().to(device=dev, dtype=torch.float32).contiguous()
    plan.run(rays, ph, i0f, z, cfg, terms)
    return terms, plan.pix


def train_step_composite(static_model, temp_model, rays: torch.Tensor, phases: torch.Tensor, i0: torch.Tensor,
                         depth: torch.Tensor, output_activation: str, cfg: LossConfig, terms: Optional[torch.Tensor] = None):
    """One fused composite training step (train/run_composite.py:262-305 minus the optimizer):

        rays [B,4,3] float64 (rows: origin, direction, pixel x3, weight x3 -- data_helpers.py:161-163), phases [B],
        depth = the already-jittered depth vector [N].

    Runs both fields forward in one launch (points formed in-kernel), the fused integral + loss + dL/d_raw kernel and
    the backward passes; parameter gradients are ACCUMULATED into `.grad`.  Returns (terms[16] float64 device sums, pix[B]).
    """
    return _fused_step(static_model, temp_model, rays, phases, i0, depth, output_activation, cfg, terms)


def train_step_static(static_model, rays: torch.Tensor, i0: torch.Tensor, depth: torch.Tensor, output_activation: str,
                      occl_weight: float, n_rays_global: int = 0, terms: Optional[torch.Tensor] = None):
    """Fused static training step (train/run_nerf.py:205-233 minus the optimizer): loss = wMSE + occl_weight * occlusion."""
    cfg = LossConfig(occl_weight=occl_weight, n_rays_global=n_rays_global)
    return _fused_step(static_model, None, rays, None, i0, depth, output_activation, cfg, terms)


def render_frame(static_model, temp_model, origins: torch.Tensor, dirs: torch.Tensor, depth: torch.Tensor, phase,
                 i0_value: float, output_activation: str = "softplus", rays_per_pass: int = 1 << 18):
    """No-grad full-frame render (train/run_composite.py:346-361, 407-413): float32 rays -> (pix, pix_static, pix_dynamic)
    [n_rays] float32.  Rays are processed in passes so no [W*H*N,3] point tensor of the whole frame ever exists.
    fused tcgen05 path (bf16, hidden 128): nerfca_render_rays -- the line integral is fused into the output layer's epilogue, so
    neither the per-sample field outputs nor the sigma arrays reach HBM (one forward launch + a tiny finalize per pass).
    fp32 path and bf16 shapes on the layer-wise GEMM path: nerfca_fields_forward + three nerfca_integrate launches per pass."""
    o = origins.reshape(-1, 3).to(torch.float32).contiguous()
    d = dirs.reshape(-1, 3).to(torch.float32).contiguous()
    n = o.shape[0]
    dev = o.device
    act = activation_code(output_activation)
    z = depth.to(device=dev, dtype=torch.float32).contiguous()
    N = z.shape[0]
    outs = [torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(3)]
    spec_s, par_s, prec_s = static_model._spec(), [p.detach() for p in static_model._param_list()], static_model._precision_code()
    dyn = temp_model is not None
    if dyn:
        spec_d, par_d, prec_d = temp_model._spec(), [p.detach() for p in temp_model._param_list()], temp_model._precision_code()
        if prec_d != prec_s:
            raise ValueError("both fields must use the same precision")
    lib = L.load()
    fs_s = spec_s.struct(_check_params(spec_s, par_s))
    fs_d = spec_d.struct(_check_params(spec_d, par_d)) if dyn else None
    for r0 in range(0, n, rays_per_pass):
        r1 = min(n, r0 + rays_per_pass)
        B = r1 - r0
        ph = None
        if dyn:
            ph = phase[r0:r1] if torch.is_tensor(phase) and phase.numel() > 1 else torch.full((B,), int(phase), device=dev)
        smp = Samples.from_rays(o[r0:r1], d[r0:r1], z, ph)
        i0 = _Scratch.get(("render_i0", B, float(i0_value)), (B,), torch.float32, dev)
        i0.fill_(i0_value)
        st = L.stream_ptr()
        # (0 = these fields are not served by the fused kernels -- fp32, or bf16 shapes on the layer-wise GEMM path)
        need = lib.nerfca_render_workspace_bytes(C.byref(fs_s), C.byref(fs_d) if dyn else None, C.byref(smp.struct()), prec_s)
        if need > 0:
            ws = _Scratch.get(("render_ws", B * N, prec_s, dyn), need, torch.uint8, dev)
            L.check(lib.nerfca_render_rays(C.byref(fs_s), C.byref(fs_d) if dyn else None, C.byref(smp.struct()), prec_s, L.ptr(i0), act,
                                           L.ptr(outs[0][r0:r1]), L.ptr(outs[1][r0:r1]) if dyn else None,
                                           L.ptr(outs[2][r0:r1]) if dyn else None, L.ptr(ws), st), "nerfca_render_rays")
            if not dyn:
                outs[1][r0:r1] = outs[0][r0:r1]
            continue
        raws = _Scratch.get(("render_raw", B * N), (2, B * N), torch.float32, dev)
        raw_s, raw_d = raws[0], (raws[1] if dyn else None)
        stp = L.StepStruct()
        stp.static_field = C.pointer(fs_s)
        if dyn:
            stp.dynamic_field = C.pointer(fs_d)
        stp.samples = C.pointer(smp.struct())
        stp.precision = prec_s
        ws = _Scratch.get(("render_ws", B * N, prec_s, dyn), lib.nerfca_step_workspace_bytes(C.byref(stp)), torch.uint8, dev)
        L.check(lib.nerfca_fields_forward(C.byref(fs_s), C.byref(fs_d) if dyn else None, C.byref(smp.struct()), prec_s, L.ptr(raw_s),
                                          L.ptr(raw_d), L.ptr(ws), st), "nerfca_fields_forward")
        sig = _Scratch.get(("render_sig", B * N), (2, B, N), torch.float32, dev)
        sig_a, sig_b = sig[0], sig[1]
        if dyn:
            L.check(lib.nerfca_integrate(L.ptr(raw_s), L.ptr(raw_d), L.ptr(z), L.ptr(i0), B, N, act, L.F32, L.ptr(outs[0][r0:r1]),
                                         L.ptr(sig_a), L.ptr(sig_b), None, st), "nerfca_integrate")
            L.check(lib.nerfca_integrate(L.ptr(raw_d), None, L.ptr(z), L.ptr(i0), B, N, act, L.F32, L.ptr(outs[2][r0:r1]),
                                         L.ptr(sig_a), None, None, st), "nerfca_integrate")
        L.check(lib.nerfca_integrate(L.ptr(raw_s), None, L.ptr(z), L.ptr(i0), B, N, act, L.F32, L.ptr(outs[1][r0:r1]),
                                     L.ptr(sig_a), None, None, st), "nerfca_integrate")
        if not dyn:
            outs[0][r0:r1] = outs[1][r0:r1]
    return outs[0], outs[1], outs[2]


def normalize_image(img: torch.Tensor):
    """N4 (train/run_composite.py:394-413): (img - min) / (max - min) on the device -> (normalised image, float32 [2] = (min, max))."""
    x = img.detach().to(torch.float32).contiguous()
    out = torch.empty_like(x)
    mm = torch.empty((2,), dtype=torch.float32, device=x.device)
    scratch = torch.empty((2,), dtype=torch.int32, device=x.device)
    L.check(L.load().nerfca_normalize_image(L.ptr(x), x.numel(), L.ptr(out), L.ptr(mm), L.ptr(scratch), L.stream_ptr()), "nerfca_normalize_image")
    return out, mm


def eval_frame(static_model, temp_model, origins, dirs, depth, phase, i0_value, gt_img=None, weights=None,
               output_activation: str = "softplus"):
    """The display block of train/run_composite.py:346-444 for one test frame: the composite / static / dynamic projections, their
    min-max normalised display versions, and -- when the ground-truth frame is given -- the weighted pixel loss (:363) and its PSNR
    -10 log10(loss).  Everything stays on the device; returns a dict of tensors."""
    pix, pix_s, pix_d = render_frame(static_model, temp_model, origins, dirs, depth, phase, i0_value, output_activation)
    out = {"pix": pix, "pix_static": pix_s, "pix_dynamic": pix_d}
    for k in ("pix", "pix_static", "pix_dynamic"):
        out[k + "_norm"], out[k + "_minmax"] = normalize_image(out[k])
    if gt_img is not None:
        gt = gt_img.reshape(-1).to(pix.device, torch.float32)
        w = torch.ones_like(gt) if weights is None else weights.reshape(-1).to(pix.device, torch.float32)
        out["pixel_loss"] = (((pix - gt) ** 2) * w).mean()
        out["psnr"] = -10.0 * torch.log10(out["pixel_loss"])
    return out

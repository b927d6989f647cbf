"""CPU oracle for the NeRF-CA training / rendering inner loop.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`nerf-ca_b200/`) may import this
file; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs do, and there only as the checker / the CPU arm that is timed *beside* the
GPU path.

It is a functional restatement (numpy + CPU torch) of the reference's algorithm for the hot
path named in SURVEY.md section 8(a), rows A1-A11.  Every function cites the reference
file:line it follows (paths relative to the upstream repo root).  The arithmetic of the
reference lives in PyTorch eager ops (reference pins no version; torch 2.11.0 here), so the
restatement uses the same CPU torch primitives in the same order and dtypes, including the
float64 promotions of the training path.

Parity pin: the reference has no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the authoring container by
`tests/golden/make_golden.py` (imports the unmodified modules from /root/reference) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every fixture.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = np.float32

# --------------------------------------------------------------------------------------------
# A1  pose  (train/proj_helpers.py:5-63)
# --------------------------------------------------------------------------------------------


def _rot_x(a: float) -> np.ndarray:
    # train/proj_helpers.py:5-11
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def _rot_z(a: float) -> np.ndarray:
    # train/proj_helpers.py:21-27
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)


def pose_tigre(theta_deg: float, phi_deg: float, dso: float) -> np.ndarray:
    """4x4 float64 source pose.  train/proj_helpers.py:50-63 (larm is accepted and ignored there).

    rot = Rz(-theta) . Rz(pi/2) . Rx(phi) . Rx(-pi/2);  pose = rot . T(0, 0, -DSO)
    """
    r1 = _rot_x(-np.pi / 2)
    r2 = _rot_x(np.deg2rad(phi_deg))
    r3 = _rot_z(np.pi / 2)
    r4 = _rot_z(-np.deg2rad(theta_deg))
    rot = np.dot(np.dot(r4, np.dot(r3, r2)), r1)
    t = np.identity(4)
    t[:3, 3] = [0.0, 0.0, -dso]
    return rot.dot(t)


# --------------------------------------------------------------------------------------------
# A2  per-pixel rays  (train/proj_helpers.py:65-90)
# --------------------------------------------------------------------------------------------


def rays_tigre(theta_deg: float, phi_deg: float, geo: dict) -> Tuple[np.ndarray, np.ndarray]:
    """(origins[W,H,3], directions[W,H,3]) float32.  train/proj_helpers.py:65-90.

    Restated as explicit fp32 scalar arithmetic: pose and every geometry scalar are rounded
    to fp32 first (:68, torch scalar promotion); u = ((i + 0.5) - W/2) * du + off_u with each
    op rounded (:79); dx = u / DSD (true division, :81); dir_k = (R_k0*dx + R_k1*dy) + R_k2*1
    with separate multiply and add (the batched 3x3 matmul of :83, no FMA).
    """
    pose = pose_tigre(theta_deg, phi_deg, geo["DSO"]).astype(F32)
    w, h = int(geo["nDetector"][0]), int(geo["nDetector"][1])
    du, dv = F32(geo["dDetector"][0]), F32(geo["dDetector"][1])
    ou, ov = F32(geo["offDetector"][0]), F32(geo["offDetector"][1])
    dsd = F32(geo["DSD"])
    i = np.arange(w, dtype=F32)
    j = np.arange(h, dtype=F32)
    u = ((i + F32(0.5)) - F32(w / 2)) * du + ou          # [W]
    v = ((j + F32(0.5)) - F32(h / 2)) * dv + ov          # [H]
    dx = (u / dsd)[:, None].astype(F32)                  # [W,1]
    dy = (v / dsd)[None, :].astype(F32)                  # [1,H]
    rot = pose[:3, :3]
    dirs = np.empty((w, h, 3), dtype=F32)
    for k in range(3):
        a = (rot[k, 0] * dx).astype(F32)
        b = (rot[k, 1] * dy).astype(F32)
        dirs[:, :, k] = ((a + b).astype(F32) + rot[k, 2] * F32(1.0)).astype(F32)
    origins = np.broadcast_to(pose[:3, 3], dirs.shape).copy()
    return origins, dirs


# --------------------------------------------------------------------------------------------
# A0  ray table layout  (train/data_helpers.py:129-165)
# --------------------------------------------------------------------------------------------


def build_ray_table(frames: Sequence[dict], geo: dict, weighted_loss_max: float = 1.0):
    """rays_train[N_img*W*H, 4, 3] float64 and phases_train[N_img*W*H] int64.

    train/data_helpers.py:141-165.  `frames` rows carry theta, phi, heart_phase, `image`
    (already denormalised and transposed to [W,H], :129-139) and `weight` ([W,H] in [1,2]).
    Rows of the table: origin, direction, pixel x3, weight x3; ray id = img*W*H + i*H + j.
    """
    w, h = int(geo["nDetector"][0]), int(geo["nDetector"][1])
    rows, phases = [], []
    for fr in frames:
        o, d = rays_tigre(fr["theta"], fr["phi"], geo)
        img = np.repeat(np.asarray(fr["image"], dtype=np.float64)[:, :, None], 3, axis=-1)
        wimg = (np.asarray(fr["weight"], dtype=np.float64) - 1) * weighted_loss_max + 1
        wimg = np.repeat(wimg[:, :, None], 3, axis=-1)
        tab = np.stack([o.astype(np.float64), d.astype(np.float64), img, wimg], axis=2)  # [W,H,4,3]
        rows.append(tab.reshape(w * h, 4, 3))
        phases.append(np.full((w * h,), int(fr["heart_phase"]), dtype=np.int64))
    return np.concatenate(rows, 0), np.concatenate(phases, 0)


# --------------------------------------------------------------------------------------------
# A3  depth grid + stratified jitter  (train/data_helpers.py:167-171, train/model_helpers.py:3-12)
# --------------------------------------------------------------------------------------------


def depth_values(near: float, far: float, n: int) -> torch.Tensor:
    # train/data_helpers.py:167-171
    t = torch.linspace(0.0, 1.0, n)
    return near * (1.0 - t) + far * t


def jitter_depth(z: torch.Tensor, t_rand: torch.Tensor) -> torch.Tensor:
    """train/model_helpers.py:3-12 with the U[0,1) draw passed in (the reference draws it from
    the CPU generator, :8; callers reproduce that with torch.manual_seed + torch.rand)."""
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.concat([mids, z[..., -1:]], -1)
    lower = torch.concat([z[..., :1], mids], -1)
    return lower + (upper - lower) * t_rand


# --------------------------------------------------------------------------------------------
# A4  sample points  (train/model_helpers.py:101-104,118-121; eval train/run_composite.py:351-352)
# --------------------------------------------------------------------------------------------


def sample_points(origins: torch.Tensor, dirs: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """[B*N,3] float32.  Same expression for both paths; the dtypes of the inputs decide the
    rounding: float64 o,d (training, ray table) -> fl32(fl64(o + d*z)); float32 o,d (eval)
    -> o + fl32(d*z)."""
    q = origins[..., None, :] + dirs[..., None, :] * z[..., :, None]
    return q.reshape((-1, 3)).float()


# --------------------------------------------------------------------------------------------
# A5  positional encoding  (model/CPPN.py:112-159, model/Temporal.py:153-204)
# --------------------------------------------------------------------------------------------


def freq_mask(n_freq: int, cur_iter: int, max_iter: int, window_start: float):
    """(freq_mask_alpha float32 [L], windowed_alpha).  model/CPPN.py:144-159."""
    if cur_iter < max_iter:
        m = np.zeros(n_freq)
        ptr = (n_freq * cur_iter) / max_iter + window_start
        ip = int(ptr)
        m[: ip + 1] = 1.0
        m[ip: ip + 1] = ptr - ip
        return torch.clip(torch.from_numpy(m), 1e-8, 1 - 1e-8).float(), ptr
    return torch.ones(n_freq).float(), n_freq + 1


def nerfies_window(n_freq: int, alpha: float) -> torch.Tensor:
    # model/CPPN.py:137-142
    bands = torch.arange(0, n_freq)
    x = torch.clip(alpha - bands, 0.0, 1.0)
    return 0.5 * (1 + torch.cos(torch.pi * x + torch.pi))


def pos_enc(x: torch.Tensor, n_freq: int, mode: str, window: Optional[torch.Tensor] = None,
            fourier_coeff: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/CPPN.py:112-135.  mode 'fourier' -> [sin | cos](2 pi x_rep * coeff); any other
    non-'none' mode -> [x, per band: sin(2^l x) x3, sin(2^l x + pi/2) x3] times `window[l]`
    for the two windowed modes ('free_windowed': freq_mask_alpha, 'nerfies_windowed': eased)."""
    if mode == "none" or n_freq <= 0:
        return x
    if mode == "fourier":
        basis = torch.cat(n_freq * [x], dim=-1)
        val = 2 * np.pi * basis * fourier_coeff
        return torch.cat([torch.sin(val), torch.cos(val)], dim=-1)
    shape = x.shape[:-1]
    scales = (2.0 ** torch.arange(0, n_freq)).to(x.device)
    xb = x[..., None, :] * scales[:, None]
    feat = torch.sin(torch.stack([xb, xb + 0.5 * torch.pi], axis=-2))
    if mode in ("free_windowed", "nerfies_windowed"):
        feat = window[..., None, None] * feat
    feat = feat.reshape((*shape, -1))
    return torch.cat([x, feat], dim=-1)


# --------------------------------------------------------------------------------------------
# A6 / A7  fields  (model/CPPN.py:88-110, model/Temporal.py:113-151)
# --------------------------------------------------------------------------------------------


def _mlp(inp: torch.Tensor, sd: Dict[str, torch.Tensor], n_hidden: int) -> torch.Tensor:
    """input -> H, n_hidden x (H -> H) with ReLU, H -> 1.  Keys are the reference state_dict
    keys: early_pts_layers.{0,2,..}.{weight,bias} (ReLU modules sit at the odd indices of the
    ModuleList, CPPN.py:40-50) and output_linear.0.{weight,bias} (:63-65)."""
    h = inp
    for k in range(n_hidden + 1):
        w = sd[f"early_pts_layers.{2 * k}.weight"]
        b = sd.get(f"early_pts_layers.{2 * k}.bias")
        h = torch.relu(torch.nn.functional.linear(h, w, b))
    return torch.nn.functional.linear(h, sd["output_linear.0.weight"], sd.get("output_linear.0.bias"))


def static_field(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict) -> torch.Tensor:
    """CPPN.forward, num_late_layers == 0 branch.  model/CPPN.py:88-110."""
    enc = pos_enc(x, cfg["n_freq"], cfg["pos_enc"], cfg.get("window"), cfg.get("fourier_coeff"))
    return _mlp(enc, sd, cfg["n_hidden"])


def dynamic_field(x: torch.Tensor, phases: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict) -> torch.Tensor:
    """Temporal.forward_composite -> query_time.  model/Temporal.py:138-151, :113-136."""
    lat = sd["time_latents"][phases.flatten().long()]
    enc = pos_enc(x, cfg["n_freq"], cfg["pos_enc"], cfg.get("window"), cfg.get("fourier_coeff"))
    return _mlp(torch.cat([enc, lat], dim=-1), sd, cfg["n_hidden"])


def chunked(fn, chunk: int, *tensors):
    """The reference's chunk loop (train/model_helpers.py:14-61): split dim 0, cat results."""
    n = tensors[0].shape[0]
    outs = [fn(*[t[i:i + chunk] for t in tensors]) for i in range(0, n, chunk)]
    return torch.cat(outs, dim=0)


# --------------------------------------------------------------------------------------------
# A9  line integral  (train/model_helpers.py:63-97)
# --------------------------------------------------------------------------------------------


def activation(name: str):
    # train/model_helpers.py:63-70: anything but 'softplus' / 'clamp' selects Sigmoid
    if name == "softplus":
        return torch.nn.functional.softplus
    if name == "clamp":
        return lambda v: torch.nn.functional.hardtanh(torch.nn.functional.softplus(v), min_val=0.0, max_val=1.0)
    return torch.sigmoid


def dists_from_depth(z: torch.Tensor, like_dtype: torch.dtype) -> torch.Tensor:
    # train/model_helpers.py:73-74: last delta = 1e-10 in ray_directions.dtype
    e = torch.tensor([1e-10], dtype=like_dtype, device=z.device)
    return torch.cat((z[..., 1:] - z[..., :-1], e.expand(z[..., :1].shape)), dim=-1)


def integrate_composite(raw_s: torch.Tensor, raw_d: torch.Tensor, i0: torch.Tensor, dirs_dtype: torch.dtype,
                        z: torch.Tensor, act: str = "softplus", scale: float = 1e-2):
    """train/model_helpers.py:72-84.  raw_*: [B,N,1].  Returns pix[B], sigma_s, sigma_d [B,N], dists[N]."""
    d = dists_from_depth(z, dirs_dtype)
    f = activation(act)
    ss = f(raw_s[..., -1]) * scale
    sd_ = f(raw_d[..., -1]) * scale
    wts = (ss + sd_) * d
    return i0 - torch.sum(wts, dim=-1), ss, sd_, d


def integrate_single(raw: torch.Tensor, i0: torch.Tensor, dirs_dtype: torch.dtype, z: torch.Tensor,
                     act: str = "softplus", scale: float = 1e-2):
    """train/model_helpers.py:86-97.  NOTE sigma is returned UNscaled here (:91-92)."""
    d = dists_from_depth(z, dirs_dtype)
    sig = activation(act)(raw[..., -1])
    wts = sig * d * scale
    return i0 - torch.sum(wts, dim=-1), sig, d


# --------------------------------------------------------------------------------------------
# A10  losses  (train/model_helpers.py:189-289)
# --------------------------------------------------------------------------------------------


def weighted_mse(pred, gt, w):
    # train/model_helpers.py:284-289 (callers take .mean())
    return ((pred - gt) ** 2) * w


def blend_ratio(ss, sd_):
    # train/model_helpers.py:189-198
    return sd_ / (ss + sd_ + 1e-10)


def blend_entropy(blendw, clip=1e-19, skew=1):
    # train/model_helpers.py:200-204
    b = torch.clip(blendw ** skew, min=clip, max=1 - clip)
    r = torch.clip(1 - b, min=clip)
    return torch.mean(-(b * torch.log(b) + r * torch.log(r)), dim=-1).mean()


def ray_entropy(sig, d, mask_thre=0.1, clip=1e-19, use_weighting=False, wpix=(), wthresh=0.25):
    # train/model_helpers.py:206-224
    sdist = sig * d
    ssum = torch.sum(sdist, dim=-1, keepdim=True)
    mask = torch.where(ssum < mask_thre, 0.0, 1.0).flatten().int()
    if len(wpix) > 0 and use_weighting:
        wm = torch.zeros(mask.shape, device=mask.device).int()
        wm[: wpix.shape[0]] = torch.where(wpix > 1 + wthresh, 1.0, 0.0).int()
        mask = torch.bitwise_or(wm, mask)
    p = sdist / torch.clip(ssum, min=clip)
    ent = mask * -torch.sum(p * torch.log(p + 1e-10), dim=-1)
    return ent.mean(), ssum.mean()


def occlusion(sig, d, reg_perc=0.1, use_back=False):
    # train/model_helpers.py:226-248 (mask is all ones unless use_back; no caller sets it)
    cum = torch.cumsum(d, dim=0).unsqueeze(dim=0).repeat((sig.shape[0], 1))
    front = torch.where(cum < reg_perc * cum[-1, -1], 1.0, 0.0).int()
    back = torch.ones(front.shape, device=front.device)
    if use_back:
        back = torch.where(cum > (1 - reg_perc) * cum[-1, -1], 1.0, 0.0)
    mask = torch.bitwise_or(front, back.int())
    return torch.sum(sig * d * mask, dim=-1).mean()


def composite_losses(ss, sd_, d, wpix, hp: dict):
    """train/model_helpers.py:250-262 -> the 11-tuple, same order."""
    bw = blend_ratio(ss, sd_)
    with torch.no_grad():
        smax, dmax = torch.max(ss), torch.max(sd_)
    favor = blend_entropy(bw, skew=hp.get("skewness_val", 1))
    s_ent, s_sum = ray_entropy(ss, d, mask_thre=hp["entro_mask_thre"])
    d_ent, d_sum = ray_entropy(sd_, d, mask_thre=hp["entro_mask_thre"], use_weighting=hp["entro_use_weighting"],
                               wpix=wpix, wthresh=hp["entro_weighted_thresh"])
    occl = occlusion(sd_, d, hp["occl_reg_perc"])
    l1 = torch.sum(ss * d, dim=-1).sum()
    l2 = torch.sum((ss * d) ** 2, dim=-1).sum()
    return bw.mean(), smax, dmax, favor, s_ent, s_sum, d_ent, d_sum, occl, l1, l2


def linear_decay(it, start, end, steps, delay=0):
    # train/model_helpers.py:264-269
    if it < delay:
        return 0
    a = min((it - delay) / steps, 1.0)
    return (1.0 - a) * start + a * end


def schedule_weights(it: int, hp: dict) -> dict:
    # train/run_composite.py:276-279
    return {
        "favor_s": linear_decay(it, hp["favor_s_weight_start"], hp["favor_s_weight_end"], hp["hyperparam_decay_steps"],
                                hp["favor_s_weight_delay_steps"]),
        "dyn_entro": linear_decay(it, hp["dynamic_entro_weight_start"], hp["dynamic_entro_weight_end"],
                                  hp["hyperparam_decay_steps"]),
        "occl": linear_decay(it, hp["occl_weight_start"], hp["occl_weight_end"], hp["hyperparam_decay_steps"],
                             hp["favor_s_weight_delay_steps"]),
        "l1": linear_decay(it, hp["l1_weight_start"], hp["l1_weight_end"], hp["hyperparam_decay_steps"]),
    }


COMPOSITE_HP = {  # train/composite.txt:45-66 (the shipped config-2 hyper-parameters)
    "entro_mask_thre": 1e-4, "entro_use_weighting": True, "entro_weighted_thresh": 0.03,
    "favor_s_weight_start": 1e-12, "favor_s_weight_end": 1e-10, "favor_s_weight_delay_steps": 40000,
    "dynamic_entro_weight_start": 1e-10, "dynamic_entro_weight_end": 1e-8,
    "occl_weight_start": 1e-8, "occl_weight_end": 1e-4,
    "l1_weight_start": 1e-8, "l1_weight_end": 1e-15,
    "hyperparam_decay_steps": 100000, "occl_reg_perc": 0.2, "skewness_val": 1,
}


# --------------------------------------------------------------------------------------------
# whole steps (the functions the drivers call)
# --------------------------------------------------------------------------------------------


def composite_forward(sd_static, sd_dyn, cfg_s, cfg_d, origins, dirs, phases_ray, i0, z, act="softplus",
                      chunk=32768):
    """obtain_train_predictions_iter without the fine pass.  train/model_helpers.py:115-129.
    `z` is the already-jittered depth vector (A3)."""
    b, n = origins.shape[0], z.shape[0]
    pts = sample_points(origins, dirs, z)
    ph = phases_ray[:, None].repeat(1, n).flatten().int()           # run_composite.py:265, model_helpers.py:122
    raw_s = chunked(lambda p: static_field(p, sd_static, cfg_s), chunk, pts)
    raw_d = chunked(lambda p, t: dynamic_field(p, t, sd_dyn, cfg_d), chunk, pts, ph)
    raw_s = raw_s.reshape(b, n, 1)
    raw_d = raw_d.reshape(b, n, 1)
    return integrate_composite(raw_s, raw_d, i0, dirs.dtype, z, act)


def composite_step_loss(sd_static, sd_dyn, cfg_s, cfg_d, origins, dirs, phases_ray, i0, z, gt, wpix, hp, it,
                        act="softplus", chunk=32768):
    """One training-step loss (train/run_composite.py:283-292).  Returns (loss, dict of terms)."""
    pix, ss, sd_, d = composite_forward(sd_static, sd_dyn, cfg_s, cfg_d, origins, dirs, phases_ray, i0, z, act, chunk)
    pixel = weighted_mse(pix, gt, wpix).mean()
    terms = composite_losses(ss, sd_, d, wpix, hp)
    w = schedule_weights(it, hp)
    loss = pixel + w["favor_s"] * terms[3] + w["dyn_entro"] * terms[6] + w["occl"] * terms[8] \
        + w["l1"] * terms[10] + w["l1"] * terms[9]
    names = ["blendw", "sigma_s_max", "sigma_d_max", "favor_s", "s_entropy", "s_entropy_sum", "d_entropy",
             "d_entropy_sum", "d_occl", "s_l1", "s_l2"]
    out = {k: v for k, v in zip(names, terms)}
    out.update(pixel=pixel, pix=pix, sigma_s=ss, sigma_d=sd_, dists=d)
    return loss, out


def sample_pdf(bins, weights, u):
    """Inverse-transform sampling of the fine depths (train/model_helpers.py:162-187) with the U[0,1) draw `u` [B, n_fine]
    passed in (the reference draws it from the CPU generator, :170)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = u.to(weights)
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


def fine_pass(sd_static_f, sd_dyn_f, cfg_s, cfg_d, origins, dirs, phases_ray, i0, z, ss_c, sd_c, u, act="softplus", chunk=32768):
    """The hierarchical fine pass of obtain_train_predictions_iter (train/model_helpers.py:131-158): importance weights from
    |delta (sigma_s + sigma_d)| of the coarse pass, n_fine extra depths per ray by sample_pdf, merged and sorted PER RAY with the
    coarse depths, both fine nets on the per-ray points, and the line integral with RAY 0's depths for every ray (:150, a quirk of
    the reference).  Returns pix_f, sigma_s_f, sigma_d_f, dists_f, z_fine [B, N + n_fine]."""
    b, n = origins.shape[0], z.shape[0]
    eps = torch.ones_like(ss_c[:, :1]) * 1e-10
    w = torch.cat([eps, torch.abs((ss_c[:, 1:] + sd_c[:, 1:]) - (ss_c[:, :-1] + sd_c[:, :-1]))], dim=-1)
    w = w / torch.max(w)
    zb = z[None, :].repeat(b, 1)
    mid = .5 * (zb[..., 1:] + zb[..., :-1])
    pdf_z = sample_pdf(mid, w[..., 1:-1], u)
    z_fine, _ = torch.sort(torch.cat([pdf_z, zb.detach()], -1), -1)
    total = z_fine.shape[-1]
    pts = (origins[..., None, :] + dirs[..., None, :] * z_fine[..., :, None]).reshape((-1, 3)).float()
    ph = phases_ray[:, None].repeat(1, total).flatten()
    raw_s = chunked(lambda p: static_field(p, sd_static_f, cfg_s), chunk, pts).reshape(b, total, 1)
    raw_d = chunked(lambda p, t: dynamic_field(p, t, sd_dyn_f, cfg_d), chunk, pts, ph).reshape(b, total, 1)
    pix, ss, sd_, d = integrate_composite(raw_s, raw_d, i0, dirs.dtype, z_fine[0, :], act)
    return pix, ss, sd_, d, z_fine


def static_step_loss(sd_static, cfg_s, origins, dirs, i0, z, gt, wpix, occl_weight, act="softplus", chunk=32768):
    """run_nerf.py training-step loss: obtain_train_predictions_static (model_helpers.py:99-113)
    + weighted MSE + occl_weight_start * compute_occl_loss (run_nerf.py:227-230)."""
    b, n = origins.shape[0], z.shape[0]
    pts = sample_points(origins, dirs, z)
    raw = chunked(lambda p: static_field(p, sd_static, cfg_s), chunk, pts).reshape(b, n, 1)
    pix, sig, d = integrate_single(raw, i0, dirs.dtype, z, act)
    pixel = weighted_mse(pix, gt, wpix).mean()
    occl = occlusion(sig, d)
    return pixel + occl_weight * occl, {"pixel": pixel, "occl": occl, "pix": pix, "sigma": sig, "dists": d}


def dloss_dsigma(ss, sd_, d, pix, gt, wpix, hp, it):
    """Closed-form dL/dsigma_s, dL/dsigma_d of the composite loss (SURVEY 8(a'));
    derived from train/model_helpers.py:189-262 + run_composite.py:287-292.  float64."""
    ss = ss.double(); sd_ = sd_.double(); d = d.double()
    b, n = ss.shape
    w = schedule_weights(it, hp)
    res = (pix.double() - gt.double())
    g_px = -(2.0 / b) * (wpix.double() * res)[:, None] * d[None, :]
    tot = ss + sd_ + 1e-10
    bw = sd_ / tot
    bh = torch.clip(bw, 1e-19, 1 - 1e-19)
    rh = torch.clip(1 - bh, min=1e-19)
    inside = ((bw > 1e-19) & (bw < 1 - 1e-19)).double()
    # d(-b ln b)/db = -(ln b + 1);  d(-r ln r)/db = +(ln r + 1) when r = 1 - b is not clipped
    r_live = ((1 - bh) > 1e-19).double()
    e1 = (-(torch.log(bh) + 1) + (torch.log(rh) + 1) * r_live) * inside
    g_fd = w["favor_s"] / (b * n) * e1 * (ss + 1e-10) / tot ** 2
    g_fs = w["favor_s"] / (b * n) * e1 * (-sd_) / tot ** 2
    a = sd_ * d[None, :]
    s_sum = a.sum(-1, keepdim=True)
    s_hat = torch.clip(s_sum, min=1e-19)
    p = a / s_hat
    m = (s_sum.flatten() >= hp["entro_mask_thre"])
    if hp["entro_use_weighting"]:
        m = m | (wpix > 1 + hp["entro_weighted_thresh"])
    m = m.double()[:, None]
    h = -(torch.log(p + 1e-10) + p / (p + 1e-10))
    live = (s_sum > 1e-19).double()
    g_ed = w["dyn_entro"] / b * m * (h - live * (h * p).sum(-1, keepdim=True)) / s_hat * d[None, :]
    g_od = w["occl"] / b * d[None, :].expand(b, n)
    g_ls = w["l1"] * (d[None, :] + 2 * ss * d[None, :] ** 2)
    return g_px + g_fs + g_ls, g_px + g_fd + g_ed + g_od


# --------------------------------------------------------------------------------------------
# helpers shared by tests / bench: model construction with reference-compatible state dicts
# --------------------------------------------------------------------------------------------


def init_field_state(in_dim: int, hidden: int, n_hidden: int, n_latent_rows: int = 0, n_latent: int = 0,
                     seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random weights with nn.Linear's default init distribution and the reference key names.
    (Used for synthetic benches; parity tests that compare against the reference modules load
    the reference's own state_dict from the golden fixtures instead.)"""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    if n_latent_rows:
        sd["time_latents"] = torch.rand((n_latent_rows, n_latent), generator=g)
    dims = [in_dim] + [hidden] * (n_hidden + 1)
    for k in range(n_hidden + 1):
        bound = 1.0 / math.sqrt(dims[k])
        sd[f"early_pts_layers.{2 * k}.weight"] = (torch.rand((hidden, dims[k]), generator=g) * 2 - 1) * bound
        sd[f"early_pts_layers.{2 * k}.bias"] = (torch.rand((hidden,), generator=g) * 2 - 1) * bound
    bound = 1.0 / math.sqrt(hidden)
    sd["output_linear.0.weight"] = (torch.rand((1, hidden), generator=g) * 2 - 1) * bound
    sd["output_linear.0.bias"] = (torch.rand((1,), generator=g) * 2 - 1) * bound
    return sd

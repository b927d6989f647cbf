"""Parity harness shared by the GPU tests, __graft_entry__.smoke() and bench.py's self-check:
builds the CUDA-backed drop-in modules with given weights, runs the CUDA path and the CPU oracle on the same
seeded synthetic inputs and returns / asserts the deviations.

Tolerances (north_star: "within a stated fp32/bf16 tolerance, ray indexing and sample positions bit-exact"):
  fp32 path : pix rtol 2e-5, sigma rtol 1e-4, every gradient tensor rel-L2 <= 2e-4
  bf16 path : pix atol 1e-4 * I0, sigma rtol 5e-2 / atol 2e-4, every gradient tensor rel-L2 <= 5e-2 and cosine >= 0.998
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nerf-ca_b200")
for _p in (ROOT, PKG, os.path.join(PKG, "train")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import nerfca_oracle as orc  # noqa: E402

I0 = float(np.log(8.670397))
GEO64 = {"DSD": 20.0, "DSO": 6.0, "nDetector": [64, 64], "dDetector": [200 * 0.01 / 64] * 2, "offDetector": [0.0, 0.0, 0.0]}
VIEWS = [(-30.0, 30.0), (-30.0, -30.0), (60.0, -30.0), (60.0, 30.0)]

TOL = {
    "fp32": dict(pix_rtol=2e-5, pix_atol=1e-7, sig_rtol=1e-4, sig_atol=1e-8, grad_rel=2e-4, grad_cos=0.999999),
    "bf16": dict(pix_rtol=0.0, pix_atol=1e-4 * I0, sig_rtol=5e-2, sig_atol=2e-4, grad_rel=5e-2, grad_cos=0.998),
}


def static_definition(device, hidden=128, n_early=4, n_freq=12, mode="free_windowed", precision=None, window_start=1):
    return {"num_early_layers": n_early, "num_late_layers": 0, "num_filters": hidden, "num_input_channels": 3,
            "num_output_channels": 1, "use_bias": True, "pos_enc": mode, "pos_enc_window_start": window_start,
            "pos_enc_basis": n_freq, "fourier_sigma": 0.0, "fourier_gaussian": None, "act_func": "relu", "device": device,
            "precision": precision}


def temporal_definition(device, hidden=128, n_early=4, n_freq=12, n_latent=8, mode="free_windowed", precision=None):
    d = static_definition(device, hidden, n_early, n_freq, mode, precision)
    d.update({"num_input_times": 1, "use_time_latents": True, "num_time_dim": n_latent})
    return d


def build_models(sd_s, sd_d, device, precision, hidden=128, n_early=4, n_freq=12, n_latent=8, mask=None):
    """CUDA-backed drop-in modules loaded with the given reference-keyed state dicts."""
    from model.CPPN import CPPN
    from model.Temporal import Temporal
    static = CPPN(static_definition(device, hidden, n_early, n_freq, precision=precision))
    static.load_state_dict({k: torch.as_tensor(v) for k, v in sd_s.items()})
    static.to(device)
    temp = None
    if sd_d is not None:
        temp = Temporal(temporal_definition(device, hidden, n_early, n_freq, n_latent, precision=precision))
        if sd_d["time_latents"].shape[0] != temp.time_latents.shape[0]:
            temp.time_latents = torch.nn.Parameter(torch.zeros(tuple(sd_d["time_latents"].shape)))
        temp.load_state_dict({k: torch.as_tensor(v) for k, v in sd_d.items()})
        temp.to(device)
    for m in (static, temp):
        if m is not None and mask is not None:
            m.freq_mask_alpha = torch.as_tensor(mask).float().clone()
    return static, temp


def synthetic_batch(n_rays, n_depth, seed, n_phases=10, near=3.2, far=8.8):
    """Ray-table rows [B,4,3] f64 + phases + jittered depth, reference layout (data_helpers.py:161-163)."""
    rng = np.random.default_rng(seed)
    tabs = []
    for th, phi in VIEWS:
        o, d = orc.rays_tigre(th, phi, GEO64)
        tabs.append(np.stack([o.reshape(-1, 3), d.reshape(-1, 3)], 1).astype(np.float64))
    tab = np.concatenate(tabs, 0)
    ids = rng.integers(0, tab.shape[0], size=n_rays)
    rays = np.zeros((n_rays, 4, 3))
    rays[:, :2] = tab[ids]
    rays[:, 2] = (1.2 + rng.random(n_rays))[:, None]
    rays[:, 3] = (1.0 + 0.1 * rng.random(n_rays))[:, None]
    phases = rng.integers(0, n_phases, size=n_rays).astype(np.int64)
    g = torch.Generator().manual_seed(seed)
    z = orc.jitter_depth(orc.depth_values(near, far, n_depth), torch.rand((n_depth,), generator=g))
    return torch.from_numpy(rays), torch.from_numpy(phases), z


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cosine(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def oracle_composite_step(sd_s, sd_d, cfg_s, cfg_d, rays, phases, z, hp, it):
    sd_s = {k: v.clone().requires_grad_(True) for k, v in sd_s.items()}
    sd_d = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
    i0 = torch.full((rays.shape[0],), I0, dtype=torch.float32)
    loss, out = orc.composite_step_loss(sd_s, sd_d, cfg_s, cfg_d, rays[:, 0, :], rays[:, 1, :], phases, i0, z, rays[:, 2, 0],
                                        rays[:, 3, 0], hp, it)
    loss.backward()
    return loss.detach(), out, {k: v.grad for k, v in sd_s.items()}, {k: v.grad for k, v in sd_d.items()}


def compare_grads(got: dict, want: dict, tol: dict, prefix=""):
    """Asserts the per-tensor bounds and returns the worst rel-L2 / cosine with the names of the tensors that set them."""
    worst = {"rel": 0.0, "cos": 1.0, "rel_name": "", "cos_name": ""}
    for k, w in want.items():
        g = got[k].detach().cpu().numpy()
        w = w.numpy()
        if np.linalg.norm(w) < 1e-30:
            continue
        r, c = rel_l2(g, w), cosine(g, w)
        assert r <= tol["grad_rel"], f"{prefix}{k}: gradient rel-L2 {r:.3e} > {tol['grad_rel']}"
        assert c >= tol["grad_cos"], f"{prefix}{k}: gradient cosine {c:.6f} < {tol['grad_cos']}"
        if r > worst["rel"]:
            worst["rel"], worst["rel_name"] = r, prefix + k
        if c < worst["cos"]:
            worst["cos"], worst["cos_name"] = c, prefix + k
    return worst


def run_composite_step_parity(n_rays=64, n_depth=40, precision="bf16", seed=0, hidden=128, n_early=4, n_freq=12, n_latent=8,
                              it=50000, fused=True, device="cuda:0", n_phases=10):
    """One composite training step on the GPU (fused path or autograd drop-in path) against the oracle."""
    from nerfca import ops
    import model_helpers as mh
    tol = TOL[precision]
    enc_dim = 3 + 6 * n_freq
    sd_s = orc.init_field_state(enc_dim, hidden, n_early, seed=seed + 1)
    sd_d = orc.init_field_state(enc_dim + n_latent, hidden, n_early, n_phases, n_latent, seed=seed + 2)
    sd_d["output_linear.0.bias"] = sd_d["output_linear.0.bias"] + 0.5
    mask, _ = orc.freq_mask(n_freq, it, 150000, 1)
    cfg = {"n_freq": n_freq, "n_hidden": n_early, "pos_enc": "free_windowed", "window": mask}
    rays, phases, z = synthetic_batch(n_rays, n_depth, seed, n_phases=n_phases)
    hp = orc.COMPOSITE_HP
    loss_o, out_o, gs_o, gd_o = oracle_composite_step(sd_s, sd_d, cfg, cfg, rays, phases, z, hp, it)

    dev = torch.device(device)
    static, temp = build_models(sd_s, sd_d, dev, precision, hidden, n_early, n_freq, n_latent, mask)
    w = orc.schedule_weights(it, hp)
    lc = ops.LossConfig(w["favor_s"], w["dyn_entro"], w["occl"], w["l1"], hp["entro_mask_thre"], hp["entro_weighted_thresh"],
                        hp["entro_use_weighting"], n_rays)
    rays_d, phases_d, z_d = rays.to(dev), phases.to(dev), z.to(dev)
    i0 = torch.full((n_rays,), I0, dtype=torch.float32, device=dev)
    if fused:
        terms, pix = ops.train_step_composite(static, temp, rays_d, phases_d, i0, z_d, "softplus", lc)
        loss = ops.loss_from_terms(terms, lc, n_rays, n_depth)
    else:
        class Args:
            favor_s_opt = None; skewness_val = 1; entro_mask_thre = hp["entro_mask_thre"]
            entro_use_weighting = hp["entro_use_weighting"]; entro_weighted_thresh = hp["entro_weighted_thresh"]
            occl_reg_perc = hp["occl_reg_perc"]
        # the drop-in path draws its own jitter from the CPU generator; feed it the same uniform numbers
        z0 = orc.depth_values(3.2, 8.8, n_depth)
        torch.manual_seed(1234)
        t_rand = torch.rand(z0.shape)
        z = orc.jitter_depth(z0, t_rand)
        loss_o, out_o, gs_o, gd_o = oracle_composite_step(sd_s, sd_d, cfg, cfg, rays, phases, z, hp, it)
        torch.manual_seed(1234)
        bps = phases_d[:, None].repeat(1, n_depth)
        pix, ss, sd_, dists, *_ = mh.obtain_train_predictions_iter(static, temp, None, None, rays_d[:, 0, :], rays_d[:, 1, :], bps, i0,
                                                                  z0.to(dev), "softplus", 32768, 0, dev)
        pixel = mh.weighted_MSELoss()(pix, rays_d[:, 2, 0], rays_d[:, 3, 0]).mean()
        t = mh.compute_losses(ss, sd_, dists, rays_d[:, 3, 0], Args)
        loss = pixel + w["favor_s"] * t[3] + w["dyn_entro"] * t[6] + w["occl"] * t[8] + w["l1"] * t[10] + w["l1"] * t[9]
        loss.backward()
        assert pix.dtype == torch.float64 and dists.dtype == torch.float64
        np.testing.assert_allclose(ss.detach().cpu().numpy(), out_o["sigma_s"].detach().numpy(), rtol=tol["sig_rtol"], atol=tol["sig_atol"])
        np.testing.assert_allclose(sd_.detach().cpu().numpy(), out_o["sigma_d"].detach().numpy(), rtol=tol["sig_rtol"], atol=tol["sig_atol"])
    torch.cuda.synchronize()
    np.testing.assert_allclose(pix.detach().cpu().numpy(), out_o["pix"].detach().numpy(), rtol=tol["pix_rtol"], atol=tol["pix_atol"])
    loss_err = abs(float(loss) - float(loss_o)) / abs(float(loss_o))
    assert loss_err <= (1e-4 if precision == "fp32" else 2e-2), f"loss {float(loss)} vs oracle {float(loss_o)}"
    gs = {k: p.grad for k, p in static.named_parameters()}
    gd = {k: p.grad for k, p in temp.named_parameters()}
    ws = compare_grads(gs, gs_o, tol, "static.")
    wd = compare_grads(gd, gd_o, tol, "dynamic.")
    wr = ws if ws["rel"] >= wd["rel"] else wd
    wc = ws if ws["cos"] <= wd["cos"] else wd
    return {"precision": precision, "loss": float(loss), "loss_oracle": float(loss_o), "loss_rel_err": loss_err,
            "pix_max_abs_err": float(np.max(np.abs(pix.detach().cpu().numpy() - out_o["pix"].detach().numpy()))),
            "grad_rel_l2_max": wr["rel"], "grad_rel_l2_worst_tensor": wr["rel_name"], "grad_cos_min": wc["cos"],
            "grad_cos_worst_tensor": wc["cos_name"]}


def run_static_step_parity(n_rays=200, n_depth=96, precision="bf16", seed=0, hidden=128, n_early=4, n_freq=12, it=50000,
                           occl_weight=1e-4, device="cuda:0"):
    """One static-only training step (run_nerf.py:205-230, BASELINE config 1 shapes) on the GPU through the fused static step
    against the oracle: one field, so the tensor-core kernels run with a single net per launch."""
    from nerfca import ops
    tol = TOL[precision]
    enc_dim = 3 + 6 * n_freq
    sd_s = orc.init_field_state(enc_dim, hidden, n_early, seed=seed + 1)
    mask, _ = orc.freq_mask(n_freq, it, 150000, 1)
    cfg = {"n_freq": n_freq, "n_hidden": n_early, "pos_enc": "free_windowed", "window": mask}
    rays, _, z = synthetic_batch(n_rays, n_depth, seed)
    sd_o = {k: v.clone().requires_grad_(True) for k, v in sd_s.items()}
    i0 = torch.full((n_rays,), I0, dtype=torch.float32)
    loss_o, out_o = orc.static_step_loss(sd_o, cfg, rays[:, 0, :], rays[:, 1, :], i0, z, rays[:, 2, 0], rays[:, 3, 0], occl_weight)
    loss_o.backward()
    dev = torch.device(device)
    static, _ = build_models(sd_s, None, dev, precision, hidden, n_early, n_freq, mask=mask)
    tv, pix = ops.train_step_static(static, rays.to(dev), i0.to(dev), z.to(dev), "softplus", occl_weight)
    loss = ops.loss_from_terms(tv, ops.LossConfig(occl_weight=occl_weight), n_rays, n_depth, static_only=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(pix.detach().cpu().numpy(), out_o["pix"].detach().numpy(), rtol=tol["pix_rtol"], atol=tol["pix_atol"])
    loss_err = abs(float(loss) - float(loss_o)) / abs(float(loss_o))
    assert loss_err <= (1e-4 if precision == "fp32" else 2e-2), f"loss {float(loss)} vs oracle {float(loss_o)}"
    w = compare_grads({k: p.grad for k, p in static.named_parameters()}, {k: v.grad for k, v in sd_o.items()}, tol, "static.")
    return {"precision": precision, "loss_rel_err": loss_err, "grad_rel_l2_max": w["rel"], "grad_cos_min": w["cos"]}

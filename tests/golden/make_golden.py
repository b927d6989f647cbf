"""Generate the golden fixtures by running the UNMODIFIED reference (kirstenmaas/NeRF-CA).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports model/CPPN.py, model/Temporal.py, train/model_helpers.py, train/proj_helpers.py and
train/data_helpers.py straight from /root/reference, drives them on small seeded synthetic
inputs and stores inputs + outputs as compressed .npz next to this file.  The fixtures pin
the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("NERFCA_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REF, "train")]
os.environ.setdefault("WANDB_MODE", "disabled")

import model_helpers as mh  # noqa: E402  (reference)
import proj_helpers as ph  # noqa: E402  (reference)
from model.CPPN import CPPN  # noqa: E402  (reference)
from model.Temporal import Temporal  # noqa: E402  (reference)

HERE = os.path.dirname(os.path.abspath(__file__))
DEV = torch.device("cpu")
torch.set_num_threads(4)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {sum(a.nbytes for a in out.values()) / 1e3:.1f} kB raw")


def static_params(h=128, n_early=4, L=12, mode="free_windowed", window_start=1, fourier=None, sigma=0.0):
    return {"num_early_layers": n_early, "num_late_layers": 0, "num_filters": h, "num_input_channels": 3,
            "num_output_channels": 1, "use_bias": True, "pos_enc": mode, "pos_enc_window_start": window_start,
            "pos_enc_basis": L, "fourier_sigma": sigma, "fourier_gaussian": fourier, "act_func": "relu",
            "device": DEV}


def temp_params(h=128, n_early=4, L=12, T=8, mode="free_windowed", window_start=1, fourier=None, sigma=0.0):
    p = static_params(h, n_early, L, mode, window_start, fourier, sigma)
    p.update({"num_input_times": 1, "use_time_latents": True, "num_time_dim": T})
    return p


def sd_arrays(prefix, model):
    return {f"{prefix}{k}": v for k, v in model.state_dict().items()}


GEOS = [
    {"DSD": 20.0, "DSO": 6.0, "nDetector": [16, 12], "dDetector": [200 * 0.01 / 16, 200 * 0.01 / 12], "offDetector": [0.0, 0.0, 0.0]},
    {"DSD": 11.98, "DSO": 7.65, "nDetector": [9, 14], "dDetector": [0.0308, 0.0291], "offDetector": [0.013, -0.027, 0.0]},
    {"DSD": 20.0, "DSO": 6.0, "nDetector": [64, 64], "dDetector": [200 * 0.01 / 64, 200 * 0.01 / 64], "offDetector": [0.0, 0.0, 0.0]},
]
VIEWS = [(-30.0, 30.0), (-30.0, -30.0), (60.0, -30.0), (60.0, 30.0), (-5.0, 40.0), (17.3, -12.9)]


def gen_geometry():
    out = {}
    for gi, geo in enumerate(GEOS):
        for vi, (th, phi) in enumerate(VIEWS):
            o, d = ph.get_ray_values_tigre(th, phi, 0, geo, DEV)
            out[f"g{gi}_v{vi}_o"] = o
            out[f"g{gi}_v{vi}_d"] = d
            out[f"g{gi}_v{vi}_pose"] = ph.source_matrix_tigre(np.array([0, 0, -geo["DSO"]]), th, phi, 0)
    save("geometry", **out)


def gen_ray_table():
    import data_helpers as dh  # reference (imports wandb at module top)
    # the reference's denormalize_image / concatenate only line up for square detectors (data_helpers.py:131,161)
    geo = {"DSD": 20.0, "DSO": 6.0, "nDetector": [10, 10], "dDetector": [0.2, 0.2], "offDetector": [0.0, 0.0, 0.0]}
    w, h = geo["nDetector"]
    rng = np.random.default_rng(3)
    frames, store = [], {}
    with tempfile.TemporaryDirectory() as td:
        for k, (th, phi) in enumerate(VIEWS[:3]):
            img = rng.random((h, w))        # file layout is [H, W]; denormalize_image transposes (data_helpers.py:131)
            img[0, 0], img[-1, -1] = 0.0, 1.0
            wimg = 1.0 + rng.random((h, w))
            fp, wp = os.path.join(td, f"i{k}.npy"), os.path.join(td, f"w{k}.npy")
            np.save(fp, img); np.save(wp, wimg)
            mm = [0.3 + 0.1 * k, 2.1 + 0.2 * k]
            frames.append({"theta": th, "phi": phi, "larm": 0, "file_path": fp, "weighted_file_path": wp,
                           "img_min_max": mm, "heart_phase": (3 * k + 1) % 10})
            store[f"img{k}"], store[f"wimg{k}"], store[f"minmax{k}"] = img, wimg, np.array(mm)
        rays, phases = dh.prepare_data_for_loader_tigre(frames, geo, w, h, 8, 1, DEV)
    save("ray_table", rays=rays, phases=phases, **store)


def gen_depth():
    out = {}
    for k, (near, far, n) in enumerate([(3.2, 8.8, 500), (0.0, 12.0, 37), (4.585786, 7.414214, 256)]):
        import data_helpers as dh
        z = dh.create_depth_values(near, far, n, DEV)
        torch.manual_seed(10 + k)
        zj = mh.randomize_depth(z, DEV)
        torch.manual_seed(10 + k)
        t = torch.rand(z.shape)
        out.update({f"c{k}_nf": np.array([near, far, n]), f"c{k}_z": z, f"c{k}_t": t, f"c{k}_zj": zj})
    save("depth", **out)


def gen_encoding():
    torch.manual_seed(1)
    x = (torch.rand(257, 3) * 2 - 1) * 2.2
    out = {"x": x}
    for name, mode in [("free", "free_windowed"), ("nerfies", "nerfies_windowed"), ("plain", "windowed")]:
        m = CPPN(static_params(h=8, n_early=0, L=12, mode=mode))
        if mode == "free_windowed":
            for tag, it in [("half", 75000), ("early", 1234), ("open", 150000)]:
                m.update_freq_mask_alpha(it, 150000)
                out[f"{name}_{tag}_mask"] = m.freq_mask_alpha
                out[f"{name}_{tag}_alpha"] = np.float64(m.windowed_alpha)
                out[f"{name}_{tag}_enc"] = m.pos_enc(x, 12, "pts")
        elif mode == "nerfies_windowed":
            m.update_windowed_alpha(40000, 150000)
            out[f"{name}_alpha"] = np.float64(m.windowed_alpha)
            out[f"{name}_enc"] = m.pos_enc(x, 12, "pts")
        else:
            out[f"{name}_enc"] = m.pos_enc(x, 12, "pts")
    g = torch.randn([3 * 6])
    m = CPPN(static_params(h=8, n_early=0, L=6, mode="fourier", fourier=g, sigma=1.7))
    out["fourier_g"] = g
    out["fourier_enc"] = m.pos_enc(x, 6, "pts")
    save("encoding", **out)


def gen_fields():
    torch.manual_seed(2)
    x = (torch.rand(300, 3) * 2 - 1) * 1.5
    ph_ = torch.randint(0, 10, (300,))
    out = {"x": x, "phases": ph_}
    for tag, h, ne, L in [("small", 32, 2, 4), ("full", 128, 4, 12)]:
        torch.manual_seed(20)
        s = CPPN(static_params(h=h, n_early=ne, L=L)); s.update_freq_mask_alpha(60000, 150000)
        t = Temporal(temp_params(h=h, n_early=ne, L=L)); t.update_freq_mask_alpha(60000, 150000)
        out.update(sd_arrays(f"{tag}_s.", s)); out.update(sd_arrays(f"{tag}_d.", t))
        out[f"{tag}_mask"] = s.freq_mask_alpha
        out[f"{tag}_raw_s"] = s(x)
        out[f"{tag}_raw_d"] = t.forward_composite(x, ph_.int())
        out[f"{tag}_raw_d_floatphase"] = t.forward_composite(x, ph_.float())
    save("fields", **out)


def _hp_namespace():
    return SimpleNamespace(favor_s_opt=None, skewness_val=1, entro_mask_thre=1e-4, entro_use_weighting=True,
                           entro_weighted_thresh=0.03, occl_reg_perc=0.2)


def gen_composite_step():
    """One run_composite.py training step (lines 262-305) on 24 rays x 40 samples, full-size nets."""
    geo = GEOS[0]
    B, N, it = 24, 40, 50000
    rng = np.random.default_rng(7)
    o, d = ph.get_ray_values_tigre(-30.0, 30.0, 0, geo, DEV)
    ids = rng.integers(0, o.shape[0] * o.shape[1], size=B)
    rays = np.zeros((B, 4, 3))
    rays[:, 0] = o.reshape(-1, 3)[ids]; rays[:, 1] = d.reshape(-1, 3)[ids]
    rays[:, 2] = (1.2 + rng.random(B))[:, None]
    rays[:, 3] = (1.0 + rng.random(B) * 0.1)[:, None]      # some above 1.03, some below
    phases = rng.integers(0, 10, size=B).astype(np.int64)
    torch.manual_seed(5)
    static = CPPN(static_params()); temp = Temporal(temp_params())
    static.update_freq_mask_alpha(it, 150000); temp.update_freq_mask_alpha(it, 150000)
    with torch.no_grad():  # make sigma_d comparable to sigma_s and push a few rays over the entropy threshold
        temp.output_linear[0].bias += 0.5
    import data_helpers as dh
    z0 = dh.create_depth_values(3.2, 8.8, N, DEV)
    batch_rays = torch.from_numpy(rays)
    batch_phases = torch.from_numpy(phases)
    bps = batch_phases[:, None].repeat(1, N)
    i0 = torch.Tensor([np.log(8.670397)] * B)
    torch.manual_seed(77)
    pix, ss, sd, dists, *_ = mh.obtain_train_predictions_iter(static, temp, None, None, batch_rays[:, 0, :], batch_rays[:, 1, :],
                                                              bps, i0, z0, "softplus", 32768, 0, DEV)
    torch.manual_seed(77)
    t_rand = torch.rand(z0.shape)
    gt, wpix = batch_rays[:, 2, 0], batch_rays[:, 3, 0]
    pixel = mh.weighted_MSELoss()(pix, gt, wpix).mean()
    terms = mh.compute_losses(ss, sd, dists, wpix, _hp_namespace())
    fw = mh.linear_param_decay(it, 1e-12, 1e-10, 100000, delay_steps=40000)
    ew = mh.linear_param_decay(it, 1e-10, 1e-8, 100000)
    ow = mh.linear_param_decay(it, 1e-8, 1e-4, 100000, delay_steps=40000)
    lw = mh.linear_param_decay(it, 1e-8, 1e-15, 100000)
    loss = pixel + fw * terms[3] + ew * terms[6] + ow * terms[8] + lw * terms[10] + lw * terms[9]
    ss.retain_grad(); sd.retain_grad()
    loss.backward()
    out = {"rays": rays, "phases": phases, "z0": z0, "t_rand": t_rand, "i0": i0, "iter": np.int64(it),
           "mask": static.freq_mask_alpha, "pix": pix, "sigma_s": ss, "sigma_d": sd, "dists": dists,
           "pixel_loss": pixel, "loss": loss, "terms": np.array([float(t) for t in terms]),
           "weights": np.array([fw, ew, ow, lw]), "dsigma_s": ss.grad, "dsigma_d": sd.grad}
    out.update(sd_arrays("s.", static)); out.update(sd_arrays("d.", temp))
    for k, p in static.named_parameters():
        out[f"gs.{k}"] = p.grad
    for k, p in temp.named_parameters():
        out[f"gd.{k}"] = p.grad
    save("composite_step", **out)


def gen_static_step():
    """One run_nerf.py training step (lines 205-230): static field, 3d.txt-style, small net."""
    geo = GEOS[1]
    B, N = 20, 33
    rng = np.random.default_rng(9)
    o, d = ph.get_ray_values_tigre(60.0, -30.0, 0, geo, DEV)
    ids = rng.integers(0, o.shape[0] * o.shape[1], size=B)
    rays = np.zeros((B, 4, 3))
    rays[:, 0] = o.reshape(-1, 3)[ids]; rays[:, 1] = d.reshape(-1, 3)[ids]
    rays[:, 2] = (1.0 + rng.random(B))[:, None]; rays[:, 3] = 1.0
    torch.manual_seed(6)
    static = CPPN(static_params(h=64, n_early=3, L=8))
    static.update_freq_mask_alpha(10, 20)
    import data_helpers as dh
    z0 = dh.create_depth_values(2.0, 13.0, N, DEV)
    br = torch.from_numpy(rays)
    i0 = torch.Tensor([np.log(8.670397)] * B)
    torch.manual_seed(78)
    pix, sig, dists = mh.obtain_train_predictions_static(static, br[:, 0, :], br[:, 1, :], i0, z0, "softplus", 256, DEV)
    torch.manual_seed(78)
    t_rand = torch.rand(z0.shape)
    pixel = mh.weighted_MSELoss()(pix, br[:, 2, 0], br[:, 3, 0]).mean()
    occl = mh.compute_occl_loss(sig, dists)
    loss = pixel + 1e-4 * occl
    loss.backward()
    out = {"rays": rays, "z0": z0, "t_rand": t_rand, "i0": i0, "mask": static.freq_mask_alpha, "pix": pix, "sigma": sig,
           "dists": dists, "pixel_loss": pixel, "occl": occl, "loss": loss}
    out.update(sd_arrays("s.", static))
    for k, p in static.named_parameters():
        out[f"gs.{k}"] = p.grad
    save("static_step", **out)


def gen_render():
    """Eval / full-frame render path of run_composite.py:346-361, 407-413 (float32 rays, float phases)."""
    geo = GEOS[0]
    N = 48
    w, h = geo["nDetector"]
    o, d = ph.get_ray_values_tigre(60.0, 30.0, 0, geo, DEV)
    to = torch.Tensor(o).reshape((-1, 3)); td = torch.Tensor(d).reshape((-1, 3))
    torch.manual_seed(8)
    static = CPPN(static_params(h=64, n_early=2, L=10)); temp = Temporal(temp_params(h=64, n_early=2, L=10))
    static.update_freq_mask_alpha(150000, 150000); temp.update_freq_mask_alpha(150000, 150000)
    static.eval(); temp.eval()
    import data_helpers as dh
    z0 = dh.create_depth_values(3.2, 8.8, N, DEV)
    torch.manual_seed(79)
    z = mh.randomize_depth(z0, DEV)
    phase = torch.full((w * h, 1), 4.0)
    i0 = torch.Tensor([np.log(8.670397)] * (w * h))
    with torch.no_grad():
        q = to[..., None, :] + td[..., None, :] * z[..., :, None]
        q = q.reshape((-1, 3)).float()
        bp = phase.repeat(N, 1).flatten()
        rs, rd = mh.get_predictions_composite(static, temp, q, bp, 4096)
        rs = rs.reshape(w * h, N, 1); rd = rd.reshape(w * h, N, 1)
        pix, ss, sd, dists = mh.render_volume_density_composite(rs, rd, i0, td, z, "softplus")
        pix_d, _, _ = mh.render_volume_density(rd, i0, td, z, "softplus")
        pix_s, _, _ = mh.render_volume_density(rs, i0, td, z, "softplus")
    out = {"theta_phi": np.array([60.0, 30.0]), "z": z, "phase": np.float32(4.0), "points": q, "pix": pix, "pix_static": pix_s,
           "pix_dynamic": pix_d, "sigma_s": ss, "sigma_d": sd, "dists": dists, "origins": to, "dirs": td}
    out.update(sd_arrays("s.", static)); out.update(sd_arrays("d.", temp))
    save("render", **out)


def gen_activations():
    torch.manual_seed(11)
    raw_s = torch.randn(6, 10, 1) * 6; raw_d = torch.randn(6, 10, 1) * 6
    raw_s[0, 0, 0] = 25.0; raw_d[0, 1, 0] = -30.0
    z = torch.sort(torch.rand(10) * 5 + 3).values
    i0 = torch.full((6,), float(np.log(8.670397)))
    out = {"raw_s": raw_s, "raw_d": raw_d, "z": z, "i0": i0}
    for act in ["softplus", "clamp", "Softplus"]:   # the last one selects Sigmoid
        for tag, dt in [("f32", torch.float32), ("f64", torch.float64)]:
            dirs = torch.zeros(6, 3, dtype=dt)
            pix, ss, sd, dists = mh.render_volume_density_composite(raw_s, raw_d, i0, dirs, z, act)
            p1, s1, _ = mh.render_volume_density(raw_s, i0, dirs, z, act)
            out.update({f"{act}_{tag}_pix": pix, f"{act}_{tag}_ss": ss, f"{act}_{tag}_sd": sd, f"{act}_{tag}_dists": dists,
                        f"{act}_{tag}_pix1": p1, f"{act}_{tag}_sig1": s1})
    save("activations", **out)


def gen_fine_pass():
    """obtain_train_predictions_iter with the hierarchical fine pass on (model_helpers.py:131-158, sample_pdf :162-187):
    20 rays x 24 coarse + 16 fine samples, two pairs of small nets; the sample_pdf draw comes from the CPU generator right
    after randomize_depth's draw, so one torch.manual_seed reproduces both."""
    geo = GEOS[0]
    B, N, NF = 20, 24, 16
    rng = np.random.default_rng(17)
    o, d = ph.get_ray_values_tigre(60.0, 30.0, 0, geo, DEV)
    ids = rng.integers(0, o.shape[0] * o.shape[1], size=B)
    rays = np.zeros((B, 2, 3))
    rays[:, 0] = o.reshape(-1, 3)[ids]; rays[:, 1] = d.reshape(-1, 3)[ids]
    phases = rng.integers(0, 10, size=B).astype(np.int64)
    torch.manual_seed(21)
    nets = [CPPN(static_params(h=64, n_early=2, L=6)), Temporal(temp_params(h=64, n_early=2, L=6)),
            CPPN(static_params(h=64, n_early=2, L=6)), Temporal(temp_params(h=64, n_early=2, L=6))]
    for m in nets:
        m.update_freq_mask_alpha(150000, 150000)
    with torch.no_grad():
        nets[1].output_linear[0].bias += 0.5; nets[3].output_linear[0].bias += 0.5
    import data_helpers as dh
    z0 = dh.create_depth_values(3.2, 8.8, N, DEV)
    batch_rays = torch.from_numpy(rays)
    bps = torch.from_numpy(phases)[:, None].repeat(1, N)
    i0 = torch.Tensor([np.log(8.670397)] * B)
    torch.manual_seed(99)
    outs = mh.obtain_train_predictions_iter(nets[0], nets[1], nets[2], nets[3], batch_rays[:, 0, :], batch_rays[:, 1, :], bps, i0, z0,
                                            "softplus", 32768, NF, DEV)
    torch.manual_seed(99)
    t_rand = torch.rand(z0.shape)
    u = torch.rand([B, NF])
    out = {"rays": rays, "phases": phases, "z0": z0, "i0": i0, "t_rand": t_rand, "u": u, "n_fine": np.int64(NF)}
    for k, v in zip(["pix_c", "ss_c", "sd_c", "dists_c", "pix_f", "ss_f", "sd_f", "dists_f"], outs):
        out[k] = v
    for tag, m in zip(["sc.", "dc.", "sf.", "df."], nets):
        out.update(sd_arrays(tag, m))
    save("fine_pass", **out)


if __name__ == "__main__":
    gen_geometry(); gen_ray_table(); gen_depth(); gen_encoding(); gen_fields()
    gen_composite_step(); gen_static_step(); gen_render(); gen_activations(); gen_fine_pass()

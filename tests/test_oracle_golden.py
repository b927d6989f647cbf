"""Pin the CPU oracle against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

from conftest import state_dict_from
from oracle import nerfca_oracle as orc

GEOS = [
    {"DSD": 20.0, "DSO": 6.0, "nDetector": [16, 12], "dDetector": [200 * 0.01 / 16, 200 * 0.01 / 12], "offDetector": [0.0, 0.0, 0.0]},
    {"DSD": 11.98, "DSO": 7.65, "nDetector": [9, 14], "dDetector": [0.0308, 0.0291], "offDetector": [0.013, -0.027, 0.0]},
    {"DSD": 20.0, "DSO": 6.0, "nDetector": [64, 64], "dDetector": [200 * 0.01 / 64, 200 * 0.01 / 64], "offDetector": [0.0, 0.0, 0.0]},
]
VIEWS = [(-30.0, 30.0), (-30.0, -30.0), (60.0, -30.0), (60.0, 30.0), (-5.0, 40.0), (17.3, -12.9)]


def cfg(L, n_hidden, mask, mode="free_windowed"):
    return {"n_freq": L, "n_hidden": n_hidden, "pos_enc": mode, "window": None if mask is None else torch.as_tensor(mask)}


def test_geometry_bit_exact(golden):
    g = golden("geometry")
    for gi, geo in enumerate(GEOS):
        for vi, (th, phi) in enumerate(VIEWS):
            assert np.array_equal(orc.pose_tigre(th, phi, geo["DSO"]), g[f"g{gi}_v{vi}_pose"])
            o, d = orc.rays_tigre(th, phi, geo)
            assert o.dtype == np.float32 and np.array_equal(o, g[f"g{gi}_v{vi}_o"])
            assert np.array_equal(d, g[f"g{gi}_v{vi}_d"]), (gi, vi)


def test_ray_table_bit_exact(golden):
    g = golden("ray_table")
    geo = {"DSD": 20.0, "DSO": 6.0, "nDetector": [10, 10], "dDetector": [0.2, 0.2], "offDetector": [0.0, 0.0, 0.0]}
    frames = []
    for k, (th, phi) in enumerate(VIEWS[:3]):
        mm = g[f"minmax{k}"]
        img = g[f"img{k}"].reshape(10, 10).T * (mm[1] - mm[0]) + mm[0]       # data_helpers.py:129-139
        frames.append({"theta": th, "phi": phi, "heart_phase": (3 * k + 1) % 10, "image": img,
                       "weight": g[f"wimg{k}"].reshape(10, 10).T})
    rays, phases = orc.build_ray_table(frames, geo, 1.0)
    assert rays.dtype == np.float64 and np.array_equal(rays, g["rays"])
    assert phases.dtype == np.int64 and np.array_equal(phases, g["phases"])


def test_depth_bit_exact(golden):
    g = golden("depth")
    for k in range(3):
        near, far, n = g[f"c{k}_nf"]
        z = orc.depth_values(float(near), float(far), int(n))
        assert np.array_equal(z.numpy(), g[f"c{k}_z"])
        zj = orc.jitter_depth(z, torch.from_numpy(g[f"c{k}_t"]))
        assert np.array_equal(zj.numpy(), g[f"c{k}_zj"])


def test_encoding_bit_exact(golden):
    g = golden("encoding")
    x = torch.from_numpy(g["x"])
    for tag, it in [("half", 75000), ("early", 1234), ("open", 150000)]:
        mask, alpha = orc.freq_mask(12, it, 150000, 1)
        assert np.array_equal(mask.numpy(), g[f"free_{tag}_mask"]) and alpha == float(g[f"free_{tag}_alpha"])
        assert np.array_equal(orc.pos_enc(x, 12, "free_windowed", mask).numpy(), g[f"free_{tag}_enc"])
    alpha = (12 * 40000) / 150000
    assert alpha == float(g["nerfies_alpha"])
    assert np.array_equal(orc.pos_enc(x, 12, "nerfies_windowed", orc.nerfies_window(12, alpha)).numpy(), g["nerfies_enc"])
    assert np.array_equal(orc.pos_enc(x, 12, "windowed").numpy(), g["plain_enc"])
    coeff = torch.from_numpy(g["fourier_g"]) * 1.7
    assert np.array_equal(orc.pos_enc(x, 6, "fourier", fourier_coeff=coeff).numpy(), g["fourier_enc"])


def test_fields_bit_exact(golden):
    g = golden("fields")
    x, ph = torch.from_numpy(g["x"]), torch.from_numpy(g["phases"])
    for tag, ne, L in [("small", 2, 4), ("full", 4, 12)]:
        c = cfg(L, ne, g[f"{tag}_mask"])
        rs = orc.static_field(x, state_dict_from(g, f"{tag}_s."), c)
        rd = orc.dynamic_field(x, ph, state_dict_from(g, f"{tag}_d."), c)
        # fp32 sgemm blocking depends on the BLAS thread count, so the MLP outputs are pinned to a few ulp, not bits
        np.testing.assert_allclose(rs.numpy(), g[f"{tag}_raw_s"], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(rd.numpy(), g[f"{tag}_raw_d"], rtol=2e-5, atol=1e-7)
        assert np.array_equal(g[f"{tag}_raw_d"], g[f"{tag}_raw_d_floatphase"])


def test_activations_and_integral(golden):
    g = golden("activations")
    rs, rd, z, i0 = (torch.from_numpy(g[k]) for k in ("raw_s", "raw_d", "z", "i0"))
    for act in ["softplus", "clamp", "Softplus"]:
        for tag, dt in [("f32", torch.float32), ("f64", torch.float64)]:
            pix, ss, sd, d = orc.integrate_composite(rs, rd, i0, dt, z, act)
            assert pix.dtype == dt and d.dtype == dt
            for got, key in [(pix, "pix"), (ss, "ss"), (sd, "sd"), (d, "dists")]:
                assert np.array_equal(got.numpy(), g[f"{act}_{tag}_{key}"]), (act, tag, key)
            p1, s1, _ = orc.integrate_single(rs, i0, dt, z, act)
            assert np.array_equal(p1.numpy(), g[f"{act}_{tag}_pix1"]) and np.array_equal(s1.numpy(), g[f"{act}_{tag}_sig1"])


def _composite_inputs(g):
    rays = torch.from_numpy(g["rays"])
    z = orc.jitter_depth(torch.from_numpy(g["z0"]), torch.from_numpy(g["t_rand"]))
    return rays, z


def test_composite_step(golden):
    g = golden("composite_step")
    rays, z = _composite_inputs(g)
    sd_s = {k: v.requires_grad_(True) for k, v in state_dict_from(g, "s.").items()}
    sd_d = {k: v.requires_grad_(True) for k, v in state_dict_from(g, "d.").items()}
    c = cfg(12, 4, g["mask"])
    it = int(g["iter"])
    loss, out = orc.composite_step_loss(sd_s, sd_d, c, c, rays[:, 0, :], rays[:, 1, :], torch.from_numpy(g["phases"]),
                                        torch.from_numpy(g["i0"]), z, rays[:, 2, 0], rays[:, 3, 0], orc.COMPOSITE_HP, it)
    out["sigma_s"].retain_grad(); out["sigma_d"].retain_grad()
    loss.backward()
    assert out["pix"].dtype == torch.float64 and out["dists"].dtype == torch.float64
    assert np.array_equal(out["dists"].numpy(), g["dists"])
    for key in ("pix", "sigma_s", "sigma_d"):
        np.testing.assert_allclose(out[key].detach().numpy(), g[key], rtol=2e-5, atol=1e-9, err_msg=key)
    w = orc.schedule_weights(it, orc.COMPOSITE_HP)
    assert np.array_equal(np.array([w["favor_s"], w["dyn_entro"], w["occl"], w["l1"]]), g["weights"])
    names = ["blendw", "sigma_s_max", "sigma_d_max", "favor_s", "s_entropy", "s_entropy_sum", "d_entropy", "d_entropy_sum",
             "d_occl", "s_l1", "s_l2"]
    np.testing.assert_allclose(np.array([float(out[n]) for n in names]), g["terms"], rtol=2e-5)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5)
    np.testing.assert_allclose(float(out["pixel"]), float(g["pixel_loss"]), rtol=1e-5)
    np.testing.assert_allclose(out["sigma_s"].grad.numpy(), g["dsigma_s"], rtol=1e-4, atol=1e-14)
    np.testing.assert_allclose(out["sigma_d"].grad.numpy(), g["dsigma_d"], rtol=1e-4, atol=1e-14)
    for k, v in sd_s.items():
        np.testing.assert_allclose(v.grad.numpy(), g[f"gs.{k}"], rtol=2e-4, atol=1e-9, err_msg=k)
    for k, v in sd_d.items():
        np.testing.assert_allclose(v.grad.numpy(), g[f"gd.{k}"], rtol=2e-4, atol=1e-9, err_msg=k)
    # closed-form dL/dsigma (what the fused backward kernel implements) against the reference autograd
    gs, gd = orc.dloss_dsigma(out["sigma_s"].detach(), out["sigma_d"].detach(), out["dists"], out["pix"].detach(),
                              rays[:, 2, 0], rays[:, 3, 0], orc.COMPOSITE_HP, it)
    np.testing.assert_allclose(gs.numpy(), g["dsigma_s"], rtol=1e-4, atol=1e-14)
    np.testing.assert_allclose(gd.numpy(), g["dsigma_d"], rtol=1e-4, atol=1e-14)


def test_static_step(golden):
    g = golden("static_step")
    rays, z = _composite_inputs(g)
    sd_s = {k: v.requires_grad_(True) for k, v in state_dict_from(g, "s.").items()}
    loss, out = orc.static_step_loss(sd_s, cfg(8, 3, g["mask"]), rays[:, 0, :], rays[:, 1, :], torch.from_numpy(g["i0"]), z,
                                     rays[:, 2, 0], rays[:, 3, 0], 1e-4, chunk=256)
    loss.backward()
    assert np.array_equal(out["dists"].numpy(), g["dists"])
    for key in ("pix", "sigma"):
        np.testing.assert_allclose(out[key].detach().numpy(), g[key], rtol=2e-5, atol=1e-9, err_msg=key)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5)
    np.testing.assert_allclose(float(out["occl"]), float(g["occl"]), rtol=1e-5)
    for k, v in sd_s.items():
        np.testing.assert_allclose(v.grad.numpy(), g[f"gs.{k}"], rtol=2e-4, atol=1e-9, err_msg=k)


def test_render_path(golden):
    g = golden("render")
    geo = GEOS[0]
    o, d = orc.rays_tigre(60.0, 30.0, geo)
    to, td = torch.from_numpy(o).reshape(-1, 3), torch.from_numpy(d).reshape(-1, 3)
    assert np.array_equal(to.numpy(), g["origins"]) and np.array_equal(td.numpy(), g["dirs"])
    z = torch.from_numpy(g["z"])
    pts = orc.sample_points(to, td, z)
    assert np.array_equal(pts.numpy(), g["points"])
    c = cfg(10, 2, np.ones(10, dtype=np.float32))
    n = z.shape[0]
    ph = torch.full((pts.shape[0],), float(g["phase"]))
    with torch.no_grad():
        rs = orc.chunked(lambda p: orc.static_field(p, state_dict_from(g, "s."), c), 4096, pts).reshape(-1, n, 1)
        rd = orc.chunked(lambda p, t: orc.dynamic_field(p, t, state_dict_from(g, "d."), c), 4096, pts, ph).reshape(-1, n, 1)
        i0 = torch.full((rs.shape[0],), float(np.float32(np.log(8.670397))))
        pix, ss, sd, dists = orc.integrate_composite(rs, rd, i0, torch.float32, z)
        pix_d, _, _ = orc.integrate_single(rd, i0, torch.float32, z)
        pix_s, _, _ = orc.integrate_single(rs, i0, torch.float32, z)
    assert np.array_equal(dists.numpy(), g["dists"])
    for got, key in [(pix, "pix"), (ss, "sigma_s"), (sd, "sigma_d"), (pix_s, "pix_static"), (pix_d, "pix_dynamic")]:
        np.testing.assert_allclose(got.numpy(), g[key], rtol=2e-5, atol=1e-8, err_msg=key)


def test_fine_pass_matches_reference(golden):
    """N3: the oracle's restatement of the hierarchical fine pass (sample_pdf + per-ray sorted depths + the ray-0 dists quirk)
    against the reference's own obtain_train_predictions_iter with depth_samples_per_ray_fine = 16."""
    g = golden("fine_pass")
    rays, phases = torch.from_numpy(g["rays"]), torch.from_numpy(g["phases"])
    z = orc.jitter_depth(torch.from_numpy(g["z0"]), torch.from_numpy(g["t_rand"]))
    cfg = {"n_freq": 6, "n_hidden": 2, "pos_enc": "free_windowed", "window": torch.ones(6)}
    sd = {t: {k[len(t):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(t)} for t in ("sc.", "dc.", "sf.", "df.")}
    i0 = torch.from_numpy(g["i0"])
    pix_c, ss_c, sd_c, d_c = orc.composite_forward(sd["sc."], sd["dc."], cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z)
    np.testing.assert_allclose(pix_c.numpy(), g["pix_c"], rtol=1e-6)
    pix_f, ss_f, sd_f, d_f, z_fine = orc.fine_pass(sd["sf."], sd["df."], cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z, ss_c, sd_c,
                                                   torch.from_numpy(g["u"]))
    assert z_fine.shape == (20, 40) and bool((z_fine[:, 1:] >= z_fine[:, :-1]).all())
    assert np.array_equal(d_f.numpy(), g["dists_f"])
    np.testing.assert_allclose(pix_f.numpy(), g["pix_f"], rtol=1e-6)
    np.testing.assert_allclose(ss_f.numpy(), g["ss_f"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(sd_f.numpy(), g["sd_f"], rtol=1e-5, atol=1e-12)

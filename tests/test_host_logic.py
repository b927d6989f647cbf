"""CPU-side checks: the C-ABI library loads and exports every symbol include/nerfca.h declares, and the host logic of
the drop-in modules (schedules, pose algebra, parameter layout, seeded initialisation) matches the reference fixtures."""
import os
import re
import types

import numpy as np
import pytest
import torch

from conftest import ROOT, state_dict_from
from oracle import nerfca_oracle as orc
import parity


def test_library_exports_every_declared_symbol():
    from nerfca import _lib
    header = open(os.path.join(ROOT, "include", "nerfca.h")).read()
    declared = set(re.findall(r"NERFCA_API\s+[\w\s\*]+?\b(nerfca_\w+)\s*\(", header))
    assert declared and declared == set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nerfca_abi_version() == _lib.ABI_VERSION
    out = os.popen(f"nm -D --defined-only {_lib.LIB_PATH}").read()
    for name in declared:
        assert f" T {name}" in out


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    from nerfca import _lib
    assert C.sizeof(_lib.FieldStruct) == 6 * 4 + 3 * 8 + 2 * 8 * _lib.MAX_LAYERS
    assert C.sizeof(_lib.FieldGradsStruct) == 8 + 2 * 8 * _lib.MAX_LAYERS
    assert C.sizeof(_lib.SamplesStruct) == 8 + 8 + 4 + 4 + 8 + 8 + 4 + 4 + 8 + 8 + 8
    assert C.sizeof(_lib.LossCfgStruct) == 6 * 8 + 2 * 4
    assert C.sizeof(_lib.PeersStruct) == 2 * 4 + 3 * 8          # nerfca_peers_t
    assert C.sizeof(_lib.AdamStepStruct) == 6 * 8               # nerfca_adam_step_t
    assert C.sizeof(_lib.RepackStruct) == 3 * 8                 # nerfca_repack_t


def test_cpu_tensors_fail_loudly():
    from model.CPPN import CPPN
    m = CPPN(parity.static_definition("cpu", hidden=16, n_early=1, n_freq=2))
    m.update_freq_mask_alpha(1, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 3))


def test_pose_matches_reference(golden):
    import proj_helpers as ph
    g = golden("geometry")
    geos_dso = [6.0, 7.65, 6.0]
    views = [(-30.0, 30.0), (-30.0, -30.0), (60.0, -30.0), (60.0, 30.0), (-5.0, 40.0), (17.3, -12.9)]
    for gi, dso in enumerate(geos_dso):
        for vi, (th, phi) in enumerate(views):
            pose = ph.source_matrix_tigre(np.array([0, 0, -dso]), th, phi, 0)
            assert np.array_equal(pose, g[f"g{gi}_v{vi}_pose"])
    assert np.allclose(ph.get_rotation(10, 20, 30) @ np.linalg.inv(ph.get_rotation(10, 20, 30)), np.eye(4))
    assert np.array_equal(ph.y_rotation_matrix(0.3)[[0, 0, 2, 2], [0, 2, 0, 2]],
                          np.array([np.cos(0.3), np.sin(0.3), -np.sin(0.3), np.cos(0.3)]))


def test_frequency_schedules_match_reference(golden):
    from model.CPPN import CPPN
    from model.Temporal import Temporal
    g = golden("encoding")
    for cls, d in [(CPPN, parity.static_definition("cpu", 8, 0, 12)), (Temporal, parity.temporal_definition("cpu", 8, 0, 12))]:
        m = cls(d)
        for tag, it in [("half", 75000), ("early", 1234), ("open", 150000)]:
            m.update_freq_mask_alpha(it, 150000)
            assert np.array_equal(m.freq_mask_alpha.numpy(), g[f"free_{tag}_mask"])
            assert float(m.windowed_alpha) == float(g[f"free_{tag}_alpha"])
        m.update_windowed_alpha(40000, 150000)
        assert m.windowed_alpha == float(g["nerfies_alpha"])


def test_state_dict_keys_and_seeded_init_match_reference(golden):
    from model.CPPN import CPPN
    from model.Temporal import Temporal
    g = golden("fields")
    for tag, h, ne, L in [("small", 32, 2, 4), ("full", 128, 4, 12)]:
        torch.manual_seed(20)
        s = CPPN(parity.static_definition("cpu", h, ne, L))
        t = Temporal(parity.temporal_definition("cpu", h, ne, L))
        ref_s, ref_d = state_dict_from(g, f"{tag}_s."), state_dict_from(g, f"{tag}_d.")
        assert list(s.state_dict().keys()) == list(ref_s.keys())
        assert list(t.state_dict().keys()) == list(ref_d.keys())
        for k, v in s.state_dict().items():
            assert torch.equal(v, ref_s[k]), k
        for k, v in t.state_dict().items():
            assert torch.equal(v, ref_d[k]), k
        assert [k for k, _ in t.named_parameters()][0] == "time_latents"
        assert s.input_features == 3 + 6 * L and t.input_features == 3 + 6 * L + 8


def test_schedules_and_losses_host_side():
    import model_helpers as mh
    for it in [0, 39999, 40000, 50000, 100000, 250000]:
        assert mh.linear_param_decay(it, 1e-12, 1e-10, 100000, delay_steps=40000) == orc.linear_decay(it, 1e-12, 1e-10, 100000, 40000)
        assert mh.linear_param_decay(it, 1e-8, 1e-15, 100000) == orc.linear_decay(it, 1e-8, 1e-15, 100000)
    assert mh.exp_param_decay(5, 1.0, 0.01, 11) == pytest.approx(1.0 * 0.01 ** 0.5)
    assert mh.exp_param_decay(20, 1.0, 0.01, 11) == 0.01 and mh.exp_param_decay(1, 1.0, 0.01, 11, delay_steps=3) == 0
    # the torch-expression regularisers agree with the oracle on CPU tensors (they are device-agnostic)
    torch.manual_seed(0)
    ss, sd = torch.rand(6, 9) * 0.02, torch.rand(6, 9) * 0.02
    sd[0] = 0.0
    d = torch.cat([torch.rand(8, dtype=torch.float64) * 0.1, torch.tensor([1e-10], dtype=torch.float64)])
    w = 1 + 0.1 * torch.rand(6, dtype=torch.float64)
    args = types.SimpleNamespace(favor_s_opt=None, skewness_val=1, entro_mask_thre=1e-4, entro_use_weighting=True,
                                 entro_weighted_thresh=0.03, occl_reg_perc=0.2)
    got = mh.compute_losses(ss, sd, d, w, args)
    want = orc.composite_losses(ss, sd, d, w, orc.COMPOSITE_HP)
    for a, b in zip(got, want):
        assert float(a) == pytest.approx(float(b), rel=1e-12)
    assert torch.equal(mh.weighted_MSELoss()(ss, sd, ss), ((ss - sd) ** 2) * ss)
    assert [len(c) for c in mh.get_minibatches_time(torch.zeros(10, 3), torch.zeros(10), 4)] == [2, 2, 2]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exactly one JSON line on stdout carrying the
    contract's keys, for the same metric / unit / workload as the GPU arm."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training rays/sec (fwd+bwd)" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]

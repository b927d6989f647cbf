#!/usr/bin/env python
"""bench.py -- NeRF-CA composite training throughput (rays/s, forward + losses + backward + Adam) on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N > 1 is launched with torch.distributed.run, one rank per GPU (NCCL); rays are sharded across ranks
  (weak scaling: every rank runs the config's 1024-ray batch), the only collective is the gradient all-reduce.

One JSON line is printed by rank 0.  `value` = rays/s with the ray batches already resident in HBM; `e2e` = the same
step driven through the public host API with pinned HOST buffers (H2D of the batch rows + D2H of the loss terms in
the timed region); `roofline` = the dominant kernel against the measured bf16 tensor peak; `cpu_baseline` = the CPU
oracle port (oracle/nerfca_oracle.py, CPU torch, all host threads) timed on a bounded sample of the same workload.

`--impl reference` times the reference's CPU algorithm (the oracle port -- the reference is pure Python/PyTorch and
/root/reference does not exist on the GPU box) on the box's host cores for the same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "nerf-ca_b200")
for _p in (ROOT, PKG, os.path.join(PKG, "train"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ---- workload: BASELINE.json configs[1] = train/composite.txt (default); --config 3 = configs[2] -----------------------------
N_RAYS = 1024          # composite.txt:40  rays per step (per GPU: weak scaling)
N_DEPTH = 500          # composite.txt:25
N_FREQ, HIDDEN, N_EARLY, N_LATENT, N_PHASES = 12, 128, 4, 8, 10
DET = 200              # 200 x 200 detector, 4 views x 10 cardiac phases = 1.6 M rays
VIEWS = [(-30.0, 30.0), (-30.0, -30.0), (60.0, -30.0), (60.0, 30.0)]
VIEWS_8 = VIEWS + [(-5.0, 40.0), (-5.0, -40.0), (90.0, 0.0), (-30.0, 0.0)]      # preprocess/general_helpers.py:94,132
GEO = {"DSD": 20.0, "DSO": 6.0, "nDetector": [DET, DET], "dDetector": [200 * 0.01 / DET] * 2, "offDetector": [0.0, 0.0, 0.0]}
NEAR, FAR = 3.2, 8.8
I0 = float(np.log(8.670397))
ITER = 50000           # schedule point: all regularisers active, 5 of 12 bands masked by the frequency window
FLOP_PER_SAMPLE = 870912      # fwd + bwd, 1 MAC = 2 FLOP, unpadded K, no recompute credit
FLOP_PER_SAMPLE_FWD = 303104
LR, LR_END_FACTOR, LR_DECAY_STEPS = 1e-3, 0.01, 150000   # composite.txt:33-35


def flops_per_sample(n_freq, hidden, n_early, n_latent):
    """SURVEY 8(d) accounting (1 MAC = 2 FLOP, unpadded K, no recompute credit) for the static + dynamic pair: forward, and forward +
    backward (weight gradients of every layer, input gradients of every layer but the first, + the latent columns of the first)."""
    d_s = 3 + 6 * n_freq
    fwd = sum(d * hidden + n_early * hidden * hidden + hidden for d in (d_s, d_s + n_latent))
    dgrad = 2 * (n_early * hidden * hidden + hidden) + hidden * n_latent
    return 2 * fwd, 2 * (2 * fwd + dgrad)


def select_config(cfg: int):
    """--config 3: BASELINE.json configs[2] -- 512^2 projections, 8 views x 30 cardiac phases (62.9 M rays, 6.5 GB ray table
    resident in HBM), 30 time latents; same nets and per-step sizes as config 2.
    --config 5: configs[4], the widened stress config -- hidden width 256, 16 Fourier bands, 256 samples per ray (config 2's phantom);
    outside the fused kernels' tile shapes, so the bf16 path is the layer-wise tcgen05 GEMM path (csrc/mlp_wide.cu)."""
    global DET, VIEWS, N_PHASES, GEO, N_DEPTH, N_FREQ, HIDDEN, FLOP_PER_SAMPLE, FLOP_PER_SAMPLE_FWD
    if cfg == 3:
        DET, VIEWS, N_PHASES = 512, VIEWS_8, 30
        GEO = dict(GEO, nDetector=[DET, DET], dDetector=[200 * 0.01 / DET] * 2)
    if cfg == 5:
        N_DEPTH, N_FREQ, HIDDEN = 256, 16, 256
    FLOP_PER_SAMPLE_FWD, FLOP_PER_SAMPLE = flops_per_sample(N_FREQ, HIDDEN, N_EARLY, N_LATENT)


# The driver reads ONE JSON line from stdout: everything else that libraries print there (e.g. NCCL's version banner) is sent to
# stderr by pointing fd 1 at fd 2 for the whole run; emit() writes the line to the saved original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


def measured_traffic(family: str):
    """DRAM bytes per launch of a kernel family from the newest committed ncu capture (profiles/*/traffic.json), or None."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
            return {"dram_bytes_per_launch": t["per_launch"][family]["dram_bytes"], "source": os.path.relpath(path, ROOT)}
        except (OSError, KeyError, ValueError):
            continue
    return None


# ---- synthetic phantom: a chain of Gaussian blobs (closed-form line integrals) ---------------------------------------

def phantom_blobs(phase: int):
    """Background blob + a branching 'vessel' of small blobs whose centreline moves with the cardiac phase."""
    rng = np.random.default_rng(0)
    t = np.linspace(0, 1, 24)
    wob = 0.09 * np.sin(2 * np.pi * (phase / N_PHASES) + 3 * t)
    c1 = np.stack([-0.5 + 1.0 * t + wob, 0.35 * np.sin(3 * t) + wob, -0.4 + 0.8 * t], 1)
    c2 = np.stack([0.1 + 0.5 * t, -0.2 - 0.5 * t + wob, 0.1 + 0.3 * t - wob], 1)
    centres = np.concatenate([np.zeros((1, 3)), c1, c2[4:]], 0)
    sig = np.concatenate([[0.8], np.full(len(c1) + len(c2) - 4, 0.045)])
    amp = np.concatenate([[0.02], 0.05 + 0.1 * rng.random(len(sig) - 1)])
    return centres, sig, amp


def project_blobs(o: torch.Tensor, d: torch.Tensor, phase: int) -> torch.Tensor:
    """log-intensity pixel  log I0 - sum_blobs A s sqrt(2 pi) / |d| exp(-dist^2 / 2 s^2)  (float64, on o's device)."""
    c, s, a = phantom_blobs(phase)
    c = torch.as_tensor(c, dtype=torch.float64, device=o.device)
    s = torch.as_tensor(s, dtype=torch.float64, device=o.device)
    a = torch.as_tensor(a, dtype=torch.float64, device=o.device)
    o, d = o.double(), d.double()
    dn = d.norm(dim=-1, keepdim=True)
    u = d / dn
    oc = c[None, :, :] - o[:, None, :]
    along = (oc * u[:, None, :]).sum(-1)
    dist2 = (oc * oc).sum(-1) - along * along
    integ = (a * s * np.sqrt(2 * np.pi))[None, :] * torch.exp(-dist2 / (2 * s * s)[None, :])
    return I0 - integ.sum(-1)


def build_ray_table(device):
    """rays_train [R,4,3] float64 (origin, direction, pixel x3, weight x3) + phases_train [R] int64 in the reference's
    layout (train/data_helpers.py:141-165), ray id = frame * W * H + i * H + j.  Built on `device`."""
    import proj_helpers as ph
    rows, phases = [], []
    for (theta, phi) in VIEWS:
        o, d = ph.ray_values_tigre_device(theta, phi, 0, GEO, device)
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
        pix = torch.stack([project_blobs(o, d, p) for p in range(N_PHASES)], 0)          # [phases, W*H]
        var = pix.var(dim=0)
        wmap = 1.0 + var / var.max().clamp_min(1e-30)                                     # [1,2] temporal-variance weights
        for p in range(N_PHASES):
            r = torch.empty((o.shape[0], 4, 3), dtype=torch.float64, device=device)
            r[:, 0], r[:, 1] = o.double(), d.double()
            r[:, 2], r[:, 3] = pix[p][:, None], wmap[:, None]
            rows.append(r)
            phases.append(torch.full((o.shape[0],), p, dtype=torch.int64, device=device))
    return torch.cat(rows, 0), torch.cat(phases, 0)


# ---- clocks ------------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU arm: the oracle port of the reference step ------------------------------------------------------------------------

def cpu_reference_step_fn(n_rays: int, seed: int = 0):
    """Returns (step, n_rays): one reference training step (obtain_train_predictions_iter + wMSE + compute_losses +
    backward + Adam, train/run_composite.py:283-308) on CPU torch through the oracle port."""
    from oracle import nerfca_oracle as orc
    import parity          # tests/parity.py (a foreign `tests` package may shadow `from tests import ...` on the box)
    torch.set_num_threads(os.cpu_count() or 1)
    enc = 3 + 6 * N_FREQ
    sd_s = {k: v.requires_grad_(True) for k, v in orc.init_field_state(enc, HIDDEN, N_EARLY, seed=1).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in orc.init_field_state(enc + N_LATENT, HIDDEN, N_EARLY, N_PHASES, N_LATENT, seed=2).items()}
    mask, _ = orc.freq_mask(N_FREQ, ITER, 150000, 1)
    cfg = {"n_freq": N_FREQ, "n_hidden": N_EARLY, "pos_enc": "free_windowed", "window": mask}
    opt = torch.optim.Adam(list(sd_s.values()) + list(sd_d.values()), lr=LR)
    i0 = torch.full((n_rays,), I0, dtype=torch.float32)
    state = {"k": 0}

    def step():
        rays, phases, z = parity.synthetic_batch(n_rays, N_DEPTH, seed + state["k"], N_PHASES, NEAR, FAR)
        state["k"] += 1
        loss, _ = orc.composite_step_loss(sd_s, sd_d, cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z, rays[:, 2, 0], rays[:, 3, 0],
                                          orc.COMPOSITE_HP, ITER)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)
    return step


def time_cpu_arm(n_rays: int, steps: int, warmup: int):
    step = cpu_reference_step_fn(n_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: rays per step sized from a 16-ray probe so that (steps + warmup) steps take about <= 150 s
    probe_rate, _ = time_cpu_arm(16, 1, 1)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = int(min(N_RAYS, max(16, 2 ** int(np.floor(np.log2(max(16.0, probe_rate * budget)))))))
    rate, sec = time_cpu_arm(n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "training rays/sec (fwd+bwd)", "value": rate, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(extra={"cpu_sample_rays_per_step": n}),
            "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{n} rays x {N_DEPTH} samples per step, {args.steps} steps, oracle port of the reference step "
                                       f"(CPU torch {torch.__version__}, {cores} threads)"},
            "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(extra=None):
    fused = HIDDEN == 128
    c = {"workload": f"NeRF-CA composite training step, train/composite.txt: static CPPN + dynamic Temporal, "
                     f"{N_RAYS} rays x {N_DEPTH} samples per step per GPU, {N_FREQ} bands, 2 x [in->{HIDDEN}, {N_EARLY} x {HIDDEN}->{HIDDEN}, {HIDDEN}->1], "
                     f"{DET}x{DET} detector x {len(VIEWS)} views x {N_PHASES} phases synthetic blob phantom",
         "rays_per_step_per_gpu": N_RAYS, "samples_per_ray": N_DEPTH, "hidden": HIDDEN, "n_freq": N_FREQ, "n_phases": N_PHASES,
         "step": ("ONE CUDA-graph launch per step: fields fwd (clears the loss sums) + line integral + 11 loss terms + closed-form "
                  "dL/draw + fields bwd (wgrad/dgrad/latent) + Adam/LinearLR (+ gradient clearing + bf16 re-pack; N>1: fused with the "
                  "gradient and loss-sum exchange over NVLink peer memory)") if fused else
                 ("ONE CUDA-graph launch per step; the fields run on the layer-wise tcgen05 GEMM path (one GEMM kernel per layer and "
                  "pass, bf16 activations in HBM) + line integral / losses / dL/draw kernel + Adam/LinearLR"),
         "l2": ("per-step working set (activation stash, ~1.0 GB written by the forward and read back by the backward) exceeds the 126 MB L2 and every step uses a distinct ray batch; no explicit flush")
               if fused else "per-step working set (bf16 activations of every layer, ~1.5 GB) exceeds the 126 MB L2 and every step uses a distinct ray batch; no explicit flush"}
    if extra:
        c.update(extra)
    return c


def dropin_loop_rate(dev, rays_host_tab, phases_host_tab, n_rays: int, steps: int, warmup: int):
    """rays/s of the UNMODIFIED driver's iteration (train/run_composite.py:227-308) over the drop-in modules: the same sequence of calls
    the stock script makes -- update_freq_mask_alpha, numpy ray-id draw + host fancy index + H2D, obtain_train_predictions_iter
    (two FieldFunction launches + the line integral through autograd), weighted_MSELoss, compute_losses (eager torch on the returned
    sigma tensors), loss.backward(), torch.optim.Adam.step(), LinearLR.step(), and the early-stop read of the loss (:310, a D2H
    sync every iteration).  (The stock script itself cannot run on the GPU box: /root/reference is not there, and it must not be
    copied; this restates its loop line by line against the same module surface.)"""
    import types
    import model_helpers as mh
    from model.CPPN import CPPN
    from model.Temporal import Temporal
    import parity
    torch.manual_seed(0)
    static = CPPN(parity.static_definition(dev, HIDDEN, N_EARLY, N_FREQ, precision="bf16")).to(dev)
    temp = Temporal(parity.temporal_definition(dev, HIDDEN, N_EARLY, N_FREQ, N_LATENT, precision="bf16"))
    if N_PHASES != temp.time_latents.shape[0]:
        temp.time_latents = torch.nn.Parameter(torch.rand((N_PHASES, N_LATENT)))
    temp.to(dev)
    opt = torch.optim.Adam(list(temp.parameters()) + list(static.parameters()), lr=LR)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=LR_END_FACTOR, total_iters=LR_DECAY_STEPS)
    from nerfca import trainer as tr
    hp = tr.COMPOSITE_HP
    args = types.SimpleNamespace(favor_s_opt=None, skewness_val=1, entro_mask_thre=hp["entro_mask_thre"],
                                 entro_use_weighting=hp["entro_use_weighting"], entro_weighted_thresh=hp["entro_weighted_thresh"], occl_reg_perc=0.2)
    t = torch.linspace(0., 1., N_DEPTH)
    depth_values = (NEAR * (1. - t) + FAR * t).to(dev)
    rays_np, phases_np = rays_host_tab.numpy(), phases_host_tab.numpy()
    mse = mh.weighted_MSELoss()
    rng = np.random.RandomState(0)
    i0 = torch.full((n_rays,), I0, dtype=torch.float32, device=dev)
    t0 = None
    for k in range(warmup + steps):
        if k == warmup:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        it = ITER + k
        static.update_freq_mask_alpha(it, hp["static_window_decay_steps"])
        temp.update_freq_mask_alpha(it, hp["temp_window_decay_steps"])
        ids = rng.randint(0, rays_np.shape[0], size=n_rays)
        batch_rays = torch.from_numpy(rays_np[ids]).to(dev)
        batch_phases = torch.from_numpy(phases_np[ids]).to(dev)
        phases_samples = batch_phases[:, None].repeat(1, N_DEPTH)
        d = hp["hyperparam_decay_steps"]
        fw = mh.linear_param_decay(it, hp["favor_s_weight_start"], hp["favor_s_weight_end"], d, hp["favor_s_weight_delay_steps"])
        ew = mh.linear_param_decay(it, hp["dynamic_entro_weight_start"], hp["dynamic_entro_weight_end"], d)
        ow = mh.linear_param_decay(it, hp["occl_weight_start"], hp["occl_weight_end"], d, hp["favor_s_weight_delay_steps"])
        lw = mh.linear_param_decay(it, hp["l1_weight_start"], hp["l1_weight_end"], d)
        pix, ss, sd, dists, *_ = mh.obtain_train_predictions_iter(static, temp, None, None, batch_rays[:, 0, :], batch_rays[:, 1, :], phases_samples,
                                                                 i0, depth_values, "softplus", 32768, 0, dev)
        pixel = mse(pix, batch_rays[:, 2, 0], batch_rays[:, 3, 0]).mean()
        terms = mh.compute_losses(ss, sd, dists, batch_rays[:, 3, 0], args)
        loss = pixel + fw * terms[3] + ew * terms[6] + ow * terms[8] + lw * terms[10] + lw * terms[9]
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        stop = bool(loss < 1e-12)            # the early-stop test of :310 (a D2H sync per iteration)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n_rays * steps / dt, "unit": "rays/s", "ms_per_step": dt / steps * 1e3, "loss_last_step": float(loss),
            "api": "the stock driver's iteration (run_composite.py:227-308) restated over the drop-in modules: obtain_train_predictions_iter + "
                   "compute_losses + loss.backward() + torch.optim.Adam + LinearLR, host ray-table gather + H2D + per-iteration loss read"}


def gpu_eager_step_fn(dev, n_rays: int):
    """The reference's own eager-PyTorch step on the GPU (SURVEY 8(d) last row, "the practical beat-this number"): the oracle port
    -- which times identically to the unmodified reference -- with every tensor on `dev`, true-fp32 sgemm (allow_tf32 off, torch's
    default), autograd backward and torch.optim.Adam + LinearLR, one step per call on a fresh synthetic batch resident in HBM."""
    from oracle import nerfca_oracle as orc
    import parity
    torch.backends.cuda.matmul.allow_tf32 = False
    enc = 3 + 6 * N_FREQ
    sd_s = {k: v.to(dev).requires_grad_(True) for k, v in orc.init_field_state(enc, HIDDEN, N_EARLY, seed=1).items()}
    sd_d = {k: v.to(dev).requires_grad_(True) for k, v in orc.init_field_state(enc + N_LATENT, HIDDEN, N_EARLY, N_PHASES, N_LATENT, seed=2).items()}
    mask, _ = orc.freq_mask(N_FREQ, ITER, 150000, 1)
    cfg = {"n_freq": N_FREQ, "n_hidden": N_EARLY, "pos_enc": "free_windowed", "window": mask.to(dev)}
    opt = torch.optim.Adam(list(sd_s.values()) + list(sd_d.values()), lr=LR)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=LR_END_FACTOR, total_iters=LR_DECAY_STEPS)
    i0 = torch.full((n_rays,), I0, dtype=torch.float32, device=dev)
    batches = [tuple(t.to(dev) for t in parity.synthetic_batch(n_rays, N_DEPTH, 900 + k, N_PHASES, NEAR, FAR)) for k in range(4)]
    state = {"k": 0}

    def step():
        rays, phases, z = batches[state["k"] % len(batches)]
        state["k"] += 1
        loss, _ = orc.composite_step_loss(sd_s, sd_d, cfg, cfg, rays[:, 0, :], rays[:, 1, :], phases, i0, z, rays[:, 2, 0], rays[:, 3, 0],
                                          orc.COMPOSITE_HP, ITER)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        return loss
    return step


# ---- GPU arm ------------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NERFCA_PRECISION", "bf16"))
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="BASELINE.json configs index + 1 (2: composite.txt, 3: 512^2 x 8 views x 30 phases, 5: hidden 256 / 16 bands / 256 samples)")
    ap.add_argument("--strong", type=int, default=0, help="fixed GLOBAL batch of this many rays split over the ranks (strong scaling) instead of 1024 per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    select_config(args.config)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    from nerfca import trainer as tr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"

    torch.manual_seed(0)
    np.random.seed(0)
    trainer = tr.CompositeTrainer.from_config(device=dev, precision=args.precision, n_freq=N_FREQ, hidden=HIDDEN, n_early=N_EARLY,
                                              n_latent=N_LATENT, n_phases=N_PHASES, lr=LR, lr_end_factor=LR_END_FACTOR,
                                              lr_decay_steps=LR_DECAY_STEPS, i0=I0, near=NEAR, far=FAR, n_depth=N_DEPTH,
                                              world_size=world)
    trainer.set_iteration(ITER)
    rays_tab, phases_tab = build_ray_table(dev)
    R = rays_tab.shape[0]
    trainer.attach_ray_table(rays_tab, phases_tab)
    rays_tab, phases_tab = trainer.rays_table, trainer.phases_table
    scaling = "strong" if args.strong else "weak"
    n_global = args.strong if args.strong else world * N_RAYS
    sl = tr.shard_slice(n_global, rank, world)
    B = sl.stop - sl.start

    n_total = args.steps + args.warmup
    # same host RNG stream on every rank; each rank takes its contiguous slice of the global batch
    ids_all = np.ascontiguousarray(np.random.randint(0, R, size=(n_total, n_global))[:, sl])
    ids_dev = torch.from_numpy(ids_all).to(dev)
    gen = torch.Generator().manual_seed(1234)
    t_rand = torch.rand((n_total, N_DEPTH), generator=gen)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(trainer.stream):      # the step's graph is launched on the trainer's stream: time on that stream
        # ---------------- value: batches resident in HBM ----------------
        batches = [(rays_tab[ids_dev[k]].contiguous(), phases_tab[ids_dev[k]].to(torch.int32).contiguous(), trainer.jitter(t_rand[k]))
                   for k in range(n_total)]
        torch.cuda.synchronize()
        for k in range(args.warmup):
            trainer.step_device(*batches[k], n_rays_global=n_global)
        clocks = ClockSampler(local)
        launches0 = trainer.launch_count
        barrier()
        clocks.start()
        cuprof = os.environ.get("NERFCA_CUPROF") == "1"      # ncu --profile-from-start off: list exactly the timed region's launches
        if cuprof:
            torch.cuda.profiler.start()
        e0.record()
        for k in range(args.warmup, n_total):
            trainer.step_device(*batches[k], n_rays_global=n_global)
        e1.record()
        barrier()
        if cuprof:
            torch.cuda.profiler.stop()
        clk = clocks.stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = trainer.launch_count - launches0
        ms_per_step = ms / args.steps
        value = n_global * args.steps / (ms * 1e-3)
        last_terms = trainer.last_terms.clone()
        graph_stats = trainer.graph_stats()

        # ---------------- per-kernel timing (second pass: kernels launched one by one, events inside the library) ----------------
        kt = trainer.kernel_times(lambda k: trainer.step_device(*batches[args.warmup + (k % args.steps)], n_rays_global=n_global),
                                  min(args.steps, 20))
        del batches

        # ---------------- e2e: what upstream does per iteration (run_composite.py:250-308) through the public host API: fancy-index
        # the HOST ray table with the step's ids, H2D of the rows, the step, D2H of the loss sums -- all inside the timed region ------
        rays_host_tab, phases_host_tab = rays_tab.cpu(), phases_tab.cpu()
        ids_host = [torch.from_numpy(ids_all[k]) for k in range(n_total)]
        ring = [(torch.empty((B, 4, 3), dtype=torch.float64).pin_memory(), torch.empty((B,), dtype=torch.int64).pin_memory(),
                 torch.empty((B,), dtype=torch.int32).pin_memory(), torch.empty((N_DEPTH,), dtype=torch.float32).pin_memory()) for _ in range(8)]

        def host_step(k):
            r, p64, p32, tr_ = ring[k % len(ring)]
            torch.index_select(rays_host_tab, 0, ids_host[k], out=r)         # rays_train[ids]      (run_composite.py:262)
            torch.index_select(phases_host_tab, 0, ids_host[k], out=p64)     # phases_train[ids]    (:263)
            p32.copy_(p64)                                                   # .int()               (:265)
            tr_.copy_(t_rand[k])                                             # the CPU generator's draw of randomize_depth
            return trainer.step_host_async(r, p32, tr_, n_rays_global=n_global)
        h2d = B * 96 + B * 4 + N_DEPTH * 4
        for k in range(args.warmup):
            host_step(k).loss()
        barrier()
        e0.record()
        losses, pending = [], []
        for k in range(args.warmup, n_total):
            # the loss of step k is read (host wait on that step's event only) while step k + 1 is already queued, as a driver
            # that logs one iteration late would
            pending.append(host_step(k))
            if len(pending) > 1:
                losses.append(pending.pop(0).loss())
        losses.extend(h.loss() for h in pending)
        e1.record()
        barrier()
        e2e_value = n_global * args.steps / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        d2h = trainer.d2h_bytes_per_step
        dropin = None
        if world == 1 and args.config == 2 and not args.no_dropin:
            dropin = dropin_loop_rate(dev, rays_host_tab, phases_host_tab, N_RAYS, min(args.steps, 30), 5)
        del rays_host_tab, phases_host_tab

        # ---------------- e2e, N1 path: ray table resident in HBM, only the step's ray ids + the jitter draw cross PCIe ----------------
        ids_pinned = [t.pin_memory() for t in ids_host]
        trand_pinned = [t_rand[k].pin_memory() for k in range(n_total)]
        for k in range(args.warmup):
            trainer.step_ids_async(ids_pinned[k], trand_pinned[k], n_rays_global=n_global).loss()
        barrier()
        e0.record()
        pending = []
        for k in range(args.warmup, n_total):
            pending.append(trainer.step_ids_async(ids_pinned[k], trand_pinned[k], n_rays_global=n_global))
            if len(pending) > 1:
                pending.pop(0).loss()
        for h in pending:
            h.loss()
        e1.record()
        barrier()
        e2e_ids_value = n_global * args.steps / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
        h2d_ids = B * 8 + N_DEPTH * 4

    # ---------------- replicas must be bit-identical after all of the above (the fused exchange keeps them so) ----------------
    replicas_identical = None
    if dist is not None:
        digest = torch.stack([trainer.flat_p.double().sum(), trainer.flat_p.double().abs().sum()])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        replicas_identical = all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])

    # ---------------- secondary metric (BASELINE.json): render ms/frame, config 4 (1024^2 frame, static + dynamic, no grad);
    # N > 1: detector rows sharded over the ranks (no collective, every rank renders its strip) ----------------
    render = None
    if not args.no_render:
        import proj_helpers as ph
        from nerfca import ops
        geo_r = dict(GEO, nDetector=[1024, 1024], dDetector=[200 * 0.01 / 1024] * 2)
        o_r, d_r = ph.ray_values_tigre_device(VIEWS[0][0], VIEWS[0][1], 0, geo_r, dev)
        rs = tr.shard_slice(1024, rank, world)                         # detector rows [u index] of this rank
        o_r, d_r = o_r[rs.start:rs.stop].reshape(-1, 3), d_r[rs.start:rs.stop].reshape(-1, 3)
        z_r = trainer.depth_uniform
        with torch.no_grad():
            ops.render_frame(trainer.static, trainer.temp, o_r, d_r, z_r, 3, I0)        # warm-up frame
            barrier()
            e0.record()
            n_frames = 3
            for ph_id in range(n_frames):
                pix_r, _, _ = ops.render_frame(trainer.static, trainer.temp, o_r, d_r, z_r, ph_id, I0)
            e1.record()
            barrier()
        ms_frame = max_over_ranks(e0.elapsed_time(e1)) / n_frames
        flop_frame = 1024 * 1024 * N_DEPTH * FLOP_PER_SAMPLE_FWD
        pk_ = peaks()
        render = {"ms_per_frame": ms_frame, "n_gpus": world,
                  "frame": f"1024x1024 rays x {N_DEPTH} samples, static + dynamic composite + both component images; detector rows sharded over the ranks",
                  "rays_per_s": 1024 * 1024 / (ms_frame * 1e-3), "tflops": flop_frame / (ms_frame * 1e-3) / 1e12,
                  "frac_of_tensor_peak_burst": flop_frame / (ms_frame * 1e-3) / 1e12 / (world * pk_["bf16_burst"]),
                  "frac_of_tensor_peak_sustained": flop_frame / (ms_frame * 1e-3) / 1e12 / (world * pk_["bf16_sustained"]),
                  "finite": bool(torch.isfinite(pix_r).all().item())}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    dom = max(kt, key=lambda k: kt[k]["ms_per_step"]) if kt else None
    roof = None
    if dom:
        d = kt[dom]
        # algorithmic FLOPs of the family per step / launches of that family per step
        fam_flop = {"field_forward": FLOP_PER_SAMPLE_FWD, "field_backward": FLOP_PER_SAMPLE - FLOP_PER_SAMPLE_FWD}.get(dom, 0) * B * N_DEPTH
        d["flop_per_launch"] = fam_flop / d["launches_per_step"]
        achieved = d["flop_per_launch"] / (d["ms_per_launch"] * 1e-3) / 1e12
        tr_ = measured_traffic(dom) if HIDDEN == 128 else None     # (the committed ncu captures are of the fused kernels)
        step_tflops = B * N_DEPTH * FLOP_PER_SAMPLE / (ms_per_step * 1e-3) / 1e12
        # the timed region is a few hundred ms at full clocks: the burst cuBLAS figure is the comparator (the sustained one was
        # measured power-throttled at ~1.3 GHz); both fractions are reported
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_burst"], "frac_of_sustained_peak": achieved / pk["bf16_sustained"],
                "peak_sustained": pk["bf16_sustained"], "traffic": tr_["dram_bytes_per_launch"] if tr_ else None,
                "traffic_source": tr_["source"] if tr_ else None, "peak_source": pk["source"] + " (burst cuBLAS bf16)",
                "ms_per_launch": d["ms_per_launch"], "launches_per_step": d["launches_per_step"],
                "share_of_step": d["ms_per_step"] / ms_per_step,
                "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / pk["bf16_burst"],
                "whole_step_frac_of_sustained_peak": step_tflops / pk["bf16_sustained"],
                "kernels": {k: {"ms_per_step": v["ms_per_step"], "launches_per_step": v["launches_per_step"]} for k, v in kt.items()}}

    eager = None
    if world == 1 and not args.no_eager_baseline:
        step = gpu_eager_step_fn(dev, N_RAYS)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0.record()
        n_e = 5
        for _ in range(n_e):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1) / n_e
        eager = {"value": N_RAYS / (ms_e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e,
                 "what": f"the reference's eager PyTorch step (oracle port; it times identically to the unmodified reference) on this GPU: fp32 "
                         f"sgemm (allow_tf32 off), autograd, torch.optim.Adam + LinearLR, {N_RAYS} x {N_DEPTH} batches resident in HBM, "
                         f"torch {torch.__version__}"}
        del step
        torch.cuda.empty_cache()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 128
        rate, sec = time_cpu_arm(n_cpu, 3, 1)
        cpu = {"value": rate, "unit": "rays/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{n_cpu} rays x {N_DEPTH} samples per step, 3 timed steps after 1 warm-up, oracle port of the reference "
                         f"training step on CPU torch ({sec:.2f} s/step)"}

    if world > 1:
        coll = ("gradient + loss-sum exchange fused with the Adam kernel over NVLink peer memory (nerfca_allreduce_adam_step, no NCCL call in the step)"
                if trainer.peer_grads is not None else "NCCL all-reduce of the flat gradient buffer + separate Adam kernel")
        par = f"rays sharded x{world} ({B} per GPU, {scaling} scaling), {coll}"
    else:
        par = "single GPU"
    line = {"metric": "training rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config({"parallelism": par, "baseline_config": args.config, "global_batch_rays": n_global,
                                       "collective": None if world == 1 else ("peer-memory fused" if trainer.peer_grads is not None else "nccl"),
                                       "replicas_bit_identical": replicas_identical, "cuda_graph": graph_stats,
                                       "loss_last_step": float(trainer.loss_from(last_terms, n_global))}),
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "loss_last_step": losses[-1] if losses else None,
                    "api": "per step: rays_train[ids] / phases_train[ids] gathered on the HOST from the host ray table (as run_composite.py:250-265 does), "
                           "CompositeTrainer.step_host_async(rays[B,4,3] f64, phases[B], t_rand[N]), loss read one step late",
                    "dropin_modules": dropin,
                    "device_ray_table": {"value": e2e_ids_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_ids, "d2h_bytes_per_step": d2h,
                                         "api": "CompositeTrainer.step_ids_async(ids[B] i64, t_rand[N]) -- ray table resident in HBM, "
                                                "batch rows gathered by nerfca_gather_batch inside the step's graph"}},
            "roofline": roof, "render": render, "cpu_baseline": cpu, "gpu_eager_baseline": eager}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Shared implementation of the two coordinate-MLP fields behind model/CPPN.py and model/Temporal.py.

The modules own ordinary fp32 nn.Parameters under the reference's state_dict keys
(early_pts_layers.{0,2,..}.{weight,bias}, output_linear.0.{weight,bias}, time_latents), so
torch.optim.Adam(list(model.parameters())) and .save() of the drivers keep working
(reference: model/CPPN.py:40-65,164-180, model/Temporal.py:23-26,62-93,206-222).
All arithmetic runs in the CUDA library through nerfca.ops.FieldFunction.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class CoordinateField(nn.Module):
    """Encoder + `in -> H, n x (H -> H), H -> out` ReLU MLP; optional learned per-phase latent appended to the input."""

    version = "v0.00"
    n_latent_rows = 10  # the reference hard-codes torch.arange(0, 10) phases (model/Temporal.py:25)

    def _setup(self, model_definition: dict, temporal: bool) -> None:
        md = model_definition
        self.model_definition = md
        self.device = md["device"]
        self.num_early_layers = md["num_early_layers"]
        self.num_late_layers = md["num_late_layers"]
        self.num_filters = md["num_filters"]
        self.num_input_channels = md["num_input_channels"]
        self.num_output_channels = md["num_output_channels"]
        self.use_bias = md["use_bias"]
        self.use_pos_enc = md["pos_enc"]
        self.first_act_func = nn.ReLU()
        self.act_func = nn.ReLU()
        self.precision = md.get("precision", None)  # None -> NERFCA_PRECISION env / library default
        if self.num_input_channels != 3:
            raise NotImplementedError("the CUDA fields take (x, y, z) inputs (num_input_channels == 3)")

        feats = self.num_input_channels
        if self.use_pos_enc != "none":
            self.pos_enc_basis = md["pos_enc_basis"]
            self.pos_enc_window_start = md["pos_enc_window_start"]
            feats = self.num_input_channels * (1 + 2 * self.pos_enc_basis)
            if self.use_pos_enc == "fourier":
                feats = self.num_input_channels * 2 * self.pos_enc_basis
                self.fourier_sigma = md["fourier_sigma"]
                self.fourier_coefficients = (md["fourier_gaussian"] * self.fourier_sigma).to(self.device)
            if temporal:
                self.windowed_alpha = 0

        self.use_time_latents = bool(md.get("use_time_latents", False)) if temporal else False
        self.num_time_dim = 0
        if temporal:
            self.num_input_times = md["num_input_times"]
            self.input_features_pts = feats
            self.input_features_time = self.num_input_times
            if self.use_time_latents:
                self.num_time_dim = md["num_time_dim"]
                self.fixed_frame_ids = torch.arange(0, self.n_latent_rows)
                self.time_latents = nn.Parameter(torch.rand((self.n_latent_rows, self.num_time_dim)))
            feats = feats + (self.num_time_dim if self.use_time_latents else self.num_input_times)
        self.input_features = feats

        # same construction order as the reference so that a given torch seed yields the same initial weights
        seq = [nn.Linear(self.input_features, self.num_filters, bias=self.use_bias), self.first_act_func]
        for _ in range(self.num_early_layers):
            seq += [nn.Linear(self.num_filters, self.num_filters, bias=self.use_bias), self.act_func]
        self.early_pts_layers = nn.ModuleList(seq)
        if self.num_late_layers > 0:
            self.skip_connection = nn.Sequential(nn.Linear(self.num_filters + self.input_features, self.num_filters,
                                                           bias=self.use_bias), self.act_func)
            late = []
            for _ in range(self.num_late_layers - 1):
                late += [nn.Linear(self.num_filters, self.num_filters, bias=self.use_bias), self.act_func]
            self.late_pts_layers = nn.ModuleList(late)
        self.output_linear = nn.Sequential(nn.Linear(self.num_filters, self.num_output_channels, bias=self.use_bias))

        self.store_activations = False
        self.activation_dictionary = {}
        self._band_cache = (None, None)

    # ---- model-understanding API kept for interface parity (model/CPPN.py:82-86) ----
    def activations(self, store_activations: bool) -> None:
        self.store_activations = store_activations
        if not store_activations:
            self.activation_dictionary = {}

    # ---- frequency schedules (host scalars; model/CPPN.py:137-162) ----
    def update_freq_mask_alpha(self, current_iter, max_iter):
        n = self.pos_enc_basis
        if current_iter >= max_iter:
            self.freq_mask_alpha = torch.ones(n).float()
            self.windowed_alpha = n + 1
            return
        ptr = (n * current_iter) / max_iter + self.pos_enc_window_start
        whole = int(ptr)
        mask = np.zeros(n)
        mask[: whole + 1] = 1.0
        mask[whole: whole + 1] = ptr - whole
        self.freq_mask_alpha = torch.clip(torch.from_numpy(mask), 1e-8, 1 - 1e-8).float()
        self.windowed_alpha = ptr

    def update_windowed_alpha(self, current_iter, max_iter):
        self.windowed_alpha = (self.pos_enc_basis * current_iter) / max_iter

    def _eased_window(self, n_bands):
        bands = torch.arange(0, n_bands)
        x = torch.clip(self.windowed_alpha - bands, 0.0, 1.0)
        return (0.5 * (1 + torch.cos(torch.pi * x + torch.pi))).float()

    # ---- descriptor plumbing ----
    def _device(self):
        return self.output_linear[0].weight.device

    def _band_weight(self, n_bands):
        """Per-band window on the device: freq_mask_alpha / eased window / None (unwindowed)."""
        if self.use_pos_enc == "free_windowed":
            host = self.freq_mask_alpha
        elif self.use_pos_enc == "nerfies_windowed":
            host = self._eased_window(n_bands)
        else:
            return None
        key, dev_t = self._band_cache
        dev = self._device()
        if key is None or dev_t.device != dev or key.shape != host.shape or not torch.equal(key, host.cpu()):
            dev_t = host.to(device=dev, dtype=torch.float32).contiguous()
            self._band_cache = (host.detach().cpu().clone(), dev_t)
        return dev_t

    def _spec(self, n_bands=None) -> ops.FieldSpec:
        if self.num_late_layers > 0:
            raise NotImplementedError("num_late_layers > 0 (skip connection) is not built; the shipped configs use 0 "
                                      "and Temporal.query_time is broken for it upstream (model/Temporal.py:128-136)")
        if self.num_output_channels != 1:
            raise NotImplementedError("num_output_channels must be 1")
        if self.use_pos_enc == "none":
            mode, n_freq = L.ENC_NONE, 0
        else:
            n_freq = self.pos_enc_basis if n_bands is None else n_bands
            mode = L.ENC_FOURIER if self.use_pos_enc == "fourier" else L.ENC_BANDS
            if n_freq <= 0:
                mode = L.ENC_NONE
        spec = ops.FieldSpec(enc_mode=mode, n_freq=n_freq, n_latent=self.num_time_dim,
                             n_phases=self.time_latents.shape[0] if self.num_time_dim else 0, hidden=self.num_filters,
                             n_hidden=self.num_early_layers, use_bias=self.use_bias)
        if mode == L.ENC_BANDS:
            spec.band_weight = self._band_weight(n_freq)
        elif mode == L.ENC_FOURIER:
            spec.fourier_coeff = self.fourier_coefficients.to(device=self._device(), dtype=torch.float32).contiguous()
        return spec

    def _param_list(self):
        """Canonical order of the C descriptor: [time_latents], then (weight, bias) per Linear, output last."""
        out = [self.time_latents] if self.num_time_dim else []
        for lin in list(self.early_pts_layers)[0::2] + [self.output_linear[0]]:
            out.append(lin.weight)
            if self.use_bias:
                out.append(lin.bias)
        return out

    def _precision_code(self) -> int:
        name = self.precision or ops.default_precision()
        if name not in ops.PRECISIONS:
            raise ValueError(f"unknown precision {name!r}; expected one of {sorted(ops.PRECISIONS)}")
        return ops.PRECISIONS[name]

    def _evaluate(self, samples: ops.Samples) -> torch.Tensor:
        return ops.FieldFunction.apply(self._spec(), samples, self._precision_code(), *self._param_list())

    def _encode(self, values: torch.Tensor, pos_enc_basis: int) -> torch.Tensor:
        """pos_enc of the reference (model/CPPN.py:112-135): encoding only, no latent columns."""
        if self.use_pos_enc == "none" or pos_enc_basis <= 0:
            return values
        spec = self._spec(pos_enc_basis)
        spec.n_latent, spec.n_phases = 0, 0
        flat = values.reshape(-1, 3)
        params = self._param_list()[(1 if self.num_time_dim else 0):]
        enc = ops.encode(spec, ops.Samples.from_points(flat), params)
        return enc.reshape(*values.shape[:-1], spec.enc_dim)

    def _checkpoint(self, filename: str, training_information: dict) -> None:
        payload = {"version": self.version, "parameters": self.model_definition,
                   "training_information": training_information, "model": self.state_dict()}
        if "nerfies_windowed" in self.use_pos_enc:
            payload["windowed_alpha"] = self.windowed_alpha
        if "free_windowed" in self.use_pos_enc:
            payload["freq_mask_alpha"] = self.freq_mask_alpha
        torch.save(payload, f=filename)

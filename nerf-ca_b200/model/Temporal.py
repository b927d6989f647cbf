"""Drop-in for the reference's model/Temporal.py: the dynamic vessel field with a learned
per-cardiac-phase latent appended to the encoded position (reference model/Temporal.py:5-222).
"""
import torch

from nerfca import ops
from nerfca.fields import CoordinateField


class Temporal(CoordinateField):
    def __init__(self, model_definition: dict) -> None:
        super().__init__()
        self._setup(model_definition, temporal=True)

    def create_time_net(self):  # layers are built in __init__; kept so callers probing the attribute still find it
        return None

    def forward_composite(self, x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        """[P,3], phase per sample (int or float tensor) -> [P,1] (reference model/Temporal.py:138-151)."""
        if not self.use_time_latents:
            raise UnboundLocalError("forward_composite needs use_time_latents=True (reference model/Temporal.py:142-149)")
        return self._evaluate(ops.Samples.from_points(x, ts))

    def forward_rays(self, samples: ops.Samples) -> torch.Tensor:
        return self._evaluate(samples)

    def query_time(self, xs: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        """MLP on [enc(xs) | ts] with explicit per-sample latent rows ts[P,T] (reference model/Temporal.py:113-136):
        evaluated as a gather from a P-row latent table so the same kernels serve it."""
        if ts.dim() != 2 or ts.shape[0] != xs.shape[0] or ts.shape[1] != self.num_time_dim:
            raise ValueError("query_time expects one latent row of num_time_dim per sample")
        spec = self._spec()
        spec.n_phases = ts.shape[0]
        idx = torch.arange(ts.shape[0], device=xs.device, dtype=torch.int32)
        params = [ts.to(torch.float32).contiguous()] + self._param_list()[1:]
        return ops.FieldFunction.apply(spec, ops.Samples.from_points(xs, idx), self._precision_code(), *params)

    def pos_enc(self, values, pos_enc_basis):
        return self._encode(values, pos_enc_basis)

    def windowed_pos_enc(self, pos_enc_basis):
        return self._eased_window(pos_enc_basis).to(self.device)

    def save(self, filename: str, training_information: dict) -> None:
        self._checkpoint(filename, training_information)

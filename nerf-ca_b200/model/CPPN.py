"""Drop-in for the reference's model/CPPN.py: the static background field.

Same constructor dict, attributes, state_dict keys and methods (reference model/CPPN.py:5-180);
the arithmetic runs in libnerfca_b200.so (tcgen05 bf16 or fp32 SIMT, see nerfca.fields).
"""
import torch

from nerfca import ops
from nerfca.fields import CoordinateField


class CPPN(CoordinateField):
    def __init__(self, model_definition: dict) -> None:
        super().__init__()
        self._setup(model_definition, temporal=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[P,3] -> [P,1] raw attenuation (reference model/CPPN.py:88-110)."""
        return self._evaluate(ops.Samples.from_points(x))

    def forward_rays(self, samples: ops.Samples) -> torch.Tensor:
        """Same field evaluated on a ray-generated sample set; points are formed inside the kernel."""
        return self._evaluate(samples)

    def pos_enc(self, values, pos_enc_basis, type):
        return self._encode(values, pos_enc_basis)

    def windowed_pos_enc(self, pos_enc_basis, type):
        return self._eased_window(pos_enc_basis).to(self.device)

    def save(self, filename: str, training_information: dict) -> None:
        self._checkpoint(filename, training_information)

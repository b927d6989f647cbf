"""nerfca: host side of the B200-native NeRF-CA inner loop (ctypes over libnerfca_b200.so)."""
from . import _lib  # noqa: F401
from .ops import (FieldFunction, FieldSpec, IntegrateFunction, LossConfig, Samples, composite_loss,  # noqa: F401
                  default_precision, loss_from_terms)

"""ctypes binding of libnerfca_b200.so (C ABI declared in include/nerfca.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Device pointers are passed as raw integers taken from torch tensors;
torch is only the owner of device memory and of the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

ABI_VERSION = 2
MAX_LAYERS = 8
N_LOSS_TERMS = 16

OK, E_ARG, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = 0, -1, -2, -3, -4
F32, F64 = 0, 1
ENC_NONE, ENC_BANDS, ENC_FOURIER = 0, 1, 2
ACT_SIGMOID, ACT_SOFTPLUS, ACT_CLAMP = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1

# indices of nerfca_composite_loss's term vector (include/nerfca.h)
T_PIXEL_SUM, T_BLENDW_SUM, T_SIGMA_S_MAX, T_SIGMA_D_MAX, T_FAVOR_SUM, T_S_ENT_SUM, T_S_SUM_SUM, T_D_ENT_SUM, \
    T_D_SUM_SUM, T_OCCL_SUM, T_L1_SUM, T_L2_SUM = range(12)


class FieldStruct(C.Structure):
    _fields_ = [("enc_mode", C.c_int32), ("n_freq", C.c_int32), ("n_latent", C.c_int32), ("n_phases", C.c_int32),
                ("hidden", C.c_int32), ("n_hidden", C.c_int32),
                ("band_weight", C.c_void_p), ("fourier_coeff", C.c_void_p), ("latents", C.c_void_p),
                ("weight", C.c_void_p * MAX_LAYERS), ("bias", C.c_void_p * MAX_LAYERS)]


class FieldGradsStruct(C.Structure):
    _fields_ = [("latents", C.c_void_p), ("weight", C.c_void_p * MAX_LAYERS), ("bias", C.c_void_p * MAX_LAYERS)]


class SamplesStruct(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("points", C.c_void_p), ("n_rays", C.c_int32), ("n_depth", C.c_int32),
                ("origins", C.c_void_p), ("dirs", C.c_void_p), ("ray_dtype", C.c_int32), ("ray_stride", C.c_int32),
                ("depth", C.c_void_p), ("phase_point", C.c_void_p), ("phase_ray", C.c_void_p)]


class LossCfgStruct(C.Structure):
    _fields_ = [("favor_s_weight", C.c_double), ("dyn_entropy_weight", C.c_double), ("occl_weight", C.c_double),
                ("l1_weight", C.c_double), ("entro_mask_thre", C.c_double), ("entro_weighted_thresh", C.c_double),
                ("entro_use_weighting", C.c_int32), ("n_rays_global", C.c_int32)]


class StepStruct(C.Structure):
    _fields_ = [("static_field", C.POINTER(FieldStruct)), ("dynamic_field", C.POINTER(FieldStruct)),
                ("static_grads", C.POINTER(FieldGradsStruct)), ("dynamic_grads", C.POINTER(FieldGradsStruct)),
                ("samples", C.POINTER(SamplesStruct)), ("precision", C.c_int32), ("activation", C.c_int32),
                ("i0", C.c_void_p), ("gt", C.c_void_p), ("wpix", C.c_void_p), ("gw_stride", C.c_int32), ("flags", C.c_int32),
                ("loss", C.POINTER(LossCfgStruct)),
                ("raw_s", C.c_void_p), ("raw_d", C.c_void_p), ("d_raw_s", C.c_void_p), ("d_raw_d", C.c_void_p),
                ("stash", C.c_void_p), ("workspace", C.c_void_p), ("pix_out", C.c_void_p), ("terms_out", C.c_void_p)]


class AdamStepStruct(C.Structure):       # nerfca_adam_step_t: the scalars of ONE update, computed by the host as torch does
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("bias_correction1", C.c_double), ("bias_correction2_sqrt", C.c_double)]


class RepackStruct(C.Structure):         # nerfca_repack_t
    _fields_ = [("static_field", C.POINTER(FieldStruct)), ("dynamic_field", C.POINTER(FieldStruct)), ("workspace", C.c_void_p)]


STEP_PACKED, STEP_ZERO_TERMS = 1, 2


class PeersStruct(C.Structure):          # nerfca_peers_t
    _fields_ = [("rank", C.c_int32), ("world_size", C.c_int32), ("grads", C.c_void_p), ("signals", C.c_void_p),
                ("own_signals", C.c_void_p)]


K_RAYS, K_PACK, K_FIELD_FWD, K_LOSS, K_FIELD_BWD, K_ADAM, K_COUNT = range(7)
KERNEL_FAMILY_NAMES = {K_RAYS: "rays", K_PACK: "pack_params", K_FIELD_FWD: "field_forward", K_LOSS: "integral_loss",
                       K_FIELD_BWD: "field_backward", K_ADAM: "adam"}




LIB_NAME = os.environ.get("NERFCA_LIB", "libnerfca_b200.so")      # developer aid: NERFCA_LIB=libnerfca_b200_tl.so (make TL=1) prints which bounded wait gave up
LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), LIB_NAME)

_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    "nerfca_last_error": (C.c_char_p, []),
    "nerfca_abi_version": (C.c_int, []),
    "nerfca_gen_rays": (C.c_int, [_P, _I32, _I32, _F, _F, _F, _F, _F, _P, _P, _P]),
    "nerfca_jitter_depth": (C.c_int, [_P, _P, _I32, _P, _P]),
    "nerfca_gather_batch": (C.c_int, [_P, _P, _I64, _P, _I32, _P, _P, _P, _P]),
    "nerfca_sample_points": (C.c_int, [C.POINTER(SamplesStruct), _P, _P]),
    "nerfca_encode": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _P, _P]),
    "nerfca_field_stash_bytes": (C.c_size_t, [C.POINTER(FieldStruct), _I64, _I32]),
    "nerfca_field_workspace_bytes": (C.c_size_t, [C.POINTER(FieldStruct), _I64, _I32, _I32]),
    "nerfca_field_forward": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32, _P, _P, _P, _P]),
    "nerfca_field_backward": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32, _P, _P, _P,
                                        C.POINTER(FieldGradsStruct), _P]),
    "nerfca_integrate": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "nerfca_integrate_backward": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "nerfca_composite_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, C.POINTER(LossCfgStruct), _P, _P,
                                        _P, _P, _P]),
    "nerfca_step_stash_bytes": (C.c_size_t, [C.POINTER(StepStruct)]),
    "nerfca_step_workspace_bytes": (C.c_size_t, [C.POINTER(StepStruct)]),
    "nerfca_train_step": (C.c_int, [C.POINTER(StepStruct), _P]),
    "nerfca_fields_forward": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32, _P, _P, _P, _P]),
    "nerfca_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, C.POINTER(AdamStepStruct), _F, _I32, C.POINTER(RepackStruct), _P]),
    "nerfca_allreduce_adam_step": (C.c_int, [C.POINTER(PeersStruct), C.c_uint32, _P, _P, _P, _P, _I64, C.POINTER(AdamStepStruct),
                                             C.POINTER(RepackStruct), _I64, _P, _P]),
    "nerfca_graph_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "nerfca_graph_begin": (C.c_int, [_P, _P]),
    "nerfca_graph_end_launch": (C.c_int, [_P, _P]),
    "nerfca_graph_abort": (C.c_int, [_P, _P]),
    "nerfca_graph_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nerfca_graph_destroy": (C.c_int, [_P]),
    "nerfca_debug_x0": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32, _P, C.POINTER(C.c_int32), _P]),
    "nerfca_render_workspace_bytes": (C.c_size_t, [C.POINTER(FieldStruct), C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32]),
    "nerfca_render_rays": (C.c_int, [C.POINTER(FieldStruct), C.POINTER(FieldStruct), C.POINTER(SamplesStruct), _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "nerfca_normalize_image": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "nerfca_launch_count": (C.c_int64, []),
    "nerfca_profile_enable": (C.c_int, [_I32]),
    "nerfca_profile_read": (C.c_int, [_I32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library once; raise if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_NAME} not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
                               f"g.build()'` or `make -C nerf-ca_b200/csrc` (there is no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.nerfca_abi_version() != ABI_VERSION:
            raise RuntimeError("libnerfca_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != OK:
        msg = load().nerfca_last_error().decode("utf-8", "replace")
        if rc == E_UNSUPPORTED:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == E_ARG:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Raw device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("nerfca_b200 kernels need CUDA tensors; this build has no CPU path")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream

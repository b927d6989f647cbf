"""Drop-in for the reference's train/proj_helpers.py (cone-beam pose + per-pixel rays + depth grid).

Pose algebra is host-side float64 numpy exactly as upstream (train/proj_helpers.py:5-63); the
per-pixel ray table is produced on the GPU by nerfca_gen_rays, bit-identical to the reference's
fp32 torch expression (train/proj_helpers.py:65-90).
"""
import ctypes as C

import numpy as np
import torch

from nerfca import _lib as L
from nerfca import ops


def _rotation(axis: int, angle: float) -> np.ndarray:
    """Homogeneous right-handed rotation about coordinate axis 0/1/2."""
    c, s = np.cos(angle), np.sin(angle)
    m = np.identity(4)
    a, b = [(1, 2), (2, 0), (0, 1)][axis]
    m[a, a], m[a, b], m[b, a], m[b, b] = c, -s, s, c
    return m


def x_rotation_matrix(angle):
    return _rotation(0, angle)


def y_rotation_matrix(angle):
    return _rotation(1, angle)


def z_rotation_matrix(angle):
    return _rotation(2, angle)


def translation_matrix(vec):
    m = np.identity(4)
    m[:3, 3] = vec[:3]
    return m


def get_rotation(theta, phi, larm):
    # roadmap-run geometry: inverse of Rz(larm) Rx(theta) Ry(phi)   (reference :34-37)
    fwd = z_rotation_matrix(np.deg2rad(larm)) @ x_rotation_matrix(np.deg2rad(theta)) @ y_rotation_matrix(np.deg2rad(phi))
    return np.linalg.inv(fwd)


def source_matrix(source_pt, theta, phi, larm=0, translation=[0, 0, 0]):
    table = translation_matrix([translation[0], translation[1], translation[2], 1])
    return table.dot(get_rotation(theta, phi, larm).dot(translation_matrix(source_pt)))


def get_rotation_matrix_tigre(theta, phi, larm=0):
    # Rz(-theta) . Rz(pi/2) . Rx(phi) . Rx(-pi/2); larm is accepted and ignored, as upstream (:50-57)
    r1, r2 = x_rotation_matrix(-np.pi / 2), x_rotation_matrix(np.deg2rad(phi))
    r3, r4 = z_rotation_matrix(np.pi / 2), z_rotation_matrix(-np.deg2rad(theta))
    return np.dot(np.dot(r4, np.dot(r3, r2)), r1)


def source_matrix_tigre(source_pt, theta, phi, larm=0):
    return get_rotation_matrix_tigre(theta, phi, larm).dot(translation_matrix(source_pt))


def ray_values_tigre_device(theta, phi, larm, geo, device):
    """(origins, directions) [W,H,3] float32 CUDA tensors for one projection view."""
    pose = source_matrix_tigre(np.array([0, 0, -geo["DSO"]]), theta, phi, larm).astype(np.float32)
    pose = np.ascontiguousarray(pose)
    w, h = int(geo["nDetector"][0]), int(geo["nDetector"][1])
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("nerfca_b200 generates rays on the GPU; pass a CUDA device")
    origins = torch.empty((w, h, 3), dtype=torch.float32, device=dev)
    dirs = torch.empty((w, h, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().nerfca_gen_rays(pose.ctypes.data_as(C.c_void_p), w, h, float(np.float32(geo["dDetector"][0])),
                                      float(np.float32(geo["dDetector"][1])), float(np.float32(geo["offDetector"][0])),
                                      float(np.float32(geo["offDetector"][1])), float(np.float32(geo["DSD"])), L.ptr(origins),
                                      L.ptr(dirs), L.stream_ptr())
    L.check(rc, "nerfca_gen_rays")
    return origins, dirs


def get_ray_values_tigre(theta, phi, larm, geo, device):
    """Same return convention as upstream: two numpy arrays [W,H,3] (reference :87-90)."""
    origins, dirs = ray_values_tigre_device(theta, phi, larm, geo, device)
    return origins.cpu().numpy(), dirs.cpu().numpy()


def get_depth_values(near_thresh, far_thresh, depth_samples_per_ray, device, stratified=True):
    t_vals = torch.linspace(0., 1., depth_samples_per_ray)
    z_vals = near_thresh * (1. - t_vals) + far_thresh * t_vals
    if not stratified:
        return z_vals.to(device)
    t_rand = torch.rand(z_vals.shape)  # CPU generator, same stream as the reference (:102)
    return ops.jitter_depth(z_vals.to(device), t_rand)

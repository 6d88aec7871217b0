"""rot6d -> rotation matrix -- mirror of ``utils/rotation_conversions.py:513-534``.

Only ``rotation_6d_to_matrix`` (and its trivial inverse ``matrix_to_rotation_6d``) are on the
sampling path (model/rotation2xyz.py:202,270).  ``quaternion_to_matrix`` / ``axis_angle_to_matrix`` exist because
``Rotation2xyz.__call__`` reaches them for ``pose_rep='rotquat' | 'rotvec'`` and for ``glob=False`` (a constant
``glob_rot``); they are a handful of torch ops, off the hot path.  The Euler utilities of the reference file serve data
preparation and training losses and are out of scope.
"""
import torch

from . import _lib


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """Gram-Schmidt per Zhou et al. 2019: d6 (*, 6) -> (*, 3, 3) with rows (b1, b2, b3)."""
    if d6.shape[-1] != 6:
        raise ValueError("rotation_6d_to_matrix expects (*, 6), got %s" % (tuple(d6.shape),))
    _lib.require_cuda_f32(d6, "d6")
    src = d6.contiguous()
    out = torch.empty(d6.shape[:-1] + (3, 3), device=d6.device, dtype=torch.float32)
    n = src.numel() // 6
    _lib.check(_lib.lib().regen_rot6d_to_matrix(_lib.ptr(src), _lib.ptr(out), n, _lib.stream_ptr(d6.device)),
               "regen_rot6d_to_matrix")
    return out


def matrix_to_rotation_6d(matrix: torch.Tensor) -> torch.Tensor:
    """utils/rotation_conversions.py:537-552 (drops the last row; a view + copy, no arithmetic)."""
    return matrix[..., :2, :].clone().reshape(*matrix.size()[:-2], 6)


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """Unit-free quaternion (real part first) (*, 4) -> rotation matrix (*, 3, 3); the standard formula
    R = I + 2s(...) with s = 2 / |q|^2 (what utils/rotation_conversions.py:41-69 computes)."""
    w, x, y, z = torch.unbind(quaternions, -1)
    s = 2.0 / (quaternions * quaternions).sum(-1)
    rows = (1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
            s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
            s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y))
    return torch.stack(rows, -1).reshape(quaternions.shape[:-1] + (3, 3))


def axis_angle_to_quaternion(axis_angle: torch.Tensor) -> torch.Tensor:
    """Rotation vector (*, 3) -> quaternion (*, 4), real part first (utils/rotation_conversions.py:456-486: series
    expansion of sin(a/2)/a below 1e-6)."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angles), angles)
    ratio = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half) / safe)
    return torch.cat([torch.cos(half), axis_angle * ratio], dim=-1)


def axis_angle_to_matrix(axis_angle: torch.Tensor) -> torch.Tensor:
    """utils/rotation_conversions.py:420-434."""
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))

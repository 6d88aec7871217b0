"""rot6d -> rotation matrix -- mirror of ``utils/rotation_conversions.py:513-534``.

Only ``rotation_6d_to_matrix`` (and its trivial inverse ``matrix_to_rotation_6d``) are on the
sampling path (model/rotation2xyz.py:202,270); the quaternion / axis-angle / Euler utilities of
the reference file serve data preparation and training losses and are out of scope.
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

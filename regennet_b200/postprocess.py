"""Post-sampling tail on the GPU.

The reference round-trips every generated batch through the CPU for
``scipy.ndimage.gaussian_filter1d(sample.cpu().numpy(), sigma=1, axis=-1)`` (sample/cgenerate.py:142-143; sigma=3 in
render/crendermotion.py:79) and then converts rot6d to rotation matrices inside ``model.rot2xyz``
(model/rotation2xyz.py:253-270).  Here both run as library kernels on the sampler's output, without leaving the device.
"""
import torch

from . import _lib
from .gaussian_diffusion import _layout_of


def gaussian_filter1d_time(sample: torch.Tensor, sigma: float = 1.0, truncate: float = 4.0) -> torch.Tensor:
    """``gaussian_filter1d(sample, sigma, axis=-1)`` (mode='reflect') for a [..., T] CUDA float32 tensor.
    Accepts the contiguous layout and the permuted ([T,B,J,F]-ordered) layout the sampler returns; the result has the
    same layout."""
    _lib.require_cuda_f32(sample, "sample")
    T = sample.shape[-1]
    lay = _layout_of(sample) if sample.dim() == 4 else ("bjft" if sample.is_contiguous() else None)
    if lay is None:
        sample = sample.contiguous()
        lay = "bjft"
    out = torch.empty_like(sample)  # preserves the (dense) strides
    n_cols = sample.numel() // max(T, 1)
    rc = _lib.lib().regen_gaussian_filter1d_time(_lib.ptr(sample), _lib.ptr(out), n_cols, T, 1 if lay == "tbi" else 0,
                                                 float(sigma), float(truncate), _lib.stream_ptr(sample.device))
    _lib.check(rc, "regen_gaussian_filter1d_time")
    return out


def smooth_rot6d_to_matrix(sample: torch.Tensor, sigma: float = 1.0, truncate: float = 4.0,
                           translation: bool = True) -> torch.Tensor:
    """sample [B, J, 6, T] (rot6d, last joint = translation if ``translation``) -> temporally smoothed rotation matrices
    [B, T, J-1 (or J), 3, 3]: the fused equivalent of cgenerate's gaussian_filter1d followed by rot2xyz's
    ``rotation_6d_to_matrix(x[:, :-1].permute(0, 3, 1, 2))``."""
    from .gaussian_diffusion import _to_layout
    _lib.require_cuda_f32(sample, "sample")
    B, J, F, T = sample.shape
    if F != 6:
        raise ValueError("smooth_rot6d_to_matrix expects rot6d features (F=6), got %d" % F)
    x_tbi = _to_layout(sample, "tbi").permute(3, 0, 1, 2)  # [T,B,J,6] contiguous
    drop = 1 if translation else 0
    R = torch.empty((B, T, J - drop, 3, 3), device=sample.device, dtype=torch.float32)
    rc = _lib.lib().regen_smooth_rot6d_to_matrix(_lib.ptr(x_tbi), _lib.ptr(R), T, B, J, drop, float(sigma),
                                                 float(truncate), _lib.stream_ptr(sample.device))
    _lib.check(rc, "regen_smooth_rot6d_to_matrix")
    return R

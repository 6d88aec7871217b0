"""GPU parity: the post-sampling tail (temporal Gaussian filter, fused filter + rot6d->rotmat) vs scipy / the oracle.
scipy.ndimage.gaussian_filter1d is the third-party routine the reference calls at sample/cgenerate.py:142."""
import numpy as np
import pytest
import torch
from scipy.ndimage import gaussian_filter1d

from oracle import sampler_ref
from regennet_b200 import postprocess

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sigma", [1.0, 3.0])
@pytest.mark.parametrize("shape", [(3, 56, 6, 60), (2, 263, 1, 196), (1, 4, 6, 3), (2, 5, 6, 1)])
@pytest.mark.parametrize("layout", ["bjft", "tbi"])
def test_gaussian_filter_matches_scipy(built_lib, sigma, shape, layout):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    want = gaussian_filter1d(x.numpy(), sigma=sigma, axis=-1)
    xc = x.cuda()
    if layout == "tbi":
        xc = xc.permute(3, 0, 1, 2).contiguous().permute(1, 2, 3, 0)
    got = postprocess.gaussian_filter1d_time(xc, sigma=sigma)
    assert got.shape == x.shape and got.stride() == xc.stride()
    # scipy accumulates in double and rounds once to float32, as the kernel does
    assert np.abs(got.cpu().numpy() - want).max() < 2e-7 * max(1.0, float(np.abs(want).max()))


def test_fused_smooth_rot6d_matches_reference_sequence(built_lib):
    g = torch.Generator().manual_seed(3)
    B, J, T = 4, 56, 60
    sample = torch.randn(B, J, 6, T, generator=g)
    # reference sequence: scipy filter on the CPU, then rot2xyz's rot6d -> matrix on x[:, :-1].permute(0, 3, 1, 2)
    gf = torch.from_numpy(gaussian_filter1d(sample.numpy(), sigma=1, axis=-1))
    want = sampler_ref.rotation_6d_to_matrix(gf[:, :-1].permute(0, 3, 1, 2))
    got = postprocess.smooth_rot6d_to_matrix(sample.cuda().permute(3, 0, 1, 2).contiguous().permute(1, 2, 3, 0), sigma=1.0)
    assert got.shape == (B, T, J - 1, 3, 3)
    assert torch.allclose(got.cpu(), want, atol=2e-6)
    eye = got @ got.transpose(-1, -2)
    assert torch.allclose(eye.cpu(), torch.eye(3).expand_as(eye), atol=1e-5)


def test_filter_preserves_constants_and_rejects_bad_sigma(built_lib):
    x = torch.full((2, 3, 6, 17), 2.5, device="cuda")
    assert torch.allclose(postprocess.gaussian_filter1d_time(x, 1.0), x, atol=1e-6)   # weights sum to one
    with pytest.raises(RuntimeError):
        postprocess.gaussian_filter1d_time(x, sigma=0.0)
    with pytest.raises(RuntimeError):
        postprocess.gaussian_filter1d_time(x.cpu(), sigma=1.0)

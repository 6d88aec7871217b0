"""GPU parity: handle-free elementwise kernels (through the C ABI) vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_ref
from regennet_b200 import gaussian_diffusion as gd
from regennet_b200 import respace, rotation_conversions

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _diffusion(rs="ddim20"):
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    return respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, rs if rs else [1000]), betas=betas,
                                   model_mean_type=gd.ModelMeanType.START_X,
                                   model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)


def _perm(t):
    """same logical tensor, reference 'model output' memory layout ([T,B,J,F] order)"""
    return t.permute(3, 0, 1, 2).contiguous().permute(1, 2, 3, 0)


@pytest.mark.parametrize("shape", [(3, 56, 6, 60), (2, 263, 1, 37), (1, 5, 3, 1)])
@pytest.mark.parametrize("layout", ["bjft", "tbi"])
@pytest.mark.parametrize("clip", [False, True])
def test_p_sample_update_matches_oracle(built_lib, shape, layout, clip):
    d = _diffusion()
    smp = sampler_ref.Sampler(timestep_respacing="ddim20")
    g = torch.Generator().manual_seed(1)
    x, x0, noise = (torch.randn(shape, generator=g) for _ in range(3))
    x0 = x0 * 1.5
    t = torch.tensor([0, 7, 19][:shape[0]])
    want, want_x0 = smp.p_sample(lambda xx, tt: x0, x, t, lambda xx: noise, clip_denoised=clip)
    xc, x0c, nc = x.cuda(), x0.cuda(), noise.cuda()
    if layout == "tbi":
        x0c = _perm(x0c)
    got, pred = d._update("p", xc, x0c, nc, t.cuda(), clip)
    # identical fp32 operation order; only expf may differ by an ulp between libm and CUDA
    assert torch.allclose(got.cpu(), want, rtol=0, atol=2e-6)
    assert torch.equal(pred.cpu(), want_x0)
    assert (gd._layout_of(got) == layout) or shape[3] == 1


@pytest.mark.parametrize("eta", [0.0, 0.7])
@pytest.mark.parametrize("shape", [(3, 56, 6, 60), (2, 263, 1, 37)])
def test_ddim_update_matches_oracle(built_lib, eta, shape):
    d = _diffusion("ddim10")
    smp = sampler_ref.Sampler(timestep_respacing="ddim10")
    g = torch.Generator().manual_seed(2)
    x, x0, noise = (torch.randn(shape, generator=g) for _ in range(3))
    t = torch.tensor([0, 4, 9][:shape[0]])
    want, _ = smp.ddim_sample(lambda xx, tt: x0, x, t, lambda xx: noise, eta=eta)
    got, _ = d._update("ddim", x.cuda(), _perm(x0.cuda()), noise.cuda(), t.cuda(), False, eta=eta)
    assert torch.allclose(got.cpu(), want, rtol=0, atol=3e-6)


def test_mean_only_and_p_mean_variance_dict(built_lib):
    d = _diffusion()
    g = torch.Generator().manual_seed(3)
    x, x0 = torch.randn(2, 56, 6, 60, generator=g), torch.randn(2, 56, 6, 60, generator=g)
    t = torch.tensor([3, 11])
    model = lambda xx, tt, **kw: x0.cuda()  # noqa: E731
    out = d.p_mean_variance(model, x.cuda(), t.cuda(), clip_denoised=False, model_kwargs={"y": {}})
    smp = sampler_ref.Sampler(timestep_respacing="ddim20")
    mean, logvar, _ = smp.p_mean_variance(lambda xx, tt: x0, x, t)
    assert torch.allclose(out["mean"].cpu(), mean, rtol=0, atol=1e-6)
    assert torch.equal(out["log_variance"].cpu(), logvar.expand_as(x))
    assert set(out) == {"mean", "variance", "log_variance", "pred_xstart"}


def test_generic_route_loop_with_foreign_model_matches_oracle(built_lib):
    """Any callable model goes through the generic route; RNG handling must equal the reference's."""
    d = _diffusion("ddim20")

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x, t, y=None):
            # permuted output like the reference's OutputProcess
            return (0.5 * x + t.view(-1, 1, 1, 1).float() * 1e-3).permute(3, 0, 1, 2).contiguous().permute(1, 2, 3, 0)

    m = Toy().cuda()
    shape = (2, 7, 6, 9)
    noises = []
    orig = torch.randn_like

    def rec(x, **kw):
        n = orig(x, **kw)
        noises.append(n.cpu())
        return n

    torch.manual_seed(5)
    init = torch.randn(*shape, device="cuda")
    torch.randn_like = rec
    try:
        got = d.p_sample_loop(m, shape, noise=init, clip_denoised=False, model_kwargs={"y": {}})
    finally:
        torch.randn_like = orig
    assert len(noises) == 20
    # from step 2 on the noise must have been drawn in the permuted layout (reference behaviour)
    assert gd._layout_of(noises[0]) == "bjft" and gd._layout_of(noises[1]) == "tbi"
    smp = sampler_ref.Sampler(timestep_respacing="ddim20")
    it = iter(noises)
    cpu_model = lambda x, t: 0.5 * x + t.view(-1, 1, 1, 1).float() * 1e-3  # noqa: E731
    want, _ = smp.loop(cpu_model, shape, noise_fn=lambda x: next(it), init_noise=init.cpu())
    assert torch.allclose(got.cpu(), want, rtol=0, atol=1e-5)


def test_cfg_combine_matches_oracle(built_lib):
    from regennet_b200 import _lib
    g = torch.Generator().manual_seed(4)
    c, u = torch.randn(3, 56, 6, 60, generator=g), torch.randn(3, 56, 6, 60, generator=g)
    s = torch.tensor([2.5, 0.0, -1.0])
    out = torch.empty_like(c, device="cuda")
    cc, uc, sc = c.cuda(), u.cuda(), s.cuda()
    _lib.check(built_lib.regen_cfg_combine(_lib.ptr(cc), _lib.ptr(uc), _lib.ptr(sc), _lib.ptr(out), c.numel(),
                                           c[0].numel(), 3, _lib.stream_ptr()), "cfg")
    assert torch.equal(out.cpu(), u + s.view(-1, 1, 1, 1) * (c - u))


def test_rot6d_matches_reference_golden_and_oracle(built_lib):
    g = np.load(os.path.join(HERE, "rot6d.npz"))
    d6 = torch.from_numpy(g["d6"])
    R = rotation_conversions.rotation_6d_to_matrix(d6.cuda()).cpu()
    assert R.shape == d6.shape[:-1] + (3, 3)
    assert np.allclose(R.numpy(), g["R"], atol=1e-6, equal_nan=True)
    # larger, ragged size (not a multiple of the 256-rotation block) against the oracle
    gen = torch.Generator().manual_seed(7)
    big = torch.randn(100003, 6, generator=gen)
    got = rotation_conversions.rotation_6d_to_matrix(big.cuda()).cpu()
    want = sampler_ref.rotation_6d_to_matrix(big)
    # Gram-Schmidt amplifies rounding when a2 is nearly parallel to a1; compare both against float64
    ref64 = sampler_ref.rotation_6d_to_matrix(big.double())
    err_gpu = (got.double() - ref64).abs().max().item()
    err_cpu = (want.double() - ref64).abs().max().item()
    print("rot6d max abs err vs fp64: gpu %.3e, cpu oracle %.3e" % (err_gpu, err_cpu))
    assert err_gpu < max(2.0 * err_cpu, 1e-5)
    assert torch.allclose(got, want, atol=1e-4)
    # size-independent property: orthonormal rows with det +1
    eye = got @ got.transpose(-1, -2)
    assert torch.allclose(eye, torch.eye(3).expand_as(eye), atol=1e-4)
    assert torch.allclose(torch.linalg.det(got), torch.ones(big.shape[0]), atol=1e-4)
    # empty input
    assert rotation_conversions.rotation_6d_to_matrix(torch.empty(0, 6, device="cuda")).shape == (0, 3, 3)


def test_layout_round_trip(built_lib):
    g = torch.Generator().manual_seed(8)
    for shape in [(3, 56, 6, 60), (2, 263, 1, 196), (1, 1, 1, 1), (5, 33, 1, 31)]:
        x = torch.randn(shape, generator=g).cuda()
        p = gd._to_layout(x, "tbi")
        assert gd._layout_of(p) in ("tbi", "bjft") and torch.equal(p, x)
        assert torch.equal(p.permute(3, 0, 1, 2).contiguous(), x.permute(3, 0, 1, 2).contiguous())
        back = gd._to_layout(p, "bjft")
        assert back.is_contiguous() and torch.equal(back, x)


def test_cpu_tensors_are_rejected(built_lib):
    d = _diffusion()
    x = torch.zeros(1, 2, 3, 4)
    with pytest.raises(RuntimeError, match="no CPU route"):
        d._update("p", x, x, x, torch.zeros(1, dtype=torch.long), False)

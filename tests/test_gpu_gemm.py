"""GPU parity: the tcgen05/TMA GEMM kernel (through the C-ABI test hook) vs float64 matmul."""
import math

import pytest
import torch

from regennet_b200 import _lib

pytestmark = pytest.mark.gpu

# bf16x3 keeps ~16 significand bits per operand: |err| ~ 2^-16 * sum|a||w| / sqrt(K) -> < 1e-4 for O(1) outputs.
TOL = {0: 1e-4, 1: 6e-2}

SHAPES = [
    (128, 256, 64),       # exactly one tile, one k-block
    (128, 256, 512),      # k loop wraps the smem ring
    (300, 512, 1024),     # ragged M, two n tiles, K=1024 (linear2)
    (60, 1536, 512),      # B=1, T=60 (config 1): M smaller than a tile
    (2000, 336, 512),     # output projection, N not a multiple of the tile
    (257, 263, 512),      # hml_vec output projection: N % 4 != 0 (scalar epilogue)
    (384, 512, 336),      # input projection: K padded 336 -> 384
    (130, 1024, 263),     # K = 263 padded to 320
]


def _run(lib, A, W, bias, res, gelu, precision):
    M, K = A.shape
    N = W.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.regen_test_gemm(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), M, N, K,
                             int(gelu), precision, _lib.stream_ptr())
    _lib.check(rc, "regen_test_gemm")
    return out


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp64(built_lib, M, N, K, precision):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda() * 0.1
    out = _run(built_lib, A, W, bias, None, False, precision)
    want = (A.double() @ W.double().t() + bias.double())
    err = (out.double() - want).abs().max().item()
    print("M=%d N=%d K=%d precision=%d max abs err %.3e" % (M, N, K, precision, err))
    assert not torch.isnan(out).any()
    assert err < TOL[precision]


def test_gemm_epilogue_residual_gelu(built_lib):
    g = torch.Generator().manual_seed(11)
    M, N, K = 200, 512, 512
    A = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)).cuda()
    bias = (torch.randn(N, generator=g) * 0.1).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    out = _run(built_lib, A, W, bias, res, True, 0)
    pre = A.double() @ W.double().t() + bias.double() + res.double()
    want = 0.5 * pre * (1 + torch.erf(pre / math.sqrt(2.0)))
    assert (out.double() - want).abs().max().item() < 1e-4


def test_gemm_linearity_at_full_size(built_lib):
    """Size-independent property at config-2 scale (M = 256*60): G(a*A1 + A2) == a*G(A1) + G(A2)."""
    g = torch.Generator().manual_seed(12)
    M, N, K = 15360, 1536, 512
    A1 = torch.randn(M, K, generator=g).cuda()
    A2 = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)).cuda()
    o1 = _run(built_lib, A1, W, None, None, False, 0)
    o2 = _run(built_lib, A2, W, None, None, False, 0)
    o3 = _run(built_lib, 2.0 * A1 + A2, W, None, None, False, 0)
    assert (o3 - (2.0 * o1 + o2)).abs().max().item() < 3e-4
    # spot-check rows against fp64
    rows = torch.tensor([0, 127, 128, 7777, 15359], device="cuda")
    want = A1[rows].double() @ W.double().t()
    assert (o1[rows].double() - want).abs().max().item() < 1e-4


# precision 2 = the mixed8 main loop of the pair kernel (fp16 MMA + two e4m3 correction MMAs per product; the test hook packs
# the operands from fp32): M > 128, K % 128 == 0.  Tail slices, ragged M / N, many tiles per cluster, the GELU epilogue.
M8_SHAPES = [(256, 256, 128), (300, 512, 1024), (2000, 336, 512), (15360, 1536, 512), (15360, 1024, 512)]


@pytest.mark.parametrize("M,N,K", M8_SHAPES)
def test_gemm_mixed8_matches_fp64(built_lib, M, N, K):
    g = torch.Generator().manual_seed(M * 5 + N * 3 + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda() * 0.1
    out = _run(built_lib, A, W, bias, None, False, 2)
    rows = torch.arange(M, device="cuda") if M <= 2000 else torch.tensor([0, 1, 127, 128, 255, 256, 7777, M - 1], device="cuda")
    want = A[rows].double() @ W.double().t() + bias.double()
    err = (out[rows].double() - want).abs().max().item()
    print("mixed8 M=%d N=%d K=%d max abs err %.3e" % (M, N, K, err))
    assert not torch.isnan(out).any()
    assert err < 2e-4


def test_gemm_mixed8_gelu_and_determinism(built_lib):
    g = torch.Generator().manual_seed(13)
    M, N, K = 15360, 1024, 512
    A = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)).cuda()
    bias = (torch.randn(N, generator=g) * 0.1).cuda()
    out = _run(built_lib, A, W, bias, None, True, 2)
    rows = torch.tensor([0, 300, 9999, M - 1], device="cuda")
    pre = A[rows].double() @ W.double().t() + bias.double()
    want = 0.5 * pre * (1 + torch.erf(pre / math.sqrt(2.0)))
    assert (out[rows].double() - want).abs().max().item() < 2e-4
    for _ in range(20):   # pipeline-hazard guard: same bits every time
        assert torch.equal(_run(built_lib, A, W, bias, None, True, 2), out)

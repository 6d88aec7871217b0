"""Host-side check of the mixed8 operand arithmetic: the exact operand formats the sm_100a kernels
use -- fp16 main operands, e4m3 correction operands with the power-of-two scales of gemm_ln_sm100.cuh / layers.cuh
(activations: residual * 2^9, hi * 2^-2; weights: hi * 2^6, residual * 2^17; both products scaled by 2^15 and folded
in by D = A.B + D * 2^-15) -- emulated with torch's float8_e4m3fn / float16 on the CPU against float64.  It pins the
choice of scales (ranges, saturation behaviour) independently of the GPU; the GPU parity tests are in test_gpu_mixed8.py."""
import math

import pytest
import torch


def _e4m3(v):
    return v.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()   # cvt.rn.satfinite.e4m3x2.f32


def _f16(v):
    return v.clamp(-65504.0, 65504.0).to(torch.float16).float()      # cvt.rn.satfinite.f16x2.f32


def mixed8_matmul(a, w):
    """a [M, K] activations, w [N, K] weights -> a @ w.T the way the mixed8 kernels compute it (fp32 accumulation)."""
    a16, w16 = _f16(a), _f16(w)
    corr = _e4m3((a - a16) * 2.0 ** 9) @ _e4m3(w16 * 2.0 ** 6).t() + _e4m3(a16 * 2.0 ** -2) @ _e4m3((w - w16) * 2.0 ** 17).t()
    return a16 @ w16.t() + corr * 2.0 ** -15


def fp16_x1(a, w):
    return _f16(a) @ _f16(w).t()


def _case(seed, M, N, K, act_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g) * act_scale
    a = torch.nn.functional.gelu(a)                     # FFN activations: many small negatives, a positive tail
    w = (torch.rand(N, K, generator=g) * 2 - 1) / math.sqrt(K)
    return a, w, a.double() @ w.double().t()


@pytest.mark.parametrize("K", [512, 1024])
def test_mixed8_matches_fp64_at_bf16x3_class_accuracy(K):
    a, w, want = _case(K, 192, 128, K)
    err = (mixed8_matmul(a, w).double() - want).abs().max().item()
    err1 = (fp16_x1(a, w).double() - want).abs().max().item()
    print("K=%d: mixed8 %.3e, fp16 single %.3e (|result| max %.2f)" % (K, err, err1, want.abs().max()))
    assert err < 5e-5            # the GPU kernels measure 2.5e-5 ... 4e-5 on such products
    assert err < err1 / 5        # the two e4m3 correction products are what buys the accuracy


def test_large_activations_inside_the_exact_range_keep_their_corrections():
    """|a| up to ~1700: above e4m3's 448, still exact for the correction operands thanks to the 2^-2 / 2^9 scales."""
    a, w, want = _case(7, 128, 64, 512)
    a[:, :8] *= 400.0                                    # outlier channels, |a| up to ~1700
    want = a.double() @ w.double().t()
    assert 448.0 < a.abs().max().item() < 1792.0
    rel = ((mixed8_matmul(a, w).double() - want).abs().max() / want.abs().max()).item()
    rel1 = ((fp16_x1(a, w).double() - want).abs().max() / want.abs().max()).item()
    print("outlier channels: mixed8 rel %.3e, fp16 single rel %.3e" % (rel, rel1))
    assert rel < 4e-5 and rel < rel1 / 5


def test_beyond_the_range_degrades_to_fp16_accuracy_not_worse():
    a, w, _ = _case(9, 128, 64, 512)
    a[:, :4] *= 6000.0                                   # |a| ~ 2e4: the fp8 copies saturate
    want = a.double() @ w.double().t()
    assert a.abs().max().item() > 1792.0
    rel = ((mixed8_matmul(a, w).double() - want).abs().max() / want.abs().max()).item()
    rel1 = ((fp16_x1(a, w).double() - want).abs().max() / want.abs().max()).item()
    print("saturated: mixed8 rel %.3e, fp16 single rel %.3e" % (rel, rel1))
    assert math.isfinite(rel) and rel < 3 * rel1 + 1e-6


# ---------------------------------------------------------------- precision 'mixed8h': the residual stream in that pack
def h8_roundtrip(h):
    """What gemm_ln_kernel<.., H8> stores for the residual stream h and what its next residual load rebuilds:
    fp16(h) + e4m3((h - fp16(h)) * 2^9) * 2^-9  (the third byte, e4m3(fp16(h) / 4), only feeds the correction MMAs)."""
    h16 = _f16(h)
    return h16 + _e4m3((h - h16) * 2.0 ** 9) * 2.0 ** -9


def test_h8_residual_stream_keeps_about_15_significand_bits():
    g = torch.Generator().manual_seed(3)
    h = torch.randn(4096, 512, generator=g) * 1.5          # LayerNorm outputs: O(1) with a tail of a few units
    back = h8_roundtrip(h)
    err = (back - h).abs()
    # residual <= 2^-11 |h| (half an fp16 ulp), kept to 4 significand bits by e4m3 -> 2^-15 |h|; below |h| ~ 1/16 the scaled
    # residual falls into e4m3's subnormals (step 2^-9): absolute error <= 2^-10 * 2^-9
    bound = torch.maximum(h.abs() * 2.0 ** -15, torch.tensor(2.0 ** -19))
    bf16_pair = h.to(torch.bfloat16).float()
    bf16_pair = bf16_pair + (h - bf16_pair).to(torch.bfloat16).float()
    print("h8 round trip: max abs err %.3e (bf16 pair: %.3e), max err / bound %.3f" % (
        err.max(), (bf16_pair - h).abs().max(), (err / bound).max()))
    assert (err <= bound).all()
    assert err.max().item() < 2e-4 * 1.5


def test_h8_subnormal_and_large_values_degrade_gracefully():
    h = torch.tensor([0.0, 1e-6, -3e-5, 0.37, -5.0, 250.0, 1500.0, 40000.0, 1e6])
    back = h8_roundtrip(h)
    assert torch.isfinite(back).all()
    assert back[0].item() == 0.0
    assert (back[:7] - h[:7]).abs().max().item() <= 2.0 ** -13 * 1500.0 / 2   # residual of 1500: (ulp 1) * 2^-5
    assert abs(back[7].item() - 40000.0) <= 16.0 + 1.0       # fp16 ulp 32 at 4e4; the scaled residual (<= 16 * 2^9) saturates at 448
    assert back[8].item() == pytest.approx(65504.0 + 448.0 / 512.0)   # satfinite fp16, saturated residual byte


def test_h8_operands_give_the_same_product_accuracy():
    """QKV / FFN1 under mixed8h: A = the stored pack of h (fp16 | residual byte | hi byte), i.e. mixed8_matmul on h itself."""
    g = torch.Generator().manual_seed(11)
    h = torch.randn(256, 512, generator=g)
    w = (torch.rand(384, 512, generator=g) * 2 - 1) / math.sqrt(512)
    want = h.double() @ w.double().t()
    err = (mixed8_matmul(h, w).double() - want).abs().max().item()
    print("mixed8h QKV-shaped product: %.3e" % err)
    assert err < 5e-5

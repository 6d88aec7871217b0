"""GPU parity: the multi-chunk compact attention kernel (64 < T <= 256, 64-query blocks, 64-key chunks, two CTAs per SM;
C-ABI test hook flag 8) vs float64, causal (arch 'online') and unmasked (arch 'offline', flag 2)."""
import math

import pytest
import torch

from test_gpu_attention import ref_attention, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T", [(2, 65), (2, 128), (3, 129), (3, 150), (2, 192), (2, 196), (1, 256), (64, 150)])
def test_mc_attention_causal_matches_fp64(built_lib, B, T):
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    out = run(built_lib, qkv, B, T, dbg=8)
    want = ref_attention(qkv, B, T)
    err = (out.double() - want).abs().max().item()
    print("MC causal B=%d T=%d max abs err %.3e" % (B, T, err))
    assert not torch.isnan(out).any()
    assert err < 1e-4


@pytest.mark.parametrize("B,T", [(2, 65), (2, 151), (1, 197), (2, 256)])
def test_mc_attention_unmasked_matches_fp64(built_lib, B, T):
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    out = run(built_lib, qkv, B, T, dbg=8 | 2)
    x = qkv.double().view(T, B, 3, 4, 128)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    p = torch.softmax(torch.einsum("ibhd,jbhd->bhij", q, k) / math.sqrt(128.0), dim=-1)
    want = torch.einsum("bhij,jbhd->ibhd", p, v).reshape(T * B, 512)
    err = (out.double() - want).abs().max().item()
    print("MC unmasked B=%d T=%d max abs err %.3e" % (B, T, err))
    assert not torch.isnan(out).any()
    assert err < 1e-4


def test_mc_attention_equals_chunk128_kernel(built_lib):
    """Same inputs through the 128-key-chunk kernel and the multi-chunk compact kernel: both are bf16x3 with an exact row
    maximum, so they agree to accumulation-order noise."""
    g = torch.Generator().manual_seed(77)
    B, T = 8, 150
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    qkv[:, :1024] *= 2.0   # peaked softmax
    a = run(built_lib, qkv, B, T, dbg=0)
    b = run(built_lib, qkv, B, T, dbg=8)
    err = (a - b).abs().max().item()
    print("MC vs chunk-128 kernel: max abs diff %.3e" % err)
    assert err < 1e-4

"""GPU parity: the tcgen05 causal attention kernel (C-ABI test hook) vs a float64 restatement of
nn.MultiheadAttention's scaled dot-product with the causal mask of model/cmdm.py:168-171."""
import math

import pytest
import torch

from regennet_b200 import _lib

pytestmark = pytest.mark.gpu


def ref_attention(qkv, B, T):
    """qkv [T*B, 1536] (row = t*B + b) -> [T*B, 512], float64."""
    x = qkv.double().view(T, B, 3, 4, 128)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]                   # [T, B, H, hd]
    s = torch.einsum("ibhd,jbhd->bhij", q, k) / math.sqrt(128.0)
    mask = torch.triu(torch.full((T, T), float("-inf"), dtype=torch.float64, device=qkv.device), diagonal=1)
    p = torch.softmax(s + mask, dim=-1)
    o = torch.einsum("bhij,jbhd->ibhd", p, v)
    return o.reshape(T * B, 512)


def test_long_sequence_non_causal(built_lib):
    """dbg bit 1 = no mask (arch 'offline') on the long-sequence kernel."""
    B, T = 2, 301
    qkv = torch.randn(T * B, 1536, generator=torch.Generator().manual_seed(9)).cuda()
    out = run(built_lib, qkv, B, T, dbg=2)
    x = qkv.double().view(T, B, 3, 4, 128)
    s_ = torch.einsum("ibhd,jbhd->bhij", x[:, :, 0], x[:, :, 1]) / math.sqrt(128.0)
    want = torch.einsum("bhij,jbhd->ibhd", torch.softmax(s_, dim=-1), x[:, :, 2]).reshape(T * B, 512)
    assert (out.double() - want).abs().max().item() < 1e-4


def run(lib, qkv, B, T, dbg=0):
    out = torch.full((T * B, 512), float("nan"), device="cuda")
    _lib.check(lib.regen_test_attention(_lib.ptr(qkv), _lib.ptr(out), B, T, dbg, _lib.stream_ptr()), "test_attention")
    return out


# T > 256: the streaming CUDA-core kernel (attention_long_kernel) -- the tcgen05 kernels keep a key row's scores in TMEM
@pytest.mark.parametrize("B,T", [(1, 1), (2, 7), (3, 60), (2, 64), (2, 65), (2, 128), (3, 150), (2, 196), (1, 256),
                                 (2, 257), (1, 300), (2, 513)])
def test_attention_matches_fp64(built_lib, B, T):
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    out = run(built_lib, qkv, B, T)
    want = ref_attention(qkv, B, T)
    err = (out.double() - want).abs().max().item()
    print("B=%d T=%d max abs err %.3e" % (B, T, err))
    assert not torch.isnan(out).any()
    assert err < 1e-4


def test_attention_large_scores_and_full_batch(built_lib):
    """Peaked softmax (large |q.k|) and the config-2 batch (B=256, T=60); spot-check rows in fp64."""
    g = torch.Generator().manual_seed(5)
    B, T = 256, 60
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    qkv[:, :1024] *= 2.0
    out = run(built_lib, qkv, B, T)
    want = ref_attention(qkv, B, T)
    err = (out.double() - want).abs().max().item()
    print("B=256 T=60 (peaked) max abs err %.3e" % err)
    assert err < 3e-4

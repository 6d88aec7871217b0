"""Bring-up tool (GPU box): per-tile pipeline timeline of the CTA-pair GEMM kernel for the layer shapes."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regennet_b200 import _lib

lib = _lib.lib()
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
M = 15360
PREC = int(os.environ.get("GEMM_TIMELINE_PRECISION", "0"))   # 2 = mixed8 main loop (shapes without residual only)
for name, N, K, res, gelu in [("qkv", 1536, 512, False, False), ("out_proj", 512, 512, True, False),
                              ("ffn1", 1024, 512, False, True), ("ffn2", 512, 1024, True, False)]:
    if PREC == 2 and res:
        continue
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / math.sqrt(K)
    b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda")
    for rep in range(3):
        tl.zero_()
        lib.regen_test_gemm_timeline(_lib.ptr(tl))
        _lib.check(lib.regen_test_gemm(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(R), _lib.ptr(out), M, N, K,
                                       int(gelu), PREC, _lib.stream_ptr()), "gemm")
    t = tl.cpu().tolist()
    t0 = t[0]
    print("== %s  N=%d K=%d: setup %d, total %d cycles" % (name, N, K, t[1] - t0, t[2] - t0))
    for i in range(16):
        if t[8 + 2 * i] == 0:
            break
        print("  tile %d: mma start %7d issue-end %7d | epi start %7d end %7d (epi %6d)" % (
            i, t[8 + 2 * i] - t0, t[9 + 2 * i] - t0, t[40 + 2 * i] - t0, t[41 + 2 * i] - t0, t[41 + 2 * i] - t[40 + 2 * i]))
    for i in range(3):
        kb = [x - t0 for x in t[104 + 8 * i:112 + 8 * i] if x]
        print("  tile %d k-block arrivals %s cadence %s" % (i, kb, [b - a for a, b in zip(kb, kb[1:])]))
    for sc in range(3):
        b0 = 80 + sc * 8
        if t[b0]:
            print("  sub-chunk %d: ld-issue 0 | res staged %d | tmem ready %d | bias+res added %d | gelu %d | staged %d | stored %d" % (
                sc, t[b0 + 1] - t[b0] if t[b0 + 1] else -1, t[b0 + 2] - t[b0], t[b0 + 3] - t[b0], t[b0 + 4] - t[b0],
                t[b0 + 5] - t[b0], t[b0 + 6] - t[b0]))
lib.regen_test_gemm_timeline(None)

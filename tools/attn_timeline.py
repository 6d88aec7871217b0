"""Bring-up tool (GPU box): event timeline of one attention CTA (test hook, B=256, T=60)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regennet_b200 import _lib

lib = _lib.lib()
B, T = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 60
qkv = torch.randn(T * B, 1536, device="cuda")
out = torch.empty(T * B, 512, device="cuda")
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    tl.zero_()
    lib.regen_test_gemm_timeline(_lib.ptr(tl))
    _lib.check(lib.regen_test_attention(_lib.ptr(qkv), _lib.ptr(out), B, T, 0, _lib.stream_ptr()), "attn")
lib.regen_test_gemm_timeline(None)
v = tl.cpu().tolist()
names = ["entry", "setup done", "Q,K landed", "S done", "softmax max done", "P written", "PV issue (P,V ready)",
         "O ready", "epilogue done", "exit"]
for k, n in enumerate(names):
    print("%-22s %7d" % (n, v[k] - v[0]))

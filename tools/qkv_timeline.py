"""Bring-up tool (GPU box): pipeline stamps of the QKV GEMM (last layer) inside a real B=256 forward: per tile and per k-block.
    REGEN_DEBUG_QKV_TIMELINE=1 python tools/qkv_timeline.py"""
import os
import sys
os.environ["REGEN_DEBUG_QKV_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import cases
from regennet_b200 import _lib, synthetic
from regennet_b200.cmdm import CMDM

lib = _lib.lib()
m = CMDM(precision=os.environ.get("REGEN_PRECISION", "mixed8"), **cases.MODELS["ntu"])
m.load_state_dict(synthetic.make_state_dict(seed=0, **cases.synth_kw("ntu")), strict=False)
m = m.cuda().eval()
x, y = synthetic.make_inputs(256, 56, 6, 60, seed=1)
xc, yc = x.cuda(), {"cmotion": y["cmotion"].cuda()}
t = torch.full((256,), 500, dtype=torch.long, device="cuda")
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
for rep in range(3):
    tl.zero_()
    lib.regen_test_gemm_timeline(_lib.ptr(tl))
    with torch.no_grad():
        m(xc, t, yc)
    torch.cuda.synchronize()
lib.regen_test_gemm_timeline(None)
v = tl.cpu().tolist()
t0 = v[0]
print("== QKV (layer 8) inside a forward: setup %d, total %d cycles" % (v[1] - t0, v[2] - t0))
for i in range(16):
    if v[8 + 2 * i] == 0:
        break
    print("  tile %d: mma start %7d issue-end %7d | epi start %7d end %7d (epi %6d)" % (
        i, v[8 + 2 * i] - t0, v[9 + 2 * i] - t0, v[40 + 2 * i] - t0, v[41 + 2 * i] - t0, v[41 + 2 * i] - v[40 + 2 * i]))
for i in range(3):
    kb = [a - t0 for a in v[104 + 8 * i:112 + 8 * i] if a]
    print("  tile %d k-block arrivals %s cadence %s" % (i, kb, [b - a for a, b in zip(kb, kb[1:])]))

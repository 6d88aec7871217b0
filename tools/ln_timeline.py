"""Bring-up tool (GPU box): pass-level timeline of the fused GEMM+LayerNorm kernel inside a real denoise call."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import cases
from regennet_b200 import _lib, synthetic
from regennet_b200.cmdm import CMDM

lib = _lib.lib()
m = CMDM(precision=os.environ.get("REGEN_PRECISION", "bf16x3"), **cases.MODELS["ntu"])
m.load_state_dict(synthetic.make_state_dict(seed=0, **cases.synth_kw("ntu")), strict=False)
m = m.cuda().eval()
NB = int(os.environ.get("LN_TIMELINE_B", "256"))
x, y = synthetic.make_inputs(NB, 56, 6, 60, seed=1)
xc, yc = x.cuda(), {"cmotion": y["cmotion"].cuda()}
t = torch.full((NB,), 500, dtype=torch.long, device="cuda")
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
for rep in range(3):
    tl.zero_()
    lib.regen_test_gemm_timeline(_lib.ptr(tl))
    with torch.no_grad():
        m(xc, t, yc)
    torch.cuda.synchronize()
lib.regen_test_gemm_timeline(None)
v = tl.cpu().tolist()
names = ["entry", "mma first operands", "mma issue end", "epi: acc ready", "pass1 done", "exchange1 done",
         "pass2 done", "exchange2 done", "final pass done", "stores drained", "exit",
         "p3 sc2: enter", "p3 sc2: buffer free", "p3 sc2: computed+staged", "p3 sc2: fenced", "p3 sc2: stores issued",
         "p3 sc3: tmem ready"]
for base, label in [(0, "out_proj + LN1 + c + LN2 (K=512)"), (20, "linear2 + LN3 (K=1024)")]:
    print("==", label)
    t0 = v[base]
    for k, n in enumerate(names):
        if v[base + k]:
            print("   %-22s %8d" % (n, v[base + k] - t0))
st = [x - v[20] for x in v[48:88] if x]
print("== linear2 main loop: stage arrival times (cycles after entry) and cadence")
print("   ", st)
print("   ", [b - a for a, b in zip(st, st[1:])])

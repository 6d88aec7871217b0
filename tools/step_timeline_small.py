"""Bring-up tool (GPU box): GPU-side timeline of one denoising step on the small-batch route (B given, default 1):
spans of the tcgen05 kernels (GEMMs, attention) and the gaps between them (LayerNorm / elementwise kernels + launch gaps)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import bench
from regennet_b200 import _lib, synthetic

dev = torch.device("cuda")
B, T, U = (int(sys.argv[1]) if len(sys.argv) > 1 else 1), 60, 2
model, mkdiff = bench.build_ours(dev)
d = mkdiff([1000])
_, y = synthetic.make_inputs(B, 56, 6, T, seed=10)
yc = {"cmotion": y["cmotion"].to(dev)}
shape = (B, 56, 6, T)
img = torch.randn(*shape, device=dev)
sess = d._fast_session(model, shape, {"y": yc}, None, None, False, False, img)
lib = _lib.lib()
CAP = 256
log = torch.zeros(2 * CAP + 160 * CAP, dtype=torch.int64, device=dev)
gen = sess.run(d, "p", img, list(range(1000))[::-1], False, 0.0, graph=True, unroll=U)
next(gen)
_lib.check(lib.regen_test_step_log(model._handle.ptr, _lib.ptr(log), CAP), "step_log")
for _ in range(4):
    next(gen)
torch.cuda.synchronize()
init = torch.zeros(2 * CAP + 160 * CAP, dtype=torch.int64)
init[0:2 * CAP:2] = torch.iinfo(torch.int64).max
log.copy_(init)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
next(gen)
e1.record()
torch.cuda.synchronize()
_lib.check(lib.regen_test_step_log(model._handle.ptr, None, 0), "off")
v = log.cpu()[:2 * CAP].view(CAP, 2)
n = int((v[:, 1] > 0).sum())
print("B=%d: graph of %d steps: %.3f ms per step by events; %d kernel records" % (B, U, e0.elapsed_time(e1) / U, n))
names = ["in_proj"] + [k for l in range(8) for k in ("qkv", "attn", "out", "ffn1", "lin2")] + ["out_proj"]
per = len(names)
agg = {}
t0 = int(v[0, 0])
for i in range(n):
    s, e = int(v[i, 0]), int(v[i, 1])
    gap = s - int(v[i - 1, 1]) if i else 0
    a = agg.setdefault(names[i % per], [0, 0, 0])
    a[0] += 1; a[1] += e - s; a[2] += gap
    if i < 8:
        print("  %-9s start %8.2f us span %6.2f us gap before %6.2f us" % (names[i % per], (s - t0) / 1e3, (e - s) / 1e3, gap / 1e3))
for nm, (c, sp, gp) in agg.items():
    print("  %-9s n=%3d  span %6.2f us  gap before %6.2f us" % (nm, c, sp / c / 1e3, gp / c / 1e3))
print("sum of spans %.1f us/step, sum of gaps %.1f us/step" % (sum(a[1] for a in agg.values()) / U / 1e3,
                                                                sum(a[2] for a in agg.values()) / U / 1e3))
gen.close()

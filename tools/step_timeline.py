"""Bring-up tool (GPU box): GPU-side timeline of whole denoising steps inside a replayed CUDA graph.

Every tcgen05 kernel logs (earliest inputs-available time after griddepcontrol.wait, latest CTA exit) from the GPU's
global nanosecond timer (regen_test_step_log), so kernel spans and the gaps BETWEEN kernels can be read without nsys.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import bench
from regennet_b200 import _lib, synthetic

dev = torch.device("cuda")
B, T, U = 256, 60, 2
model, mkdiff = bench.build_ours(dev, B, T)
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
next(gen)                                      # first step (host-enqueued); creates the handle
_lib.check(lib.regen_test_step_log(model._handle.ptr, _lib.ptr(log), CAP), "step_log")
next(gen)                                      # capture (slots are baked into the graph) + first replay
for _ in range(3):
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
_lib.check(lib.regen_test_step_log(model._handle.ptr, None, 0), "step_log off")
full = log.cpu()
v = full[:2 * CAP].view(CAP, 2)
cta = full[2 * CAP:].view(CAP, 160)
n = int((v[:, 1] > 0).sum())
names = ["in_proj"] + [k for l in range(8) for k in ("qkv", "attn", "out+LN", "ffn1", "lin2+LN")] + ["out_proj"]
per = len(names)
print("graph of %d steps: %.3f ms per step by events; %d kernel records" % (U, e0.elapsed_time(e1) / U, n))
t0 = int(v[0, 0])
tot_span = tot_gap = 0
agg = {}
for i in range(n):
    s, e = int(v[i, 0]), int(v[i, 1])
    gap = s - int(v[i - 1, 1]) if i else 0
    nm = names[i % per]
    a = agg.setdefault(nm.split("+")[0] if False else nm, [0, 0, 0])
    a[0] += 1; a[1] += e - s; a[2] += gap
    if i < per + 2:
        print("  %-9s start %9.2f us  span %7.2f us  gap before %6.2f us" % (nm, (s - t0) / 1e3, (e - s) / 1e3, gap / 1e3))
    tot_span += e - s
    tot_gap += gap
print("per kernel type (avg over %d steps): " % U)
for nm, (c, sp, gp) in agg.items():
    print("  %-9s n=%3d  span %7.2f us  gap before %6.2f us" % (nm, c, sp / c / 1e3, gp / c / 1e3))
print("sum of spans %.3f ms/step, sum of gaps %.3f ms/step (gaps include the small elementwise kernels + randn)" % (
    tot_span / U / 1e6, tot_gap / U / 1e6))

# per-CTA exit times (relative to the kernel's inputs-available time) for one QKV, FFN1 and fused-LN launch
for slot in (1 + 5, 1 + 5 + 3, 1 + 5 + 2, 1 + 5 + 4):
    nm = names[slot % per]
    rel = [(int(c) - int(v[slot, 0])) / 1e3 for c in cta[slot] if int(c) > 0]
    pairs = [max(rel[i], rel[i + 1]) for i in range(0, len(rel) - 1, 2)]
    srt = sorted(pairs)
    print("%-8s %d clusters: exit time us min %.1f p25 %.1f median %.1f p75 %.1f max %.1f | first 20 clusters: %s" % (
        nm, len(pairs), srt[0], srt[len(srt) // 4], srt[len(srt) // 2], srt[3 * len(srt) // 4], srt[-1],
        " ".join("%.0f" % x for x in pairs[:20])))
    print("         last 12 clusters: %s" % " ".join("%.0f" % x for x in pairs[-12:]))
gen.close()

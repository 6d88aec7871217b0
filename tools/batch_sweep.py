"""GPU box: device-resident step time (graph driver) of the NTU online model over batch sizes, fused vs separate LayerNorm."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import bench
from regennet_b200 import synthetic

dev = torch.device("cuda")
T = 60
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128, 256]:
    model, mkdiff = bench.build_ours(dev, B, T)
    d = mkdiff([1000])
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=10)
    yc = {"cmotion": y["cmotion"].to(dev)}
    shape = (B, 56, 6, T)
    img = torch.randn(*shape, device=dev)
    sess = d._fast_session(model, shape, {"y": yc}, None, None, False, False, img)
    gen = sess.run(d, "p", img, list(range(1000))[::-1], False, 0.0, graph=True, unroll=10)
    done = 0
    while done < 21:
        done += next(gen)["steps"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = 0
    while k < 50:
        k += next(gen)["steps"]
    e1.record()
    torch.cuda.synchronize()
    gen.close()
    ms = e0.elapsed_time(e1) / k
    print("B=%4d M=%6d: %.3f ms/step  (%.1f steps/s, %.0f poses/s)" % (B, B * T, ms, 1e3 / ms, B * T * 1e3 / ms))
    del model, sess, gen

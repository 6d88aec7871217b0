"""Probe (GPU box): wall time of repeated public-API p_sample_loop calls, graph driver on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import bench

dev = torch.device("cuda")
B, T, K = 256, 60, int(sys.argv[1]) if len(sys.argv) > 1 else 100
model, mkdiff = bench.build_ours(dev, B, T)
from regennet_b200 import synthetic
_, y = synthetic.make_inputs(B, 56, 6, T, seed=10)
cm_host = y["cmotion"].pin_memory()
out_host = torch.empty((B, 56, 6, T)).pin_memory()
d = mkdiff([K])
for mode in ["", "0", ""]:
    os.environ["REGEN_CUDA_GRAPH"] = mode
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ycm = {"cmotion": cm_host.to(dev, non_blocking=True)}
        s = d.p_sample_loop(model, (B, 56, 6, T), clip_denoised=False, model_kwargs={"y": ycm})
        t1 = time.perf_counter()
        out_host.copy_(s, non_blocking=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("graph=%r rep %d: enqueue %.1f ms, total %.1f ms -> %.1f steps/s" % (mode, rep, (t1 - t0) * 1e3, (t2 - t0) * 1e3, K / (t2 - t0)))

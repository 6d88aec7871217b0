"""Profiling target (GPU box): a few host-enqueued denoising steps of BASELINE config 2 (B=256, T=60) through the public
sampler -- every kernel of the step is an ordinary launch (no CUDA graph), so `ncu -s/-c` can pick steady-state launches.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
        python tools/ncu_target.py 4
    ncu --set full --clock-control none --import-source on -s 330 -c 60 -f -o gpurun_out/r02_full python tools/ncu_target.py 4
(setup + step 1 launch ~110 kernels, every later step 55: 46 library kernels + torch's Philox / fill kernels.)
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import bench
from regennet_b200 import synthetic

os.environ["REGEN_CUDA_GRAPH"] = "0"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
B, T = 256, 60
model, mkdiff = bench.build_ours(dev)
d = mkdiff("ddim%d" % steps)
_, y = synthetic.make_inputs(B, 56, 6, T, seed=10)
torch.manual_seed(10)
out = d.p_sample_loop(model, (B, 56, 6, T), clip_denoised=False, model_kwargs={"y": {"cmotion": y["cmotion"].to(dev)}})
R = out.permute(0, 3, 1, 2)[:, :, :-1].contiguous()
from regennet_b200 import rotation_conversions
rotation_conversions.rotation_6d_to_matrix(R)
torch.cuda.synchronize()
print("done", float(out.abs().max()))

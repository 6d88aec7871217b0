"""Experiment (GPU box): one B=256 sampling session vs S concurrent sessions of B=256/S on separate streams."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import cases
from regennet_b200 import gaussian_diffusion as gd, respace, synthetic
from regennet_b200.cmdm import CMDM

dev = torch.device("cuda")
betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
diff = respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, [1000]), betas=betas,
                               model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                               loss_type=gd.LossType.MSE)
K, W = 40, 5


def make(B):
    m = CMDM(**cases.MODELS["ntu"])
    m.load_state_dict(synthetic.make_state_dict(seed=0, **cases.synth_kw("ntu")), strict=False)
    m = m.to(dev).eval()
    _, y = synthetic.make_inputs(B, 56, 6, 60, seed=10)
    yc = {"cmotion": y["cmotion"].to(dev)}
    shape = (B, 56, 6, 60)
    img = torch.randn(*shape, device=dev)
    sess = diff._fast_session(m, shape, {"y": yc}, None, None, False, False, img)
    return m, sess.run(diff, "p", img, list(range(1000))[::-1], False, 0.0)


for S in [1, 2, 4]:
    B = 256 // S
    streams = [torch.cuda.Stream() for _ in range(S)]
    gens = []
    keep = []
    for s in streams:
        with torch.cuda.stream(s):
            m, g = make(B)
            keep.append(m)
            gens.append(g)
    for _ in range(W):
        for s, g in zip(streams, gens):
            with torch.cuda.stream(s):
                next(g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        for s, g in zip(streams, gens):
            with torch.cuda.stream(s):
                next(g)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("S=%d streams x B=%d: %.3f ms per full-batch step -> %.1f steps/s" % (S, B, dt / K * 1e3, K / dt))
    for g in gens:
        g.close()
    del gens, keep

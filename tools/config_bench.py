"""GPU box: device-resident step time of the other BASELINE.json configs (per-GPU shapes)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import cases
from regennet_b200 import gaussian_diffusion as gd, respace, synthetic, _lib
from regennet_b200.cmdm import CMDM
from regennet_b200.cfg_sampler import ClassifierFreeSampleModel

dev = torch.device("cuda")
betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
lib = _lib.lib()


def run(name, model_name, B, T, cfg, ddim, respacing, K=30, W=5):
    mk = cases.MODELS[model_name]
    m = CMDM(**mk)
    m.load_state_dict(synthetic.make_state_dict(seed=0, **cases.synth_kw(model_name)), strict=False)
    m = m.to(dev).eval()
    inner = m
    diff = respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, respacing), betas=betas,
                                   model_mean_type=gd.ModelMeanType.START_X,
                                   model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    _, y = synthetic.make_inputs(B, mk["njoints"], mk["nfeats"], T, seed=10, cond_mode=mk["cond_mode"],
                                 num_actions=mk["num_actions"], scale=2.5 if cfg else None)
    yc = {k: v.to(dev) for k, v in y.items()}
    run_model = ClassifierFreeSampleModel(m) if cfg else m
    shape = (B, mk["njoints"], mk["nfeats"], T)
    img = torch.randn(*shape, device=dev)
    sess = diff._fast_session(run_model, shape, {"y": yc}, None, None, False, False, img)
    n = diff.num_timesteps
    gen = sess.run(diff, "ddim" if ddim else "p", img, list(range(n))[::-1], False, 0.0)
    for _ in range(W):
        next(gen)
    torch.cuda.synchronize()
    _lib.check(lib.regen_profile_begin(inner._handle.ptr), "pb")
    t0 = time.perf_counter()
    for _ in range(K):
        next(gen)
    ms = (_lib.c_float * 4)()
    nn = (_lib.c_int * 4)()
    _lib.check(lib.regen_profile_end(inner._handle.ptr, ms, nn), "pe")
    dt = time.perf_counter() - t0
    gen.close()
    # throughput through the CUDA-graph driver (what p_sample_loop / ddim_sample_loop use)
    img = torch.randn(*shape, device=dev)
    gen = sess.run(diff, "ddim" if ddim else "p", img, list(range(n))[::-1], False, 0.0, graph=True, unroll=10)
    done = 0
    while done < 11:
        done += next(gen)["steps"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    KG = 0
    while KG < (60 if n >= 100 else 20):
        KG += next(gen)["steps"]
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / KG
    gen.close()
    print("%-34s B=%3d T=%3d cfg=%d: graph driver %.3f ms/step (%.1f steps/s) | host-enqueued %.3f ms/step | "
          "gemm %.3f attn %.3f ln %.3f other %.3f ms" % (name, B, T, cfg, gms, 1e3 / gms, dt / K * 1e3, ms[0] / K,
                                                         ms[1] / K, ms[2] / K, ms[3] / K))


only = set(sys.argv[1:])   # e.g. `config_bench.py 3 5` runs configs 3 and 5 only
if not only or "1" in only:
    run("config1 NTU B=1 (latency)", "ntu", 1, 60, False, False, [1000], K=100)
if not only or "2" in only:
    run("config2 NTU B=256", "ntu", 256, 60, False, False, [1000])
if not only or "3" in only:
    run("config3 Chi3D CFG B=128 T=150", "chi3d", 128, 150, True, False, [1000])
if not only or "5" in only:
    run("config5 HML text CFG DDIM100 B=64", "hml", 64, 196, True, True, "ddim100")

"""GPU parity: full sampling loops through the public API (fast route) vs the oracle and vs the
golden samples the imported reference produced on CPU with the same RNG stream."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import cmdm_ref, sampler_ref
from regennet_b200 import gaussian_diffusion as gd
from regennet_b200 import respace, synthetic
from regennet_b200.cfg_sampler import ClassifierFreeSampleModel
from test_gpu_denoiser import _kw, get_model, to_cuda

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def _diffusion(rs):
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    return respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, rs if rs else [1000]), betas=betas,
                                   model_mean_type=gd.ModelMeanType.START_X,
                                   model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)


class cpu_rng_stream:
    """Make torch.randn_like(x_cuda) return what the reference's CPU run drew: a CPU tensor with the
    same shape AND strides filled by the CPU generator, moved to the GPU."""

    def __enter__(self):
        self.orig = torch.randn_like

        def fake(x, **kw):
            if not x.is_cuda:
                return self.orig(x, **kw)
            cpu = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype)
            n = self.orig(cpu)
            out = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
            out.copy_(n)
            return out

        torch.randn_like = fake
        return self

    def __exit__(self, *a):
        torch.randn_like = self.orig


@pytest.mark.parametrize("name", sorted(cases.LOOP_CASES))
def test_loop_reproduces_reference_golden(built_lib, name):
    c = cases.LOOP_CASES[name]
    mk = cases.MODELS[c["model"]]
    gold = torch.from_numpy(np.load(os.path.join(HERE, "loops.npz"))[name])
    model, sd = get_model(c["model"], c["wseed"])
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    d = _diffusion(c["respacing"])
    run = ClassifierFreeSampleModel(model) if "cfg_scale" in c else model
    shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
    torch.manual_seed(c["seed"])
    init = torch.randn(*shape)  # the reference's th.randn(*shape) on CPU
    fn = d.ddim_sample_loop if c["ddim"] else d.p_sample_loop
    with cpu_rng_stream():
        out = fn(run, shape, noise=init.cuda(), clip_denoised=False, model_kwargs={"y": to_cuda(y)})
    assert out.shape == gold.shape
    err = (out.cpu() - gold).abs().max().item()
    print("%s: %d steps, max abs err vs reference golden %.3e" % (name, d.num_timesteps, err))
    assert err < TOL


def test_fast_route_is_taken_and_matches_generic_route(built_lib):
    model, sd = get_model("ntu", 0)
    _, y = synthetic.make_inputs(3, 56, 6, 60, seed=21)
    d = _diffusion("ddim5")
    shape = (3, 56, 6, 60)
    yc = to_cuda(y)
    assert d._fast_session(model, shape, {"y": yc}, None, None, False, False, torch.zeros(1, device="cuda")) is not None
    torch.manual_seed(1)
    a = d.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": yc})
    # generic route: hide the fast hook behind a plain callable wrapper
    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, t, y=None):
            return self.m(x, t, y)

    torch.manual_seed(1)
    b = d.p_sample_loop(Wrap(model), shape, clip_denoised=False, model_kwargs={"y": yc})
    assert torch.equal(a, b)  # same kernels, same noise stream
    assert a.permute(3, 0, 1, 2).is_contiguous()


def test_progressive_yields_every_step_and_clip(built_lib):
    model, sd = get_model("ntu", 0)
    _, y = synthetic.make_inputs(2, 56, 6, 60, seed=22)
    d = _diffusion("ddim5")
    torch.manual_seed(2)
    outs = list(d.p_sample_loop_progressive(model, (2, 56, 6, 60), clip_denoised=True, model_kwargs={"y": to_cuda(y)}))
    assert len(outs) == 5
    for o in outs:
        assert set(o) == {"sample", "pred_xstart"}
        assert o["pred_xstart"].abs().max().item() <= 1.0
    # last step has t == 0: no noise, coef1 == 1, coef2 == 0 -> the sample IS the (clipped) prediction
    assert torch.allclose(outs[-1]["sample"], outs[-1]["pred_xstart"], atol=1e-6)


def test_seeded_gpu_loop_vs_oracle_with_recorded_noise(built_lib):
    """CUDA RNG stream: record the noise our loop drew, replay it through the oracle."""
    mk = cases.MODELS["ntu"]
    model, sd = get_model("ntu", 0)
    _, y = synthetic.make_inputs(2, 56, 6, 60, seed=23)
    d = _diffusion("ddim20")
    shape = (2, 56, 6, 60)
    noises = []
    orig = torch.randn_like

    def rec(x, **kw):
        n = orig(x, **kw)
        noises.append(n.cpu())
        return n

    torch.manual_seed(10)
    init = torch.randn(*shape, device="cuda")
    torch.randn_like = rec
    try:
        out = d.p_sample_loop(model, shape, noise=init, clip_denoised=False, model_kwargs={"y": to_cuda(y)})
    finally:
        torch.randn_like = orig
    assert len(noises) == 20
    it = iter(noises)
    smp = sampler_ref.Sampler(timestep_respacing="ddim20")
    want, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **_kw(mk)), shape,
                       noise_fn=lambda x: next(it), init_noise=init.cpu())
    err = (out.cpu() - want).abs().max().item()
    print("20-step loop max abs err vs oracle %.3e" % err)
    assert err < TOL


def _loop_eager_and_graph(monkeypatch, d, run, shape, yc, seed, ddim=False, unroll=None, clip=False):
    """Same seed through the step-by-step driver (REGEN_CUDA_GRAPH=0) and the graph-replay driver."""
    fn = d.ddim_sample_loop if ddim else d.p_sample_loop
    monkeypatch.setenv("REGEN_CUDA_GRAPH", "0")
    torch.manual_seed(seed)
    a = fn(run, shape, clip_denoised=clip, model_kwargs={"y": yc})
    monkeypatch.setenv("REGEN_CUDA_GRAPH", str(unroll) if unroll else "")
    torch.manual_seed(seed)
    b = fn(run, shape, clip_denoised=clip, model_kwargs={"y": yc})
    return a, b


@pytest.mark.parametrize("steps,unroll", [(25, 4), (27, 4), (30, None), (13, 6)])
def test_graph_replay_is_bit_identical_to_step_by_step(built_lib, monkeypatch, steps, unroll):
    """CUDA-graph replay of the loop (regen_step_tables + denoiser + randn_like + update captured once) must give
    exactly the samples of the host-enqueued loop: same kernels, same Philox noise stream, integer timestep
    bookkeeping on the device (respace.py:125-126) bit-exact.  Covers trailing step-by-step steps (27 = 1 + 6*4 + 2)."""
    from regennet_b200 import _lib
    model, sd = get_model("ntu", 0)
    _, y = synthetic.make_inputs(3, 56, 6, 60, seed=31)
    d = _diffusion("ddim%d" % steps)
    n0 = _lib.lib().regen_launch_count()
    a, b = _loop_eager_and_graph(monkeypatch, d, model, (3, 56, 6, 60), to_cuda(y), seed=5, unroll=unroll)
    assert torch.equal(a, b)
    assert b.permute(3, 0, 1, 2).is_contiguous()
    assert model.__dict__.get("_graph_cache"), "graph driver was not used"
    # replayed launches are credited to the library's counter: both drivers launch ~the same number of kernels
    assert _lib.lib().regen_launch_count() - n0 > 2 * steps * 40


def test_graph_replay_cfg_ddim_clip_and_cache_reuse(built_lib, monkeypatch):
    mk = cases.MODELS["chi3d"]
    model, sd = get_model("chi3d", 2)
    run = ClassifierFreeSampleModel(model)
    shape = (2, 56, 6, 60)
    d = _diffusion("ddim20")
    _, y1 = synthetic.make_inputs(2, 56, 6, 60, seed=41, cond_mode="action", num_actions=mk["num_actions"], scale=2.5)
    _, y2 = synthetic.make_inputs(2, 56, 6, 60, seed=42, cond_mode="action", num_actions=mk["num_actions"], scale=1.5)
    for ddim, clip in [(True, False), (False, True)]:
        a1, b1 = _loop_eager_and_graph(monkeypatch, d, run, shape, to_cuda(y1), seed=7, ddim=ddim, unroll=5, clip=clip)
        assert torch.equal(a1, b1)
        n_graphs = len(model.__dict__["_graph_cache"])
        # second loop, new conditioning + guidance scale: the cached graph is reused on its static buffers
        a2, b2 = _loop_eager_and_graph(monkeypatch, d, run, shape, to_cuda(y2), seed=8, ddim=ddim, unroll=5, clip=clip)
        assert torch.equal(a2, b2)
        assert len(model.__dict__["_graph_cache"]) == n_graphs
        assert not torch.equal(a1, a2)
        # results are detached from the static buffers
        assert torch.equal(a1, b1)


def _edit_inputs(B, T, seed):
    """sample/edit.py:75-90 'in_between': keep a prefix and a suffix of the input motion, generate the middle."""
    g = torch.Generator().manual_seed(seed)
    motion = torch.randn(B, 56, 6, T, generator=g)
    mask = torch.ones(B, 56, 6, T, dtype=torch.bool)
    mask[:, :, :, int(0.25 * T):int(0.75 * T)] = False
    mask[1, :10] = True      # ragged: sample 1 also keeps some joints everywhere
    return mask, motion


def test_inpainting_fast_route_matches_generic_route_and_oracle(built_lib, monkeypatch):
    """Motion editing (gaussian_diffusion.py:319-323): x0 <- x0 * ~mask + motion * mask every step.  The fused route
    (regen_inpaint_blend, also inside the captured graph) must equal the generic route bit for bit and the oracle to 1e-3."""
    mk = cases.MODELS["ntu"]
    model, sd = get_model("ntu", 0)
    B, T = 2, 60
    shape = (B, 56, 6, T)
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=61)
    mask, motion = _edit_inputs(B, T, 62)
    yc = to_cuda(dict(y, inpainting_mask=mask, inpainted_motion=motion))
    d = _diffusion("ddim16")
    assert d._fast_session(model, shape, {"y": yc}, None, None, False, False, torch.zeros(1, device="cuda")) is not None

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, t, y=None):
            return self.m(x, t, y)

    outs = {}
    for name, run, graph in [("generic", Wrap(model), "0"), ("fast", model, "0"), ("graph", model, "5")]:
        monkeypatch.setenv("REGEN_CUDA_GRAPH", graph)
        torch.manual_seed(3)
        outs[name] = d.p_sample_loop(run, shape, clip_denoised=False, model_kwargs={"y": yc})
    assert torch.equal(outs["generic"], outs["fast"])
    assert torch.equal(outs["fast"], outs["graph"])
    # last step (t = 0): sample == blended prediction, so the kept region is the input motion exactly
    assert torch.equal(outs["fast"].cpu()[mask], motion[mask])

    # oracle on the recorded noise
    noises = []
    orig = torch.randn_like

    def rec(x, **kw):
        n = orig(x, **kw)
        noises.append(n.cpu())
        return n

    torch.manual_seed(10)
    init = torch.randn(*shape, device="cuda")
    torch.randn_like = rec
    try:
        out = d.p_sample_loop(model, shape, noise=init, clip_denoised=False, model_kwargs={"y": yc})
    finally:
        torch.randn_like = orig
    it = iter(noises)
    smp = sampler_ref.Sampler(timestep_respacing="ddim16", inpaint=(mask, motion))
    want, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **_kw(mk)), shape,
                       noise_fn=lambda x: next(it), init_noise=init.cpu())
    err = (out.cpu() - want).abs().max().item()
    print("16-step in-betweening loop max abs err vs oracle %.3e" % err)
    assert err < TOL


def test_graph_driver_with_editing_arguments(built_lib, monkeypatch):
    """skip_timesteps / init_image (q_sample start, sample/edit.py), a caller-supplied noise tensor and progress=True go
    through the same graph driver; all must equal the step-by-step driver bit for bit."""
    model, sd = get_model("ntu", 0)
    shape = (2, 56, 6, 60)
    _, y = synthetic.make_inputs(2, 56, 6, 60, seed=71)
    yc = to_cuda(y)
    d = _diffusion("ddim30")
    g = torch.Generator().manual_seed(72)
    init_image = torch.randn(*shape, generator=g).cuda()
    noise = torch.randn(*shape, generator=g).cuda()
    outs = []
    for mode in ("0", "6"):
        monkeypatch.setenv("REGEN_CUDA_GRAPH", mode)
        torch.manual_seed(4)
        outs.append(d.p_sample_loop(model, shape, noise=noise.clone(), clip_denoised=False, model_kwargs={"y": yc},
                                    skip_timesteps=7, init_image=init_image, progress=True))
    assert torch.equal(outs[0], outs[1])
    # dump_steps keeps every requested intermediate sample: step-by-step driver, values equal to the plain loop's prefix
    monkeypatch.setenv("REGEN_CUDA_GRAPH", "")
    torch.manual_seed(4)
    dumps = d.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": yc}, dump_steps=[0, 10, 29])
    torch.manual_seed(4)
    final = d.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": yc})
    assert len(dumps) == 3 and torch.equal(dumps[-1], final)


def test_full_size_loop_graph_equals_step_by_step_and_is_deterministic(built_lib, monkeypatch):
    """BASELINE config 2 (B=256, T=60): the graph driver equals the step-by-step driver bit for bit at full size, a repeated
    run reproduces itself exactly (no atomics / order-dependent reductions anywhere on the path), and the result is a
    finite motion whose last step returned the prediction itself (coef1[0] = 1, coef2[0] = 0, no noise at t = 0)."""
    model, sd = get_model("ntu", 0)
    B, T = 256, 60
    shape = (B, 56, 6, T)
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=81)
    yc = to_cuda(y)
    d = _diffusion("ddim14")
    outs = []
    for mode in ("0", "", ""):
        monkeypatch.setenv("REGEN_CUDA_GRAPH", mode)
        torch.manual_seed(12)
        outs.append(d.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": yc}))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert torch.isfinite(outs[0]).all()
    last = list(d.p_sample_loop_progressive(model, shape, noise=outs[0], clip_denoised=False, model_kwargs={"y": yc},
                                            skip_timesteps=13))[-1]
    assert torch.equal(last["sample"], last["pred_xstart"])


# ------------------------------------------------------------------------------------------ PLMS
@pytest.mark.parametrize("name", sorted(cases.PLMS_LOOP_CASES))
def test_plms_loop_reproduces_reference_golden(built_lib, name):
    """plms_sample_loop (diffusion/gaussian_diffusion.py:1100-1202) through the public API vs the samples the imported
    reference produced on CPU (tests/golden/make_golden_plms.py).  PLMS draws x_T only, so no RNG patching is needed."""
    c = cases.PLMS_LOOP_CASES[name]
    mk = cases.MODELS[c["model"]]
    gold = torch.from_numpy(np.load(os.path.join(HERE, "loops_plms.npz"))[name])
    model, sd = get_model(c["model"], c["wseed"])
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    d = _diffusion(c["respacing"])
    run = ClassifierFreeSampleModel(model) if "cfg_scale" in c else model
    shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
    torch.manual_seed(c["seed"])
    init = torch.randn(*shape)
    launches0 = built_lib.regen_launch_count()
    out = d.plms_sample_loop(run, shape, noise=init.cuda(), clip_denoised=bool(c.get("clip")),
                             model_kwargs={"y": to_cuda(y)}, order=c["order"])
    assert built_lib.regen_launch_count() > launches0
    assert out.shape == gold.shape
    err = (out.cpu() - gold).abs().max().item()
    print("%s: %d steps order %d, max abs err vs reference golden %.3e" % (name, d.num_timesteps, c["order"], err))
    assert err < TOL


def test_plms_step_arithmetic_matches_oracle_on_injected_model_outputs(built_lib):
    """The three PLMS kernels alone: a fake model (fixed linear map of x, per-sample timestep-dependent) makes the whole
    difference arithmetic, so the fp32 operation order must reproduce the oracle to rounding."""
    B, shape = 3, (3, 5, 6, 17)
    smp = sampler_ref.Sampler(timestep_respacing="ddim7")
    d = _diffusion("ddim7")

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x, t, y=None):
            return 0.7 * x + 0.001 * t.float().view(-1, 1, 1, 1)

    torch.manual_seed(3)
    init = torch.randn(*shape)
    for order in (2, 3, 4):
        ora = smp.plms_loop(lambda xx, tt: 0.7 * xx + 0.001 * tt.float().view(-1, 1, 1, 1), shape, order=order,
                            init_noise=init.clone())
        out = d.plms_sample_loop(Fake().cuda(), shape, noise=init.cuda(), clip_denoised=False, order=order)
        err = (out.cpu() - ora).abs().max().item()
        print("order %d: max abs err %.3e (absmax %.3f)" % (order, err, ora.abs().max()))
        assert err < 1e-5 * max(1.0, ora.abs().max().item())


def test_plms_error_behaviour(built_lib):
    model, _ = get_model("ntu", 0)
    d = _diffusion("ddim5")
    x = torch.randn(1, 56, 6, 60, device="cuda")
    t = torch.tensor([4], device="cuda")
    _, y = synthetic.make_inputs(1, 56, 6, 60, seed=1)
    with pytest.raises(ValueError, match="order is invalid"):
        d.plms_sample(model, x, t, model_kwargs={"y": to_cuda(y)}, order=5)
    with pytest.raises(TypeError):  # the reference dereferences old_out=None at :1067 when order == 1
        d.plms_sample(model, x, t, model_kwargs={"y": to_cuda(y)}, order=1)


def test_add_mode_loop_reproduces_reference_golden(built_lib):
    """arch='online' with cm_mode='add': 10-step ancestral loop vs the reference's samples (make_golden_add.py)."""
    from test_gpu_denoiser import get_add_model
    name = "add_loop_ntu_p10"
    c = cases.ADD_LOOP_CASES[name]
    mk = cases.ADD_MODELS[c["model"]]
    gold = torch.from_numpy(np.load(os.path.join(HERE, "loops_add.npz"))[name])
    model, sd = get_add_model(c["model"], c["wseed"])
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"])
    d = _diffusion(c["respacing"])
    shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
    torch.manual_seed(c["seed"])
    init = torch.randn(*shape)
    with cpu_rng_stream():
        out = d.p_sample_loop(model, shape, noise=init.cuda(), clip_denoised=False, model_kwargs={"y": to_cuda(y)})
    err = (out.cpu() - gold).abs().max().item()
    print("%s: max abs err vs reference golden %.3e" % (name, err))
    assert err < TOL

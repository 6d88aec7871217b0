"""GPU parity: auto-regressive ("online") inference (regennet_b200.autoregressive) vs the oracle's literal
restatement of eval/a2m/stgcn_eval.py:50-67 -- T full-length sampling loops, one per revealed actor frame --
on injected noise.  Loop f of the reference draws noise for all T frames; the truncated / stacked drivers must
reproduce frame f of its result from that noise restricted to frames <= f (tolerance 1e-3, as everywhere)."""
import pytest
import torch

import cases
from oracle import cmdm_ref, sampler_ref
from regennet_b200 import synthetic
from regennet_b200.autoregressive import auto_regressive_sample
from regennet_b200.cfg_sampler import ClassifierFreeSampleModel
from test_gpu_denoiser import _kw, get_model, to_cuda
from test_gpu_sampler import _diffusion

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _noise_bank(T, steps, shape, seed):
    g = torch.Generator().manual_seed(seed)
    init = [torch.randn(*shape, generator=g) for _ in range(T)]
    per_step = [[torch.randn(*shape, generator=g) for _ in range(steps)] for _ in range(T)]
    return init, per_step


def _run_ours(d, run, shape, yc, init, per_step, G, truncate, setting):
    """sample_fn wrapper: feeds call k (frames kG .. kG+g-1 stacked along the batch) the banked noise of those loops."""
    B, V, C, T = shape
    state = {"call": 0}
    orig = torch.randn_like

    def sample_fn(model, shp, clip_denoised, model_kwargs):
        k = state["call"]
        state["call"] += 1
        frames = list(range(k * G, min(k * G + G, T)))
        Tc = shp[-1]
        assert shp[0] == len(frames) * B and Tc == (frames[-1] + 1 if truncate else T)
        x_T = torch.cat([init[f][..., :Tc] for f in frames], 0).cuda()
        step = {"s": 0}

        def fake(x, **kw):
            n = torch.cat([per_step[f][step["s"]][..., :Tc] for f in frames], 0)
            step["s"] += 1
            out = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
            out.copy_(n)
            return out

        torch.randn_like = fake
        try:
            return d.p_sample_loop(model, shp, noise=x_T, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        finally:
            torch.randn_like = orig

    out = auto_regressive_sample(sample_fn, run, shape, {"y": yc}, setting=setting, truncate=truncate,
                                 frames_per_call=G)
    assert state["call"] == (T + G - 1) // G
    return out


@pytest.mark.parametrize("G,truncate", [(1, True), (3, True), (7, True), (1, False), (4, False)])
def test_auto_regressive_matches_reference_semantics(built_lib, G, truncate):
    mk = cases.MODELS["ntu"]
    model, sd = get_model("ntu", 0)
    B, T, steps = 2, 7, 3
    shape = (B, 56, 6, T)
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=51)
    d = _diffusion("ddim%d" % steps)
    init, per_step = _noise_bank(T, steps, shape, seed=77)
    smp = sampler_ref.Sampler(timestep_respacing="ddim%d" % steps)

    def loop_fn(f, cmotion):
        it = iter(per_step[f])
        yy = dict(y, cmotion=cmotion.clone())
        out, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, yy, **_kw(mk)), shape,
                          noise_fn=lambda x: next(it), init_noise=init[f])
        return out

    want = sampler_ref.auto_regressive(loop_fn, y["cmotion"], setting="cmdm")
    yc = to_cuda(y)
    got = _run_ours(d, model, shape, yc, init, per_step, G, truncate, "cmdm")
    assert got.shape == (B, 56, 12, T)
    assert torch.equal(got[:, :, :6].cpu(), y["cmotion"])            # actor half is copied through
    assert torch.equal(yc["cmotion"].cpu(), y["cmotion"])            # left in model_kwargs as the reference does
    err = (got.cpu() - want).abs().max().item()
    print("auto-regressive G=%d truncate=%s: max abs err vs oracle %.3e" % (G, truncate, err))
    assert err < TOL


def test_auto_regressive_guided_action_model(built_lib):
    """Stacked loops replicate the per-sample conditioning (action, guidance scale) along the batch."""
    mk = cases.MODELS["chi3d"]
    model, sd = get_model("chi3d", 2)
    run = ClassifierFreeSampleModel(model)
    B, T, steps, G = 2, 5, 2, 2
    shape = (B, 56, 6, T)
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=52, cond_mode="action", num_actions=mk["num_actions"], scale=2.5)
    d = _diffusion("ddim%d" % steps)
    init, per_step = _noise_bank(T, steps, shape, seed=78)
    smp = sampler_ref.Sampler(timestep_respacing="ddim%d" % steps)

    def loop_fn(f, cmotion):
        it = iter(per_step[f])
        yy = dict(y, cmotion=cmotion.clone())
        out, _ = smp.loop(lambda xx, tt: cmdm_ref.cfg_forward(sd, xx, tt, yy, **_kw(mk)), shape,
                          noise_fn=lambda x: next(it), init_noise=init[f])
        return out

    want = sampler_ref.auto_regressive(loop_fn, y["cmotion"], setting="sample")
    got = _run_ours(d, run, shape, to_cuda(y), init, per_step, G, True, "sample")
    assert got.shape == shape
    err = (got.cpu() - want).abs().max().item()
    print("auto-regressive CFG: max abs err vs oracle %.3e" % err)
    assert err < TOL

"""GPU parity for arch='offline' (model/cmdm.py:63-71, 228-238: nn.TransformerEncoder over [condition token | frames],
no attention mask) vs the golden outputs of the imported reference (tests/golden/make_golden_offline.py) and the oracle."""
import math
import os

import numpy as np
import pytest
import torch

import cases
from oracle import cmdm_ref, sampler_ref
from regennet_b200 import _lib, synthetic
from regennet_b200.cfg_sampler import ClassifierFreeSampleModel
from regennet_b200.cmdm import CMDM
from test_gpu_attention import run as run_attention
from test_gpu_denoiser import to_cuda
from test_gpu_sampler import _diffusion, cpu_rng_stream

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL, TOL_TIGHT = 1e-3, 2e-4
_models = {}


def get_model(name, wseed):
    key = (name, wseed)
    if key not in _models:
        m = CMDM(**cases.OFFLINE_MODELS[name])
        sd = synthetic.make_state_dict(seed=wseed, **cases.synth_kw_offline(name))
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("clip_model.") for k in missing)
        _models[key] = (m.cuda().eval(), sd)
    return _models[key]


def _kw(mk):
    return dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"],
                arch="offline")


@pytest.mark.parametrize("B,T", [(1, 1), (2, 61), (3, 64), (2, 65), (2, 151), (1, 197), (2, 256)])
def test_unmasked_attention_matches_fp64(built_lib, B, T):
    """The attention kernel without the causal mask (test-hook flag 2), token counts of the offline configs (T + 1)."""
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(T * B, 1536, generator=g).cuda()
    out = run_attention(built_lib, qkv, B, T, dbg=2)
    x = qkv.double().view(T, B, 3, 4, 128)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    p = torch.softmax(torch.einsum("ibhd,jbhd->bhij", q, k) / math.sqrt(128.0), dim=-1)
    want = torch.einsum("bhij,jbhd->ibhd", p, v).reshape(T * B, 512)
    err = (out.double() - want).abs().max().item()
    print("unmasked B=%d T=%d max abs err %.3e" % (B, T, err))
    assert not torch.isnan(out).any()
    assert err < 1e-4


@pytest.mark.parametrize("name", sorted(cases.OFFLINE_FORWARD_CASES))
def test_offline_forward_matches_reference_golden(built_lib, name):
    c = cases.OFFLINE_FORWARD_CASES[name]
    mk = cases.OFFLINE_MODELS[c["model"]]
    gold = torch.from_numpy(np.load(os.path.join(HERE, "forward_offline.npz"))[name])
    model, sd = get_model(c["model"], c["wseed"])
    x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    run = ClassifierFreeSampleModel(model) if "cfg_scale" in c else model
    with torch.no_grad():
        out = run(x.cuda(), torch.tensor(c["t"], dtype=torch.long).cuda(), to_cuda(y))
    assert out.shape == gold.shape and out.permute(3, 0, 1, 2).is_contiguous()
    err = (out.cpu() - gold).abs().max().item()
    print("%s: max abs err vs reference golden %.3e" % (name, err))
    assert err < TOL_TIGHT


@pytest.mark.parametrize("B,T", [(1, 1), (5, 3), (7, 63), (7, 64), (33, 60), (256, 60)])
def test_offline_forward_matches_oracle_various_sizes(built_lib, B, T):
    mk = cases.OFFLINE_MODELS["ntu_off"]
    model, sd = get_model("ntu_off", 4)
    x, y = synthetic.make_inputs(B, 56, 6, T, seed=200 + B)
    g = torch.Generator().manual_seed(B)
    t = torch.randint(0, 1000, (B,), generator=g)
    with torch.no_grad():
        out = model(x.cuda(), t.cuda(), to_cuda(y)).cpu()
    sel = torch.arange(B) if B <= 33 else torch.tensor([0, 1, 63, 127, 128, 200, 255])
    with torch.no_grad():
        want = cmdm_ref.cmdm_forward(sd, x[sel], t[sel], {"cmotion": y["cmotion"][sel]}, **_kw(mk))
    err = (out[sel] - want).abs().max().item()
    print("offline B=%d T=%d: max abs err vs oracle %.3e" % (B, T, err))
    assert err < TOL_TIGHT


def test_offline_is_not_causal(built_lib):
    """Every frame attends to every frame: changing late frames changes early outputs (unlike arch='online')."""
    model, _ = get_model("ntu_off", 4)
    x, y = synthetic.make_inputs(2, 56, 6, 60, seed=3)
    t = torch.tensor([400, 30]).cuda()
    with torch.no_grad():
        a = model(x.cuda(), t, to_cuda(y)).cpu()
        x2 = x.clone()
        x2[..., 40:] += 1.0
        b = model(x2.cuda(), t, to_cuda(y)).cpu()
    assert not torch.allclose(a[..., :40], b[..., :40], atol=1e-4)


def test_offline_loop_reproduces_reference_golden_and_graph_driver(built_lib, monkeypatch):
    name = "off_loop_ntu_p10"
    c = cases.OFFLINE_LOOP_CASES[name]
    mk = cases.OFFLINE_MODELS[c["model"]]
    gold = torch.from_numpy(np.load(os.path.join(HERE, "loops_offline.npz"))[name])
    model, sd = get_model(c["model"], c["wseed"])
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
    # CUDA-graph driver == step-by-step driver, bit for bit
    d2 = _diffusion("ddim13")
    res = []
    for mode in ("0", "4"):
        monkeypatch.setenv("REGEN_CUDA_GRAPH", mode)
        torch.manual_seed(9)
        res.append(d2.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": to_cuda(y)}))
    assert torch.equal(res[0], res[1])

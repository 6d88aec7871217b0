"""The oracle (oracle/) pinned against golden outputs of the imported reference (tests/golden)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import cmdm_ref, sampler_ref, schedule
from regennet_b200 import synthetic

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# fp32 restatement vs the reference's torch.nn layers: reassociation-level differences only
TOL = 2e-5

_sd_cache = {}


def _sd(model, wseed):
    key = (model, wseed)
    if key not in _sd_cache:
        _sd_cache[key] = synthetic.make_state_dict(seed=wseed, **cases.synth_kw(model))
    return _sd_cache[key]


def _kw(mk):
    return dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])


@pytest.mark.parametrize("name", sorted(cases.FORWARD_CASES))
def test_forward_matches_reference(name):
    c = cases.FORWARD_CASES[name]
    mk = cases.MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "forward.npz"))[name]
    x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    if c.get("uncond"):
        y["uncond"] = True
    t = torch.tensor(c["t"], dtype=torch.long)
    fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
    with torch.no_grad():
        out = fwd(_sd(c["model"], c["wseed"]), x, t, y, **_kw(mk))
    assert out.shape == gold.shape
    assert np.abs(out.numpy() - gold).max() < TOL


@pytest.mark.parametrize("name", sorted(cases.LOOP_CASES))
def test_sampling_loop_matches_reference(name):
    c = cases.LOOP_CASES[name]
    mk = cases.MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "loops.npz"))[name]
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    sd = _sd(c["model"], c["wseed"])
    fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
    smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
    torch.manual_seed(c["seed"])
    out, _ = smp.loop(lambda xx, tt: fwd(sd, xx, tt, y, **_kw(mk)), (c["B"], mk["njoints"], mk["nfeats"], c["T"]),
                      ddim=c["ddim"])
    assert np.abs(out.numpy() - gold).max() < TOL


@pytest.mark.parametrize("rs", cases.RESPACINGS)
def test_schedule_tables_bit_exact(rs):
    g = np.load(os.path.join(HERE, "schedule.npz"))
    betas = schedule.named_beta_schedule("cosine", 1000)
    assert np.array_equal(betas, g["betas_cosine_1000"])
    assert np.array_equal(schedule.named_beta_schedule("linear", 1000), g["betas_linear_1000"])
    tab, tmap = schedule.spaced_tables(betas, schedule.space_timesteps(1000, rs if rs else [1000]))
    tag = "rs[%s]" % rs
    assert np.array_equal(np.array(tmap, dtype=np.int64), g[tag + ".timestep_map"])  # integer: bit-exact
    for f in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]:
        assert np.array_equal(getattr(tab, f), g[tag + "." + f]), f  # fp64: bit-exact


def test_rot6d_matches_reference():
    g = np.load(os.path.join(HERE, "rot6d.npz"))
    R = sampler_ref.rotation_6d_to_matrix(torch.from_numpy(g["d6"])).numpy()
    assert np.allclose(R, g["R"], atol=1e-6, equal_nan=True)


def _kw_off(mk):
    return dict(_kw(mk), arch="offline")


@pytest.mark.parametrize("name", sorted(cases.OFFLINE_FORWARD_CASES))
def test_offline_forward_matches_reference(name):
    """arch='offline' (model/cmdm.py:63-71, 228-238) against the reference's outputs (make_golden_offline.py)."""
    c = cases.OFFLINE_FORWARD_CASES[name]
    mk = cases.OFFLINE_MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "forward_offline.npz"))[name]
    x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    sd = synthetic.make_state_dict(seed=c["wseed"], **cases.synth_kw_offline(c["model"]))
    fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
    with torch.no_grad():
        out = fwd(sd, x, torch.tensor(c["t"], dtype=torch.long), y, **_kw_off(mk))
    assert out.shape == gold.shape
    assert np.abs(out.numpy() - gold).max() < TOL


def test_offline_sampling_loop_matches_reference():
    name = "off_loop_ntu_p10"
    c = cases.OFFLINE_LOOP_CASES[name]
    mk = cases.OFFLINE_MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "loops_offline.npz"))[name]
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"])
    sd = synthetic.make_state_dict(seed=c["wseed"], **cases.synth_kw_offline(c["model"]))
    smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
    torch.manual_seed(c["seed"])
    out, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **_kw_off(mk)),
                      (c["B"], mk["njoints"], mk["nfeats"], c["T"]))
    assert np.abs(out.numpy() - gold).max() < TOL


@pytest.mark.parametrize("name", sorted(cases.PLMS_LOOP_CASES))
def test_plms_loop_matches_reference(name):
    """PLMS (diffusion/gaussian_diffusion.py:1007-1202) against the reference's plms_sample_loop (make_golden_plms.py)."""
    c = cases.PLMS_LOOP_CASES[name]
    mk = cases.MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "loops_plms.npz"))[name]
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    sd = _sd(c["model"], c["wseed"])
    fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
    smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
    torch.manual_seed(c["seed"])
    out = smp.plms_loop(lambda xx, tt: fwd(sd, xx, tt, y, **_kw(mk)), (c["B"], mk["njoints"], mk["nfeats"], c["T"]),
                        order=c["order"], clip_denoised=bool(c.get("clip")))
    assert out.shape == gold.shape
    assert np.abs(out.numpy() - gold).max() < TOL


def test_plms_order1_from_fresh_loop_raises_like_reference():
    # the reference dereferences old_out (None) at :1067 when order == 1 on the first step
    smp = sampler_ref.Sampler(timestep_respacing="ddim5")
    with pytest.raises(TypeError):
        smp.plms_loop(lambda xx, tt: xx, (1, 2, 3, 4), order=1)


@pytest.mark.parametrize("name", sorted(cases.ADD_FORWARD_CASES))
def test_add_mode_forward_matches_reference(name):
    """arch='online', cm_mode='add' (model/cmdm.py:207-211) against the reference's outputs (make_golden_add.py)."""
    c = cases.ADD_FORWARD_CASES[name]
    mk = cases.ADD_MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "forward_add.npz"))[name]
    x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                 cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
    sd = synthetic.make_state_dict(seed=c["wseed"], **cases.synth_kw_add(c["model"]))
    assert "fuse_process.weight" not in sd
    fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
    with torch.no_grad():
        out = fwd(sd, x, torch.tensor(c["t"], dtype=torch.long), y, **_kw(mk))
    assert out.shape == gold.shape
    assert np.abs(out.numpy() - gold).max() < TOL


def test_add_mode_sampling_loop_matches_reference():
    name = "add_loop_ntu_p10"
    c = cases.ADD_LOOP_CASES[name]
    mk = cases.ADD_MODELS[c["model"]]
    gold = np.load(os.path.join(HERE, "loops_add.npz"))[name]
    _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"])
    sd = synthetic.make_state_dict(seed=c["wseed"], **cases.synth_kw_add(c["model"]))
    smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
    torch.manual_seed(c["seed"])
    out, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **_kw(mk)),
                      (c["B"], mk["njoints"], mk["nfeats"], c["T"]))
    assert np.abs(out.numpy() - gold).max() < TOL


@pytest.mark.parametrize("name", sorted(cases.STGCN_CASES))
def test_stgcn_oracle_matches_reference(name):
    """Evaluation feature extractor (oracle only so far): oracle/stgcn_ref.py against the reference's
    STGCN.forward outputs (eval/a2m/recognition/models/stgcn.py:76-126; tests/golden/make_golden_stgcn.py)."""
    from oracle import stgcn_ref
    c = cases.STGCN_CASES[name]
    g = np.load(os.path.join(HERE, "stgcn.npz"))
    A = torch.from_numpy(g[name + ".A"])
    sd = stgcn_ref.make_state_dict(A, c["in_channels"], c["num_class"], c["num_person"], seed=c["wseed"])
    x = torch.randn(c["N"], A.shape[1], c["in_channels"], c["T"], generator=torch.Generator().manual_seed(c["xseed"]))
    with torch.no_grad():
        feat, yhat = stgcn_ref.stgcn_forward(sd, x, c["num_person"])
    assert feat.shape == (c["N"], 256) and yhat.shape == (c["N"], c["num_class"])
    assert np.abs(feat.numpy() - g[name + ".features"]).max() < TOL
    assert np.abs(yhat.numpy() - g[name + ".yhat"]).max() < TOL

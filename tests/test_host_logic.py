"""Host-side logic of the product package (no GPU): schedules, respacing, API surface."""
import inspect
import os

import numpy as np
import pytest
import torch

import cases
from regennet_b200 import gaussian_diffusion as gd
from regennet_b200 import respace

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _spaced(rs):
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    return respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, rs if rs else [1000]), betas=betas,
                                   model_mean_type=gd.ModelMeanType.START_X,
                                   model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)


@pytest.mark.parametrize("rs", cases.RESPACINGS)
def test_tables_and_timestep_map_bit_exact_vs_reference(rs):
    g = np.load(os.path.join(HERE, "schedule.npz"))
    d = _spaced(rs)
    tag = "rs[%s]" % rs
    assert np.array_equal(np.array(d.timestep_map, dtype=np.int64), g[tag + ".timestep_map"])
    for f in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]:
        assert np.array_equal(getattr(d, f), g[tag + "." + f]), f


def test_beta_schedules_bit_exact():
    g = np.load(os.path.join(HERE, "schedule.npz"))
    assert np.array_equal(gd.get_named_beta_schedule("cosine", 1000), g["betas_cosine_1000"])
    assert np.array_equal(gd.get_named_beta_schedule("linear", 1000), g["betas_linear_1000"])
    with pytest.raises(NotImplementedError):
        gd.get_named_beta_schedule("sqrt", 10)


def test_space_timesteps_errors_and_edges():
    assert respace.space_timesteps(1000, "ddim5") == set(range(0, 1000, 200))
    assert respace.space_timesteps(10, [10]) == set(range(10))
    assert respace.space_timesteps(300, "10,15,20") == respace.space_timesteps(300, [10, 15, 20])
    with pytest.raises(ValueError):
        respace.space_timesteps(1000, "ddim999")   # no integer stride
    with pytest.raises(ValueError):
        respace.space_timesteps(10, [20])          # section too small


def test_loop_signatures_match_reference():
    want_p = ["self", "model", "shape", "noise", "clip_denoised", "denoised_fn", "cond_fn", "model_kwargs", "device",
              "progress", "skip_timesteps", "init_image", "randomize_class", "cond_fn_with_grad", "dump_steps",
              "const_noise"]
    assert list(inspect.signature(gd.GaussianDiffusion.p_sample_loop).parameters) == want_p
    want_d = ["self", "model", "shape", "noise", "clip_denoised", "denoised_fn", "cond_fn", "model_kwargs", "device",
              "progress", "eta", "skip_timesteps", "init_image", "randomize_class", "cond_fn_with_grad", "dump_steps",
              "const_noise"]
    assert list(inspect.signature(gd.GaussianDiffusion.ddim_sample_loop).parameters) == want_d
    sig = inspect.signature(gd.GaussianDiffusion.p_sample_loop).parameters
    assert sig["clip_denoised"].default is True and sig["noise"].default is None


def test_unsupported_variants_raise():
    betas = gd.get_named_beta_schedule("cosine", 10)
    d = gd.GaussianDiffusion(betas=betas, model_mean_type=gd.ModelMeanType.EPSILON,
                             model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    with pytest.raises(NotImplementedError):
        d._check_supported()
    with pytest.raises(ValueError):
        gd.GaussianDiffusion(betas=betas, model_mean_type=gd.ModelMeanType.START_X,
                             model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE, lambda_pose=2.0)


def test_cmdm_state_dict_keys_match_reference_layout():
    """The parameter container must expose exactly the reference's state-dict keys so that
    load_model_wo_clip(model, checkpoint) works: no unexpected keys, only clip_model.* may be missing."""
    import torch
    from regennet_b200 import synthetic
    from regennet_b200.cmdm import CMDM
    from regennet_b200.model_util import load_model_wo_clip
    for name in ["ntu", "chi3d", "hml"]:
        m = CMDM(**cases.MODELS[name])
        sd = synthetic.make_state_dict(seed=0, **cases.synth_kw(name))
        want = set(sd)
        have = set(m.state_dict())
        assert have == want, (sorted(want - have), sorted(have - want))
        load_model_wo_clip(m, sd)
        assert torch.equal(m.output_process.poseFinal.weight, sd["output_process.poseFinal.weight"])
        assert m.njoints * m.nfeats == m.input_feats
    assert abs(sum(p.numel() for p in CMDM(**cases.MODELS["ntu"]).parameters()) - 26.80e6) < 0.01e6


def test_cmdm_rejects_out_of_scope_variants_and_cpu_forward():
    import torch
    from regennet_b200.cmdm import CMDM
    with pytest.raises(NotImplementedError):
        CMDM(**dict(cases.MODELS["ntu"], arch="trans_enc"))
    with pytest.raises(NotImplementedError):
        CMDM(**dict(cases.MODELS["ntu"], latent_dim=256))
    m = CMDM(**cases.MODELS["ntu"])
    x = torch.zeros(1, 56, 6, 60)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, torch.zeros(1, dtype=torch.long), {"cmotion": x})   # no CPU fallback


def test_install_as_reference_modules_aliases():
    import sys
    import regennet_b200
    saved = {k: sys.modules.get(k) for k in regennet_b200._ALIASES}
    try:
        regennet_b200.install_as_reference_modules()
        from model.cmdm import CMDM as A            # noqa: the reference's import paths
        from diffusion.respace import SpacedDiffusion as B
        from utils.rotation_conversions import rotation_6d_to_matrix as C
        from regennet_b200.cmdm import CMDM
        from regennet_b200.respace import SpacedDiffusion
        from regennet_b200.rotation_conversions import rotation_6d_to_matrix
        assert A is CMDM and B is SpacedDiffusion and C is rotation_6d_to_matrix
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_model_util_factory_matches_reference_defaults():
    from argparse import Namespace
    from regennet_b200 import model_util
    args = Namespace(unconstrained=True, dataset="ntu", pose_rep="rot6d", body_model="smplx", latent_dim=512, layers=8,
                     cond_mask_prob=0.0, arch="online", cm_mode="concat", wo_pos_emb=False, emb_trans_dec=False,
                     setting="cmdm", timestep_respacing="ddim5", noise_schedule="cosine", sigma_small=True,
                     lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0, lambda_orient=0.0, lambda_body=0.0,
                     lambda_transl=0.0, num_person=1, vel_threshold=0.01)

    class D:
        class dataset:
            num_actions = 26
    model, diffusion = model_util.create_model_and_diffusion(args, D)
    assert (model.njoints, model.nfeats, model.cond_mode, model.num_frames) == (56, 6, "no_cond", 60)
    assert diffusion.num_timesteps == 5 and diffusion.timestep_map == [0, 200, 400, 600, 800]
    assert diffusion.model_var_type == gd.ModelVarType.FIXED_SMALL


def test_auto_regressive_driver_indexing():
    """Host logic of regennet_b200.autoregressive (eval/a2m/stgcn_eval.py:50-67): with a deterministic causal
    stand-in for the sampling loop, the literal, truncated and batch-stacked drivers all equal the oracle's loop."""
    import torch
    from oracle import sampler_ref
    from regennet_b200.autoregressive import auto_regressive_sample
    B, V, C, T = 3, 4, 2, 9
    g = torch.Generator().manual_seed(0)
    cm = torch.randn(B, V, C, T, generator=g)
    action = torch.arange(B).view(B, 1)
    calls = []

    def fake_loop(model, shape, clip_denoised, model_kwargs):
        y = model_kwargs["y"]
        assert tuple(y["cmotion"].shape) == tuple(shape) and y["action"].shape[0] == shape[0]
        assert len(y["action_text"]) == shape[0] and y["tag"] == "kept"
        calls.append(tuple(shape))
        # causal in time, independent across samples, conditioned on the per-sample action
        return torch.cumsum(y["cmotion"], dim=-1) * (1.0 + y["action"].view(-1, 1, 1, 1).float())

    def oracle_loop(f, cmotion):
        return torch.cumsum(cmotion, dim=-1) * (1.0 + action.view(-1, 1, 1, 1).float())

    from types import SimpleNamespace
    online = SimpleNamespace(model=SimpleNamespace(arch="online"))     # as seen through ClassifierFreeSampleModel
    offline = SimpleNamespace(arch="offline")
    for setting in ("cmdm", "sample"):
        want = sampler_ref.auto_regressive(oracle_loop, cm, setting=setting)
        for G, trunc in [(1, False), (1, True), (4, True), (4, False), (9, True), (20, True), (3, None)]:
            calls.clear()
            y = {"cmotion": cm.clone(), "action": action, "action_text": ["a"] * B, "tag": "kept"}
            got = auto_regressive_sample(fake_loop, online, (B, V, C, T), {"y": y}, setting=setting, truncate=trunc,
                                         frames_per_call=G)
            assert torch.equal(got, want), (setting, G, trunc)
            assert torch.equal(y["cmotion"], cm)
            Ge = min(G, T)
            assert len(calls) == (T + Ge - 1) // Ge
            if trunc or trunc is None:     # None resolves to True for the causal arch
                assert [c[-1] for c in calls] == [min(k * Ge + Ge, T) for k in range(len(calls))]
    # truncation is only valid for the causal denoiser: a bidirectional / unknown model must run the literal loop
    y = {"cmotion": cm.clone(), "action": action, "action_text": ["a"] * B, "tag": "kept"}
    for mdl in (offline, None):
        with pytest.raises(ValueError, match="causal"):
            auto_regressive_sample(fake_loop, mdl, (B, V, C, T), {"y": y}, truncate=True)
        calls.clear()
        auto_regressive_sample(fake_loop, mdl, (B, V, C, T), {"y": y}, frames_per_call=2)
        assert all(c[-1] == T for c in calls)
    # per-frame conditioning (motion editing) follows the truncated time axis and the stacked batch
    seen = []

    def edit_loop(model, shape, clip_denoised, model_kwargs):
        yy = model_kwargs["y"]
        assert tuple(yy["inpainting_mask"].shape) == tuple(shape) == tuple(yy["inpainted_motion"].shape)
        seen.append(tuple(shape))
        return yy["cmotion"]
    y = {"cmotion": cm.clone(), "inpainting_mask": torch.zeros(B, V, C, T, dtype=torch.bool),
         "inpainted_motion": torch.zeros(B, V, C, T)}
    auto_regressive_sample(edit_loop, online, (B, V, C, T), {"y": y}, frames_per_call=2)
    assert seen[0] == (2 * B, V, C, 2) and seen[-1] == (B, V, C, T)


def test_oracle_inpainting_blend_matches_reference_expression():
    """oracle Sampler(inpaint=...) restates gaussian_diffusion.py:319-323: where the mask is set the prediction is the
    input motion, elsewhere the model output, and the last step (t=0) returns exactly that blend."""
    import torch
    from oracle import sampler_ref
    g = torch.Generator().manual_seed(1)
    shape = (2, 3, 2, 5)
    motion = torch.randn(*shape, generator=g)
    mask = torch.rand(*shape, generator=g) > 0.5
    model = lambda x, t: 0.5 * x + 1.0
    smp = sampler_ref.Sampler(timestep_respacing="ddim4", inpaint=(mask, motion))
    torch.manual_seed(0)
    out, x0 = smp.loop(model, shape)
    assert torch.equal(out[mask], motion[mask]) and torch.equal(x0[mask], motion[mask])
    assert not torch.equal(out[~mask], motion[~mask])


@pytest.mark.reference
def test_model_util_and_respace_equal_the_imported_reference():
    """Where /root/reference exists (build container): get_model_args returns the reference's kwargs for a matrix of
    command lines, space_timesteps the same sets, and create_gaussian_diffusion the same fp64 tables."""
    from argparse import Namespace
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    import importlib
    ref_mu = importlib.import_module("utils.model_util")
    ref_rs = importlib.import_module("diffusion.respace")
    from regennet_b200 import model_util, respace

    def ds(**kw):
        return type("D", (), {"dataset": type("S", (), kw)})

    base = dict(pose_rep="rot6d", latent_dim=512, layers=8, cond_mask_prob=0.1, cm_mode="concat", wo_pos_emb=False,
                emb_trans_dec=False, setting="cmdm")
    for dataset, body, unc, arch, data in [("ntu", "smplx", True, "online", ds(num_actions=26, num_person=2)),
                                           ("chi3d", "smplx", False, "online", ds(num_actions=8)),
                                           ("ntu", "smpl", False, "offline", ds()),
                                           ("chi3d", "smplx", True, "offline", ds(num_person=2))]:
        args = Namespace(dataset=dataset, body_model=body, unconstrained=unc, arch=arch, **base)
        assert model_util.get_model_args(args, data) == ref_mu.get_model_args(args, data), (dataset, body, unc, arch)
    for n, spec in [(1000, "ddim100"), (1000, "ddim7"), (1000, "10,15,20"), (1000, [250]), (1000, "1000"), (50, "13")]:
        assert respace.space_timesteps(n, spec) == ref_rs.space_timesteps(n, spec), spec
    with pytest.raises(ValueError):
        respace.space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        respace.space_timesteps(10, "6,6")
    dargs = Namespace(noise_schedule="cosine", sigma_small=False, timestep_respacing="ddim20", lambda_vel=0.0,
                      lambda_rcxyz=0.0, lambda_fc=0.0, lambda_orient=0.0, lambda_body=0.0, lambda_transl=0.0,
                      pose_rep="rot6d", num_person=1, body_model="smplx", vel_threshold=0.01)
    ours, ref = model_util.create_gaussian_diffusion(dargs), ref_mu.create_gaussian_diffusion(dargs)
    assert ours.timestep_map == ref.timestep_map and ours.model_var_type.name == ref.model_var_type.name
    for f in ["betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"]:
        assert np.array_equal(getattr(ours, f), getattr(ref, f)), f


def test_results_writer_matches_cgenerate_format(tmp_path):
    """sample/cgenerate.py:167-192: concatenation over repetitions, cut to num_samples * num_repetitions, pickled dict with
    the reference's keys, caption / length side files, existing directory replaced."""
    from regennet_b200.results import ResultsWriter, load_results
    bs, J, F, T = 3, 56, 6, 8
    w = ResultsWriter(num_samples=2, num_repetitions=2)     # batches of 3 are cut to 2 * 2 = 4 entries in total
    g = torch.Generator().manual_seed(0)
    reps = []
    for rep in range(2):
        motion = torch.randn(bs, J, 3, T, generator=g)
        output = torch.randn(bs, J, F, T, generator=g)
        cmotion = torch.randn(bs, J, F, T, generator=g)
        lengths = torch.tensor([T, T - 1, T - 2])
        text = ["a%d_%d" % (rep, i) for i in range(bs)]
        w.add(motion, output, cmotion, lengths, text)
        reps.append((motion, output, cmotion, lengths, text))
    out_dir = tmp_path / "samples"
    out_dir.mkdir()
    (out_dir / "stale.txt").write_text("old")               # the reference rmtree()s an existing output directory
    path = w.save(str(out_dir))
    assert not (out_dir / "stale.txt").exists()
    r = load_results(path)
    assert sorted(r) == sorted(['motion', 'output', 'cmotion', 'text', 'lengths', 'num_samples', 'num_repetitions'])
    assert r['motion'].shape == (4, J, 3, T) and r['output'].shape == (4, J, F, T) and r['cmotion'].shape == (4, J, F, T)
    want = np.concatenate([reps[0][1].numpy(), reps[1][1].numpy()], axis=0)[:4]
    assert np.array_equal(r['output'], want)
    assert r['text'] == ["a0_0", "a0_1", "a0_2", "a1_0"]
    assert r['lengths'].tolist() == [T, T - 1, T - 2, T]
    assert r['num_samples'] == 2 and r['num_repetitions'] == 2
    assert (out_dir / "results.txt").read_text().split("\n") == r['text']
    assert (out_dir / "results_len.txt").read_text().split("\n") == [str(T), str(T - 1), str(T - 2), str(T)]
    with pytest.raises(ValueError):
        ResultsWriter(1, 1).save(str(tmp_path / "empty"))


def test_stgcn_graph_adjacency_bit_exact():
    """regennet_b200/stgcn_graph.py against the reference's Graph(...).A (stgcnutils/graph.py) stored by
    tests/golden/make_golden_stgcn.py: every file-free layout x strategy x max_hop, and 'smpl' through a kinematic tree."""
    from regennet_b200 import stgcn_graph
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stgcn.npz"))
    n = 0
    for layout in cases.STGCN_GRAPH_LAYOUTS:
        for strategy in ("uniform", "distance", "spatial"):
            for hop in (1, 2):
                want = g["graph.%s.%s.%d" % (layout, strategy, hop)]
                got = stgcn_graph.adjacency(layout, strategy, max_hop=hop)
                assert got.shape == want.shape and np.array_equal(got, want), (layout, strategy, hop)
                n += 1
    assert n == 18
    kt = np.stack([np.array(cases.STGCN_SMPL_PARENTS), np.arange(24)])
    assert np.array_equal(stgcn_graph.adjacency("smpl", "spatial", kintree=kt), g["graph.smpl.spatial.1"])
    # the evaluation's own layout: 55 joints + the translation node, tree of cases.STGCN_SMPLX_PARENTS (the reference read it
    # from a temporary SMPLX_NEUTRAL.npz when the golden was written)
    ktx = np.stack([np.array(cases.STGCN_SMPLX_PARENTS), np.arange(55)])
    Ax = stgcn_graph.adjacency("smplx", "spatial", kintree=ktx)
    assert Ax.shape == (3, 56, 56) and np.array_equal(Ax, g["graph.smplx.spatial.1"])
    assert np.array_equal(g["stgcn_p2.A"], stgcn_graph.adjacency("ntu-rgb+d", "spatial").astype(np.float32))
    with pytest.raises(ValueError):
        stgcn_graph.adjacency("smplx", "spatial")          # needs the body model's kinematic tree
    with pytest.raises(NotImplementedError):
        stgcn_graph.adjacency("nope")


def test_eval_metrics_match_reference_goldens():
    """regennet_b200/eval_metrics.py against eval/a2m/stgcn/{fid,diversity,accuracy}.py run on the same seeded features
    (tests/golden/make_golden_stgcn.py); the pair sampling consumes numpy's global generator in the reference's order."""
    from regennet_b200 import eval_metrics as em
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stgcn.npz"))
    f1, f2, labels = cases.metric_inputs()
    s1, s2 = em.calculate_activation_statistics(f1), em.calculate_activation_statistics(f2)
    assert abs(em.calculate_fid(s1, s2) - float(g["metrics.fid"])) < 1e-9 * max(1.0, abs(float(g["metrics.fid"])))
    assert abs(em.calculate_fid(s1, s1) - float(g["metrics.fid_self"])) < 1e-6
    div, mm = em.calculate_diversity_multimodality(f1, labels, cases.METRIC_LABELS, seed=7)
    assert np.allclose([div, mm], g["metrics.div_mm"], rtol=0, atol=1e-6)
    loader = [{"yhat": f1[i:i + 40, :cases.METRIC_LABELS], "y": labels[i:i + 40]} for i in range(0, f1.shape[0], 40)]
    acc, conf = em.calculate_accuracy(None, loader, cases.METRIC_LABELS, lambda b: b, "cpu")
    assert abs(acc - float(g["metrics.accuracy"])) < 1e-7
    assert np.array_equal(conf.numpy(), g["metrics.confusion"])       # integer counts: bit-exact

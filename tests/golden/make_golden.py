"""Generate golden vectors by running the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Inputs are not stored: they are re-created from seeds by
regennet_b200.synthetic (torch CPU generators are deterministic for a given torch build);
only reference OUTPUTS are stored.  While writing, the script also checks the oracle
restatement against the reference and prints the max abs differences.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import cmdm_ref, ref_shim, sampler_ref, schedule  # noqa: E402
from regennet_b200 import synthetic  # noqa: E402
import cases  # noqa: E402


def ref_y(y, model_name):
    """our y -> the reference's y (text features travel as y['text'] through the identity encode_text)."""
    ry = {k: v for k, v in y.items() if k != "text_embed"}
    if "text_embed" in y:
        ry["text"] = y["text_embed"]
    return ry


def main():
    torch.set_num_threads(8)
    ref_shim.install()
    from model.cfg_sampler import ClassifierFreeSampleModel
    from diffusion import gaussian_diffusion as gd
    from diffusion.respace import SpacedDiffusion, space_timesteps
    from utils.rotation_conversions import rotation_6d_to_matrix

    models = {}

    def get_model(name, wseed):
        key = (name, wseed)
        if key not in models:
            m, _ = ref_shim.build_reference(cases.MODELS[name], {})
            sd = synthetic.make_state_dict(seed=wseed, **cases.synth_kw(name))
            missing, unexpected = m.load_state_dict(sd, strict=False)
            assert not unexpected, unexpected
            assert all(k.startswith("clip_model.") for k in missing), missing
            models[key] = (m, sd)
        return models[key]

    out = {}
    # ---- forward cases -------------------------------------------------------------------------
    for name, c in cases.FORWARD_CASES.items():
        mk = cases.MODELS[c["model"]]
        model, sd = get_model(c["model"], c["wseed"])
        x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                     cond_mode=mk["cond_mode"], num_actions=mk["num_actions"],
                                     scale=c.get("cfg_scale"))
        t = torch.tensor(c["t"], dtype=torch.long)
        if c.get("uncond"):
            y["uncond"] = True
        kw = dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])
        with torch.no_grad():
            if "cfg_scale" in c:
                ref = ClassifierFreeSampleModel(model)(x, t, ref_y(y, c["model"]))
                ora = cmdm_ref.cfg_forward(sd, x, t, y, **kw)
            else:
                ref = model(x, t, ref_y(y, c["model"]))
                ora = cmdm_ref.cmdm_forward(sd, x, t, y, **kw)
        print("%-24s ref absmax %.3f  oracle-vs-ref max abs %.3e" % (name, ref.abs().max(), (ref - ora).abs().max()))
        out[name] = ref.numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "forward.npz"), **out)

    # ---- sampling loops ------------------------------------------------------------------------
    out = {}
    for name, c in cases.LOOP_CASES.items():
        mk = cases.MODELS[c["model"]]
        model, sd = get_model(c["model"], c["wseed"])
        _, diffusion = ref_shim.build_reference(cases.MODELS["ntu"], dict(timestep_respacing=c["respacing"])) \
            if False else (None, None)
        # build only the diffusion (cheap) with the requested respacing
        from argparse import Namespace
        from utils.model_util import create_gaussian_diffusion
        args = Namespace(noise_schedule="cosine", sigma_small=True, timestep_respacing=c["respacing"],
                         lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0, lambda_orient=0.0, lambda_body=0.0,
                         lambda_transl=0.0, pose_rep="rot6d", num_person=1, body_model="smplx", vel_threshold=0.01)
        diffusion = create_gaussian_diffusion(args)
        _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                     cond_mode=mk["cond_mode"], num_actions=mk["num_actions"],
                                     scale=c.get("cfg_scale"))
        shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
        run_model = ClassifierFreeSampleModel(model) if "cfg_scale" in c else model
        fn = diffusion.ddim_sample_loop if c["ddim"] else diffusion.p_sample_loop
        torch.manual_seed(c["seed"])
        ref = fn(run_model, shape, clip_denoised=False, model_kwargs={"y": ref_y(y, c["model"])}, device="cpu")
        # oracle on the same RNG stream
        kw = dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])
        smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
        assert smp.timestep_map == diffusion.timestep_map
        fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
        torch.manual_seed(c["seed"])
        ora, _ = smp.loop(lambda xx, tt: fwd(sd, xx, tt, y, **kw), shape, ddim=c["ddim"])
        print("%-24s steps %4d ref absmax %.3f  oracle-vs-ref max abs %.3e" %
              (name, diffusion.num_timesteps, ref.abs().max(), (ref - ora).abs().max()))
        out[name] = ref.numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "loops.npz"), **out)

    # ---- schedule tables / respacing (fp64 + integer, bit-exact) -----------------------------
    out = {}
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    assert np.array_equal(betas, schedule.named_beta_schedule("cosine", 1000))
    out["betas_cosine_1000"] = betas
    out["betas_linear_1000"] = gd.get_named_beta_schedule("linear", 1000, 1.0)
    for rs in cases.RESPACINGS:
        use = space_timesteps(1000, rs if rs else [1000])
        d = SpacedDiffusion(use_timesteps=use, betas=betas, model_mean_type=gd.ModelMeanType.START_X,
                            model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
        tag = "rs[%s]" % rs
        out[tag + ".timestep_map"] = np.array(d.timestep_map, dtype=np.int64)
        for f in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"]:
            out[tag + "." + f] = getattr(d, f)
        tab, tmap = schedule.spaced_tables(betas, schedule.space_timesteps(1000, rs if rs else [1000]))
        assert tmap == d.timestep_map
        assert np.array_equal(tab.posterior_mean_coef1, d.posterior_mean_coef1)
    np.savez(os.path.join(HERE, "schedule.npz"), **out)
    print("schedule: %d arrays" % len(out))

    # ---- rot6d -> rotmat ---------------------------------------------------------------------
    g = torch.Generator().manual_seed(99)
    d6 = torch.randn(4, 7, 55, 6, generator=g)
    d6[0, 0, 0] = 0.0                       # degenerate: zero vector (F.normalize eps path)
    d6[0, 0, 1, 3:] = d6[0, 0, 1, :3] * 2   # degenerate: a2 parallel to a1
    d6[0, 0, 2] *= 1e-20                    # tiny norm
    d6[0, 0, 3] *= 1e6                      # large norm
    ref = rotation_6d_to_matrix(d6)
    ora = sampler_ref.rotation_6d_to_matrix(d6)
    same = torch.isclose(ref, ora, atol=1e-6, equal_nan=True).all().item()
    print("rot6d: oracle-vs-ref allclose(1e-6, nan==nan):", same)
    np.savez(os.path.join(HERE, "rot6d.npz"), d6=d6.numpy(), R=ref.numpy())


if __name__ == "__main__":
    main()

"""Golden vectors for arch=online, cm_mode=add from the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden_add.py

Writes tests/golden/forward_add.npz and loops_add.npz (reference OUTPUTS only; inputs are re-created from
seeds by regennet_b200.synthetic) and prints the oracle-vs-reference differences.  Kept separate from
make_golden.py so the other goldens are never rewritten.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import cmdm_ref, ref_shim, sampler_ref  # noqa: E402
from regennet_b200 import synthetic  # noqa: E402
import cases  # noqa: E402
from make_golden import ref_y  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref_shim.install()
    from model.cfg_sampler import ClassifierFreeSampleModel
    models = {}

    def get_model(name, wseed):
        key = (name, wseed)
        if key not in models:
            m, _ = ref_shim.build_reference(cases.ADD_MODELS[name], {})
            sd = synthetic.make_state_dict(seed=wseed, **cases.synth_kw_add(name))
            missing, unexpected = m.load_state_dict(sd, strict=False)
            assert not unexpected, unexpected
            assert all(k.startswith("clip_model.") for k in missing), missing
            models[key] = (m, sd)
        return models[key]

    def kw_of(mk):
        return dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"],
                    cm_mode=mk["cm_mode"])

    out = {}
    for name, c in cases.ADD_FORWARD_CASES.items():
        mk = cases.ADD_MODELS[c["model"]]
        model, sd = get_model(c["model"], c["wseed"])
        x, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                     cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
        t = torch.tensor(c["t"], dtype=torch.long)
        with torch.no_grad():
            if "cfg_scale" in c:
                ref = ClassifierFreeSampleModel(model)(x, t, ref_y(y, c["model"]))
                ora = cmdm_ref.cfg_forward(sd, x, t, y, **kw_of(mk))
            else:
                ref = model(x, t, ref_y(y, c["model"]))
                ora = cmdm_ref.cmdm_forward(sd, x, t, y, **kw_of(mk))
        print("%-24s ref absmax %.3f  oracle-vs-ref max abs %.3e" % (name, ref.abs().max(), (ref - ora).abs().max()))
        out[name] = ref.numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "forward_add.npz"), **out)

    out = {}
    from argparse import Namespace
    from utils.model_util import create_gaussian_diffusion
    for name, c in cases.ADD_LOOP_CASES.items():
        mk = cases.ADD_MODELS[c["model"]]
        model, sd = get_model(c["model"], c["wseed"])
        args = Namespace(noise_schedule="cosine", sigma_small=True, timestep_respacing=c["respacing"],
                         lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0, lambda_orient=0.0, lambda_body=0.0,
                         lambda_transl=0.0, pose_rep="rot6d", num_person=1, body_model="smplx", vel_threshold=0.01)
        diffusion = create_gaussian_diffusion(args)
        _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                     cond_mode=mk["cond_mode"], num_actions=mk["num_actions"])
        shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
        torch.manual_seed(c["seed"])
        ref = diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": ref_y(y, c["model"])},
                                      device="cpu")
        smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
        torch.manual_seed(c["seed"])
        ora, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **kw_of(mk)), shape)
        print("%-24s steps %4d ref absmax %.3f  oracle-vs-ref max abs %.3e" %
              (name, diffusion.num_timesteps, ref.abs().max(), (ref - ora).abs().max()))
        out[name] = ref.numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "loops_add.npz"), **out)


if __name__ == "__main__":
    main()

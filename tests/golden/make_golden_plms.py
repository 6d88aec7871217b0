"""Golden vectors for the PLMS sampler from the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden_plms.py

Writes tests/golden/loops_plms.npz (reference OUTPUTS of plms_sample_loop, diffusion/gaussian_diffusion.py:1100-1202;
inputs are re-created from seeds by regennet_b200.synthetic) and prints the oracle-vs-reference differences.
Kept separate from make_golden.py so the other goldens are never rewritten.
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
    out = {}
    for name, c in cases.PLMS_LOOP_CASES.items():
        mk = cases.MODELS[c["model"]]
        key = (c["model"], c["wseed"])
        if key not in models:
            m, _ = ref_shim.build_reference(mk, {})
            sd = synthetic.make_state_dict(seed=c["wseed"], **cases.synth_kw(c["model"]))
            missing, unexpected = m.load_state_dict(sd, strict=False)
            assert not unexpected and all(k.startswith("clip_model.") for k in missing)
            models[key] = (m, sd)
        model, sd = models[key]
        _, diffusion = ref_shim.build_reference(cases.MODELS["ntu"], dict(timestep_respacing=c["respacing"]))
        _, y = synthetic.make_inputs(c["B"], mk["njoints"], mk["nfeats"], c["T"], seed=c["xseed"],
                                     cond_mode=mk["cond_mode"], num_actions=mk["num_actions"], scale=c.get("cfg_scale"))
        shape = (c["B"], mk["njoints"], mk["nfeats"], c["T"])
        run_model = ClassifierFreeSampleModel(model) if "cfg_scale" in c else model
        clip = bool(c.get("clip"))
        torch.manual_seed(c["seed"])
        ref = diffusion.plms_sample_loop(run_model, shape, clip_denoised=clip, model_kwargs={"y": ref_y(y, c["model"])},
                                         device="cpu", order=c["order"])
        kw = dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])
        fwd = cmdm_ref.cfg_forward if "cfg_scale" in c else cmdm_ref.cmdm_forward
        smp = sampler_ref.Sampler(timestep_respacing=c["respacing"])
        assert smp.timestep_map == diffusion.timestep_map
        torch.manual_seed(c["seed"])
        ora = smp.plms_loop(lambda xx, tt: fwd(sd, xx, tt, y, **kw), shape, order=c["order"], clip_denoised=clip)
        print("%-24s steps %4d order %d ref absmax %.3f  oracle-vs-ref max abs %.3e" %
              (name, diffusion.num_timesteps, c["order"], ref.abs().max(), (ref - ora).abs().max()))
        out[name] = ref.numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "loops_plms.npz"), **out)


if __name__ == "__main__":
    main()

"""Import shim for the UNMODIFIED reference (liangxuy/ReGenNet) on CPU.

Test / measurement infrastructure, never imported by the product path.  The reference tree is taken from
``REGEN_REFERENCE_ROOT``, else /root/reference (the build container; used by tests/golden/make_golden*.py to generate
golden vectors and by the ``reference``-marked tests), else ``oracle/_ref`` -- the 17 unmodified reference files of the
sampling path staged by ``oracle/make_ref.sh`` (git-ignored; travels to the GPU box, where it lets ``bench.py --impl
reference`` time the reference's own p_sample_loop on the host cores and on the GPU as the torch-eager baseline).

The reference imports three packages that are absent here and irrelevant to the
hot path (timm's DropPath, OpenAI clip, smplx body-model layers) and uses the
numpy aliases removed in numpy>=1.24; those are stubbed, nothing else is touched.
"""
import os
import sys
import types

import numpy as np
import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _find_root():
    for cand in (os.environ.get("REGEN_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isdir(os.path.join(cand, "model")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def install():
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float"):
        np.float = float  # data_loaders/humanml/common/quaternion.py:13
    if not hasattr(np, "int"):
        np.int = int

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _DropPath(nn.Identity):
        def __init__(self, p=0.0):
            super().__init__()

    class _Layer(nn.Module):  # stands in for smplx.SMPLLayer / SMPLXLayer
        def __init__(self, *a, **k):
            super().__init__()

    if "timm" not in sys.modules:
        layers = _mod("timm.models.layers", DropPath=_DropPath)
        _mod("timm.models", layers=layers)
        _mod("timm", models=sys.modules["timm.models"])
    if "clip" not in sys.modules:
        _mod("clip")
    if "smplx" not in sys.modules:
        lbs = _mod("smplx.lbs", vertices2joints=lambda *a, **k: None)
        _mod("smplx", SMPLLayer=_Layer, SMPLXLayer=_Layer, lbs=lbs)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def build_reference(model_kw, diffusion_kw):
    """-> (reference CMDM in eval mode, reference SpacedDiffusion)."""
    install()
    import contextlib
    import io
    from argparse import Namespace

    from model.cmdm import CMDM
    from utils.model_util import create_gaussian_diffusion

    with contextlib.redirect_stdout(io.StringIO()):
        if "text" in model_kw.get("cond_mode", ""):
            # CLIP weights are unavailable; text features are injected (see encode_text patch below)
            CMDM.load_and_freeze_clip = lambda self, v: nn.Identity()
        model = CMDM(**model_kw)
    model.eval()
    if "text" in model_kw.get("cond_mode", ""):
        model.encode_text = lambda feats: feats  # y['text'] carries precomputed [B,512] features
    args = Namespace(noise_schedule="cosine", sigma_small=True, timestep_respacing="",
                     lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0, lambda_orient=0.0, lambda_body=0.0,
                     lambda_transl=0.0, pose_rep="rot6d", num_person=1, body_model="smplx", vel_threshold=0.01)
    for k, v in diffusion_kw.items():
        setattr(args, k, v)
    return model, create_gaussian_diffusion(args)

"""regennet_b200 -- B200-native (sm_100a) implementation of ReGenNet's diffusion sampling hot path.

Public surface mirrors the reference modules it replaces:

    reference module                       this package
    diffusion/gaussian_diffusion.py   ->   regennet_b200.gaussian_diffusion
    diffusion/respace.py              ->   regennet_b200.respace
    model/cmdm.py                     ->   regennet_b200.cmdm
    model/cfg_sampler.py              ->   regennet_b200.cfg_sampler
    utils/rotation_conversions.py     ->   regennet_b200.rotation_conversions
    utils/model_util.py               ->   regennet_b200.model_util

``install_as_reference_modules()`` registers those aliases in ``sys.modules`` so the reference's
own ``sample/cgenerate.py`` / ``eval/eval_cmdm.py`` import this implementation unchanged
(see INTEGRATION.md).
"""

__version__ = "0.1.0"

_ALIASES = {
    "diffusion.gaussian_diffusion": "regennet_b200.gaussian_diffusion",
    "diffusion.respace": "regennet_b200.respace",
    "model.cmdm": "regennet_b200.cmdm",
    "model.cfg_sampler": "regennet_b200.cfg_sampler",
    "utils.rotation_conversions": "regennet_b200.rotation_conversions",
    "utils.model_util": "regennet_b200.model_util",
}


def install_as_reference_modules():
    """Make ``from model.cmdm import CMDM`` etc. resolve to this package (drop-in for the
    reference's entry scripts).  Parent packages of the reference (``model``, ``diffusion``,
    ``utils``) keep working for everything that is not on the hot path."""
    import importlib
    import sys
    for ref_name, ours in _ALIASES.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent, _, child = ref_name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)

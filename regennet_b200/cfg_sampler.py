"""Classifier-free guidance around a CMDM (API of the reference's ``model/cfg_sampler.py``).

The reference evaluates the denoiser twice per step -- conditioned, and with ``y['uncond'] = True`` on a deep copy of
``y`` -- and returns ``uncond + y['scale'] * (cond - uncond)`` (model/cfg_sampler.py:24-31).  Here both halves are ONE
doubled batch inside the library (rows [0, B) conditional, [B, 2B) unconditional) and the combination is a library
kernel with the same operation order.
"""
import torch.nn as nn

# attributes callers read from the wrapper as if it were the model (sample/cgenerate.py, eval/a2m/stgcn/evaluate.py)
_FORWARDED = ("rot2xyz", "translation", "njoints", "nfeats", "data_rep", "cond_mode")


class ClassifierFreeSampleModel(nn.Module):

    def __init__(self, model):
        super().__init__()
        self.model = model
        assert self.model.cond_mask_prob > 0, \
            'Cannot run a guided diffusion on a model that has not been trained with no conditions'
        for name in _FORWARDED:
            setattr(self, name, getattr(model, name))

    def _check(self):
        assert self.model.cond_mode in ['text', 'action']

    def forward(self, x, timesteps, y=None):
        self._check()
        return self.model._forward_impl(x, timesteps, y, scale=y['scale'])

    def regen_sampling_session(self, shape, y, timestep_map):
        """Hook of the fused sampling route (GaussianDiffusion._fast_session)."""
        self._check()
        return self.model.regen_sampling_session(shape, y, timestep_map, scale=y['scale'])

"""Classifier-free guidance wrapper -- mirror of the reference's ``model/cfg_sampler.py``.

The reference runs two sequential forwards (conditional, unconditional) with a ``deepcopy(y)`` per
step and combines them in PyTorch (model/cfg_sampler.py:24-31).  Here both halves run as ONE
doubled batch inside the library (rows [0,B) conditional, [B,2B) unconditional) and the combine
``u + s*(c-u)`` is a library kernel; the result is identical arithmetic.
"""
import torch.nn as nn


class ClassifierFreeSampleModel(nn.Module):

    def __init__(self, model):
        super().__init__()
        self.model = model  # model is the actual model to run

        assert self.model.cond_mask_prob > 0, \
            'Cannot run a guided diffusion on a model that has not been trained with no conditions'

        # pointers to inner model
        self.rot2xyz = self.model.rot2xyz
        self.translation = self.model.translation
        self.njoints = self.model.njoints
        self.nfeats = self.model.nfeats
        self.data_rep = self.model.data_rep
        self.cond_mode = self.model.cond_mode

    def forward(self, x, timesteps, y=None):
        cond_mode = self.model.cond_mode
        assert cond_mode in ['text', 'action']
        return self.model._forward_impl(x, timesteps, y, scale=y['scale'])

    def regen_sampling_session(self, shape, y, timestep_map):
        assert self.model.cond_mode in ['text', 'action']
        return self.model.regen_sampling_session(shape, y, timestep_map, scale=y['scale'])

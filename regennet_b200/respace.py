"""Timestep respacing for the sampler (API of the reference's ``diffusion/respace.py``).

A respaced process keeps a subset ``S`` of the base process' timesteps.  Its betas follow from the base
cumulative products, ``beta'_k = 1 - abar[S_k] / abar[S_{k-1}]`` (respace.py:73-87), and the denoiser is always
called with ORIGINAL timesteps: ``timestep_map[k] = S_k``.  That remap is an int64 gather (respace.py:125-126) and is
bit-exact here -- the fused route applies it on the host / in ``regen_step_tables`` (all samples of a step share
the index), the generic route through ``_WrappedModel`` like the reference.

Public names and call signatures are the reference's: ``space_timesteps(num_timesteps, section_counts)``,
``SpacedDiffusion(use_timesteps, **gaussian_diffusion_kwargs)``.
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def _ddim_stride_steps(num_timesteps, wanted):
    """'ddimK': the smallest integer stride whose arithmetic progression has exactly K members (respace.py:29-37)."""
    for stride in range(1, num_timesteps):
        steps = range(0, num_timesteps, stride)
        if len(steps) == wanted:
            return set(steps)
    raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")


def _section_steps(first, length, count):
    """`count` steps spread over the section [first, first + length): positions advance by a fractional stride that is
    ACCUMULATED (not multiplied) and rounded half-to-even, exactly the float sequence of respace.py:49-58."""
    if length < count:
        raise ValueError(f"cannot divide section of {length} steps into {count}")
    stride = 1 if count <= 1 else (length - 1) / (count - 1)
    picked, pos = [], 0.0
    while len(picked) < count:
        picked.append(first + round(pos))
        pos += stride
    return picked


def space_timesteps(num_timesteps, section_counts):
    """Timesteps to keep (a set).  ``section_counts`` is a list of per-section counts, the same as a comma separated
    string, or ``"ddimK"`` for the fixed-stride DDIM spacing (respace.py:8-61)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            return _ddim_stride_steps(num_timesteps, int(section_counts[len("ddim"):]))
        section_counts = [int(tok) for tok in section_counts.split(",")]
    n_sec = len(section_counts)
    base, longer = divmod(num_timesteps, n_sec)      # the first `longer` sections get one extra step
    kept, first = [], 0
    for k, count in enumerate(section_counts):
        length = base + (1 if k < longer else 0)
        kept.extend(_section_steps(first, length, count))
        first += length
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process over the timesteps in ``use_timesteps`` of the base process described by ``kwargs``."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        abar = GaussianDiffusion(**kwargs).alphas_cumprod            # fp64 tables of the base process
        self.timestep_map = [i for i in range(len(abar)) if i in self.use_timesteps]
        prev = np.concatenate(([1.0], abar[self.timestep_map[:-1]])) if self.timestep_map else np.array([])
        kwargs["betas"] = 1 - abar[self.timestep_map] / prev       # same fp64 operations as respace.py:80-83
        super().__init__(**kwargs)

    # every entry point that hands the model (or a guidance function) timesteps goes through the remapping wrapper
    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def _call_model(self, model, *args, **kwargs):
        return super()._call_model(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t   # the wrapper rescales (respace.py:113-115)

    def _timestep_map_for_model(self):
        return list(self.timestep_map)


class _WrappedModel:
    """Callable that converts respaced step indices to original timesteps before calling the model (respace.py:117-129)."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._device_maps = {}   # (device, dtype) -> tensor; the reference re-uploads the list on every call

    def _map_on(self, like):
        key = (str(like.device), like.dtype)
        if key not in self._device_maps:
            self._device_maps[key] = th.tensor(self.timestep_map, device=like.device, dtype=like.dtype)
        return self._device_maps[key]

    def __call__(self, x, ts, **kwargs):
        original_ts = self._map_on(ts)[ts]
        if self.rescale_timesteps:
            original_ts = original_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, original_ts, **kwargs)

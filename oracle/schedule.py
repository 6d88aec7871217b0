"""Oracle (test infrastructure): fp64 diffusion schedule tables and respacing.

Restates, in numpy float64 exactly as the reference computes them on the host:
  * diffusion/gaussian_diffusion.py:21-65   named beta schedules (cosine / linear)
  * diffusion/gaussian_diffusion.py:172-209 coefficient tables
  * diffusion/respace.py:8-61               space_timesteps
  * diffusion/respace.py:73-87              SpacedDiffusion betas / timestep_map
"""
import math

import numpy as np


def named_beta_schedule(name, n, scale_betas=1.0):
    # gaussian_diffusion.py:21-45
    if name == "linear":
        scale = scale_betas * 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if name == "cosine":
        # gaussian_diffusion.py:48-65 (betas_for_alpha_bar, max_beta=0.999)
        def alpha_bar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

        out = []
        for i in range(n):
            t1 = i / n
            t2 = (i + 1) / n
            out.append(min(1 - alpha_bar(t2) / alpha_bar(t1), 0.999))
        return np.array(out)
    raise NotImplementedError(name)


def space_timesteps(num_timesteps, section_counts):
    # respace.py:8-61
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError("no integer stride gives %d steps" % want)
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start = 0
    steps = []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError("cannot divide section of %d steps into %d" % (size, count))
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += frac
        start += size
    return set(steps)


class Tables:
    """fp64 tables of one (possibly respaced) diffusion process."""

    def __init__(self, betas):
        # gaussian_diffusion.py:172-209
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def spaced_tables(base_betas, use_timesteps):
    """respace.py:73-87 -> (Tables of the respaced process, timestep_map list)."""
    base = Tables(base_betas)
    use = set(use_timesteps)
    last = 1.0
    new_betas = []
    tmap = []
    for i, ac in enumerate(base.alphas_cumprod):
        if i in use:
            new_betas.append(1 - ac / last)
            last = ac
            tmap.append(i)
    return Tables(np.array(new_betas)), tmap

"""Oracle (test infrastructure): ancestral / DDIM sampling loops and rot6d->rotmat.

Follows
  * diffusion/gaussian_diffusion.py:289-400  p_mean_variance (START_X, FIXED_SMALL/LARGE)
  * diffusion/gaussian_diffusion.py:265-287  q_posterior_mean_variance
  * diffusion/gaussian_diffusion.py:508-560  p_sample
  * diffusion/gaussian_diffusion.py:675-742  p_sample_loop_progressive
  * diffusion/gaussian_diffusion.py:744-794  ddim_sample
  * diffusion/gaussian_diffusion.py:1007-1098, 1132-1202  plms_sample / plms_sample_loop
  * diffusion/gaussian_diffusion.py:1604-1617 _extract_into_tensor (fp64 gather, THEN cast to fp32)
  * diffusion/respace.py:117-129             _WrappedModel timestep remap (integer gather)
  * utils/rotation_conversions.py:513-534    rotation_6d_to_matrix
"""
import numpy as np
import torch

from . import schedule


def _extract(arr, t, ndim):
    res = torch.from_numpy(arr)[t].float()
    while res.dim() < ndim:
        res = res[..., None]
    return res


class Sampler:
    """SpacedDiffusion restated: START_X mean, fixed variance, clip_denoised off by default."""

    def __init__(self, noise_schedule="cosine", steps=1000, timestep_respacing="", sigma_small=True, inpaint=None):
        # inpaint = (mask bool, motion): diffusion/gaussian_diffusion.py:319-323, the model output is overwritten
        # where the mask is set (y['inpainting_mask'], y['inpainted_motion']; sample/edit.py:75-90)
        self.inpaint = inpaint
        base = schedule.named_beta_schedule(noise_schedule, steps)
        use = schedule.space_timesteps(steps, timestep_respacing if timestep_respacing else [steps])
        self.tab, self.timestep_map = schedule.spaced_tables(base, use)
        self.num_timesteps = self.tab.num_timesteps
        self.sigma_small = sigma_small

    def remap(self, t):
        # respace.py:125-126, integer gather
        return torch.tensor(self.timestep_map, dtype=t.dtype)[t]

    def p_mean_variance(self, model, x, t, clip_denoised=False):
        tab = self.tab
        x0 = model(x, self.remap(t))
        if self.inpaint is not None:
            mask, motion = self.inpaint
            x0 = (x0 * ~mask) + (motion * mask)
        if self.sigma_small:
            logvar = tab.posterior_log_variance_clipped
        else:  # FIXED_LARGE, gaussian_diffusion.py:345-350
            logvar = np.log(np.append(tab.posterior_variance[1], tab.betas[1:]))
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        mean = _extract(tab.posterior_mean_coef1, t, x.dim()) * x0 + _extract(tab.posterior_mean_coef2, t, x.dim()) * x
        return mean, _extract(logvar, t, x.dim()), x0

    def p_sample(self, model, x, t, noise_fn, clip_denoised=False):
        mean, logvar, x0 = self.p_mean_variance(model, x, t, clip_denoised)
        noise = noise_fn(x)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        return mean + nonzero * torch.exp(0.5 * logvar) * noise, x0

    def ddim_sample(self, model, x, t, noise_fn, eta=0.0, clip_denoised=False):
        tab = self.tab
        _, _, x0 = self.p_mean_variance(model, x, t, clip_denoised)
        nd = x.dim()
        eps = (_extract(tab.sqrt_recip_alphas_cumprod, t, nd) * x - x0) / _extract(tab.sqrt_recipm1_alphas_cumprod, t, nd)
        ab = _extract(tab.alphas_cumprod, t, nd)
        abp = _extract(tab.alphas_cumprod_prev, t, nd)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        noise = noise_fn(x)
        mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
        nonzero = (t != 0).float().view(-1, *([1] * (nd - 1)))
        return mean + nonzero * sigma * noise, x0

    def plms_sample(self, model, x, t, order=2, old_out=None, clip_denoised=False):
        """diffusion/gaussian_diffusion.py:1007-1098 (no cond_fn): pseudo improved Euler for the first step when
        order > 1, Adams-Bashforth of order min(order, len(history)) afterwards."""
        tab = self.tab
        nd = x.dim()

        def model_eps(xx, tt):
            _, _, x0 = self.p_mean_variance(model, xx, tt, clip_denoised)
            eps = (_extract(tab.sqrt_recip_alphas_cumprod, tt, nd) * xx - x0) / _extract(tab.sqrt_recipm1_alphas_cumprod, tt, nd)
            return eps, x0

        def x0_from_eps(eps):
            return _extract(tab.sqrt_recip_alphas_cumprod, t, nd) * x - _extract(tab.sqrt_recipm1_alphas_cumprod, t, nd) * eps

        abp = _extract(tab.alphas_cumprod_prev, t, nd)
        eps, x0 = model_eps(x, t)
        if order > 1 and old_out is None:
            old_eps = [eps]
            mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp) * eps
            eps_2, _ = model_eps(mean_pred, t - 1)
            eps_prime = (eps + eps_2) / 2
        else:
            old_eps = old_out["old_eps"]
            old_eps.append(eps)
            k = min(order, len(old_eps))
            if k == 1:
                eps_prime = old_eps[-1]
            elif k == 2:
                eps_prime = (3 * old_eps[-1] - old_eps[-2]) / 2
            elif k == 3:
                eps_prime = (23 * old_eps[-1] - 16 * old_eps[-2] + 5 * old_eps[-3]) / 12
            else:
                eps_prime = (55 * old_eps[-1] - 59 * old_eps[-2] + 37 * old_eps[-3] - 9 * old_eps[-4]) / 24
        mean_pred = x0_from_eps(eps_prime) * torch.sqrt(abp) + torch.sqrt(1 - abp) * eps_prime
        if len(old_eps) >= order:
            old_eps.pop(0)
        nonzero = (t != 0).float().view(-1, *([1] * (nd - 1)))
        return {"sample": mean_pred * nonzero + x0 * (1 - nonzero), "pred_xstart": x0, "old_eps": old_eps}

    def plms_loop(self, model, shape, order=2, clip_denoised=False, init_noise=None):
        """:1132-1202."""
        img = init_noise if init_noise is not None else torch.randn(*shape)
        old = None
        with torch.no_grad():
            for i in list(range(self.num_timesteps))[::-1]:
                old = self.plms_sample(model, img, torch.tensor([i] * shape[0]), order, old, clip_denoised)
                img = old["sample"]
        return img

    def loop(self, model, shape, noise_fn=None, ddim=False, eta=0.0, clip_denoised=False, init_noise=None):
        """model(x, t_original) -> x0.  Noise is drawn as the reference does: th.randn(*shape) for x_N
        (gaussian_diffusion.py:706), then th.randn_like(x) once per step AFTER the model call, including
        t == 0 (:544, :781).  randn_like inherits x's memory layout, and x inherits the permuted layout of
        the model output from the second step on (the first operand of `coef1*x0 + coef2*x` wins), so the
        RNG stream is layout dependent; this restatement keeps operand order and layouts identical.
        noise_fn(x) -> tensor overrides the per-step draw (used to feed GPU-generated noise)."""
        noise_fn = noise_fn or torch.randn_like
        img = init_noise if init_noise is not None else torch.randn(*shape)
        x0 = None
        with torch.no_grad():
            for i in list(range(self.num_timesteps))[::-1]:
                t = torch.tensor([i] * shape[0])
                if ddim:
                    img, x0 = self.ddim_sample(model, img, t, noise_fn, eta, clip_denoised)
                else:
                    img, x0 = self.p_sample(model, img, t, noise_fn, clip_denoised)
        return img, x0


def rotation_6d_to_matrix(d6):
    # utils/rotation_conversions.py:529-534; F.normalize(v) = v / max(||v||_2, 1e-12)
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = a1 / a1.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = b2 / b2.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def auto_regressive(loop_fn, cmotion_bak, setting='cmdm'):
    """eval/a2m/stgcn_eval.py:50-67 restated.  loop_fn(f, cmotion) -> sample [B,V,C,T] is one full sampling loop with
    the actor motion revealed up to frame f (zeros afterwards); frame f of its result is kept."""
    B, V, C, T = cmotion_bak.shape
    cmotion = torch.zeros_like(cmotion_bak)
    output = torch.zeros((B, V, C * 2, T)) if setting == 'cmdm' else torch.zeros((B, V, C, T))
    for f in range(T):
        cmotion[:, :, :, f] = cmotion_bak[:, :, :, f]
        sample = loop_fn(f, cmotion)
        tmp = torch.cat((cmotion, sample), dim=2) if setting == 'cmdm' else sample
        output[:, :, :, f] = tmp[:, :, :, f]
    return output

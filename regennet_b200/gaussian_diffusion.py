"""Gaussian diffusion sampler -- host-side mirror of the reference's
``diffusion/gaussian_diffusion.py`` (sampling half), backed by sm_100a kernels.

Same names, argument meaning and error behaviour as the reference for everything the
sampling callers use (sample/cgenerate.py:121-135, eval/a2m/stgcn_eval.py:38-69):
``get_named_beta_schedule``, ``GaussianDiffusion(...)`` with its fp64 tables,
``p_mean_variance``, ``p_sample``, ``p_sample_loop(_progressive)``, ``ddim_sample``,
``ddim_sample_loop(_progressive)``, ``plms_sample`` / ``plms_sample_loop(_progressive)``.  Training losses and the
learned-variance / epsilon-prediction variants are out of scope (utils/model_util.py:75-117 hard-wires START_X with
a fixed variance) and raise NotImplementedError.

Two routes:
  * fast route -- ``model`` is this package's CMDM (or its ClassifierFreeSampleModel wrapper),
    no cond_fn / denoised_fn / inpainting: the whole step (denoiser + posterior update) runs as
    library kernels on a token-major [T,B,I] state; Python only draws the noise and enqueues.
  * generic route -- any callable model: the model is called as in the reference and the update
    runs as one fused kernel (regen_p_sample_update / regen_ddim_update).
Both draw noise exactly as the reference does (``th.randn(*shape)`` once, then ``th.randn_like(x)``
per step in x's memory layout), so a reference run and this run on the same device and seed see
the same noise.
"""
import enum
import math
from copy import deepcopy

import numpy as np
import torch as th

from . import _lib


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.):
    """diffusion/gaussian_diffusion.py:21-45."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    elif schedule_name == "cosine":
        return betas_for_alpha_bar(
            num_diffusion_timesteps,
            lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2,
        )
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """diffusion/gaussian_diffusion.py:48-65."""
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self == LossType.KL or self == LossType.RESCALED_KL


# ---------------------------------------------------------------------------------------------
# memory layouts of a logical [B,J,F,T] tensor
# ---------------------------------------------------------------------------------------------

def _layout_of(t):
    """'bjft' (contiguous), 'tbi' (memory order [T,B,J,F], what the reference's permuted model
    output carries) or None."""
    if t.is_contiguous():
        return "bjft"
    if t.dim() == 4 and t.permute(3, 0, 1, 2).is_contiguous():
        return "tbi"
    return None


def _empty_in_layout(shape, layout, device):
    if layout == "tbi":
        B, J, F, T = shape
        return th.empty((T, B, J, F), device=device, dtype=th.float32).permute(1, 2, 3, 0)
    return th.empty(tuple(shape), device=device, dtype=th.float32)


def _to_layout(t, layout):
    """Same logical tensor, memory re-ordered to `layout` with the library's transpose kernels."""
    cur = _layout_of(t)
    if cur is None:
        t = t.contiguous()
        cur = "bjft"
    if cur == layout:
        return t
    B, J, F, T = t.shape
    out = _empty_in_layout(t.shape, layout, t.device)
    fn = _lib.lib().regen_bjft_to_tbi if layout == "tbi" else _lib.lib().regen_tbi_to_bjft
    _lib.check(fn(_lib.ptr(t), _lib.ptr(out), B, J * F, T, _lib.stream_ptr(t.device)), "layout conversion")
    return out


class GaussianDiffusion:
    """diffusion/gaussian_diffusion.py:120-209 (constructor and fp64 tables)."""

    def __init__(
        self,
        *,
        betas,
        model_mean_type,
        model_var_type,
        loss_type,
        rescale_timesteps=False,
        lambda_rcxyz=0.,
        lambda_vel=0.,
        lambda_pose=1.,
        lambda_loc=1.,
        data_rep='rot6d',
        lambda_root_vel=0.,
        lambda_vel_rcxyz=0.,
        lambda_fc=0.,
        lambda_orient=0.,
        lambda_body=0.,
        lambda_transl=0.,
        num_person=1,
        body_model='smpl',
        vel_threshold=0.01,
    ):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        self.data_rep = data_rep
        if data_rep != 'rot_vel' and lambda_pose != 1.:
            raise ValueError('lambda_pose is relevant only when training on velocities!')
        # the lambda_* weights only parameterise training losses (out of scope); kept as attributes
        self.lambda_pose, self.lambda_orient, self.lambda_loc = lambda_pose, lambda_orient, lambda_loc
        self.lambda_rcxyz, self.lambda_vel, self.lambda_root_vel = lambda_rcxyz, lambda_vel, lambda_root_vel
        self.lambda_vel_rcxyz, self.lambda_fc, self.lambda_body = lambda_vel_rcxyz, lambda_fc, lambda_body
        self.lambda_transl = lambda_transl
        self.num_person, self.body_model, self.vel_threshold = num_person, body_model, vel_threshold

        # Use float64 for accuracy (gaussian_diffusion.py:172-209).
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])

        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        assert self.alphas_cumprod_prev.shape == (self.num_timesteps,)

        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)

        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

        self._dev_tables = {}

    # ------------------------------------------------------------------------------------ tables
    def _model_variance_tables(self):
        """(variance, log_variance) fp64 tables for the fixed-variance types (:338-364)."""
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            v = np.append(self.posterior_variance[1], self.betas[1:])
            return v, np.log(v)
        if self.model_var_type == ModelVarType.FIXED_SMALL:
            return self.posterior_variance, self.posterior_log_variance_clipped
        raise NotImplementedError("learned variance (%s) is not on the sampling hot path" % self.model_var_type)

    def _tables(self, device):
        """fp32 device copies of the fp64 tables: value == float32(fp64 entry), which is what
        _extract_into_tensor (:1604-1617) yields after its gather-then-.float()."""
        key = str(device)
        tabs = self._dev_tables.get(key)
        if tabs is None:
            var, logvar = self._model_variance_tables()
            src = dict(coef1=self.posterior_mean_coef1, coef2=self.posterior_mean_coef2, var=var, logvar=logvar,
                       sqrt_recip_ac=self.sqrt_recip_alphas_cumprod, sqrt_recipm1_ac=self.sqrt_recipm1_alphas_cumprod,
                       ac=self.alphas_cumprod, ac_prev=self.alphas_cumprod_prev,
                       sqrt_ac=self.sqrt_alphas_cumprod, sqrt_1m_ac=self.sqrt_one_minus_alphas_cumprod)
            tabs = {k: th.from_numpy(np.ascontiguousarray(v)).to(device=device).float() for k, v in src.items()}
            self._dev_tables[key] = tabs
        return tabs

    # ------------------------------------------------------------------------- q(x_t | x_0) etc.
    def q_sample(self, x_start, t, noise=None):
        """:239-256.  Only reached through init_image / skip_timesteps (editing), off the hot path."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    # ----------------------------------------------------------------------------- generic route
    def _check_supported(self):
        if self.model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError("only ModelMeanType.START_X is on the sampling hot path "
                                      "(utils/model_util.py:77 hard-wires predict_xstart=True)")
        self._model_variance_tables()

    def _call_model(self, model, x, t, denoised_fn, model_kwargs):
        """model call + inpainting blend + denoised_fn (:317-323, :366-370); clip happens in-kernel."""
        if model_kwargs is None:
            model_kwargs = {}
        B = x.shape[0]
        assert t.shape == (B,)
        model_output = model(x, self._scale_timesteps(t), **model_kwargs)
        y = model_kwargs.get('y', None)
        if isinstance(y, dict) and 'inpainting_mask' in y and 'inpainted_motion' in y:
            inpainting_mask, inpainted_motion = y['inpainting_mask'], y['inpainted_motion']
            assert model_output.shape == inpainting_mask.shape == inpainted_motion.shape
            model_output = (model_output * ~inpainting_mask) + (inpainted_motion * inpainting_mask)
        if denoised_fn is not None:
            model_output = denoised_fn(model_output)
        assert model_output.shape == x.shape
        return model_output

    def _update(self, kind, x, x0, noise, t, clip_denoised, eta=0.0, out=None):
        """One fused update kernel over (x, x0, noise) -> (sample, pred_xstart).  `out` (optional, fast route
        only) receives the sample instead of a fresh tensor; it must have x0's memory layout and may alias x.

        Memory layout follows the reference's TensorIterator rule: the result takes the layout of
        the model output (first operand of ``coef1*x0 + coef2*x``), x / noise are re-laid-out if
        they differ (only at the first step of a loop)."""
        _lib.require_cuda_f32(x, "x")
        _lib.require_cuda_f32(x0, "model output")
        if x.dim() == 4:
            lay = _layout_of(x0)
            if lay is None:
                x0 = x0.contiguous()
                lay = "bjft"
            x_l = _to_layout(x, lay)
            n_l = _to_layout(noise, lay) if noise is not None else None
            if out is None:
                out = _empty_in_layout(x.shape, lay, x.device)
            else:
                assert _layout_of(out) == lay and out.shape == x.shape
            pred = _empty_in_layout(x.shape, lay, x.device) if clip_denoised else None
            B = x.shape[0]
            inner = x[0].numel() if lay == "bjft" else x.shape[1] * x.shape[2]
        else:
            x0 = x0.contiguous()
            x_l = x.contiguous()
            n_l = noise.contiguous() if noise is not None else None
            out = th.empty_like(x0)
            pred = th.empty_like(x0) if clip_denoised else None
            B = x.shape[0]
            inner = x[0].numel() if x.dim() > 1 else 1
        t = t.to(device=x.device, dtype=th.int64).contiguous()
        tab = self._tables(x.device)
        L = _lib.lib()
        if kind == "p":
            rc = L.regen_p_sample_update(_lib.ptr(x_l), _lib.ptr(x0), _lib.ptr(n_l), _lib.ptr(out), _lib.ptr(pred),
                                         _lib.ptr(t), _lib.ptr(tab["coef1"]), _lib.ptr(tab["coef2"]),
                                         _lib.ptr(tab["logvar"]), x.numel(), inner, B, int(bool(clip_denoised)),
                                         _lib.stream_ptr(x.device))
            _lib.check(rc, "regen_p_sample_update")
        else:
            rc = L.regen_ddim_update(_lib.ptr(x_l), _lib.ptr(x0), _lib.ptr(n_l), _lib.ptr(out), _lib.ptr(pred),
                                     _lib.ptr(t), _lib.ptr(tab["sqrt_recip_ac"]), _lib.ptr(tab["sqrt_recipm1_ac"]),
                                     _lib.ptr(tab["ac"]), _lib.ptr(tab["ac_prev"]), float(eta), x.numel(), inner, B,
                                     int(bool(clip_denoised)), _lib.stream_ptr(x.device))
            _lib.check(rc, "regen_ddim_update")
        return out, (pred if clip_denoised else x0)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """:265-287."""
        assert x_start.shape == x_t.shape
        mean, _ = self._update("p", x_t, x_start, None, t, False)
        tab = self._tables(x_t.device)
        return (mean, _expand(tab["var"][t], x_t.shape), _expand(tab["logvar"][t], x_t.shape))

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """:289-400 for START_X + fixed variance: dict(mean, variance, log_variance, pred_xstart)."""
        self._check_supported()
        x0 = self._call_model(model, x, t, denoised_fn, model_kwargs)
        mean, pred = self._update("p", x, x0, None, t, clip_denoised)
        tab = self._tables(x.device)
        return {"mean": mean, "variance": _expand(tab["var"][t], x.shape),
                "log_variance": _expand(tab["logvar"][t], x.shape), "pred_xstart": pred}

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        """:418-423 (torch ops; only used by cond_fn score conditioning, off the hot path)."""
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart) \
            / _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        """:430-446 (classifier guidance; off the hot path, torch ops)."""
        gradient = cond_fn(x, self._scale_timesteps(t), **model_kwargs)
        return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 const_noise=False):
        """:508-560 -> {"sample", "pred_xstart"}."""
        self._check_supported()
        if cond_fn is not None:
            out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                       model_kwargs=model_kwargs)
            noise = th.randn_like(x)
            if const_noise:
                noise = noise[[0]].repeat(x.shape[0], 1, 1, 1)
            nonzero_mask = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
            out["mean"] = self.condition_mean(cond_fn, out, x, t, model_kwargs=model_kwargs)
            sample = out["mean"] + nonzero_mask * th.exp(0.5 * out["log_variance"]) * noise
            return {"sample": sample, "pred_xstart": out["pred_xstart"]}
        x0 = self._call_model(model, x, t, denoised_fn, model_kwargs)
        noise = th.randn_like(x)
        if const_noise:
            noise = noise[[0]].repeat(x.shape[0], 1, 1, 1)
        sample, pred = self._update("p", x, x0, noise, t, clip_denoised)
        return {"sample": sample, "pred_xstart": pred}

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                    eta=0.0):
        """:744-794 -> {"sample", "pred_xstart"}."""
        self._check_supported()
        if cond_fn is not None:
            raise NotImplementedError("cond_fn score conditioning (condition_score) is off the sampling hot path")
        x0 = self._call_model(model, x, t, denoised_fn, model_kwargs)
        noise = th.randn_like(x)
        sample, pred = self._update("ddim", x, x0, noise, t, clip_denoised, eta=eta)
        return {"sample": sample, "pred_xstart": pred}

    # ------------------------------------------------------------------------------------- loops
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """:610-673."""
        final = None
        if dump_steps is not None:
            dump = []
        if cond_fn_with_grad:
            raise NotImplementedError("p_sample_with_grad is off the sampling hot path")
        # only the final sample is consumed here, so the fast route may replay CUDA graphs of several steps at a
        # time (the progressive generator below hands out every intermediate tensor and stays step-by-step)
        for i, sample in enumerate(self._loop(
                "p", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
                skip_timesteps, init_image, randomize_class, const_noise, 0.0, graph_ok=dump_steps is None)):
            if dump_steps is not None and i in dump_steps:
                dump.append(deepcopy(sample["sample"]))
            final = sample
        if dump_steps is not None:
            return dump
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  cond_fn=None, model_kwargs=None, device=None, progress=False, skip_timesteps=0,
                                  init_image=None, randomize_class=False, cond_fn_with_grad=False,
                                  const_noise=False):
        """:675-742."""
        if cond_fn_with_grad:
            raise NotImplementedError("p_sample_with_grad is off the sampling hot path")
        yield from self._loop("p", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
                              progress, skip_timesteps, init_image, randomize_class, const_noise, 0.0)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """:891-939."""
        if dump_steps is not None:
            raise NotImplementedError()
        if const_noise == True:  # noqa: E712  (mirrors the reference check)
            raise NotImplementedError()
        final = None
        if cond_fn_with_grad:
            raise NotImplementedError("ddim_sample_with_grad is off the sampling hot path")
        for sample in self._loop("ddim", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
                                 progress, skip_timesteps, init_image, randomize_class, False, eta, graph_ok=True):
            final = sample
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                                     skip_timesteps=0, init_image=None, randomize_class=False,
                                     cond_fn_with_grad=False):
        """:941-1005."""
        if cond_fn_with_grad:
            raise NotImplementedError("ddim_sample_with_grad is off the sampling hot path")
        yield from self._loop("ddim", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
                              progress, skip_timesteps, init_image, randomize_class, False, eta)

    # ------------------------------------------------------------------------------------- PLMS
    def _plms_align(self, x, x0):
        """Operands of the PLMS kernels in the model output's memory layout (same rule as _update)."""
        _lib.require_cuda_f32(x, "x")
        _lib.require_cuda_f32(x0, "model output")
        if x.dim() == 4:
            lay = _layout_of(x0)
            if lay is None:
                x0, lay = x0.contiguous(), "bjft"
            new = lambda: _empty_in_layout(x.shape, lay, x.device)
            inner = x[0].numel() if lay == "bjft" else x.shape[1] * x.shape[2]
            return _to_layout(x, lay), x0, new, inner
        x0 = x0.contiguous()
        return x.contiguous(), x0, (lambda: th.empty_like(x0)), (x[0].numel() if x.dim() > 1 else 1)

    def plms_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                    cond_fn_with_grad=False, order=2, old_out=None):
        """Pseudo linear multistep step (:1007-1098) -> {"sample", "pred_xstart", "old_eps"}.  The model runs through
        _call_model (fused denoiser kernels for this package's CMDM); the eps / Adams-Bashforth / x_{t-1} arithmetic
        runs in regen_plms_eps / regen_plms_combine / regen_plms_finish."""
        if not int(order) or not 1 <= order <= 4:
            raise ValueError('order is invalid (should be int from 1-4).')
        if cond_fn is not None or cond_fn_with_grad:
            raise NotImplementedError("cond_fn score conditioning is off the sampling hot path")
        self._check_supported()
        L, tab, B = _lib.lib(), self._tables(x.device), x.shape[0]
        n_table = self.num_timesteps
        t = t.to(device=x.device, dtype=th.int64).contiguous()
        sp = _lib.stream_ptr(x.device)

        def model_eps(xin, tt, shift):
            """eps and (clipped) pred_xstart of the model at (xin, tt); tables at tt (= t + shift)."""
            x0 = self._call_model(model, xin, tt, denoised_fn, model_kwargs)
            xl, x0, new, inner = self._plms_align(xin, x0)
            eps, pred = new(), new()
            _lib.check(L.regen_plms_eps(_lib.ptr(xl), _lib.ptr(x0), _lib.ptr(eps), _lib.ptr(pred), _lib.ptr(t),
                                        _lib.ptr(tab["sqrt_recip_ac"]), _lib.ptr(tab["sqrt_recipm1_ac"]), xl.numel(), inner,
                                        B, n_table, shift, int(bool(clip_denoised)), sp), "regen_plms_eps")
            return xl, eps, pred, new, inner

        def finish(xl, epsp, pred, new, inner, mode):
            out = new()
            _lib.check(L.regen_plms_finish(_lib.ptr(xl), _lib.ptr(epsp), _lib.ptr(pred), _lib.ptr(out), _lib.ptr(t),
                                           _lib.ptr(tab["sqrt_recip_ac"]), _lib.ptr(tab["sqrt_recipm1_ac"]),
                                           _lib.ptr(tab["ac_prev"]), xl.numel(), inner, B, n_table, mode, sp),
                       "regen_plms_finish")
            return out

        def combine(hist, code, new):
            out = new()
            e = [hist[-1 - i] if i < len(hist) else None for i in range(4)]
            _lib.check(L.regen_plms_combine(_lib.ptr(e[0]), _lib.ptr(e[1]), _lib.ptr(e[2]), _lib.ptr(e[3]), _lib.ptr(out),
                                            out.numel(), code, sp), "regen_plms_combine")
            return out

        xl, eps, pred, new, inner = model_eps(x, t, 0)
        if order > 1 and old_out is None:
            # pseudo improved Euler: predictor with eps, second model evaluation at t - 1, average of the two eps
            old_eps = [eps]
            mean_pred = finish(xl, eps, pred, new, inner, 1)
            _, eps_2, _, _, _ = model_eps(mean_pred, t - 1, -1)
            eps_prime = combine([eps, eps_2], 5, new)
        else:
            old_eps = old_out["old_eps"]
            old_eps.append(eps)
            eps_prime = combine(old_eps, min(order, len(old_eps)), new)
        sample = finish(xl, eps_prime, pred, new, inner, 0)
        if len(old_eps) >= order:
            old_eps.pop(0)
        return {"sample": sample, "pred_xstart": pred, "old_eps": old_eps}

    def plms_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, order=2):
        """:1100-1130."""
        final = None
        for sample in self.plms_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                model_kwargs=model_kwargs, device=device, progress=progress, skip_timesteps=skip_timesteps,
                init_image=init_image, randomize_class=randomize_class, cond_fn_with_grad=cond_fn_with_grad, order=order):
            final = sample
        return final["sample"]

    def plms_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                                     randomize_class=False, cond_fn_with_grad=False, order=2):
        """:1132-1202 (deterministic: the only random draw is x_T)."""
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        indices = list(range(self.num_timesteps - skip_timesteps))[::-1]
        if init_image is not None:
            my_t = th.ones([shape[0]], device=device, dtype=th.long) * indices[0]
            img = self.q_sample(init_image, my_t, img)
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        old_out = None
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            if randomize_class and 'y' in model_kwargs:
                model_kwargs['y'] = th.randint(low=0, high=model.num_classes, size=model_kwargs['y'].shape,
                                               device=model_kwargs['y'].device)
            with th.no_grad():
                out = self.plms_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                       cond_fn=cond_fn, model_kwargs=model_kwargs, cond_fn_with_grad=cond_fn_with_grad,
                                       order=order, old_out=old_out)
                yield out
                old_out = out
                img = out["sample"]

    def _loop(self, kind, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
              skip_timesteps, init_image, randomize_class, const_noise, eta, graph_ok=False):
        self._check_supported()
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        if noise is not None:
            img = noise
        else:
            img = th.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        indices = list(range(self.num_timesteps - skip_timesteps))[::-1]
        if init_image is not None:
            my_t = th.ones([shape[0]], device=device, dtype=th.long) * indices[0]
            img = self.q_sample(init_image, my_t, img)

        fast = self._fast_session(model, shape, model_kwargs, denoised_fn, cond_fn, randomize_class, const_noise,
                                  img)
        if fast is not None:
            for out in fast.run(self, kind, img, indices, clip_denoised, eta, graph=graph_ok, progress=progress):
                out.pop("steps", None)   # the session's own bookkeeping; callers see the reference's two keys
                yield out
            return

        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)

        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            if randomize_class and 'y' in model_kwargs:
                model_kwargs['y'] = th.randint(low=0, high=model.num_classes, size=model_kwargs['y'].shape,
                                               device=model_kwargs['y'].device)
            with th.no_grad():
                if kind == "p":
                    out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                        cond_fn=cond_fn, model_kwargs=model_kwargs, const_noise=const_noise)
                else:
                    out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                           cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)
                yield out
                img = out["sample"]

    def _fast_session(self, model, shape, model_kwargs, denoised_fn, cond_fn, randomize_class, const_noise, img):
        """Return a fused sampling session when `model` is this package's denoiser and nothing in
        the call needs per-step Python hooks; otherwise None (generic route)."""
        maker = getattr(model, "regen_sampling_session", None)
        if maker is None or denoised_fn is not None or cond_fn is not None or randomize_class or const_noise:
            return None
        if self.rescale_timesteps or len(shape) != 4 or not img.is_cuda:
            return None
        y = (model_kwargs or {}).get('y', None)
        if not isinstance(y, dict):
            return None
        if ('inpainting_mask' in y) != ('inpainted_motion' in y):
            return None      # the reference blends only when both keys are present (:319); leave odd inputs to the generic route
        if 'inpainting_mask' in y:
            m, w = y['inpainting_mask'], y['inpainted_motion']
            if not (th.is_tensor(m) and th.is_tensor(w) and m.dtype == th.bool and tuple(m.shape) == tuple(shape)
                    and tuple(w.shape) == tuple(shape) and w.dtype == th.float32):
                return None
        return maker(shape, y, self._timestep_map_for_model())

    def _timestep_map_for_model(self):
        """Step index -> timestep handed to the model (identity here; SpacedDiffusion remaps)."""
        return list(range(self.num_timesteps))


def _expand(v, shape):
    while len(v.shape) < len(shape):
        v = v[..., None]
    return v.expand(shape)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """:1604-1617."""
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)

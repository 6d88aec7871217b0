"""Host-side logic of the product package (no GPU): schedules, respacing, API surface."""
import inspect
import os

import numpy as np
import pytest

import cases
from regennet_b200 import gaussian_diffusion as gd
from regennet_b200 import respace

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _spaced(rs):
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    return respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, rs if rs else [1000]), betas=betas,
                                   model_mean_type=gd.ModelMeanType.START_X,
                                   model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)


@pytest.mark.parametrize("rs", cases.RESPACINGS)
def test_tables_and_timestep_map_bit_exact_vs_reference(rs):
    g = np.load(os.path.join(HERE, "schedule.npz"))
    d = _spaced(rs)
    tag = "rs[%s]" % rs
    assert np.array_equal(np.array(d.timestep_map, dtype=np.int64), g[tag + ".timestep_map"])
    for f in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]:
        assert np.array_equal(getattr(d, f), g[tag + "." + f]), f


def test_beta_schedules_bit_exact():
    g = np.load(os.path.join(HERE, "schedule.npz"))
    assert np.array_equal(gd.get_named_beta_schedule("cosine", 1000), g["betas_cosine_1000"])
    assert np.array_equal(gd.get_named_beta_schedule("linear", 1000), g["betas_linear_1000"])
    with pytest.raises(NotImplementedError):
        gd.get_named_beta_schedule("sqrt", 10)


def test_space_timesteps_errors_and_edges():
    assert respace.space_timesteps(1000, "ddim5") == set(range(0, 1000, 200))
    assert respace.space_timesteps(10, [10]) == set(range(10))
    assert respace.space_timesteps(300, "10,15,20") == respace.space_timesteps(300, [10, 15, 20])
    with pytest.raises(ValueError):
        respace.space_timesteps(1000, "ddim999")   # no integer stride
    with pytest.raises(ValueError):
        respace.space_timesteps(10, [20])          # section too small


def test_loop_signatures_match_reference():
    want_p = ["self", "model", "shape", "noise", "clip_denoised", "denoised_fn", "cond_fn", "model_kwargs", "device",
              "progress", "skip_timesteps", "init_image", "randomize_class", "cond_fn_with_grad", "dump_steps",
              "const_noise"]
    assert list(inspect.signature(gd.GaussianDiffusion.p_sample_loop).parameters) == want_p
    want_d = ["self", "model", "shape", "noise", "clip_denoised", "denoised_fn", "cond_fn", "model_kwargs", "device",
              "progress", "eta", "skip_timesteps", "init_image", "randomize_class", "cond_fn_with_grad", "dump_steps",
              "const_noise"]
    assert list(inspect.signature(gd.GaussianDiffusion.ddim_sample_loop).parameters) == want_d
    sig = inspect.signature(gd.GaussianDiffusion.p_sample_loop).parameters
    assert sig["clip_denoised"].default is True and sig["noise"].default is None


def test_unsupported_variants_raise():
    betas = gd.get_named_beta_schedule("cosine", 10)
    d = gd.GaussianDiffusion(betas=betas, model_mean_type=gd.ModelMeanType.EPSILON,
                             model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    with pytest.raises(NotImplementedError):
        d._check_supported()
    with pytest.raises(ValueError):
        gd.GaussianDiffusion(betas=betas, model_mean_type=gd.ModelMeanType.START_X,
                             model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE, lambda_pose=2.0)

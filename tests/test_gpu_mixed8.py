"""precision='mixed8': on the large-batch (fused) route the two fused GEMM+LayerNorm kernels (attention out_proj + norm1 + norm2,
linear2 + norm3) run as one fp16 MMA plus two e4m3 correction MMAs per product (gemm_ln_sm100.cuh, M8 = true); their A
operands leave the producing epilogues in that operand format (FFN1: gemm_sm100.cuh Params::m8; attention kernels:
attention_sm100.cuh Params::m8).  Checked against the fp32 oracle, against the bf16x3 path on the same inputs, under
outlier channels, and over a 50-step sampling loop; the small-batch route must be untouched (bit-identical to bf16x3)."""
import pytest
import torch

import cases
from oracle import cmdm_ref
from regennet_b200 import synthetic
from regennet_b200.cfg_sampler import ClassifierFreeSampleModel
from regennet_b200.cmdm import CMDM
from test_gpu_denoiser import _kw, get_model, to_cuda
from test_gpu_parity_stress import _outlier_state_dict
from test_gpu_sampler import _diffusion

pytestmark = pytest.mark.gpu
TOL = 1e-3       # north-star tolerance
TOL_M8 = 1e-4    # what the mixed8 route is expected to hold on the forward (measured ~4e-5)

_m8 = {}


def _mixed8_model(sd, key, precision="mixed8"):
    """precision 'mixed8': the two fused GEMM+LayerNorm kernels; 'mixed8h': every GEMM of the fused route, residual stream
    kept as fp16 + e4m3 residual bytes (gemm_sm100.cuh / gemm_ln_sm100.cuh, M8 / H8)."""
    key = (key, precision)
    if key not in _m8:
        m = CMDM(precision=precision, **cases.MODELS["ntu"])
        m.load_state_dict(sd, strict=False)
        _m8[key] = m.cuda().eval()
    return _m8[key]


def test_bad_precision_rejected():
    with pytest.raises(ValueError):
        CMDM(precision="fp8", **cases.MODELS["ntu"])


# T = 60: compact attention kernel; T = 150: multi-chunk compact kernel; T = 196: 128-key-chunk kernel
@pytest.mark.parametrize("precision", ["mixed8", "mixed8h"])
@pytest.mark.parametrize("B,T", [(256, 60), (180, 60), (64, 150), (48, 196)])
def test_forward_matches_oracle_and_bf16x3(built_lib, B, T, precision):
    mk = cases.MODELS["ntu"]
    ref_model, sd = get_model("ntu", 0)
    model = _mixed8_model(sd, "w0", precision)
    x, y = synthetic.make_inputs(B, 56, 6, T, seed=400 + B)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(B + 1))
    with torch.no_grad():
        out = model(x.cuda(), t.cuda(), to_cuda(y)).cpu()
        out3 = ref_model(x.cuda(), t.cuda(), to_cuda(y)).cpu()
    sel = torch.tensor([0, 1, 63, 127, 128, B - 1]) if B > 128 else torch.tensor([0, 1, B // 2, B - 1])
    with torch.no_grad():
        want = cmdm_ref.cmdm_forward(sd, x[sel], t[sel], {"cmotion": y["cmotion"][sel]}, **_kw(mk))
    err = (out[sel] - want).abs().max().item()
    err3 = (out3[sel] - want).abs().max().item()
    d = (out - out3).abs().max().item()
    print("%s B=%d T=%d: max abs err vs oracle %.3e (bf16x3: %.3e); vs bf16x3 %.3e" % (precision, B, T, err, err3, d))
    assert torch.isfinite(out).all()
    assert err < (TOL_M8 if precision == "mixed8" else 2 * TOL_M8)
    assert d < 2 * TOL_M8
    assert d > 0.0   # the route really ran (a silent bf16x3 fallback would be bit-identical)


@pytest.mark.parametrize("precision", ["mixed8", "mixed8h"])
@pytest.mark.parametrize("name,model_name,B,T", [("config3", "chi3d", 128, 150), ("config5", "hml", 64, 196)])
def test_guided_configs_match_oracle_on_a_subset(built_lib, name, model_name, B, T, precision):
    """BASELINE configs 3 (Chi3D, action-conditioned) and 5 (HumanML-shaped text model, 263 input features: not a multiple of
    4, so 'mixed8h' keeps the bf16-pair residual stream there and runs like 'mixed8') under classifier-free guidance at their
    full per-GPU sizes -- the shapes bench.py's other_configs runs with the GPU arm's precision."""
    mk = cases.MODELS[model_name]
    _, sd = get_model(model_name, 0)
    key = (model_name, precision)
    if key not in _m8:
        m = CMDM(precision=precision, **mk)
        m.load_state_dict(sd, strict=False)
        _m8[key] = m.cuda().eval()
    run = ClassifierFreeSampleModel(_m8[key])
    x, y = synthetic.make_inputs(B, mk["njoints"], mk["nfeats"], T, seed=400 + B, cond_mode=mk["cond_mode"],
                                 num_actions=mk["num_actions"], scale=2.5)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(B + T))
    with torch.no_grad():
        out = run(x.cuda(), t.cuda(), to_cuda(y)).cpu()
    assert out.shape == x.shape and torch.isfinite(out).all()
    sel = torch.tensor([0, B // 2 - 1, B - 1])
    ysel = {k: (v[sel] if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in y.items()}
    with torch.no_grad():
        want = cmdm_ref.cfg_forward(sd, x[sel], t[sel], ysel, **_kw(mk))
    err = (out[sel] - want).abs().max().item()
    print("%s %s (B=%d, T=%d, CFG 2.5): max abs err vs oracle on 3 samples %.3e (absmax %.2f)" % (
        precision, name, B, T, err, want.abs().max()))
    assert err < TOL   # guidance amplifies the difference of the two forwards (2.5 x cond - 1.5 x uncond)


@pytest.mark.parametrize("precision", ["mixed8", "mixed8h"])
def test_small_batch_route_is_untouched(built_lib, precision):
    ref_model, sd = get_model("ntu", 0)
    model = _mixed8_model(sd, "w0", precision)
    x, y = synthetic.make_inputs(3, 56, 6, 60, seed=77)
    t = torch.tensor([1, 500, 999])
    with torch.no_grad():
        a = model(x.cuda(), t.cuda(), to_cuda(y))
        b = ref_model(x.cuda(), t.cuda(), to_cuda(y))
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["mixed8", "mixed8h"])
@pytest.mark.parametrize("l1_scale", [8.0, 100.0])
def test_outlier_channels_stay_within_tolerance(built_lib, l1_scale, precision):
    """x100 on four linear1 rows pushes FFN activations past 448 (e4m3's largest finite value): the operand scales of the
    fp8 copies keep |a| <= 1792 exact, anything larger saturates and falls back to fp16 accuracy for that element."""
    mk = cases.MODELS["ntu"]
    sd = _outlier_state_dict(5, l1_scale)
    model = _mixed8_model(sd, "outlier5_%g" % l1_scale, precision)
    B, T = 256, 60
    x, y = synthetic.make_inputs(B, 56, 6, T, seed=900 + B)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(B))
    with torch.no_grad():
        out = model(x.cuda(), t.cuda(), to_cuda(y)).cpu()
    sel = torch.tensor([0, 1, B // 2, B - 1])
    with torch.no_grad():
        want = cmdm_ref.cmdm_forward(sd, x[sel], t[sel], {"cmotion": y["cmotion"][sel]}, **_kw(mk))
    err = (out[sel] - want).abs().max().item()
    print("%s outlier stress (linear1 rows x%g): max abs err vs oracle %.3e (output absmax %.2f)" % (
        precision, l1_scale, err, want.abs().max()))
    assert torch.isfinite(out).all()
    assert err < TOL


@pytest.mark.parametrize("precision", ["mixed8", "mixed8h"])
def test_50_step_loop_stays_close_to_bf16x3(built_lib, precision):
    ref_model, sd = get_model("ntu", 0)
    model = _mixed8_model(sd, "w0", precision)
    B, T = 256, 60
    shape = (B, 56, 6, T)
    _, y = synthetic.make_inputs(B, 56, 6, T, seed=83)
    d = _diffusion("50")
    outs = []
    for m in (model, ref_model):
        torch.manual_seed(7)
        init = torch.randn(*shape, device="cuda")
        outs.append(d.p_sample_loop(m, shape, noise=init, clip_denoised=False, model_kwargs={"y": to_cuda(y)}).cpu())
    diff = (outs[0] - outs[1]).abs().max().item()
    print("%s vs bf16x3, 50-step loop at B=256: max abs diff %.3e (absmax %.2f)" % (precision, diff, outs[1].abs().max()))
    assert torch.isfinite(outs[0]).all()
    assert diff < 2 * TOL_M8


def test_offline_arch_stays_close_to_bf16x3(built_lib):
    """arch='offline' (encoder layers: out_proj + norm1 and linear2 + norm2 are both the non-chained fused kernel, S = T + 1
    tokens, unmasked attention) through the same mixed8 kernels."""
    name = "ntu_off"
    sd = synthetic.make_state_dict(seed=3, **cases.synth_kw_offline(name))
    models = []
    for prec in ("mixed8h", "bf16x3"):   # 'mixed8h' on arch 'offline' = 'mixed8' (the residual-stream format is online-only)
        m = CMDM(precision=prec, **cases.OFFLINE_MODELS[name])
        m.load_state_dict(sd, strict=False)
        models.append(m.cuda().eval())
    mk = cases.OFFLINE_MODELS[name]
    B, T = 200, 60
    x, y = synthetic.make_inputs(B, mk["njoints"], mk["nfeats"], T, seed=17, cond_mode=mk["cond_mode"],
                                 num_actions=mk.get("num_actions", 1))
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a, b = (m(x.cuda(), t.cuda(), to_cuda(y)).cpu() for m in models)
    d = (a - b).abs().max().item()
    print("offline mixed8 vs bf16x3 at B=%d: max abs diff %.3e (absmax %.2f)" % (B, d, b.abs().max()))
    assert torch.isfinite(a).all()
    assert 0.0 < d < 2 * TOL_M8


@pytest.mark.parametrize("precision", ["bf16x3", "mixed8", "mixed8h"])
def test_repeated_forwards_are_bit_identical(built_lib, precision):
    """Race / pipeline-hazard guard: 60 forwards of the same B = 256 input through the fused route (persistent tcgen05
    kernels, mbarrier rings, TMEM double buffers, no atomics anywhere) must return the same bits every time."""
    ref_model, sd = get_model("ntu", 0)
    model = ref_model if precision == "bf16x3" else _mixed8_model(sd, "w0", precision)
    B, T = 256, 60
    x, y = synthetic.make_inputs(B, 56, 6, T, seed=5)
    xc, yc = x.cuda(), to_cuda(y)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        first = model(xc, t, yc).clone()
        for _ in range(60):
            assert torch.equal(model(xc, t, yc), first)

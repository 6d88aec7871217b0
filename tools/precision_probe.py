"""CPU study (no GPU needed): forward error of CMDM under emulated tensor-core operand formats.

Every matmul of the oracle forward (Linear layers AND the two attention products) is replaced by a sum of products of
rounded operands with fp32 accumulation, which is what a tcgen05 kind::f16 pipeline with fp32 TMEM accumulators computes.
Answers the question "does anything cheaper than bf16x3 (3 MMAs per product) meet the 1e-3 tolerance with margin?".

    python tools/precision_probe.py            # B=2, T=60, NTU model, 3 weight / input seeds
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402

import cases  # noqa: E402
from oracle import cmdm_ref  # noqa: E402
from regennet_b200 import synthetic  # noqa: E402


def split(v, dt):
    hi = v.to(dt).float()
    lo = (v - hi).to(dt).float()
    return hi, lo


def q8_blocks(v, dim):
    """e4m3 rounding with a power-of-two scale per block of 32 elements along `dim` (what tcgen05 kind::mxf8f6f4 with
    UE8M0 block scales represents): value = e4m3(v / 2^e) * 2^e, e chosen so that the block maximum lands in [128, 256)."""
    v = v.transpose(dim, -1)
    shp = v.shape
    K = shp[-1]
    pad = (-K) % 32
    if pad:
        v = torch.nn.functional.pad(v, (0, pad))
    blk = v.reshape(*v.shape[:-1], -1, 32)
    amax = blk.abs().amax(-1, keepdim=True).clamp_min(1e-30)
    e = torch.floor(torch.log2(amax)) - 7.0
    scale = torch.exp2(e)
    q = (blk / scale).to(torch.float8_e4m3fn).float() * scale
    q = q.reshape(*v.shape[:-1], -1)[..., :K].reshape(shp)
    return q.transpose(dim, -1)


def make_mm(scheme):
    """-> f(a, b) computing a @ b under `scheme` (a: activations / left operand, b: weights / right operand)."""
    if scheme == "fp32":
        return lambda a, b: a @ b
    fmt, terms = scheme.split(":")
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]

    def mm(a, b):
        ah, al = split(a, dt)
        bh, bl = split(b, dt)
        if terms == "x1":
            return ah @ bh
        if terms == "a1_w2":        # activations single, weights split: 2 MMAs
            return ah @ bh + ah @ bl
        if terms == "a2_w1":        # activations split, weights single: 2 MMAs
            return ah @ bh + al @ bh
        if terms == "x3":
            return al @ bh + ah @ bl + ah @ bh
        if terms == "hi16_corr8g":  # corrections as PLAIN e4m3 products with global power-of-two scales whose sum is 15,
            # accumulated first and folded in by the first main-term MMA's scale-input-d (D = A.B + D * 2^-15):
            #   (A_lo * 2^11) . (W_hi * 2^4)  +  (A_hi * 2^0) . (W_lo * 2^15)
            e4 = lambda v: v.to(torch.float8_e4m3fn).float()
            corr = e4(al * 2.0 ** 11) @ e4(bh * 2.0 ** 4) + e4(ah) @ e4(bl * 2.0 ** 15)
            return ah @ bh + corr * 2.0 ** -15
        if terms == "hi16_corr8":   # main term in fp16, both correction terms as block-scaled e4m3 products (2x MMA rate)
            return ah @ bh + q8_blocks(al, -1) @ q8_blocks(bh, -2) + q8_blocks(ah, -1) @ q8_blocks(bl, -2)
        raise ValueError(terms)
    return mm


class patched:
    """Route every product of oracle.cmdm_ref through `lin_mm` (Linear layers) / `att_mm` (QK^T and PV)."""

    def __init__(self, lin_mm, att_mm, residual16=False):
        self.lin_mm, self.att_mm, self.residual16 = lin_mm, att_mm, residual16

    def __enter__(self):
        import math
        self.saved = (cmdm_ref._lin, cmdm_ref._mha, cmdm_ref._layer_norm)
        lin_mm, att_mm = self.lin_mm, self.att_mm
        cmdm_ref._lin = lambda a, w, b: lin_mm(a, w.t()) + b

        def mha(q_in, kv_in, w, b, wo, bo, nhead, mask):
            L, B, D = q_in.shape
            S = kv_in.shape[0]
            hd = D // nhead
            q = cmdm_ref._lin(q_in, w[:D], b[:D]).reshape(L, B * nhead, hd).transpose(0, 1)
            k = cmdm_ref._lin(kv_in, w[D:2 * D], b[D:2 * D]).reshape(S, B * nhead, hd).transpose(0, 1)
            v = cmdm_ref._lin(kv_in, w[2 * D:], b[2 * D:]).reshape(S, B * nhead, hd).transpose(0, 1)
            s = att_mm(q, k.transpose(1, 2)) * (1.0 / math.sqrt(hd))
            if mask is not None:
                s = s + mask
            p = torch.softmax(s, dim=-1)
            a = att_mm(p, v).transpose(0, 1).reshape(L, B, D)
            return cmdm_ref._lin(a, wo, bo)
        cmdm_ref._mha = mha
        if self.residual16:
            ln = self.saved[2]

            def ln16(a, w, b, eps=1e-5):   # the LayerNorm output (= next residual) keeps 16 significand bits
                hi, lo = split(ln(a, w, b, eps), torch.bfloat16)
                return hi + lo
            cmdm_ref._layer_norm = ln16
        return self

    def __exit__(self, *a):
        cmdm_ref._lin, cmdm_ref._mha, cmdm_ref._layer_norm = self.saved


SCHEMES = [
    ("fp32 (restatement)", "fp32", "fp32", False, 0),
    ("bf16 x1", "bf16:x1", "bf16:x1", False, 1),
    ("bf16 x3 (product path)", "bf16:x3", "bf16:x3", False, 3),
    ("bf16 x3 + (hi, lo) residual (fused route)", "bf16:x3", "bf16:x3", True, 3),
    ("fp16 x1", "fp16:x1", "fp16:x1", False, 1),
    ("fp16: A single, W split (2 MMAs); attention fp16 x3", "fp16:a1_w2", "fp16:x3", False, 2),
    ("fp16: A split, W single (2 MMAs); attention fp16 x3", "fp16:a2_w1", "fp16:x3", False, 2),
    ("fp16 x3", "fp16:x3", "fp16:x3", False, 3),
    ("fp16 hi.hi + two block-scaled e4m3 correction products", "fp16:hi16_corr8", "fp16:hi16_corr8", False, 2),
    ("bf16 hi.hi + two block-scaled e4m3 correction products", "bf16:hi16_corr8", "bf16:hi16_corr8", False, 2),
    ("fp16 hi.hi + plain e4m3 corrections, global scales 2^11/2^4, 2^0/2^15", "fp16:hi16_corr8g", "fp16:hi16_corr8g", False, 2),
    ("same for the Linear layers only; attention fp16 x3", "fp16:hi16_corr8g", "fp16:x3", False, 2),
]


def main():
    torch.set_num_threads(8)
    mk = cases.MODELS["ntu"]
    kw = dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])
    rows = []
    for label, lin, att, r16, mmas in SCHEMES:
        worst = 0.0
        for seed in range(3):
            sd = synthetic.make_state_dict(seed=seed, **cases.synth_kw("ntu"))
            x, y = synthetic.make_inputs(2, 56, 6, 60, seed=50 + seed)
            t = torch.tensor([900 - 300 * seed, 7 + seed])
            with torch.no_grad():
                ref = cmdm_ref.cmdm_forward(sd, x.double(), t, {"cmotion": y["cmotion"].double()},
                                            **kw) if False else cmdm_ref.cmdm_forward(sd, x, t, y, **kw)
                with patched(make_mm(lin), make_mm(att), r16):
                    out = cmdm_ref.cmdm_forward(sd, x, t, y, **kw)
            worst = max(worst, (out - ref).abs().max().item())
        rows.append((label, mmas, worst))
        print("%-58s MMAs/product %d   max abs err vs fp32 oracle %.3e" % (label, mmas, worst), flush=True)
    return rows


if __name__ == "__main__" and "--loop" not in sys.argv:
    main()


def loop_errors(respacing="ddim50"):
    """Final-sample error of a whole respaced p_sample_loop (same noise stream) under a scheme vs the fp32 oracle."""
    from oracle import sampler_ref
    mk = cases.MODELS["ntu"]
    kw = dict(num_layers=mk["num_layers"], nhead=mk["num_heads"], cond_mode=mk["cond_mode"], cm_mode=mk["cm_mode"])
    sd = synthetic.make_state_dict(seed=0, **cases.synth_kw("ntu"))
    _, y = synthetic.make_inputs(2, 56, 6, 60, seed=50)
    shape = (2, 56, 6, 60)

    def run():
        smp = sampler_ref.Sampler(timestep_respacing=respacing)
        torch.manual_seed(10)
        out, _ = smp.loop(lambda xx, tt: cmdm_ref.cmdm_forward(sd, xx, tt, y, **kw), shape)
        return out

    ref = run()
    for label, lin, att, r16, mmas in SCHEMES:
        if lin in ("fp32", "bf16:x1", "fp16:x1", "fp16:a1_w2", "fp16:a2_w1"):
            continue
        with patched(make_mm(lin), make_mm(att), r16):
            out = run()
        print("%-58s %s loop: final-sample max abs err %.3e" % (label, respacing, (out - ref).abs().max().item()), flush=True)


if __name__ == "__main__" and "--loop" in sys.argv:
    loop_errors()

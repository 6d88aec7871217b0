"""Oracle (test infrastructure): CMDM.forward, arch='online' and arch='offline', restated on CPU fp32.

Follows model/cmdm.py:173-252 (forward), :265-281 (PositionalEncoding),
:284-298 (TimestepEmbedder), :301-317 (InputProcess), :329-355 (OutputProcess),
:358-366 (EmbedAction), :129-137 (mask_cond), :168-171 (causal mask), and
model/cfg_sampler.py:24-31 (classifier-free guidance).

The decoder layer arithmetic is torch.nn.TransformerDecoderLayer (post-norm,
norm_first=False, eps=1e-5, exact-erf GELU, dropout = identity in eval) as the
reference constructs it at model/cmdm.py:75-81 and calls it at :227; torch is a
third-party dependency of the reference (pinned 1.7.1 / 1.12.0, not vendored), so
its published algorithm is restated here with plain matmul / softmax / mean / var.

Works on a plain ``state_dict`` (name -> tensor) with the reference's key names.
"""
import math

import numpy as np
import torch


def positional_table(max_len, d):
    # model/cmdm.py:266-276
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-np.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1)  # [max_len, 1, d]


def _lin(a, w, b):
    return a @ w.t() + b


def _layer_norm(a, w, b, eps=1e-5):
    mu = a.mean(-1, keepdim=True)
    var = ((a - mu) ** 2).mean(-1, keepdim=True)
    return (a - mu) / torch.sqrt(var + eps) * w + b


def _gelu(a):
    return 0.5 * a * (1.0 + torch.erf(a / math.sqrt(2.0)))


def _mha(q_in, kv_in, w, b, wo, bo, nhead, mask):
    """nn.MultiheadAttention, seq-first [L,B,D] query, [S,B,D] key/value."""
    L, B, D = q_in.shape
    S = kv_in.shape[0]
    hd = D // nhead
    q = _lin(q_in, w[:D], b[:D])
    k = _lin(kv_in, w[D:2 * D], b[D:2 * D])
    v = _lin(kv_in, w[2 * D:], b[2 * D:])
    q = q.reshape(L, B * nhead, hd).transpose(0, 1)
    k = k.reshape(S, B * nhead, hd).transpose(0, 1)
    v = v.reshape(S, B * nhead, hd).transpose(0, 1)
    s = (q * (1.0 / math.sqrt(hd))) @ k.transpose(1, 2)  # [B*H, L, S]
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    a = (p @ v).transpose(0, 1).reshape(L, B, D)
    return _lin(a, wo, bo)


def causal_mask(sz):
    # model/cmdm.py:168-171: 0 where j <= i, -inf elsewhere
    m = torch.full((sz, sz), float("-inf"))
    return torch.triu(m, diagonal=1)


def embed(sd, timesteps, y, cond_mode):
    """emb [1,B,D]: model/cmdm.py:179-187."""
    pe = sd["sequence_pos_encoder.pe"]
    e = _lin(pe[timesteps], sd["embed_timestep.time_embed.0.weight"], sd["embed_timestep.time_embed.0.bias"])
    e = e * torch.sigmoid(e)  # SiLU
    e = _lin(e, sd["embed_timestep.time_embed.2.weight"], sd["embed_timestep.time_embed.2.bias"])
    e = e.permute(1, 0, 2)  # [1,B,D]
    force_mask = y.get("uncond", False)
    if "text" in cond_mode:
        # encode_text (CLIP) is outside the path; features are injected as y['text_embed'] [B,clip_dim]
        enc = y["text_embed"]
        enc = torch.zeros_like(enc) if force_mask else enc
        e = e + _lin(enc, sd["embed_text.weight"], sd["embed_text.bias"])
    if "action" in cond_mode:
        idx = y["action"][:, 0].to(torch.long)
        a = sd["embed_action.action_embedding"][idx]
        a = torch.zeros_like(a) if force_mask else a
        e = e + a
    return e


def cmdm_forward(sd, x, timesteps, y, *, num_layers=8, nhead=4, cond_mode="no_cond", cm_mode="concat", arch="online"):
    """x [B,J,F,T], timesteps int64 [B] -> x0_hat [B,J,F,T].  model/cmdm.py:173-252."""
    B, J, F, T = x.shape
    emb = embed(sd, timesteps, y, cond_mode)
    X = x.permute(3, 0, 1, 2).reshape(T, B, J * F)
    C = y["cmotion"].permute(3, 0, 1, 2).reshape(T, B, J * F)
    hx = _lin(X, sd["input_process.poseEmbedding.weight"], sd["input_process.poseEmbedding.bias"])
    hc = _lin(C, sd["cmo_process.poseEmbedding.weight"], sd["cmo_process.poseEmbedding.bias"])
    if cm_mode == "add":
        h = hx + hc
    else:
        h = _lin(torch.cat((hx, hc), -1), sd["fuse_process.weight"], sd["fuse_process.bias"])
    if arch == "offline":
        # model/cmdm.py:228-238: the condition embedding is token 0, positions 0..T, nn.TransformerEncoder
        # (post-norm encoder layers, no mask: every frame attends to every frame), token 0 dropped at the end
        h = torch.cat((emb, h), 0) + sd["sequence_pos_encoder.pe"][:T + 1]
        for l in range(num_layers):
            p = "seqTransEncoder.layers.%d." % l
            sa = _mha(h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                      sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, None)
            h = _layer_norm(h + sa, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            ff = _lin(_gelu(_lin(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                      sd[p + "linear2.weight"], sd[p + "linear2.bias"])
            h = _layer_norm(h + ff, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        out = _lin(h[1:], sd["output_process.poseFinal.weight"], sd["output_process.poseFinal.bias"])
        return out.reshape(T, B, J, F).permute(1, 2, 3, 0)
    h = h + sd["sequence_pos_encoder.pe"][:T]
    mask = causal_mask(T)
    for l in range(num_layers):
        p = "seqTransDecoder.layers.%d." % l
        sa = _mha(h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                  sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, mask)
        h = _layer_norm(h + sa, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        ca = _mha(h, emb, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                  sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], nhead, None)
        h = _layer_norm(h + ca, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        ff = _lin(_gelu(_lin(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                  sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        h = _layer_norm(h + ff, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    out = _lin(h, sd["output_process.poseFinal.weight"], sd["output_process.poseFinal.bias"])
    # NOT made contiguous: the reference returns the permuted view (model/cmdm.py:353-354) and the
    # sampler's randn_like() inherits that memory layout, which changes the CPU/GPU noise stream.
    return out.reshape(T, B, J, F).permute(1, 2, 3, 0)


def cfg_forward(sd, x, timesteps, y, **kw):
    """model/cfg_sampler.py:24-31: u + s*(c-u)."""
    y_un = dict(y)
    y_un["uncond"] = True
    c = cmdm_forward(sd, x, timesteps, y, **kw)
    u = cmdm_forward(sd, x, timesteps, y_un, **kw)
    return u + y["scale"].view(-1, 1, 1, 1) * (c - u)

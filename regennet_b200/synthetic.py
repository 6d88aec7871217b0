"""Deterministic synthetic CMDM weights and inputs.

The reference publishes checkpoints only as Google-Drive links (README.md:67) and
the datasets are not distributable (README.md:76), so tests, golden vectors and the
benchmark all use weights drawn from a per-key seeded CPU generator.  The factory
depends only on the state-dict key names/shapes, never on module
construction order, so the same tensors can be loaded into the reference CMDM,
the oracle and this package's CMDM.
"""
import math

import numpy as np
import torch


def positional_table(max_len, d):
    """Sinusoid buffer ``sequence_pos_encoder.pe`` [max_len,1,d] (model/cmdm.py:266-276)."""
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-np.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def state_dict_spec(njoints=56, nfeats=6, latent_dim=512, ff_size=1024, num_layers=8,
                    cond_mode="no_cond", num_actions=1, clip_dim=512, cm_mode="concat", arch="online"):
    """Ordered list of (key, shape, kind); kind in {'w','b','ln_w','ln_b','emb'}.  arch 'online' gives the
    nn.TransformerDecoder keys (model/cmdm.py:75-81), 'offline' the nn.TransformerEncoder keys (:63-71)."""
    D, I, Fd = latent_dim, njoints * nfeats, ff_size
    spec = [
        ("input_process.poseEmbedding.weight", (D, I), "w"),
        ("input_process.poseEmbedding.bias", (D,), "b"),
        ("cmo_process.poseEmbedding.weight", (D, I), "w"),
        ("cmo_process.poseEmbedding.bias", (D,), "b"),
    ]
    if cm_mode == "concat":
        spec += [("fuse_process.weight", (D, 2 * D), "w"), ("fuse_process.bias", (D,), "b")]
    for l in range(num_layers if arch == "offline" else 0):
        p = "seqTransEncoder.layers.%d." % l
        spec += [
            (p + "self_attn.in_proj_weight", (3 * D, D), "w"),
            (p + "self_attn.in_proj_bias", (3 * D,), "b"),
            (p + "self_attn.out_proj.weight", (D, D), "w"),
            (p + "self_attn.out_proj.bias", (D,), "b"),
            (p + "linear1.weight", (Fd, D), "w"),
            (p + "linear1.bias", (Fd,), "b"),
            (p + "linear2.weight", (D, Fd), "w"),
            (p + "linear2.bias", (D,), "b"),
            (p + "norm1.weight", (D,), "ln_w"), (p + "norm1.bias", (D,), "ln_b"),
            (p + "norm2.weight", (D,), "ln_w"), (p + "norm2.bias", (D,), "ln_b"),
        ]
    for l in range(num_layers if arch != "offline" else 0):
        p = "seqTransDecoder.layers.%d." % l
        spec += [
            (p + "self_attn.in_proj_weight", (3 * D, D), "w"),
            (p + "self_attn.in_proj_bias", (3 * D,), "b"),
            (p + "self_attn.out_proj.weight", (D, D), "w"),
            (p + "self_attn.out_proj.bias", (D,), "b"),
            (p + "multihead_attn.in_proj_weight", (3 * D, D), "w"),
            (p + "multihead_attn.in_proj_bias", (3 * D,), "b"),
            (p + "multihead_attn.out_proj.weight", (D, D), "w"),
            (p + "multihead_attn.out_proj.bias", (D,), "b"),
            (p + "linear1.weight", (Fd, D), "w"),
            (p + "linear1.bias", (Fd,), "b"),
            (p + "linear2.weight", (D, Fd), "w"),
            (p + "linear2.bias", (D,), "b"),
            (p + "norm1.weight", (D,), "ln_w"), (p + "norm1.bias", (D,), "ln_b"),
            (p + "norm2.weight", (D,), "ln_w"), (p + "norm2.bias", (D,), "ln_b"),
            (p + "norm3.weight", (D,), "ln_w"), (p + "norm3.bias", (D,), "ln_b"),
        ]
    spec += [
        ("embed_timestep.time_embed.0.weight", (D, D), "w"),
        ("embed_timestep.time_embed.0.bias", (D,), "b"),
        ("embed_timestep.time_embed.2.weight", (D, D), "w"),
        ("embed_timestep.time_embed.2.bias", (D,), "b"),
    ]
    if "text" in cond_mode:
        spec += [("embed_text.weight", (D, clip_dim), "w"), ("embed_text.bias", (D,), "b")]
    if "action" in cond_mode:
        spec += [("embed_action.action_embedding", (num_actions, D), "emb")]
    spec += [
        ("output_process.poseFinal.weight", (I, D), "w"),
        ("output_process.poseFinal.bias", (I,), "b"),
    ]
    return spec


def make_state_dict(seed=0, max_len=5000, **model_kw):
    """Seeded weights: Linear ~ U(+-1/sqrt(fan_in)) like nn.Linear's default, LayerNorm
    affine perturbed away from (1,0) so the affine path is exercised, action table ~ N(0,1)."""
    spec = state_dict_spec(**model_kw)
    D = model_kw.get("latent_dim", 512)
    sd = {}
    for i, (key, shape, kind) in enumerate(spec):
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        if kind == "w":
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g)
        sd[key] = t
    # the reference shares one PositionalEncoding module between the model and the timestep
    # embedder (model/cmdm.py:50,92), so its buffer appears under two keys
    sd["sequence_pos_encoder.pe"] = positional_table(max_len, D)
    sd["embed_timestep.sequence_pos_encoder.pe"] = sd["sequence_pos_encoder.pe"]
    return sd


def make_inputs(B, njoints, nfeats, T, seed=10, cond_mode="no_cond", num_actions=1, clip_dim=512,
                scale=None):
    """Seeded x, cmotion and conditioning dict (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, njoints, nfeats, T, generator=g)
    y = {"cmotion": torch.randn(B, njoints, nfeats, T, generator=g)}
    if "action" in cond_mode:
        y["action"] = torch.randint(0, num_actions, (B, 1), generator=g)
    if "text" in cond_mode:
        y["text_embed"] = torch.randn(B, clip_dim, generator=g)
    if scale is not None:
        y["scale"] = torch.ones(B) * scale
    return x, y

"""Conditional motion diffusion model -- host-side mirror of the reference's ``model/cmdm.py``
for ``arch='online'`` (causal decoder) and ``arch='offline'`` (encoder with the condition as first token), cm_mode
'concat' or 'add', backed by libregen_sm100.

The class keeps the reference constructor signature, attribute names and state-dict keys
(model/cmdm.py:13-111), so ``utils/model_util.create_model_and_diffusion`` + ``load_model_wo_clip``
and the reference's checkpoints work unchanged.  The torch modules created here are parameter
containers only: ``forward`` never calls them -- all arithmetic runs in the library's sm_100a
kernels (tcgen05 GEMMs, fused epilogues, attention, LayerNorm).  There is no CPU or PyTorch
fallback: calling ``forward`` on a CPU tensor raises.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class PositionalEncoding(nn.Module):
    """model/cmdm.py:265-281 (buffer ``pe`` [max_len, 1, d_model])."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        pe = pe.unsqueeze(0).transpose(0, 1)
        self.register_buffer('pe', pe)


class TimestepEmbedder(nn.Module):
    """model/cmdm.py:284-298 (parameter container; evaluated inside the library as a table)."""

    def __init__(self, latent_dim, sequence_pos_encoder):
        super().__init__()
        self.latent_dim = latent_dim
        self.sequence_pos_encoder = sequence_pos_encoder
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.SiLU(), nn.Linear(latent_dim, latent_dim))


class InputProcess(nn.Module):
    """model/cmdm.py:301-317 (parameter container)."""

    def __init__(self, data_rep, input_feats, latent_dim):
        super().__init__()
        self.data_rep, self.input_feats, self.latent_dim = data_rep, input_feats, latent_dim
        self.poseEmbedding = nn.Linear(input_feats, latent_dim)


class OutputProcess(nn.Module):
    """model/cmdm.py:329-355 (parameter container)."""

    def __init__(self, data_rep, input_feats, latent_dim, njoints, nfeats):
        super().__init__()
        self.data_rep, self.input_feats, self.latent_dim = data_rep, input_feats, latent_dim
        self.njoints, self.nfeats = njoints, nfeats
        self.poseFinal = nn.Linear(latent_dim, input_feats)


class EmbedAction(nn.Module):
    """model/cmdm.py:358-366 (parameter container)."""

    def __init__(self, num_actions, latent_dim):
        super().__init__()
        self.action_embedding = nn.Parameter(torch.randn(num_actions, latent_dim))


def _cfg_scale(scale, B, device):
    """y['scale'] as a contiguous CUDA fp32 [B] vector.  The reference broadcasts ``y['scale'].view(-1, 1, 1, 1)``
    (model/cfg_sampler.py:31), so one element (shared by the batch) or B elements are legal; anything else would make the
    kernel read past the buffer and is rejected here."""
    if not torch.is_tensor(scale):
        scale = torch.as_tensor(scale, dtype=torch.float32)
    scale = _lib.require_cuda_f32(scale.to(device), "y['scale']").reshape(-1)
    if scale.numel() == 1 and B != 1:
        scale = scale.expand(B)
    if scale.numel() != B:
        raise ValueError("y['scale'] must hold 1 or B=%d guidance scales, got %d" % (B, scale.numel()))
    return scale.contiguous()


class _Handle:
    """Owns one regen_handle (device buffers sized for max_batch x max_frames)."""

    def __init__(self, ptr, max_batch, max_frames, device, key):
        self.ptr, self.max_batch, self.max_frames, self.device, self.key = ptr, max_batch, max_frames, device, key

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().regen_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


_PRECISIONS = {'bf16x3': 0, 'bf16': 1, 'mixed8': 2, 'mixed8h': 3}   # regen_model_desc.precision


class CMDM(nn.Module):
    def __init__(self, modeltype, njoints, nfeats, num_actions, translation, pose_rep, glob, glob_rot,
                 num_frames=60, latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1,
                 ablation=None, activation="gelu", legacy=False, data_rep='rot6d', dataset='amass', clip_dim=512,
                 arch='trans_enc', cm_mode='add', body_model='smpl', wo_pos_emb=False, emb_trans_dec=False,
                 clip_version=None, **kargs):
        super().__init__()
        self.legacy = legacy
        self.modeltype = modeltype
        self.njoints = njoints
        self.nfeats = nfeats
        self.num_actions = num_actions
        self.data_rep = data_rep
        self.dataset = dataset
        self.pose_rep = pose_rep
        self.glob = glob
        self.glob_rot = glob_rot
        self.translation = translation
        self.latent_dim = latent_dim
        self.ff_size = ff_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.ablation = ablation
        self.activation = activation
        self.clip_dim = clip_dim
        self.action_emb = kargs.get('action_emb', None)
        self.input_feats = self.njoints * self.nfeats
        self.normalize_output = kargs.get('normalize_encoder_output', False)
        self.cond_mode = kargs.get('cond_mode', 'no_cond')
        self.cond_mask_prob = kargs.get('cond_mask_prob', 0.)
        self.arch = arch
        self.cm_mode = cm_mode
        self.num_frames = num_frames
        self.emb_trans_dec = emb_trans_dec
        self.wo_pos_emb = wo_pos_emb
        self.body_model = body_model
        #: 'bf16x3' (parity mode: three bf16 MMAs per product, ~2e-5 abs error), 'mixed8' (bf16x3 except linear2 of the
        #: large-batch route: one fp16 MMA + two e4m3 correction MMAs per product, ~4e-5), 'mixed8h' (arch 'online': every
        #: GEMM of the large-batch route that way, the residual stream kept as fp16 + e4m3 residual bytes) or 'bf16' (single pass, ~1e-2)
        self.precision = kargs.get('precision', 'bf16x3')
        if self.precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s (got %r)" % (sorted(_PRECISIONS), self.precision))

        # --- scope of the B200 path: everything else in the reference is a different model
        if arch not in ('online', 'offline'):
            raise NotImplementedError("regennet_b200.CMDM implements arch='online' and arch='offline' (got %r); the "
                                      "trans_enc / trans_dec / gru / mlp variants of model/cmdm.py:63-89 are outside "
                                      "the sampling hot path" % arch)
        if cm_mode not in ('concat', 'add'):
            raise ValueError("cm_mode must be 'concat' or 'add'")
        if emb_trans_dec or wo_pos_emb:
            raise NotImplementedError("emb_trans_dec / wo_pos_emb variants are not on the hot path "
                                      "(defaults False, utils/parser_util.py:116)")
        if latent_dim != 512 or num_heads != 4 or ff_size != 1024 or activation != 'gelu':
            raise NotImplementedError("kernels are specialised for latent_dim=512, num_heads=4, ff_size=1024, gelu "
                                      "(utils/model_util.py:69-70 and the --latent_dim default)")
        if data_rep not in ('rot6d', 'xyz', 'hml_vec'):
            raise ValueError(data_rep)

        self.input_process = InputProcess(self.data_rep, self.input_feats, self.latent_dim)
        self.cmo_process = InputProcess(self.data_rep, self.input_feats, self.latent_dim)
        self.sequence_pos_encoder = PositionalEncoding(self.latent_dim, self.dropout)
        if self.cm_mode == 'concat':
            self.fuse_process = nn.Linear(self.latent_dim * 2, self.latent_dim)
        if arch == 'offline':   # model/cmdm.py:63-71: encoder over [condition token | frames], no mask
            layer = nn.TransformerEncoderLayer(d_model=self.latent_dim, nhead=self.num_heads,
                                               dim_feedforward=self.ff_size, dropout=self.dropout,
                                               activation=activation)
            self.seqTransEncoder = nn.TransformerEncoder(layer, num_layers=self.num_layers, enable_nested_tensor=False)
        else:                   # model/cmdm.py:75-81: causal decoder, the condition is a 1-token memory
            layer = nn.TransformerDecoderLayer(d_model=self.latent_dim, nhead=self.num_heads,
                                               dim_feedforward=self.ff_size, dropout=self.dropout,
                                               activation=activation)
            self.seqTransDecoder = nn.TransformerDecoder(layer, num_layers=self.num_layers)
        self.embed_timestep = TimestepEmbedder(self.latent_dim, self.sequence_pos_encoder)
        if self.cond_mode != 'no_cond':
            if 'text' in self.cond_mode:
                self.embed_text = nn.Linear(self.clip_dim, self.latent_dim)
                self.clip_version = clip_version
                self.clip_model = self.load_and_freeze_clip(clip_version)
            if 'action' in self.cond_mode:
                self.embed_action = EmbedAction(self.num_actions, self.latent_dim)
        self.output_process = OutputProcess(self.data_rep, self.input_feats, self.latent_dim, self.njoints,
                                            self.nfeats)
        from .rotation2xyz import Rotation2xyz
        self.rot2xyz = Rotation2xyz(device='cpu', dataset=self.dataset, body_model=body_model)

        self._handle = None
        self._cond_key = None
        self._weights_version = None

    # ----------------------------------------------------------------------------- reference API
    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith('clip_model.')]

    def load_and_freeze_clip(self, clip_version):
        """model/cmdm.py:116-127.  CLIP is a third-party text encoder outside the hot path; when the
        package is absent the model still samples from precomputed features (y['text_embed'])."""
        try:
            import clip  # noqa: F401
        except Exception:
            return None
        clip_model, _ = clip.load(clip_version, device='cpu', jit=False)
        clip.model.convert_weights(clip_model)
        clip_model.eval()
        for p in clip_model.parameters():
            p.requires_grad = False
        return clip_model

    def encode_text(self, raw_text):
        """model/cmdm.py:153-166."""
        if getattr(self, 'clip_model', None) is None:
            raise RuntimeError("CLIP is not available; pass precomputed text features as y['text_embed'] [B,%d]"
                               % self.clip_dim)
        import clip
        device = next(self.parameters()).device
        max_text_len = 20 if self.dataset in ['humanml', 'kit'] else None
        if max_text_len is not None:
            default_context_length = 77
            context_length = max_text_len + 2
            texts = clip.tokenize(raw_text, context_length=context_length, truncate=True).to(device)
            zero_pad = torch.zeros([texts.shape[0], default_context_length - context_length], dtype=texts.dtype,
                                   device=texts.device)
            texts = torch.cat([texts, zero_pad], dim=1)
        else:
            texts = clip.tokenize(raw_text, truncate=True).to(device)
        return self.clip_model.encode_text(texts).float()

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        body = getattr(getattr(self, 'rot2xyz', None), 'smpl_model', None)
        if body is not None:        # model/cmdm.py:255-257: the body model follows .to() / .cuda()
            body._apply(fn)
        self._invalidate()
        return out

    def train(self, *args, **kwargs):
        out = super().train(*args, **kwargs)
        body = getattr(getattr(self, 'rot2xyz', None), 'smpl_model', None)
        if body is not None:        # model/cmdm.py:260-262
            body.train(*args, **kwargs)
        return out

    def _invalidate(self):
        self._handle = None
        self._cond_key = None
        self.__dict__.pop("_graph_cache", None)   # captured graphs hold the old handle's buffers

    # --------------------------------------------------------------------------- library plumbing
    def _get_handle(self, batch_eff, frames, device):
        """Create (or grow) the library handle and pack the current weights into it."""
        wkey = tuple((p.data_ptr(), p._version) for p in self.parameters_wo_clip())
        h = self._handle
        if (h is not None and h.device == device and h.max_batch >= batch_eff and h.max_frames >= frames
                and h.key == (wkey, self.precision)):
            return h
        self._invalidate()
        if device.type != 'cuda':
            raise RuntimeError("regennet_b200.CMDM runs on CUDA (sm_100a) only; move the model and inputs to a GPU "
                               "-- there is no CPU fallback for the sampling hot path")
        L = _lib.lib()
        table_steps = self._table_steps()
        desc = _lib.ModelDesc(latent_dim=self.latent_dim, num_heads=self.num_heads, ff_size=self.ff_size,
                              num_layers=self.num_layers, input_feats=self.input_feats,
                              cm_mode=1 if self.cm_mode == 'concat' else 0,
                              max_batch=max(batch_eff, h.max_batch if h else 0),
                              max_frames=max(frames, h.max_frames if h else 0, min(self.num_frames, 196)),
                              num_table_steps=table_steps, precision=_PRECISIONS[self.precision],
                              arch=1 if self.arch == 'offline' else 0)
        hp = ctypes.c_void_p()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(L.regen_create(ctypes.byref(hp), dev_index, ctypes.byref(desc)), "regen_create")
        handle = _Handle(hp, desc.max_batch, desc.max_frames, device, (wkey, self.precision))

        def P(t):
            t = t.detach()
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("CMDM parameters must be contiguous CUDA float32 tensors")
            return t.data_ptr()

        w = _lib.WeightPtrs()
        D = self.latent_dim
        w.in_w, w.in_b = P(self.input_process.poseEmbedding.weight), P(self.input_process.poseEmbedding.bias)
        w.cmo_w, w.cmo_b = P(self.cmo_process.poseEmbedding.weight), P(self.cmo_process.poseEmbedding.bias)
        if self.cm_mode == 'concat':
            w.fuse_w, w.fuse_b = P(self.fuse_process.weight), P(self.fuse_process.bias)
        te = self.embed_timestep.time_embed
        w.t0_w, w.t0_b, w.t2_w, w.t2_b = P(te[0].weight), P(te[0].bias), P(te[2].weight), P(te[2].bias)
        pe = self.sequence_pos_encoder.pe
        w.pe, w.pe_len = P(pe), pe.shape[0]
        w.out_w, w.out_b = P(self.output_process.poseFinal.weight), P(self.output_process.poseFinal.bias)
        if 'action' in self.cond_mode:
            w.action_emb, w.num_actions = P(self.embed_action.action_embedding), self.num_actions
        if 'text' in self.cond_mode:
            w.text_w, w.text_b, w.clip_dim = P(self.embed_text.weight), P(self.embed_text.bias), self.clip_dim
        esz = 4
        offline = self.arch == 'offline'
        for l, layer in enumerate((self.seqTransEncoder if offline else self.seqTransDecoder).layers):
            lw = w.layers[l]
            lw.qkv_w, lw.qkv_b = P(layer.self_attn.in_proj_weight), P(layer.self_attn.in_proj_bias)
            lw.o_w, lw.o_b = P(layer.self_attn.out_proj.weight), P(layer.self_attn.out_proj.bias)
            lw.l1_w, lw.l1_b = P(layer.linear1.weight), P(layer.linear1.bias)
            lw.l2_w, lw.l2_b = P(layer.linear2.weight), P(layer.linear2.bias)
            lw.n1_w, lw.n1_b = P(layer.norm1.weight), P(layer.norm1.bias)
            lw.n2_w, lw.n2_b = P(layer.norm2.weight), P(layer.norm2.bias)
            if offline:
                continue
            # 1-token memory: only the value rows [2D:3D] of the cross-attention in-projection matter
            lw.xv_w = P(layer.multihead_attn.in_proj_weight) + 2 * D * D * esz
            lw.xv_b = P(layer.multihead_attn.in_proj_bias) + 2 * D * esz
            lw.xo_w, lw.xo_b = P(layer.multihead_attn.out_proj.weight), P(layer.multihead_attn.out_proj.bias)
            lw.n3_w, lw.n3_b = P(layer.norm3.weight), P(layer.norm3.bias)
        _lib.check(L.regen_load_weights(hp, ctypes.byref(w), _lib.stream_ptr(device)), "regen_load_weights")
        self._handle = handle
        return handle

    def _prepare(self, y, B, T, device, guidance):
        """Loop-invariant conditioning (actor motion, action / text embedding); cached while the
        same tensors are passed again, which is what a sampling loop does every step."""
        if y is None:
            raise TypeError("CMDM.forward needs y (a dict with at least 'cmotion')")  # the reference crashes at y.get
        uncond = bool(y.get('uncond', False))
        cmotion = _lib.require_cuda_f32(y['cmotion'], "y['cmotion']")
        if tuple(cmotion.shape) != (B, self.njoints, self.nfeats, T):
            raise ValueError("y['cmotion'] must have shape %s, got %s" % ((B, self.njoints, self.nfeats, T),
                                                                          tuple(cmotion.shape)))
        action = text = None
        if 'action' in self.cond_mode:
            action = y['action'][:, 0].to(device=device, dtype=torch.long).contiguous()
            # the reference's embedding lookup raises on a bad class id (model/cmdm.py:363-365); checked once per
            # conditioning tensor (cached below), not per step
            akey = (id(y['action']), y['action']._version)
            if getattr(self, '_action_checked', None) != akey:
                lo, hi = int(action.min()), int(action.max())
                if lo < 0 or hi >= self.num_actions:
                    raise IndexError("y['action'] holds class ids in [%d, %d]; the model has %d actions"
                                     % (lo, hi, self.num_actions))
                self._action_checked = akey
        if 'text' in self.cond_mode:
            if 'text_embed' in y:
                text = y['text_embed']
            else:
                text = self.encode_text(y['text'])
            text = _lib.require_cuda_f32(text, "text features").contiguous()
        handle = self._get_handle(2 * B if guidance else B, T, device)
        # Cache on the IDENTITY of the conditioning tensors (kept alive below, so an id cannot be recycled) plus their
        # in-place version counters.  Keying on data_ptr() would be wrong: the caching allocator hands the address
        # of a freed tensor to the next one of the same size.
        srcs = (y['cmotion'], y.get('action') if action is not None else None,
                (y.get('text_embed') if 'text_embed' in y else text) if text is not None else None)
        key = (B, T, bool(guidance), uncond) + tuple((id(t), t._version) if t is not None else None for t in srcs)
        if self._cond_key != key:
            cm = cmotion.contiguous()
            rc = _lib.lib().regen_prepare_cond(handle.ptr, _lib.ptr(cm), _lib.ptr(action), _lib.ptr(text), B, T,
                                               int(bool(guidance)), int(uncond), _lib.stream_ptr(device))
            _lib.check(rc, "regen_prepare_cond")
            self._cond_key = key
            self._cond_keepalive = (srcs, cm, action, text)
        return handle

    def _table_steps(self):
        pe_len = self.sequence_pos_encoder.pe.shape[0]
        return 1000 if pe_len >= 1000 else pe_len

    def _check_timesteps(self, lo, hi):
        """The timestep embedding is a table of ``_table_steps()`` rows of pe (model/cmdm.py:291-298 indexes pe[t]):
        out-of-range steps raise here instead of being clamped in the kernel."""
        n = self._table_steps()
        if lo < 0 or hi >= n:
            raise IndexError("timestep %d outside the denoiser's timestep table [0, %d) (negative steps wrap in the "
                             "reference's pe[t] lookup and are not supported)" % (lo if lo < 0 else hi, n))

    def _denoise_tbi(self, handle, x_tbi, t, scale, B, T):
        """x_tbi [T,B,I] contiguous -> x0 [T,B,I] (new tensor)."""
        out = torch.empty_like(x_tbi)
        rc = _lib.lib().regen_denoise(handle.ptr, _lib.ptr(x_tbi), _lib.ptr(t), _lib.ptr(scale), _lib.ptr(out), B, T,
                                      _lib.stream_ptr(x_tbi.device))
        _lib.check(rc, "regen_denoise")
        return out

    def _forward_impl(self, x, timesteps, y, scale=None):
        from .gaussian_diffusion import _to_layout
        _lib.require_cuda_f32(x, "x")
        bs, njoints, nfeats, nframes = x.shape
        if (njoints, nfeats) != (self.njoints, self.nfeats):
            raise ValueError("x must be [B,%d,%d,T], got %s" % (self.njoints, self.nfeats, tuple(x.shape)))
        guidance = scale is not None
        handle = self._prepare(y, bs, nframes, x.device, guidance)
        x_tbi = _to_layout(x, "tbi").permute(3, 0, 1, 2)          # [T,B,J,F] contiguous
        if timesteps.numel() != bs:
            raise ValueError("timesteps must hold one step per sample (%d), got %d" % (bs, timesteps.numel()))
        t = timesteps.to(device=x.device, dtype=torch.long).contiguous()
        if not torch.cuda.is_current_stream_capturing():
            self._check_timesteps(int(t.min()), int(t.max()))   # generic (host-driven) route: one sync per call
        if guidance:
            scale = _cfg_scale(scale, bs, x.device)
        out = self._denoise_tbi(handle, x_tbi, t, scale, bs, nframes)
        # [T,B,J,F] -> [B,J,F,T] as a permuted view, like the reference (model/cmdm.py:353-354)
        return out.view(nframes, bs, njoints, nfeats).permute(1, 2, 3, 0)

    def forward(self, x, timesteps, y=None):
        """x: [batch_size, njoints, nfeats, max_frames] (x_t); timesteps: [batch_size] int -> x_0 prediction.
        model/cmdm.py:173-252."""
        return self._forward_impl(x, timesteps, y, None)

    # ------------------------------------------------------------------------------- fast route
    def regen_sampling_session(self, shape, y, timestep_map, scale=None):
        return SamplingSession(self, shape, y, timestep_map, scale)


_ORIG_RANDN_LIKE = torch.randn_like   # a replaced torch.randn_like (noise-recording test hooks) disables graph replay
_GRAPH_CACHE_MAX = 4
_SEQ_CAP = 4096                       # capacity of the per-graph step sequences (loop length - 1)


def _graph_unroll(n_after_first):
    """Steps per captured graph: REGEN_CUDA_GRAPH=0 disables graphs, =U forces U; default = the U in [6, 12]
    leaving the fewest trailing (step-by-step) steps, larger U on ties."""
    import os
    env = os.environ.get("REGEN_CUDA_GRAPH", "")
    if env == "0":
        return 0
    if env.isdigit() and int(env) > 0:
        return int(env)
    best = None
    for u in range(12, 5, -1):
        rem = n_after_first % u
        if best is None or rem < best[0]:
            best = (rem, u)
    return best[1]


class _StepGraph:
    """A captured CUDA graph of `unroll` consecutive sampling steps on static buffers."""

    def __init__(self):
        self.graph = None
        self.launches = 0


class SamplingSession:
    """Fused sampling loop state: token-major x, hoisted conditioning, one denoise + one update
    kernel sequence per step.  Created by GaussianDiffusion._fast_session.

    Two drivers over the same kernels:
      * step-by-step: the host enqueues every step (the progressive generators, short loops, dump_steps);
      * graph replay (``graph=True``, only from the non-progressive loops, which consume just the final
        sample): after the first step, `unroll` consecutive steps -- device-side timestep bookkeeping
        (regen_step_tables), denoiser, ``torch.randn_like`` noise, posterior update -- are captured once into
        a CUDA graph on static buffers and replayed; this removes the host launch path and most of the
        kernel-to-kernel gaps.  The noise stream is the one the step-by-step driver draws (torch's
        graph-safe Philox offsets), so both drivers return bit-identical samples for a given seed.
    """

    def __init__(self, model, shape, y, timestep_map, scale):
        self.model, self.shape, self.y, self.scale = model, tuple(shape), y, scale
        self.timestep_map = timestep_map

    # ------------------------------------------------------------------------------------------------
    def run(self, diffusion, kind, img, indices, clip_denoised, eta, graph=False, progress=False, unroll=None):
        """Generator over the loop.  Step-by-step it yields once per step; with graph replay it yields once
        per replayed graph (dict key "steps" = steps advanced by that yield)."""
        from .gaussian_diffusion import _to_layout
        m = self.model
        B, J, F, T = self.shape
        dev = img.device
        _lib.require_cuda_f32(img, "initial noise")
        guidance = self.scale is not None
        handle = m._prepare(self.y, B, T, dev, guidance)
        scale = None
        if guidance:
            scale = _cfg_scale(self.scale, B, dev)
        idx = [int(i) for i in indices]
        n = len(idx)
        if n:
            tm = [int(self.timestep_map[i]) for i in idx]
            m._check_timesteps(min(tm), max(tm))                    # host integers: free
        bar = None
        if progress:
            from tqdm.auto import tqdm
            bar = tqdm(total=n)
        x = _to_layout(img, "tbi")                                  # logical [B,J,F,T], memory [T,B,J,F]
        t_model = torch.empty(B, dtype=torch.long, device=dev)
        t_idx = torch.empty(B, dtype=torch.long, device=dev)
        # motion editing (gaussian_diffusion.py:319-323): mask / motion re-laid-out once to the model-output layout
        inpaint = None
        if 'inpainting_mask' in self.y:
            inpaint = (_to_layout(self.y['inpainting_mask'].to(device=dev, dtype=torch.float32), "tbi"),
                       _to_layout(_lib.require_cuda_f32(self.y['inpainted_motion'].to(dev), "y['inpainted_motion']"), "tbi"))

        def blend(x0_tbi, mask_tbi, motion_tbi):
            _lib.check(_lib.lib().regen_inpaint_blend(_lib.ptr(x0_tbi), _lib.ptr(motion_tbi), _lib.ptr(mask_tbi),
                                                      x0_tbi.numel(), _lib.stream_ptr(dev)), "regen_inpaint_blend")
        self._blend = blend

        def eager_step(x, i, first):
            t_idx.fill_(i)
            t_model.fill_(int(self.timestep_map[i]))                # respace.py:125-126 (integer remap)
            x0 = m._denoise_tbi(handle, x.permute(3, 0, 1, 2), t_model, scale, B, T)
            if inpaint is not None:
                blend(x0, inpaint[0], inpaint[1])
            x0 = x0.view(T, B, J, F).permute(1, 2, 3, 0)
            # the reference draws randn_like(x) AFTER the model call, in x's memory layout: contiguous
            # [B,J,F,T] at the first step, the permuted model-output layout from then on
            noise = torch.randn_like(img if first else x)
            return diffusion._update("p" if kind == "p" else "ddim", x, x0, noise, t_idx, clip_denoised, eta=eta)

        with torch.no_grad():
            U = 0
            if graph and n >= 2 and torch.randn_like is _ORIG_RANDN_LIKE and n - 1 <= _SEQ_CAP:
                U = unroll if unroll is not None else _graph_unroll(n - 1)
                if U and (n - 1) // U < 2:
                    U = 0                                           # not worth a capture
            k = 0
            if n:
                x, pred = eager_step(x, idx[0], True)
                k = 1
                if bar is not None:
                    bar.update(1)
                yield {"sample": x, "pred_xstart": pred, "steps": 1}
            if U:
                st = self._graph_state(diffusion, kind, handle, scale, clip_denoised, eta, U, dev, inpaint is not None)
                st["x"].copy_(x)
                if inpaint is not None:
                    st["inp_mask"].copy_(inpaint[0])
                    st["inp_motion"].copy_(inpaint[1])
                rest = idx[1:]
                st["seq_idx"][:len(rest)].copy_(torch.tensor(rest, dtype=torch.long), non_blocking=False)
                st["seq_model"][:len(rest)].copy_(torch.tensor([int(self.timestep_map[i]) for i in rest],
                                                               dtype=torch.long))
                st["pos"].zero_()
                if scale is not None:
                    st["scale"].copy_(scale)
                sg = st["graph"]
                if sg.graph is None:
                    self._capture(st, diffusion, kind, handle, clip_denoised, eta, U)
                L = _lib.lib()
                for _ in range((n - 1) // U):
                    sg.graph.replay()
                    L.regen_launch_count_add(sg.launches)
                    k += U
                    if bar is not None:
                        bar.update(U)
                    yield {"sample": st["x"], "pred_xstart": st["pred"], "steps": U}
                # detach the result from the static buffers (the next loop on this model reuses them)
                x, pred = st["x"].clone(), st["pred"].clone()
                if k == n:
                    yield {"sample": x, "pred_xstart": pred, "steps": 0}
            while k < n:
                x, pred = eager_step(x, idx[k], False)
                k += 1
                if bar is not None:
                    bar.update(1)
                yield {"sample": x, "pred_xstart": pred, "steps": 1}
        if bar is not None:
            bar.close()

    # ------------------------------------------------------------------------------------------------
    def _graph_state(self, diffusion, kind, handle, scale, clip_denoised, eta, U, dev, inpainting=False):
        """Static buffers + graph for this (handle, problem, sampler) combination, cached on the model."""
        m = self.model
        B, J, F, T = self.shape
        key = (id(handle), id(diffusion), B, T, kind, bool(clip_denoised), float(eta), U, scale is not None,
               m._cond_key[:4] if m._cond_key else None, handle.ptr.value, bool(inpainting))
        cache = m.__dict__.setdefault("_graph_cache", {})
        st = cache.get(key)
        if st is not None and st["handle"] is handle and st["diffusion"] is diffusion:
            cache[key] = cache.pop(key)                              # most recently used last
            return st
        while len(cache) >= _GRAPH_CACHE_MAX:
            cache.pop(next(iter(cache)))
        st = {
            "handle": handle, "diffusion": diffusion, "graph": _StepGraph(),
            "x": torch.empty((T, B, J, F), device=dev, dtype=torch.float32).permute(1, 2, 3, 0),
            "pred": None,
            "seq_idx": torch.zeros(_SEQ_CAP, dtype=torch.long, device=dev),
            "seq_model": torch.zeros(_SEQ_CAP, dtype=torch.long, device=dev),
            "pos": torch.zeros(1, dtype=torch.long, device=dev),
            "t_idx": torch.empty(B, dtype=torch.long, device=dev),
            "t_model": torch.empty(B, dtype=torch.long, device=dev),
            "scale": torch.empty_like(scale) if scale is not None else None,
            "inp_mask": torch.empty((T, B, J, F), device=dev).permute(1, 2, 3, 0) if inpainting else None,
            "inp_motion": torch.empty((T, B, J, F), device=dev).permute(1, 2, 3, 0) if inpainting else None,
        }
        cache[key] = st
        return st

    def _capture(self, st, diffusion, kind, handle, clip_denoised, eta, U):
        m = self.model
        B, J, F, T = self.shape
        L = _lib.lib()
        dev = st["x"].device
        diffusion._tables(dev)                                       # device tables exist before capture
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = L.regen_launch_count()
        with torch.cuda.graph(g):
            xs = st["x"]
            pred = None
            for u in range(U):
                _lib.check(L.regen_step_tables(_lib.ptr(st["seq_idx"]), _lib.ptr(st["seq_model"]), _lib.ptr(st["pos"]),
                                               _lib.ptr(st["t_idx"]), _lib.ptr(st["t_model"]), B, _SEQ_CAP,
                                               _lib.stream_ptr(dev)), "regen_step_tables")
                x0 = m._denoise_tbi(handle, xs.permute(3, 0, 1, 2), st["t_model"], st["scale"], B, T)
                if st["inp_mask"] is not None:
                    self._blend(x0, st["inp_mask"], st["inp_motion"])
                x0 = x0.view(T, B, J, F).permute(1, 2, 3, 0)
                noise = torch.randn_like(xs)
                xs, pred = diffusion._update("p" if kind == "p" else "ddim", xs, x0, noise, st["t_idx"], clip_denoised,
                                             eta=eta, out=st["x"] if u == U - 1 else None)
            st["pred"] = pred
        st["graph"].graph = g
        import os
        if os.environ.get("REGEN_DEBUG_GRAPH"):
            import sys
            print("[regen] captured %d-step graph (B=%d T=%d, %d kernels)" % (U, B, T, L.regen_launch_count() - n0),
                  file=sys.stderr)
        st["graph"].launches = int(L.regen_launch_count() - n0)
        # the capture pass enqueued nothing: undo its contribution to the executed-launch counter
        L.regen_launch_count_add(-st["graph"].launches)

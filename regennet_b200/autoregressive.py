"""Auto-regressive ("online") inference -- the reference's per-frame sampling of
``eval/a2m/stgcn_eval.py:50-67`` (``auto_regressive=True``), re-designed for the causal denoiser.

Reference semantics (one full sampling loop per frame, T loops)::

    cmotion = zeros_like(cmotion_bak)
    for f in range(T):
        cmotion[..., f] = cmotion_bak[..., f]          # the actor's motion is revealed frame by frame
        sample = sample_fn(model, shape, clip_denoised=False, model_kwargs={'y': {..., 'cmotion': cmotion}})
        output[..., f] = cat((cmotion, sample), axis=2)[..., f]      # 'cmdm' setting; else sample[..., f]

Every loop f draws fresh noise and keeps only frame f of its result, so the T loops are independent
sampling problems.  Two facts of ``arch='online'`` (model/cmdm.py:168-171, 220-227: causal self-attention,
per-frame embeddings, 1-token memory) let this be restructured without changing what frame f is computed from:

  * frame f of the denoiser output depends on frames <= f of x_t and cmotion only, at every step of the
    loop -- so loop f can run on the first f + 1 frames alone (``truncate=True``; Sum_f (f+1) = T(T+1)/2
    instead of T*T frame-steps);
  * independent loops can share launches: ``frames_per_call = G`` stacks G consecutive loops along the batch
    axis (sub-batch g sees cmotion revealed up to its own frame f_g, exactly as in the reference, and runs
    f_max + 1 frames), which turns T small launches per step into T/G full ones.

With ``truncate=False, frames_per_call=1`` this is the reference loop literally.  In every mode frame f of
the output is the reference's function of (the noise of loop f restricted to frames <= f, cmotion[..., :f+1]);
tests/test_gpu_autoregressive.py checks that against the oracle with injected noise.
"""
import torch


def _rep_batch(v, B, G):
    """Conditioning entries with a leading batch axis are repeated for the G stacked loops."""
    if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B:
        return v.repeat((G,) + (1,) * (v.dim() - 1))
    if isinstance(v, (list, tuple)) and len(v) == B:
        return type(v)(list(v) * G)
    return v


def _model_arch(model):
    """arch of the denoiser behind `model`, unwrapping ClassifierFreeSampleModel (``.model``); None when unknown."""
    seen = 0
    while model is not None and seen < 4:
        arch = getattr(model, 'arch', None)
        if arch is not None:
            return arch
        model = getattr(model, 'model', None)
        seen += 1
    return None


def _cut_time(v, shape, Tc):
    """Per-frame conditioning entries (same [B, V, C, T] shape as the motion: inpainting mask / motion) follow the
    truncated time axis."""
    if torch.is_tensor(v) and tuple(v.shape) == tuple(shape) and Tc < shape[-1]:
        return v[..., :Tc]
    return v


def auto_regressive_sample(sample_fn, model, shape, model_kwargs, setting='cmdm', truncate=None, frames_per_call=1,
                           clip_denoised=False, **sample_kwargs):
    """-> output [B, V, 2C, T] (``setting='cmdm'``: actor motion | generated reaction, eval/a2m/stgcn_eval.py:61-62)
    or [B, V, C, T].

    sample_fn: ``diffusion.p_sample_loop`` / ``ddim_sample_loop`` (called as in the reference, :58).
    model_kwargs['y']['cmotion'] is the full actor motion [B, V, C, T]; as in the reference it is left in
    model_kwargs on return.  Extra keyword arguments go to sample_fn.

    truncate: run loop f on the first f + 1 frames only.  That is the same function of the inputs ONLY for a causal
    denoiser (``arch='online'``); for ``arch='offline'`` (bidirectional encoder, model/cmdm.py:228-238) or an unknown
    model, frame f of the reference also depends on the noise of later frames, so ``truncate=True`` raises there and
    the default (``None``) resolves to True for 'online' and to False -- the literal reference loop -- otherwise.
    """
    arch = _model_arch(model)
    causal = arch == 'online'
    if truncate is None:
        truncate = causal
    if truncate and not causal:
        raise ValueError("auto_regressive_sample(truncate=True) needs a causal denoiser (arch='online'); this model has "
                         "arch=%r, whose frame f depends on later frames (eval/a2m/stgcn_eval.py:50-67 semantics): "
                         "pass truncate=False" % (arch,))
    y = model_kwargs['y']
    cmotion_bak = y['cmotion']
    B, V, C, T = cmotion_bak.shape
    assert tuple(shape) == (B, V, C, T), "shape must be cmotion's shape (eval/a2m/stgcn_eval.py:52-58)"
    G = max(1, min(int(frames_per_call), T))
    dev = cmotion_bak.device
    out_c = 2 * C if setting == 'cmdm' else C
    output = torch.zeros((B, V, out_c, T), device=dev, dtype=cmotion_bak.dtype)
    for f0 in range(0, T, G):
        frames = list(range(f0, min(f0 + G, T)))
        g = len(frames)
        Tc = frames[-1] + 1 if truncate else T
        # sub-batch j of the stacked call sees the actor's frames 0..frames[j], zeros afterwards (:53-56)
        cm = torch.zeros((g, B, V, C, Tc), device=dev, dtype=cmotion_bak.dtype)
        for j, f in enumerate(frames):
            cm[j, ..., :f + 1] = cmotion_bak[..., :f + 1]
        yk = {k: _rep_batch(_cut_time(v, (B, V, C, T), Tc), B, g) for k, v in y.items() if k != 'cmotion'}
        yk['cmotion'] = cm.view(g * B, V, C, Tc)
        kw = dict(model_kwargs)
        kw['y'] = yk
        sample = sample_fn(model, (g * B, V, C, Tc), clip_denoised=clip_denoised, model_kwargs=kw, **sample_kwargs)
        sample = sample.reshape(g, B, V, C, Tc)
        for j, f in enumerate(frames):
            if setting == 'cmdm':
                output[:, :, :C, f] = cmotion_bak[..., f]
                output[:, :, C:, f] = sample[j, ..., f]
            else:
                output[..., f] = sample[j, ..., f]
    y['cmotion'] = cmotion_bak.clone()
    return output

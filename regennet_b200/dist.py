"""Multi-GPU sampling: independent motion sequences sharded over ranks, one all-gather at the end.

The reference's sampling path issues no collective; its only parallelism is DDP for
training (train/training_loop.py:115-124) bootstrapped through mpi4py (utils/dist_util.py:20-42).
Samples in a batch never interact (attention is within a sample, LayerNorm per token, the update
elementwise), so inference shards the batch: rank r owns rows [r*B/W, (r+1)*B/W) of every
per-sample tensor in ``model_kwargs['y']``, seeds its RNG with ``seed + r``, runs the whole loop with
zero communication, and ONE ``all_gather_into_tensor`` reassembles [B,J,F,T] in rank order.

One process per GPU (torchrun / env:// rendezvous); backend 'nccl' on GPUs, 'gloo' in CPU tests.
"""
import os

import torch
import torch.distributed as dist


def setup_dist(backend=None):
    """torchrun-style bootstrap (replaces utils/dist_util.py:20-42's mpi4py bootstrap)."""
    if dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        # binding the communicator to the device up front avoids the lazy init inside the first collective
        dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        return
    dist.init_process_group(backend, rank=rank, world_size=world)


def dev():
    """utils/dist_util.py:45-51."""
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_bounds(B, rank, world):
    """Contiguous, balanced split of B samples; the first B % world ranks take one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_kwargs(y, B, rank, world, device=None):
    """Slice every tensor in the conditioning dict whose leading dim is the batch; share the rest.  With `device`, the
    tensor entries (sliced or not) are moved there -- the shard of a (pinned) host-resident global batch is the only part
    of it this rank ever copies."""
    lo, hi = shard_bounds(B, rank, world)
    out = {}
    for k, v in y.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == B:
            out[k] = v[lo:hi]
        else:
            out[k] = v
        if device is not None and torch.is_tensor(out[k]):
            out[k] = out[k].to(device, non_blocking=True)
    return out


def gather_batch(local, B, group=None, timing=None):
    """All-gather the per-rank shards [b_r, ...] into [B, ...] in rank order (shards may be ragged).  `timing` (a dict,
    CUDA only) receives a CUDA event pair around the collective under "collective_events"."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(B, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    local = local.contiguous()
    assert local.shape[0] == counts[rank]
    if len(set(counts)) == 1:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        ev = None
        if timing is not None and local.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        dist.all_gather_into_tensor(out, local, group=group)
        if ev is not None:
            ev[1].record()
            timing["collective_events"] = ev
        return out
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:counts[rank]] = local
    buf = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)


def sharded_sample(sample_fn, model, shape, model_kwargs, seed=None, device=None, timing=None, **kw):
    """Run ``sample_fn`` (``diffusion.p_sample_loop`` or ``ddim_sample_loop``) on this rank's shard and
    return the full batch on every rank.  ``shape`` and ``model_kwargs`` describe the GLOBAL batch; the conditioning
    tensors may live on the host (pass ``device`` and only this rank's shard is copied to it).  ``timing``: see
    gather_batch; it also receives this rank's shard under "local"."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B = shape[0]
    lo, hi = shard_bounds(B, rank, world)
    if seed is not None:
        torch.manual_seed(seed + rank)
    local_kwargs = dict(model_kwargs or {})
    if isinstance(local_kwargs.get("y"), dict):
        local_kwargs["y"] = shard_kwargs(local_kwargs["y"], B, rank, world, device=device)
    local = sample_fn(model, (hi - lo,) + tuple(shape[1:]), model_kwargs=local_kwargs, **kw)
    if timing is not None:
        timing["local"] = local
    return gather_batch(local, B, timing=timing)

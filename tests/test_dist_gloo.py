"""world_size-2 gloo test (CPU) of the sharding / gather host logic used for multi-GPU sampling."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from regennet_b200 import dist as rdist


def test_shard_bounds_cover_batch_exactly():
    for B in [1, 2, 7, 256, 2048, 2049]:
        for W in [1, 2, 3, 8]:
            spans = [rdist.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    rdist.setup_dist("gloo")
    y = {"cmotion": torch.arange(B * 6, dtype=torch.float32).view(B, 1, 2, 3), "scale": torch.arange(B).float(),
         "flag": True, "text": ["s%d" % i for i in range(B)]}

    def fake_sample(model, shape, model_kwargs=None, **kw):
        yy = model_kwargs["y"]
        assert yy["cmotion"].shape[0] == shape[0] == len(yy["text"]) == yy["scale"].shape[0]
        assert yy["flag"] is True
        # deterministic per-sample function + the per-rank RNG stream
        return yy["cmotion"] * 2.0 + torch.randn(shape) * 0.0 + yy["scale"].view(-1, 1, 1, 1)

    out = rdist.sharded_sample(fake_sample, None, (B, 1, 2, 3), {"y": y}, seed=10)
    want = y["cmotion"] * 2.0 + y["scale"].view(-1, 1, 1, 1)
    ok = torch.equal(out, want)
    # per-rank seeds differ
    torch.manual_seed(10 + rank)
    r = torch.rand(1)
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, r)
    ok = ok and (gathered[0] != gathered[1]).item()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])  # even and ragged shards
def test_sharded_sample_world2_gloo(tmp_path, B):
    port = 29600 + (os.getpid() % 200) + B
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"

"""CPU, world_size 2, gloo: the batch-shard + all-gather plumbing of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muggled_dpt_b200.distributed import all_gather_depth, shard_range, sharded_forward


def test_shard_range_covers_batch_exactly():
    for B in (1, 4, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _FakeModel:
    """stands in for DPTModel on CPU: a per-frame function, so sharding must not change the result"""

    def __call__(self, x):
        return x.sum(dim=1) * 2.0 + 1.0


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        imgs = torch.randn(B, 3, 6, 8, generator=g)
        full = _FakeModel()(imgs)
        got = sharded_forward(_FakeModel(), imgs)
        ok = torch.equal(got, full)
        lo, hi = shard_range(B, rank, world)
        got2 = all_gather_depth(full[lo:hi].clone(), B)
        ok = ok and torch.equal(got2, full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])  # even and uneven shards
def test_sharded_forward_gloo_world2(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]

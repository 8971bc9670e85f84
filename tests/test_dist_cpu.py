"""CPU, world_size 2 over gloo: region sharding and the ragged all-gather of predictions (the N>1 path of bench.py)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pepper_b200.dist import shard_regions, gather_predictions


def test_shard_regions_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    work = rng.integers(1000, 5000, size=37)
    for world in (1, 2, 4, 8):
        blocks = shard_regions(work, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 37
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        loads = [work[b:e].sum() for b, e in blocks]
        assert max(loads) <= work.sum() / world + work.max()
    assert shard_regions(np.zeros(0), 4) == [(0, 0)] * 4
    assert shard_regions(np.array([5, 5]), 4)[-1][1] == 2


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5 + 3 * rank                                  # ragged
    probs = torch.zeros((16, 3), dtype=torch.float32)
    probs[:n] = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) + 100 * rank
    allp, counts = gather_predictions(probs, n, world)
    q.put((rank, allp.numpy().copy(), counts))
    dist.destroy_process_group()


def test_gather_predictions_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    want = np.concatenate([np.arange(5 * 3, dtype=np.float32).reshape(5, 3), np.arange(8 * 3, dtype=np.float32).reshape(8, 3) + 100])
    for rank, allp, counts in res:
        assert counts == [5, 8]
        assert np.array_equal(allp, want)          # rank-major == region order

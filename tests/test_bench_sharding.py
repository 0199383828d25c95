"""world_size-2 gloo test (CPU) of the bench's multi-rank plumbing: frame sharding and the MAX reduction of times."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = bench.shard_frames(5, rank, world)
    t = bench.reduce_max(1.0 + rank, dist)
    dist.destroy_process_group()
    q.put((rank, seeds, t))


def test_two_rank_sharding_and_max_reduce():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == [0, 1, 2, 3, 4] and res[1][1] == [5, 6, 7, 8, 9]      # disjoint, weak scaling
    assert res[0][2] == 2.0 and res[1][2] == 2.0                                # max over ranks


def test_single_process_helpers():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.shard_frames(3, 0, 1) == [0, 1, 2]
    assert bench.reduce_max(3.5) == 3.5

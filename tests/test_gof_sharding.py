"""Frame sharding of the inter-frame path (BASELINE configs[2]): frames round-robin over ranks, the
simplified cloud of frame f sent to the owner of frame f + 1.  World-size-2 gloo test on the CPU with the oracle as the
per-rank codec: the P streams every rank produces equal those of a single process."""
import os
import sys

import numpy as np

from cwi_pcl_codec_b200 import gof, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NF, NP, BITS = 5, 8000, 7


def _recs(c):
    return np.ascontiguousarray(c).view(np.uint8).reshape(-1, 32)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = [_recs(c) for c in synth.gen_gof(NP, seed=11, frames=NF)]
    p = O.default_params(octree_bits=BITS)
    local = {}
    for f in gof.owned_frames(NF, rank, world):                  # the oracle stands in for the per-rank GPU encoder
        _, _, dbg = O.encode(frames[f], p, debug=True)
        local[f] = torch.from_numpy(dbg["output_cloud"].reshape(-1).copy())
    pred = gof.exchange_predictors(local, NF, dist)
    out = {}
    for g, t in pred.items():
        i_s, p_s, info = O.encode_delta(t.numpy().reshape(-1, 32), frames[g], p)
        out[g] = (i_s, p_s)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_helpers():
    assert gof.owned_frames(30, 3, 8) == [3, 11, 19, 27] and gof.owner(17, 8) == 1
    assert gof.exchange_predictors({0: "a", 1: "b", 2: "c"}, 3, None) == {1: "a", 2: "b"}


def test_predictors_travel_to_the_neighbour_rank(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, out = q.get(timeout=180)
        assert sorted(out) == [g for g in gof.owned_frames(NF, rank, 2) if g >= 1]
        got.update(out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = [_recs(c) for c in synth.gen_gof(NP, seed=11, frames=NF)]
    prm = oracle.default_params(octree_bits=BITS)
    for g in range(1, NF):
        _, _, dbg = oracle.encode(frames[g - 1], prm, debug=True)
        i_s, p_s, _ = oracle.encode_delta(dbg["output_cloud"], frames[g], prm)
        assert got[g] == (i_s, p_s), g


def _worker3(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nf = 8
    local = {f: torch.full((32 * (f + 1),), f, dtype=torch.uint8) for f in gof.owned_frames(nf, rank, world)}   # "cloud" f: f + 1 records of value f
    pred = gof.exchange_predictors(local, nf, dist)
    q.put((rank, {g: (int(t.numel()), int(t[0]) if t.numel() else -1, bool((t == t[0]).all())) for g, t in pred.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_three_ranks_every_frame_gets_its_predecessor():
    """World size 3: every rank receives from its left neighbour and sends to its right one, several frames per pair, in frame order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker3, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(3):
        rank, out = q.get(timeout=120)
        assert sorted(out) == [g for g in gof.owned_frames(8, rank, 3) if g >= 1]
        got.update(out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g in range(1, 8):
        assert got[g] == (32 * g, g - 1, True), g

"""Tile mode (BASELINE configs[3]): every tile stream is bit-exact against the oracle run on that tile's
points in their original relative order; the union of the decoded tiles is the decoded frame; tiles shard round-robin over
ranks and the streams are gathered on a writer rank (world-size-2 gloo test on the CPU, with the oracle as the encoder)."""
import os
import sys

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth, tiles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_ids_partition_the_unit_cube():
    cl = synth.gen_surface(5000, 1)
    xyz = np.stack([cl["x"], cl["y"], cl["z"]], 1)
    t3, t6 = tiles.tile_ids(xyz, 3), tiles.tile_ids(xyz, 6)
    assert t3.min() >= 0 and t3.max() <= 7 and t6.max() <= 63
    assert np.array_equal(t6 >> 3, t3)                                  # a 64-tile id refines the 8-tile id (octree child order)
    assert np.array_equal(t3, ((xyz[:, 0] >= 0.5).astype(int) << 2) | ((xyz[:, 1] >= 0.5).astype(int) << 1) | (xyz[:, 2] >= 0.5).astype(int))
    far = np.array([[-3.0, 0.2, 0.2], [7.0, 0.9, 0.1], [np.nan, 0.9, 0.9]], np.float32)
    assert list(tiles.tile_ids(far, 3)) == [0, 6, 0]                    # clamped; non-finite -> tile 0
    assert tiles.owned_tiles(3, 1, 4) == [1, 5] and tiles.owned_tiles(6, 7, 8) == list(range(7, 64, 8))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cl = synth.gen_surface(30000, 9)
    tid = tiles.tile_ids(np.stack([cl["x"], cl["y"], cl["z"]], 1), 3)
    p = O.default_params(octree_bits=8)
    local = {}
    for k, t in enumerate(tiles.owned_tiles(3, rank, world)):          # the oracle stands in for the per-rank GPU encoder
        sub = cl[tid == t]
        if sub.shape[0]:
            local[t] = O.encode(sub, p, frame_id=k + 1)[0]
    allt = tiles.gather_tile_streams(local, 3, dist)
    if rank == 0:
        q.put({t: s for t, s in allt.items()})
    else:
        assert allt is None
    dist.barrier()
    dist.destroy_process_group()


def test_tiles_shard_over_two_ranks_and_gather_on_the_writer(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cl = synth.gen_surface(30000, 9)
    tid = tiles.tile_ids(np.stack([cl["x"], cl["y"], cl["z"]], 1), 3)
    prm = oracle.default_params(octree_bits=8)
    for t in range(8):
        sub = cl[tid == t]
        if not sub.shape[0]:
            assert t not in got
            continue
        assert got[t] == oracle.encode(sub, prm, frame_id=t // 2 + 1)[0], t      # rank r's k-th tile carries frame id k + 1


@pytest.mark.gpu
@pytest.mark.parametrize("tile_bits", [3, 6])
def test_tile_streams_are_bit_exact_per_tile_and_decode_to_the_frame(oracle, tile_bits):
    from cwi_pcl_codec_b200 import codec as K
    cl = synth.gen_surface(200000, 4)
    cl["x"][17] = np.nan
    kp = K.default_params(octree_bits=10)
    c = K.Codec(kp)
    part, offs = c.split_tiles(cl, tile_bits)
    tid = tiles.tile_ids(np.stack([cl["x"], cl["y"], cl["z"]], 1), tile_bits)
    raw = cl.view(np.uint8).reshape(-1, 32)
    for t in range(1 << tile_bits):
        assert np.array_equal(part[offs[t]:offs[t + 1]], raw[tid == t]), t            # stable: original relative order inside a tile
    c.frame_id = 0
    streams, npts = c.encode_tiles(cl, tile_bits)
    assert npts == [int((tid == t).sum()) for t in range(1 << tile_bits)]
    prm = oracle.default_params(octree_bits=10)
    fid, dec_all = 0, []
    for t in range(1 << tile_bits):
        sub = cl[tid == t]
        ref = oracle.encode(sub, prm, frame_id=fid + 1)[0] if sub.shape[0] else b""
        if ref:
            fid += 1
            assert streams[t] == ref, t
            dec_all.append(oracle.decode(ref)[0])
        else:
            assert t not in streams
    dec = c.decode_batch([streams[t] for t in sorted(streams)])
    assert all(np.array_equal(a, b) for a, b in zip(dec, dec_all))
    # every finite input point lies in exactly one tile, so the tiles' point counts add up to the frame
    assert sum(npts) == cl.shape[0]
    # rank sharding: two "ranks" on this GPU produce the same streams as one
    c.frame_id = 0
    s0, _ = c.encode_tiles(cl, tile_bits, first_tile=0, tile_step=2)
    c.frame_id = 0
    s1, _ = c.encode_tiles(cl, tile_bits, first_tile=1, tile_step=2)
    assert set(s0) | set(s1) == set(streams) and not (set(s0) & set(s1))
    for t, s in list(s0.items()) + list(s1.items()):
        assert s[52:] == streams[t][52:]                                               # same bytes but for the rank-local frame id
    c.close()

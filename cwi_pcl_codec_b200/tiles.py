"""Tile mode across the ranks of a node (BASELINE configs[3]): the tiles of one frame are encoded
round-robin by the ranks (tile t on rank t mod world, no communication while encoding) and the variable-length streams are
gathered on a writer rank -- the one real exchange step of the intra path: sizes by all_gather, payloads by grouped
point-to-point sends (NCCL over NVLink when the tensors are on the GPU, gloo on the CPU).

The per-rank work is ``Codec.encode_tiles(cloud, tile_bits, first_tile=rank, tile_step=world)`` (CUDA, through the C ABI);
this module only holds the host-side logic: the tile definition in numpy (for tests) and the gather.
"""
import numpy as np


def tile_ids(xyz, tile_bits):
    """Tile of every point: Morton index (x most significant, like an octree child) of floor(p * 2^k) per axis, clamped
    to the unit cube, k = tile_bits / 3; non-finite points go to tile 0 (the encoder drops them)."""
    k = tile_bits // 3
    xyz = np.asarray(xyz, np.float32)
    fin = np.isfinite(xyz).all(1)
    q = np.clip(np.floor(np.where(fin[:, None], xyz, 0) * np.float32(1 << k)), 0, (1 << k) - 1).astype(np.int64)
    t = np.zeros(xyz.shape[0], np.int64)
    for b in range(k - 1, -1, -1):
        t = (t << 3) | (((q[:, 0] >> b) & 1) << 2) | (((q[:, 1] >> b) & 1) << 1) | ((q[:, 2] >> b) & 1)
    return np.where(fin, t, 0)


def owned_tiles(tile_bits, rank, world):
    return list(range(rank, 1 << tile_bits, world))


def gather_tile_streams(local, tile_bits, dist, device="cpu", dst=0):
    """local: {tile: bytes} of the tiles this rank encoded.  Returns {tile: bytes} of ALL tiles on rank `dst` (None
    elsewhere).  One all_gather of the size table, then one batch of point-to-point transfers into the writer."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    nt = 1 << tile_bits
    sizes = torch.zeros(nt, dtype=torch.int64, device=device)
    for t, s in local.items():
        sizes[t] = len(s)
    table = [torch.zeros(nt, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(table, sizes)
    table = torch.stack(table).cpu().numpy()                # table[r][t] = bytes of tile t on rank r (0 if not its tile)
    mine = b"".join(local[t] for t in sorted(local))
    ops, recv = [], {}
    if rank == dst:
        for r in range(world):
            tot = int(table[r].sum())
            if r == dst or tot == 0:
                continue
            recv[r] = torch.empty(tot, dtype=torch.uint8, device=device)
            ops.append(dist.P2POp(dist.irecv, recv[r], r))
    elif len(mine):
        payload = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
        ops.append(dist.P2POp(dist.isend, payload, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != dst:
        return None
    out = dict(local)
    for r, buf in recv.items():
        b = buf.cpu().numpy().tobytes()
        o = 0
        for t in range(nt):
            ln = int(table[r][t])
            if ln:
                out[t] = b[o:o + ln]
                o += ln
    return out

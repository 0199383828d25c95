"""Inter-frame coding of a group of frames across the ranks of a node (BASELINE configs[2]): frame f lives
on rank f mod world; the P coding of frame f + 1 needs the simplified cloud of frame f ([PCL] getOutputCloud, eval.hpp:862),
which its owner sends to the neighbour -- the one exchange of this path, point to point (NCCL over NVLink when the tensors
are on the GPU, gloo on the CPU): a table of point counts by all_gather, then one batch of sends and receives.

The per-rank work is the codec's (``Codec.encode_batch`` + ``Codec.output_cloud`` for the I frames, ``Codec.encode_delta`` /
``decode_delta`` for the P coding, all CUDA behind the C ABI); this module only holds the host-side plumbing.
"""


def owner(frame, world):
    return frame % world


def owned_frames(nframes, rank, world):
    return list(range(rank, nframes, world))


def exchange_predictors(local, nframes, dist, device="cpu"):
    """local: {f: uint8 tensor (V_f * 32 bytes)} = the simplified clouds of the frames this rank coded intra.
    Returns {g: tensor} for every frame g >= 1 this rank owns: the simplified cloud of frame g - 1 (its own tensor when
    both frames are local)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return {g: local[g - 1] for g in range(1, nframes) if g - 1 in local}
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = torch.zeros(nframes, dtype=torch.int64, device=device)
    for f, t in local.items():
        sizes[f] = t.numel()
    table = [torch.zeros(nframes, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(table, sizes)
    table = torch.stack(table).sum(0).cpu().tolist()            # bytes of every frame's simplified cloud
    ops, out = [], {}
    for g in owned_frames(nframes, rank, world):
        if g == 0:
            continue
        src = owner(g - 1, world)
        if src == rank:
            out[g] = local[g - 1]
        else:
            out[g] = torch.empty(int(table[g - 1]), dtype=torch.uint8, device=device)
            if table[g - 1]:
                ops.append(dist.P2POp(dist.irecv, out[g], src))
    for f, t in sorted(local.items()):
        if f + 1 < nframes and owner(f + 1, world) != rank and t.numel():
            ops.append(dist.P2POp(dist.isend, t, owner(f + 1, world)))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out

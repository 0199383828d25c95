"""bench.py -- Mpoints/s encode+decode @ depth-11 intra on synthetic 1M-point XYZRGB frames (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (the oracle port, all host cores)

  python bench.py --mode decode ...                        BASELINE configs[4]: decode-only on oracle-produced streams

A "step" is one pass of the hot path over one batch of F frames: every cloud is encoded and the stream just produced is
decoded again, like evaluate_compression's per-frame loop (eval.hpp:818-843) -- one ccv2_submit_roundtrip call per step.
Steps are submitted back to back (two calls in flight, ccv2_submit_* / ccv2_wait), so the next step's uploads and
parallel kernels overlap the serial range-coder stages of the previous one; the timed region is bracketed by barrier +
synchronize and holds exactly K complete steps.  --value-api separate runs ccv2_encode_batch then ccv2_decode_batch per
step instead (synchronous; it also provides the encode / decode split that is always reported).
`value` is measured with the clouds already resident in HBM (device pointers in, device pointers out); `e2e` is the same
work through the C ABI with pinned HOST buffers, so the host->device copy of every cloud and the device->host copy of
every stream and decoded cloud are inside the timed region (the stream itself stays on the device between encoder and
decoder: 1.2 MB per frame that a file-based caller would upload again are not counted).  For N > 1 the frames are sharded
over the ranks (no data-path collective: intra frames are independent); times are reduced with MAX.
"""
import argparse
import ctypes as C
import json
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see csrc/ccv2_api.cu (stream -> hardware queue aliasing)
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpoints/s encode+decode @ depth-11 intra; bitstream bit-exact vs ref"
INFLIGHT_CALLS = int(os.environ.get("BENCH_INFLIGHT_CALLS", "3"))     # steps submitted ahead (ccv2_submit_*): the third queues behind the first on the 16 work streams, so a set is taken again the moment it is free


def shard_frames(n_frames_per_rank, rank, world):
    """Seeds of the frames rank `rank` owns (weak scaling: every rank processes n_frames_per_rank distinct frames)."""
    return [rank * n_frames_per_rank + i for i in range(n_frames_per_rank)]


def reduce_max(value, dist=None, device=None):
    """MAX over ranks of a python float (torch.distributed all_reduce when initialised)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gen_frames(kind, n, seeds):
    from concurrent.futures import ThreadPoolExecutor
    from cwi_pcl_codec_b200 import synth
    gen = synth.gen_surface if kind == "surf" else synth.gen_uniform
    with ThreadPoolExecutor(max_workers=min(len(seeds), os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda s: gen(n, s), seeds))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_run(frames, bits, threads):
    """encode+decode every frame with the oracle on `threads` host threads; returns (seconds, total points)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    p = O.default_params(octree_bits=bits)

    def one(f):
        data, _ = O.encode(f, p, frame_id=1)
        O.decode(data)
        return f.shape[0]
    t = time.perf_counter()
    if threads <= 1:
        pts = sum(one(f) for f in frames)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            pts = sum(ex.map(one, frames))
    return time.perf_counter() - t, pts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = max(2, min(2 * cores, 64))
    frames = gen_frames(args.kind, args.points, list(range(min(per_step, 16))))
    frames = [frames[i % len(frames)] for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_oracle_run(frames[:cores], args.bits, cores)
    tot_t, tot_p = 0.0, 0
    for _ in range(args.steps):
        t, p = cpu_oracle_run(frames, args.bits, cores)
        tot_t += t; tot_p += p
    v = tot_p / tot_t / 1e6
    sample = "%d frames of %d points per step, oracle port (restated reference algorithm, gcc -O3), %d host threads" % (per_step, args.points, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32/f64 (integer codec, FP64 keys)",
            "data": "synthetic", "config": workload_config(args, per_step),
            "cpu_baseline": {"value": v, "unit": "Mpoints/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, frames):
    return {"workload": "BASELINE.json configs[1]: %d-point synthetic XYZRGB frames (G-%s, SURVEY 8d), intra, octree_bits %d (realised depth 12-13), colour JPEG snake Q85"
            % (args.points, args.kind, args.bits), "frames_per_step_per_gpu": frames, "points_per_frame": args.points,
            "cache": "inputs (%.0f MB per step per GPU) are larger than the 126 MB L2; no flush needed" % (frames * args.points * 32 / 1e6)}


def pcie_ceiling(torch, dist, world, dev, nbytes=1 << 30, reps=3):
    """What the box gives: pinned cudaMemcpyAsync H2D and D2H at the same time, on every rank at once (the ranks of a
    node share host memory and PCIe root complexes).  Returns per-rank GB/s (min over ranks)."""
    h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = [0.0, 0.0]
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s1):
            e[0].record(); d_a.copy_(h_a, non_blocking=True); e[1].record()
        with torch.cuda.stream(s2):
            e[2].record(); h_b.copy_(d_b, non_blocking=True); e[3].record()
        torch.cuda.synchronize()
        best = [max(best[0], nbytes / 1e6 / e[0].elapsed_time(e[1])), max(best[1], nbytes / 1e6 / e[2].elapsed_time(e[3]))]
    if world > 1:
        t = torch.tensor(best, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        best = [float(t[0]), float(t[1])]
    del h_a, h_b, d_a, d_b
    return {"h2d_gbs": best[0], "d2h_gbs": best[1], "how": "1 GiB pinned cudaMemcpyAsync each way at the same time, all %d ranks at once, best of %d, min over ranks" % (world, reps)}


def pcie_ceiling_same_buffers(torch, dist, world, dev, h_in, h_out, nframes, frame_bytes):
    """The ceiling for THIS step's traffic: every frame of the e2e leg's own pinned input buffer goes up and a frame-sized
    block comes down into its own pinned output buffer, one cudaMemcpyAsync per frame and direction like the library issues
    them, both directions at once, all ranks at once.  (A 100+ GB pinned working set does not move like a hot 1 GiB one.)"""
    src = torch.from_numpy(h_in.array); dst = torch.from_numpy(h_out.array)
    ring = 8
    d_up = torch.empty(ring * frame_bytes, dtype=torch.uint8, device=dev); d_dn = torch.empty(ring * frame_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s1):
        e[0].record()
        for i in range(nframes):
            d_up[(i % ring) * frame_bytes:(i % ring + 1) * frame_bytes].copy_(src[i * frame_bytes:(i + 1) * frame_bytes], non_blocking=True)
        e[1].record()
    with torch.cuda.stream(s2):
        e[2].record()
        for i in range(nframes):
            dst[i * frame_bytes:(i + 1) * frame_bytes].copy_(d_dn[(i % ring) * frame_bytes:(i % ring + 1) * frame_bytes], non_blocking=True)
        e[3].record()
    torch.cuda.synchronize()
    up, dn = nframes * frame_bytes / 1e6 / e[0].elapsed_time(e[1]), nframes * frame_bytes / 1e6 / e[2].elapsed_time(e[3])
    if world > 1:
        t = torch.tensor([up, dn], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        up, dn = float(t[0]), float(t[1])
    del d_up, d_dn
    return {"h2d_gbs": up, "d2h_gbs": dn, "pinned": bool(src.is_pinned()),
            "how": "the e2e leg's own pinned buffers (%d frames of %d bytes), one cudaMemcpyAsync per frame each way at the same time, all %d ranks at once, min over ranks" % (nframes, frame_bytes, world)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default="roundtrip", choices=["roundtrip", "decode", "tiles", "inter"],
                    help="roundtrip: the headline metric (encode+decode); decode: BASELINE configs[4], decode-only on oracle-produced streams resident in device memory; "
                         "tiles: BASELINE configs[3], one dense 4M-point frame at octree_bits 12 cut into root-octant tiles (sharded over the ranks, streams gathered on rank 0); "
                         "inter: BASELINE configs[2], a 30-frame group of 1M-point frames, every frame coded intra and as a P frame against its predecessor, frames sharded over the ranks, "
                         "the predictor clouds sent to the neighbour rank over NCCL")
    ap.add_argument("--gof-frames", type=int, default=30)
    ap.add_argument("--tile-bits", type=int, default=3, choices=[3, 6])
    ap.add_argument("--frames", type=int, default=int(os.environ.get("BENCH_FRAMES", "0")),
                    help="frames per step per GPU (0: as many as fit in device memory, at most 1024 -- the serial entropy stage is latency bound, so throughput grows with the frames in flight)")
    ap.add_argument("--points", type=int, default=1000000)
    ap.add_argument("--bits", type=int, default=11)
    ap.add_argument("--kind", default="surf", choices=["surf", "unif"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--value-api", default="roundtrip", choices=["roundtrip", "separate"],
                    help="device-resident step through pipelined ccv2_submit_roundtrip calls (default) or synchronous ccv2_encode_batch + ccv2_decode_batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cwi_pcl_codec_b200 import codec as K
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.mode == "decode":
        return run_decode_mode(args, torch, dist, K, rank, world, local, dev)
    if args.mode == "tiles":
        return run_tiles_mode(args, torch, dist, K, rank, world, local, dev)
    if args.mode == "inter":
        return run_inter_mode(args, torch, dist, K, rank, world, local, dev)
    NP = args.points
    F = args.frames
    cap = 4 * NP + (1 << 16)
    if F <= 0:       # the caller's side of a device-resident step: cloud in, cloud out, stream (32 + 32 + 4 B/point); the library's rings take what is left (~36 B/point per frame in flight)
        free_b, _total_b = torch.cuda.mem_get_info(local)
        F = int(max(8, min(1024, (0.55 * free_b) // (2 * NP * 32 + cap)))) // 8 * 8
    codec = K.Codec(K.default_params(octree_bits=args.bits), device=local)
    lib = K.load_library()

    # ---- host memory budget: the e2e leg needs pinned input + two sets of stream / cloud buffers on every rank of the node
    F_e2e = 0
    if not args.no_e2e:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        per_frame = NP * 32 * (1 + INFLIGHT_CALLS) + INFLIGHT_CALLS * cap
        F_e2e = int(max(8, min(F, (0.6 * avail / world) // per_frame)))
    h_in = K.PinnedBuffer(F_e2e * NP * 32) if F_e2e else None

    # ---- synthetic frames, generated in chunks and moved straight to the device (and to the pinned e2e input buffer);
    # only the first 16 stay on the host for the oracle checks / cpu_baseline
    # at most 256 distinct clouds per rank (generation is host-bound: ~0.1 s of numpy per frame); later frames of the step
    # reuse them in their own device / pinned buffers, so the step still moves and codes F separate 32 MB clouds
    U = min(F, 256)
    seeds = shard_frames(U, rank, world)
    d_in, frames = [], []
    for c0 in range(0, U, 32):
        chunk = gen_frames(args.kind, NP, seeds[c0:c0 + 32])
        for j, fr in enumerate(chunk):
            i = c0 + j
            flat = fr.view(np.uint8).reshape(-1)
            d_in.append(torch.from_numpy(flat).to(dev))
            if i < F_e2e:
                h_in.array[i * NP * 32:(i + 1) * NP * 32] = flat
            if i < 16:
                frames.append(fr)
        del chunk
    for i in range(U, F):
        d_in.append(d_in[i % U].clone())
        if i < F_e2e:
            h_in.array[i * NP * 32:(i + 1) * NP * 32] = h_in.array[(i % U) * NP * 32:(i % U + 1) * NP * 32]

    # ---- device-resident buffers (value) ----
    d_str = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(F)]
    d_out = [torch.empty(NP * 32, dtype=torch.uint8, device=dev) for _ in range(F)]
    in_ptrs = [t.data_ptr() for t in d_in]; str_ptrs = [t.data_ptr() for t in d_str]; out_ptrs = [t.data_ptr() for t in d_out]

    split = {"enc": 0.0, "dec": 0.0}

    def step_separate():                         # ccv2_encode_batch then ccv2_decode_batch (reported as the encode/decode split)
        lens = codec.encode_batch_raw(in_ptrs, [NP] * F, str_ptrs, [cap] * F)
        ms, launches = codec.last_device_ms, codec.last_launch_count
        ns = codec.decode_batch_raw(str_ptrs, lens, out_ptrs, [NP] * F)
        split["enc"] += ms; split["dec"] += codec.last_device_ms
        return launches + codec.last_launch_count, lens, ns

    def run_steps(k, submit):                    # k steps, at most two in flight; returns (launches, last lens, last counts)
        launches, pend, res = 0, [], None
        for _ in range(k):
            pend.append(submit())
            if len(pend) == INFLIGHT_CALLS:
                res = pend.pop(0).wait(); launches += codec.last_launch_count
        while pend:
            res = pend.pop(0).wait(); launches += codec.last_launch_count
        return launches, res[0], res[1]

    def device_steps(k):
        if args.value_api == "separate":
            launches = 0
            for _ in range(k):
                l, lens, ns = step_separate(); launches += l
            return launches, lens, ns
        return run_steps(k, lambda: codec.submit_roundtrip_raw(in_ptrs, [NP] * F, str_ptrs, [cap] * F, out_ptrs, [NP] * F))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    device_steps(args.warmup)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    codec.timer_start()
    t0 = time.perf_counter()
    launches, lens, ns = device_steps(args.steps)
    dev_ms = codec.timer_stop()                  # CUDA events on the library's control stream: first submit .. completion of the last call
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    t_dev = reduce_max(dev_ms / 1e3, dist if world > 1 else None, dev)
    t_wall = reduce_max(wall, dist if world > 1 else None, dev)
    value = world * F * NP * args.steps / t_dev / 1e6

    # ---- encode / decode split: two untimed synchronous steps through ccv2_encode_batch + ccv2_decode_batch
    split["enc"] = split["dec"] = 0.0
    split_steps = 2
    for _ in range(split_steps):
        step_separate()
    split_final = dict(split)
    codec.frame_id = 0
    _, lens, ns = device_steps(1)                # streams with frame ids 1..F again, for the oracle check below

    # ---- bit-exactness spot check against the oracle (outside the timed region, rank 0, first frame) ----
    bit_exact = None
    if rank == 0:
        from oracle import oracle as O
        ref, info = O.encode(frames[0], O.default_params(octree_bits=args.bits), frame_id=1)
        got = d_str[0][:lens[0]].cpu().numpy().tobytes()
        bit_exact = (got[:48] == ref[:48] and got[52:] == ref[52:])          # frame_ID (bytes 48..51) advances with every step
        rdec, _ = O.decode(ref)
        bit_exact = bool(bit_exact and np.array_equal(d_out[0][:ns[0] * 32].cpu().numpy().reshape(-1, 32), rdec))
    S = float(np.mean(lens)); V = float(np.mean(ns))
    alg_bytes_frame = 32 * NP + 32 * V + 2 * S                                # encode+decode, per frame

    # ---- single-frame latency: one cloud per call, device resident, best of 3 ----
    lat = {"encode": 1e30, "decode": 1e30}
    for _ in range(3):
        l1 = codec.encode_batch_raw(in_ptrs[:1], [NP], str_ptrs[:1], [cap])
        lat["encode"] = min(lat["encode"], codec.last_device_ms)
        codec.decode_batch_raw(str_ptrs[:1], l1, out_ptrs[:1], [NP])
        lat["decode"] = min(lat["decode"], codec.last_device_ms)

    # ---- e2e: the same steps with pinned host buffers ----
    e2e = None
    if F_e2e:
        FE = F_e2e
        del d_in, d_str, d_out, in_ptrs, str_ptrs, out_ptrs      # make room: the rings now also stage host inputs and outputs
        torch.cuda.empty_cache()
        ceiling_1g = pcie_ceiling(torch, dist, world, dev)
        # one pinned allocation per role, sliced per frame; two sets of result buffers (consecutive steps are in flight together)
        h_str = [K.PinnedBuffer(FE * cap) for _ in range(INFLIGHT_CALLS)]; h_out = [K.PinnedBuffer(FE * NP * 32) for _ in range(INFLIGHT_CALLS)]
        hi = [h_in.ptr + i * NP * 32 for i in range(FE)]
        hs = [[b.ptr + i * cap for i in range(FE)] for b in h_str]; ho = [[b.ptr + i * NP * 32 for i in range(FE)] for b in h_out]
        ceiling = pcie_ceiling_same_buffers(torch, dist, world, dev, h_in, h_out[0], FE, NP * 32)
        flip = [0]

        def submit_host():
            k = flip[0]; flip[0] = (k + 1) % INFLIGHT_CALLS
            return codec.submit_roundtrip_raw(hi, [NP] * FE, hs[k], [cap] * FE, ho[k], [NP] * FE)
        run_steps(2, submit_host)
        barrier()
        if os.environ.get("CCV2_TRACE"):                         # developer aid: the library's timelines then carry times since this point, across calls
            codec.timer_start()
        t0 = time.perf_counter()
        _, l2, n2 = run_steps(args.steps, submit_host)
        barrier()
        t_e2e = reduce_max(time.perf_counter() - t0, dist if world > 1 else None, dev)
        h2d_b, d2h_b = int(FE * NP * 32), int(sum(l2) + 32 * FE * NP)
        gbs = (h2d_b / 1e9 * args.steps / t_e2e, d2h_b / 1e9 * args.steps / t_e2e)
        e2e = {"value": world * FE * NP * args.steps / t_e2e / 1e6, "unit": "Mpoints/s", "frames_per_step_per_gpu": FE,
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "d2h_note": "decoded clouds come down as one capacity-sized transfer per frame (32 B x %d records, of which %.0f are voxels); streams by zero-copy stores at their exact size" % (NP, float(np.mean(n2))),
               "api": "ccv2_submit_roundtrip / ccv2_wait, %d calls in flight" % INFLIGHT_CALLS,
               "timing": "wall clock around K pipelined C-ABI calls, pinned host buffers in and out, max over ranks; the stream stays on the device between encoder and decoder (its 1.2 MB/frame re-upload is not part of the step)",
               "pcie_ceiling_gbs": ceiling, "pcie_hot_1gib_gbs": ceiling_1g, "achieved_h2d_gbs": gbs[0], "achieved_d2h_gbs": gbs[1],
               "frac_of_pcie": max(gbs[0] / ceiling["h2d_gbs"], gbs[1] / ceiling["d2h_gbs"])}
        for b in [h_in] + h_str + h_out:
            b.close()

    # ---- per-kernel profile (one extra, untimed, single-stream step over one group) -> roofline of the dominant kernel ----
    roofline = None
    if rank == 0 and not args.no_profile:
        roofline = profile_roofline(K, lib, args, frames, alg_bytes_frame, F, t_dev / args.steps)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nb = min(F, len(frames))
        t, p = cpu_oracle_run(frames[:nb], args.bits, 1)
        cpu_baseline = {"value": p / t / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
                        "sample": "%d of the step's %d frames, encode+decode, oracle port single thread (the reference's intra path is single-threaded)" % (nb, F)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps * 1e3, "wall_ms_per_step": t_wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/u32/f64 (integer codec, FP64 keys)", "data": "synthetic",
                "config": dict(workload_config(args, F), distinct_clouds_per_gpu=U), "bit_exact_vs_oracle": bit_exact, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "value_api": "ccv2_submit_roundtrip x K (%d calls in flight), ccv2_wait" % INFLIGHT_CALLS if args.value_api == "roundtrip" else "ccv2_encode_batch + ccv2_decode_batch",
                "encode_ms_per_step": split_final["enc"] / split_steps, "decode_ms_per_step": split_final["dec"] / split_steps,
                "encode_only_mpoints_s": F * NP * split_steps / max(split_final["enc"], 1e-9) / 1e3, "decode_only_mpoints_s": F * NP * split_steps / max(split_final["dec"], 1e-9) / 1e3,
                "single_frame_latency_ms": lat, "stream_bytes_per_frame": S, "voxels_per_frame": V}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_decode_mode(args, torch, dist, K, rank, world, local, dev):
    """BASELINE configs[4]: decode-only throughput on streams the ORACLE produced (seeds 0..63 of the config-2 inputs,
    sharded over the ranks), resident in device memory; decoded clouds stay in device memory."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    NP = args.points
    U = max(1, 64 // world)
    seeds = [rank * U + i for i in range(U)]
    op = O.default_params(octree_bits=args.bits)

    def make(s):
        fr = gen_frames(args.kind, NP, [s])[0]
        return O.encode(fr, op, frame_id=s + 1)[0]
    with ThreadPoolExecutor(max_workers=min(U, os.cpu_count() or 1)) as ex:
        streams = list(ex.map(make, seeds))
    F = args.frames
    if F <= 0:
        free_b, _ = torch.cuda.mem_get_info(local)
        F = int(max(8, min(2048, (0.4 * free_b) // (NP * 32 + max(len(s) for s in streams))))) // 8 * 8
    vox = [int.from_bytes(s[55:63], "little") for s in streams]
    d_str = [torch.from_numpy(np.frombuffer(streams[i % U], np.uint8).copy()).to(dev) for i in range(F)]
    d_out = [torch.empty(NP * 32, dtype=torch.uint8, device=dev) for _ in range(F)]
    sp = [t.data_ptr() for t in d_str]; sl = [len(streams[i % U]) for i in range(F)]; op_ = [t.data_ptr() for t in d_out]
    codec = K.Codec(K.default_params(octree_bits=args.bits), device=local)

    def run(k):
        pend, res, launches = [], None, 0
        for _ in range(k):
            pend.append(codec.submit_decode_raw(sp, sl, op_, [NP] * F))
            if len(pend) == INFLIGHT_CALLS:
                res = pend.pop(0).wait(); launches += codec.last_launch_count
        while pend:
            res = pend.pop(0).wait(); launches += codec.last_launch_count
        return res, launches

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    run(args.warmup)
    sampler = ClockSampler(local)
    barrier(); sampler.start(); codec.timer_start()
    ns, launches = run(args.steps)
    dev_ms = codec.timer_stop()
    barrier()
    clocks = sampler.stop()
    t_dev = reduce_max(dev_ms / 1e3, dist if world > 1 else None, dev)
    ok = None
    if rank == 0:
        rdec, _ = O.decode(streams[0])
        ok = bool(ns[0] == rdec.shape[0] and np.array_equal(d_out[0][:ns[0] * 32].cpu().numpy().reshape(-1, 32), rdec))
    lat = 1e30
    for _ in range(3):
        codec.decode_batch_raw(sp[:1], sl[:1], op_[:1], [NP]); lat = min(lat, codec.last_device_ms)
    if rank == 0:
        S, V = float(np.mean(sl)), float(np.mean([vox[i % U] for i in range(F)]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg = S + 32 * V                                                          # decode reads the stream, writes 32-B points
        line = {"metric": "Mpoints/s decode-only @ depth-11 intra, oracle-produced streams (BASELINE configs[4]); points = input points the streams represent",
                "value": world * F * NP * args.steps / t_dev / 1e6, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/u32/f64 (integer codec, FP64 positions)", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[4]: decode-only, streams of %d-point G-%s frames (seeds %d..%d on this rank) encoded by the CPU oracle, octree_bits %d, JPEG snake Q85; streams and decoded clouds resident in device memory"
                           % (NP, args.kind, seeds[0], seeds[-1], args.bits), "frames_per_step_per_gpu": F, "distinct_streams_per_gpu": U,
                           "cache": "streams + decoded clouds (%.0f MB per step per GPU) exceed the 126 MB L2" % (F * (S + 32 * V) / 1e6)},
                "decoded_bit_exact_vs_oracle": ok, "voxels_per_s_M": world * F * V * args.steps / t_dev / 1e6, "clocks": clocks, "gpu_launches": int(launches),
                "single_frame_latency_ms": lat, "stream_bytes_per_frame": S, "voxels_per_frame": V,
                "roofline": {"bound": "hbm", "achieved": alg * F / (t_dev / args.steps) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg * F / (t_dev / args.steps) / 1e9 / peak,
                             "traffic": None, "note": "step level: algorithmic decode bytes (S + 32 V) x frames per step / step time"},
                "api": "ccv2_submit_decode x K (%d calls in flight), ccv2_wait" % INFLIGHT_CALLS}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_tiles_mode(args, torch, dist, K, rank, world, local, dev):
    """BASELINE configs[3]: a dense 4M-point frame (G-surf), octree_bits 12, JPEG quality sweep 60..95, cut into root-octant
    tiles.  A step = one frame: every rank partitions the frame (device resident), encodes ITS tiles (tile t on rank
    t mod world), the streams are gathered on rank 0 (all_gather of sizes + point-to-point payloads over NCCL), and every
    rank decodes its own tiles again.  Strong scaling of ONE frame: what is reported is the frame's latency."""
    from cwi_pcl_codec_b200 import tiles as T
    NP = args.points if args.points != 1000000 else 4000000
    bits = args.bits if args.bits != 11 else 12
    tb = args.tile_bits
    nt = 1 << tb
    fr = gen_frames(args.kind, NP, [0])[0]
    d_in = torch.from_numpy(fr.view(np.uint8).reshape(-1)).to(dev)
    cap = 2 * NP + (1 << 20)
    mine = T.owned_tiles(tb, rank, world)
    probe = K.Codec(K.default_params(octree_bits=bits), device=local)
    _, tile_off = probe.split_tiles(fr, tb)                        # points per tile: sizes of this rank's stream and cloud buffers
    probe.close()
    tn = [tile_off[t + 1] - tile_off[t] for t in range(nt)]
    d_str = {t: torch.empty(4 * tn[t] + (1 << 18), dtype=torch.uint8, device=dev) for t in mine}
    d_out = {t: torch.empty(32 * max(tn[t], 1), dtype=torch.uint8, device=dev) for t in mine}
    results = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for q in (85, 60, 65, 70, 75, 80, 90, 95):
        codec = K.Codec(K.default_params(octree_bits=bits, jpeg_quality=q), device=local)
        outp = [d_str[t].data_ptr() if t in d_str else None for t in range(nt)]
        caps = [d_str[t].numel() if t in d_str else 0 for t in range(nt)]

        def step():
            lens, npts = codec.encode_tiles_raw(d_in.data_ptr(), NP, tb, outp, caps, rank, world)
            launches = codec.last_launch_count
            if world > 1:                                           # the exchange step: variable-length streams to the writer rank
                T.gather_tile_streams({t: d_str[t][:lens[t]].cpu().numpy().tobytes() for t in mine if lens[t]}, tb, dist, device=dev)
            live = [t for t in mine if lens[t]]
            ns = codec.decode_batch_raw([d_str[t].data_ptr() for t in live], [lens[t] for t in live], [d_out[t].data_ptr() for t in live], [d_out[t].numel() // 32 for t in live]) if live else []
            return lens, npts, sum(ns), launches + codec.last_launch_count
        steps = args.steps if q == 85 else 2
        for _ in range(args.warmup if q == 85 else 1):
            step()
        sampler = ClockSampler(local) if q == 85 else None
        barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        launches = 0
        for _ in range(steps):
            lens, npts, nvox, l = step(); launches += l
        barrier()
        wall = reduce_max(time.perf_counter() - t0, dist if world > 1 else None, dev)
        clocks = sampler.stop() if sampler else None
        tot = torch.tensor([float(sum(lens[t] for t in mine)), float(nvox)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        results[q] = {"ms_per_frame": wall / steps * 1e3, "stream_bytes": float(tot[0]), "voxels": float(tot[1]), "launches": launches, "clocks": clocks, "steps": steps}
        codec.close()
    # the same frame as ONE stream (no tiles), for the latency comparison (rank 0 only, Q85)
    plain = None
    if rank == 0:
        codec = K.Codec(K.default_params(octree_bits=bits, jpeg_quality=85), device=local)
        s1 = torch.empty(cap, dtype=torch.uint8, device=dev); o1 = torch.empty(NP * 32, dtype=torch.uint8, device=dev)
        best = [1e30, 1e30]
        for _ in range(3):
            l1 = codec.encode_batch_raw([d_in.data_ptr()], [NP], [s1.data_ptr()], [cap]); best[0] = min(best[0], codec.last_device_ms)
            codec.decode_batch_raw([s1.data_ptr()], l1, [o1.data_ptr()], [NP]); best[1] = min(best[1], codec.last_device_ms)
        plain = {"encode_ms": best[0], "decode_ms": best[1], "stream_bytes": l1[0]}
        codec.close()
    if rank == 0:
        r = results[85]
        line = {"metric": "Mpoints/s encode+decode of ONE dense frame in root-octant tile mode (BASELINE configs[3]); every tile stream bit-exact vs the reference encoder on that tile",
                "value": NP / (r["ms_per_frame"] / 1e3) / 1e6, "unit": "Mpoints/s", "n_gpus": world, "steps": r["steps"], "warmup": args.warmup,
                "ms_per_step": r["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32/f64 (integer codec, FP64 keys)", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[3]: one %d-point G-%s frame, octree_bits %d, JPEG snake Q85 (sweep 60..95 below), %d tiles (tile t on rank t mod %d), streams gathered on rank 0, every rank decodes its tiles"
                           % (NP, args.kind, bits, nt, world), "tile_bits": tb, "points_per_frame": NP,
                           "cache": "the frame (%.0f MB) exceeds the 126 MB L2" % (NP * 32 / 1e6)},
                "timing": "wall clock per frame (synchronous calls: partition + encode + gather + decode), max over ranks",
                "single_frame_latency_ms": {"tiled_encode_plus_decode": r["ms_per_frame"], "untiled_encode": plain["encode_ms"], "untiled_decode": plain["decode_ms"]},
                "speedup_vs_untiled": (plain["encode_ms"] + plain["decode_ms"]) / r["ms_per_frame"],
                "stream_bytes_tiled": r["stream_bytes"], "stream_bytes_untiled": plain["stream_bytes"], "voxels": r["voxels"],
                "jpeg_quality_sweep": {str(q): {"ms_per_frame": results[q]["ms_per_frame"], "stream_bytes": results[q]["stream_bytes"]} for q in sorted(results)},
                "clocks": r["clocks"], "gpu_launches": int(r["launches"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_inter_mode(args, torch, dist, K, rank, world, local, dev):
    """BASELINE configs[2]: a group of frames (30 x 1M points, synth.gen_gof), do_delta_coding as evaluate_compression runs it
    (eval.hpp:818-889): every frame is coded intra (encode + decode), and every frame but the first is ALSO coded as a P frame
    against the encoder's simplified cloud of its predecessor (encode + decode).  Frame f lives on rank f mod world; the
    predictor cloud of frame f travels to the owner of frame f + 1 over NCCL point to point (gof.exchange_predictors).
    Strong scaling of ONE group: value = points of the group / time of the step."""
    from cwi_pcl_codec_b200 import gof as G, synth
    NP, NF, bits = args.points, args.gof_frames, args.bits
    mine = G.owned_frames(NF, rank, world)
    clouds = synth.gen_gof(NP, seed=0, frames=NF)
    d_in = {f: torch.from_numpy(clouds[f].view(np.uint8).reshape(-1)).to(dev) for f in mine}
    first = clouds[0] if rank == 0 else None
    del clouds
    cap = 4 * NP + (1 << 16)
    codec = K.Codec(K.default_params(octree_bits=bits), device=local)
    d_str = {f: torch.empty(cap, dtype=torch.uint8, device=dev) for f in mine}
    d_dec = {f: torch.empty(NP * 32, dtype=torch.uint8, device=dev) for f in mine}
    d_oc = {f: torch.empty(NP * 32, dtype=torch.uint8, device=dev) for f in mine}
    d_is = {f: torch.empty(cap, dtype=torch.uint8, device=dev) for f in mine if f >= 1}
    d_ps = {f: torch.empty(30 * NP // 16 + (1 << 20), dtype=torch.uint8, device=dev) for f in mine if f >= 1}
    d_pdec = {f: torch.empty(2 * NP * 32, dtype=torch.uint8, device=dev) for f in mine if f >= 1}
    lib = K.load_library()
    import ctypes as C
    stats = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        launches = 0
        t = {}
        t0 = time.perf_counter()
        lens = codec.encode_batch_raw([d_in[f].data_ptr() for f in mine], [NP] * len(mine), [d_str[f].data_ptr() for f in mine], [cap] * len(mine)) if mine else []
        launches += codec.last_launch_count if mine else 0
        local_oc = {}
        for k, f in enumerate(mine):                             # [PCL] getOutputCloud of every I frame, straight into device memory
            n = C.c_size_t()
            codec._check(lib.ccv2_get_output_cloud(codec._h, k, d_oc[f].data_ptr(), NP, C.byref(n)))
            local_oc[f] = d_oc[f][:32 * n.value]
            launches += 1
        t["intra_encode"] = time.perf_counter() - t0; t0 = time.perf_counter()
        # the exchange comes BEFORE the asynchronous decode is submitted: NCCL's kernels would otherwise share a hardware queue with one of
        # the (by then ~60) streams of the codec and its child and sit behind a 0.18 s serial decode kernel (measured: 180 ms instead of 9)
        pred = G.exchange_predictors(local_oc, NF, dist if world > 1 else None, device=dev)
        torch.cuda.current_stream().synchronize()               # the received clouds are complete
        t["exchange"] = time.perf_counter() - t0; t0 = time.perf_counter()
        # the I frames' decode is submitted and collected at the end of the step: its serial entropy stage (latency bound, a few
        # dozen warps) runs while the P frames are predicted and coded
        pend = codec.submit_decode_raw([d_str[f].data_ptr() for f in mine], lens, [d_dec[f].data_ptr() for f in mine], [NP] * len(mine)) if mine else None
        t["intra_decode_submit"] = time.perf_counter() - t0; t0 = time.perf_counter()
        tot = {"i": 0, "p": 0, "mb": 0, "shared": 0, "conv": 0, "intra_pts": 0, "ppts": 0, "dec_pts": 0, "predict_ms": 0.0, "intra_ms": 0.0, "xbytes": 0}
        pf = [g for g in mine if g >= 1]
        td = 0.0
        if pf:                                                   # all P frames of this rank in one call each way: the intra parts run as one pipelined batch
            ics = [pred[g] for g in pf]
            nis = [ic.numel() // 32 for ic in ics]
            tot["xbytes"] = sum(ic.numel() for g, ic in zip(pf, ics) if G.owner(g - 1, world) != rank)
            ils, pls, infos = codec.encode_delta_batch_raw([ic.data_ptr() for ic in ics], nis, [d_in[g].data_ptr() for g in pf], [NP] * len(pf),
                                                           [d_is[g].data_ptr() for g in pf], [d_is[g].numel() for g in pf], [d_ps[g].data_ptr() for g in pf], [d_ps[g].numel() for g in pf])
            launches += codec.last_launch_count
            t1 = time.perf_counter()
            nds, nbs = codec.decode_delta_batch_raw([ic.data_ptr() for ic in ics], nis, [d_is[g].data_ptr() for g in pf], ils, [d_ps[g].data_ptr() for g in pf], pls,
                                                    [d_pdec[g].data_ptr() for g in pf], [d_pdec[g].numel() // 32 for g in pf])
            launches += codec.last_launch_count
            td = time.perf_counter() - t1
            for il, pl, info, n in zip(ils, pls, infos, nds):
                tot["i"] += il; tot["p"] += pl; tot["mb"] += info.macro_blocks; tot["shared"] += info.shared_blocks; tot["conv"] += info.converged_blocks
                tot["intra_pts"] += info.n_intra_points; tot["ppts"] += info.n_p_points; tot["dec_pts"] += n; tot["predict_ms"] += info.predict_ms; tot["intra_ms"] += info.intra_ms
        t["delta_decode"] = td; t["delta_encode"] = time.perf_counter() - t0 - td
        t0 = time.perf_counter()
        ns = pend.wait() if pend is not None else []
        launches += codec.last_launch_count if mine else 0
        t["intra_decode_wait"] = time.perf_counter() - t0
        tot["intra_bytes"] = int(sum(lens)); tot["intra_voxels"] = int(sum(ns))
        return launches, t, tot

    for _ in range(max(1, args.warmup)):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    launches = 0
    for _ in range(args.steps):
        l, tparts, tot = step(); launches += l
    barrier()
    wall = reduce_max(time.perf_counter() - t0, dist if world > 1 else None, dev)
    clocks = sampler.stop()
    keys = sorted(k for k in tot if k not in ("predict_ms", "intra_ms"))
    agg = torch.tensor([float(tot[k]) for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg)
    tot_all = dict(zip(keys, [float(v) for v in agg]))
    parts = torch.tensor([tparts[k] for k in sorted(tparts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(parts, op=dist.ReduceOp.MAX)
    # CPU side of the comparison (rank 0, one delta frame of the same group at reduced size would not be the same workload:
    # the oracle codes frame 1 against frame 0 at FULL size once; ~10-30 s)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cl = synth.gen_gof(NP, seed=0, frames=2)
        op = O.default_params(octree_bits=bits)
        t1 = time.perf_counter()
        ref, _, dbg = O.encode(cl[0], op, debug=True)
        rd, _ = O.decode(ref)
        ri, rp, rinfo = O.encode_delta(dbg["output_cloud"], cl[1], op)
        pdec, _ = O.decode_delta(dbg["output_cloud"], ri, rp, op)
        dt = time.perf_counter() - t1
        cpu = {"value": NP / dt / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
               "sample": "frames 0 and 1 of the same group: intra encode + decode of frame 0, P encode + decode of frame 1 against it, CPU oracle, 1 thread, %.1f s" % dt}
    if rank == 0:
        step_s = wall / args.steps
        line = {"metric": "Mpoints/s of a group of frames with inter-frame prediction (BASELINE configs[2]): every frame intra encode+decode, every frame but the first also P encode+decode against its predecessor; P and I streams bit-exact vs the oracle",
                "value": NF * NP / step_s / 1e6, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": step_s * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32/f32/f64 (integer codec, FP64 keys, FP32 ICP distances, FP64 alignment)", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[2]: GOF(0, %d) of %d-point frames (synth.gen_gof: 0.4 deg/frame rotation, 0.002/frame shift, 1 %% re-sampled), octree_bits %d, macroblock_size 16, "
                                       "icp_on_original 0; frame f on rank f mod %d, the simplified cloud of frame f sent to the owner of frame f+1 (NCCL point to point)" % (NF, NP, bits, world),
                           "points_per_frame": NP, "frames": NF, "cache": "each frame (32 MB) and its simplified cloud (30 MB) exceed what stays in L2 between uses; frames of a step are distinct"},
                "timing": "wall clock of the step (synchronous calls), max over ranks; device-resident inputs and outputs",
                "phase_ms_max_over_ranks": dict(zip(sorted(tparts), [float(v) * 1e3 for v in parts])),
                "prediction": {"macro_blocks": tot_all["mb"], "shared": tot_all["shared"], "predicted": tot_all["conv"], "points_coded_intra": tot_all["intra_pts"], "p_cloud_points": tot_all["ppts"],
                               "p_stream_bytes": tot_all["p"], "i_stream_bytes": tot_all["i"], "intra_only_bytes_all_frames": tot_all["intra_bytes"], "decoded_points": tot_all["dec_pts"],
                               "predictor_bytes_exchanged_per_step": tot_all["xbytes"],
                               "device_ms_per_p_frame_rank0": {"prediction_stage": tot["predict_ms"] / max(1, len([g for g in mine if g >= 1])), "intra_coder": tot["intra_ms"] / max(1, len([g for g in mine if g >= 1]))}},
                "cpu_baseline": cpu, "clocks": clocks, "gpu_launches": int(launches)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def profile_roofline(K, lib, args, frames, alg_bytes_frame, frames_per_step, step_s):
    """Runs one encode+decode of one group on a single stream with CUDA events around every kernel (inside the
    library, on the launching stream) and reports the dominant kernel against the measured HBM peak."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    which = "measured (MEASURED_PEAKS.json, copy kernel)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    prof = K.profile_step(frames[:min(len(frames), 16)], args.bits)
    if not prof:
        return None
    name, tot_ms, count, frames_per_launch = max(prof, key=lambda r: r[1])
    avg_s = tot_ms / count / 1e3
    achieved = alg_bytes_frame * frames_per_launch / avg_s / 1e9
    total_ms = sum(r[1] for r in prof)
    traffic = None
    try:                                                        # dram bytes per launch from the committed ncu --set full capture
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json" if os.path.exists(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")) else "ncu_traffic_r1.json")))
        traffic = int(t[name]) * frames_per_launch if name in t else None
    except Exception:
        pass
    # the HBM-bound parallel kernels against their OWN algorithmic bytes (per frame: N points, V voxels), for context:
    # the dominant kernel above is a serial dependency chain whose HBM fraction is tiny by construction
    N_, V_ = float(len(frames[0])), (alg_bytes_frame - 32.0 * len(frames[0])) / 32.0 * 0.96
    # sort: packed (code, colour) words, 16 B per point and pass; the library always launches 8 pass kernels of which
    # ceil((3 depth + 1) / 8) do work (the others return at once), so the per-launch average carries that factor
    depth_ = getattr(K.profile_step, "depth", 13)
    passes_ = (3 * depth_ + 1 + 7) // 8 if depth_ <= 13 else (3 * depth_ + 7) // 8
    own = {"keygen_kernel": 32 * N_ + 8 * N_, "sort_pass_kernel": (16 if depth_ <= 13 else 24) * N_ * passes_ / 8.0, "leaf_scan_kernel": 8 * N_ + 17 * V_,
           "dec_leaves_kernel": 32 * V_ + 13 * V_ / 2.9 + 1.6 * V_, "hist_kernel": 1.6 * V_ + 0.25 * V_}
    hbm_kernels = {}
    for n_, ms_, k_, fpl_ in prof:
        if n_ in own and ms_ > 0:
            gbs = own[n_] * fpl_ / (ms_ / k_ / 1e3) / 1e9
            hbm_kernels[n_] = {"own_algorithmic_bytes_per_frame": int(own[n_]), "avg_launch_ms": ms_ / k_, "achieved_GBs": gbs, "frac": gbs / peak}
    step_gbs = alg_bytes_frame * frames_per_step / step_s / 1e9
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "frac_step": step_gbs / peak, "achieved_step": step_gbs,
            "frac_note": "frac / achieved: the dominant kernel's own launch in the profiling step (frames_per_launch frames, one stream; a latency-bound serial kernel, so this only reflects how few frames one launch holds; calls of 256+ device-resident frames, like the timed steps, run the same stage as rc_decode_lps_kernel, 8 frames per warp). frac_step / achieved_step: algorithmic bytes x frames per step / measured ms_per_step -- the whole pipeline against the HBM peak",
            "hbm_bound_kernels_vs_own_bytes": hbm_kernels,
            "peak_source": which, "avg_launch_ms": avg_s * 1e3, "frames_per_launch": frames_per_launch,
            "algorithmic_bytes_per_frame": alg_bytes_frame, "share_of_step": tot_ms / total_ms,
            "kernels_ms": {r[0]: round(r[1], 4) for r in prof}}


if __name__ == "__main__":
    sys.exit(main())

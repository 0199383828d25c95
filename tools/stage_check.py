"""Developer tool (GPU box): stage-by-stage comparison of the CUDA path with the CPU oracle.

Runs a list of synthetic frames through ccv2_encode_batch, fetches the intermediates through the test hook and
diffs them against oracle/ccv2_oracle.c's; then decodes both the oracle's and the GPU's streams on the GPU and
diffs against the oracle's decoder.  Never stops at the first mismatch -- prints a table.
usage: python tools/stage_check.py [--big]
"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see csrc/ccv2_api.cu (stream -> hardware queue aliasing)
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K  # noqa: E402
from cwi_pcl_codec_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def first_diff(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    if d.size:
        i = int(d[0])
        return "first diff @%d: gpu=%s ref=%s (%d diffs)" % (i, a[i], b[i], d.size)
    if a.size != b.size:
        return "sizes differ gpu=%d ref=%d" % (a.size, b.size)
    return None


def cmp(name, got, ref, res):
    msg = first_diff(got, ref)
    ok = msg is None
    res.append((name, ok))
    print("    %-18s %s %s" % (name, "OK " if ok else "BAD", "" if ok else msg))
    return ok


def oparams(kp):
    return O.default_params(octree_resolution=kp.octree_resolution, point_resolution=kp.point_resolution,
                            do_color=kp.do_color_encoding, color_bit_resolution=kp.color_bit_resolution,
                            color_coding_type=kp.color_coding_type, do_centroid=kp.do_voxel_grid_centroid,
                            jpeg_quality=kp.jpeg_quality, create_scalable=kp.create_scalable_stream,
                            code_connectivity=kp.code_connectivity, macroblock_size=kp.macroblock_size,
                            do_icp_color_offset=kp.do_icp_color_offset)


def run_case(name, clouds, kp, results):
    print("== %s (%d frames)" % (name, len(clouds)))
    cdc = K.Codec(kp)
    t = time.time()
    try:
        streams = cdc.encode_batch(clouds)
    except K.Ccv2Error as e:
        print("   ENCODE FAILED:", e)
        results.append((name + ":encode", False))
        streams = None
    print("   encode wall %.1f ms, device %.2f ms, launches %d" % ((time.time() - t) * 1e3, cdc.last_device_ms, cdc.last_launch_count))
    op = oparams(kp)
    ref_streams = []
    next_id = 1
    for i, cl in enumerate(clouds):
        res = []
        data, info, dbg = O.encode(cl, op, frame_id=next_id, debug=True)
        if len(data):
            next_id += 1                     # the codec pre-increments frame_ID_ only for non-empty frames (impl.hpp:133)
        ref_streams.append(data)
        gi = cdc.debug_fetch(i, 5)
        print("  frame %d: n=%d ref depth=%d V=%d B=%d J=%d S=%d | gpu depth=%d V=%d B=%d J=%d err=0x%x" % (
            i, np.asarray(cl).nbytes // 32, info.depth, info.n_leaves, info.n_tree_bytes, info.n_color_bytes, len(data),
            gi.depth, gi.n_leaves, gi.n_tree_bytes, gi.n_color_bytes, gi.error))
        cmp("depth", [gi.depth], [info.depth], res)
        cmp("bbox", list(gi.bb_min) + list(gi.bb_max), list(info.bb_min) + list(info.bb_max), res)
        cmp("leaf_keys", cdc.debug_fetch(i, 0), dbg["leaf_keys"], res)
        cmp("tree_bytes", cdc.debug_fetch(i, 1), dbg["tree_bytes"], res)
        if kp.do_color_encoding:
            cmp("avg_colors", cdc.debug_fetch(i, 2), dbg["avg_colors"], res)
            cmp("color_payload", cdc.debug_fetch(i, 3), dbg["color_payload"], res)
        cmp("output_cloud", cdc.output_cloud(i).reshape(-1), dbg["output_cloud"].reshape(-1), res)
        if streams is not None:
            cmp("stream", np.frombuffer(streams[i], np.uint8), np.frombuffer(data, np.uint8), res)
        results.extend((name + ":f%d:" % i + k, v) for k, v in res)
    # decode: oracle streams on the GPU vs oracle decode
    nonempty = [s for s in ref_streams if len(s)]
    if nonempty:
        t = time.time()
        try:
            dec = cdc.decode_batch(nonempty)
            print("   decode wall %.1f ms, device %.2f ms" % ((time.time() - t) * 1e3, cdc.last_device_ms))
            for i, s in enumerate(nonempty):
                ref, _ = O.decode(s)
                res = []
                g = dec[i]
                cmp("dec_count", [g.shape[0]], [ref.shape[0]], res)
                m = min(g.shape[0], ref.shape[0])
                cmp("dec_xyz", g[:m, :16], ref[:m, :16], res)
                cmp("dec_rgba", g[:m, 16:20], ref[:m, 16:20], res)
                cmp("dec_pad", g[:m, 20:], ref[:m, 20:], res)
                results.extend((name + ":dec%d:" % i + k, v) for k, v in res)
        except K.Ccv2Error as e:
            print("   DECODE FAILED:", e)
            results.append((name + ":decode", False))
    cdc.close()


def main():
    big = "--big" in sys.argv
    results = []
    P = K.default_params
    cases = [
        ("surf10k_d8", [synth.gen_surface(10000, 0)], P(octree_bits=8)),
        ("unif2k_d6", [synth.gen_uniform(2000, 1)], P(octree_bits=6)),
        ("surf100k_d10_x3", [synth.gen_surface(100000, s) for s in (1, 2, 3)], P(octree_bits=10)),
        ("unif100k_d11", [synth.gen_uniform(100000, 4)], P(octree_bits=11)),
        ("tiny", [synth.gen_uniform(1, 5), synth.gen_uniform(2, 6), synth.gen_uniform(257, 7), synth.gen_uniform(513, 8)], P(octree_bits=7)),
        ("raw_type3", [synth.gen_surface(20000, 9)], P(octree_bits=9, color_coding_type=3)),
        ("pcl_type0_6bit", [synth.gen_surface(20000, 10)], P(octree_bits=9, color_coding_type=0, color_bits=6)),
        ("nocolor", [synth.gen_surface(20000, 11)], P(octree_bits=9, color_bits=0)),
        ("centroid", [synth.gen_surface(30000, 12)], P(octree_bits=7, keep_centroid=1)),
        ("q50", [synth.gen_surface(50000, 13)], P(octree_bits=10, jpeg_quality=50)),
        ("lines_type2", [synth.gen_surface(20000, 20), synth.gen_uniform(9000, 21)], P(octree_bits=9, color_coding_type=2)),
        ("lines_small", [synth.gen_surface(1500, 22), synth.gen_surface(2, 23), synth.gen_uniform(2049, 24), synth.gen_uniform(4097, 25)], P(octree_bits=10, color_coding_type=2, jpeg_quality=60)),
    ]
    # non-finite points and a late bbox violator (slow path)
    a = synth.gen_surface(40000, 14)
    a["x"][100] = np.nan
    a["y"][20000] = np.inf
    cases.append(("nonfinite", [a], P(octree_bits=9)))
    b = synth.gen_surface(60000, 15)
    b["x"][50000] = 7.5
    b["z"][55000] = -3.25
    cases.append(("late_violator", [b], P(octree_bits=9)))
    cases.append(("empty_mix", [np.zeros(0, synth.POINT_DTYPE), synth.gen_surface(5000, 16), np.full(3, np.nan, np.float32).repeat(8).view(synth.POINT_DTYPE)], P(octree_bits=8)))
    if big:
        cases.append(("surf1M_d11", [synth.gen_surface(1000000, 0)], P(octree_bits=11)))
        cases.append(("unif1M_d11", [synth.gen_uniform(1000000, 0)], P(octree_bits=11)))
        cases.append(("lines1M_d11", [synth.gen_surface(1000000, 1)], P(octree_bits=11, color_coding_type=2)))
    for name, clouds, kp in cases:
        try:
            run_case(name, clouds, kp, results)
        except Exception as e:  # keep going: this is a diagnosis tool
            import traceback
            traceback.print_exc()
            results.append((name + ":exception:" + type(e).__name__, False))
    bad = [k for k, v in results if not v]
    print("\nSUMMARY: %d checks, %d failed" % (len(results), len(bad)))
    for k in bad:
        print("  FAILED", k)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

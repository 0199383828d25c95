"""Generates the committed fixtures of the inter-frame path:

  delta_inputs.npz   -- two frozen frames of a synthetic group (synth.gen_gof(12000, seed=3, frames=2); xyz float32 + b,g,r
                        uint8), so the tests do not depend on numpy's generators
  delta_hashes.json  -- per configuration: SHA-256 of the ORACLE's I stream and P stream of frame 1 coded against the
                        simplified cloud of frame 0, the block statistics, and the SHA-256 of the decoded frame.
                        Like stream_hashes.json these do not pin the oracle against PCL (oracle/ccv2_oracle_inter.c explains
                        why the registration cannot be pinned bit for bit at all); they freeze today's behaviour so that the
                        oracle and the CUDA path cannot drift TOGETHER unnoticed.
usage: python tests/golden/make_delta_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cwi_pcl_codec_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

DELTA_CASES = {
    "b9": dict(octree_bits=9),
    "b9_centroid_offsets": dict(octree_bits=9, do_centroid=1, do_icp_color_offset=1),
    "b8_mb8": dict(octree_bits=8, macroblock_size=8),
    "res0.004_original": dict(octree_resolution=0.004, point_resolution=0.004, _icp_on_original=1),
}


def load_inputs(path=os.path.join(HERE, "delta_inputs.npz")):
    z = np.load(path)
    out = []
    for k in range(2):
        xyz, bgr = z["xyz%d" % k], z["bgr%d" % k]
        p = synth.pack_points(xyz, bgr[:, ::-1])
        out.append(np.ascontiguousarray(p).view(np.uint8).reshape(-1, 32))
    return out


def run_case(frames, kw):
    kw = dict(kw)
    orig = bool(kw.pop("_icp_on_original", 0))
    p = O.default_params(**kw)
    _, _, dbg = O.encode(frames[0], p, debug=True)
    i_s, p_s, info = O.encode_delta(dbg["output_cloud"], frames[1], p, icp_on_original=orig)
    dec, nb = O.decode_delta(dbg["output_cloud"], i_s, p_s, p)
    return {"i_sha256": hashlib.sha256(i_s).hexdigest(), "p_sha256": hashlib.sha256(p_s).hexdigest(), "i_len": len(i_s), "p_len": len(p_s),
            "macro_blocks": int(info.macro_blocks), "shared_blocks": int(info.shared_blocks), "converged_blocks": int(info.converged_blocks),
            "n_intra_points": int(info.n_intra_points), "decoded_points": int(dec.shape[0]), "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest()}


if __name__ == "__main__":
    g = synth.gen_gof(12000, seed=3, frames=2)
    np.savez_compressed(os.path.join(HERE, "delta_inputs.npz"),
                        **{"xyz%d" % k: np.stack([g[k]["x"], g[k]["y"], g[k]["z"]], 1) for k in range(2)},
                        **{"bgr%d" % k: np.stack([g[k]["b"], g[k]["g"], g[k]["r"]], 1) for k in range(2)})
    frames = load_inputs()
    assert all(np.array_equal(f, np.ascontiguousarray(c).view(np.uint8).reshape(-1, 32)) for f, c in zip(frames, g))
    out = {name: run_case(frames, kw) for name, kw in DELTA_CASES.items()}
    json.dump(out, open(os.path.join(HERE, "delta_hashes.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v["i_len"], v["p_len"], v["converged_blocks"], "/", v["shared_blocks"], "/", v["macro_blocks"])

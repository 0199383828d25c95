"""Generates the committed golden fixtures (run in the build container, which has Pillow with libjpeg-turbo):

  jpeg_vectors.npz   -- small RGB images, the JPEG bytes libjpeg-turbo (via Pillow) produces for them with exactly
                        the settings jpeg_io uses (jpeg_io.hpp:290-292: jpeg_set_defaults + jpeg_set_quality(q, TRUE),
                        JCS_RGB, 3 components), and the pixels libjpeg-turbo decodes from those bytes.
                        These PIN the oracle's JPEG encoder/decoder.
  stream_inputs.npz  -- the frozen INPUT clouds of every case of at most 50k points (xyz float32 + b,g,r uint8; the 32-byte
                        records are rebuilt by cases.load_case): the golden tests read these files, so they cannot skip
                        when numpy's generators change, and tools/export_golden_inputs.py turns them into the raw files
                        tools/real_pcl_dump.cpp feeds to the unmodified reference (tools/diff_against_real_pcl.md).
  stream_hashes.json -- SHA-256 of the ORACLE's compressed frames for frozen synthetic inputs.  These do NOT pin the
                        oracle against PCL (the reference ships no vectors, PCL is unavailable offline); they freeze
                        today's behaviour so that (a) the GPU path can be checked without running the oracle and
                        (b) whoever has PCL 1.10 can diff real reference output against the same inputs in one command.
usage: python tests/golden/make_golden.py
"""
import hashlib
import io
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cwi_pcl_codec_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

JPEG_CASES = [  # (h, w, quality, smooth)
    (16, 256, 85, 0), (37, 256, 85, 0), (37, 256, 60, 1), (1, 2048, 85, 1), (1, 777, 95, 0), (9, 256, 85, 1),
    (24, 256, 75, 1), (1, 776, 85, 0), (5, 264, 75, 0), (1, 9, 85, 0), (1, 1, 85, 0), (3, 2047, 85, 0),
    (40, 256, 1, 1), (17, 256, 100, 1), (8, 8, 50, 0), (33, 40, 85, 1), (1, 3, 85, 0), (2, 4, 85, 0), (1, 5, 85, 0)]

STREAM_CASES = [  # (name, generator, n, seed, oracle params)
    ("surf10k_b8_snake85", "gen_surface", 10000, 0, dict(octree_bits=8)),
    ("unif20k_b9_snake85", "gen_uniform", 20000, 1, dict(octree_bits=9)),
    ("surf50k_b10_q60", "gen_surface", 50000, 2, dict(octree_bits=10, jpeg_quality=60)),
    ("surf20k_b9_raw", "gen_surface", 20000, 3, dict(octree_bits=9, color_coding_type=3)),
    ("surf20k_b9_pcl6", "gen_surface", 20000, 4, dict(octree_bits=9, color_coding_type=0, color_bit_resolution=6)),
    ("surf20k_b9_nocolor", "gen_surface", 20000, 5, dict(octree_bits=9, do_color=0, color_bit_resolution=0)),
    ("surf30k_b7_centroid", "gen_surface", 30000, 6, dict(octree_bits=7, do_centroid=1)),
    ("surf30k_res0p01_growth", "late_growth", 30000, 2, dict(octree_resolution=0.01, point_resolution=0.01)),       # non-power-of-two resolution: PCL's key order matters
    ("surf30k_res0p003_centroid_pcl6", "late_growth", 30000, 1, dict(octree_resolution=0.003, point_resolution=0.003, do_centroid=1, color_coding_type=0, color_bit_resolution=6)),
    ("surf30k_b9_q0", "gen_surface", 30000, 27, dict(octree_bits=9, jpeg_quality=0)),                                # the CLI's default quality (eval.hpp:161)
    ("surf20k_b7_detail_pcl", "gen_surface", 20000, 7, dict(octree_bits=7, enh_bits=3, do_voxel_grid=0, color_coding_type=0, color_bit_resolution=8)),   # the class default: detail mode
    ("surf20k_b7_detail_snake_nocentroid", "gen_surface", 20000, 8, dict(octree_bits=7, enh_bits=2, do_voxel_grid=0)),                                # detail mode, JPEG averages (decode is undefined in the reference, App. C-7)
    ("surf1M_b11_snake85", "gen_surface", 1000000, 0, dict(octree_bits=11)),
    ("unif1M_b11_snake85", "gen_uniform", 1000000, 0, dict(octree_bits=11)),
]


def late_growth(n, seed):
    """G-surf with two points far outside, beyond index 16384: bounding-box growth late in the cloud."""
    b = synth.gen_surface(n, seed)
    b["x"][25000] = 3.5
    b["z"][28000] = -2.25
    return b


def image(rng, h, w, smooth):
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([128 + 100 * np.sin(x / 17 + y / 5), 128 + 90 * np.cos(x / 11), 128 + 60 * np.sin(y / 3 + x / 29)], -1)
    return np.clip(a + rng.integers(-6, 7, a.shape), 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    for k, (h, w, q, sm) in enumerate(JPEG_CASES):
        img = image(rng, h, w, sm)
        b = io.BytesIO()
        Image.fromarray(img, "RGB").save(b, "JPEG", quality=q)
        data = np.frombuffer(b.getvalue(), np.uint8)
        dec = np.array(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))
        out["img%d" % k] = img
        out["jpg%d" % k] = data
        out["dec%d" % k] = dec
        out["q%d" % k] = np.int32(q)
    np.savez_compressed(os.path.join(HERE, "jpeg_vectors.npz"), **out)
    hashes = {}
    frozen = {}
    for name, gen, n, seed, kw in STREAM_CASES:
        pts = late_growth(n, seed) if gen == "late_growth" else getattr(synth, gen)(n, seed)
        if n <= 50000:
            frozen[name + ".xyz"] = np.stack([pts["x"], pts["y"], pts["z"]], 1)
            frozen[name + ".bgr"] = np.stack([pts["b"], pts["g"], pts["r"]], 1)
        data, info = O.encode(pts, O.default_params(**kw), frame_id=1)
        try:
            dec, _ = O.decode(data)
        except RuntimeError:                                         # detail mode with JPEG colour: the reference's decoder is undefined there
            dec = np.zeros((0, 32), np.uint8)
        hashes[name] = dict(gen=gen, n=n, seed=seed, params=kw, stream_bytes=len(data), stream_sha256=hashlib.sha256(data).hexdigest(),
                            input_sha256=hashlib.sha256(pts.tobytes()).hexdigest(), decoded_sha256=hashlib.sha256(dec.tobytes()).hexdigest(),
                            depth=int(info.depth), leaves=int(info.n_leaves), tree_bytes=int(info.n_tree_bytes), color_bytes=int(info.n_color_bytes))
        print(name, len(data), hashes[name]["stream_sha256"][:16])
    json.dump(hashes, open(os.path.join(HERE, "stream_hashes.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "stream_inputs.npz"), **frozen)


if __name__ == "__main__":
    main()

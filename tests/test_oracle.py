"""CPU tests of the oracle (oracle/ccv2_oracle.c): pinned against libjpeg-turbo golden vectors where the
reference's arithmetic lives in libjpeg, self-consistency elsewhere (the reference ships no tests -- SURVEY 4)."""
import hashlib
import io
import json
import os
import sys

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth


def _vectors(golden_dir):
    z = np.load(os.path.join(golden_dir, "jpeg_vectors.npz"))
    k = 0
    while "img%d" % k in z:
        yield z["img%d" % k], z["jpg%d" % k], z["dec%d" % k], int(z["q%d" % k])
        k += 1


def test_jpeg_encode_matches_libjpeg_turbo_golden(oracle, golden_dir):
    n = 0
    for img, jpg, _, q in _vectors(golden_dir):
        got = oracle.jpeg_encode(img, q)
        assert got.size == jpg.size and np.array_equal(got, jpg), "shape %s q=%d" % (img.shape, q)
        n += 1
    assert n >= 15


def test_jpeg_decode_matches_libjpeg_turbo_golden(oracle, golden_dir):
    for img, jpg, dec, q in _vectors(golden_dir):
        got = oracle.jpeg_decode(jpg)
        assert got.shape == dec.shape and np.array_equal(got, dec), "shape %s q=%d" % (img.shape, q)


def test_jpeg_live_pillow(oracle):
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    for h, w, q in [(13, 256, 85), (64, 256, 30), (1, 300, 92), (2, 2, 85)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b = io.BytesIO()
        PIL.fromarray(img, "RGB").save(b, "JPEG", quality=q)
        assert oracle.jpeg_encode(img, q).tobytes() == b.getvalue()
        assert np.array_equal(oracle.jpeg_decode(b.getvalue()), np.array(PIL.open(io.BytesIO(b.getvalue())).convert("RGB")))


@pytest.mark.parametrize("n,alphabet", [(0, 1), (1, 1), (1000, 4), (70000, 256), (200000, 256), (300000, 3), (66000, 1)])
def test_range_coder_roundtrip_and_exact_consumption(oracle, n, alphabet):
    rng = np.random.default_rng(n + alphabet)
    data = rng.integers(0, alphabet, n, dtype=np.uint8) if alphabet < 256 else (rng.normal(128, 20, n).clip(0, 255)).astype(np.uint8)
    enc = oracle.range_encode(data)
    assert enc.size >= 1028 + 4                       # table + 4 flush bytes, even for n == 0
    table = enc[:1028].view("<u4")
    assert table[0] == 0 and np.all(np.diff(table.astype(np.int64)) >= 1) and table[256] < (1 << 16)
    tail = np.concatenate([enc, np.arange(17, dtype=np.uint8)])       # decoder must stop exactly at the encoder's end
    dec, used = oracle.range_decode(tail, n)
    assert used == enc.size
    assert np.array_equal(dec, data)


def test_range_coder_known_answer_empty(oracle):
    enc = oracle.range_encode(np.zeros(0, np.uint8))
    assert enc.size == 1032
    assert np.array_equal(enc[:1028].view("<u4"), np.arange(257, dtype=np.uint32))   # "+1 if empty" rule
    assert np.array_equal(enc[1028:], np.zeros(4, np.uint8))


@pytest.mark.parametrize("w,h", [(256, 1), (256, 8), (256, 13), (256, 16), (256, 37), (64, 5), (8, 3)])
def test_snake_closed_form_equals_literal_iterator(oracle, w, h):
    lit = oracle.snake_literal(w, h)
    assert np.array_equal(np.sort(lit), np.arange(w * h))            # a permutation
    assert np.array_equal(oracle.snake_closed(w, h), lit)


@pytest.mark.parametrize("d,n", [(1, 3), (3, 40), (5, 500), (7, 3000), (11, 2000)])
def test_sort_based_serialisation_equals_recursive_dfs(oracle, d, n):
    rng = np.random.default_rng(d * 1000 + n)
    codes = np.unique(rng.integers(0, 1 << (3 * d), n, dtype=np.uint64))
    pts = np.zeros(codes.size, synth.POINT_DTYPE)
    # place one point at the centre of every chosen voxel of a unit-resolution grid, first point pinned so that
    # the bbox growth is deterministic; then compare the encoder's tree bytes with the recursive DFS over its leaves
    k = np.zeros((codes.size, 3))
    for l in range(d):
        c = (codes >> np.uint64(3 * (d - 1 - l))) & np.uint64(7)
        k[:, 0] = k[:, 0] * 2 + (c >> np.uint64(2)).astype(float)
        k[:, 1] = k[:, 1] * 2 + ((c >> np.uint64(1)) & np.uint64(1)).astype(float)
        k[:, 2] = k[:, 2] * 2 + (c & np.uint64(1)).astype(float)
    pts["x"], pts["y"], pts["z"] = k[:, 0] + 0.5, k[:, 1] + 0.5, k[:, 2] + 0.5
    p = oracle.default_params(octree_resolution=1.0, point_resolution=1.0, color_coding_type=3)
    _, info, dbg = oracle.encode(pts, p, debug=True)
    assert info.n_leaves == codes.size
    assert np.array_equal(oracle.dfs_recursive(dbg["leaf_keys"], info.depth), dbg["tree_bytes"])


def test_bbox_growth_properties(oracle):
    for seed in range(4):
        pts = synth.gen_uniform(20000, seed)
        res = 2.0 ** -11
        bmin, bmax, depth, keys, fin = oracle.bbox_keys(pts, res)
        assert 12 <= depth <= 15                                     # 11 bits -> depth 12-14 typically
        assert fin.all() and int(keys.max()) < (1 << depth)
        x0 = np.array([pts["x"][0], pts["y"][0], pts["z"][0]], np.float64)
        # first point centres the grid: min = p0 - res - k * res * 2^j  => (p0 - min)/res is an odd-ish integer + 1 exactly
        t = (x0 - bmin) / res
        assert np.allclose(t, np.round(t)) and np.all(bmax - bmin == (1 << depth) * res - 1.1920928955078125e-07)
        # keys recomputed from the final box equal the sequentially maintained ones
        xyz = np.stack([pts["x"], pts["y"], pts["z"]], 1).astype(np.float64)
        assert np.array_equal(((xyz - bmin) / res).astype(np.uint32), keys)


def test_nonfinite_points_are_skipped_and_empty_cloud_writes_nothing(oracle):
    pts = synth.gen_surface(5000, 1)
    pts["x"][17] = np.nan
    pts["z"][99] = np.inf
    data, info = oracle.encode(pts, oracle.default_params(octree_bits=8))
    assert info.n_finite == 4998
    dec, _ = oracle.decode(data)
    assert dec.shape[0] == info.n_leaves
    empty, _ = oracle.encode(np.zeros(0, synth.POINT_DTYPE), oracle.default_params())
    assert empty == b""
    allnan = np.zeros(4, synth.POINT_DTYPE)
    allnan["x"] = np.nan
    assert oracle.encode(allnan, oracle.default_params())[0] == b""


@pytest.mark.parametrize("kw", [dict(octree_bits=8), dict(octree_bits=9, color_coding_type=3), dict(octree_bits=9, color_coding_type=0, color_bit_resolution=5),
                                dict(octree_bits=8, do_color=0), dict(octree_bits=7, do_centroid=1), dict(octree_bits=9, color_coding_type=2)])
def test_roundtrip_geometry_and_header(oracle, kw):
    pts = synth.gen_surface(15000, 5)
    p = oracle.default_params(**kw)
    data, info, dbg = oracle.encode(pts, p, frame_id=3, debug=True)
    assert data[:48] == b"<PCL-OCT-CODECV2-COMPRESSED><PCL-OCT-COMPRESSED>"
    assert int.from_bytes(data[48:52], "little") == 3 and data[52] == 1 and data[53] == 1
    assert int.from_bytes(data[55:63], "little") == info.n_leaves
    assert int.from_bytes(data[140:148], "little") == info.n_tree_bytes
    dec, dinfo = oracle.decode(data)
    assert dec.shape[0] == info.n_leaves and dinfo.depth == info.depth
    xyz = dec[:, :12].copy().view(np.float32).reshape(-1, 3)
    d, keys = info.depth, dbg["leaf_keys"]
    k = np.zeros((keys.size, 3), np.uint64)
    for l in range(d):
        c = (keys >> np.uint64(3 * (d - 1 - l))) & np.uint64(7)
        k[:, 0] = (k[:, 0] << np.uint64(1)) | (c >> np.uint64(2))
        k[:, 1] = (k[:, 1] << np.uint64(1)) | ((c >> np.uint64(1)) & np.uint64(1))
        k[:, 2] = (k[:, 2] << np.uint64(1)) | (c & np.uint64(1))
    off = 0.0 if kw.get("do_centroid") else 0.5
    exp = ((k.astype(np.float64) + off) * p.octree_resolution + np.array(info.bb_min))
    if kw.get("do_centroid"):
        exp = exp + dbg["centroid_bytes"].reshape(-1, 3).astype(np.float32) * np.float32(0.001)
    assert np.array_equal(exp.astype(np.float32), xyz)
    rgb = dec[:, 16:19]
    if kw.get("do_color", 1) == 0:
        assert np.all(rgb == 255) and np.all(dec[:, 19] == 0)
    elif kw.get("color_coding_type", 1) == 3:
        assert np.array_equal(rgb, dbg["avg_colors"].reshape(-1, 3))
    elif kw.get("color_coding_type", 1) == 0:
        red = 8 - kw["color_bit_resolution"]
        assert np.array_equal(rgb, (dbg["avg_colors"].reshape(-1, 3).astype(np.uint16) << red).astype(np.uint8))
    else:
        err = rgb.astype(np.float64) - dbg["avg_colors"].reshape(-1, 3)
        psnr = 10 * np.log10(255.0 ** 2 / (err ** 2).mean())
        assert psnr > 14.0                                           # lossy JPEG of a Morton-ordered colour strip


def test_frozen_stream_hashes(oracle, golden_dir):
    """The committed SHA-256 of the oracle's streams for the FROZEN input files (tests/golden/stream_inputs.npz): no
    generator is involved, so the test cannot skip."""
    sys.path.insert(0, golden_dir)
    import cases
    table = json.load(open(os.path.join(golden_dir, "stream_hashes.json")))
    names = cases.frozen_names()
    assert len(names) >= 12 and all(n in table for n in names)
    for name in names:
        e = table[name]
        pts = cases.load_case(name)
        assert hashlib.sha256(pts.tobytes()).hexdigest() == e["input_sha256"], name
        data, _ = oracle.encode(pts, oracle.default_params(**e["params"]), frame_id=1)
        assert hashlib.sha256(data).hexdigest() == e["stream_sha256"], name
        if name == "surf20k_b7_detail_snake_nocentroid":
            with pytest.raises(RuntimeError):                        # detail mode + JPEG colour: undefined in the reference's decoder (App. C-7)
                oracle.decode(data)
            continue
        dec, _ = oracle.decode(data)
        assert hashlib.sha256(dec.tobytes()).hexdigest() == e["decoded_sha256"], name


def test_int_vector_range_coder_round_trip(oracle):
    """[PCL] StaticRangeCoder::encodeIntVectorToStream / decodeStreamToIntVector (detail mode's point counts, impl.hpp:1738)."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(5)
    for n, hi in ((1, 2), (1000, 3), (50000, 40), (20000, 70000), (300, 1)):
        v = rng.integers(0 if hi > 1 else 1, hi + 1, n).astype(np.uint32) if hi > 1 else np.ones(n, np.uint32)
        out = C.POINTER(C.c_uint8)(); ol = C.c_size_t()
        L.orc_range_encode_int.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
        L.orc_range_decode_int.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        assert L.orc_range_encode_int(v.ctypes.data, n, C.byref(out), C.byref(ol)) == 0
        coded = np.ctypeslib.as_array(out, shape=(ol.value,)).copy()
        L.orc_free(out)
        tsize = int.from_bytes(coded[:8].tobytes(), "little")
        assert tsize - 1 >= int(v.max()) + 1 and (tsize - 1) & (tsize - 2) == 0      # power-of-two table that covers max symbol + 1
        back = np.zeros(n, np.uint32); used = C.c_size_t()
        padded = np.concatenate([coded, np.full(16, 0xAB, np.uint8)])
        assert L.orc_range_decode_int(padded.ctypes.data, padded.size, back.ctypes.data, n, C.byref(used)) == 0
        assert np.array_equal(back, v) and used.value == coded.size                  # exact byte consumption: the next size word follows


def test_detail_mode_round_trip(oracle):
    """doVoxelGridDownDownSampling=false (the class default, codec.h:108-143): every point survives, quantised to
    point_resolution relative to its voxel corner (impl.hpp:1525-1541, 1592-1613); colours are exact for type 0 at 8 bits."""
    cl = synth.gen_surface(20000, 77)
    cl["x"][5] = np.nan
    p = oracle.default_params(octree_bits=7, enh_bits=3, do_voxel_grid=0, color_coding_type=0, color_bit_resolution=8)
    data, info = oracle.encode(cl, p, frame_id=1)
    assert data[53] == 0 and int.from_bytes(data[55:63], "little") == 19999           # header: detail mode, object count
    dec, _ = oracle.decode(data)
    assert dec.shape[0] == 19999
    fin = np.isfinite(cl["x"])
    src = np.stack([cl["x"], cl["y"], cl["z"]], 1)[fin].astype(np.float64)
    xyz = dec[:, :12].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
    res, pres = 2.0 ** -7, 2.0 ** -10
    bmin = np.frombuffer(data[80:104], "<f8")
    # decoded points are in voxel (Morton) order; match through (voxel key, quantised residual) multisets
    def sig(a):
        k = np.floor((a - bmin) / res + 1e-9).astype(np.int64)
        q = np.floor((a - (k * res + bmin)) / pres + 1e-6).astype(np.int64)
        return np.sort((k * 8 + np.clip(q, 0, 7)).view([("a", "i8"), ("b", "i8"), ("c", "i8")]).reshape(-1), order=["a", "b", "c"])
    assert np.array_equal(sig(src), sig(xyz))
    assert np.abs(np.sort(src[:, 0]) - np.sort(xyz[:, 0])).max() < 2 * pres
    # colour multiset is preserved exactly (XOR diffs against the voxel average, no bit reduction)
    csrc = np.sort((cl["b"][fin].astype(np.int64) << 16) | (cl["g"][fin].astype(np.int64) << 8) | cl["r"][fin])
    cdec = np.sort((dec[:, 16].astype(np.int64) << 16) | (dec[:, 17].astype(np.int64) << 8) | dec[:, 18])
    assert np.array_equal(csrc, cdec)
    # 6-bit colours: two low bits lost; and the JPEG-average + diff combination encodes but is undefined on decode
    p6 = oracle.default_params(octree_bits=7, enh_bits=3, do_voxel_grid=0, color_coding_type=0, color_bit_resolution=6)
    d6, _ = oracle.decode(oracle.encode(cl, p6, frame_id=1)[0])
    assert d6.shape[0] == 19999 and set(np.unique(d6[:, 16:19] & 3).tolist()) <= {0, 1, 2, 3}
    pj = oracle.default_params(octree_bits=7, enh_bits=3, do_voxel_grid=0, color_coding_type=1)
    sj, _ = oracle.encode(cl, pj, frame_id=1)
    with pytest.raises(RuntimeError):
        oracle.decode(sj)


def test_output_cloud_is_the_simplified_cloud_of_the_encoder():
    """impl.hpp:1549-1576 / [PCL] getOutputCloud (eval.hpp:862): one point per voxel in stream order, voxel centre (or the
    float centroid), average colour before JPEG, alpha 255."""
    from oracle import oracle as O
    from cwi_pcl_codec_b200 import synth
    cl = synth.gen_surface(20000, 3)
    for cen in (0, 1):
        data, info, dbg = O.encode(cl, O.default_params(octree_bits=8, do_centroid=cen, color_coding_type=3), frame_id=1, debug=True)
        oc = dbg["output_cloud"]
        dec, _ = O.decode(data)
        assert oc.shape == (info.n_leaves, 32) and dec.shape == oc.shape
        xyz_o = oc[:, :12].copy().view(np.float32).reshape(-1, 3)
        xyz_d = dec[:, :12].copy().view(np.float32).reshape(-1, 3)
        if cen == 0:
            assert np.array_equal(xyz_o, xyz_d)                     # same voxel centres as the decoder reconstructs
        else:
            assert np.abs(xyz_o - xyz_d).max() <= 0.001 + 1e-6      # the centroid coder quantises to 1 mm (pcv2.h:83-118)
        # colour type 3 ships the raw averages: the decoder's colours are the encoder's averages
        assert np.array_equal(oc[:, 16:19], dec[:, 16:19])
        assert np.array_equal(oc[:, 16:19].reshape(-1), dbg["avg_colors"])
        assert set(oc[:, 19].tolist()) == {255}
        assert np.all(oc[:, 12:16].copy().view(np.float32) == 1.0)


def test_quality_metrics_match_an_independent_kdtree(oracle):
    """computeQualityMetric (quality_metrics_impl.hpp:82-239): the oracle's exhaustive search against scipy's kd-tree."""
    scipy_spatial = pytest.importorskip("scipy.spatial")
    cl = synth.gen_surface(6000, 1)
    dec, _ = oracle.decode(oracle.encode(cl, oracle.default_params(octree_bits=7), frame_id=1)[0])
    q = oracle.quality_metrics(cl, dec)
    a = np.stack([cl["x"], cl["y"], cl["z"]], 1).astype(np.float64)
    b = dec[:, :12].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
    da, ia = scipy_spatial.cKDTree(b).query(a)
    db, _ = scipy_spatial.cKDTree(a).query(b)
    assert q.in_point_count == 6000 and q.out_point_count == dec.shape[0]
    assert abs(q.left_rms - np.sqrt((da ** 2).mean())) < 1e-7 and abs(q.right_rms - np.sqrt((db ** 2).mean())) < 1e-7
    assert abs(q.symm_hausdorff - max(da.max(), db.max())) < 1e-7
    rms = max(np.sqrt((da ** 2).mean()), np.sqrt((db ** 2).mean()))
    assert abs(q.psnr_db - 10 * np.log10((a.max(0) ** 2).sum() / rms ** 2)) < 1e-3
    def yuv(r, g, bl):
        return np.stack([0.299 * r + 0.587 * g + 0.114 * bl, -0.147 * r - 0.289 * g + 0.436 * bl, 0.615 * r - 0.515 * g - 0.100 * bl], 1) / 255.0
    ya = yuv(cl["r"].astype(np.float64), cl["g"].astype(np.float64), cl["b"].astype(np.float64))
    yb = yuv(dec[ia, 18].astype(np.float64), dec[ia, 17].astype(np.float64), dec[ia, 16].astype(np.float64))
    psnr = 10 * np.log10(1.0 / ((ya - yb) ** 2).mean(0))
    assert np.abs(np.array(list(q.psnr_yuv)) - psnr).max() < 1e-3

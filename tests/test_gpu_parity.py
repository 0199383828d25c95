"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libccv2.so), against the CPU oracle on the
same seeded inputs -- bit-exact streams, bit-exact decoded clouds -- plus the committed golden hashes, the
edge cases the format has (empty / non-finite / single-point / late bbox growth), batches spanning several
stream groups, device-pointer I/O and error behaviour."""
import hashlib
import json
import os

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from cwi_pcl_codec_b200 import codec
    codec.load_library()
    return codec


def oparams(O, kp):
    return O.default_params(octree_resolution=kp.octree_resolution, point_resolution=kp.point_resolution,
                            do_color=kp.do_color_encoding, color_bit_resolution=kp.color_bit_resolution,
                            color_coding_type=kp.color_coding_type, do_centroid=kp.do_voxel_grid_centroid,
                            jpeg_quality=kp.jpeg_quality, do_voxel_grid=kp.do_voxel_grid_downsampling)


def check_batch(K, O, clouds, kp):
    c = K.Codec(kp)
    try:
        streams = c.encode_batch(clouds)
        op = oparams(O, kp)
        fid = 0
        refs = []
        for i, cl in enumerate(clouds):
            ref, info = O.encode(cl, op, frame_id=fid + 1)
            if ref:
                fid += 1
            refs.append(ref)
            assert streams[i] == ref, "frame %d: GPU stream (%d B) != oracle stream (%d B)" % (i, len(streams[i]), len(ref))
        assert c.frame_id == fid
        live = [s for s in refs if s]
        if live:
            dec = c.decode_batch(live)
            for s, d in zip(live, dec):
                rd, _ = O.decode(s)
                assert d.shape == rd.shape and np.array_equal(d, rd)
            m = c.metrics()
            assert m[0] > 1028
        return streams
    finally:
        c.close()


CASES = {
    "surf10k_b8": (lambda: [synth.gen_surface(10000, 0)], dict(octree_bits=8)),                      # BASELINE configs[0] shape
    "unif2k_b6": (lambda: [synth.gen_uniform(2000, 1)], dict(octree_bits=6)),
    "surf100k_b10_x3": (lambda: [synth.gen_surface(100000, s) for s in (1, 2, 3)], dict(octree_bits=10)),
    "unif100k_b11": (lambda: [synth.gen_uniform(100000, 4)], dict(octree_bits=11)),
    "tiny": (lambda: [synth.gen_uniform(1, 5), synth.gen_uniform(2, 6), synth.gen_uniform(255, 7), synth.gen_uniform(256, 17), synth.gen_uniform(257, 7), synth.gen_uniform(513, 8)], dict(octree_bits=7)),
    "raw_type3": (lambda: [synth.gen_surface(20000, 9)], dict(octree_bits=9, color_coding_type=3)),
    "pcl_type0_6bit": (lambda: [synth.gen_surface(20000, 10)], dict(octree_bits=9, color_coding_type=0, color_bits=6)),
    "nocolor": (lambda: [synth.gen_surface(20000, 11)], dict(octree_bits=9, color_bits=0)),
    "centroid": (lambda: [synth.gen_surface(30000, 12)], dict(octree_bits=7, keep_centroid=1)),
    "q50": (lambda: [synth.gen_surface(50000, 13)], dict(octree_bits=10, jpeg_quality=50)),
    "q100_random_colour": (lambda: [synth.gen_uniform(30000, 14)], dict(octree_bits=9, jpeg_quality=100)),
    "lines_type2": (lambda: [synth.gen_surface(20000, 20), synth.gen_uniform(9000, 21)], dict(octree_bits=9, color_coding_type=2)),
    "lines_type2_edge_widths": (lambda: [synth.gen_surface(1500, 22), synth.gen_surface(2, 23), synth.gen_uniform(2049, 24), synth.gen_uniform(4097, 25), synth.gen_uniform(6143, 26)],
                                dict(octree_bits=10, color_coding_type=2, jpeg_quality=60)),
    "q0_cli_default": (lambda: [synth.gen_surface(30000, 27)], dict(octree_bits=9, jpeg_quality=0)),          # eval.hpp:161: the CLI default, libjpeg clamps to 1
    "duplicates_one_voxel": (lambda: [np.repeat(synth.gen_surface(7, 15), 3000)], dict(octree_bits=5)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_stream_and_decode_bit_exact(K, oracle, name):
    make, kw = CASES[name]
    check_batch(K, oracle, make(), K.default_params(**kw))


def test_nonfinite_points_and_late_bbox_growth(K, oracle):
    a = synth.gen_surface(40000, 14)
    a["x"][100] = np.nan
    a["y"][20000] = np.inf
    b = synth.gen_surface(60000, 15)      # violators far beyond the exact 16k prefix: exercises the slow sequential path
    b["x"][50000] = 7.5
    b["z"][55000] = -3.25
    lead_nan = synth.gen_surface(30000, 16)
    lead_nan["x"][:20000] = np.nan        # no finite point inside the prefix at all
    check_batch(K, oracle, [a, b, lead_nan], K.default_params(octree_bits=9))


def _late_growth_cloud(seed):
    b = synth.gen_surface(30000, seed)
    b["x"][25000] = 3.5                    # beyond the exact 16k prefix: slow sequential path, two more growth steps
    b["z"][28000] = -2.25
    return b


@pytest.mark.parametrize("res,seed", [(0.01, 2), (0.003, 1), (0.007, 0), (0.01, 3)])
def test_non_power_of_two_resolution_follows_pcl_key_order(K, oracle, res, seed):
    """octreeResolution that is not a power of two (the class default is 0.01, codec.h:108-143): PCL keys a point against
    the box in force when it was added and re-roots on growth; recomputing from the final box gives a different key for
    the first point of these clouds (checked below with the oracle), so only the sequential order is bit-exact."""
    cl = _late_growth_cloud(seed)
    bmin, _, _, keys, _ = oracle.bbox_keys(cl, res)
    xyz = np.stack([cl["x"], cl["y"], cl["z"]], 1).astype(np.float64)
    differs = int((((xyz - bmin) / res).astype(np.uint32) != keys).any(1).sum())
    if seed != 3:
        assert differs > 0                 # the case is a real discriminator
    kp = K.default_params(octree_resolution=res, point_resolution=res)
    check_batch(K, oracle, [cl, synth.gen_uniform(5000, seed + 40)], kp)
    kp = K.default_params(octree_resolution=res, point_resolution=res, keep_centroid=1, color_coding_type=0, color_bits=6)
    check_batch(K, oracle, [cl], kp)


def test_junk_before_the_magic_and_ff_quirk(K, oracle):
    """syncToHeader (impl.hpp:1660-1676) scans for the magic: leading junk is skipped, unless it holds a 0xFF byte
    (`readChar == EOF` on a char), which aborts the sync."""
    cl = synth.gen_surface(12000, 28)
    kp = K.default_params(octree_bits=9)
    c = K.Codec(kp)
    s = c.encode_batch([cl])[0]
    ref = oracle.decode(s)[0]
    junk = bytes([1, 2, 3, ord("<"), ord("P"), 0, ord("<")]) * 5
    out = np.zeros((ref.shape[0], 32), np.uint8)
    a = np.frombuffer(junk + s, np.uint8)
    n = c.decode_batch_raw([a.ctypes.data], [a.size], [out.ctypes.data], [out.shape[0]])
    assert n == [ref.shape[0]] and np.array_equal(out, ref)
    assert np.array_equal(oracle.decode(junk + s)[0], ref)
    b = np.frombuffer(b"\x01\xff\x02" + s, np.uint8)
    with pytest.raises(K.Ccv2Error) as ei:
        c.decode_batch_raw([b.ctypes.data], [b.size], [out.ctypes.data], [out.shape[0]])
    assert ei.value.status == -6
    with pytest.raises(RuntimeError):
        oracle.decode(b"\x01\xff\x02" + s)
    c.close()


def test_empty_and_all_nonfinite_frames_write_nothing_and_keep_frame_ids(K, oracle):
    nanf = np.zeros(3, synth.POINT_DTYPE)
    nanf["y"] = np.nan
    clouds = [np.zeros(0, synth.POINT_DTYPE), synth.gen_surface(5000, 16), nanf, synth.gen_surface(3000, 17)]
    streams = check_batch(K, oracle, clouds, K.default_params(octree_bits=8))
    assert streams[0] == b"" and streams[2] == b""
    assert int.from_bytes(streams[1][48:52], "little") == 1 and int.from_bytes(streams[3][48:52], "little") == 2


def test_batch_spanning_groups_and_streams_matches_one_by_one(K, oracle):
    clouds = [synth.gen_surface(3000 + 500 * i, 100 + i) for i in range(37)]   # > group size * streams mix
    kp = K.default_params(octree_bits=8)
    streams = check_batch(K, oracle, clouds, kp)
    c = K.Codec(kp)
    one = [c.encode_batch([cl])[0] for cl in clouds]
    c.close()
    assert one == streams


def kparams_from_golden(K, prm):
    """oracle parameter names of stream_hashes.json -> ccv2_params."""
    kw = dict(prm)
    m = {"do_centroid": "do_voxel_grid_centroid", "do_voxel_grid": "do_voxel_grid_downsampling", "do_color": "do_color_encoding",
         "color_bit_resolution": "color_bit_resolution"}
    out = {}
    if "octree_bits" in kw:
        out["octree_bits"] = kw.pop("octree_bits"); out["enh_bits"] = kw.pop("enh_bits", 0)
    for k, v in kw.items():
        out[m.get(k, k)] = v
    return K.default_params(**out)


def test_golden_stream_hashes_frozen_inputs(K, golden_dir):
    """GPU streams and decoded clouds against the committed SHA-256 for the frozen input FILES (no generator involved)."""
    import sys
    sys.path.insert(0, golden_dir)
    import cases
    table = json.load(open(os.path.join(golden_dir, "stream_hashes.json")))
    checked = 0
    for name in cases.frozen_names():
        e = table[name]
        c = K.Codec(kparams_from_golden(K, e["params"]))
        s = c.encode_batch([cases.load_case(name)])[0]
        assert len(s) == e["stream_bytes"] and hashlib.sha256(s).hexdigest() == e["stream_sha256"], name
        if name == "surf20k_b7_detail_snake_nocentroid":           # detail mode + JPEG colour: the reference's decoder is undefined there (App. C-7)
            with pytest.raises(K.Ccv2Error) as ei:
                c.decode_batch([s])
            assert ei.value.status == -3
        else:
            d = c.decode_batch([s])[0]
            assert hashlib.sha256(d.tobytes()).hexdigest() == e["decoded_sha256"], name
        c.close()
        checked += 1
    assert checked >= 12


def test_golden_stream_hashes_full_size(K, golden_dir):
    """BASELINE configs[1] size (1M points, depth 11, Q85): GPU streams against the committed SHA-256."""
    table = json.load(open(os.path.join(golden_dir, "stream_hashes.json")))
    c = K.Codec(K.default_params(octree_bits=11))
    checked = 0
    for name in ("surf1M_b11_snake85", "unif1M_b11_snake85"):
        e = table[name]
        pts = getattr(synth, e["gen"])(e["n"], e["seed"])
        if hashlib.sha256(pts.tobytes()).hexdigest() != e["input_sha256"]:
            continue
        c.frame_id = 0
        s = c.encode_batch([pts])[0]
        assert len(s) == e["stream_bytes"] and hashlib.sha256(s).hexdigest() == e["stream_sha256"], name
        d = c.decode_batch([s])[0]
        assert hashlib.sha256(d.tobytes()).hexdigest() == e["decoded_sha256"], name
        checked += 1
    c.close()
    if not checked:
        pytest.skip("synthetic generator differs on this machine")


def test_full_size_against_oracle_and_roundtrip_properties(K, oracle):
    pts = synth.gen_surface(1000000, 21)
    kp = K.default_params(octree_bits=11)
    s = check_batch(K, oracle, [pts], kp)[0]
    c = K.Codec(kp)
    d = c.decode_batch([s])[0]
    # size-independent properties: count == header point_count; positions are voxel centres of the bbox grid; sorted in
    # Morton (DFS) order; every input point lies in exactly one decoded voxel
    assert d.shape[0] == int.from_bytes(s[55:63], "little")
    bmin = np.frombuffer(s[80:104], "<f8")
    xyz = d[:, :12].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
    k = (xyz - bmin) * 2048.0 - 0.5
    assert np.abs(k - np.round(k)).max() < 2e-3                    # positions are float32: voxel centres up to rounding
    k = np.round(k)
    kin = np.floor((np.stack([pts["x"], pts["y"], pts["z"]], 1).astype(np.float64) - bmin) * 2048.0).astype(np.int64)
    assert np.array_equal(np.unique(kin, axis=0), np.unique(k.astype(np.int64), axis=0))
    c.close()


def test_device_pointer_io(K, oracle):
    torch = pytest.importorskip("torch")
    clouds = [synth.gen_surface(50000, 30 + i) for i in range(3)]
    kp = K.default_params(octree_bits=10)
    c = K.Codec(kp)
    d_in = [torch.from_numpy(cl.view(np.uint8).reshape(-1)).cuda() for cl in clouds]
    cap = 6 * 50000 + 65536
    d_str = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in clouds]
    lens = c.encode_batch_raw([t.data_ptr() for t in d_in], [50000] * 3, [t.data_ptr() for t in d_str], [cap] * 3)
    d_out = [torch.empty(50000 * 32, dtype=torch.uint8, device="cuda") for _ in clouds]
    ns = c.decode_batch_raw([t.data_ptr() for t in d_str], lens, [t.data_ptr() for t in d_out], [50000] * 3)
    op = oparams(oracle, kp)
    for i, cl in enumerate(clouds):
        ref, _ = oracle.encode(cl, op, frame_id=i + 1)
        assert d_str[i][:lens[i]].cpu().numpy().tobytes() == ref
        rd, _ = oracle.decode(ref)
        assert np.array_equal(d_out[i][:ns[i] * 32].cpu().numpy().reshape(-1, 32), rd)
    assert c.last_launch_count > 0
    c.close()


def test_error_behaviour(K, oracle):
    c = K.Codec(K.default_params(octree_bits=8))
    cl = synth.gen_surface(5000, 40)
    a = np.ascontiguousarray(cl)
    small = np.zeros(100, np.uint8)
    with pytest.raises(K.Ccv2Error) as ei:
        c.encode_batch_raw([a.ctypes.data], [5000], [small.ctypes.data], [100])
    assert ei.value.status == -4                                   # CCV2_ERR_CAPACITY
    s = c.encode_batch([cl])[0]
    junk = bytes(300)
    out = np.zeros((10, 32), np.uint8)
    jb = np.frombuffer(junk, np.uint8)
    with pytest.raises(K.Ccv2Error) as ei:
        c.decode_batch_raw([jb.ctypes.data], [300], [out.ctypes.data], [10])
    assert ei.value.status == -6                                   # CCV2_ERR_STREAM (no frame header)
    trunc = np.frombuffer(s[:len(s) // 2], np.uint8)
    big = np.zeros((5000, 32), np.uint8)
    with pytest.raises(K.Ccv2Error):
        c.decode_batch_raw([trunc.ctypes.data], [trunc.size], [big.ctypes.data], [5000])
    sb = np.frombuffer(s, np.uint8)
    with pytest.raises(K.Ccv2Error) as ei:
        c.decode_batch_raw([sb.ctypes.data], [sb.size], [out.ctypes.data], [10])
    assert ei.value.status == -4                                   # caller's point buffer too small
    # the codec is still usable afterwards
    assert c.decode_batch([s])[0].shape[0] == int.from_bytes(s[55:63], "little")
    c.close()


def test_reference_facade(K, oracle):
    cdc = K.OctreePointCloudCodecV2(K.MANUAL_CONFIGURATION, False, 2.0 ** -9, 2.0 ** -9, True, 0, True, 8, 1, False, False, False, 85, 1)
    cl = synth.gen_surface(8000, 50)
    s = cdc.encodePointCloud(cl)
    ref, _ = oracle.encode(cl, oracle.default_params(octree_bits=9), frame_id=1)
    assert s == ref
    assert np.array_equal(cdc.decodePointCloud(s), oracle.decode(ref)[0])
    assert cdc.decodePointCloud(b"garbage" * 20).shape[0] == 0
    assert cdc.getPerformanceMetrics()[0] > 0


@pytest.mark.parametrize("env", [{"CCV2_LPS_DEC": "1"}, {"CCV2_LPS_ENC": "0", "CCV2_LPS_DEC": "0"}, {"CCV2_PACKED": "0"}, {"CCV2_GROUP": "3", "CCV2_STREAMS": "2", "CCV2_INFLIGHT": "6"}])
def test_both_entropy_stage_implementations_are_bit_exact(K, oracle, env, monkeypatch):
    """The lane-per-stream coders and the CTA-per-frame ones are interchangeable (the library picks by call type; the
    knobs are read when a handle is created): same streams, same clouds, whichever is forced.  Likewise the packed
    (code, colour) sort elements against (code, index) pairs, and a tiny ring (2 long-lived sets of 3 frames on 2 streams:
    every set is handed on several times inside one call)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    clouds = [synth.gen_surface(40000, 70 + i) for i in range(11)] + [np.zeros(0, synth.POINT_DTYPE), synth.gen_uniform(3000, 90)]
    for kw in ({"octree_bits": 9}, {"octree_bits": 8, "keep_centroid": True}, {"octree_bits": 9, "color_coding_type": 2}):
        check_batch(K, oracle, clouds, K.default_params(**kw))


@pytest.mark.parametrize("centroid", [False, True])
def test_output_cloud_matches_oracle(K, oracle, centroid):
    """[PCL] getOutputCloud() after encodePointCloud (eval.hpp:862): the encoder's simplified cloud, impl.hpp:1549-1576."""
    clouds = [synth.gen_surface(30000, 60), synth.gen_uniform(5000, 61), np.zeros(0, synth.POINT_DTYPE)]
    kp = K.default_params(octree_bits=9, keep_centroid=centroid)
    c = K.Codec(kp)
    streams = c.encode_batch(clouds)
    op = oracle.default_params(octree_bits=9, do_centroid=int(centroid))
    for i, cl in enumerate(clouds):
        got = c.output_cloud(i)
        if np.asarray(cl).size == 0:
            assert got.shape[0] == 0
            continue
        _, _, dbg = oracle.encode(cl, op, frame_id=1, debug=True)
        assert np.array_equal(got, dbg["output_cloud"])
    c.decode_batch([streams[0]])                                   # any later call invalidates it
    with pytest.raises(K.Ccv2Error) as ei:
        c.output_cloud(0)
    assert ei.value.status == -3
    c.close()
    cdc = K.OctreePointCloudCodecV2(K.MANUAL_CONFIGURATION, False, 2.0 ** -9, 2.0 ** -9, True, 0, True, 8, 1, False, False, False, 85, 1)
    cdc.encodePointCloud(clouds[0])
    assert cdc.getOutputCloud().shape[0] == c_leaves(oracle, clouds[0])


def c_leaves(oracle, cl):
    return oracle.encode(cl, oracle.default_params(octree_bits=9), frame_id=1, debug=True)[1].n_leaves


def test_roundtrip_call_matches_separate_calls(K, oracle):
    clouds = [synth.gen_surface(20000 + 1000 * i, 60 + i) for i in range(11)] + [np.zeros(0, synth.POINT_DTYPE)]
    kp = K.default_params(octree_bits=9)
    c = K.Codec(kp)
    arrs = [np.ascontiguousarray(cl) for cl in clouds]
    ns = [a.nbytes // 32 for a in arrs]
    caps = [6 * n + 65536 for n in ns]
    strs = [np.zeros(cp, np.uint8) for cp in caps]
    outs = [np.zeros((max(1, n), 32), np.uint8) for n in ns]
    lens, cnts = c.roundtrip_batch_raw([a.ctypes.data if a.size else None for a in arrs], ns, [s.ctypes.data for s in strs], caps,
                                       [o.ctypes.data for o in outs], [max(1, n) for n in ns])
    op = oparams(oracle, kp)
    fid = 0
    for i, cl in enumerate(clouds):
        ref, _ = oracle.encode(cl, op, frame_id=fid + 1)
        if ref:
            fid += 1
        assert strs[i][:lens[i]].tobytes() == ref
        if ref:
            rd, _ = oracle.decode(ref)
            assert cnts[i] == rd.shape[0] and np.array_equal(outs[i][:cnts[i]], rd)
        else:
            assert cnts[i] == 0 and lens[i] == 0
    # streams optional: sizes are still reported
    c.frame_id = 0
    lens2, cnts2 = c.roundtrip_batch_raw([a.ctypes.data if a.size else None for a in arrs], ns, None, None,
                                         [o.ctypes.data for o in outs], [max(1, n) for n in ns])
    assert lens2 == lens and cnts2 == cnts
    c.close()


@pytest.mark.parametrize("quality", [60, 75, 95])
def test_config4_shape_4m_points_depth12(K, oracle, quality):
    """BASELINE configs[3] shape on one GPU: dense 4M-point frame, octree_bits 12, points of the JPEG quality sweep 60..95."""
    pts = synth.gen_surface(4000000, 0)
    check_batch(K, oracle, [pts], K.default_params(octree_bits=12, jpeg_quality=quality))


def test_lines_mode_stream_as_the_reference_writes_it_for_a_small_cloud(K, oracle):
    """color_coding_type 2 with fewer than 2048 voxels: the reference leaves the line 2048 pixels wide and reads past its
    buffer (cjpeg.h:255-274), so ITS stream carries one 2048 x 1 image whose first V pixels are the voxel colours; oracle
    and CUDA encoder write a V x 1 image instead (documented divergence).  A decoder must still read the reference's
    form: build it (same tree layer, colour layer = one 2048-wide line with arbitrary pixels after the first V)."""
    cl = synth.gen_surface(1500, 22)
    op = oracle.default_params(octree_bits=10, color_coding_type=2, jpeg_quality=85)
    ref, info, dbg = oracle.encode(cl, op, frame_id=1, debug=True)
    V = info.n_leaves
    assert V < 2048
    rng = np.random.default_rng(3)
    line = np.concatenate([dbg["avg_colors"].reshape(-1, 3), rng.integers(0, 256, (2048 - V, 3), dtype=np.uint8)])[None, :, :]
    jpg = oracle.jpeg_encode(line, 85).tobytes()
    payload = (1).to_bytes(4, "little") + len(jpg).to_bytes(4, "little") + jpg
    s = ref[:148 + info.coded[0]] + len(payload).to_bytes(8, "little") + oracle.range_encode(np.frombuffer(payload, np.uint8)).tobytes()
    rd, _ = oracle.decode(s)
    assert rd.shape[0] == V
    c = K.Codec(K.default_params(octree_bits=10, color_coding_type=2))
    d = c.decode_batch([s])[0]
    assert np.array_equal(d, rd)
    # and the V x 1 form both encoders write
    assert np.array_equal(c.decode_batch([ref])[0], oracle.decode(ref)[0])
    c.close()


def test_many_frames_in_one_call_all_groups_and_streams(K, oracle):
    """more frames than streams x 1, mixed sizes: every stream gets a multi-frame group; spot-check against the oracle."""
    clouds = [synth.gen_surface(40000 + 997 * (i % 7), 200 + i) for i in range(70)]
    kp = K.default_params(octree_bits=9)
    c = K.Codec(kp)
    streams = c.encode_batch(clouds)
    dec = c.decode_batch(streams)
    op = oparams(oracle, kp)
    for i in (0, 8, 9, 33, 69):
        ref, _ = oracle.encode(clouds[i], op, frame_id=i + 1)
        assert streams[i] == ref
        assert np.array_equal(dec[i], oracle.decode(ref)[0])
    assert c.frame_id == 70
    c.close()


def test_pinned_host_buffers_and_async_calls_share_the_rings(K, oracle):
    """The fast path of the C ABI: pinned host memory in and out (copy engines up and down, zero-copy stream export) and
    two calls in flight on one handle (ccv2_submit_* / ccv2_wait) -- same bytes as the synchronous, pageable calls."""
    clouds = [synth.gen_surface(30000 + 700 * (i % 5), 300 + i) for i in range(40)] + [np.zeros(0, synth.POINT_DTYPE)]
    kp = K.default_params(octree_bits=9)
    c = K.Codec(kp)
    ref_streams = c.encode_batch(clouds)
    ref_dec = c.decode_batch([s for s in ref_streams if s])
    ns = [cl.shape[0] for cl in clouds]
    cap = 6 * max(ns) + 65536
    F = len(clouds)
    h_in = K.PinnedBuffer(sum(ns) * 32 + 32)
    offs = np.concatenate([[0], np.cumsum(ns)]) * 32
    for i, cl in enumerate(clouds):
        h_in.array[offs[i]:offs[i + 1]] = np.ascontiguousarray(cl).view(np.uint8).reshape(-1)
    bufs = []
    pend = []
    c.frame_id = 0
    for rep in range(3):                                   # three calls back to back, at most two in flight
        h_str, h_out = K.PinnedBuffer(F * cap), K.PinnedBuffer(F * max(ns) * 32)
        bufs.append((h_str, h_out))
        pend.append(c.submit_roundtrip_raw([h_in.ptr + int(offs[i]) if ns[i] else None for i in range(F)], ns,
                                           [h_str.ptr + i * cap for i in range(F)], [cap] * F,
                                           [h_out.ptr + i * max(ns) * 32 for i in range(F)], [max(n, 1) for n in ns]))
    live = 0
    for rep, p in enumerate(pend):
        lens, cnts = p.wait()
        h_str, h_out = bufs[rep]
        j = 0
        for i in range(F):
            got = h_str.array[i * cap:i * cap + lens[i]].tobytes()
            if not ref_streams[i]:
                assert lens[i] == 0 and cnts[i] == 0
                continue
            assert got[:48] == ref_streams[i][:48] and got[52:] == ref_streams[i][52:]        # frame ids run on over the three calls
            assert int.from_bytes(got[48:52], "little") == int.from_bytes(ref_streams[i][48:52], "little") + rep * 40
            d = h_out.array[i * max(ns) * 32:i * max(ns) * 32 + cnts[i] * 32].reshape(-1, 32)
            assert np.array_equal(d, ref_dec[j])
            j += 1
            live += 1
    assert live == 120
    # decode-only from pinned streams into pinned clouds, asynchronously, while an encode of other frames is in flight
    h_str, h_out = bufs[0]
    lens = [len(s) for s in ref_streams]
    for i, s in enumerate(ref_streams):
        h_str.array[i * cap:i * cap + len(s)] = np.frombuffer(s, np.uint8)
    h_out.array[:] = 0
    h_str2 = K.PinnedBuffer(F * cap)
    pe = c.submit_encode_raw([h_in.ptr + int(offs[i]) if ns[i] else None for i in range(F)], ns, [h_str2.ptr + i * cap for i in range(F)], [cap] * F)
    pd = c.submit_decode_raw([h_str.ptr + i * cap for i in range(F - 1)], lens[:-1], [h_out.ptr + i * max(ns) * 32 for i in range(F - 1)], [int.from_bytes(s[55:63], "little") for s in ref_streams[:-1]])
    cnts = pd.wait()
    lens2 = pe.wait()
    for i in range(F - 1):
        assert np.array_equal(h_out.array[i * max(ns) * 32:i * max(ns) * 32 + cnts[i] * 32].reshape(-1, 32), ref_dec[i])
        assert h_str2.array[i * cap + 52:i * cap + lens2[i]].tobytes() == ref_streams[i][52:]
    c.close()


def test_sparse_deep_octree_overflows_the_default_workspace_and_is_retried(K, oracle):
    """Sparse clouds in a deep octree have up to depth x V tree bytes (here ~10 per point against a default bound of 4):
    the frame is flagged on the device and encoded again with bounds that cannot overflow; decoding sizes its workspace
    from the stream's own header, so reference-produced streams of such clouds decode as well."""
    clouds = [synth.gen_uniform(300000, 31), synth.gen_surface(20000, 32), synth.gen_uniform(60000, 33)]
    kp = K.default_params(octree_bits=16)
    streams = check_batch(K, oracle, clouds, kp)
    assert int.from_bytes(streams[0][140:148], "little") > 4 * 300000 + 1024          # B beyond the default tree bound
    # the same through a round trip (device-side link between encoder and decoder) and with device-resident streams
    c = K.Codec(kp)
    arrs = [np.ascontiguousarray(cl) for cl in clouds]
    caps = [len(s) + 64 for s in streams]
    strs = [np.zeros(cp, np.uint8) for cp in caps]
    outs = [np.zeros((a.shape[0], 32), np.uint8) for a in arrs]
    lens, cnts = c.roundtrip_batch_raw([a.ctypes.data for a in arrs], [a.shape[0] for a in arrs], [s.ctypes.data for s in strs], caps,
                                       [o.ctypes.data for o in outs], [a.shape[0] for a in arrs])
    for i in range(3):
        assert strs[i][:lens[i]].tobytes() == streams[i]
        assert np.array_equal(outs[i][:cnts[i]], oracle.decode(streams[i])[0])
    torch = pytest.importorskip("torch")
    d_str = [torch.from_numpy(np.frombuffer(s, np.uint8).copy()).cuda() for s in streams]
    d_out = [torch.empty(a.shape[0] * 32, dtype=torch.uint8, device="cuda") for a in arrs]
    ns = c.decode_batch_raw([t.data_ptr() for t in d_str], [len(s) for s in streams], [t.data_ptr() for t in d_out], [a.shape[0] for a in arrs])
    for i in range(3):
        assert np.array_equal(d_out[i][:ns[i] * 32].cpu().numpy().reshape(-1, 32), oracle.decode(streams[i])[0])
    c.close()


def test_quality_metrics_and_the_stated_colour_tolerance(K, oracle):
    """computeQualityMetric on the GPU (exact exhaustive nearest neighbours) against the oracle's, and the colour
    tolerance the build states: the decoded colours are BYTE-IDENTICAL to the reference algorithm's
    lossy JPEG path (libjpeg-turbo's ISLOW decoder with fancy upsampling, restated and pinned to its golden vectors), so
    the PSNR of the GPU-decoded cloud equals the PSNR of the oracle-decoded cloud: tolerance 0 dB."""
    cl = synth.gen_surface(9000, 1)
    kp = K.default_params(octree_bits=7)
    c = K.Codec(kp)
    s = c.encode_batch([cl])[0]
    d = c.decode_batch([s])[0]
    rd, _ = oracle.decode(s)
    assert np.array_equal(d[:, 16:20], rd[:, 16:20])                  # the tolerance is zero: same bytes
    qg, qo = c.quality_metrics(cl, d), oracle.quality_metrics(cl, rd)
    assert (qg.in_point_count, qg.out_point_count) == (qo.in_point_count, qo.out_point_count)
    for f in ("symm_rms", "symm_hausdorff", "left_hausdorff", "right_hausdorff", "left_rms", "right_rms"):
        assert abs(getattr(qg, f) - getattr(qo, f)) <= 1e-7 * max(1.0, abs(getattr(qo, f))), f
    assert abs(qg.psnr_db - qo.psnr_db) < 1e-4
    assert max(abs(a - b) for a, b in zip(qg.psnr_yuv, qo.psnr_yuv)) < 1e-6
    # device-resident clouds, a cloud with non-finite points, and a larger pair (tiles of 2048 candidates, partial last tile)
    torch = pytest.importorskip("torch")
    big = synth.gen_surface(70001, 2)
    big["x"][7] = np.nan
    s2 = c.encode_batch([big])[0]
    d2 = c.decode_batch([s2])[0]
    q_host = c.quality_metrics(big, d2)
    ta, tb = torch.from_numpy(big.view(np.uint8).reshape(-1)).cuda(), torch.from_numpy(d2.reshape(-1)).cuda()
    q_dev = K.Quality()
    c._check(c._L.ccv2_quality_metrics(c._h, ta.data_ptr(), 70001, tb.data_ptr(), d2.shape[0], __import__("ctypes").byref(q_dev)))
    assert q_dev.symm_rms == q_host.symm_rms and max(abs(a - b) for a, b in zip(q_dev.psnr_yuv, q_host.psnr_yuv)) < 1e-9   # double atomics: summation order
    scipy_spatial = pytest.importorskip("scipy.spatial")
    fin = np.isfinite(big["x"])
    a = np.stack([big["x"], big["y"], big["z"]], 1)[fin].astype(np.float64)
    b = d2[:, :12].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
    da, _ = scipy_spatial.cKDTree(b).query(a)
    db, _ = scipy_spatial.cKDTree(a).query(b)
    assert abs(q_host.left_rms - np.sqrt((da ** 2).sum() / 70001)) < 1e-6 and abs(q_host.right_rms - np.sqrt((db ** 2).mean())) < 1e-6
    assert abs(q_host.symm_hausdorff - max(da.max(), db.max())) < 1e-6
    assert q_host.psnr_yuv[0] > 20.0                                  # Q85 JPEG of a Morton-ordered colour strip: sanity floor, not the tolerance
    c.close()


DETAIL_CASES = {
    "pcl8": dict(octree_bits=7, enh_bits=3, color_coding_type=0, color_bits=8),
    "pcl6_centroid_flag": dict(octree_bits=8, enh_bits=2, color_coding_type=0, color_bits=6, keep_centroid=1),
    "nocolor": dict(octree_bits=7, enh_bits=4, color_bits=0),
    "coarse_many_points_per_voxel": dict(octree_bits=4, enh_bits=4, color_coding_type=0, color_bits=8),
    "res_0p01": dict(octree_resolution=0.01, point_resolution=0.001, color_coding_type=0),
}


@pytest.mark.parametrize("name", sorted(DETAIL_CASES))
def test_detail_mode_bit_exact(K, oracle, name):
    """doVoxelGridDownDownSampling = false (the class default, codec.h:108-143): per-voxel point counts through the 64-bit
    int-vector range coder, per-point residuals, XOR colour differences (impl.hpp:1525-1541, 1728-1757, 1802-1832)."""
    clouds = [synth.gen_surface(40000, 400), synth.gen_uniform(3000, 401), np.zeros(0, synth.POINT_DTYPE), synth.gen_surface(1, 402), np.repeat(synth.gen_surface(5, 403), 700)]
    clouds[0]["y"][11] = np.inf
    kp = K.default_params(do_voxel_grid_downsampling=0, **DETAIL_CASES[name])
    streams = check_batch(K, oracle, clouds, kp)
    assert streams[0][53] == 0 and int.from_bytes(streams[0][55:63], "little") == 39999      # header: detail mode, object count


def test_detail_mode_round_trip_call_jpeg_encode_and_output_cloud(K, oracle):
    clouds = [synth.gen_surface(30000 + 500 * i, 410 + i) for i in range(5)]
    kp = K.default_params(do_voxel_grid_downsampling=0, octree_bits=7, enh_bits=3, color_coding_type=0)
    c = K.Codec(kp)
    arrs = [np.ascontiguousarray(cl) for cl in clouds]
    ns = [a.shape[0] for a in arrs]
    caps = [12 * n + (1 << 17) for n in ns]
    strs = [np.zeros(cp, np.uint8) for cp in caps]
    outs = [np.zeros((n, 32), np.uint8) for n in ns]
    lens, cnts = c.roundtrip_batch_raw([a.ctypes.data for a in arrs], ns, [s.ctypes.data for s in strs], caps, [o.ctypes.data for o in outs], ns)
    op = oparams(oracle, kp)
    for i, cl in enumerate(clouds):
        ref, _ = oracle.encode(cl, op, frame_id=i + 1)
        assert strs[i][:lens[i]].tobytes() == ref
        rd, _ = oracle.decode(ref)
        assert cnts[i] == ns[i] and np.array_equal(outs[i], rd)                              # every point comes back
    c.encode_batch(clouds[:1])
    assert c.output_cloud(0).shape[0] == 0                                                  # the reference's callback leaves output_ empty in detail mode
    c.close()
    # JPEG colour types encode (average through the JPEG coder + raw differences) but their decode is undefined in the reference
    kj = K.default_params(do_voxel_grid_downsampling=0, octree_bits=7, enh_bits=3, color_coding_type=1)
    cj = K.Codec(kj)
    s = cj.encode_batch(clouds[:1])[0]
    assert s == oracle.encode(clouds[0], oparams(oracle, kj), frame_id=1)[0]
    with pytest.raises(K.Ccv2Error) as ei:
        cj.decode_batch([s])
    assert ei.value.status == -3
    cj.close()
    # device-resident detail stream whose header the host cannot read: the first attempt has no detail buffers, the retry does
    torch = pytest.importorskip("torch")
    c = K.Codec(kp)
    ref, _ = oracle.encode(clouds[0], op, frame_id=1)
    d_s = torch.from_numpy(np.frombuffer(ref, np.uint8).copy()).cuda()
    d_o = torch.empty(ns[0] * 32, dtype=torch.uint8, device="cuda")
    n = c.decode_batch_raw([d_s.data_ptr()], [len(ref)], [d_o.data_ptr()], [ns[0]])
    assert n == [ns[0]] and np.array_equal(d_o.cpu().numpy().reshape(-1, 32), oracle.decode(ref)[0])
    c.close()

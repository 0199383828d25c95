"""GPU parity tests (-m gpu) of the inter-frame (predictive) path, called through the C ABI: simplifyPCloud, the
delta-frame encoder (both streams) and decoder against the CPU oracle on the same seeded inputs, bit for bit; device
pointers; the reference-shaped Python facade; properties at a size the oracle does not reach."""
import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from cwi_pcl_codec_b200 import codec
    codec.load_library()
    return codec


def recs(cloud):
    return np.ascontiguousarray(cloud).view(np.uint8).reshape(-1, 32)


def oparams(O, kp):
    return O.default_params(octree_resolution=kp.octree_resolution, point_resolution=kp.point_resolution,
                            do_color=kp.do_color_encoding, color_bit_resolution=kp.color_bit_resolution,
                            color_coding_type=kp.color_coding_type, do_centroid=kp.do_voxel_grid_centroid,
                            jpeg_quality=kp.jpeg_quality, do_voxel_grid=kp.do_voxel_grid_downsampling,
                            macroblock_size=kp.macroblock_size, do_icp_color_offset=kp.do_icp_color_offset)


@pytest.fixture(scope="module")
def gof():
    return [recs(c) for c in synth.gen_gof(60000, seed=7, frames=3)]


@pytest.mark.parametrize("kw", [dict(octree_bits=9), dict(octree_bits=8, keep_centroid=1), dict(octree_resolution=0.003, point_resolution=0.003),
                                dict(octree_bits=10, color_coding_type=0)])
def test_simplify_bit_exact(K, oracle, gof, kw):
    kp = K.default_params(**kw)
    c = K.Codec(kp)
    s = c.simplify(gof[0])
    ref = oracle.simplify(gof[0], oparams(oracle, kp))
    assert s.shape == ref.shape and np.array_equal(s, ref)
    # points outside the unit box grow it like any octree's ([PCL] adoptBoundingBoxToPoint after defineBoundingBox)
    far = gof[0][:5000].copy()
    x = np.ascontiguousarray(far[:, :12]).view(np.float32).reshape(-1, 3).copy()
    x[1234] = [1.7, 0.2, -0.4]; x[4000] = [-2.5, 0.5, 0.5]
    far[:, :12] = x.view(np.uint8).reshape(-1, 12)
    assert np.array_equal(c.simplify(far), oracle.simplify(far, oparams(oracle, kp)))
    assert c.simplify(np.zeros((0, 32), np.uint8)).shape[0] == 0
    c.close()


CASES = [
    dict(octree_bits=9),
    dict(octree_bits=9, keep_centroid=1, do_icp_color_offset=1),
    dict(octree_bits=8, do_icp_color_offset=1, macroblock_size=8),
    dict(octree_bits=10, color_coding_type=2, jpeg_quality=60),
    dict(octree_resolution=0.004, point_resolution=0.004),
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("orig", [False, True])
def test_delta_frame_streams_and_decoder_bit_exact(K, oracle, gof, kw, orig):
    """encodePointCloudDeltaFrame: P stream and I stream equal the oracle's byte for byte (the registration included: same
    neighbours, same summation order, same quantisation); the predicted frame too; decodePointCloudDeltaFrame of the
    oracle's streams equals the oracle's decode."""
    kp = K.default_params(**kw)
    op = oparams(oracle, kp)
    c = K.Codec(kp)
    s0 = c.encode_batch([gof[0]])[0]
    icloud = c.output_cloud(0)                                    # what evaluate_compression predicts from (eval.hpp:862)
    _, _, dbg = oracle.encode(gof[0], op, debug=True)
    assert np.array_equal(icloud, dbg["output_cloud"])
    i_s, p_s, info, oc = c.encode_delta(icloud, gof[1], icp_on_original=orig, want_out_cloud=True)
    ri, rp, rinfo, roc = oracle.encode_delta(icloud, gof[1], op, icp_on_original=orig, want_out_cloud=True)
    assert (info.macro_blocks, info.shared_blocks, info.converged_blocks, info.n_intra_points) == (rinfo.macro_blocks, rinfo.shared_blocks, rinfo.converged_blocks, rinfo.n_intra_points)
    assert info.converged_blocks > 0 and info.n_intra_points > 0
    assert p_s == rp, "P stream differs (%d vs %d bytes)" % (len(p_s), len(rp))
    assert i_s == ri, "I stream differs (%d vs %d bytes)" % (len(i_s), len(ri))
    assert oc.shape == roc.shape and np.array_equal(oc, roc)
    assert abs(info.shared_percentage - rinfo.shared_percentage) < 1e-7 and abs(info.convergence_percentage - rinfo.convergence_percentage) < 1e-7
    dec, nb = c.decode_delta(icloud, ri, rp)
    rdec, rnb = oracle.decode_delta(icloud, ri, rp, op)
    assert nb == rnb == info.converged_blocks
    assert dec.shape == rdec.shape and np.array_equal(dec, rdec)
    # with the reference's 16-voxel macroblocks the delta frame is smaller than the intra frame of the same cloud; always close to it
    if kp.macroblock_size == 16 and kp.octree_resolution <= 2.0 ** -9:
        assert len(i_s) + len(p_s) < len(s0)
    q = c.quality_metrics(gof[1], dec)
    assert q.symm_rms < 6 * kp.octree_resolution
    c.close()


def test_delta_frame_edge_cases(K, oracle, gof):
    kp = K.default_params(octree_bits=8)
    op = oparams(oracle, kp)
    c = K.Codec(kp)
    empty = np.zeros((0, 32), np.uint8)
    # nothing to predict from: every block is exclusive, the P stream is empty, the I stream is the intra frame of the simplified cloud
    i_s, p_s, info = c.encode_delta(empty, gof[1])
    ri, rp, rinfo = oracle.encode_delta(empty, gof[1], op)
    assert p_s == rp == b"" and i_s == ri and info.shared_blocks == 0
    dec, nb = c.decode_delta(empty, i_s, p_s)
    assert nb == 0 and np.array_equal(dec, oracle.decode_delta(empty, ri, rp, op)[0])
    # an empty P frame: both streams empty, nothing decoded
    i_s, p_s, info = c.encode_delta(gof[0], empty)
    assert i_s == b"" and p_s == b"" and info.macro_blocks == 0
    dec, nb = c.decode_delta(gof[0], b"", b"")
    assert dec.shape[0] == 0
    # identical clouds: every shared block converges at once and nothing is left for the intra coder... except blocks the gates refuse
    ic = oracle.simplify(gof[0], op)
    i_s, p_s, info = c.encode_delta(ic, gof[0])
    ri, rp, rinfo = oracle.encode_delta(ic, gof[0], op)
    assert p_s == rp and i_s == ri and info.shared_blocks == info.macro_blocks
    # a P stream with a key no I block has, a truncated chunk and trailing zero: the decoder skips / stops like the oracle
    ri, rp, _ = oracle.encode_delta(ic, gof[1], op)
    bad = bytearray(rp[:19 * 3]); bad[1:7] = (1000).to_bytes(2, "little") * 3
    for tail in (b"", b"\x00junk", bytes([18]) + b"\x01" * 7):
        s = bytes(bad) + tail
        dec, nb = c.decode_delta(ic, ri, s)
        rdec, rnb = oracle.decode_delta(ic, ri, s, op)
        assert nb == rnb and np.array_equal(dec, rdec)
    # too small a point buffer is an error, not an overrun
    with pytest.raises(K.Ccv2Error) as ei:
        c.decode_delta(ic, ri, rp, cap_points=10)
    assert ei.value.status == -4
    c.close()


def test_delta_frame_device_pointers_and_facade(K, oracle, gof):
    import torch
    kp = K.default_params(octree_bits=9)
    op = oparams(oracle, kp)
    dev = torch.device("cuda", 0)
    _, _, dbg = oracle.encode(gof[0], op, debug=True)
    icloud = dbg["output_cloud"]
    ri, rp, rinfo = oracle.encode_delta(icloud, gof[1], op)
    c = K.Codec(kp)
    d_i = torch.from_numpy(icloud.copy()).to(dev); d_p = torch.from_numpy(gof[1].copy()).to(dev)
    d_is = torch.empty(6 * gof[1].shape[0] + 65536, dtype=torch.uint8, device=dev); d_ps = torch.empty(30 * gof[1].shape[0] + 64, dtype=torch.uint8, device=dev)
    il, pl, no, info = c.encode_delta_raw(d_i.data_ptr(), icloud.shape[0], d_p.data_ptr(), gof[1].shape[0], d_is.data_ptr(), d_is.numel(), d_ps.data_ptr(), d_ps.numel())
    assert bytes(d_is[:il].cpu().numpy()) == ri and bytes(d_ps[:pl].cpu().numpy()) == rp
    d_out = torch.zeros((icloud.shape[0] + gof[1].shape[0]) * 32, dtype=torch.uint8, device=dev)
    n, nb = c.decode_delta_raw(d_i.data_ptr(), icloud.shape[0], d_is.data_ptr(), il, d_ps.data_ptr(), pl, d_out.data_ptr(), icloud.shape[0] + gof[1].shape[0])
    rdec, _ = oracle.decode_delta(icloud, ri, rp, op)
    assert n == rdec.shape[0] and np.array_equal(d_out[:32 * n].cpu().numpy().reshape(-1, 32), rdec)
    # the predicted frame into device memory, and simplifyPCloud device to device
    _, _, _, roc = oracle.encode_delta(icloud, gof[1], op, want_out_cloud=True)
    d_oc = torch.zeros(32 * (roc.shape[0] + 5), dtype=torch.uint8, device=dev)
    il, pl, no, info = c.encode_delta_raw(d_i.data_ptr(), icloud.shape[0], d_p.data_ptr(), gof[1].shape[0], d_is.data_ptr(), d_is.numel(), d_ps.data_ptr(), d_ps.numel(),
                                          out_ptr=d_oc.data_ptr(), out_cap=roc.shape[0] + 5)
    assert no == roc.shape[0] and np.array_equal(d_oc[:32 * no].cpu().numpy().reshape(-1, 32), roc)
    import ctypes as C
    nv = C.c_size_t()
    d_sv = torch.zeros(32 * gof[1].shape[0], dtype=torch.uint8, device=dev)
    c._check(c._L.ccv2_simplify(c._h, d_p.data_ptr(), gof[1].shape[0], d_sv.data_ptr(), gof[1].shape[0], C.byref(nv)))
    assert np.array_equal(d_sv[:32 * nv.value].cpu().numpy().reshape(-1, 32), oracle.simplify(gof[1], op))
    c.close()
    # the reference-shaped class: evaluate_compression's constructor arguments (eval.hpp:377-395), then its calls
    cdc = K.OctreePointCloudCodecV2(K.MANUAL_CONFIGURATION, False, 2.0 ** -9, 2.0 ** -9, True, 0, True, 8, 1, False, False, False, 85, 1)
    cdc.setMacroblockSize(16); cdc.setDoICPColorOffset(False)
    cdc.encodePointCloud(gof[0])
    i2, p2, oc = cdc.encodePointCloudDeltaFrame(cdc.getOutputCloud(), gof[1], False, False)
    assert i2 == ri and p2 == rp and oc.shape[0] == 0
    assert abs(cdc.getMacroBlockPercentage() - rinfo.shared_percentage) < 1e-7 and abs(cdc.getMacroBlockConvergencePercentage() - rinfo.convergence_percentage) < 1e-7
    assert np.array_equal(cdc.decodePointCloudDeltaFrame(icloud, i2, p2), rdec)


def test_delta_frame_at_full_size_properties(K):
    """BASELINE configs[2] shape: 1M-point frames of a GOF at 11 bits.  Beyond the oracle's reach in test time, so size-
    independent properties: the decoder reproduces the encoder's predicted frame bit for bit, every predicted block's key is
    a macroblock of the P frame, block counts add up, the frame stays within a voxel of the input, the result does not
    depend on where the buffers live."""
    import torch
    g = [recs(c) for c in synth.gen_gof(1000000, seed=0, frames=2)]
    kp = K.default_params(octree_bits=11)
    c = K.Codec(kp)
    c.encode_batch([g[0]])
    icloud = c.output_cloud(0)
    i_s, p_s, info, oc = c.encode_delta(icloud, g[1], want_out_cloud=True)
    assert info.macro_blocks >= info.shared_blocks >= info.converged_blocks > 1000
    dec, nb = c.decode_delta(icloud, i_s, p_s)
    assert nb == info.converged_blocks
    npred = dec.shape[0] - info.n_intra_points
    a = np.sort(np.ascontiguousarray(dec[:npred, :12]).view("V12").ravel())
    b = np.sort(np.ascontiguousarray(oc[:, :12]).view("V12").ravel())
    assert np.all(np.isin(a, b)) and oc.shape[0] == npred + info.n_intra_points
    assert len(i_s) + len(p_s) < len(c.encode_batch([g[1]])[0])
    q = c.quality_metrics(g[1], dec)
    assert q.symm_rms < 4 * 2.0 ** -11
    dev = torch.device("cuda", 0)
    d_i = torch.from_numpy(icloud.copy()).to(dev); d_p = torch.from_numpy(g[1].copy()).to(dev)
    d_is = torch.empty(len(i_s) + 4096, dtype=torch.uint8, device=dev); d_ps = torch.empty(len(p_s) + 4096, dtype=torch.uint8, device=dev)
    il, pl, _, info2 = c.encode_delta_raw(d_i.data_ptr(), icloud.shape[0], d_p.data_ptr(), g[1].shape[0], d_is.data_ptr(), d_is.numel(), d_ps.data_ptr(), d_ps.numel())
    assert bytes(d_is[:il].cpu().numpy()) == i_s and bytes(d_ps[:pl].cpu().numpy()) == p_s
    print("1M-point delta frame: %d macroblocks, %d shared, %d predicted, %d of %d points intra; %d + %d bytes; predict %.1f ms, intra coder %.1f ms" % (
        info.macro_blocks, info.shared_blocks, info.converged_blocks, info.n_intra_points, info.n_p_points, len(i_s), len(p_s), info2.predict_ms, info2.intra_ms))
    c.close()


def test_delta_batch_calls_equal_frame_by_frame_calls(K, oracle, gof):
    """ccv2_encode_delta_batch / ccv2_decode_delta_batch: the intra parts of all frames go through the child codec as one
    pipelined call; streams and clouds are those of frame-by-frame calls (every frame's I stream carries frame id 1)."""
    import torch
    kp = K.default_params(octree_bits=9)
    op = oparams(oracle, kp)
    dev = torch.device("cuda", 0)
    c = K.Codec(kp)
    empty = np.zeros((0, 32), np.uint8)
    ics = [oracle.simplify(gof[0], op), oracle.simplify(gof[1], op), oracle.simplify(gof[0], op), empty]
    pcs = [gof[1], gof[2], gof[0], gof[1][:5000]]                   # frame 2: identical clouds (almost nothing intra); frame 3: nothing to predict from
    single = [c.encode_delta(i, p) for i, p in zip(ics, pcs)]
    d_i = [torch.from_numpy(np.ascontiguousarray(i).reshape(-1).copy()).to(dev) for i in ics]
    d_p = [torch.from_numpy(np.ascontiguousarray(p).reshape(-1).copy()).to(dev) for p in pcs]
    d_is = [torch.empty(6 * p.shape[0] + 65536, dtype=torch.uint8, device=dev) for p in pcs]
    d_ps = [torch.empty(30 * p.shape[0] + 64, dtype=torch.uint8, device=dev) for p in pcs]
    ils, pls, infos = c.encode_delta_batch_raw([t.data_ptr() if t.numel() else None for t in d_i], [i.shape[0] for i in ics], [t.data_ptr() for t in d_p], [p.shape[0] for p in pcs],
                                               [t.data_ptr() for t in d_is], [t.numel() for t in d_is], [t.data_ptr() for t in d_ps], [t.numel() for t in d_ps])
    for k in range(4):
        assert bytes(d_is[k][:ils[k]].cpu().numpy()) == single[k][0], k
        assert bytes(d_ps[k][:pls[k]].cpu().numpy()) == single[k][1], k
        assert infos[k].converged_blocks == single[k][2].converged_blocks
        ri, rp, _ = oracle.encode_delta(ics[k], pcs[k], op)
        assert single[k][0] == ri and single[k][1] == rp
    d_out = [torch.zeros(32 * (i.shape[0] + p.shape[0] + 1), dtype=torch.uint8, device=dev) for i, p in zip(ics, pcs)]
    ns, nbs = c.decode_delta_batch_raw([t.data_ptr() if t.numel() else None for t in d_i], [i.shape[0] for i in ics], [t.data_ptr() for t in d_is], ils, [t.data_ptr() if pl else None for t, pl in zip(d_ps, pls)], pls,
                                       [t.data_ptr() for t in d_out], [t.numel() // 32 for t in d_out])
    for k in range(4):
        rdec, rnb = oracle.decode_delta(ics[k], single[k][0], single[k][1], op)
        assert ns[k] == rdec.shape[0] and nbs[k] == rnb
        assert np.array_equal(d_out[k][:32 * ns[k]].cpu().numpy().reshape(-1, 32), rdec)
    c.close()


def test_delta_calls_run_beside_submitted_intra_calls(K, oracle, gof):
    """The delta path shares no workspace with submitted intra calls of the same handle: a decode in flight and a delta
    encode / decode issued meanwhile both give their usual results (bench.py --mode inter relies on it)."""
    import torch
    kp = K.default_params(octree_bits=9)
    op = oparams(oracle, kp)
    dev = torch.device("cuda", 0)
    c = K.Codec(kp)
    streams = c.encode_batch([gof[0], gof[1], gof[2]])
    icloud = c.output_cloud(0)
    d_s = [torch.from_numpy(np.frombuffer(s, np.uint8).copy()).to(dev) for s in streams]
    d_o = [torch.zeros(32 * g.shape[0], dtype=torch.uint8, device=dev) for g in gof]
    pend = c.submit_decode_raw([t.data_ptr() for t in d_s], [t.numel() for t in d_s], [t.data_ptr() for t in d_o], [g.shape[0] for g in gof])
    i_s, p_s, info = c.encode_delta(icloud, gof[1])
    dec, nb = c.decode_delta(icloud, i_s, p_s)
    ns = pend.wait()
    ri, rp, _ = oracle.encode_delta(icloud, gof[1], op)
    assert i_s == ri and p_s == rp and np.array_equal(dec, oracle.decode_delta(icloud, ri, rp, op)[0])
    for k in range(3):
        rd, _ = oracle.decode(streams[k])
        assert ns[k] == rd.shape[0] and np.array_equal(d_o[k][:32 * ns[k]].cpu().numpy().reshape(-1, 32), rd)
    c.close()


def test_delta_frame_golden_hashes_on_the_gpu(K, golden_dir):
    """The same frozen inputs through the CUDA path WITHOUT running the oracle: stream hashes, statistics, decoded frame."""
    import hashlib
    import importlib.util
    import json
    import os
    spec = importlib.util.spec_from_file_location("make_delta_golden", os.path.join(golden_dir, "make_delta_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    want = json.load(open(os.path.join(golden_dir, "delta_hashes.json")))
    frames = m.load_inputs()
    for name, kw in m.DELTA_CASES.items():
        kw = dict(kw)
        orig = bool(kw.pop("_icp_on_original", 0))
        if "do_centroid" in kw:
            kw["do_voxel_grid_centroid"] = kw.pop("do_centroid")
        c = K.Codec(K.default_params(**kw))
        c.encode_batch([frames[0]])
        ic = c.output_cloud(0)
        i_s, p_s, info = c.encode_delta(ic, frames[1], icp_on_original=orig)
        dec, nb = c.decode_delta(ic, i_s, p_s)
        got = {"i_sha256": hashlib.sha256(i_s).hexdigest(), "p_sha256": hashlib.sha256(p_s).hexdigest(), "i_len": len(i_s), "p_len": len(p_s),
               "macro_blocks": int(info.macro_blocks), "shared_blocks": int(info.shared_blocks), "converged_blocks": int(info.converged_blocks),
               "n_intra_points": int(info.n_intra_points), "decoded_points": int(dec.shape[0]), "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest()}
        assert got == want[name], name
        c.close()


def test_delta_decoder_with_a_block_named_several_times(K, oracle, gof):
    """A P stream may name the same I macroblock more than once (nothing in the format forbids it): the decoder then writes
    more predicted points than the I cloud has.  Host and device destinations, against the oracle."""
    import torch
    kp = K.default_params(octree_bits=8)
    op = oparams(oracle, kp)
    ic_full = oracle.simplify(gof[0], op)
    ri, rp, _ = oracle.encode_delta(ic_full, gof[1], op)
    chunk = rp[:19]                                               # [18][key][6 words]
    key = np.frombuffer(chunk[1:7], np.int16).astype(np.int64)
    mres = 16.0 / 256
    xyz = np.ascontiguousarray(ic_full[:, :12]).view(np.float32).reshape(-1, 3).astype(np.float64)
    sel = np.all(np.floor(xyz / mres).astype(np.int64) == key, axis=1)
    ic = np.ascontiguousarray(ic_full[sel])                       # the I cloud is just that macroblock
    assert 0 < ic.shape[0] < 600
    ps = chunk * 5
    rdec, rnb = oracle.decode_delta(ic, ri, ps, op)
    assert rnb == 5 and rdec.shape[0] > 5 * ic.shape[0]
    c = K.Codec(kp)
    dec, nb = c.decode_delta(ic, ri, ps, cap_points=rdec.shape[0] + 7)          # host destination
    assert nb == 5 and np.array_equal(dec, rdec)
    dev = torch.device("cuda", 0)
    d_i = torch.from_numpy(ic.reshape(-1).copy()).to(dev); d_is = torch.from_numpy(np.frombuffer(ri, np.uint8).copy()).to(dev); d_ps = torch.from_numpy(np.frombuffer(ps, np.uint8).copy()).to(dev)
    d_o = torch.zeros(32 * (rdec.shape[0] + 7), dtype=torch.uint8, device=dev)
    n, nb = c.decode_delta_raw(d_i.data_ptr(), ic.shape[0], d_is.data_ptr(), len(ri), d_ps.data_ptr(), len(ps), d_o.data_ptr(), rdec.shape[0] + 7)
    assert n == rdec.shape[0] and nb == 5 and np.array_equal(d_o[:32 * n].cpu().numpy().reshape(-1, 32), rdec)
    c.close()


def test_delta_decoder_two_row_transform_mode(K, oracle, gof):
    """A chunk whose transform is stored as two matrix rows + sign word (10 words: what compressRigidTransform writes when the
    quantised quaternion misses the matrix, rigid_transform_coding_impl.hpp:110-122): decoder against the oracle."""
    kp = K.default_params(octree_bits=8)
    op = oparams(oracle, kp)
    ic = oracle.simplify(gof[0], op)
    ri, rp, _ = oracle.encode_delta(ic, gof[1], op)
    M = np.eye(4, dtype=np.float32)
    c_, s_ = np.cos(2.0), np.sin(-2.0)
    M[:2, :2] = [[c_, -s_], [s_, c_]]                             # rotation about -z by 2 rad: the coder's signed comparison drops into two-row mode
    M[:3, 3] = [0.01, -0.02, 0.005]
    w = oracle.compress_rigid_transform(M)
    assert w.size == 10
    chunks = [bytes([6 + 20]) + rp[1:7] + w.tobytes(), rp[:19], bytes([6 + 20]) + rp[20:26] + w.tobytes()]
    ps = b"".join(chunks)
    rdec, rnb = oracle.decode_delta(ic, ri, ps, op)
    assert rnb == 3
    c = K.Codec(kp)
    dec, nb = c.decode_delta(ic, ri, ps)
    assert nb == 3 and dec.shape == rdec.shape and np.array_equal(dec, rdec)
    c.close()

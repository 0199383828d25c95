"""CPU tests of the oracle's inter-frame (predictive) path (oracle/ccv2_oracle_inter.c): the restatement is checked
against independent numpy statements of the same definitions (voxel grid, Kabsch alignment, chunk format) and against
the properties the reference's own encoder/decoder pair has (the decoder reproduces the encoder's predicted cloud)."""
import struct

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth
from oracle import oracle as O


def recs(cloud):
    return np.ascontiguousarray(cloud).view(np.uint8).reshape(-1, 32)


def xyz_of(r):
    return np.ascontiguousarray(r[:, :12]).view(np.float32).reshape(-1, 3)


def rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


@pytest.fixture(scope="module")
def gof():
    return [recs(c) for c in synth.gen_gof(40000, seed=5, frames=3)]


def test_simplify_is_the_voxel_grid_of_the_unit_box(gof):
    """simplifyPCloud (impl.hpp:318-400): one point per voxel of floor(p / res) in Morton (x major) order, centre of the
    voxel, colour = integer mean per channel."""
    p = O.default_params(octree_bits=8)
    s = O.simplify(gof[0], p)
    xyz = xyz_of(gof[0]).astype(np.float64)
    k = np.floor(xyz * 256).astype(np.int64)
    lin = (k[:, 0] << 16) | (k[:, 1] << 8) | k[:, 2]
    uniq, inv = np.unique(lin, return_inverse=True)
    assert s.shape[0] == uniq.size
    sk = np.floor(xyz_of(s).astype(np.float64) * 256).astype(np.int64)
    assert np.array_equal(np.sort((sk[:, 0] << 16) | (sk[:, 1] << 8) | sk[:, 2]), uniq)
    # centres, and DFS order = ascending Morton code
    assert np.array_equal(xyz_of(s), ((sk + 0.5) / 256).astype(np.float32))
    def morton(k3):
        m = np.zeros(k3.shape[0], np.int64)
        for b in range(8):
            m |= (((k3[:, 0] >> b) & 1) << (3 * b + 2)) | (((k3[:, 1] >> b) & 1) << (3 * b + 1)) | (((k3[:, 2] >> b) & 1) << (3 * b))
        return m
    assert np.all(np.diff(morton(sk)) > 0)
    # colours
    order = np.argsort((sk[:, 0] << 16) | (sk[:, 1] << 8) | sk[:, 2])
    cnt = np.bincount(inv)
    for ch in range(3):
        sums = np.bincount(inv, weights=gof[0][:, 16 + ch].astype(np.float64)).astype(np.int64)
        assert np.array_equal(s[order, 16 + ch], (sums // cnt).astype(np.uint8))
    assert np.all(s[:, 19] == 255) and np.all(s[:, 12:16].view(np.float32) == 1.0)


def test_simplify_centroid_mode(gof):
    p = O.default_params(octree_bits=7, do_centroid=1)
    s = O.simplify(gof[0], p)
    xyz = xyz_of(gof[0])
    k = np.floor(xyz.astype(np.float64) * 128).astype(np.int64)
    lin = (k[:, 0] << 14) | (k[:, 1] << 7) | k[:, 2]
    sx = xyz_of(s)
    sk = np.floor(sx.astype(np.float64) * 128).astype(np.int64)
    slin = (sk[:, 0] << 14) | (sk[:, 1] << 7) | sk[:, 2]
    for j in range(0, s.shape[0], 97):
        m = xyz[lin == slin[j]]
        acc = np.zeros(3, np.float32)
        for q in m:
            acc = (acc + q).astype(np.float32)
        assert np.array_equal(sx[j], (acc / np.float32(m.shape[0])).astype(np.float32))


def test_rigid_transform_coder_round_trip_and_modes():
    """RigidTransformCoding (rigid_transform_coding_impl.hpp:63-203): 6 words when the quantised quaternion reproduces the
    matrix to 1e-3, the translation is quantised to 2.5 / 32767."""
    rng = np.random.default_rng(2)
    for ang in [0.0, 1e-3, 0.05, 0.7, 1.5, 2.5, 3.1]:
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = rot(rng.normal(size=3), ang)
        M[:3, 3] = rng.uniform(-0.3, 0.3, 3)
        w = O.compress_rigid_transform(M)
        assert w.size in (6, 10)
        D = O.decompress_rigid_transform(w)
        assert np.abs(D[:3, :3] - M[:3, :3]).max() < 2e-3 and np.abs(D[:3, 3] - M[:3, 3]).max() < 2.5 / 32766 * 1.01
        assert np.array_equal(D[3], [0, 0, 0, 1])
    # translations clamp at +-2.5
    M = np.eye(4, dtype=np.float32); M[:3, 3] = [7, -9, 2.4]
    D = O.decompress_rigid_transform(O.compress_rigid_transform(M))
    assert np.allclose(D[:3, 3], [2.5, -2.5, 2.4], atol=2e-4)
    # a matrix that is no rotation (ICP never returns one, the coder does not care): the quaternion cannot reproduce it -> two-row mode
    M = np.eye(4, dtype=np.float32); M[0, 0] = 0.5; M[1, 2] = 0.3
    w = O.compress_rigid_transform(M)
    assert w.size == 10
    D = O.decompress_rigid_transform(w)
    assert np.abs(D[:2, :3] - M[:2, :3]).max() < 1e-4            # the two stored rows; the third is rebuilt from column norms


def test_icp_step_is_the_kabsch_optimum():
    """One iteration on exact correspondences = the closed-form optimum (what TransformationEstimationSVD returns)."""
    rng = np.random.default_rng(3)
    src = (rng.random((200, 3)) * 0.06 + 0.45).astype(np.float32)
    R = rot([0.3, -0.2, 0.9], 0.004); t = np.array([2e-5, -1e-5, 3e-5]); c = src.astype(np.float64).mean(0)
    tgt = ((src.astype(np.float64) - c) @ R.T + c + t).astype(np.float32)      # motion far below the point spacing: nearest neighbour = the true partner
    d = np.linalg.norm(src[:, None, :].astype(np.float64) - tgt[None, :, :], axis=2)
    assert np.array_equal(d.argmin(1), np.arange(200))
    F, conv, fit, it = O.icp(src, tgt, max_iter=1)
    assert conv and it == 1
    s64, t64 = src.astype(np.float64), tgt.astype(np.float64)
    ms, mt = s64.mean(0), t64.mean(0)
    U, S, Vt = np.linalg.svd((t64 - mt).T @ (s64 - ms))
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    Rk = U @ D @ Vt
    assert np.abs(F[:3, :3] - Rk).max() < 2e-7
    assert np.abs(F[:3, 3] - (mt - Rk @ ms)).max() < 2e-7
    assert fit < 1e-12


def test_icp_recovers_a_motion_and_reports_convergence_like_pcl():
    rng = np.random.default_rng(4)
    src = (rng.random((400, 3)) * [0.06, 0.06, 0.01] + 0.4).astype(np.float32)    # a slab, like a surface patch
    R = rot([0, 0, 1], 0.01); t = np.array([1.5e-3, -1e-3, 2e-4])
    tgt = (src.astype(np.float64) @ R.T + t).astype(np.float32)
    F, conv, fit, it = O.icp(src, tgt)
    assert conv and 1 < it <= 50
    assert fit < 1e-9
    moved = src.astype(np.float64) @ F[:3, :3].astype(np.float64).T + F[:3, 3]
    assert np.abs(moved - tgt).max() < 2e-5
    # fewer than three source points: no correspondences to estimate from -> not converged (PCL: min_number_correspondences_ = 3)
    F, conv, fit, it = O.icp(src[:2], tgt)
    assert not conv and it == 0 and np.array_equal(F, np.eye(4, dtype=np.float32))


def parse_p_stream(p_s, with_offsets):
    out, pos = [], 0
    while pos < len(p_s):
        n = p_s[pos]
        body = p_s[pos + 1:pos + 1 + n]
        key = struct.unpack("<3h", body[:6])
        nw = (n - 6 - (3 if with_offsets else 0)) // 2
        words = np.frombuffer(body[6:6 + 2 * nw], np.int16)
        off = struct.unpack("<3b", body[6 + 2 * nw:]) if with_offsets else (0, 0, 0)
        out.append((key, words, off))
        pos += 1 + n
    assert pos == len(p_s)
    return out


@pytest.mark.parametrize("cen,off,orig", [(0, 0, False), (1, 1, False), (0, 1, True)])
def test_delta_frame_format_and_decoder_pair(gof, cen, off, orig):
    """encodePointCloudDeltaFrame / decodePointCloudDeltaFrame (impl.hpp:787-1235): chunk layout
    [u8 size][3 x i16 key][6 | 10 x i16][3 x i8 offsets], keys ascending in DFS order, the decoder's predicted points =
    the encoder's own (up to the colour-offset doubling of impl.hpp:1187-1189), the rest = an ordinary intra frame."""
    p = O.default_params(octree_bits=9, do_centroid=cen, do_icp_color_offset=off)
    _, _, dbg = O.encode(gof[0], p, debug=True)
    icloud = dbg["output_cloud"]
    i_s, p_s, info, oc = O.encode_delta(icloud, gof[1], p, icp_on_original=orig, want_out_cloud=True)
    chunks = parse_p_stream(p_s, off)
    assert len(chunks) == info.converged_blocks > 0
    assert info.converged_blocks <= info.shared_blocks <= info.macro_blocks
    assert all(c[1].size in (6, 10) for c in chunks)
    def mort(k):
        m = 0
        for b in range(15):
            m |= (((k[0] >> b) & 1) << (3 * b + 2)) | (((k[1] >> b) & 1) << (3 * b + 1)) | (((k[2] >> b) & 1) << (3 * b))
        return m
    codes = [mort(c[0]) for c in chunks]
    assert codes == sorted(codes) and len(set(codes)) == len(codes)
    dec, nb = O.decode_delta(icloud, i_s, p_s, p)
    assert nb == info.converged_blocks
    intra, _ = O.decode(i_s) if len(i_s) else (np.zeros((0, 32), np.uint8), None)
    npred = dec.shape[0] - intra.shape[0]
    assert np.array_equal(dec[npred:], intra)
    # the encoder's out cloud interleaves predicted blocks and unpredicted points in block order; the predicted points are
    # the I points of the block moved by the DEcoded transform: compare as multisets of xyz
    a = np.sort(np.ascontiguousarray(dec[:npred, :12]).view("V12").ravel())
    bb = np.ascontiguousarray(oc[:, :12]).view("V12").ravel()
    assert np.all(np.isin(a, bb))
    assert oc.shape[0] == npred + info.n_intra_points
    # every predicted point lies near its source block (a 16-voxel macroblock moved by a small transform)
    mres = 16 / 512
    mk = np.floor(xyz_of(dec[:npred]).astype(np.float64) / mres).astype(int)
    keys = {c[0] for c in chunks}
    near = sum(1 for k in map(tuple, mk) if any((k[0] + d0, k[1] + d1, k[2] + d2) in keys for d0 in (-1, 0, 1) for d1 in (-1, 0, 1) for d2 in (-1, 0, 1)))
    assert near == npred
    q = O.quality_metrics(gof[1], dec)
    assert q.symm_rms < 4 * 2.0 ** -9
    if not off:
        assert q.psnr_yuv[0] > 25


def test_delta_frame_of_disjoint_clouds_is_all_intra(gof):
    p = O.default_params(octree_bits=8)
    far = gof[1].copy()
    x = xyz_of(far).copy(); x[:, 0] = x[:, 0] * 0.2 + 0.01          # squeezed to the box's edge: no shared macroblock
    far[:, :12] = x.view(np.uint8).reshape(-1, 12)
    ic = gof[0].copy(); xi = xyz_of(ic).copy(); xi[:, 0] = xi[:, 0] * 0.2 + 0.79; ic[:, :12] = xi.view(np.uint8).reshape(-1, 12)
    i_s, p_s, info = O.encode_delta(ic, far, p)
    assert info.shared_blocks == 0 and len(p_s) == 0
    dec, nb = O.decode_delta(ic, i_s, p_s, p)
    simp = O.simplify(far, p)
    ref, _ = O.decode(O.encode(simp, O.default_params(octree_bits=8, create_scalable=1, jpeg_quality=75))[0])
    assert nb == 0 and np.array_equal(dec, ref)


def test_delta_frame_golden_hashes(golden_dir):
    """Frozen inputs (tests/golden/delta_inputs.npz) and the SHA-256 of the oracle's I and P streams, block statistics and
    decoded frame (tests/golden/delta_hashes.json, written by make_delta_golden.py)."""
    import importlib.util
    import json
    import os
    spec = importlib.util.spec_from_file_location("make_delta_golden", os.path.join(golden_dir, "make_delta_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    want = json.load(open(os.path.join(golden_dir, "delta_hashes.json")))
    frames = m.load_inputs()
    assert sorted(want) == sorted(m.DELTA_CASES)
    for name, kw in m.DELTA_CASES.items():
        assert m.run_case(frames, kw) == want[name], name


def test_generate_delta_frame_layout_is_the_chunk_list_without_size_bytes(gof):
    """generatePointCloudDeltaFrame (impl.hpp:650-660) writes key, transform words and offsets of every predicted macroblock
    with no size byte in front; the facades produce it by re-framing encodePointCloudDeltaFrame's stream."""
    from cwi_pcl_codec_b200 import codec as K
    p = O.default_params(octree_bits=9, do_icp_color_offset=1)
    _, _, dbg = O.encode(gof[0], p, debug=True)
    _, p_s, info = O.encode_delta(dbg["output_cloud"], gof[1], p)
    g = K.strip_chunk_sizes(p_s)
    chunks = parse_p_stream(p_s, True)
    assert len(g) == len(p_s) - len(chunks)
    pos = 0
    for key, words, off in chunks:
        n = 6 + 2 * words.size + 3
        assert struct.unpack("<3h", g[pos:pos + 6]) == key and np.array_equal(np.frombuffer(g[pos + 6:pos + 6 + 2 * words.size], np.int16), words)
        assert struct.unpack("<3b", g[pos + n - 3:pos + n]) == off
        pos += n
    assert pos == len(g)
    with pytest.raises(ValueError):
        K.strip_chunk_sizes(p_s[:-1])


def test_quaternion_coder_branches_and_the_signed_comparison_quirk():
    """QuaternionCoding (quaternion_coding_impl.hpp:55-222) picks the component it drops by SIGNED comparisons
    (`w > x && w > y && w > z`, ...): all four cases are reachable, and a rotation whose largest component is negative (about
    -z by 2 rad: w = 0.54, z = -0.84) lands in the w case, clamps z and fails the 1e-3 check of compressRigidTransform
    (rigid_transform_coding_impl.hpp:99-108), which then stores two matrix rows and a sign word instead."""
    def which(words):
        return ((int(words[1]) & 1) << 1) | (int(words[2]) & 1)
    seen = set()
    for axis, ang, want in [([1, 0, 0], 0.3, 3), ([1, 0, 0], 2.0, 0), ([0, 1, 0], 2.9, 1), ([0, 0, 1], 3.1, 2), ([1, 1, 1], 2.0, 3), ([1, 1, 1], 2.9, 0)]:
        M = np.eye(4, dtype=np.float32); M[:3, :3] = rot(axis, ang)
        w = O.compress_rigid_transform(M)
        assert w.size == 6 and which(w) == want, (axis, ang)
        seen.add(want)
        assert np.abs(O.decompress_rigid_transform(w)[:3, :3] - M[:3, :3]).max() < 1e-4
    assert seen == {0, 1, 2, 3}
    M = np.eye(4, dtype=np.float32); M[:3, :3] = rot([0, 0, -1], 2.0)
    w = O.compress_rigid_transform(M)
    assert w.size == 10                                           # two rows + sign word + translation
    D = O.decompress_rigid_transform(w)
    assert np.abs(D[:2, :3] - M[:2, :3]).max() < 1e-4 and np.abs(D[2, :3] - M[2, :3]).max() < 5e-3   # third row rebuilt from column norms
    assert int(w[6]) == sum(1 << l for l in range(3) if M[2, l] < 0)

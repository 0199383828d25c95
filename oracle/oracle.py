"""ctypes binding of the CPU oracle (oracle/ccv2_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libccv2_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [("point_resolution", C.c_double), ("octree_resolution", C.c_double),
                ("do_voxel_grid", C.c_int), ("do_color", C.c_int), ("color_bit_resolution", C.c_int),
                ("color_coding_type", C.c_int), ("do_centroid", C.c_int), ("create_scalable", C.c_int),
                ("code_connectivity", C.c_int), ("jpeg_quality", C.c_int), ("macroblock_size", C.c_int),
                ("do_icp_color_offset", C.c_int)]


class OrcInfo(C.Structure):
    _fields_ = [("depth", C.c_uint32), ("bb_min", C.c_double * 3), ("bb_max", C.c_double * 3),
                ("n_finite", C.c_uint64), ("n_leaves", C.c_uint64), ("n_tree_bytes", C.c_uint64),
                ("n_color_bytes", C.c_uint64), ("coded", C.c_uint64 * 3), ("t_ms", C.c_double * 8)]


class OrcDebug(C.Structure):
    _fields_ = [("leaf_keys", C.POINTER(C.c_uint64)), ("tree_bytes", C.POINTER(C.c_uint8)),
                ("avg_colors", C.POINTER(C.c_uint8)), ("color_payload", C.POINTER(C.c_uint8)),
                ("centroid_bytes", C.POINTER(C.c_uint8)), ("output_cloud", C.POINTER(C.c_uint8))]


class OrcDeltaInfo(C.Structure):
    _fields_ = [("macro_blocks", C.c_uint64), ("shared_blocks", C.c_uint64), ("converged_blocks", C.c_uint64),
                ("n_intra_points", C.c_uint64), ("n_p_points", C.c_uint64),
                ("shared_percentage", C.c_float), ("convergence_percentage", C.c_float)]


class OrcQuality(C.Structure):
    _fields_ = [("in_point_count", C.c_uint64), ("out_point_count", C.c_uint64), ("symm_rms", C.c_float), ("symm_hausdorff", C.c_float),
                ("left_hausdorff", C.c_float), ("right_hausdorff", C.c_float), ("left_rms", C.c_float), ("right_rms", C.c_float),
                ("psnr_db", C.c_double), ("psnr_yuv", C.c_double * 3)]


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "ccv2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "ccv2_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
        L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
        L.orc_encode.argtypes = [C.POINTER(OrcParams), C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(u8p), szp,
                                 C.POINTER(OrcInfo), C.POINTER(OrcDebug)]
        L.orc_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), szp, C.POINTER(OrcInfo)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free_debug.argtypes = [C.POINTER(OrcDebug)]
        L.orc_range_encode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(u8p), szp]
        L.orc_range_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
        L.orc_jpeg_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(u8p), szp]
        L.orc_jpeg_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_snake_positions_literal.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_snake_pos_closed.argtypes = [C.c_int, C.c_int, C.c_int64]
        L.orc_snake_pos_closed.restype = C.c_int32
        L.orc_bbox_keys.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
        L.orc_dfs_recursive.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(u8p), szp]
        L.orc_quality_metrics.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(OrcQuality)]
        L.orc_simplify.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), szp]
        L.orc_encode_delta.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                                       C.POINTER(u8p), szp, C.POINTER(u8p), szp, C.POINTER(C.c_void_p), szp, C.POINTER(OrcDeltaInfo)]
        L.orc_decode_delta.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_void_p), szp, C.POINTER(C.c_uint64)]
        L.orc_compress_rigid_transform.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_decompress_rigid_transform.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_icp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_void_p,
                              C.POINTER(C.c_int), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def default_params(**kw):
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    if "octree_bits" in kw:
        bits = kw.pop("octree_bits")
        enh = kw.pop("enh_bits", 0)
        p.octree_resolution = 2.0 ** -bits
        p.point_resolution = 2.0 ** -(bits + enh)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _take(ptr, n):
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.uint8)
    return out


def encode(points, params=None, frame_id=1, debug=False):
    """points: (n, 8) float32 view of 32-byte PointXYZRGB records (or uint8 (n,32)). Returns (bytes, info[, dbg])."""
    L = lib()
    p = params or default_params()
    a = np.ascontiguousarray(points)
    assert a.nbytes % 32 == 0
    n = a.nbytes // 32
    out = C.POINTER(C.c_uint8)()
    ol = C.c_size_t()
    info = OrcInfo()
    dbg = OrcDebug()
    rc = L.orc_encode(C.byref(p), frame_id, a.ctypes.data, n, C.byref(out), C.byref(ol), C.byref(info),
                      C.byref(dbg) if debug else None)
    if rc:
        raise RuntimeError("orc_encode rc=%d" % rc)
    data = _take(out, ol.value).tobytes()
    L.orc_free(out)
    if not debug:
        return data, info
    V, B = info.n_leaves, info.n_tree_bytes
    d = dict(
        leaf_keys=np.ctypeslib.as_array(dbg.leaf_keys, shape=(V,)).copy() if V else np.zeros(0, np.uint64),
        tree_bytes=_take(dbg.tree_bytes, B),
        avg_colors=_take(dbg.avg_colors, 3 * V) if dbg.avg_colors else np.zeros(0, np.uint8),
        color_payload=_take(dbg.color_payload, info.n_color_bytes) if dbg.color_payload else np.zeros(0, np.uint8),
        centroid_bytes=_take(dbg.centroid_bytes, 3 * V) if dbg.centroid_bytes else np.zeros(0, np.uint8),
        output_cloud=_take(dbg.output_cloud, 32 * V).reshape(-1, 32) if dbg.output_cloud else np.zeros((0, 32), np.uint8),
    )
    L.orc_free_debug(C.byref(dbg))
    return data, info, d


def decode(data):
    L = lib()
    buf = np.frombuffer(data, np.uint8)
    pts = C.c_void_p()
    n = C.c_size_t()
    info = OrcInfo()
    rc = L.orc_decode(buf.ctypes.data, buf.size, C.byref(pts), C.byref(n), C.byref(info))
    if rc:
        raise RuntimeError("orc_decode rc=%d" % rc)
    arr = np.ctypeslib.as_array(C.cast(pts, C.POINTER(C.c_uint8)), shape=(n.value * 32,)).copy() if n.value else np.zeros(0, np.uint8)
    L.orc_free(pts)
    return arr.reshape(-1, 32), info


def range_encode(data):
    L = lib()
    a = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    out = C.POINTER(C.c_uint8)()
    ol = C.c_size_t()
    L.orc_range_encode(a.ctypes.data if a.size else None, a.size, C.byref(out), C.byref(ol))
    r = _take(out, ol.value)
    L.orc_free(out)
    return r


def range_decode(stream, n):
    L = lib()
    a = np.ascontiguousarray(stream, np.uint8)
    out = np.zeros(max(n, 1), np.uint8)
    used = C.c_size_t()
    rc = L.orc_range_decode(a.ctypes.data, a.size, out.ctypes.data, n, C.byref(used))
    if rc:
        raise RuntimeError("orc_range_decode rc=%d" % rc)
    return out[:n], used.value


def jpeg_encode(rgb, quality):
    L = lib()
    a = np.ascontiguousarray(rgb, np.uint8)
    h, w = a.shape[:2]
    out = C.POINTER(C.c_uint8)()
    ol = C.c_size_t()
    rc = L.orc_jpeg_encode(a.ctypes.data, w, h, quality, C.byref(out), C.byref(ol))
    if rc:
        raise RuntimeError("orc_jpeg_encode rc=%d" % rc)
    r = _take(out, ol.value)
    L.orc_free(out)
    return r


def jpeg_decode(data):
    L = lib()
    a = np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8))
    out = C.POINTER(C.c_uint8)()
    w, h = C.c_int(), C.c_int()
    rc = L.orc_jpeg_decode(a.ctypes.data, a.size, C.byref(out), C.byref(w), C.byref(h))
    if rc:
        raise RuntimeError("orc_jpeg_decode rc=%d" % rc)
    r = _take(out, w.value * h.value * 3).reshape(h.value, w.value, 3)
    L.orc_free(out)
    return r


def snake_literal(w, h):
    pos = np.zeros(w * h, np.int32)
    lib().orc_snake_positions_literal(w, h, pos.ctypes.data)
    return pos


def snake_closed(w, h):
    L = lib()
    return np.array([L.orc_snake_pos_closed(w, h, i) for i in range(w * h)], np.int32)


def bbox_keys(points, resolution):
    L = lib()
    a = np.ascontiguousarray(points)
    n = a.nbytes // 32
    bmin, bmax = (C.c_double * 3)(), (C.c_double * 3)()
    depth = C.c_uint32()
    keys = np.zeros((n, 3), np.uint32)
    fin = np.zeros(n, np.uint8)
    rc = L.orc_bbox_keys(a.ctypes.data, n, resolution, bmin, bmax, C.byref(depth), keys.ctypes.data, fin.ctypes.data)
    if rc < 0:
        raise RuntimeError("orc_bbox_keys rc=%d" % rc)
    return np.array(bmin), np.array(bmax), depth.value, keys, fin.astype(bool)


def dfs_recursive(leaf_codes, depth):
    L = lib()
    a = np.ascontiguousarray(leaf_codes, np.uint64)
    out = C.POINTER(C.c_uint8)()
    ol = C.c_size_t()
    L.orc_dfs_recursive(a.ctypes.data, a.size, depth, C.byref(out), C.byref(ol))
    r = _take(out, ol.value)
    L.orc_free(out)
    return r


def quality_metrics(cloud_a, cloud_b):
    """computeQualityMetric(original, decoded) (quality_metrics_impl.hpp:82-239); O(na * nb), test sizes only."""
    a, b = np.ascontiguousarray(cloud_a), np.ascontiguousarray(cloud_b)
    q = OrcQuality()
    lib().orc_quality_metrics(a.ctypes.data, a.nbytes // 32, b.ctypes.data, b.nbytes // 32, C.byref(q))
    return q


# ---------------------------------------------------------------- inter-frame (predictive) path
def _take_points(ptr, n):
    if not n:
        return np.zeros((0, 32), np.uint8)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * 32,)).copy()
    lib().orc_free(ptr)
    return arr.reshape(-1, 32)


def simplify(points, params):
    """simplifyPCloud (impl.hpp:318-400): (V, 32) uint8 records."""
    a = np.ascontiguousarray(points)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().orc_simplify(C.byref(params), a.ctypes.data, a.nbytes // 32, C.byref(out), C.byref(n))
    if rc:
        raise RuntimeError("orc_simplify rc=%d" % rc)
    return _take_points(out, n.value)


def encode_delta(icloud, pcloud, params, icp_on_original=False, want_out_cloud=False):
    """encodePointCloudDeltaFrame (impl.hpp:787-1112) -> (i_stream, p_stream, info[, predicted cloud])."""
    L = lib()
    ic, pc = np.ascontiguousarray(icloud), np.ascontiguousarray(pcloud)
    io, po = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
    il, pl = C.c_size_t(), C.c_size_t()
    oc, on = C.c_void_p(), C.c_size_t()
    info = OrcDeltaInfo()
    rc = L.orc_encode_delta(C.byref(params), ic.ctypes.data, ic.nbytes // 32, pc.ctypes.data, pc.nbytes // 32, int(icp_on_original),
                            C.byref(io), C.byref(il), C.byref(po), C.byref(pl),
                            C.byref(oc) if want_out_cloud else None, C.byref(on) if want_out_cloud else None, C.byref(info))
    if rc:
        raise RuntimeError("orc_encode_delta rc=%d" % rc)
    i_s, p_s = _take(io, il.value).tobytes(), _take(po, pl.value).tobytes()
    L.orc_free(io); L.orc_free(po)
    if want_out_cloud:
        return i_s, p_s, info, _take_points(oc, on.value)
    return i_s, p_s, info


def decode_delta(icloud, i_stream, p_stream, params):
    """decodePointCloudDeltaFrame (impl.hpp:1120-1235) -> ((n, 32) uint8 records, decoded macroblocks)."""
    ic = np.ascontiguousarray(icloud)
    a, b = np.frombuffer(i_stream, np.uint8), np.frombuffer(p_stream, np.uint8)
    out, n, nb = C.c_void_p(), C.c_size_t(), C.c_uint64()
    rc = lib().orc_decode_delta(C.byref(params), ic.ctypes.data, ic.nbytes // 32, a.ctypes.data if a.size else None, a.size,
                                b.ctypes.data if b.size else None, b.size, C.byref(out), C.byref(n), C.byref(nb))
    if rc:
        raise RuntimeError("orc_decode_delta rc=%d" % rc)
    return _take_points(out, n.value), nb.value


def compress_rigid_transform(m):
    """4x4 float32 (row-major) -> int16 words (6: quaternion mode, 10: two-row mode)."""
    a = np.ascontiguousarray(m, np.float32)
    out = np.zeros(10, np.int16)
    n = C.c_int()
    lib().orc_compress_rigid_transform(a.ctypes.data, out.ctypes.data, C.byref(n))
    return out[:n.value].copy()


def decompress_rigid_transform(words):
    w = np.ascontiguousarray(words, np.int16)
    m = np.zeros((4, 4), np.float32)
    lib().orc_decompress_rigid_transform(w.ctypes.data, w.size, m.ctypes.data)
    return m


def icp(src_xyz, tgt_xyz, max_iter=50, tf_eps=float(np.float32(1e-8)), fit_eps=float(np.float32(3) * np.float32(1e-8))):
    """-> (4x4 float32, converged, fitness, iterations)."""
    s, t = np.ascontiguousarray(src_xyz, np.float32), np.ascontiguousarray(tgt_xyz, np.float32)
    F = np.zeros((4, 4), np.float32)
    conv, fit = C.c_int(), C.c_double()
    it = lib().orc_icp(s.ctypes.data, s.shape[0], t.ctypes.data, t.shape[0], max_iter, tf_eps, fit_eps, F.ctypes.data, C.byref(conv), C.byref(fit))
    return F, bool(conv.value), fit.value, it

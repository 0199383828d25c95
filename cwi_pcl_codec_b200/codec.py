"""ctypes binding of libccv2.so (the C ABI in include/ccv2.h) and a Python mirror of the reference's
``pcl::io::OctreePointCloudCodecV2<PointXYZRGB>`` interface (cloud_codec_v2/include/pcl/cloud_codec_v2/
point_cloud_codec_v2.h:70-368) for tests and bench.

There is no CPU path here: if libccv2.so is missing or no CUDA device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCV2_LIBRARY") or os.path.join(_HERE, "libccv2.so")   # CCV2_LIBRARY: developer override (A/B builds of the same ABI)

MANUAL_CONFIGURATION = 13  # pcl::io::compression_Profiles_e (12 profiles, COMPRESSION_PROFILE_COUNT, MANUAL_CONFIGURATION)


class Ccv2Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("ccv2 status %d (%s): %s" % (status, _status_string(status), msg))
        self.status = status


class Params(C.Structure):
    """ccv2_params == the constructor surface of OctreePointCloudCodecV2 (codec.h:108-143)."""
    _fields_ = [("profile", C.c_int32), ("show_statistics", C.c_int32),
                ("point_resolution", C.c_double), ("octree_resolution", C.c_double),
                ("do_voxel_grid_downsampling", C.c_int32), ("i_frame_rate", C.c_uint32),
                ("do_color_encoding", C.c_int32), ("color_bit_resolution", C.c_uint8),
                ("color_coding_type", C.c_uint8), ("_pad0", C.c_uint8 * 2),
                ("do_voxel_grid_centroid", C.c_int32), ("create_scalable_stream", C.c_int32),
                ("code_connectivity", C.c_int32), ("jpeg_quality", C.c_int32), ("num_threads", C.c_int32),
                ("macroblock_size", C.c_int32), ("do_icp_color_offset", C.c_int32)]


class FrameInfo(C.Structure):
    _fields_ = [("depth", C.c_uint32), ("n_finite", C.c_uint32), ("n_leaves", C.c_uint32),
                ("n_tree_bytes", C.c_uint32), ("n_color_bytes", C.c_uint32), ("error", C.c_uint32),
                ("bb_min", C.c_double * 3), ("bb_max", C.c_double * 3), ("coded", C.c_uint64 * 3)]


class Quality(C.Structure):
    """ccv2_quality == QualityMetric's computed fields (quality_metrics.h:53-75)."""
    _fields_ = [("in_point_count", C.c_uint64), ("out_point_count", C.c_uint64), ("symm_rms", C.c_float), ("symm_hausdorff", C.c_float),
                ("left_hausdorff", C.c_float), ("right_hausdorff", C.c_float), ("left_rms", C.c_float), ("right_rms", C.c_float),
                ("psnr_db", C.c_double), ("psnr_yuv", C.c_double * 3)]


class DeltaInfo(C.Structure):
    """ccv2_delta_info: the prediction statistics of one delta frame (impl.hpp:803-805, 1105-1106)."""
    _fields_ = [("macro_blocks", C.c_uint64), ("shared_blocks", C.c_uint64), ("converged_blocks", C.c_uint64),
                ("n_intra_points", C.c_uint64), ("n_p_points", C.c_uint64),
                ("shared_percentage", C.c_float), ("convergence_percentage", C.c_float),
                ("predict_ms", C.c_float), ("intra_ms", C.c_float)]


_lib = None


def load_library():
    """Loads libccv2.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libccv2.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C cwi_pcl_codec_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vpp, szp = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)
    L.ccv2_default_params.argtypes = [C.POINTER(Params)]
    L.ccv2_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]
    L.ccv2_destroy.argtypes = [C.c_void_p]
    L.ccv2_destroy.restype = None
    L.ccv2_max_compressed_size.argtypes = [C.c_size_t]
    L.ccv2_max_compressed_size.restype = C.c_size_t
    L.ccv2_encode_batch.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp]
    L.ccv2_decode_batch.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp]
    L.ccv2_roundtrip_batch.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp, vpp, szp, szp]
    ip = C.POINTER(C.c_int)
    L.ccv2_submit_encode.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp, ip]
    L.ccv2_submit_decode.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp, ip]
    L.ccv2_submit_roundtrip.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, szp, vpp, szp, szp, ip]
    L.ccv2_wait.argtypes = [C.c_void_p, C.c_int]
    L.ccv2_timer_start.argtypes = [C.c_void_p]
    L.ccv2_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.ccv2_split_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, szp]
    L.ccv2_encode_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, vpp, szp, szp, szp]
    L.ccv2_quality_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Quality)]
    L.ccv2_peek_point_count.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    L.ccv2_get_metrics.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ccv2_set_frame_id.argtypes = [C.c_void_p, C.c_uint32]
    L.ccv2_get_frame_id.argtypes = [C.c_void_p]
    L.ccv2_get_frame_id.restype = C.c_uint32
    L.ccv2_last_launch_count.argtypes = [C.c_void_p]
    L.ccv2_last_launch_count.restype = C.c_uint64
    L.ccv2_last_device_ms.argtypes = [C.c_void_p]
    L.ccv2_last_device_ms.restype = C.c_float
    L.ccv2_last_error.argtypes = [C.c_void_p]
    L.ccv2_last_error.restype = C.c_char_p
    L.ccv2_status_string.argtypes = [C.c_int]
    L.ccv2_status_string.restype = C.c_char_p
    L.ccv2_host_alloc.argtypes = [C.c_size_t]
    L.ccv2_host_alloc.restype = C.c_void_p
    L.ccv2_host_free.argtypes = [C.c_void_p]
    L.ccv2_host_free.restype = None
    L.ccv2_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, szp]
    L.ccv2_get_output_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, szp]
    L.ccv2_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.ccv2_encode_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, szp,
                                    C.c_void_p, C.c_size_t, szp, C.c_void_p, C.c_size_t, szp, C.POINTER(DeltaInfo)]
    L.ccv2_decode_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_size_t, szp, C.POINTER(C.c_uint64)]
    L.ccv2_simplify.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
    L.ccv2_encode_delta_batch.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, C.c_int, vpp, szp, szp, vpp, szp, szp, C.POINTER(DeltaInfo)]
    L.ccv2_decode_delta_batch.argtypes = [C.c_void_p, C.c_int, vpp, szp, vpp, szp, vpp, szp, vpp, szp, szp, C.POINTER(C.c_uint64)]
    L.ccv2_max_p_stream_size.argtypes = [C.c_size_t]
    L.ccv2_max_p_stream_size.restype = C.c_size_t
    L.ccv2_get_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    _lib = L
    return L


EXPORTED_SYMBOLS = ["ccv2_default_params", "ccv2_create", "ccv2_destroy", "ccv2_max_compressed_size",
                    "ccv2_encode_batch", "ccv2_decode_batch", "ccv2_roundtrip_batch", "ccv2_peek_point_count", "ccv2_get_metrics",
                    "ccv2_set_frame_id", "ccv2_get_frame_id", "ccv2_last_launch_count", "ccv2_last_device_ms",
                    "ccv2_last_error", "ccv2_status_string", "ccv2_host_alloc", "ccv2_host_free", "ccv2_debug_fetch", "ccv2_get_output_cloud",
                    "ccv2_set_profiling", "ccv2_get_profile", "ccv2_submit_encode", "ccv2_submit_decode", "ccv2_submit_roundtrip",
                    "ccv2_wait", "ccv2_timer_start", "ccv2_timer_stop", "ccv2_quality_metrics", "ccv2_split_tiles", "ccv2_encode_tiles",
                    "ccv2_encode_delta", "ccv2_decode_delta", "ccv2_simplify", "ccv2_max_p_stream_size", "ccv2_encode_delta_batch", "ccv2_decode_delta_batch"]


def _status_string(s):
    try:
        return load_library().ccv2_status_string(s).decode()
    except Exception:  # pragma: no cover
        return "?"


def default_params(**kw):
    """evaluate_compression's configuration (eval.hpp:377-395) with parameter_config.txt values; keyword
    overrides use the ccv2_params field names, or octree_bits / enh_bits / color_bits / keep_centroid."""
    p = Params()
    load_library().ccv2_default_params(C.byref(p))
    bits = kw.pop("octree_bits", None)
    enh = kw.pop("enh_bits", 0)
    if bits is not None:
        p.octree_resolution = 2.0 ** -bits
        p.point_resolution = 2.0 ** -(bits + enh)
    if "color_bits" in kw:
        cb = kw.pop("color_bits")
        p.color_bit_resolution = cb
        p.do_color_encoding = 1 if cb > 0 else 0      # eval.hpp:387
    if "keep_centroid" in kw:
        p.do_voxel_grid_centroid = kw.pop("keep_centroid")
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class PinnedBuffer:
    """cudaMallocHost'ed byte buffer exposed as a numpy array (for full-speed PCIe copies)."""

    def __init__(self, nbytes):
        L = load_library()
        self.nbytes = int(nbytes)
        self.ptr = L.ccv2_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError("ccv2_host_alloc(%d) failed" % nbytes)
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(1, self.nbytes),))

    def close(self):
        if self.ptr:
            load_library().ccv2_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr_len(x):
    """(address, nbytes) of a numpy array / bytes / (ptr, nbytes) tuple / object with data_ptr() (torch tensor)."""
    if isinstance(x, tuple):
        return int(x[0]), int(x[1])
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C contiguous")
        return x.ctypes.data, x.nbytes
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr()), int(x.numel() * x.element_size())
    if isinstance(x, (bytes, bytearray, memoryview)):
        a = np.frombuffer(x, np.uint8)
        return a.ctypes.data, a.size
    raise TypeError(type(x))


class Pending:
    """A submitted call (ccv2_submit_*): holds the ctypes arrays the library still reads and writes."""

    def __init__(self, codec, ticket, keep, result):
        self.codec, self.ticket, self._keep, self._result = codec, ticket, keep, result

    def wait(self):
        self.codec._check(self.codec._L.ccv2_wait(self.codec._h, self.ticket))
        return self._result()


class Codec:
    """Thin handle over ccv2_codec: batch encode/decode with host or device buffers."""

    def __init__(self, params=None, device=0):
        self._L = load_library()
        self.params = params or default_params()
        h = C.c_void_p()
        rc = self._L.ccv2_create(C.byref(self.params), device, C.byref(h))
        if rc:
            raise Ccv2Error(rc, (self._L.ccv2_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ccv2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise Ccv2Error(rc, (self._L.ccv2_last_error(self._h) or b"").decode())

    # ---- raw pointer API (host or device addresses)
    def encode_batch_raw(self, in_ptrs, npts, out_ptrs, out_caps):
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*npts)
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_len = (C.c_size_t * n)()
        rc = self._L.ccv2_encode_batch(self._h, n, a_in, a_n, a_out, a_cap, a_len)
        self._check(rc)
        return list(a_len)

    def decode_batch_raw(self, in_ptrs, in_lens, out_ptrs, out_caps):
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*in_lens)
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_len = (C.c_size_t * n)()
        rc = self._L.ccv2_decode_batch(self._h, n, a_in, a_n, a_out, a_cap, a_len)
        self._check(rc)
        return list(a_len)

    def roundtrip_batch_raw(self, in_ptrs, npts, str_ptrs, str_caps, out_ptrs, out_caps):
        """encode -> decode in one pipelined call; returns (stream lengths, decoded point counts)."""
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*npts)
        a_str = (C.c_void_p * n)(*str_ptrs) if str_ptrs is not None else None
        a_scap = (C.c_size_t * n)(*(str_caps if str_caps is not None else [0] * n))
        a_slen = (C.c_size_t * n)()
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_np = (C.c_size_t * n)()
        rc = self._L.ccv2_roundtrip_batch(self._h, n, a_in, a_n, a_str, a_scap, a_slen, a_out, a_cap, a_np)
        self._check(rc)
        return list(a_slen), list(a_np)

    # ---- asynchronous forms: submit returns a Pending (keeps the argument arrays alive); wait() returns what the synchronous call returns
    def submit_roundtrip_raw(self, in_ptrs, npts, str_ptrs, str_caps, out_ptrs, out_caps):
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*npts)
        a_str = (C.c_void_p * n)(*str_ptrs) if str_ptrs is not None else None
        a_scap = (C.c_size_t * n)(*(str_caps if str_caps is not None else [0] * n))
        a_slen = (C.c_size_t * n)()
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_np = (C.c_size_t * n)()
        t = C.c_int()
        self._check(self._L.ccv2_submit_roundtrip(self._h, n, a_in, a_n, a_str, a_scap, a_slen, a_out, a_cap, a_np, C.byref(t)))
        return Pending(self, t.value, (a_in, a_n, a_str, a_scap, a_out, a_cap), lambda: (list(a_slen), list(a_np)))

    def submit_encode_raw(self, in_ptrs, npts, out_ptrs, out_caps):
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*npts)
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_len = (C.c_size_t * n)()
        t = C.c_int()
        self._check(self._L.ccv2_submit_encode(self._h, n, a_in, a_n, a_out, a_cap, a_len, C.byref(t)))
        return Pending(self, t.value, (a_in, a_n, a_out, a_cap), lambda: list(a_len))

    def submit_decode_raw(self, in_ptrs, in_lens, out_ptrs, out_caps):
        n = len(in_ptrs)
        a_in = (C.c_void_p * n)(*in_ptrs)
        a_n = (C.c_size_t * n)(*in_lens)
        a_out = (C.c_void_p * n)(*out_ptrs)
        a_cap = (C.c_size_t * n)(*out_caps)
        a_len = (C.c_size_t * n)()
        t = C.c_int()
        self._check(self._L.ccv2_submit_decode(self._h, n, a_in, a_n, a_out, a_cap, a_len, C.byref(t)))
        return Pending(self, t.value, (a_in, a_n, a_out, a_cap), lambda: list(a_len))

    # ---- tile mode (BASELINE configs[3]): one frame -> one reference-format stream per spatial tile
    def split_tiles(self, cloud, tile_bits):
        """Stable partition of a frame into 2^tile_bits tiles: (points grouped by tile, offsets[2^tile_bits + 1])."""
        a = np.ascontiguousarray(cloud)
        n = a.nbytes // 32
        out = np.zeros((max(n, 1), 32), np.uint8)
        offs = (C.c_size_t * ((1 << tile_bits) + 1))()
        self._check(self._L.ccv2_split_tiles(self._h, a.ctypes.data if n else None, n, tile_bits, out.ctypes.data, offs))
        return out[:n], list(offs)

    def encode_tiles_raw(self, ptr, n, tile_bits, out_ptrs, out_caps, first_tile=0, tile_step=1):
        nt = 1 << tile_bits
        a_out = (C.c_void_p * nt)(*out_ptrs)
        a_cap = (C.c_size_t * nt)(*out_caps)
        a_len = (C.c_size_t * nt)()
        a_np = (C.c_size_t * nt)()
        self._check(self._L.ccv2_encode_tiles(self._h, ptr, n, tile_bits, first_tile, tile_step, a_out, a_cap, a_len, a_np))
        return list(a_len), list(a_np)

    def encode_tiles(self, cloud, tile_bits, first_tile=0, tile_step=1):
        """Streams of the tiles first_tile, first_tile + tile_step, ... of one frame: ({tile: bytes}, points per tile)."""
        a = np.ascontiguousarray(cloud)
        n = a.nbytes // 32
        nt = 1 << tile_bits
        cap = min(self._L.ccv2_max_compressed_size(n), 6 * n + (1 << 16))
        bufs = [np.empty(cap, np.uint8) if (t >= first_tile and (t - first_tile) % tile_step == 0) else None for t in range(nt)]
        lens, npts = self.encode_tiles_raw(a.ctypes.data if n else None, n, tile_bits, [b.ctypes.data if b is not None else None for b in bufs], [cap] * nt, first_tile, tile_step)
        return {t: bufs[t][:lens[t]].tobytes() for t in range(nt) if bufs[t] is not None and lens[t]}, npts

    def quality_metrics(self, cloud_a, cloud_b):
        """computeQualityMetric(original, decoded) -> Quality (quality_metrics_impl.hpp:82-239)."""
        a, b = np.ascontiguousarray(cloud_a), np.ascontiguousarray(cloud_b)
        q = Quality()
        self._check(self._L.ccv2_quality_metrics(self._h, a.ctypes.data if a.size else None, a.nbytes // 32, b.ctypes.data if b.size else None, b.nbytes // 32, C.byref(q)))
        return q

    # ---- inter-frame (predictive) coding: encodePointCloudDeltaFrame / decodePointCloudDeltaFrame (impl.hpp:787-1235)
    def encode_delta_raw(self, i_ptr, ni, p_ptr, np_, i_out, i_cap, p_out, p_cap, icp_on_original=False, out_ptr=None, out_cap=0):
        """Pointers may be host or device addresses.  Returns (i_len, p_len, n_out, DeltaInfo)."""
        il, pl, no = C.c_size_t(), C.c_size_t(), C.c_size_t()
        info = DeltaInfo()
        self._check(self._L.ccv2_encode_delta(self._h, i_ptr, ni, p_ptr, np_, int(icp_on_original), i_out, i_cap, C.byref(il),
                                              p_out, p_cap, C.byref(pl), out_ptr, out_cap, C.byref(no), C.byref(info)))
        return il.value, pl.value, no.value, info

    def encode_delta(self, icloud, pcloud, icp_on_original=False, want_out_cloud=False):
        """-> (i_stream, p_stream, DeltaInfo[, predicted frame])."""
        ic, pc = np.ascontiguousarray(icloud), np.ascontiguousarray(pcloud)
        ni, npn = ic.nbytes // 32, pc.nbytes // 32
        icap = min(self._L.ccv2_max_compressed_size(npn), 6 * npn + (1 << 16))
        pcap = self._L.ccv2_max_p_stream_size(npn)
        ib, pb = np.empty(icap, np.uint8), np.empty(pcap, np.uint8)
        ob = np.zeros((ni + npn + 1, 32), np.uint8) if want_out_cloud else None
        il, pl, no, info = self.encode_delta_raw(ic.ctypes.data if ni else None, ni, pc.ctypes.data if npn else None, npn, ib.ctypes.data, icap,
                                                 pb.ctypes.data, pcap, icp_on_original, ob.ctypes.data if want_out_cloud else None, ni + npn + 1 if want_out_cloud else 0)
        if want_out_cloud:
            return ib[:il].tobytes(), pb[:pl].tobytes(), info, ob[:no]
        return ib[:il].tobytes(), pb[:pl].tobytes(), info

    def encode_delta_batch_raw(self, i_ptrs, nis, p_ptrs, nps, i_outs, i_caps, p_outs, p_caps, icp_on_original=False):
        """n delta frames in one call (the intra parts as one pipelined batch).  Returns (i_lens, p_lens, [DeltaInfo])."""
        n = len(i_ptrs)
        A = lambda v: (C.c_void_p * n)(*v)
        Z = lambda v: (C.c_size_t * n)(*v)
        il, pl, info = (C.c_size_t * n)(), (C.c_size_t * n)(), (DeltaInfo * n)()
        self._check(self._L.ccv2_encode_delta_batch(self._h, n, A(i_ptrs), Z(nis), A(p_ptrs), Z(nps), int(icp_on_original), A(i_outs), Z(i_caps), il, A(p_outs), Z(p_caps), pl, info))
        return list(il), list(pl), list(info)

    def decode_delta_batch_raw(self, i_ptrs, nis, is_ptrs, is_lens, ps_ptrs, ps_lens, out_ptrs, out_caps):
        n = len(i_ptrs)
        A = lambda v: (C.c_void_p * n)(*v)
        Z = lambda v: (C.c_size_t * n)(*v)
        npts, nb = (C.c_size_t * n)(), (C.c_uint64 * n)()
        self._check(self._L.ccv2_decode_delta_batch(self._h, n, A(i_ptrs), Z(nis), A(is_ptrs), Z(is_lens), A(ps_ptrs), Z(ps_lens), A(out_ptrs), Z(out_caps), npts, nb))
        return list(npts), list(nb)

    def decode_delta_raw(self, i_ptr, ni, is_ptr, is_len, ps_ptr, ps_len, out_ptr, out_cap):
        n, nb = C.c_size_t(), C.c_uint64()
        self._check(self._L.ccv2_decode_delta(self._h, i_ptr, ni, is_ptr, is_len, ps_ptr, ps_len, out_ptr, out_cap, C.byref(n), C.byref(nb)))
        return n.value, nb.value

    def decode_delta(self, icloud, i_stream, p_stream, cap_points=None):
        """-> ((n, 32) uint8 records, decoded macroblocks)."""
        ic = np.ascontiguousarray(icloud)
        ni = ic.nbytes // 32
        a, b = np.frombuffer(i_stream, np.uint8), np.frombuffer(p_stream, np.uint8)
        if cap_points is None:
            cnt = C.c_uint64(0)
            if a.size and self._L.ccv2_peek_point_count(a.ctypes.data, a.size, C.byref(cnt)):
                cnt = C.c_uint64(0)
            cap_points = ni * max(1, b.size // 19 // max(1, ni) + 1) + cnt.value + 1
        out = np.zeros((cap_points, 32), np.uint8)
        n, nb = self.decode_delta_raw(ic.ctypes.data if ni else None, ni, a.ctypes.data if a.size else None, a.size,
                                      b.ctypes.data if b.size else None, b.size, out.ctypes.data, cap_points)
        return out[:n], nb

    def simplify(self, cloud):
        """simplifyPCloud (impl.hpp:318-400) -> (V, 32) uint8 records."""
        a = np.ascontiguousarray(cloud)
        n = a.nbytes // 32
        out = np.zeros((max(n, 1), 32), np.uint8)
        v = C.c_size_t()
        self._check(self._L.ccv2_simplify(self._h, a.ctypes.data if n else None, n, out.ctypes.data, out.shape[0], C.byref(v)))
        return out[:v.value]

    def timer_start(self):
        self._check(self._L.ccv2_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._L.ccv2_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    # ---- numpy convenience API
    def encode_batch(self, clouds):
        """clouds: list of arrays of 32-byte PointXYZRGB records (any dtype, nbytes % 32 == 0). Returns list of bytes."""
        ins, ns, outs, caps, keep = [], [], [], [], []
        for cl in clouds:
            a = np.ascontiguousarray(cl)
            if a.nbytes % 32:
                raise ValueError("cloud bytes must be a multiple of 32")
            n = a.nbytes // 32
            cap = min(self._L.ccv2_max_compressed_size(n), 6 * n + (1 << 16))
            o = np.empty(cap, np.uint8)
            keep.append((a, o))
            ins.append(a.ctypes.data if n else None)
            ns.append(n)
            outs.append(o.ctypes.data)
            caps.append(cap)
        lens = self.encode_batch_raw(ins, ns, outs, caps)
        return [keep[i][1][:lens[i]].tobytes() for i in range(len(clouds))]

    def decode_batch(self, streams):
        """streams: list of bytes. Returns list of (n, 32) uint8 arrays (PointXYZRGB records)."""
        ins, lens, outs, caps, keep = [], [], [], [], []
        for s in streams:
            a = np.frombuffer(s, np.uint8)
            cnt = C.c_uint64(0)
            rc = self._L.ccv2_peek_point_count(a.ctypes.data if a.size else None, a.size, C.byref(cnt))
            if rc:
                raise Ccv2Error(rc, "not a cloud_codec_v2 frame")
            o = np.zeros((max(1, cnt.value), 32), np.uint8)
            keep.append((a, o))
            ins.append(a.ctypes.data)
            lens.append(a.size)
            outs.append(o.ctypes.data)
            caps.append(max(1, cnt.value))
        ns = self.decode_batch_raw(ins, lens, outs, caps)
        return [keep[i][1][:ns[i]] for i in range(len(streams))]

    # ---- accessors
    def metrics(self):
        m = (C.c_uint64 * 3)()
        self._check(self._L.ccv2_get_metrics(self._h, m))
        return list(m)

    @property
    def frame_id(self):
        return self._L.ccv2_get_frame_id(self._h)

    @frame_id.setter
    def frame_id(self, v):
        self._check(self._L.ccv2_set_frame_id(self._h, v))

    @property
    def last_launch_count(self):
        return int(self._L.ccv2_last_launch_count(self._h))

    @property
    def last_device_ms(self):
        return float(self._L.ccv2_last_device_ms(self._h))

    def set_profiling(self, on):
        self._check(self._L.ccv2_set_profiling(self._h, int(on)))

    def profile(self):
        """[(kernel name, total ms, launches)] of the last batch call made with profiling on."""
        out, i = [], 0
        while True:
            name, ms, n = C.c_char_p(), C.c_float(), C.c_int()
            if self._L.ccv2_get_profile(self._h, i, C.byref(name), C.byref(ms), C.byref(n)):
                return out
            out.append((name.value.decode(), float(ms.value), int(n.value)))
            i += 1

    def output_cloud(self, frame=0):
        """[PCL] getOutputCloud() for frame `frame` of the last encode_batch: (V, 32) uint8 PointXYZRGB records."""
        n = C.c_size_t()
        rc = self._L.ccv2_get_output_cloud(self._h, frame, None, 0, C.byref(n))
        if rc not in (0, -4):
            self._check(rc)
        out = np.zeros((max(1, n.value), 32), np.uint8)
        self._check(self._L.ccv2_get_output_cloud(self._h, frame, out.ctypes.data, out.shape[0], C.byref(n)))
        return out[:n.value]

    def debug_fetch(self, frame, what):
        """Test hook: 0 leaf codes (u64), 1 tree bytes, 2 avg colours, 3 colour payload, 4 sorted indices, 5 info."""
        ln = C.c_size_t()
        if what == 5:
            info = FrameInfo()
            self._check(self._L.ccv2_debug_fetch(self._h, frame, 5, C.byref(info), C.sizeof(info), C.byref(ln)))
            return info
        rc = self._L.ccv2_debug_fetch(self._h, frame, what, None, 0, C.byref(ln))
        if rc not in (0, -4):
            self._check(rc)
        buf = np.zeros(max(1, ln.value), np.uint8)
        self._check(self._L.ccv2_debug_fetch(self._h, frame, what, buf.ctypes.data, buf.size, C.byref(ln)))
        buf = buf[:ln.value]
        if what == 0:
            return buf.view(np.uint64)
        if what == 4:
            return buf.view(np.uint32)
        return buf


def strip_chunk_sizes(p_stream):
    """The P stream generatePointCloudDeltaFrame writes (impl.hpp:650-660): the same chunks as encodePointCloudDeltaFrame's
    (impl.hpp:877-883) WITHOUT their leading size byte.  Host-side re-framing of a finished stream."""
    out, pos, n = bytearray(), 0, len(p_stream)
    while pos < n:
        size = p_stream[pos]
        if size == 0 or pos + 1 + size > n:
            raise ValueError("malformed P stream")
        out += p_stream[pos + 1:pos + 1 + size]
        pos += 1 + size
    return bytes(out)


class OctreePointCloudCodecV2:
    """Python mirror of pcl::io::OctreePointCloudCodecV2<PointXYZRGB> (codec.h:70-368) over the C ABI.

    Same constructor argument order and meaning as codec.h:108-143; ``encodePointCloud`` returns the bytes the
    reference writes to its ostream and ``decodePointCloud`` returns the decoded cloud (an (n, 32) uint8 array
    of PointXYZRGB records).  Like the reference, the calls return nothing useful on an empty cloud / a stream
    without a frame header (impl.hpp:206-212, :231)."""

    def __init__(self, compressionProfile=MANUAL_CONFIGURATION, showStatistics=False, pointResolution=0.001,
                 octreeResolution=0.01, doVoxelGridDownDownSampling=False, iFrameRate=0, doColorEncoding=True,
                 colorBitResolution=6, colorCodingType=0, doVoxelGridCentroid=True, createScalableStream=True,
                 codeConnectivity=False, jpeg_quality=75, num_threads=0, device=0):
        p = Params()
        p.profile = compressionProfile
        p.show_statistics = int(showStatistics)
        p.point_resolution = pointResolution
        p.octree_resolution = octreeResolution
        p.do_voxel_grid_downsampling = int(doVoxelGridDownDownSampling)
        p.i_frame_rate = iFrameRate
        p.do_color_encoding = int(doColorEncoding)
        p.color_bit_resolution = colorBitResolution
        p.color_coding_type = colorCodingType
        p.do_voxel_grid_centroid = int(doVoxelGridCentroid)
        p.create_scalable_stream = int(createScalableStream)
        p.code_connectivity = int(codeConnectivity)
        p.jpeg_quality = jpeg_quality
        p.num_threads = num_threads
        p.macroblock_size = 16            # codec.h:138
        p.do_icp_color_offset = 0         # codec.h:141
        self._params, self._device = p, device
        self._codec = Codec(p, device)
        self._mb = (0.0, 0.0)

    def _reconfigure(self):               # the setters change state the C handle was created with
        fid = self._codec.frame_id
        self._codec.close()
        self._codec = Codec(self._params, self._device)
        self._codec.frame_id = fid

    def setMacroblockSize(self, size):    # codec.h:149-152
        self._params.macroblock_size = int(size); self._reconfigure()

    def setDoICPColorOffset(self, doit):  # codec.h:164-167 (the bool overload)
        self._params.do_icp_color_offset = int(bool(doit)); self._reconfigure()

    def encodePointCloudDeltaFrame(self, icloud, pcloud, icp_on_original=False, write_out_cloud=False):
        """codec.h:180-184: returns (i_coded_data, p_coded_data, out_cloud); out_cloud is empty unless write_out_cloud."""
        r = self._codec.encode_delta(icloud, pcloud, icp_on_original, write_out_cloud)
        info = r[2]
        self._mb = (info.shared_percentage, info.convergence_percentage)
        return r[0], r[1], (r[3] if write_out_cloud else np.zeros((0, 32), np.uint8))

    def generatePointCloudDeltaFrame(self, icloud, pcloud, icp_on_original=False, write_out_cloud=True):
        """codec.h:180-182, impl.hpp:577-786: the older form of the delta encoder -- the same prediction, chunks written without
        their size byte (no decoder of the reference reads that layout; kept for callers that store it).  Returns
        (i_coded_data, p_coded_data, out_cloud)."""
        i_s, p_s, oc = self.encodePointCloudDeltaFrame(icloud, pcloud, icp_on_original, write_out_cloud)
        return i_s, strip_chunk_sizes(p_s), oc

    def decodePointCloudDeltaFrame(self, icloud, i_coded_data, p_coded_data):
        """codec.h:186-190: returns the decoded frame."""
        return self._codec.decode_delta(icloud, i_coded_data, p_coded_data)[0]

    def getMacroBlockPercentage(self):            # codec.h:200-204
        return self._mb[0]

    def getMacroBlockConvergencePercentage(self):  # codec.h:206-210
        return self._mb[1]

    def encodePointCloud(self, cloud):
        return self._codec.encode_batch([cloud])[0]

    def decodePointCloud(self, data):
        try:
            return self._codec.decode_batch([data])[0]
        except Ccv2Error as e:
            if e.status == -6:            # sync failure => silent return (impl.hpp:231)
                return np.zeros((0, 32), np.uint8)
            raise

    def getPerformanceMetrics(self):
        return self._codec.metrics()

    def getOutputCloud(self):
        """The simplified cloud of the last encodePointCloud call ([PCL] getOutputCloud, eval.hpp:862)."""
        return self._codec.output_cloud(0)


def profile_step(clouds, octree_bits=11, device=0):
    """One encode+decode of `clouds` as a single group on one stream with CUDA events around every kernel.
    Returns [(kernel, total_ms, launches, frames_per_launch)] (bench.py's roofline leg)."""
    c = Codec(default_params(octree_bits=octree_bits), device)
    try:
        c.encode_batch(clouds)                      # warm-up (allocations)
        profile_step.depth = int(c.debug_fetch(0, 5).depth)   # realised octree depth of the first frame (sort passes = ceil((3 depth + 1) / 8))
        c.set_profiling(True)
        streams = c.encode_batch(clouds)
        prof = c.profile()
        c.decode_batch(streams)
        prof += c.profile()
        return [(n, ms, k, len(clouds)) for n, ms, k in prof]
    finally:
        c.close()

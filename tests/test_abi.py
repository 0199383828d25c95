"""CPU tests of the drop-in boundary: libccv2.so loads, exports every symbol include/ccv2.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torch_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def K():
    import __graft_entry__ as g
    from cwi_pcl_codec_b200 import codec
    if not os.path.exists(codec.LIB_PATH):
        g.build()
    return codec


def test_every_declared_symbol_is_exported(K):
    hdr = open(os.path.join(ROOT, "include", "ccv2.h")).read()
    declared = sorted(set(re.findall(r"\b(ccv2_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = C.CDLL(K.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libccv2.so does not export %s" % name
    assert sorted(K.EXPORTED_SYMBOLS) == declared


def test_params_struct_matches_header_layout(K):
    p = K.default_params()
    assert C.sizeof(K.Params) == 72
    assert p.profile == K.MANUAL_CONFIGURATION and p.octree_resolution == 2.0 ** -11 and p.point_resolution == 2.0 ** -11
    assert p.do_voxel_grid_downsampling == 1 and p.color_coding_type == 1 and p.jpeg_quality == 85 and p.macroblock_size == 16
    q = K.default_params(octree_bits=8, color_bits=0, keep_centroid=1)
    assert q.octree_resolution == 2.0 ** -8 and q.do_color_encoding == 0 and q.do_voxel_grid_centroid == 1


def test_no_cpu_fallback_without_a_device(K):
    if _torch_cuda():
        pytest.skip("a CUDA device is present")
    with pytest.raises(K.Ccv2Error) as ei:
        K.Codec(K.default_params())
    assert ei.value.status == -2                                     # CCV2_ERR_CUDA


def test_unsupported_configurations_are_rejected_before_touching_cuda(K):
    lib = K.load_library()
    h = C.c_void_p()
    for kw, want in [(dict(profile=3), -3), (dict(color_coding_type=7), -1), (dict(octree_resolution=0.0), -1), (dict(color_bit_resolution=9), -1)]:
        assert lib.ccv2_create(C.byref(K.default_params(**kw)), 0, C.byref(h)) == want
        assert lib.ccv2_last_error(None)


def test_peek_point_count_and_sizes(K, oracle):
    from cwi_pcl_codec_b200 import synth
    data, info = oracle.encode(synth.gen_surface(3000, 2), oracle.default_params(octree_bits=7))
    lib = K.load_library()
    cnt = C.c_uint64()
    buf = np.frombuffer(data, np.uint8)
    assert lib.ccv2_peek_point_count(buf.ctypes.data, buf.size, C.byref(cnt)) == 0 and cnt.value == info.n_leaves
    junk = np.zeros(200, np.uint8)
    assert lib.ccv2_peek_point_count(junk.ctypes.data, junk.size, C.byref(cnt)) == -6
    assert lib.ccv2_max_compressed_size(1000000) > lib.ccv2_max_compressed_size(1000) > 2200
    assert lib.ccv2_status_string(-4) == b"output buffer too small"

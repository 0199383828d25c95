"""The C++ host side of the boundary: cwi_pcl_codec_b200/host/pcl/cloud_codec_v2/point_cloud_codec_v2.h mirrors the
reference class over the C ABI; evaluate_compression_mini is a small harness built on it."""
import os
import subprocess

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cwi_pcl_codec_b200", "host", "evaluate_compression_mini")


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    if not os.path.exists(EXE):
        g.build()
    return EXE


def test_harness_builds_and_prints_usage(exe):
    out = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--octree_bits" in out.stdout and "--jpeg_quality" in out.stdout


def test_harness_fails_loudly_without_a_device(exe):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    out = subprocess.run([exe, "--synthetic", "1000"], capture_output=True, text=True)
    assert out.returncode == 1 and "no usable CUDA device" in out.stderr


@pytest.mark.gpu
def test_cpp_facade_round_trip_is_bit_exact(exe, oracle, tmp_path):
    cloud = synth.gen_surface(30000, 77)
    src = tmp_path / "cloud.xyzrgb32"
    cloud.tofile(src)
    stream, dec = tmp_path / "s.bin", tmp_path / "d.bin"
    out = subprocess.run([exe, "--input", str(src), "--octree_bits", "9", "--jpeg_quality", "85", "--frames", "2",
                          "--dump_stream", str(stream), "--dump_cloud", str(dec)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = out.stdout.strip().splitlines()
    assert rows[0].startswith("frame;points;compressed_byte_size") and len(rows) == 3
    ref, info = oracle.encode(cloud, oracle.default_params(octree_bits=9), frame_id=1)
    assert stream.read_bytes() == ref
    rd, _ = oracle.decode(ref)
    assert np.array_equal(np.fromfile(dec, np.uint8).reshape(-1, 32), rd)
    f0 = rows[1].split(";")
    assert int(f0[1]) == 30000 and int(f0[2]) == len(ref) and int(f0[6]) == info.n_leaves
    assert int(f0[3]) == info.coded[0] and int(f0[5]) == info.coded[2]

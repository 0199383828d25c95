"""evaluate_compression (B200 harness): the reference tool's option surface, file formats, bounding-box normalisation and
CSV (apps/evaluate_compression/.../evaluate_compression_impl.hpp:137-169, 250-305, 683-897; quality_metrics_impl.hpp:242-285;
impl.hpp:1871-1986) -- CPU tests for everything around the codec, one GPU test for BASELINE configs[0] end to end."""
import os
import struct
import subprocess

import numpy as np
import pytest

from cwi_pcl_codec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cwi_pcl_codec_b200", "host", "evaluate_compression")


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    if not os.path.exists(EXE):
        g.build()
    return EXE


def write_ply(path, cl, binary):
    n = cl.shape[0]
    hdr = ("ply\nformat %s 1.0\ncomment test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "property float nx\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n"
           % ("binary_little_endian" if binary else "ascii", n))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            rec = np.zeros(n, np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))
            for k in ("x", "y", "z", "r", "g", "b"):
                rec[k] = cl[k]
            f.write(rec.tobytes())
        else:
            for p in cl:
                f.write(("%.9g %.9g %.9g 0 %d %d %d\n" % (p["x"], p["y"], p["z"], p["r"], p["g"], p["b"])).encode())


def write_pcd(path, cl, binary):
    n = cl.shape[0]
    rgb = (cl["r"].astype(np.uint32) << 16) | (cl["g"].astype(np.uint32) << 8) | cl["b"].astype(np.uint32)
    hdr = "# .PCD v0.7\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, "binary" if binary else "ascii")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            rec = np.zeros(n, np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4")]))
            rec["x"], rec["y"], rec["z"], rec["rgb"] = cl["x"], cl["y"], cl["z"], rgb
            f.write(rec.tobytes())
        else:
            for i in range(n):
                f.write(("%.9g %.9g %.9g %d\n" % (cl["x"][i], cl["y"][i], cl["z"][i], rgb[i])).encode())


def read_ply_ascii(path):
    lines = open(path).read().splitlines()
    k = lines.index("end_header")
    a = np.array([[float(v) for v in ln.split()] for ln in lines[k + 1:] if ln.strip()])
    return a[:, :3].astype(np.float32), a[:, 3:6].astype(np.uint8)


def normalize_group(clouds, f):
    """normalize_pointclouds (impl.hpp:1871-1966) in numpy float32, the reference's operation order."""
    mn_bb = np.full(3, 1000, np.float32); mx_bb = np.full(3, -1000, np.float32)
    init = False
    out = []
    for xyz in clouds:
        mn, mx = xyz.min(0), xyz.max(0)
        if not (np.all(mn > mn_bb) and np.all(mx < mx_bb)):
            init = False
        if not init:
            ext = np.abs(mx - mn).astype(np.float32)
            mn_bb = (mn.astype(np.float64) - f * ext.astype(np.float64)).astype(np.float32)
            mx_bb = (mx.astype(np.float64) + f * ext.astype(np.float64)).astype(np.float32)
            init = True
        dyn = (mx_bb - mn_bb).astype(np.float32)
        out.append(((xyz - mn_bb).astype(np.float32) / dyn).astype(np.float32))
    return out, mn_bb, mx_bb


def restore(xyz, mn_bb, mx_bb):
    dyn = (mx_bb - mn_bb).astype(np.float32)
    return ((xyz * dyn).astype(np.float32) + mn_bb).astype(np.float32)


def raw_cloud(n, seed, scale=(3.0, 2.0, 5.0), shift=(-1.0, 4.0, 0.5)):
    cl = synth.gen_surface(n, seed)
    for a, k in enumerate("xyz"):
        cl[k] = (cl[k] * np.float32(scale[a]) + np.float32(shift[a])).astype(np.float32)
    return cl


def test_option_surface_matches_the_reference(exe):
    out = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for name in ("K_outlier_filter", "radius", "group_size", "bb_expand_factor", "algorithm", "input_directories", "output_directory", "show_statistics",
                 "visualization", "point_resolution", "octree_resolution", "octree_bits", "color_bits", "enh_bits", "color_coding_type", "macroblock_size",
                 "keep_centroid", "create_scalable", "do_connectivity_coding", "icp_on_original", "jpeg_quality", "do_delta_coding", "do_quality_computation",
                 "do_icp_color_offset", "num_threads", "intra_frame_quality_csv", "predictive_quality_csv", "debug_level"):
        assert "--" + name in out.stdout, name
    assert "--jpeg_quality -j (=0)" in out.stdout and "--octree_bits -b (=11)" in out.stdout        # eval.hpp:161, 152
    bad = subprocess.run([exe, "--no_such_option", "1", "/tmp"], capture_output=True, text=True)
    assert bad.returncode != 0 and "Unrecognized options on command line" in bad.stderr


def test_file_formats_normalisation_and_outlier_filter_without_a_gpu(exe, tmp_path):
    """Four input formats, a group that keeps its first bounding box, restore_scaling with the LAST box."""
    d = tmp_path / "in"; d.mkdir()
    a, b = raw_cloud(3000, 1), raw_cloud(2500, 2, scale=(2.5, 1.8, 4.5), shift=(-0.8, 4.1, 0.7))      # b fits a's expanded box
    c = raw_cloud(2000, 3, scale=(6.0, 2.0, 5.0))                                                     # c does not: the box is re-initialised
    e = raw_cloud(1500, 4)
    write_ply(d / "0001.ply", a, True); write_ply(d / "0002.ply", b, False); write_pcd(d / "0003.pcd", c, True); write_pcd(d / "0004.pcd", e, False)
    (d / "notes.txt").write_text("ignored")
    out = tmp_path / "out"
    r = subprocess.run([exe, "--skip_coding", "-o", str(out), "-f", "0.2", "--intra_frame_quality_csv", str(tmp_path / "i.csv"),
                        "--predictive_quality_csv", "", str(d)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    clouds = [np.stack([q["x"], q["y"], q["z"]], 1) for q in (a, b, c, e)]
    norm, mn_bb, mx_bb = normalize_group(clouds, 0.2)
    assert all(n.min() >= 0 and n.max() <= 1 for n in norm)
    for i, (src, q) in enumerate(zip(norm, (a, b, c, e))):
        xyz, rgb = read_ply_ascii(out / ("pointcloud_%d.ply" % i))
        assert np.array_equal(xyz, restore(src, mn_bb, mx_bb)), i                                     # every frame un-scaled with the last box
        assert np.array_equal(rgb, np.stack([q["r"], q["g"], q["b"]], 1))
    assert (tmp_path / "i.csv").read_text().startswith("compression setting; in point count;out point count;compressed_byte_size;")
    # outlier filter: an isolated point goes, dense points stay ([PCL] RadiusOutlierRemoval: more than K points within the radius)
    d2 = tmp_path / "in2"; d2.mkdir()
    g = synth.gen_surface(4000, 5)
    g["x"][7], g["y"][7], g["z"][7] = 9.0, 9.0, 9.0
    write_ply(d2 / "a.ply", g, True)
    out2 = tmp_path / "out2"
    r = subprocess.run([exe, "--skip_coding", "-o", str(out2), "-f", "0", "-K", "2", "--radius", "0.05", "--intra_frame_quality_csv", "", "--predictive_quality_csv", "", str(d2)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    xyz, _ = read_ply_ascii(out2 / "pointcloud_0.ply")
    src = np.stack([g["x"], g["y"], g["z"]], 1)
    d2m = ((src[:, None, :].astype(np.float32) - src[None, :, :].astype(np.float32)) ** 2).sum(-1)
    keep = (d2m <= np.float32(0.05 * 0.05)).sum(1) > 2
    assert not keep[7] and np.array_equal(xyz, src[keep])


def test_parameter_config_file_and_command_line_precedence(exe, tmp_path):
    d = tmp_path / "in"; d.mkdir()
    write_ply(d / "a.ply", raw_cloud(500, 6), False)
    (tmp_path / "parameter_config.txt").write_text("# like the reference's parameter_config.txt\noctree_bits=9\nbb_expand_factor = 0.5\noutput_directory=%s\n" % (tmp_path / "cfg_out"))
    r = subprocess.run([exe, "--skip_coding", "--debug_level", "1", "-b", "7", "--intra_frame_quality_csv", "", "--predictive_quality_csv", "", str(d)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert "octree_bits=7" in r.stdout and "bb_expand_factor=0.5" in r.stdout                        # the command line wins, the file fills the rest
    assert (tmp_path / "cfg_out" / "pointcloud_0.ply").exists()
    (tmp_path / "parameter_config.txt").write_text("not_an_option=1\n")
    r = subprocess.run([exe, "--skip_coding", str(d)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "Unrecognized options in configuration file" in r.stderr


@pytest.mark.gpu
def test_baseline_config_0_end_to_end(exe, oracle, tmp_path):
    """BASELINE.json configs[0]: one synthetic 10k-point XYZRGB frame from a PLY through the CLI clone, intra, octree_bits 8,
    JPEG Q85: the written cloud and the CSV row against the CPU oracle pipeline (numpy normalisation -> oracle encode /
    decode -> oracle quality metrics -> numpy restore)."""
    d = tmp_path / "in"; d.mkdir()
    cl = raw_cloud(10000, 0)
    write_ply(d / "frame_0000.ply", cl, True)
    out = tmp_path / "out"
    r = subprocess.run([exe, "-b", "8", "-t", "1", "-j", "85", "-q", "1", "-o", str(out), "--intra_frame_quality_csv", str(tmp_path / "q.csv"), "--predictive_quality_csv", "", str(d)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout
    (norm,), mn_bb, mx_bb = normalize_group([np.stack([cl["x"], cl["y"], cl["z"]], 1)], 0.2)
    ncl = cl.copy()
    ncl["x"], ncl["y"], ncl["z"] = norm[:, 0], norm[:, 1], norm[:, 2]
    ref, info = oracle.encode(ncl, oracle.default_params(octree_bits=8, jpeg_quality=85), frame_id=1)
    assert " octreeCoding %d bytes" % len(ref) in r.stdout
    rd, _ = oracle.decode(ref)
    xyz, rgb = read_ply_ascii(out / "pointcloud_0.ply")
    assert np.array_equal(xyz, restore(rd[:, :12].copy().view(np.float32).reshape(-1, 3), mn_bb, mx_bb))
    assert np.array_equal(rgb, rd[:, [18, 17, 16]])
    q = oracle.quality_metrics(ncl, rd)
    rows = (tmp_path / "q.csv").read_text().strip().splitlines()
    assert len(rows) == 2
    f = rows[1].split(";")
    assert f[0] == "octree_bits=8 color_bits=8 enh._bits=0_colortype=1 centroid=0"
    assert int(f[1]) == 10000 and int(f[2]) == rd.shape[0] and int(f[3]) == len(ref)
    assert abs(float(f[5]) - info.coded[0] / rd.shape[0]) < 1e-4 and abs(float(f[7]) - info.coded[2] / rd.shape[0]) < 1e-4
    assert abs(float(f[8]) - q.symm_rms) < 1e-6 * max(1, q.symm_rms) + 1e-9 and abs(float(f[9]) - q.symm_hausdorff) < 1e-6
    assert abs(float(f[10]) - q.psnr_db) < 1e-3
    for k in range(3):
        assert abs(float(f[11 + k]) - q.psnr_yuv[k]) < 1e-3


@pytest.mark.gpu
def test_delta_coding_through_the_cli_clone(exe, oracle, tmp_path):
    """do_delta_coding (eval.hpp:854-889): a group of three frames; frame i+1 is coded against the encoder's simplified
    cloud of frame i and decoded against the decoded frame i.  Stream sizes, the predictive CSV rows and the written
    delta_decoded_pc_<n>.ply against the same pipeline run through the CPU oracle."""
    d = tmp_path / "in"; d.mkdir()
    gof = synth.gen_gof(30000, seed=3, frames=3)
    raw = []
    for k, cl in enumerate(gof):
        c = cl.copy()
        for a, nm in enumerate("xyz"):
            c[nm] = (c[nm] * np.float32(2.0) + np.float32(a - 1.0)).astype(np.float32)
        raw.append(c)
        write_ply(d / ("frame_%04d.ply" % k), c, True)
    out = tmp_path / "out"
    r = subprocess.run([exe, "-b", "9", "-t", "1", "-j", "85", "-q", "1", "-d", "1", "-g", "3", "-o", str(out), "--intra_frame_quality_csv", str(tmp_path / "i.csv"),
                        "--predictive_quality_csv", str(tmp_path / "p.csv"), str(d)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout
    norm, mn_bb, mx_bb = normalize_group([np.stack([c["x"], c["y"], c["z"]], 1) for c in raw], 0.2)
    ncl = []
    for c, nx in zip(raw, norm):
        q = c.copy(); q["x"], q["y"], q["z"] = nx[:, 0], nx[:, 1], nx[:, 2]; ncl.append(q)
    op = oracle.default_params(octree_bits=9, jpeg_quality=85)
    prows = (tmp_path / "p.csv").read_text().strip().splitlines()
    assert len(prows) == 3                                                                             # header + two predicted frames
    for i in range(2):
        ref, info, dbg = oracle.encode(ncl[i], op, frame_id=i + 1, debug=True)
        rd, _ = oracle.decode(ref)
        ri, rp, rinfo = oracle.encode_delta(dbg["output_cloud"], ncl[i + 1], op)
        assert " encoded a predictive frame: coded %d bytes intra and %d inter frame encoded " % (len(ri), len(rp)) in r.stdout
        pdec, _ = oracle.decode_delta(rd, ri, rp, op)
        q = oracle.quality_metrics(ncl[i + 1], pdec)
        f = prows[1 + i].split(";")
        assert int(f[1]) == 30000 and int(f[2]) == pdec.shape[0] and int(f[3]) == len(ri) + len(rp)
        assert abs(float(f[5]) - len(ri) / pdec.shape[0]) < 1e-4 and abs(float(f[6]) - len(rp) / pdec.shape[0]) < 1e-4 and float(f[7]) == 0
        assert abs(float(f[8]) - q.symm_rms) < 1e-6 and abs(float(f[10]) - q.psnr_db) < 1e-3 and abs(float(f[11]) - q.psnr_yuv[0]) < 1e-3
        xyz, rgb = read_ply_ascii(out / ("delta_decoded_pc_%d.ply" % (i + 1)))
        assert np.array_equal(xyz, restore(pdec[:, :12].copy().view(np.float32).reshape(-1, 3), mn_bb, mx_bb))
        assert np.array_equal(rgb, pdec[:, [18, 17, 16]])
    assert len((tmp_path / "i.csv").read_text().strip().splitlines()) == 4

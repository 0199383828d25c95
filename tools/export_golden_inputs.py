"""Writes the frozen golden inputs (tests/golden/stream_inputs.npz) as raw files for tools/real_pcl_dump.cpp:
<case>.bin = n x 32-byte PointXYZRGB records, <case>.args = the constructor arguments of OctreePointCloudCodecV2 in order
(point_cloud_codec_v2.h:108-143) + macroblock_size + do_icp_color_offset.   usage: python tools/export_golden_inputs.py OUTDIR"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402


def ctor_args(p):
    bits, enh = p.get("octree_bits"), p.get("enh_bits", 0)
    ores = p.get("octree_resolution", 2.0 ** -bits if bits is not None else 2.0 ** -11)
    pres = p.get("point_resolution", 2.0 ** -(bits + enh) if bits is not None else ores)
    color_bits = p.get("color_bit_resolution", 8)
    do_color = p.get("do_color", 1 if color_bits > 0 else 0)
    return [13, 0, repr(pres), repr(ores), p.get("do_voxel_grid", 1), 0, do_color, color_bits, p.get("color_coding_type", 1),
            p.get("do_centroid", 0), p.get("create_scalable", 0), 0, p.get("jpeg_quality", 85), 1, 16, 0]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_hashes.json")))
    for name in cases.frozen_names():
        cases.load_case(name).tofile(os.path.join(outdir, name + ".bin"))
        open(os.path.join(outdir, name + ".args"), "w").write(" ".join(str(a) for a in ctor_args(table[name]["params"])) + "\n")
        print(name)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "golden_export")

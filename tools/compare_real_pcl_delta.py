"""The inter-frame half of the real-PCL comparison (tools/diff_against_real_pcl.md, last section).

  python tools/compare_real_pcl_delta.py export DIR    frozen frame pair + constructor arguments + OUR (oracle) streams per case
  (run tools/real_pcl_delta_dump on every case)
  python tools/compare_real_pcl_delta.py check DIR     verdict per case

Bit-exact expectations: the reference DECODER on our streams equals our decode (no registration in the decoder); the
reference's I cloud equals ours; its shared-macroblock percentage equals ours.  Not bit-exact by nature (ICP rounding
differs between PCL / Eigen builds): which shared blocks are predicted and the transforms' last bits -- compared through
block counts, stream sizes and the quality of the decoded frame."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("make_delta_golden", os.path.join(ROOT, "tests", "golden", "make_delta_golden.py"))
G = importlib.util.module_from_spec(spec); spec.loader.exec_module(G)


def params_of(kw):
    kw = dict(kw)
    orig = int(kw.pop("_icp_on_original", 0))
    return O.default_params(**kw), orig


def export(d):
    os.makedirs(d, exist_ok=True)
    frames = G.load_inputs()
    for name, kw in G.DELTA_CASES.items():
        p, orig = params_of(kw)
        base = os.path.join(d, name)
        frames[0].tofile(base + ".f0.bin"); frames[1].tofile(base + ".f1.bin")
        args = [13, 0, repr(p.point_resolution), repr(p.octree_resolution), 1, 0, 1, p.color_bit_resolution, p.color_coding_type, p.do_centroid, p.create_scalable, 0,
                p.jpeg_quality, 1, p.macroblock_size, p.do_icp_color_offset, orig]
        open(base + ".args", "w").write(" ".join(str(a) for a in args) + "\n")
        _, _, dbg = O.encode(frames[0], p, debug=True)
        ic = frames[0] if orig else dbg["output_cloud"]
        i_s, p_s, info = O.encode_delta(ic, frames[1], p, icp_on_original=bool(orig))
        open(base + ".our_istream", "wb").write(i_s); open(base + ".our_pstream", "wb").write(p_s)
        print(name, len(i_s), len(p_s), info.shared_percentage, info.convergence_percentage)


def check(d):
    frames = G.load_inputs()
    bad = seen = 0
    for name, kw in G.DELTA_CASES.items():
        p, orig = params_of(kw)
        base = os.path.join(d, name)
        if not os.path.exists(base + ".ref_pstream"):
            print("%-24s (no reference output: run tools/real_pcl_delta_dump %s)" % (name, base)); continue
        seen += 1
        _, _, dbg = O.encode(frames[0], p, debug=True)
        ic = frames[0] if orig else dbg["output_cloud"]
        ref_ic = np.fromfile(base + ".icloud", np.uint8).reshape(-1, 32)
        same_ic = ref_ic.shape == ic.shape and np.array_equal(ref_ic[:, :20], ic[:, :20])          # bytes 20..31 are padding PCL does not define
        i_s, p_s = open(base + ".our_istream", "rb").read(), open(base + ".our_pstream", "rb").read()
        ours, _ = O.decode_delta(ic, i_s, p_s, p)
        theirs = np.fromfile(base + ".ref_decoded_ours", np.uint8).reshape(-1, 32)
        same_dec = theirs.shape == ours.shape and np.array_equal(theirs[:, :20], ours[:, :20])
        info = O.encode_delta(ic, frames[1], p, icp_on_original=bool(orig))[2]
        rs, rc = [float(v) for v in open(base + ".stats").read().split()]
        ref_dec = np.fromfile(base + ".ref_decoded", np.uint8).reshape(-1, 32)
        q_ref, q_our = O.quality_metrics(frames[1], ref_dec), O.quality_metrics(frames[1], ours)
        ri, rp = os.path.getsize(base + ".ref_istream"), os.path.getsize(base + ".ref_pstream")
        ok = same_ic and same_dec and abs(rs - info.shared_percentage) < 1e-6 and abs(q_ref.psnr_db - q_our.psnr_db) < 0.5 and abs(q_ref.psnr_yuv[0] - q_our.psnr_yuv[0]) < 0.5
        bad += not ok
        print("%-24s I cloud %s | reference decoder on our streams %s | shared %.4f vs %.4f | predicted %.4f vs %.4f | bytes %d+%d vs %d+%d | PSNR %.2f / Y %.2f vs %.2f / %.2f" % (
            name, "MATCH" if same_ic else "DIFFERS", "MATCH" if same_dec else "DIFFERS", rs, info.shared_percentage, rc, info.convergence_percentage,
            ri, rp, len(i_s), len(p_s), q_ref.psnr_db, q_ref.psnr_yuv[0], q_our.psnr_db, q_our.psnr_yuv[0]))
    if not seen:
        print("nothing to check")
        return 2
    print("inter-frame path: format and quality agree" if not bad else "%d case(s) outside the stated agreement" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("export", "check"):
        sys.exit(__doc__)
    sys.exit(export(sys.argv[2]) if sys.argv[1] == "export" else check(sys.argv[2]))

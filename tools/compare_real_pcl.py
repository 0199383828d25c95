"""Compares what tools/real_pcl_dump wrote (<case>.stream, <case>.decoded) with the oracle's committed SHA-256
(tests/golden/stream_hashes.json).   usage: python tools/compare_real_pcl.py DIR"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(d):
    table = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_hashes.json")))
    bad = seen = 0
    for name, e in sorted(table.items()):
        sp, dp = os.path.join(d, name + ".stream"), os.path.join(d, name + ".decoded")
        if not os.path.exists(sp):
            continue
        s = open(sp, "rb").read()
        seen += 1
        ok_s = hashlib.sha256(s).hexdigest() == e["stream_sha256"]
        ok_d = None
        if os.path.exists(dp):
            ok_d = hashlib.sha256(open(dp, "rb").read()).hexdigest() == e["decoded_sha256"]
        print("%-40s stream %s (%d vs %d bytes)  decoded %s" % (name, "MATCH" if ok_s else "DIFFERS", len(s), e["stream_bytes"], {True: "MATCH", False: "DIFFERS", None: "-"}[ok_d]))
        if not ok_s:                                        # first differing byte and the layer it falls in
            bad += 1
    if not seen:
        print("no <case>.stream files in %s: run tools/real_pcl_dump first" % d)
        return 2
    print("pinned" if bad == 0 else "%d case(s) differ: see tools/diff_against_real_pcl.md for how to localise the layer" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "golden_export"))

"""Developer tool (GPU box): config-4-like sizes (4M points, 12 bits, JPEG quality sweep) against the oracle."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
from oracle import oracle as O
pts = synth.gen_surface(4000000, 0)
bad = 0
for q in (60, 85, 95):
    c = K.Codec(K.default_params(octree_bits=12, jpeg_quality=q))
    t = time.time(); s = c.encode_batch([pts])[0]; te = time.time() - t
    ref, info = O.encode(pts, O.default_params(octree_bits=12, jpeg_quality=q), frame_id=1)
    t = time.time(); d = c.decode_batch([s])[0]; td = time.time() - t
    rd, _ = O.decode(ref)
    ok = (s == ref) and np.array_equal(d, rd)
    bad += not ok
    print("Q%d: depth %d V %d B %d J %d S %d | stream %s decode %s | enc %.0f ms dec %.0f ms" % (q, info.depth, info.n_leaves, info.n_tree_bytes, info.n_color_bytes, len(ref), s == ref, np.array_equal(d, rd), te * 1e3, td * 1e3))
    c.close()
u = synth.gen_uniform(2000000, 3)
c = K.Codec(K.default_params(octree_bits=12))
s = c.encode_batch([u, pts[:1500000]])
for i, cl in enumerate((u, pts[:1500000])):
    ref, _ = O.encode(cl, O.default_params(octree_bits=12), frame_id=i + 1)
    ok = s[i] == ref
    bad += not ok
    print("mixed batch frame", i, ok)
print("FAILED" if bad else "ALL OK")

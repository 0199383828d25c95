"""Developer tool (GPU box), used under ncu: one warm-up and one measured delta-frame encode + decode of a 1M-point frame
against its predecessor (synth.gen_gof, octree_bits 11).  usage: prof_inter.py [N]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
g = [np.ascontiguousarray(c).view(np.uint8).reshape(-1, 32) for c in synth.gen_gof(n, seed=0, frames=2)]
c = K.Codec(K.default_params(octree_bits=11))
c.encode_batch([g[0]])
ic = c.output_cloud(0)
for r in range(2):
    i_s, p_s, info = c.encode_delta(ic, g[1])
    dec, nb = c.decode_delta(ic, i_s, p_s)
    print("rep %d: %d macroblocks, %d shared, %d predicted, %d points intra; %d + %d bytes; predict %.2f ms, intra coder %.2f ms; decoded %d points" % (
        r, info.macro_blocks, info.shared_blocks, info.converged_blocks, info.n_intra_points, len(i_s), len(p_s), info.predict_ms, info.intra_ms, dec.shape[0]))

"""Developer tool (GPU box): encode+decode F frames of N points once or twice; prints device times. Used under ncu."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see csrc/ccv2_api.cu (stream -> hardware queue aliasing)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kind = sys.argv[4] if len(sys.argv) > 4 else "surf"
gen = synth.gen_surface if kind == "surf" else synth.gen_uniform
clouds = [gen(n, s) for s in range(F)]
c = K.Codec(K.default_params(octree_bits=11))
for r in range(reps):
    t = time.time(); s = c.encode_batch(clouds); te = time.time() - t; de = c.last_device_ms
    t = time.time(); d = c.decode_batch(s); td = time.time() - t; dd = c.last_device_ms
    print("rep %d: encode wall %.1f ms dev %.1f ms | decode wall %.1f ms dev %.1f ms | %.2f Mpts/s (device enc+dec)" % (r, te * 1e3, de, td * 1e3, dd, n * F / (de + dd) / 1e3))

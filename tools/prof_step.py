"""Developer tool (GPU box), used under ncu: one warm-up and one measured encode + decode of F frames of N points as ONE
group on one stream (CCV2_GROUP=F), so every kernel launch covers F frames.  usage: prof_step.py N F [kind] [lps_dec]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kind = sys.argv[3] if len(sys.argv) > 3 else "surf"
os.environ["CCV2_GROUP"] = str(F)
if len(sys.argv) > 4:
    os.environ["CCV2_LPS_DEC"] = sys.argv[4]
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
gen = synth.gen_surface if kind == "surf" else synth.gen_uniform
dev = torch.device("cuda", 0)
base = [gen(n, s) for s in range(min(F, 8))]
d_in = [torch.from_numpy(base[i % len(base)].view(np.uint8).reshape(-1)).to(dev) for i in range(F)]
cap = 4 * n + (1 << 16)
d_str = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(F)]
d_out = [torch.empty(n * 32, dtype=torch.uint8, device=dev) for _ in range(F)]
c = K.Codec(K.default_params(octree_bits=11))
ip = [t.data_ptr() for t in d_in]; sp = [t.data_ptr() for t in d_str]; op = [t.data_ptr() for t in d_out]
for r in range(2):
    lens = c.encode_batch_raw(ip, [n] * F, sp, [cap] * F); de = c.last_device_ms
    ns = c.decode_batch_raw(sp, lens, op, [n] * F); dd = c.last_device_ms
    print("rep %d: encode %.1f ms, decode %.1f ms (device), %d frames of %d points, stream %.0f B, %d voxels" % (r, de, dd, F, n, np.mean(lens), ns[0]))

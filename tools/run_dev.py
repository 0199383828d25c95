"""Developer tool (GPU box): device-resident encode+decode of F frames; prints device ms of each call."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see csrc/ccv2_api.cu (stream -> hardware queue aliasing)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
n = int(sys.argv[1]); F = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
base = [synth.gen_surface(n, s) for s in range(min(F, 8))]
dev = torch.device("cuda", 0)
d_in = [torch.from_numpy(base[i % len(base)].view(np.uint8).reshape(-1)).to(dev) for i in range(F)]
cap = 4 * n + (1 << 16)
d_str = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(F)]
d_out = [torch.empty(n * 32, dtype=torch.uint8, device=dev) for _ in range(F)]
c = K.Codec(K.default_params(octree_bits=11))
ip = [t.data_ptr() for t in d_in]; sp = [t.data_ptr() for t in d_str]; op = [t.data_ptr() for t in d_out]
for r in range(reps):
    t = time.time(); lens = c.encode_batch_raw(ip, [n] * F, sp, [cap] * F); te = time.time() - t; de = c.last_device_ms
    t = time.time(); ns = c.decode_batch_raw(sp, lens, op, [n] * F); td = time.time() - t; dd = c.last_device_ms
print("F=%d streams=%s group=%s: encode wall %.1f dev %.1f ms | decode wall %.1f dev %.1f ms | %.1f Mpts/s" % (
    F, os.environ.get("CCV2_STREAMS", "-"), os.environ.get("CCV2_GROUP", "-"), te * 1e3, de, td * 1e3, dd, n * F / (de + dd) / 1e3))

for r in range(reps):
    t = time.time(); lens2, ns2 = c.roundtrip_batch_raw(ip, [n] * F, sp, [cap] * F, op, [n] * F); tr = time.time() - t; dr = c.last_device_ms
print("F=%d roundtrip: wall %.1f dev %.1f ms | %.1f Mpts/s ; same lens %s same counts %s" % (F, tr * 1e3, dr, n * F / tr / 1e6, lens2 == lens or "n/a(frame ids differ)", ns2 == ns))

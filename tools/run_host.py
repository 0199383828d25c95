"""Developer tool (GPU box): pinned-host encode+decode of F frames; prints wall ms of each call and PCIe copy bandwidth."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see csrc/ccv2_api.cu (stream -> hardware queue aliasing)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwi_pcl_codec_b200 import codec as K, synth
n = int(sys.argv[1]); F = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
base = [synth.gen_surface(n, s) for s in range(min(F, 8))]
cap = 4 * n + (1 << 16)
h_in = [K.PinnedBuffer(n * 32) for _ in range(F)]
for i, b in enumerate(h_in): b.array[:] = base[i % len(base)].view(np.uint8).reshape(-1)
h_str = [K.PinnedBuffer(cap) for _ in range(F)]
ONE = os.environ.get("ONE_BUF", "0") == "1"
if ONE:
    big = K.PinnedBuffer(n * 32 * F)
    class _V:                                   # view into the single allocation
        def __init__(self, p): self.ptr = p
    h_out = [_V(big.ptr + i * n * 32) for i in range(F)]
else:
    h_out = [K.PinnedBuffer(n * 32) for _ in range(F)]
# raw PCIe bandwidth with torch
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", x, y), ("D2H", y, x)):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(4): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); print("%s %.1f GB/s" % (name, 4 * (1 << 30) / (time.time() - t) / 1e9))
c = K.Codec(K.default_params(octree_bits=11))
ip = [b.ptr for b in h_in]; sp = [b.ptr for b in h_str]; op = [b.ptr for b in h_out]
for r in range(reps):
    t = time.time(); lens = c.encode_batch_raw(ip, [n] * F, sp, [cap] * F); te = time.time() - t; de = c.last_device_ms
    t = time.time(); ns = c.decode_batch_raw(sp, lens, op, [n] * F); td = time.time() - t; dd = c.last_device_ms
print("F=%d host pinned: encode wall %.1f dev %.1f ms | decode wall %.1f dev %.1f ms | %.1f Mpts/s e2e" % (F, te * 1e3, de, td * 1e3, dd, n * F / (te + td) / 1e6))

for r in range(reps):
    t = time.time(); lens2, ns2 = c.roundtrip_batch_raw(ip, [n] * F, sp, [cap] * F, op, [n] * F); tr = time.time() - t; dr = c.last_device_ms
print("F=%d roundtrip: wall %.1f dev %.1f ms | %.1f Mpts/s ; same lens %s same counts %s" % (F, tr * 1e3, dr, n * F / tr / 1e6, lens2 == lens or "n/a(frame ids differ)", ns2 == ns))

// ccv2_api.cu -- C ABI (include/ccv2.h) and host-side orchestration of the CUDA pipelines.
//
// Execution model: a batch of frames is cut into groups; every group runs start-to-finish on one CUDA stream of
// a small round-robin pool (H2D of the inputs, the parallel kernels with blockIdx.y = frame, the serial
// range coder stage -- lane-per-stream or warp-per-stream kernels, see entropy_kernels.cuh / dec_lps_kernels.cuh --
// assembly, D2H of the results).  Groups on different streams overlap; throughput is frames in flight divided by
// the per-frame latency of the serial entropy stage.  No host synchronisation happens
// inside a group: all sizes (depth, V, B, J, stream length) are device-side values in the frame records.
#include "../../include/ccv2.h"
#include "common.cuh"
#include "enc_kernels.cuh"
#include "jpeg_enc_kernels.cuh"
#include "entropy_kernels.cuh"
#include "dec_kernels.cuh"
#include "dec_lps_kernels.cuh"
#include "lines_kernels.cuh"
#include "lines_dec_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

// The batch driver keeps up to 26 streams busy (8 or 16 group streams, 8 side streams, a control and a copy stream).  CUDA maps streams onto
// CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue serialise behind each other's
// long serial kernels (measured: 2x on the round trip).  Ask for more queues unless the user already chose -- this only
// takes effect if it happens before the process creates its CUDA context, so hosts that initialise CUDA first
// (e.g. import torch; torch.cuda.init()) should export the variable themselves (bench.py and tests/conftest.py do).
struct EnvInit { EnvInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); } } g_env_init;

struct DevBuf {
  void *p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + std::min(bytes / 8, (size_t)64 << 20) + 4096;   // slack so slowly growing batches do not reallocate every call
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
  void *p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes + 4096);
    if (e == cudaSuccess) cap = bytes + 4096;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// bump allocator over a DevBuf: first pass (base == nullptr) only measures
struct Carver {
  uint8_t *base; size_t off = 0;
  explicit Carver(void *b) : base((uint8_t *)b) {}
  template <typename T> T *take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T *r = base ? (T *)(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
};

constexpr int MAX_STREAMS = 16;
constexpr int SIDE_STREAMS = 8;          // group + side + control streams stay within the 32 hardware queues

}  // namespace

// Per-frame results published straight into pinned host memory by a kernel (zero-copy store over PCIe): the host learns
// the sizes from an event right after the group's last kernel instead of waiting for a D2H copy that would queue behind
// other groups' result copies in the copy engine.
struct FrameResult { uint64_t out_len; uint32_t enc_error, dec_error, V, _pad; };
__global__ void publish_kernel(const EncFrame *enc, const DecFrame *dec, FrameResult *res, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FrameResult r; r.out_len = 0; r.enc_error = 0; r.dec_error = 0; r.V = 0; r._pad = 0;
  if (enc) { r.out_len = enc[i].out_len; r.enc_error = enc[i].error; }
  if (dec) { r.dec_error = dec[i].error; r.V = dec[i].V; }
  res[i] = r;
}

struct ccv2_codec {
  ccv2_params prm;
  int device = 0;
  int n_sm = 148;
  int trace = 0;                          // CCV2_TRACE=1: print per-group timeline (ms since batch start) to stderr
  std::vector<cudaEvent_t> ev_trace;
  struct TraceMark { int group; const char *label; };
  std::vector<TraceMark> trace_marks;
  int use_ring = 1;                       // pipeline the DFS walk behind the range decoder (CCV2_NO_RING=1 disables: debugging)
  int n_streams = 0, group = 0;           // streams 0 = auto (8 for device-resident calls, 16 when host buffers are involved); group 0 = auto: spread the batch over all streams
  cudaStream_t main_stream = nullptr;
  cudaStream_t copy_stream = nullptr;     // all host->device input copies, in group order (see run_batch)
  std::vector<cudaEvent_t> ev_h2d;
  cudaStream_t streams[MAX_STREAMS] = {};
  cudaStream_t side_streams[SIDE_STREAMS] = {};   // colour layer of a group, concurrent with its tree layer (lane-per-stream decoder)
  std::vector<cudaEvent_t> ev_side;
  int lps_dec = -1;                       // lane-per-stream range decoder: -1 auto (round trips only), 0 off, 1 on (CCV2_LPS_DEC)
  cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_fork = nullptr;
  std::vector<cudaEvent_t> ev_group;
  JpegTables *d_tables = nullptr;
  uint32_t *d_frame_counter = nullptr;
  size_t lps_smem_enc = 0;                // dynamic shared memory that keeps the lane-per-stream coder CTAs one to an SM
  int lps_enc = 1;                        // lane-per-stream range encoder (CCV2_LPS_ENC=0: one warp per stream)
  int serial_cap = 0;                     // serial CTAs per SM; 0: ceil(frames in the call / SMs)  (CCV2_CAP overrides, -1 disables the cap)
  size_t smem_sm = 0, smem_static_enc = 0, smem_static_dec = 0;   // shared memory per SM; static use of the two serial kernels
  uint32_t frame_id = 0;
  // encode workspaces
  DevBuf work;                             // encode slots / decode workspaces (one arena: see run_batch)
  int work_mode = -1;                      // mode of the call that last wrote the arena
  EncParams last_enc_params;               // of the last encode (ccv2_get_output_cloud)
  DevBuf out_cloud;                        // staging for ccv2_get_output_cloud into host memory
  DevBuf enc_frames, enc_persist, enc_input;
  HostBuf h_frames;
  std::vector<EncFrame> enc_host;        // host mirror of the last batch's frame records (with device pointers)
  // decode workspaces
  DevBuf dec_frames, dec_input, dec_output;
  HostBuf h_dframes, h_results;
  uint64_t metrics[3] = {0, 0, 0};
  uint64_t launches = 0;
  float device_ms = 0.f;
  std::string err;
  // profiling hook: events around every kernel launch (single stream)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_pool; size_t prof_used = 0;
  struct ProfRec { const char *name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  struct ProfSum { std::string name; float ms; int launches; };
  std::vector<ProfSum> prof_sum;
};

namespace {

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e__); return CCV2_ERR_CUDA; } } while (0)

cudaEvent_t prof_event(ccv2_codec *c) {
  if (c->prof_used == c->prof_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->prof_pool.push_back(e); }
  return c->prof_pool[c->prof_used++];
}
void prof_begin(ccv2_codec *c, cudaStream_t st, const char *name) {
  if (!c->profiling) return;
  ccv2_codec::ProfRec r; r.name = name; r.e0 = prof_event(c); r.e1 = prof_event(c);
  cudaEventRecord(r.e0, st);
  c->prof_recs.push_back(r);
}
void prof_end(ccv2_codec *c, cudaStream_t st) { if (c->profiling) cudaEventRecord(c->prof_recs.back().e1, st); }
void prof_collect(ccv2_codec *c) {
  if (!c->profiling) return;
  c->prof_sum.clear();
  for (auto &r : c->prof_recs) {
    float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
    bool found = false;
    for (auto &q : c->prof_sum) if (q.name == r.name) { q.ms += ms; q.launches++; found = true; break; }
    if (!found) c->prof_sum.push_back({r.name, ms, 1});
  }
  c->prof_recs.clear(); c->prof_used = 0;
}
#define LAUNCH_S(stream, name, ...) do { prof_begin(c, stream, name); __VA_ARGS__; prof_end(c, stream); launches++; } while (0)
#define LAUNCH(name, ...) LAUNCH_S(st, name, __VA_ARGS__)

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- JPEG tables on the host (libjpeg std tables)
const uint8_t ZZ_H[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
  21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
const uint8_t QL_H[64] = { 16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
  14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
const uint8_t QC_H[64] = { 17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
  47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
const uint8_t DCL_BITS[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 };
const uint8_t DCC_BITS[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
const uint8_t DC_VALS_H[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
const uint8_t ACL_BITS[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d };
const uint8_t ACL_VALS[162] = { 0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
  0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
  0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
  0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
  0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
  0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
  0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8,
  0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
const uint8_t ACC_BITS[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };
const uint8_t ACC_VALS[162] = { 0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
  0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
  0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46,
  0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
  0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98,
  0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
  0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7,
  0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

void huff_codes(const uint8_t bits[16], const uint8_t *vals, uint16_t *code, uint8_t *len) {
  unsigned c = 0; int k = 0;
  for (int l = 1; l <= 16; l++) {
    for (int i = 0; i < bits[l - 1]; i++) { code[vals[k]] = (uint16_t)c; len[vals[k]] = (uint8_t)l; c++; k++; }
    c <<= 1;
  }
}
void build_jpeg_tables(int quality, JpegTables &T) {
  memset(&T, 0, sizeof T);
  memcpy(T.zz, ZZ_H, 64);
  int q = quality; if (q <= 0) q = 1; if (q > 100) q = 100;
  int scale = q < 50 ? 5000 / q : 200 - q * 2;
  for (int t = 0; t < 2; t++) for (int i = 0; i < 64; i++) {
    long v = ((long)(t ? QC_H[i] : QL_H[i]) * scale + 50L) / 100L;
    if (v <= 0) v = 1; if (v > 255) v = 255;
    T.q[t][i] = (uint16_t)v;
  }
  huff_codes(DCL_BITS, DC_VALS_H, T.dc_code[0], T.dc_len[0]); huff_codes(DCC_BITS, DC_VALS_H, T.dc_code[1], T.dc_len[1]);
  huff_codes(ACL_BITS, ACL_VALS, T.ac_code[0], T.ac_len[0]); huff_codes(ACC_BITS, ACC_VALS, T.ac_code[1], T.ac_len[1]);
  // header template (623 bytes): SOI APP0 DQT DQT SOF0 DHTx4 SOS ; height patched per frame, width 256
  std::vector<uint8_t> h;
  auto p8 = [&](int v) { h.push_back((uint8_t)v); };
  auto p16 = [&](int v) { h.push_back((uint8_t)(v >> 8)); h.push_back((uint8_t)v); };
  p16(0xFFD8);
  const uint8_t app0[] = { 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 };
  h.insert(h.end(), app0, app0 + sizeof app0);
  for (int t = 0; t < 2; t++) { p16(0xFFDB); p16(67); p8(t); for (int i = 0; i < 64; i++) p8(T.q[t][ZZ_H[i]]); }
  p16(0xFFC0); p16(17); p8(8); p16(0); p16(256); p8(3); p8(1); p8(0x22); p8(0); p8(2); p8(0x11); p8(1); p8(3); p8(0x11); p8(1);
  auto dht = [&](int id, const uint8_t *bits, const uint8_t *vals) { int n = 0; for (int i = 0; i < 16; i++) n += bits[i]; p16(0xFFC4); p16(19 + n); p8(id); h.insert(h.end(), bits, bits + 16); h.insert(h.end(), vals, vals + n); };
  dht(0x00, DCL_BITS, DC_VALS_H); dht(0x10, ACL_BITS, ACL_VALS); dht(0x01, DCC_BITS, DC_VALS_H); dht(0x11, ACC_BITS, ACC_VALS);
  p16(0xFFDA); p16(12); p8(3); p8(1); p8(0x00); p8(2); p8(0x11); p8(3); p8(0x11); p8(0); p8(63); p8(0);
  if (h.size() != JPEG_HDR_BYTES) { fprintf(stderr, "ccv2: jpeg header template is %zu bytes\n", h.size()); abort(); }
  memcpy(T.header, h.data(), JPEG_HDR_BYTES);
}

// ---- capacity rules (bytes) for a frame of n points
size_t tree_cap_for(size_t n) { return ((8 * n + 1024) + 255) & ~size_t(255); }
size_t cpay_cap_for(size_t n) { return ((4 * n + 8192) + 255) & ~size_t(255); }
size_t cen_cap_for(size_t n) { return ((3 * n + 256) + 255) & ~size_t(255); }
size_t rc_cap_for(size_t raw_cap) { return ((raw_cap + raw_cap / 8 + 4096) + 255) & ~size_t(255); }
size_t stream_cap_for(size_t n, bool cen) {
  return FRAME_HDR_BYTES + 8 + rc_cap_for(tree_cap_for(n)) + (cen ? 4 + rc_cap_for(cen_cap_for(n)) : 0) + 8 + rc_cap_for(cpay_cap_for(n));
}

struct EncSlotLayout { size_t bytes; size_t zero_off, zero_bytes; };

// carve the per-frame group-slot workspace; returns total bytes (measure when base == nullptr)
size_t carve_enc_slot(uint8_t *base, size_t n, EncFrame *f, const ccv2_params &prm, size_t *zero_off, size_t *zero_bytes) {
  Carver cv(base);
  const uint32_t tiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE) + 1;
  const uint32_t scan_tiles = (uint32_t)(n / 1024) + 8;
  const size_t img_h = n / 256 + 1, mcu_h = (img_h + 15) / 16;
  const bool lines = prm.color_coding_type == 2;
  const size_t lines_cap = n / LINE_PX + 2;
  const size_t jbits_words = lines ? lines_cap * LINE_BITS_WORDS : (((4 * n + 8192) / 4 + 63) & ~size_t(63));
  // --- zero-initialised region first
  cv.off = 0;
  uint32_t *ghist = cv.take<uint32_t>(8 * 256);
  size_t z0 = 0;
  uint32_t *sort_status = cv.take<uint32_t>((size_t)8 * tiles * 256);
  uint64_t *scan_status = cv.take<uint64_t>((size_t)3 * scan_tiles);
  uint32_t *jbits = cv.take<uint32_t>(jbits_words + 16);
  size_t z1 = (cv.off + 255) & ~size_t(255);
  // --- rest
  uint64_t *k0 = cv.take<uint64_t>(n + 8), *k1 = cv.take<uint64_t>(n + 8);
  uint32_t *v0 = cv.take<uint32_t>(n + 8), *v1 = cv.take<uint32_t>(n + 8);
  uint64_t *leaf_key = cv.take<uint64_t>(n + 8);
  uint32_t *leaf_start = cv.take<uint32_t>(n + 8), *leaf_off = cv.take<uint32_t>(n + 8);
  uint8_t *first_new = cv.take<uint8_t>(n + 8);
  uint8_t *avg = cv.take<uint8_t>(3 * n + 64);
  int16_t *coef = cv.take<int16_t>((lines ? (n / 16 + 260) : mcu_h * 16) * 384 + 64);
  uint8_t *line_slots = cv.take<uint8_t>(lines ? lines_cap * LINE_SLOT_BYTES : 16);
  uint32_t *line_len = cv.take<uint32_t>(lines ? lines_cap : 4), *line_off = cv.take<uint32_t>(lines ? lines_cap : 4);
  if (f) {
    f->ghist = ghist; f->sort_status = sort_status; f->tiles_max = tiles; f->scan_status = scan_status; f->scan_tiles_max = scan_tiles;
    f->jbits_buf = jbits; f->jbits_cap_words = (uint32_t)jbits_words;
    f->keys[0] = k0; f->keys[1] = k1; f->vals[0] = v0; f->vals[1] = v1;
    f->leaf_key = leaf_key; f->leaf_start = leaf_start; f->leaf_off = leaf_off; f->first_new = first_new;
    f->avg = avg; f->coef = coef;
    f->line_slots = line_slots; f->line_len = line_len; f->line_off = line_off; f->lines_cap = lines ? (uint32_t)lines_cap : 0;
  }
  if (zero_off) *zero_off = z0;
  if (zero_bytes) *zero_bytes = z1 - z0;
  return (cv.off + 255) & ~size_t(255);
}
size_t carve_enc_persist(uint8_t *base, size_t n, EncFrame *f, bool cen) {
  Carver cv(base);
  uint8_t *tree = cv.take<uint8_t>(tree_cap_for(n));
  uint8_t *cenb = cv.take<uint8_t>(cen ? cen_cap_for(n) : 256);
  uint8_t *cpay = cv.take<uint8_t>(cpay_cap_for(n));
  uint8_t *rc0 = cv.take<uint8_t>(cen ? rc_cap_for(cen_cap_for(n)) : 256);
  uint8_t *rc1 = cv.take<uint8_t>(rc_cap_for(cpay_cap_for(n)));
  uint8_t *stream = cv.take<uint8_t>(stream_cap_for(n, cen));
  if (f) {
    f->tree = tree; f->tree_cap = (uint32_t)tree_cap_for(n) - 64; f->cen = cenb; f->cpay = cpay; f->cpay_cap = (uint32_t)cpay_cap_for(n) - 64;
    f->rc_tmp[0] = rc0; f->rc_tmp[1] = rc1; f->rc_tmp_cap[0] = (uint32_t)(cen ? rc_cap_for(cen_cap_for(n)) : 0); f->rc_tmp_cap[1] = (uint32_t)rc_cap_for(cpay_cap_for(n));
    f->stream = stream; f->stream_cap = stream_cap_for(n, cen);
  }
  return (cv.off + 255) & ~size_t(255);
}

int check_params(const ccv2_params *p, std::string &err) {
  if (!p) { err = "null params"; return CCV2_ERR_ARG; }
  if (p->profile != CCV2_MANUAL_CONFIGURATION) { err = "only MANUAL_CONFIGURATION is implemented"; return CCV2_ERR_UNSUPPORTED; }
  if (!p->do_voxel_grid_downsampling) { err = "detail mode (doVoxelGridDownDownSampling=false) is not implemented"; return CCV2_ERR_UNSUPPORTED; }
  if (!(p->octree_resolution > 0)) { err = "octree_resolution must be > 0"; return CCV2_ERR_ARG; }
  if (p->color_coding_type > 3) { err = "unknown colorCodingType"; return CCV2_ERR_ARG; }
  if (p->color_bit_resolution > 8) { err = "colorBitResolution > 8"; return CCV2_ERR_ARG; }
  return CCV2_OK;
}

}  // namespace

extern "C" {

void ccv2_default_params(ccv2_params *p) {
  memset(p, 0, sizeof *p);
  p->profile = CCV2_MANUAL_CONFIGURATION;
  p->point_resolution = ldexp(1.0, -11); p->octree_resolution = ldexp(1.0, -11);
  p->do_voxel_grid_downsampling = 1; p->i_frame_rate = 0; p->do_color_encoding = 1; p->color_bit_resolution = 8;
  p->color_coding_type = 1; p->do_voxel_grid_centroid = 0; p->create_scalable_stream = 0; p->code_connectivity = 0;
  p->jpeg_quality = 85; p->num_threads = 1; p->macroblock_size = 16; p->do_icp_color_offset = 0;
}

const char *ccv2_status_string(int s) {
  switch (s) {
    case CCV2_OK: return "ok";
    case CCV2_ERR_ARG: return "bad argument";
    case CCV2_ERR_CUDA: return "CUDA error";
    case CCV2_ERR_UNSUPPORTED: return "configuration not implemented";
    case CCV2_ERR_CAPACITY: return "output buffer too small";
    case CCV2_ERR_WORKSPACE: return "internal workspace bound exceeded";
    case CCV2_ERR_STREAM: return "malformed compressed stream";
    case CCV2_ERR_DEPTH: return "octree depth > 21";
    default: return "unknown";
  }
}
const char *ccv2_last_error(const ccv2_codec *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ccv2_create(const ccv2_params *p, int device, ccv2_codec **out) {
  if (!out) return CCV2_ERR_ARG;
  *out = nullptr;
  int rc = check_params(p, g_create_error);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal");
    cudaGetLastError();
    return CCV2_ERR_CUDA;
  }
  ccv2_codec *c = new ccv2_codec();
  c->prm = *p; c->device = device;
  if (const char *s = getenv("CCV2_TRACE")) c->trace = atoi(s);
  if (const char *s = getenv("CCV2_CAP")) c->serial_cap = atoi(s);
  if (const char *s = getenv("CCV2_LPS_ENC")) c->lps_enc = atoi(s) != 0;
  if (const char *s = getenv("CCV2_LPS_DEC")) c->lps_dec = atoi(s);
  if (const char *s = getenv("CCV2_NO_RING")) c->use_ring = atoi(s) ? 0 : 1;
  if (const char *s = getenv("CCV2_STREAMS")) c->n_streams = std::max(0, std::min(MAX_STREAMS, atoi(s)));
  if (const char *s = getenv("CCV2_GROUP")) c->group = std::max(0, std::min(1024, atoi(s)));
  auto fail = [&](cudaError_t ee, const char *what) { g_create_error = std::string(what) + ": " + cudaGetErrorString(ee); ccv2_destroy(c); return CCV2_ERR_CUDA; };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  if ((e = cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail(e, "cudaDeviceGetAttribute");
  if ((e = cudaStreamCreateWithFlags(&c->main_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  {
    // optional (CCV2_PRIORITY=1): earlier groups get higher stream priority.  Measured neutral on B200 (the long serial
    // kernels are resident anyway), so it is off by default.
    int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);      // lo = least urgent (numerically largest)
    const bool prio = getenv("CCV2_PRIORITY") && atoi(getenv("CCV2_PRIORITY")) != 0;
    for (int i = 0; i < MAX_STREAMS; i++) {
      int p = prio ? std::min(lo, hi + i * (lo - hi + 1) / MAX_STREAMS) : lo;
      if ((e = cudaStreamCreateWithPriority(&c->streams[i], cudaStreamNonBlocking, p)) != cudaSuccess) return fail(e, "cudaStreamCreate");
      if (i < SIDE_STREAMS && (e = cudaStreamCreateWithPriority(&c->side_streams[i], cudaStreamNonBlocking, lo)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    }
  }
  if ((e = cudaEventCreate(&c->ev_start)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&c->ev_end)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaFuncSetAttribute(sort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_BYTES)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
  JpegTables T; build_jpeg_tables(p->jpeg_quality, T);
  if ((e = cudaMalloc(&c->d_tables, sizeof T)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(c->d_tables, &T, sizeof T, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy");
  if ((e = cudaMalloc(&c->d_frame_counter, 4)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMemset(c->d_frame_counter, 0, 4)) != cudaSuccess) return fail(e, "cudaMemset");
  {
    cudaFuncAttributes fa; int per_sm = 0, per_block = 0;
    if ((e = cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device)) != cudaSuccess) return fail(e, "cudaDeviceGetAttribute");
    if ((e = cudaDeviceGetAttribute(&per_block, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return fail(e, "cudaDeviceGetAttribute");
    c->smem_sm = (size_t)per_sm;
    if ((e = cudaFuncGetAttributes(&fa, rc_encode_kernel)) != cudaSuccess) return fail(e, "cudaFuncGetAttributes");
    c->smem_static_enc = fa.sharedSizeBytes;
    if ((e = cudaFuncSetAttribute(rc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, per_block - (int)fa.sharedSizeBytes)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = cudaFuncGetAttributes(&fa, dec_entropy_kernel)) != cudaSuccess) return fail(e, "cudaFuncGetAttributes");
    c->smem_static_dec = fa.sharedSizeBytes;
    if ((e = cudaFuncSetAttribute(dec_entropy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, per_block - (int)fa.sharedSizeBytes)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = cudaFuncGetAttributes(&fa, rc_encode_lps_kernel)) != cudaSuccess) return fail(e, "cudaFuncGetAttributes");
    c->lps_smem_enc = ((size_t)per_sm / 1024 / 2 + 1) * 1024 - fa.sharedSizeBytes - 1024;      // more than half an SM's shared memory per CTA
    if ((e = cudaFuncSetAttribute(rc_encode_lps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lps_smem_enc)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    cudaFuncSetAttribute(rc_encode_lps_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if ((e = cudaFuncSetAttribute(rc_decode_lps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpsSmem))) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = cudaFuncSetAttribute(rc_decode_lps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)offsetof(LpsSmem, rg))) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    cudaFuncSetAttribute(rc_decode_lps_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(rc_decode_lps_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(rc_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(dec_entropy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  *out = c;
  return CCV2_OK;
}

void ccv2_destroy(ccv2_codec *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto ev : c->ev_group) cudaEventDestroy(ev);
  for (auto ev : c->prof_pool) cudaEventDestroy(ev);
  for (auto ev : c->ev_trace) cudaEventDestroy(ev);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_end) cudaEventDestroy(c->ev_end);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int i = 0; i < MAX_STREAMS; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
  for (int i = 0; i < SIDE_STREAMS; i++) if (c->side_streams[i]) cudaStreamDestroy(c->side_streams[i]);
  for (auto ev : c->ev_side) cudaEventDestroy(ev);
  if (c->main_stream) cudaStreamDestroy(c->main_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto ev : c->ev_h2d) cudaEventDestroy(ev);
  if (c->d_tables) cudaFree(c->d_tables);
  if (c->d_frame_counter) cudaFree(c->d_frame_counter);
  c->work.release(); c->out_cloud.release(); c->enc_frames.release(); c->enc_persist.release(); c->enc_input.release();
  c->dec_frames.release(); c->dec_input.release(); c->dec_output.release();
  c->h_frames.release(); c->h_dframes.release(); c->h_results.release();
  delete c;
}

size_t ccv2_max_compressed_size(size_t npts) { return stream_cap_for(npts, true); }

void *ccv2_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void ccv2_host_free(void *p) { if (p) cudaFreeHost(p); }

int ccv2_get_metrics(ccv2_codec *c, uint64_t m[3]) { if (!c || !m) return CCV2_ERR_ARG; for (int i = 0; i < 3; i++) m[i] = c->metrics[i]; return CCV2_OK; }
int ccv2_set_frame_id(ccv2_codec *c, uint32_t v) {
  if (!c) return CCV2_ERR_ARG;
  cudaSetDevice(c->device);
  c->frame_id = v;
  CU(cudaMemcpy(c->d_frame_counter, &v, 4, cudaMemcpyHostToDevice));
  return CCV2_OK;
}
uint32_t ccv2_get_frame_id(const ccv2_codec *c) { return c ? c->frame_id : 0; }
uint64_t ccv2_last_launch_count(const ccv2_codec *c) { return c ? c->launches : 0; }
float ccv2_last_device_ms(const ccv2_codec *c) { return c ? c->device_ms : 0.f; }

int ccv2_set_profiling(ccv2_codec *c, int on) { if (!c) return CCV2_ERR_ARG; c->profiling = on != 0; c->prof_sum.clear(); return CCV2_OK; }
int ccv2_get_profile(const ccv2_codec *c, int idx, const char **name, float *total_ms, int *launches) {
  if (!c || idx < 0 || idx >= (int)c->prof_sum.size()) return CCV2_ERR_ARG;
  if (name) *name = c->prof_sum[idx].name.c_str();
  if (total_ms) *total_ms = c->prof_sum[idx].ms;
  if (launches) *launches = c->prof_sum[idx].launches;
  return CCV2_OK;
}

int ccv2_peek_point_count(const void *in_host, size_t len, uint64_t *npts) {
  if (!in_host || !npts || len < FRAME_HDR_BYTES) return CCV2_ERR_ARG;
  const uint8_t *b = (const uint8_t *)in_host;
  if (memcmp(b, "<PCL-OCT-CODECV2-COMPRESSED><PCL-OCT-COMPRESSED>", 48) != 0) return CCV2_ERR_STREAM;
  memcpy(npts, b + 55, 8);
  return CCV2_OK;
}

static size_t carve_dec(uint8_t *base, size_t pcap, DecFrame *f, size_t *zero_off, size_t *zero_bytes, bool lines = false) {
  Carver cv(base);
  const size_t lines_cap = lines ? pcap / LINE_PX + 2 : 0;
  const size_t img_h = pcap / 256 + 2, mcu_h = (img_h + 15) / 16, nblocks = lines ? (lines_cap * LINE_MCU_STRIDE + LINE_MCU_STRIDE) * 6 : mcu_h * 16 * 6;
  const uint32_t scan_tiles = (uint32_t)(pcap / NODE_THREADS) + 8;
  uint64_t *scan_status = cv.take<uint64_t>(2 * (size_t)scan_tiles);
  int16_t *coef = cv.take<int16_t>(nblocks * 64 + 64);
  size_t z1 = (cv.off + 255) & ~size_t(255);
  uint8_t *tree = cv.take<uint8_t>(tree_cap_for(pcap));
  uint8_t *cen = cv.take<uint8_t>(cen_cap_for(pcap));
  uint8_t *col = cv.take<uint8_t>(cpay_cap_for(pcap));
  uint64_t *node_prefix = cv.take<uint64_t>(pcap + 8);
  uint8_t *node_byte = cv.take<uint8_t>(pcap + 8);
  uint64_t *l2_prefix = cv.take<uint64_t>(pcap + 8);
  uint8_t *l2_mask = cv.take<uint8_t>(pcap + 8);
  uint32_t *l2_off = cv.take<uint32_t>(pcap + 8);
  uint8_t *planes = cv.take<uint8_t>(mcu_h * 16 * 256 * 3 / 2 + 256);
  uint16_t *qt = cv.take<uint16_t>(128);
  uint8_t *scan = cv.take<uint8_t>(cpay_cap_for(pcap));
  uint32_t *line_off = cv.take<uint32_t>(lines_cap + 4), *line_len = cv.take<uint32_t>(lines_cap + 4), *line_w = cv.take<uint32_t>(lines_cap + 4);
  uint16_t *line_qt = cv.take<uint16_t>(lines_cap * 128 + 128);
  uint8_t *line_planes = cv.take<uint8_t>(lines_cap * 8192 + 256);
  if (f) {
    f->scan_status = scan_status; f->scan_tiles_max = scan_tiles; f->coef = coef; f->coef_cap_blocks = (uint32_t)nblocks;
    f->tree = tree; f->tree_cap = (uint32_t)tree_cap_for(pcap) - 64; f->cen = cen; f->cen_cap = (uint32_t)cen_cap_for(pcap) - 64;
    f->col = col; f->col_cap = (uint32_t)cpay_cap_for(pcap) - 64;
    f->node_prefix = node_prefix; f->node_byte = node_byte; f->node_cap = (uint32_t)pcap;
    f->l2_prefix = l2_prefix; f->l2_mask = l2_mask; f->l2_off = l2_off;
    f->planes = planes; f->planes_cap = (uint32_t)(mcu_h * 16 * 256 * 3 / 2); f->qt = qt; f->scan = scan;
    f->lines_cap = (uint32_t)lines_cap; f->line_off = line_off; f->line_len = line_len; f->line_w = line_w; f->line_qt = line_qt; f->line_planes = line_planes;
  }
  if (zero_off) *zero_off = 0;
  if (zero_bytes) *zero_bytes = z1;
  return (cv.off + 255) & ~size_t(255);
}

// ================================================================================================ batch driver
// mode 0: encode, 1: decode, 2: encode -> decode round trip (decode reads the encoder's device-resident streams, so the
// host->device copies of later groups overlap the device->host copies of earlier ones: both PCIe directions busy).
static int run_batch(ccv2_codec *c, int mode, int nframes,
                     const void *const *pts, const size_t *npts, void *const *out, const size_t *out_cap, size_t *out_len,
                     const void *const *in, const size_t *in_len, void *const *pts_out, const size_t *pts_cap, size_t *npts_out) {
  const bool do_enc = mode != 1, do_dec = mode != 0, rt = mode == 2;
  c->err.clear(); c->launches = 0; c->device_ms = 0;
  if (nframes == 0) return CCV2_OK;
  c->work_mode = mode;
  CU(cudaSetDevice(c->device));
  const ccv2_params &prm = c->prm;
  const bool cen = prm.do_voxel_grid_centroid != 0, color = prm.do_color_encoding != 0;
  // Device-resident batches run best as 8 groups (fewer, larger launches; measured 1640 against 1390 Mpoints/s with 16).
  // When clouds or results cross PCIe the pipeline is paced by the copies, and 16 smaller groups start earlier and
  // leave a shorter copy-back tail (end to end 900 against 820 Mpoints/s).
  const void *first_io = do_enc ? (pts ? pts[0] : nullptr) : (in ? in[0] : nullptr);
  const void *first_out = do_dec ? (pts_out ? pts_out[0] : nullptr) : (out ? out[0] : nullptr);
  const bool host_io = (first_io && !is_device_ptr(first_io)) || (first_out && !is_device_ptr(first_out));
  const int NS = c->profiling ? 1 : (c->n_streams ? c->n_streams : (host_io ? MAX_STREAMS : 8));
  const int G = c->profiling ? std::max(1, std::min(nframes, 64)) : (c->group ? c->group : std::max(1, std::min(256, (nframes + NS - 1) / NS)));
  const int ngroups = (nframes + G - 1) / G;
  while ((int)c->ev_h2d.size() < ngroups) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); c->ev_h2d.push_back(ev); }
  while ((int)c->ev_side.size() < 2 * ngroups) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); c->ev_side.push_back(ev); }
  while ((int)c->ev_group.size() < 2 * ngroups) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); c->ev_group.push_back(ev); }
  cudaStream_t ms = c->main_stream;
  CU(c->h_results.ensure(sizeof(FrameResult) * nframes));
  FrameResult *hres = (FrameResult *)c->h_results.p;
  c->trace_marks.clear();
  auto mark = [&](int g, const char *label, cudaStream_t s) {       // CCV2_TRACE: timestamp on the group's stream
    if (!c->trace) return;
    if (c->ev_trace.size() <= c->trace_marks.size()) { cudaEvent_t ev; if (cudaEventCreate(&ev) != cudaSuccess) return; c->ev_trace.push_back(ev); }
    cudaEventRecord(c->ev_trace[c->trace_marks.size()], s);
    c->trace_marks.push_back({g, label});
  };

  // ------------------------------------------------------------------ encode side set-up
  size_t nmax = 1, zoff = 0, zbytes = 0, slot_bytes = 0, frames_bytes = 0;
  std::vector<size_t> input_off(nframes + 1, 0);
  std::vector<char> in_dev(nframes, 1), out_dev(nframes, 0);
  EncFrame *hf = nullptr, *df = nullptr;
  EncParams P; HeaderParams H;
  memset(&P, 0, sizeof P); memset(&H, 0, sizeof H);
  if (do_enc) {
    for (int i = 0; i < nframes; i++) { if (npts[i] >= (1u << 28)) { c->err = "frame too large"; return CCV2_ERR_ARG; } if (npts[i] && !pts[i]) return CCV2_ERR_ARG; nmax = std::max(nmax, npts[i]); }
    // Slots: one per (stream, frame-in-group); persist + input staging: one per frame of the batch
    slot_bytes = carve_enc_slot(nullptr, nmax, nullptr, prm, &zoff, &zbytes);
    if (rt) {                                                        // round trip: a frame's decode workspace reuses its encode slot (dead once the stream is assembled)
      size_t pmax = 0, zo, zz;
      for (int i = 0; i < nframes; i++) pmax = std::max(pmax, pts_cap[i]);
      if (pmax >= (1u << 28)) { c->err = "frame too large"; return CCV2_ERR_ARG; }
      slot_bytes = std::max(slot_bytes, carve_dec(nullptr, pmax, nullptr, &zo, &zz, prm.color_coding_type == 2));
    }
    const int nslots = std::min(ngroups, NS) * G;
    CU(c->work.ensure(slot_bytes * nslots));
    std::vector<size_t> persist_off(nframes + 1, 0);
    for (int i = 0; i < nframes; i++) {
      persist_off[i + 1] = persist_off[i] + carve_enc_persist(nullptr, npts[i], nullptr, cen);
      in_dev[i] = npts[i] ? is_device_ptr(pts[i]) : 1;
      out_dev[i] = (out && out[i]) ? is_device_ptr(out[i]) : 0;
      input_off[i + 1] = input_off[i] + (in_dev[i] ? 0 : ((32 * npts[i] + 255) & ~size_t(255)));
    }
    CU(c->enc_persist.ensure(persist_off[nframes]));
    CU(c->enc_input.ensure(input_off[nframes] + 256));
    frames_bytes = (sizeof(EncFrame) * nframes + 255) & ~size_t(255);
    CU(c->enc_frames.ensure(frames_bytes + (size_t)nframes * 3 * 256 * 4));
    CU(c->h_frames.ensure(sizeof(EncFrame) * nframes));
    hf = (EncFrame *)c->h_frames.p; df = (EncFrame *)c->enc_frames.p;
    memset(hf, 0, sizeof(EncFrame) * nframes);
    for (int i = 0; i < nframes; i++) {
      EncFrame &f = hf[i];
      const int g = i / G, slot = (g % NS) * G + (i % G);
      f.pts = in_dev[i] ? (const uint8_t *)pts[i] : (const uint8_t *)c->enc_input.p + input_off[i];
      f.n = (uint32_t)npts[i];
      f.n_finite = (uint32_t)npts[i];                               // keygen subtracts the non-finite points
      f.violator = NONE_U32;
      carve_enc_slot((uint8_t *)c->work.p + slot_bytes * slot, nmax, &f, prm, nullptr, nullptr);
      f.zero_ptr = (uint8_t *)c->work.p + slot_bytes * slot + zoff; f.zero_bytes = zbytes;
      carve_enc_persist((uint8_t *)c->enc_persist.p + persist_off[i], npts[i], &f, cen);
      f.hist = (uint32_t *)((uint8_t *)c->enc_frames.p + frames_bytes) + (size_t)i * 3 * 256;
      if (color && (prm.color_coding_type == 0 || prm.color_coding_type == 3)) f.avg = f.cpay;   // raw averages are the colour payload
    }
    P.res = prm.octree_resolution;
    { int ex; double m = frexp(P.res, &ex); P.res_pow2 = (m == 0.5); P.inv_res = P.res_pow2 ? 1.0 / P.res : 0.0; }
    P.do_color = color; P.color_type = prm.color_coding_type; P.do_centroid = cen;
    P.color_reduction = (prm.color_coding_type == 0) ? std::max(0, 8 - (int)prm.color_bit_resolution) : 0;   // jp_color_coder_ is never configured
    P.prefix_len = 16384;
    H.octree_res = prm.octree_resolution; H.point_res = (double)(float)prm.point_resolution;
    H.do_voxel_grid = 1; H.with_color = color; H.color_bits = prm.color_bit_resolution; H.do_centroid = cen;
    H.connectivity = prm.code_connectivity != 0; H.scalable = prm.create_scalable_stream != 0; H.icp_offset = prm.do_icp_color_offset != 0; H._p = 0;
    H.color_type = prm.color_coding_type; H.macroblock = prm.macroblock_size;
    c->last_enc_params = P;
  }

  // ------------------------------------------------------------------ decode side set-up
  std::vector<size_t> work_off(nframes + 1, 0), dinput_off(nframes + 1, 0), output_off(nframes + 1, 0), zb(nframes, 0);
  std::vector<char> din_dev(nframes, 1), dout_dev(nframes, 1);
  DecFrame *hd = nullptr, *dd = nullptr;
  if (do_dec) {
    for (int i = 0; i < nframes; i++) {
      if (pts_cap[i] >= (1u << 28)) { c->err = "frame too large"; return CCV2_ERR_ARG; }
      if ((!rt && in_len[i] && !in[i]) || (pts_cap[i] && !pts_out[i])) return CCV2_ERR_ARG;
      size_t zo;
      work_off[i + 1] = work_off[i] + carve_dec(nullptr, pts_cap[i], nullptr, &zo, &zb[i], prm.color_coding_type == 2);
      din_dev[i] = rt ? 1 : (in_len[i] ? is_device_ptr(in[i]) : 1);
      dout_dev[i] = pts_cap[i] ? is_device_ptr(pts_out[i]) : 1;
      dinput_off[i + 1] = dinput_off[i] + (din_dev[i] ? 0 : ((in_len[i] + 64 + 255) & ~size_t(255)));
      output_off[i + 1] = output_off[i] + (dout_dev[i] ? 0 : ((32 * pts_cap[i] + 255) & ~size_t(255)));
    }
    if (!rt) CU(c->work.ensure(work_off[nframes]));                  // one arena serves encode slots and decode workspaces: a call uses one or the other (or aliases them, round trip)
    CU(c->dec_input.ensure(dinput_off[nframes] + 256));
    CU(c->dec_output.ensure(output_off[nframes] + 256));
    CU(c->dec_frames.ensure(sizeof(DecFrame) * nframes));
    CU(c->h_dframes.ensure(sizeof(DecFrame) * nframes));
    hd = (DecFrame *)c->h_dframes.p; dd = (DecFrame *)c->dec_frames.p;
    memset(hd, 0, sizeof(DecFrame) * nframes);
    for (int i = 0; i < nframes; i++) {
      DecFrame &f = hd[i];
      if (rt) { f.in = hf[i].stream; f.in_len = 0; }       // length filled in on the device by link_kernel
      else {
        f.in = din_dev[i] ? (const uint8_t *)in[i] : (const uint8_t *)c->dec_input.p + dinput_off[i];
        f.in_len = in_len[i];
        if (in_len[i] == 0) f.error = FERR_BAD_STREAM;
      }
      f.out_pts = dout_dev[i] ? (uint8_t *)pts_out[i] : (uint8_t *)c->dec_output.p + output_off[i];
      f.out_cap = pts_cap[i];
      uint8_t *wbase = rt ? (uint8_t *)c->work.p + slot_bytes * (((i / G) % NS) * G + (i % G)) : (uint8_t *)c->work.p + work_off[i];
      carve_dec(wbase, pts_cap[i], &f, nullptr, nullptr, prm.color_coding_type == 2);
      f.zero_ptr = wbase; f.zero_bytes = zb[i];
    }
  }

  // ------------------------------------------------------------------ enqueue
  CU(cudaEventRecord(c->ev_start, ms));
  if (do_enc) {
    CU(cudaMemcpyAsync(df, hf, sizeof(EncFrame) * nframes, cudaMemcpyHostToDevice, ms));
    CU(cudaMemsetAsync((uint8_t *)c->enc_frames.p + frames_bytes, 0, (size_t)nframes * 3 * 256 * 4, ms));
  }
  if (do_dec) CU(cudaMemcpyAsync(dd, hd, sizeof(DecFrame) * nframes, cudaMemcpyHostToDevice, ms));
  // serial CTAs reserve enough shared memory (1 KB per CTA is the system's) that only serial_cap of them fit on an SM
  const int serial_cap = c->serial_cap ? c->serial_cap : (nframes + c->n_sm - 1) / c->n_sm;
  size_t serial_smem_enc = 0, serial_smem_dec = 0;
  if (serial_cap > 0 && serial_cap <= 14 && !c->profiling) {
    const size_t per_cta = (c->smem_sm / 1024 / (size_t)(serial_cap + 1) + 1) * 1024;
    if (per_cta > c->smem_static_enc + 1024) serial_smem_enc = per_cta - c->smem_static_enc - 1024;
    if (per_cta > c->smem_static_dec + 1024) serial_smem_dec = per_cta - c->smem_static_dec - 1024;
  }
  cudaEvent_t ev_setup = c->ev_fork;
  CU(cudaEventRecord(ev_setup, ms));
  uint64_t launches = 0;
  uint32_t *counter = c->d_frame_counter;
  for (int g = 0; g < ngroups; g++) {
    cudaStream_t st = c->streams[g % NS];
    const int f0 = g * G, gf = std::min(G, nframes - f0);
    const unsigned steered_grid = (unsigned)((gf + c->n_sm - 1) / c->n_sm * c->n_sm);
    CU(cudaStreamWaitEvent(st, ev_setup, 0));
    if (do_enc) {
      EncFrame *dg = df + f0;
      size_t gn = 1;
      // Host inputs go through ONE copy stream in group order: copies issued on the group streams would be
      // interleaved by the copy engine and every group's input would land at the very end (measured), which defeats
      // the pipeline.  This way group g can start as soon as its own clouds are on the device.
      bool any_h2d = false;
      for (int i = 0; i < gf; i++) {
        gn = std::max(gn, npts[f0 + i]);
        if (!in_dev[f0 + i] && npts[f0 + i]) {
          if (!any_h2d && g == 0) CU(cudaStreamWaitEvent(c->copy_stream, ev_setup, 0));
          CU(cudaMemcpyAsync((void *)hf[f0 + i].pts, pts[f0 + i], 32 * npts[f0 + i], cudaMemcpyHostToDevice, c->copy_stream));
          any_h2d = true;
        }
      }
      LAUNCH("zero_region_kernel", zero_region_kernel<EncFrame><<<dim3(128, gf), 256, 0, st>>>(dg));
      if (any_h2d) { CU(cudaEventRecord(c->ev_h2d[g], c->copy_stream)); CU(cudaStreamWaitEvent(st, c->ev_h2d[g], 0)); }
      mark(g, "inputs", st);
      const unsigned gx256 = (unsigned)((gn + 255) / 256), gtiles = (unsigned)((gn + SORT_TILE - 1) / SORT_TILE);
      LAUNCH("bbox_kernel", bbox_kernel<<<gf, 1024, 0, st>>>(dg, P, 0));
      LAUNCH("bbox_fixup_kernel", bbox_fixup_kernel<<<gf, 32, 0, st>>>(dg, P));
      LAUNCH("keygen_kernel", keygen_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P, 0));
      LAUNCH("bbox_kernel(slow path)", bbox_kernel<<<gf, 1024, 0, st>>>(dg, P, 1));
      LAUNCH("keygen_kernel(rekey)", keygen_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P, 1));
      // frame ids are sequential over the batch: group g's setup needs group g-1's setup kernel to have run
      if (g > 0) CU(cudaStreamWaitEvent(st, c->ev_group[g - 1], 0));
      LAUNCH("frame_setup_kernel", frame_setup_kernel<<<1, 32, 0, st>>>(dg, gf, counter));
      CU(cudaEventRecord(c->ev_group[g], st));
      LAUNCH("sort_hist_kernel", sort_hist_kernel<<<dim3(gtiles, gf), 256, 0, st>>>(dg));
      for (int p = 0; p < 8; p++) LAUNCH("sort_pass_kernel", sort_pass_kernel<<<dim3(gtiles, gf), SORT_THREADS, SORT_SMEM_BYTES, st>>>(dg, p));
      LAUNCH("leaf_scan_kernel", leaf_scan_kernel<<<dim3((unsigned)((gn + LEAF_TILE - 1) / LEAF_TILE), gf), LEAF_THREADS, 0, st>>>(dg));
      LAUNCH("leaf_emit_kernel", leaf_emit_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P));
      if (color && prm.color_coding_type == 1) {
        const size_t img_h = gn / 256 + 1, mcu_h = (img_h + 15) / 16, nblk = mcu_h * 16 * 6;
        LAUNCH("jpeg_mcu_kernel", jpeg_mcu_kernel<<<dim3((unsigned)(mcu_h * 16), gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("jpeg_huff_kernel", jpeg_huff_kernel<<<dim3((unsigned)((nblk + HUFF_THREADS - 1) / HUFF_THREADS), gf), HUFF_THREADS, 0, st>>>(dg, c->d_tables));
        const size_t jb = 4 * gn + 8192;
        LAUNCH("jpeg_stuff_kernel", jpeg_stuff_kernel<<<dim3((unsigned)((jb + STUFF_THREADS * STUFF_BYTES - 1) / (STUFF_THREADS * STUFF_BYTES)), gf), STUFF_THREADS, 0, st>>>(dg, c->d_tables));
      }
      if (color && prm.color_coding_type == 2) {
        const unsigned lines_max = (unsigned)(gn / LINE_PX + 1), mcus_max = (unsigned)(gn / 16 + 256);
        LAUNCH("lines_mcu_kernel", lines_mcu_kernel<<<dim3(mcus_max, gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("lines_huff_kernel", lines_huff_kernel<<<dim3(lines_max, gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("lines_offsets_kernel", lines_offsets_kernel<<<gf, 1024, 0, st>>>(dg));
        LAUNCH("lines_copy_kernel", lines_copy_kernel<<<dim3(lines_max, gf), 256, 0, st>>>(dg));
      }
      const size_t hmax = std::max(tree_cap_for(gn), cpay_cap_for(gn));
      mark(g, "leaves", st);
      LAUNCH("hist_kernel", hist_kernel<<<dim3((unsigned)((hmax + 16383) / 16384), 3, gf), 256, 0, st>>>(dg));
      if (c->lps_enc) LAUNCH("rc_encode_lps_kernel", rc_encode_lps_kernel<<<dim3((unsigned)((gf + 31) / 32), 3), 32, c->lps_smem_enc, st>>>(dg, gf, cen, color));
      else LAUNCH("rc_encode_kernel", rc_encode_kernel<<<gf, 96, serial_smem_enc, st>>>(dg, cen, color));
      LAUNCH("assemble_kernel", assemble_kernel<<<dim3(64, gf), 256, 0, st>>>(dg, H));
      mark(g, "encoded", st);
    }
    if (do_dec) {
      DecFrame *dg = dd + f0;
      size_t pmax = 1;
      bool any_h2d = false;
      for (int i = 0; i < gf; i++) {
        const int k = f0 + i;
        pmax = std::max(pmax, pts_cap[k]);
        if (!rt && !din_dev[k] && in_len[k]) {
          if (!any_h2d && g == 0) CU(cudaStreamWaitEvent(c->copy_stream, ev_setup, 0));
          CU(cudaMemcpyAsync((void *)hd[k].in, in[k], in_len[k], cudaMemcpyHostToDevice, c->copy_stream));
          any_h2d = true;
        }
      }
      LAUNCH("zero_region_kernel", zero_region_kernel<DecFrame><<<dim3(64, gf), 256, 0, st>>>(dg));
      if (any_h2d) { CU(cudaEventRecord(c->ev_h2d[g], c->copy_stream)); CU(cudaStreamWaitEvent(st, c->ev_h2d[g], 0)); }
      if (rt) LAUNCH("link_kernel", link_kernel<<<(gf + 63) / 64, 64, 0, st>>>(df + f0, dg, gf));
      // Two entropy stages.  dec_entropy_kernel (a CTA per frame) is the faster one when nothing else runs (decode-only
      // calls: 374 ms against 414 ms for 1024 frames) and when the copies pace the pipeline (host buffers: 935 against
      // 896 Mpoints/s end to end); in a device-resident round trip its 7 CTAs per SM compete with the other groups'
      // encode kernels and the lane-per-stream decoder, 8 frames to a warp on an SM of its own, wins (677 ms against 753 ms).
      if (c->lps_dec < 0 ? (rt && !host_io) : c->lps_dec != 0) {
        // tree layers on the group's stream, speculated colour layers on a side stream at the same time
        cudaStream_t s2 = c->profiling ? st : c->side_streams[g % SIDE_STREAMS];
        const unsigned lps_ctas = (unsigned)((gf + LPS_DEC_FRAMES - 1) / LPS_DEC_FRAMES);
        LAUNCH("dec_head_kernel", dec_head_kernel<<<gf, 32, 0, st>>>(dg));
        if (s2 != st) { CU(cudaEventRecord(c->ev_side[2 * g], st)); CU(cudaStreamWaitEvent(s2, c->ev_side[2 * g], 0)); }
        LAUNCH("rc_decode_lps_kernel<tree>", rc_decode_lps_kernel<true><<<lps_ctas, 32 * (1 + LPS_DEC_FRAMES), sizeof(LpsSmem), st>>>(dg, gf, c->use_ring));
        LAUNCH_S(s2, "rc_decode_lps_kernel<colour>", rc_decode_lps_kernel<false><<<lps_ctas, 32, offsetof(LpsSmem, rg), s2>>>(dg, gf, 0));
        mark(g, "tree", st);
        mark(g, "colour-rc", s2);
        LAUNCH_S(s2, "dec_jpeg_kernel", dec_jpeg_kernel<<<gf, 32, 0, s2>>>(dg));
        mark(g, "jpeg", s2);
        if (s2 != st) { CU(cudaEventRecord(c->ev_side[2 * g + 1], s2)); CU(cudaStreamWaitEvent(st, c->ev_side[2 * g + 1], 0)); }
        LAUNCH("dec_finish_kernel", dec_finish_kernel<<<gf, 32, 0, st>>>(dg));
      } else
      LAUNCH("dec_entropy_kernel", dec_entropy_kernel<<<gf, 96, serial_smem_dec, st>>>(dg, c->use_ring));
      mark(g, "entropy", st);
      LAUNCH("jpeg_destuff_kernel", jpeg_destuff_kernel<<<gf, 1024, 0, st>>>(dg));
      LAUNCH("dec_serial_kernel", dec_serial_kernel<<<steered_grid, 64, 0, st>>>(dg, f0, gf));
      if (prm.color_coding_type == 2) {
        const unsigned lines_max = (unsigned)(pmax / LINE_PX + 2);
        LAUNCH("lines_index_kernel", lines_index_kernel<<<gf, 32, 0, st>>>(dg));
        LAUNCH("lines_decode_kernel", lines_decode_kernel<<<dim3(lines_max, gf), 32, 0, st>>>(dg));
        LAUNCH("lines_idct_kernel", lines_idct_kernel<<<dim3((unsigned)(((size_t)lines_max * LINE_MCU_STRIDE + LINE_MCU_STRIDE) * 6 / 32 + 1), gf), 256, 0, st>>>(dg, c->d_tables));
      }
      const size_t img_h = pmax / 256 + 2, mcu_h = (img_h + 15) / 16, nblocks = mcu_h * 16 * 6;
      LAUNCH("jpeg_idct_kernel", jpeg_idct_kernel<<<dim3((unsigned)((nblocks + 31) / 32), gf), 256, 0, st>>>(dg, c->d_tables));
      LAUNCH("dec_leaves_kernel", dec_leaves_kernel<<<dim3((unsigned)((pmax + 255) / 256), gf), 256, 0, st>>>(dg));        // frames walked by the pipelined walkers
      LAUNCH("dec_points_kernel", dec_points_kernel<<<dim3((unsigned)((pmax + NODE_THREADS - 1) / NODE_THREADS), gf), NODE_THREADS, 0, st>>>(dg));
      mark(g, "decoded", st);
    }
    LAUNCH("publish_kernel", publish_kernel<<<(gf + 63) / 64, 64, 0, st>>>(do_enc ? df + f0 : nullptr, do_dec ? dd + f0 : nullptr, hres + f0, gf));
    CU(cudaEventRecord(c->ev_group[ngroups + g], st));
    CU(cudaGetLastError());
  }

  // ------------------------------------------------------------------ collect: per group wait for its record copies,
  // then move the results out with exact sizes (later groups keep running meanwhile)
  int rc = CCV2_OK;
  for (int g = 0; g < ngroups; g++) {
    cudaStream_t st = c->streams[g % NS];
    cudaEvent_t ev = c->ev_group[ngroups + g];
    CU(cudaEventSynchronize(ev));                            // the group's kernels are done and its FrameResults are in host memory
    const int f0 = g * G, gf = std::min(G, nframes - f0);
    for (int i = 0; i < gf; i++) {
      const int k = f0 + i;
      const FrameResult &r = hres[k];
      bool enc_ok = true;
      if (do_enc) {
        if (out_len) out_len[k] = 0;
        if (r.enc_error) {
          enc_ok = false;
          if (rc == CCV2_OK) {
            rc = (r.enc_error & FERR_DEPTH) ? CCV2_ERR_DEPTH : CCV2_ERR_WORKSPACE;
            char b[96]; snprintf(b, sizeof b, "frame %d: device error bits 0x%x (encode)", k, r.enc_error); c->err = b;
          }
        } else if (r.out_len && out && out[k]) {
          if (r.out_len > out_cap[k]) { if (rc == CCV2_OK) { rc = CCV2_ERR_CAPACITY; c->err = "output buffer too small"; } }
          else {
            CU(cudaMemcpyAsync(out[k], hf[k].stream, r.out_len, out_dev[k] ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
            out_len[k] = r.out_len;
          }
        } else if (r.out_len && !rt) { if (rc == CCV2_OK) { rc = CCV2_ERR_CAPACITY; c->err = "output buffer too small"; } }
        else if (r.out_len && out_len) out_len[k] = r.out_len;      // round trip without a stream buffer: report the size only
      }
      if (do_dec) {
        npts_out[k] = 0;
        if (rt && (!enc_ok || r.out_len == 0)) continue;           // empty frame: nothing was written, nothing to decode
        if (r.dec_error) {
          if (rc == CCV2_OK) {
            rc = (r.dec_error & FERR_OUT_CAP) ? CCV2_ERR_CAPACITY : (r.dec_error & FERR_DEPTH) ? CCV2_ERR_DEPTH : (r.dec_error & FERR_UNSUPPORTED) ? CCV2_ERR_UNSUPPORTED
               : (r.dec_error & (FERR_TREE_CAP | FERR_JPEG_CAP)) ? CCV2_ERR_WORKSPACE : CCV2_ERR_STREAM;
            char b[96]; snprintf(b, sizeof b, "frame %d: device error bits 0x%x (decode)", k, r.dec_error); c->err = b;
          }
          continue;
        }
        npts_out[k] = r.V;
        if (!dout_dev[k] && r.V) CU(cudaMemcpyAsync(pts_out[k], hd[k].out_pts, 32ull * r.V, cudaMemcpyDeviceToHost, st));
      }
    }
    // full frame records (metrics, debug hook) come back last on this stream
    if (do_enc) CU(cudaMemcpyAsync(hf + f0, df + f0, sizeof(EncFrame) * gf, cudaMemcpyDeviceToHost, st));
    if (do_dec) CU(cudaMemcpyAsync(hd + f0, dd + f0, sizeof(DecFrame) * gf, cudaMemcpyDeviceToHost, st));
    mark(g, "copied", st);
    CU(cudaEventRecord(ev, st));
    CU(cudaStreamWaitEvent(ms, ev, 0));
  }
  CU(cudaEventRecord(c->ev_end, ms));
  CU(cudaEventSynchronize(c->ev_end));
  CU(cudaEventElapsedTime(&c->device_ms, c->ev_start, c->ev_end));
  if (c->trace) {
    fprintf(stderr, "ccv2 trace mode=%d frames=%d groups=%d total %.1f ms\n", mode, nframes, ngroups, c->device_ms);
    for (int g = 0; g < ngroups; g++) {
      fprintf(stderr, "  group %2d:", g);
      for (size_t k = 0; k < c->trace_marks.size(); k++) if (c->trace_marks[k].group == g) {
        float t = -1; cudaEventElapsedTime(&t, c->ev_start, c->ev_trace[k]);
        fprintf(stderr, " %s %.1f |", c->trace_marks[k].label, t);
      }
      fprintf(stderr, "\n");
    }
  }
  if (c->trace) {                                            // how evenly did the serial CTAs spread?  (frame records are back on the host)
    for (int dir = 0; dir < 2; dir++) {
      if (dir == 0 ? !do_enc : !do_dec) continue;
      std::vector<int> per_sm(256, 0);
      for (int i = 0; i < nframes; i++) per_sm[(dir == 0 ? hf[i].serial_sm : hd[i].serial_sm) & 255]++;
      std::vector<int> hist(64, 0); int mx = 0;
      for (int s2 = 0; s2 < c->n_sm; s2++) { hist[std::min(63, per_sm[s2])]++; mx = std::max(mx, std::min(63, per_sm[s2])); }
      fprintf(stderr, "  %s serial CTAs per SM (cap %d):", dir == 0 ? "encode" : "decode", serial_cap);
      for (int k = 0; k <= mx; k++) fprintf(stderr, " %dx%d", hist[k], k);
      fprintf(stderr, "\n");
    }
  }
  prof_collect(c);
  c->launches = launches;
  if (do_enc) {
    CU(cudaMemcpy(&c->frame_id, counter, 4, cudaMemcpyDeviceToHost));
    c->enc_host.assign(hf, hf + nframes);
    for (int i = nframes - 1; i >= 0; i--) if (hf[i].out_len) { for (int k = 0; k < 3; k++) c->metrics[k] = hf[i].coded[k]; break; }
  } else {
    for (int i = nframes - 1; i >= 0; i--) if (!hd[i].error) { for (int k = 0; k < 3; k++) c->metrics[k] = hd[i].coded[k]; c->frame_id = hd[i].frame_id; break; }
  }
  return rc;
}

int ccv2_encode_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                      void *const *out, const size_t *out_cap, size_t *out_len) {
  if (!c || nframes < 0 || (nframes && (!pts || !npts || !out || !out_cap || !out_len))) return CCV2_ERR_ARG;
  return run_batch(c, 0, nframes, pts, npts, out, out_cap, out_len, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ccv2_decode_batch(ccv2_codec *c, int nframes, const void *const *in, const size_t *in_len,
                      void *const *pts_out, const size_t *pts_cap, size_t *npts_out) {
  if (!c || nframes < 0 || (nframes && (!in || !in_len || !pts_out || !pts_cap || !npts_out))) return CCV2_ERR_ARG;
  return run_batch(c, 1, nframes, nullptr, nullptr, nullptr, nullptr, nullptr, in, in_len, pts_out, pts_cap, npts_out);
}

int ccv2_roundtrip_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                         void *const *out, const size_t *out_cap, size_t *out_len,
                         void *const *pts_out, const size_t *pts_cap, size_t *npts_out) {
  if (!c || nframes < 0 || (nframes && (!pts || !npts || !pts_out || !pts_cap || !npts_out))) return CCV2_ERR_ARG;
  if (out && (!out_cap || !out_len)) return CCV2_ERR_ARG;
  return run_batch(c, 2, nframes, pts, npts, out, out_cap, out_len, nullptr, nullptr, pts_out, pts_cap, npts_out);
}

int ccv2_get_output_cloud(ccv2_codec *c, int frame, void *points_out, size_t cap_points, size_t *npoints) {
  if (!c || frame < 0 || frame >= (int)c->enc_host.size() || !npoints) return CCV2_ERR_ARG;
  if (c->work_mode != 0) { c->err = "the output cloud is only available right after ccv2_encode_batch"; return CCV2_ERR_UNSUPPORTED; }
  CU(cudaSetDevice(c->device));
  const EncFrame &f = c->enc_host[frame];
  if (f.error) { *npoints = 0; return CCV2_OK; }
  *npoints = f.V;
  if (f.V == 0) return CCV2_OK;
  if (f.V > cap_points || !points_out) return CCV2_ERR_CAPACITY;
  const bool dev = is_device_ptr(points_out);
  uint8_t *dst = (uint8_t *)points_out;
  if (!dev) { CU(c->out_cloud.ensure(32ull * f.V)); dst = (uint8_t *)c->out_cloud.p; }
  output_cloud_kernel<<<(f.V + 255) / 256, 256, 0, c->main_stream>>>(f, c->last_enc_params, dst);
  CU(cudaGetLastError());
  if (!dev) CU(cudaMemcpyAsync(points_out, dst, 32ull * f.V, cudaMemcpyDeviceToHost, c->main_stream));
  CU(cudaStreamSynchronize(c->main_stream));
  return CCV2_OK;
}

int ccv2_debug_fetch(ccv2_codec *c, int frame, int what, void *host_buf, size_t cap, size_t *len) {
  if (!c || frame < 0 || frame >= (int)c->enc_host.size() || !len) return CCV2_ERR_ARG;
  CU(cudaSetDevice(c->device));
  const EncFrame &f = c->enc_host[frame];
  const void *src = nullptr; size_t n = 0;
  if ((what == 0 || what == 2 || what == 4) && c->work_mode != 0) { c->err = "encode intermediates are only available right after ccv2_encode_batch"; return CCV2_ERR_UNSUPPORTED; }
  ccv2_frame_info info;
  switch (what) {
    case 0: src = f.leaf_key; n = (size_t)f.V * 8; break;
    case 1: src = f.tree; n = f.B; break;
    case 2: src = f.avg; n = (size_t)f.V * 3; break;
    case 3: src = f.cpay; n = f.ncolor; break;
    case 4: src = f.vals[f.npasses & 1]; n = (size_t)f.n_finite * 4; break;
    case 5:
      memset(&info, 0, sizeof info);
      info.depth = f.depth; info.n_finite = f.n_finite; info.n_leaves = f.V; info.n_tree_bytes = f.B; info.n_color_bytes = f.ncolor; info.error = f.error;
      for (int a = 0; a < 3; a++) { info.bb_min[a] = f.bmin[a]; info.bb_max[a] = f.bmax[a]; info.coded[a] = f.coded[a]; }
      *len = sizeof info;
      if (cap < sizeof info || !host_buf) return CCV2_ERR_CAPACITY;
      memcpy(host_buf, &info, sizeof info);
      return CCV2_OK;
    default: return CCV2_ERR_ARG;
  }
  *len = n;
  if (n > cap || !host_buf) return CCV2_ERR_CAPACITY;
  if (n) CU(cudaMemcpy(host_buf, src, n, cudaMemcpyDeviceToHost));
  return CCV2_OK;
}

}  // extern "C"

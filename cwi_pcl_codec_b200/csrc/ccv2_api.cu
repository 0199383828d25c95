// ccv2_api.cu -- C ABI (include/ccv2.h) and host-side orchestration of the CUDA pipelines.
//
// Execution model (round 2): a STREAMING pipeline over two rings of device workspaces.
//   * A call's frames are cut into groups of G <= 32 frames.  Group q (a sequence number that runs across calls) takes
//     front-end set q mod R_fe (sort buffers, leaf arrays, JPEG scratch, staging for host inputs: ~53-85 B/point, needed
//     for a few milliseconds) and long-lived set q mod R_ll (tree bytes, colour payload, the stream, the decoder's
//     workspace, staging for host outputs: ~36-68 B/point, needed for the 0.1-0.4 s the serial range-coder stages take).
//     A set is handed on by CUDA events, so the host never waits inside a call: everything is enqueued up front, and the
//     copy engines, the parallel front-end kernels and the latency-bound serial kernels of different groups overlap.
//   * Host inputs travel on ONE copy stream in group order (cudaMemcpyAsync from pinned or pageable memory); finished
//     streams leave through export_kernel (device memory, or zero-copy stores into pinned host memory); decoded clouds
//     are written straight into device destinations, or into the set's staging area and from there by ONE copy-engine
//     transfer per frame into pinned host memory.  Pageable host destinations get a per-call staging area and are
//     copied when the call is collected.
//   * ccv2_submit_* enqueue a call and return a ticket; ccv2_wait collects it.  Three calls may be in flight, sharing the
//     rings: the next call's uploads and front-ends run while the previous call's serial stages drain, which is what
//     hides the pipeline's fill and drain (0.3-0.5 s against 0.6 s of PCIe time per 1024-frame call).
//   * No host synchronisation inside a group: all sizes (depth, V, B, J, stream length) are device-side values in the
//     frame records; the records come back once per call, stored into their pinned host copy by a kernel (records_home_kernel:
//     a copy-engine transfer would queue behind the cloud transfers of later calls).
#include "../../include/ccv2.h"
#include "common.cuh"
#include "enc_kernels.cuh"
#include "jpeg_enc_kernels.cuh"
#include "entropy_kernels.cuh"
#include "dec_kernels.cuh"
#include "dec_lps_kernels.cuh"
#include "lines_kernels.cuh"
#include "lines_dec_kernels.cuh"
#include "quality_kernels.cuh"
#include "tile_kernels.cuh"
#include "detail_kernels.cuh"
#include "inter_kernels.cuh"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static_assert(sizeof(EncFrame) % 8 == 0 && sizeof(DecFrame) % 8 == 0, "records_home_kernel moves the frame records in 8-byte words");

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void *p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes, bool slack = true) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + (slack ? std::min(bytes / 8, (size_t)64 << 20) : 0) + 4096;   // slack so slowly growing batches do not reallocate every call
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

// bump allocator over a device region: first pass (base == nullptr) only measures
struct Carver {
  uint8_t *base; size_t off = 0;
  explicit Carver(void *b) : base((uint8_t *)b) {}
  template <typename T> T *take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T *r = base ? (T *)(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  size_t end() const { return (off + 255) & ~size_t(255); }
};

constexpr int MAX_STREAMS = 16;         // work streams: long-lived set s always runs on stream s mod 16 -- a group waits for the previous user of its set anyway,
                                         // so with at most 16 sets no group ever queues behind a stream that is busy with an unrelated one
constexpr int SIDE_STREAMS = 8;          // 16 work + 8 side + copy, D2H, control, collect, end = 29 streams <= the 32 hardware queues: with more, streams alias
                                         // onto queues and a group's front-end kernels sit behind another group's 0.4 s serial kernel (measured: 48 streams 2270, 34 streams 2340 Mpoints/s)          // group + side + copy + control streams stay within 32 hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, see ccv2.h)
constexpr int MAX_GROUP = 128;           // frames per group.  The serial range-coder kernels are latency bound (0.1-0.4 s per launch whatever the frame count),
                                         // so throughput = frames per launch x launches in flight: large groups, one stream each (measured: 32-frame groups 650, 64 1040-1270 Mpoints/s)
constexpr int N_CALLS = 4;               // call contexts: three user calls in flight + one for the retry of a frame that overflowed its workspace

struct DoneRc { int ticket, rc; };
enum PtrKind : char { PK_NONE = 0, PK_DEVICE = 1, PK_PINNED = 2, PK_PAGEABLE = 3 };

// A ring of equally sized workspace sets (G frames each) that groups take in sequence; ev_free[set] is recorded when the
// set's current user is done with it, and the next user's stream (and the copy stream, if it uploads into the set) waits on it.
struct Ring {
  DevBuf buf; size_t frame_bytes = 0; int G = 0, nsets = 0; uint64_t seq = 0; bool mem_limited = false;
  std::vector<cudaEvent_t> ev_free; std::vector<char> used;
  uint8_t *frame_base(int set, int i) const { return (uint8_t *)buf.p + ((size_t)set * G + i) * frame_bytes; }
};

struct CallCtx {
  bool busy = false; int ticket = 0, mode = 0, nframes = 0, G = 1, ngroups = 0, rc_early = 0;
  bool boost = false, timed = false;
  DevBuf enc_frames, dec_frames, stage;                     // frame records (+ histograms); per-call staging for pageable destinations
  HostBuf h_frames, h_dframes;
  cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_setup = nullptr;
  std::vector<cudaEvent_t> ev_h2d, ev_side, ev_done, ev_fin, ev_enc, ev_hop;
  // per-frame bookkeeping of the call (the caller's arrays must stay alive until the call is collected)
  std::vector<char> in_kind, out_kind, pts_kind;            // PtrKind of pts[i] / in[i], out[i], pts_out[i]
  std::vector<size_t> stage_off_stream, stage_off_pts;      // offsets into `stage` (pageable destinations)
  std::vector<int> fe_set, ll_set;                          // ring sets of every group
  const void *const *pts = nullptr; const size_t *npts = nullptr; void *const *out = nullptr; const size_t *out_cap = nullptr; size_t *out_len = nullptr;
  const void *const *in = nullptr; const size_t *in_len = nullptr; void *const *pts_out = nullptr; const size_t *pts_cap = nullptr; size_t *npts_out = nullptr;
  uint64_t launches = 0; float device_ms = 0;
  uint64_t fe_seq0 = 0, ll_seq0 = 0; int fe_nsets = 0, ll_nsets = 0;   // ring state the call started from (is a frame's workspace still intact afterwards?)
  struct TraceMark { int group; const char *label; };
  std::vector<cudaEvent_t> ev_trace; std::vector<TraceMark> trace_marks;
};

}  // namespace

struct ccv2_codec {
  ccv2_params prm;
  int device = 0;
  int n_sm = 148;
  int trace = 0;                          // CCV2_TRACE=1: print per-group timeline (ms since call start) to stderr
  int use_ring = 1;                       // pipeline the DFS walk behind the range decoder (CCV2_NO_RING=1 disables: debugging)
  int n_streams = 0, group = 0;           // CCV2_STREAMS / CCV2_GROUP overrides (0 = automatic)
  int n_side = SIDE_STREAMS;              // side streams in use (CCV2_SIDE)
  int allow_packed = 1;                   // packed sort elements (CCV2_PACKED=0: always (code, index) pairs)
  int enc_reserve = 1;                    // lane-per-stream encoder CTAs reserve half an SM's shared memory (one CTA per SM); CCV2_ENC_RESERVE=0 turns it off
  int inflight_max = 2048;                // frames the long-lived ring may hold (CCV2_INFLIGHT); memory permitting
  int fe_frames = 0;                      // frames the front-end ring holds (CCV2_FE_FRAMES; 0 = 256, or 512 when host inputs are staged in it); at least two sets
  cudaStream_t main_stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr, fin_stream = nullptr;
  cudaStream_t end_stream = nullptr;      // gathers every call's group events (calls complete in order); on the control stream that wait would hold back the next call's set-up
  cudaStream_t streams[MAX_STREAMS] = {};
  cudaStream_t ser_streams[MAX_STREAMS] = {};     // green contexts on: the same slots on the serial partition's SMs; off: aliases of streams[]
  int green_sms = 0;                      // SMs set aside for the latency-bound range-coder kernels (CCV2_GREEN; 0 = no partition)
  CUgreenCtx green_ser = nullptr, green_par = nullptr;
  CUresult (*drv_green_stream_create)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*drv_green_destroy)(CUgreenCtx) = nullptr;
  cudaStream_t side_streams[SIDE_STREAMS] = {};   // colour layer of a group, concurrent with its tree layer (lane-per-stream decoder)
  int lps_dec = -1;                       // lane-per-stream range decoder: -1 auto (round trips only), 0 off, 1 on (CCV2_LPS_DEC)
  cudaEvent_t ev_id_chain = nullptr; bool id_chain_used = false;     // frame ids are sequential: a group's setup waits for the previous group's
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr; bool timer_on = false; // ccv2_timer_* (CCV2_TRACE prints times since the timer's start when it runs)
  JpegTables *d_tables = nullptr;
  uint32_t *d_frame_counter = nullptr;
  size_t lps_smem_enc = 0;                // dynamic shared memory that keeps the lane-per-stream coder CTAs one to an SM
  int lps_enc = 1;                        // lane-per-stream range encoder (CCV2_LPS_ENC=0: one warp per stream)
  int serial_cap = 0;                     // serial CTAs per SM; 0: ceil(frames in the call / SMs)  (CCV2_CAP overrides, -1 disables the cap)
  size_t smem_sm = 0, smem_static_enc = 0, smem_static_dec = 0;   // shared memory per SM; static use of the two serial kernels
  uint32_t frame_id = 0;
  Ring fe, ll;
  CallCtx calls[N_CALLS];
  int next_ticket = 1, user_calls = 0; uint64_t stage_seq = 0;
  std::vector<DoneRc> done;              // results of collected calls, by ticket (ccv2_wait after the fact)
  int last_mode = -1;                      // mode of the call collected last
  EncParams last_enc_params;               // of the last encode (ccv2_get_output_cloud)
  DevBuf out_cloud;                        // staging for ccv2_get_output_cloud into host memory
  DevBuf tile_pts, tile_aux;               // tile mode: the partitioned frame, counters
  DevBuf inter_ws, inter_batch;            // inter-frame path: grids, macroblock results, staging; the unpredicted points of a batch of delta frames (ccv2_inter.cuh)
  ccv2_codec *intra_child = nullptr;       // the fresh intra coder a delta frame's unpredicted points go through (impl.hpp:1089-1101)
  cudaEvent_t inter_ev0 = nullptr, inter_ev1 = nullptr;
  float mb_percentage = 0.f, mb_convergence = 0.f;   // getMacroBlockPercentage / getMacroBlockConvergencePercentage of the last delta frame
  std::vector<EncFrame> enc_host;          // host mirror of the last encode call's frame records (with device pointers)
  std::vector<char> enc_host_valid;        // per frame: its ring sets were not handed on to a later group
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

PtrKind ptr_kind(const void *p) {
  if (!p) return PK_NONE;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return PK_PAGEABLE; }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return PK_DEVICE;
  if (a.type == cudaMemoryTypeHost) return PK_PINNED;
  return PK_PAGEABLE;
}
bool is_device_ptr(const void *p) { return ptr_kind(p) == PK_DEVICE; }

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

// ---- capacity rules (bytes) for a frame of n points.  The defaults hold every cloud the benchmarks and tests produce
// (uniform random 1M points at 11 bits: 3.6 tree bytes per point); a frame that overflows one (sparse deep octrees reach
// depth * n tree bytes) is flagged on the device and encoded again, alone, with the `boost` bounds, which cannot overflow.
size_t tree_cap_for(size_t n, bool boost) {
  const size_t full = 22 * n + 1024;                                    // B <= depth * V, depth <= 21
  const size_t v = boost ? full : std::max(4 * n, std::min(full, (size_t)1 << 20)) + 1024;
  return (v + 255) & ~size_t(255);
}
size_t cpay_cap_for(size_t n, bool boost) { return ((boost ? 8 * n + 65536 : 3 * n + n / 4 + 8192) + 255) & ~size_t(255); }   // raw colour types need 3 n
size_t cen_cap_for(size_t n) { return ((3 * n + 256) + 255) & ~size_t(255); }
size_t rc_cap_for(size_t raw_cap) { return ((raw_cap + raw_cap / 8 + 4096) + 255) & ~size_t(255); }
// detail mode: int-coder table entries (u64) and the coded counts; 65536 entries hold voxels of up to 32767 points
size_t itab_cap_for(size_t n, bool boost) { return boost ? 4 * (n + 2) : std::min<size_t>(4 * (n + 2), 65536); }
size_t rc_int_cap_for(size_t n, bool boost) { return ((9 + 8 * itab_cap_for(n, boost) + 8 * n + 64) + 255) & ~size_t(255); }
size_t diff_cap_for(size_t n) { return ((3 * n + 64) + 255) & ~size_t(255); }
// The stream: with `boost` the sum of every layer's worst case; by default half of that for the tree and colour layers
// (occupancy bytes code to ~0.6 of their size, a stream that does not fit is flagged and retried like the other bounds).
size_t stream_cap_for(size_t n, bool cen, bool boost, bool detail = false) {
  const size_t tc = rc_cap_for(tree_cap_for(n, boost)) + rc_cap_for(cpay_cap_for(n, boost));
  return FRAME_HDR_BYTES + 8 + (boost || n < 65536 ? tc : tc / 2 + 65536) + (cen ? 4 + rc_cap_for(cen_cap_for(n)) : 0) + 8 +
         (detail ? 8 + rc_int_cap_for(n, boost) + 2 * (8 + rc_cap_for(diff_cap_for(n))) : 0);
}

// front-end workspace of one frame (ring `fe`): dead once the colour payload is complete
size_t carve_fe(uint8_t *base, size_t n, EncFrame *f, const ccv2_params &prm, bool boost, bool host_in, size_t *zero_off, size_t *zero_bytes) {
  Carver cv(base);
  const uint32_t tiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE) + 1;
  const uint32_t scan_tiles = (uint32_t)(n / 1024) + 8;
  const size_t img_h = n / 256 + 1, mcu_h = (img_h + 15) / 16;
  const bool lines = prm.color_coding_type == 2;
  const size_t lines_cap = n / LINE_PX + 2;
  const size_t jbits_words = lines ? lines_cap * LINE_BITS_WORDS : ((cpay_cap_for(n, boost) / 4 + 63) & ~size_t(63));
  // --- zero-initialised region first
  uint32_t *ghist = cv.take<uint32_t>(8 * 256);
  uint32_t *sort_status = cv.take<uint32_t>((size_t)8 * tiles * 256);
  uint64_t *scan_status = cv.take<uint64_t>((size_t)4 * scan_tiles);
  uint32_t *jbits = cv.take<uint32_t>(jbits_words + 16);
  const size_t z1 = cv.end();
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
  uint32_t *cd_off = cv.take<uint32_t>(prm.do_voxel_grid_downsampling ? 4 : n + 8);
  uint8_t *pts_stage = cv.take<uint8_t>(host_in ? 32 * n + 32 : 16);
  if (f) {
    f->cd_off = cd_off;
    f->ghist = ghist; f->sort_status = sort_status; f->tiles_max = tiles; f->scan_status = scan_status; f->scan_tiles_max = scan_tiles;
    f->jbits_buf = jbits; f->jbits_cap_words = (uint32_t)jbits_words;
    f->keys[0] = k0; f->keys[1] = k1; f->vals[0] = v0; f->vals[1] = v1;
    f->leaf_key = leaf_key; f->leaf_start = leaf_start; f->leaf_off = leaf_off; f->first_new = first_new;
    f->avg = avg; f->coef = coef;
    f->line_slots = line_slots; f->line_len = line_len; f->line_off = line_off; f->lines_cap = lines ? (uint32_t)lines_cap : 0;
    if (host_in) f->pts = pts_stage;
  }
  if (zero_off) *zero_off = 0;
  if (zero_bytes) *zero_bytes = z1;
  return cv.end();
}
// long-lived encoder buffers of one frame (ring `ll`): the stream first (a round trip's decoder reads it while its own
// workspace lies over the rest, which is dead once the stream is assembled); *stream_end = offset where the rest starts
size_t carve_enc_ll(uint8_t *base, size_t n, EncFrame *f, bool cen, bool boost, size_t *stream_end, bool detail) {
  Carver cv(base);
  uint8_t *stream = cv.take<uint8_t>(stream_cap_for(n, cen, boost, detail));
  if (stream_end) *stream_end = cv.end();
  uint8_t *tree = cv.take<uint8_t>(tree_cap_for(n, boost));
  uint8_t *cenb = cv.take<uint8_t>(cen ? cen_cap_for(n) : 256);
  uint8_t *cpay = cv.take<uint8_t>(cpay_cap_for(n, boost));
  uint8_t *rc0 = cv.take<uint8_t>(cen ? rc_cap_for(cen_cap_for(n)) : 256);
  uint8_t *rc1 = cv.take<uint8_t>(rc_cap_for(cpay_cap_for(n, boost)));
  uint32_t *counts = cv.take<uint32_t>(detail ? n + 8 : 4);
  uint8_t *pdiff = cv.take<uint8_t>(detail ? diff_cap_for(n) : 16), *cdiff = cv.take<uint8_t>(detail ? diff_cap_for(n) : 16);
  uint64_t *itab = cv.take<uint64_t>(detail ? itab_cap_for(n, boost) + 8 : 2);
  uint8_t *rc_int = cv.take<uint8_t>(detail ? rc_int_cap_for(n, boost) : 16);
  uint8_t *rc2 = cv.take<uint8_t>(detail ? rc_cap_for(diff_cap_for(n)) : 16), *rc3 = cv.take<uint8_t>(detail ? rc_cap_for(diff_cap_for(n)) : 16);
  if (f) {
    f->counts = counts; f->pdiff = pdiff; f->cdiff = cdiff; f->itab = itab; f->itab_cap = detail ? (uint32_t)itab_cap_for(n, boost) : 0;
    f->rc_int = rc_int; f->rc_int_cap = detail ? (uint32_t)rc_int_cap_for(n, boost) : 0;
    f->rc_tmp[2] = rc2; f->rc_tmp[3] = rc3; f->rc_tmp_cap[2] = f->rc_tmp_cap[3] = detail ? (uint32_t)rc_cap_for(diff_cap_for(n)) : 0;
    f->tree = tree; f->tree_cap = (uint32_t)tree_cap_for(n, boost) - 64; f->cen = cenb; f->cpay = cpay; f->cpay_cap = (uint32_t)cpay_cap_for(n, boost) - 64;
    f->rc_tmp[0] = rc0; f->rc_tmp[1] = rc1; f->rc_tmp_cap[0] = (uint32_t)(cen ? rc_cap_for(cen_cap_for(n)) : 0); f->rc_tmp_cap[1] = (uint32_t)rc_cap_for(cpay_cap_for(n, boost));
    f->stream = stream; f->stream_cap = stream_cap_for(n, cen, boost, detail);
  }
  return cv.end();
}
// decoder workspace of one frame: ncap = voxels it may hold, tcap / ccap = tree / colour payload bytes
size_t carve_dec(uint8_t *base, size_t ncap, size_t tcap, size_t ccap, DecFrame *f, size_t *zero_bytes, bool lines, size_t dpts /* detail mode: points the frame may hold, 0 = no detail buffers */, bool boost, bool cen) {
  Carver cv(base);
  const size_t lines_cap = lines ? ncap / LINE_PX + 2 : 0;
  const size_t img_h = ncap / 256 + 2, mcu_h = (img_h + 15) / 16, nblocks = lines ? (lines_cap * LINE_MCU_STRIDE + LINE_MCU_STRIDE) * 6 : mcu_h * 16 * 6;
  const uint32_t scan_tiles = (uint32_t)(ncap / NODE_THREADS) + 8;
  uint64_t *scan_status = cv.take<uint64_t>(3 * (size_t)scan_tiles);
  int16_t *coef = cv.take<int16_t>(nblocks * 64 + 64);
  const size_t z1 = cv.end();
  uint8_t *tree = cv.take<uint8_t>(tcap + 64);
  uint8_t *cenb = cv.take<uint8_t>(cen ? cen_cap_for(ncap) : 256);
  uint8_t *col = cv.take<uint8_t>(ccap + 64);
  // The pipelined walker records level depth-2 branches (l2_*), the fallback walkers bottom-level branches (node_*): a
  // frame uses one or the other, so they share their memory.
  const size_t rec_off = cv.end();
  uint64_t *l2_prefix = cv.take<uint64_t>(ncap + 8);
  uint8_t *l2_mask = cv.take<uint8_t>(ncap + 8);
  uint32_t *l2_off = cv.take<uint32_t>(ncap + 8);
  const size_t rec_end = cv.end();
  cv.off = rec_off;
  uint64_t *node_prefix = cv.take<uint64_t>(ncap + 8);
  uint8_t *node_byte = cv.take<uint8_t>(ncap + 8);
  cv.off = std::max(rec_end, cv.end());
  uint8_t *planes = cv.take<uint8_t>(mcu_h * 16 * 256 * 3 / 2 + 256);
  uint16_t *qt = cv.take<uint16_t>(128);
  uint8_t *scan = cv.take<uint8_t>(ccap + 64);
  uint32_t *line_off = cv.take<uint32_t>(lines_cap + 4), *line_len = cv.take<uint32_t>(lines_cap + 4), *line_w = cv.take<uint32_t>(lines_cap + 4);
  uint16_t *line_qt = cv.take<uint16_t>(lines_cap * 128 + 128);
  uint8_t *line_planes = cv.take<uint8_t>(lines_cap * 8192 + 256);
  uint32_t *counts = cv.take<uint32_t>(dpts ? ncap + 8 : 4);
  uint64_t *dleaf_key = cv.take<uint64_t>(dpts ? ncap + 8 : 2);
  uint8_t *pdiff = cv.take<uint8_t>(dpts ? diff_cap_for(dpts) : 16), *cdiff = cv.take<uint8_t>(dpts ? diff_cap_for(dpts) : 16);
  uint64_t *itab = cv.take<uint64_t>(dpts ? itab_cap_for(dpts, boost) + 8 : 2);
  if (f) {
    f->counts = counts; f->counts_cap = dpts ? (uint32_t)ncap : 0; f->dleaf_key = dleaf_key; f->pdiff = pdiff; f->cdiff = cdiff; f->pdiff_cap = dpts ? 3 * dpts : 0;
    f->itab = itab; f->itab_cap = dpts ? (uint32_t)itab_cap_for(dpts, boost) : 0;
    f->scan_status = scan_status; f->scan_tiles_max = scan_tiles; f->coef = coef; f->coef_cap_blocks = (uint32_t)nblocks;
    f->tree = tree; f->tree_cap = (uint32_t)std::min(tcap, (size_t)0xFFFFFF00u); f->cen = cenb; f->cen_cap = cen ? (uint32_t)cen_cap_for(ncap) - 64 : 192;
    f->col = col; f->col_cap = (uint32_t)std::min(ccap, (size_t)0xFFFFFF00u);
    f->node_prefix = node_prefix; f->node_byte = node_byte; f->node_cap = (uint32_t)ncap;
    f->l2_prefix = l2_prefix; f->l2_mask = l2_mask; f->l2_off = l2_off;
    f->planes = planes; f->planes_cap = (uint32_t)(mcu_h * 16 * 256 * 3 / 2); f->qt = qt; f->scan = scan;
    f->lines_cap = (uint32_t)lines_cap; f->line_off = line_off; f->line_len = line_len; f->line_w = line_w; f->line_qt = line_qt; f->line_planes = line_planes;
  }
  if (zero_bytes) *zero_bytes = z1;
  return cv.end();
}

int check_params(const ccv2_params *p, std::string &err) {
  if (!p) { err = "null params"; return CCV2_ERR_ARG; }
  if (p->profile != CCV2_MANUAL_CONFIGURATION) { err = "only MANUAL_CONFIGURATION is implemented"; return CCV2_ERR_UNSUPPORTED; }
  if (!(p->octree_resolution > 0)) { err = "octree_resolution must be > 0"; return CCV2_ERR_ARG; }
  if (p->color_coding_type > 3) { err = "unknown colorCodingType"; return CCV2_ERR_ARG; }
  if (p->color_bit_resolution > 8) { err = "colorBitResolution > 8"; return CCV2_ERR_ARG; }
  return CCV2_OK;
}

// syncToHeader + the fields a decoder needs to size its workspace, on a HOST copy of the first bytes of a stream
// (impl.hpp:1660-1676).  Returns false when no header lies inside the window.
struct PeekInfo { uint64_t point_count, B; uint8_t voxel_grid, with_color, centroid; uint32_t cct; };
bool peek_header(const uint8_t *b, size_t len, PeekInfo &o) {
  static const char id2[] = "<PCL-OCT-CODECV2-COMPRESSED>", id1[] = "<PCL-OCT-COMPRESSED>";
  size_t pos = 0; unsigned hp = 0;
  while (hp < 28) {
    if (pos >= len) return false;
    const uint8_t ch = b[pos++];
    if (ch == 0xFF) return false;                           // (char)0xFF == EOF quirk, SURVEY App. C-9
    if (ch != (uint8_t)id2[hp++]) hp = ((uint8_t)id2[0] == ch) ? 1 : 0;
  }
  hp = 0;
  while (hp < 20) {
    if (pos >= len) return false;
    const uint8_t ch = b[pos++];
    if (ch != (uint8_t)id1[hp++]) hp = ((uint8_t)id1[0] == ch) ? 1 : 0;
  }
  if (pos + 92 + 8 > len) return false;
  const uint8_t *h = b + pos;
  o.voxel_grid = h[5]; o.with_color = h[6];
  memcpy(&o.point_count, h + 7, 8);
  o.centroid = h[80];
  memcpy(&o.cct, h + 83, 4);
  memcpy(&o.B, h + 92, 8);
  return true;
}

}  // namespace

// ================================================================================================ handle life cycle
static void drain(ccv2_codec *c) {                           // waits for everything the codec has enqueued
  cudaStreamSynchronize(c->end_stream);
  cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->d2h_stream);
  for (int i = 0; i < MAX_STREAMS; i++) { cudaStreamSynchronize(c->streams[i]); if (c->ser_streams[i] != c->streams[i]) cudaStreamSynchronize(c->ser_streams[i]); }
  for (int i = 0; i < SIDE_STREAMS; i++) cudaStreamSynchronize(c->side_streams[i]);
  cudaStreamSynchronize(c->main_stream); cudaStreamSynchronize(c->fin_stream);
}
static int finish_call(ccv2_codec *c, CallCtx &x);
static int finish_all(ccv2_codec *c) {                       // collects calls still in flight, oldest first
  int rc = CCV2_OK;
  for (;;) {
    CallCtx *oldest = nullptr;
    for (auto &x : c->calls) if (x.busy && (!oldest || x.ticket < oldest->ticket)) oldest = &x;
    if (!oldest) return rc;
    const int r = finish_call(c, *oldest);
    if (rc == CCV2_OK) rc = r;
  }
}

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
  if (const char *s = getenv("CCV2_GROUP")) c->group = std::max(0, std::min(256, atoi(s)));
  if (const char *s = getenv("CCV2_PACKED")) c->allow_packed = atoi(s) != 0;
  if (const char *s = getenv("CCV2_ENC_RESERVE")) c->enc_reserve = atoi(s) != 0;
  if (const char *s = getenv("CCV2_SIDE")) c->n_side = std::max(1, std::min(SIDE_STREAMS, atoi(s)));
  if (const char *s = getenv("CCV2_INFLIGHT")) c->inflight_max = std::max(1, atoi(s));
  if (const char *s = getenv("CCV2_FE_FRAMES")) c->fe_frames = std::max(1, atoi(s));
  auto fail = [&](cudaError_t ee, const char *what) { g_create_error = std::string(what) + ": " + cudaGetErrorString(ee); ccv2_destroy(c); return CCV2_ERR_CUDA; };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  if ((e = cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail(e, "cudaDeviceGetAttribute");
  for (cudaStream_t *s : { &c->main_stream, &c->copy_stream, &c->d2h_stream, &c->fin_stream, &c->end_stream })
    if ((e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  // SM partition (CUDA green contexts): the serial range-coder kernels are latency bound -- one warp per 8-32 frames, a few
  // hundred warps in all -- and lose up to half their speed when bandwidth-bound front-end kernels of other groups share
  // their SMs' schedulers.  With CCV2_GREEN=n they get n SMs of their own and everything else runs on the rest.
  if (const char *s = getenv("CCV2_GREEN")) c->green_sms = std::max(0, std::min(c->n_sm - 16, atoi(s)));
  if (c->green_sms > 0) {
    cudaFree(0);
    // driver entry points through the runtime (the library does not link libcuda: it must load on hosts without a driver)
    struct Drv {
      CUresult (*DeviceGet)(CUdevice *, int); CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType);
      CUresult (*SplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
      CUresult (*GenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int); CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    } drv = {};
    auto sym = [](const char *name, void **fp) { cudaDriverEntryPointQueryResult q; return cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fp; };
    CUdevice dev; CUdevResource all, grp, rest; unsigned int ng = 1; CUdevResourceDesc d0, d1;
    bool ok = sym("cuDeviceGet", (void **)&drv.DeviceGet) && sym("cuDeviceGetDevResource", (void **)&drv.DeviceGetDevResource) && sym("cuDevSmResourceSplitByCount", (void **)&drv.SplitByCount) &&
              sym("cuDevResourceGenerateDesc", (void **)&drv.GenerateDesc) && sym("cuGreenCtxCreate", (void **)&drv.GreenCtxCreate) &&
              sym("cuGreenCtxStreamCreate", (void **)&c->drv_green_stream_create) && sym("cuGreenCtxDestroy", (void **)&c->drv_green_destroy);
    ok = ok && drv.DeviceGet(&dev, device) == CUDA_SUCCESS && drv.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
         drv.SplitByCount(&grp, &ng, &all, &rest, 0, (unsigned)c->green_sms) == CUDA_SUCCESS && ng == 1 &&
         drv.GenerateDesc(&d0, &grp, 1) == CUDA_SUCCESS && drv.GenerateDesc(&d1, &rest, 1) == CUDA_SUCCESS &&
         drv.GreenCtxCreate(&c->green_ser, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS && drv.GreenCtxCreate(&c->green_par, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS;
    if (!ok) { g_create_error = "CCV2_GREEN: the device could not be split into two SM partitions (green contexts)"; ccv2_destroy(c); return CCV2_ERR_CUDA; }
    c->green_sms = (int)grp.sm.smCount;
  }
  for (int i = 0; i < MAX_STREAMS; i++) {
    if (c->green_sms > 0) {
      CUstream a = nullptr, b = nullptr;
      if (c->drv_green_stream_create(&a, c->green_par, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS || c->drv_green_stream_create(&b, c->green_ser, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) { g_create_error = "cuGreenCtxStreamCreate failed"; ccv2_destroy(c); return CCV2_ERR_CUDA; }
      c->streams[i] = (cudaStream_t)a; c->ser_streams[i] = (cudaStream_t)b;
      if (i < SIDE_STREAMS) { CUstream s2 = nullptr; if (c->drv_green_stream_create(&s2, c->green_ser, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) { g_create_error = "cuGreenCtxStreamCreate failed"; ccv2_destroy(c); return CCV2_ERR_CUDA; } c->side_streams[i] = (cudaStream_t)s2; }
      continue;
    }
    if ((e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    c->ser_streams[i] = c->streams[i];
    if (i < SIDE_STREAMS && (e = cudaStreamCreateWithFlags(&c->side_streams[i], cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  }
  for (auto &x : c->calls) {
    if ((e = cudaEventCreate(&x.ev_start)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&x.ev_end)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&x.ev_setup, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  }
  if ((e = cudaEventCreateWithFlags(&c->ev_id_chain, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&c->ev_t0)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&c->ev_t1)) != cudaSuccess) return fail(e, "cudaEventCreate");
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
  for (auto &x : c->calls) {
    for (auto *v : { &x.ev_h2d, &x.ev_side, &x.ev_done, &x.ev_fin, &x.ev_enc, &x.ev_hop, &x.ev_trace }) for (auto ev : *v) cudaEventDestroy(ev);
    for (cudaEvent_t ev : { x.ev_start, x.ev_end, x.ev_setup }) if (ev) cudaEventDestroy(ev);
    x.enc_frames.release(); x.dec_frames.release(); x.stage.release(); x.h_frames.release(); x.h_dframes.release();
  }
  for (Ring *r : { &c->fe, &c->ll }) { for (auto ev : r->ev_free) cudaEventDestroy(ev); r->buf.release(); }
  for (auto ev : c->prof_pool) cudaEventDestroy(ev);
  for (cudaEvent_t ev : { c->ev_id_chain, c->ev_t0, c->ev_t1 }) if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < MAX_STREAMS; i++) { if (c->ser_streams[i] && c->ser_streams[i] != c->streams[i]) cudaStreamDestroy(c->ser_streams[i]); if (c->streams[i]) cudaStreamDestroy(c->streams[i]); }
  for (int i = 0; i < SIDE_STREAMS; i++) if (c->side_streams[i]) cudaStreamDestroy(c->side_streams[i]);
  for (cudaStream_t s : { c->main_stream, c->copy_stream, c->d2h_stream, c->fin_stream, c->end_stream }) if (s) cudaStreamDestroy(s);
  if (c->green_ser && c->drv_green_destroy) c->drv_green_destroy(c->green_ser);
  if (c->green_par && c->drv_green_destroy) c->drv_green_destroy(c->green_par);
  if (c->d_tables) cudaFree(c->d_tables);
  if (c->d_frame_counter) cudaFree(c->d_frame_counter);
  c->out_cloud.release(); c->tile_pts.release(); c->tile_aux.release(); c->inter_ws.release(); c->inter_batch.release();
  if (c->intra_child) ccv2_destroy(c->intra_child);
  for (cudaEvent_t ev : { c->inter_ev0, c->inter_ev1 }) if (ev) cudaEventDestroy(ev);
  delete c;
}

size_t ccv2_max_compressed_size(size_t npts) { return stream_cap_for(npts, true, true, true); }

void *ccv2_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void ccv2_host_free(void *p) { if (p) cudaFreeHost(p); }

int ccv2_get_metrics(ccv2_codec *c, uint64_t m[3]) { if (!c || !m) return CCV2_ERR_ARG; for (int i = 0; i < 3; i++) m[i] = c->metrics[i]; return CCV2_OK; }
int ccv2_set_frame_id(ccv2_codec *c, uint32_t v) {
  if (!c) return CCV2_ERR_ARG;
  cudaSetDevice(c->device);
  finish_all(c); drain(c);
  c->frame_id = v;
  CU(cudaMemcpy(c->d_frame_counter, &v, 4, cudaMemcpyHostToDevice));
  return CCV2_OK;
}
uint32_t ccv2_get_frame_id(const ccv2_codec *c) { return c ? c->frame_id : 0; }
uint64_t ccv2_last_launch_count(const ccv2_codec *c) { return c ? c->launches : 0; }
float ccv2_last_device_ms(const ccv2_codec *c) { return c ? c->device_ms : 0.f; }

int ccv2_set_profiling(ccv2_codec *c, int on) { if (!c) return CCV2_ERR_ARG; finish_all(c); c->profiling = on != 0; c->prof_sum.clear(); return CCV2_OK; }
int ccv2_get_profile(const ccv2_codec *c, int idx, const char **name, float *total_ms, int *launches) {
  if (!c || idx < 0 || idx >= (int)c->prof_sum.size()) return CCV2_ERR_ARG;
  if (name) *name = c->prof_sum[idx].name.c_str();
  if (total_ms) *total_ms = c->prof_sum[idx].ms;
  if (launches) *launches = c->prof_sum[idx].launches;
  return CCV2_OK;
}

int ccv2_peek_point_count(const void *in_host, size_t len, uint64_t *npts) {
  if (!in_host || !npts || len < FRAME_HDR_BYTES) return CCV2_ERR_ARG;
  PeekInfo pi;
  if (!peek_header((const uint8_t *)in_host, len, pi)) return CCV2_ERR_STREAM;
  if (pi.point_count >= (1ull << 28)) return CCV2_ERR_STREAM;      // the codec's frame limit: a forged count must not size a caller's allocation
  *npts = pi.point_count;
  return CCV2_OK;
}

}  // extern "C"

// ================================================================================================ rings
// (Re)shapes a ring for sets of G frames of frame_bytes each.  An existing ring is kept when it already fits (calls of
// the same shape share it while in flight); otherwise everything in flight is collected first.
static int ensure_ring(ccv2_codec *c, Ring &r, size_t frame_bytes, int G, int want_sets, const char *what) {
  const bool fits = r.G == G && r.nsets > 0 && frame_bytes <= r.frame_bytes;
  if (fits && (r.nsets >= want_sets || r.mem_limited)) return CCV2_OK;
  finish_all(c);
  drain(c);
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t reserve = ((size_t)1 << 30) + total_b / 64;
  const size_t avail = free_b + r.buf.cap > reserve ? free_b + r.buf.cap - reserve : 0;
  const size_t set_bytes = (size_t)G * frame_bytes;
  int nsets = (int)std::min<size_t>((size_t)want_sets, avail / set_bytes);
  if (fits && nsets <= r.nsets) { r.mem_limited = true; return CCV2_OK; }                  // memory would not give more than there is
  if (nsets < 1) { c->err = std::string("not enough device memory for one ") + what + " workspace set"; return CCV2_ERR_CUDA; }
  cudaError_t e = r.buf.ensure((size_t)nsets * set_bytes, false);
  while (e != cudaSuccess && nsets > 1) { cudaGetLastError(); nsets = (nsets + 1) / 2; e = r.buf.ensure((size_t)nsets * set_bytes, false); }
  if (e != cudaSuccess) { cudaGetLastError(); c->err = std::string("cudaMalloc (") + what + " ring): " + cudaGetErrorString(e); return CCV2_ERR_CUDA; }
  r.frame_bytes = frame_bytes; r.G = G; r.nsets = nsets; r.seq = 0; r.mem_limited = nsets < want_sets;
  while ((int)r.ev_free.size() < nsets) { cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); r.ev_free.push_back(ev); }
  r.used.assign(nsets, 0);
  c->enc_host_valid.assign(c->enc_host_valid.size(), 0);     // intermediates of the last encode are gone
  return CCV2_OK;
}

static void grow_events(std::vector<cudaEvent_t> &v, size_t n, bool timing = false) {
  while (v.size() < n) { cudaEvent_t ev; if (cudaEventCreateWithFlags(&ev, timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) break; v.push_back(ev); }
}


// ================================================================================================ submit
// mode 0: encode, 1: decode, 2: encode -> decode round trip (the decoder reads the encoder's device-resident stream).
static int submit_call(ccv2_codec *c, int mode, int nframes,
                       const void *const *pts, const size_t *npts, void *const *out, const size_t *out_cap, size_t *out_len,
                       const void *const *in, const size_t *in_len, void *const *pts_out, const size_t *pts_cap, size_t *npts_out,
                       bool boost, uint32_t fixed_id, int *ticket_out) {
  const bool do_enc = mode != 1, do_dec = mode != 0, rt = mode == 2;
  c->err.clear();
  CU(cudaSetDevice(c->device));
  const ccv2_params &prm = c->prm;
  const bool cen = prm.do_voxel_grid_centroid != 0, color = prm.do_color_encoding != 0, lines = prm.color_coding_type == 2;
  const bool detail = prm.do_voxel_grid_downsampling == 0;           // the encoder's mode; a decoder learns it from the stream
  CallCtx &x = boost ? c->calls[N_CALLS - 1] : c->calls[c->user_calls++ % (N_CALLS - 1)];
  if (x.busy) finish_call(c, x);
  const int ticket = c->next_ticket++;
  if (ticket_out) *ticket_out = ticket;
  x.ticket = ticket; x.mode = mode; x.nframes = nframes; x.boost = boost; x.rc_early = CCV2_OK; x.launches = 0; x.device_ms = 0;
  x.pts = pts; x.npts = npts; x.out = out; x.out_cap = out_cap; x.out_len = out_len; x.in = in; x.in_len = in_len; x.pts_out = pts_out; x.pts_cap = pts_cap; x.npts_out = npts_out;
  x.trace_marks.clear();
  if (nframes == 0) { x.busy = false; x.ngroups = 0; return CCV2_OK; }

  // ------------------------------------------------------------------ classify the caller's buffers, size the workspaces
  x.in_kind.assign(nframes, PK_NONE); x.out_kind.assign(nframes, PK_NONE); x.pts_kind.assign(nframes, PK_NONE);
  x.stage_off_stream.assign(nframes + 1, 0); x.stage_off_pts.assign(nframes + 1, 0);
  size_t nmax = 1, ncap_max = 1, tcap_max = 0, ccap_max = 0, in_stage_max = 0, out_stage_max = 0, stage_total = 0, dpts_max = 0;
  bool host_in_any = false;
  bool dec_cen = rt ? (cen && !detail) : false;            // does any frame carry a centroid layer?  (round trip: the encoder's setting; decode: the header, or assume so)
  std::vector<size_t> dcount(nframes, 0);                  // records a pinned destination receives by one copy-engine transfer
  for (int i = 0; i < nframes; i++) {
    if (do_enc) {
      if (npts[i] >= (1u << 28)) { c->err = "frame too large"; return CCV2_ERR_ARG; }
      if (npts[i] && !pts[i]) return CCV2_ERR_ARG;
      nmax = std::max(nmax, npts[i]);
      x.in_kind[i] = npts[i] ? ptr_kind(pts[i]) : PK_DEVICE;
      if (x.in_kind[i] != PK_DEVICE) host_in_any = true;
      x.out_kind[i] = (out && out[i]) ? ptr_kind(out[i]) : PK_NONE;
      if (x.out_kind[i] == PK_PAGEABLE) { x.stage_off_stream[i] = stage_total; stage_total += (std::min(out_cap[i], stream_cap_for(npts[i], cen, boost, detail)) + 255) & ~size_t(255); }
    }
    if (do_dec) {
      if (pts_cap[i] >= (1u << 28)) { c->err = "frame too large"; return CCV2_ERR_ARG; }
      if ((!rt && in_len[i] && !in[i]) || (pts_cap[i] && !pts_out[i])) return CCV2_ERR_ARG;
      x.pts_kind[i] = pts_cap[i] ? ptr_kind(pts_out[i]) : PK_DEVICE;
      size_t ncap = pts_cap[i], tcap, ccap;
      if (rt) { ncap = std::max<size_t>(npts[i], 1); tcap = tree_cap_for(ncap, boost); ccap = cpay_cap_for(ncap, boost); dcount[i] = std::min(pts_cap[i], npts[i]); if (detail) dpts_max = std::max(dpts_max, ncap); }
      else {
        x.in_kind[i] = in_len[i] ? ptr_kind(in[i]) : PK_DEVICE;
        PeekInfo pi; bool peeked = false;
        if (x.in_kind[i] != PK_DEVICE && in_len[i]) {       // host stream: read the header here and size the workspace from it
          peeked = peek_header((const uint8_t *)in[i], in_len[i], pi);
          in_stage_max = std::max(in_stage_max, (in_len[i] + 64 + 255) & ~size_t(255));
        }
        if (peeked && pi.point_count < (1ull << 28)) {
          ncap = std::max<size_t>(1, std::min<size_t>(pts_cap[i], pi.point_count));
          tcap = (size_t)std::min<uint64_t>(pi.B, 22ull * ncap + 1024) + 1024;          // B <= depth * V: a larger size word is a malformed stream
          ccap = cpay_cap_for(ncap, boost) + (boost ? 2 * in_len[i] : 0);
          dcount[i] = ncap;
          if (!pi.voxel_grid) dpts_max = std::max(dpts_max, ncap);
          if (pi.centroid) dec_cen = true;               // detail-mode stream: room for the enhancement vectors
        } else { ncap = std::max<size_t>(ncap, 1); tcap = tree_cap_for(ncap, boost); ccap = cpay_cap_for(ncap, boost) + (boost ? 2 * in_len[i] : 0); dcount[i] = pts_cap[i]; dec_cen = true; if (boost) dpts_max = std::max(dpts_max, ncap); }   // header not read here: detail buffers only on the retry
      }
      ncap_max = std::max(ncap_max, ncap); tcap_max = std::max(tcap_max, tcap); ccap_max = std::max(ccap_max, ccap);
      if (x.pts_kind[i] == PK_PINNED) out_stage_max = std::max(out_stage_max, (32 * dcount[i] + 255) & ~size_t(255));
      if (x.pts_kind[i] == PK_PAGEABLE) { x.stage_off_pts[i] = stage_total; stage_total += (32 * pts_cap[i] + 255) & ~size_t(255); }
    }
  }
  const int NS = c->profiling ? 1 : (c->n_streams ? c->n_streams : MAX_STREAMS);
  const int G = c->profiling ? std::max(1, std::min(nframes, 64)) : (c->group ? c->group : std::max(1, std::min(MAX_GROUP, (nframes + 7) / 8)));
  const int ngroups = (nframes + G - 1) / G;
  x.G = G; x.ngroups = ngroups;

  size_t fe_bytes = 0, fe_zero = 0, ll_bytes = 0, stream_end = 0, enc_ll = 0, dec_ws = 0, dec_zero = 0;
  if (do_enc) { size_t zo; fe_bytes = carve_fe(nullptr, nmax, nullptr, prm, boost, host_in_any, &zo, &fe_zero); enc_ll = carve_enc_ll(nullptr, nmax, nullptr, cen, boost, &stream_end, detail); }
  if (do_dec) dec_ws = carve_dec(nullptr, ncap_max, tcap_max, ccap_max, nullptr, &dec_zero, lines, dpts_max, boost, dec_cen);
  const size_t ws_off = rt ? stream_end : 0;                          // round trip: the decoder's workspace lies over the encoder's dead buffers
  const size_t in_stage_off = std::max(enc_ll, ws_off + dec_ws);
  const size_t out_stage_off = in_stage_off + in_stage_max;
  ll_bytes = out_stage_off + out_stage_max;
  {
    // at most one long-lived set per work stream: sets that share a stream queue behind each other although they are unrelated
    // (measured end to end with 19 sets on 16 streams: a group's kernels 220 ms late, the copy engine idle meanwhile)
    const int want_ll = std::max(1, std::min(std::min((N_CALLS - 1) * ngroups, NS), (c->inflight_max + G - 1) / G));
    // host inputs are uploaded into the front-end set: with only two sets the copy engine waits whenever a front-end runs long
    // under load (measured: 140 ms instead of 20, the next upload 90 ms late), so staged inputs get four
    const int fe_frames = c->fe_frames ? c->fe_frames : (host_in_any ? 512 : 256);
    const int want_fe = std::max(1, std::min((N_CALLS - 1) * ngroups, std::max(2, fe_frames / G)));
    int rc;
    if (do_enc && (rc = ensure_ring(c, c->fe, fe_bytes, G, want_fe, "front-end")) != CCV2_OK) return rc;
    if ((rc = ensure_ring(c, c->ll, ll_bytes, G, want_ll, "long-lived")) != CCV2_OK) return rc;
  }
  Ring &fe = c->fe, &ll = c->ll;
  x.fe_seq0 = fe.seq; x.ll_seq0 = ll.seq; x.fe_nsets = fe.nsets; x.ll_nsets = ll.nsets;
  x.fe_set.assign(ngroups, 0); x.ll_set.assign(ngroups, 0);
  for (int g = 0; g < ngroups; g++) { x.fe_set[g] = do_enc ? (int)((fe.seq + g) % fe.nsets) : 0; x.ll_set[g] = (int)((ll.seq + g) % ll.nsets); }
  if (do_enc) fe.seq += ngroups;
  ll.seq += ngroups;
  CU(x.stage.ensure(stage_total + 256));
  grow_events(x.ev_h2d, ngroups); grow_events(x.ev_side, 2 * (size_t)ngroups); grow_events(x.ev_done, ngroups); grow_events(x.ev_fin, ngroups); grow_events(x.ev_enc, ngroups); grow_events(x.ev_hop, 4 * (size_t)ngroups);
  if ((int)x.ev_h2d.size() < ngroups || (int)x.ev_side.size() < 2 * ngroups || (int)x.ev_done.size() < ngroups || (int)x.ev_fin.size() < ngroups || (int)x.ev_enc.size() < ngroups || (int)x.ev_hop.size() < 4 * ngroups) { c->err = "cudaEventCreate failed"; return CCV2_ERR_CUDA; }

  // ------------------------------------------------------------------ frame records
  EncFrame *hf = nullptr, *df = nullptr; DecFrame *hd = nullptr, *dd = nullptr;
  EncParams P; HeaderParams H;
  memset(&P, 0, sizeof P); memset(&H, 0, sizeof H);
  size_t frames_bytes = 0;
  if (do_enc) {
    frames_bytes = (sizeof(EncFrame) * nframes + 255) & ~size_t(255);
    CU(x.enc_frames.ensure(frames_bytes + (size_t)nframes * 5 * 256 * 4));
    CU(x.h_frames.ensure(sizeof(EncFrame) * nframes));
    hf = (EncFrame *)x.h_frames.p; df = (EncFrame *)x.enc_frames.p;
    memset(hf, 0, sizeof(EncFrame) * nframes);
    for (int i = 0; i < nframes; i++) {
      EncFrame &f = hf[i];
      const int g = i / G, j = i % G;
      f.n = (uint32_t)npts[i];
      f.n_finite = (uint32_t)npts[i];                               // keygen subtracts the non-finite points
      f.violator = NONE_U32;
      f.frame_id_fixed = fixed_id;
      f.pts = (const uint8_t *)pts[i];
      uint8_t *fb = fe.frame_base(x.fe_set[g], j);
      carve_fe(fb, nmax, &f, prm, boost, x.in_kind[i] != PK_DEVICE, nullptr, nullptr);     // host input: f.pts = the set's staging area
      f.zero_ptr = fb; f.zero_bytes = fe_zero;
      carve_enc_ll(ll.frame_base(x.ll_set[g], j), nmax, &f, cen, boost, nullptr, detail);
      f.hist = (uint32_t *)((uint8_t *)x.enc_frames.p + frames_bytes) + (size_t)i * 5 * 256;
      if (color && (prm.color_coding_type == 0 || prm.color_coding_type == 3)) f.avg = f.cpay;   // raw averages are the colour payload
      switch (x.out_kind[i]) {
        case PK_DEVICE: f.out_ptr = (uint8_t *)out[i]; f.out_cap = out_cap[i]; break;
        case PK_PINNED: { void *dp = nullptr; CU(cudaHostGetDevicePointer(&dp, out[i], 0)); f.out_ptr = (uint8_t *)dp; f.out_cap = out_cap[i]; break; }
        case PK_PAGEABLE: f.out_ptr = (uint8_t *)x.stage.p + x.stage_off_stream[i]; f.out_cap = std::min(out_cap[i], stream_cap_for(npts[i], cen, boost, detail)); break;
        default: f.out_ptr = nullptr; f.out_cap = 0; break;
      }
    }
    P.res = prm.octree_resolution;
    { int ex; double m = frexp(P.res, &ex); P.res_pow2 = (m == 0.5); P.inv_res = P.res_pow2 ? 1.0 / P.res : 0.0; }
    P.do_color = color; P.color_type = prm.color_coding_type; P.do_centroid = cen;
    P.color_reduction = (prm.color_coding_type == 0) ? std::max(0, 8 - (int)prm.color_bit_resolution) : 0;   // jp_color_coder_ is never configured
    P.prefix_len = 16384;
    P.detail = detail; P.point_res_f = (float)prm.point_resolution;
    P.allow_packed = (!cen && !detail && c->allow_packed) ? 1 : 0;
    H.octree_res = prm.octree_resolution; H.point_res = (double)(float)prm.point_resolution;
    H.do_voxel_grid = detail ? 0 : 1; H.with_color = color; H.color_bits = prm.color_bit_resolution; H.do_centroid = cen;
    H.connectivity = prm.code_connectivity != 0; H.scalable = prm.create_scalable_stream != 0; H.icp_offset = prm.do_icp_color_offset != 0; H._p = 0;
    H.color_type = prm.color_coding_type; H.macroblock = prm.macroblock_size;
    c->last_enc_params = P;
  }
  if (do_dec) {
    CU(x.dec_frames.ensure(sizeof(DecFrame) * nframes));
    CU(x.h_dframes.ensure(sizeof(DecFrame) * nframes));
    hd = (DecFrame *)x.h_dframes.p; dd = (DecFrame *)x.dec_frames.p;
    memset(hd, 0, sizeof(DecFrame) * nframes);
    for (int i = 0; i < nframes; i++) {
      DecFrame &f = hd[i];
      const int g = i / G, j = i % G;
      uint8_t *lb = ll.frame_base(x.ll_set[g], j);
      if (rt) { f.in = hf[i].stream; f.in_len = 0; }       // length filled in on the device by link_kernel
      else {
        f.in = x.in_kind[i] == PK_DEVICE ? (const uint8_t *)in[i] : lb + in_stage_off;
        f.in_len = in_len[i];
        if (in_len[i] == 0) f.error = FERR_BAD_STREAM;
      }
      switch (x.pts_kind[i]) {
        case PK_PINNED: f.out_pts = lb + out_stage_off; break;
        case PK_PAGEABLE: f.out_pts = (uint8_t *)x.stage.p + x.stage_off_pts[i]; break;
        default: f.out_pts = (uint8_t *)pts_out[i]; break;
      }
      f.out_cap = pts_cap[i];
      carve_dec(lb + ws_off, ncap_max, tcap_max, ccap_max, &f, nullptr, lines, dpts_max, boost, dec_cen);
      f.zero_ptr = lb + ws_off; f.zero_bytes = dec_zero;
    }
  }

  // ------------------------------------------------------------------ enqueue
  cudaStream_t ms = c->main_stream;
  auto mark = [&](int g, const char *label, cudaStream_t s) {       // CCV2_TRACE: timestamp on the group's stream
    if (!c->trace) return;
    if (x.ev_trace.size() <= x.trace_marks.size()) grow_events(x.ev_trace, x.trace_marks.size() + 1, true);
    if (x.ev_trace.size() <= x.trace_marks.size()) return;
    cudaEventRecord(x.ev_trace[x.trace_marks.size()], s);
    x.trace_marks.push_back({g, label});
  };
  x.busy = true;
  // from here on work is in flight: an error must not leave it queued behind the caller's back
#define CUQ(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e__); drain(c); cudaGetLastError(); x.busy = false; return CCV2_ERR_CUDA; } } while (0)
  CUQ(cudaEventRecord(x.ev_start, ms));
  if (do_enc) {
    CUQ(cudaMemcpyAsync(df, hf, sizeof(EncFrame) * nframes, cudaMemcpyHostToDevice, ms));
    CUQ(cudaMemsetAsync((uint8_t *)x.enc_frames.p + frames_bytes, 0, (size_t)nframes * 5 * 256 * 4, ms));
  }
  if (do_dec) CUQ(cudaMemcpyAsync(dd, hd, sizeof(DecFrame) * nframes, cudaMemcpyHostToDevice, ms));
  // serial CTAs reserve enough shared memory (1 KB per CTA is the system's) that only serial_cap of them fit on an SM
  const int serial_cap = c->serial_cap ? c->serial_cap : (std::min(nframes, ll.nsets * G) + c->n_sm - 1) / c->n_sm;
  size_t serial_smem_enc = 0, serial_smem_dec = 0;
  if (serial_cap > 0 && serial_cap <= 14 && !c->profiling) {
    const size_t per_cta = (c->smem_sm / 1024 / (size_t)(serial_cap + 1) + 1) * 1024;
    if (per_cta > c->smem_static_enc + 1024) serial_smem_enc = per_cta - c->smem_static_enc - 1024;
    if (per_cta > c->smem_static_dec + 1024) serial_smem_dec = per_cta - c->smem_static_dec - 1024;
  }
  CUQ(cudaEventRecord(x.ev_setup, ms));
  bool host_io = host_in_any;
  for (int i = 0; i < nframes && !host_io; i++) host_io = (x.out_kind[i] != PK_NONE && x.out_kind[i] != PK_DEVICE) || (do_dec && x.pts_kind[i] != PK_DEVICE);
  // the lane-per-stream decoder wins whenever many frames are in flight on the device (2416 against 1902 Mpoints/s round trip,
  // 4329 against 3239 decode-only); with host buffers the copies pace the pipeline and the CTA-per-frame decoder's shorter
  // latency is worth more (1174 against 1103 end to end)
  // ... and a call of a few frames is a latency matter: the CTA-per-frame decoder (179 ms for a 1M-point frame against 240)
  const bool use_lps_dec = c->lps_dec < 0 ? (!host_io && nframes >= 256) : c->lps_dec != 0;
  uint64_t launches = 0;
  uint32_t *counter = c->d_frame_counter;
  bool copy_waits_setup = false;
  for (int g = 0; g < ngroups; g++) {
    const int sl = x.ll_set[g], sf = x.fe_set[g];
    cudaStream_t const sp = c->streams[sl % NS], ss = c->profiling ? sp : c->ser_streams[sl % NS];   // parallel kernels / serial range-coder kernels (same stream unless the SMs are partitioned)
    cudaStream_t st = sp;
    int hops = 0;
    auto hop = [&](cudaStream_t from, cudaStream_t to) -> cudaError_t {     // hand the group over to the other partition's stream
      if (from == to) return cudaSuccess;
      cudaEvent_t ev = x.ev_hop[4 * g + (hops++ & 3)];
      cudaError_t e2 = cudaEventRecord(ev, from);
      return e2 != cudaSuccess ? e2 : cudaStreamWaitEvent(to, ev, 0);
    };
    const int f0 = g * G, gf = std::min(G, nframes - f0);
    const unsigned steered_grid = (unsigned)((gf + c->n_sm - 1) / c->n_sm * c->n_sm);
    CUQ(cudaStreamWaitEvent(st, x.ev_setup, 0));
    bool any_h2d = false;
    for (int i = 0; i < gf; i++) any_h2d |= do_enc ? (x.in_kind[f0 + i] != PK_DEVICE && npts[f0 + i]) : (x.in_kind[f0 + i] != PK_DEVICE && in_len[f0 + i]);
    // take the sets: the previous users' streams (and copies) must be done with them
    if (do_enc && fe.used[sf]) { CUQ(cudaStreamWaitEvent(st, fe.ev_free[sf], 0)); if (any_h2d) CUQ(cudaStreamWaitEvent(c->copy_stream, fe.ev_free[sf], 0)); }
    if (ll.used[sl]) { CUQ(cudaStreamWaitEvent(st, ll.ev_free[sl], 0)); if (any_h2d && !do_enc) CUQ(cudaStreamWaitEvent(c->copy_stream, ll.ev_free[sl], 0)); }
    if (any_h2d && !copy_waits_setup) { CUQ(cudaStreamWaitEvent(c->copy_stream, x.ev_setup, 0)); copy_waits_setup = true; }
    if (do_enc) {
      EncFrame *dg = df + f0;
      size_t gn = 1;
      // Host inputs go through ONE copy stream in group order: copies issued on the group streams would be
      // interleaved by the copy engine and every group's input would land at the very end (measured), which defeats
      // the pipeline.  This way group g can start as soon as its own clouds are on the device.
      for (int i = 0; i < gf; i++) {
        gn = std::max(gn, npts[f0 + i]);
        if (x.in_kind[f0 + i] != PK_DEVICE && npts[f0 + i])
          CUQ(cudaMemcpyAsync((void *)hf[f0 + i].pts, pts[f0 + i], 32 * npts[f0 + i], cudaMemcpyHostToDevice, c->copy_stream));
      }
      LAUNCH("zero_region_kernel", zero_region_kernel<EncFrame><<<dim3(128, gf), 256, 0, st>>>(dg));
      if (any_h2d) { CUQ(cudaEventRecord(x.ev_h2d[g], c->copy_stream)); CUQ(cudaStreamWaitEvent(st, x.ev_h2d[g], 0)); }
      mark(g, "inputs", st);
      const unsigned gx256 = (unsigned)((gn + 255) / 256), gtiles = (unsigned)((gn + SORT_TILE - 1) / SORT_TILE);
      LAUNCH("bbox_kernel", bbox_kernel<<<gf, 1024, 0, st>>>(dg, P, 0));
      LAUNCH("bbox_fixup_kernel", bbox_fixup_kernel<<<gf, 32, 0, st>>>(dg, P));
      LAUNCH("keygen_kernel", keygen_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P, 0));
      LAUNCH("bbox_kernel(slow path)", bbox_kernel<<<gf, 1024, 0, st>>>(dg, P, 1));
      LAUNCH("keygen_kernel(rekey)", keygen_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P, 1));
      // frame ids are sequential over the batch (and over calls): a group's setup waits for the previous group's
      if (c->id_chain_used) CUQ(cudaStreamWaitEvent(st, c->ev_id_chain, 0));
      LAUNCH("frame_setup_kernel", frame_setup_kernel<<<1, 32, 0, st>>>(dg, gf, counter));
      CUQ(cudaEventRecord(c->ev_id_chain, st)); c->id_chain_used = true;
      LAUNCH("sort_hist_kernel", sort_hist_kernel<<<dim3(gtiles, gf), 256, 0, st>>>(dg));
      for (int p = 0; p < 8; p++) LAUNCH("sort_pass_kernel", sort_pass_kernel<<<dim3(gtiles, gf), SORT_THREADS, SORT_SMEM_BYTES, st>>>(dg, p));
      LAUNCH("leaf_scan_kernel", leaf_scan_kernel<<<dim3((unsigned)((gn + LEAF_TILE - 1) / LEAF_TILE), gf), LEAF_THREADS, 0, st>>>(dg, color && prm.color_coding_type == 1));
      LAUNCH("leaf_emit_kernel", leaf_emit_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P));
      if (color && prm.color_coding_type == 1) {
        const size_t img_h = gn / 256 + 1, mcu_h = (img_h + 15) / 16, nblk = mcu_h * 16 * 6;
        LAUNCH("jpeg_mcu_kernel", jpeg_mcu_kernel<<<dim3((unsigned)(mcu_h * 16), gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("jpeg_huff_kernel", jpeg_huff_kernel<<<dim3((unsigned)((nblk + HUFF_THREADS - 1) / HUFF_THREADS), gf), HUFF_THREADS, 0, st>>>(dg, c->d_tables));
        const size_t jb = cpay_cap_for(gn, boost);
        LAUNCH("jpeg_stuff_kernel", jpeg_stuff_kernel<<<dim3((unsigned)((jb + STUFF_THREADS * STUFF_BYTES - 1) / (STUFF_THREADS * STUFF_BYTES)), gf), STUFF_THREADS, 0, st>>>(dg, c->d_tables));
      }
      if (color && prm.color_coding_type == 2) {
        const unsigned lines_max = (unsigned)(gn / LINE_PX + 1), mcus_max = (unsigned)(gn / 16 + 256);
        LAUNCH("lines_mcu_kernel", lines_mcu_kernel<<<dim3(mcus_max, gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("lines_huff_kernel", lines_huff_kernel<<<dim3(lines_max, gf), 256, 0, st>>>(dg, c->d_tables));
        LAUNCH("lines_offsets_kernel", lines_offsets_kernel<<<gf, 1024, 0, st>>>(dg));
        LAUNCH("lines_copy_kernel", lines_copy_kernel<<<dim3(lines_max, gf), 256, 0, st>>>(dg));
      }
      if (detail) {
        LAUNCH("detail_scan_kernel", detail_scan_kernel<<<dim3((unsigned)((gn + 1023) / 1024), gf), 256, 0, st>>>(dg));
        LAUNCH("detail_emit_kernel", detail_emit_kernel<<<dim3(gx256, gf), 256, 0, st>>>(dg, P));
      }
      // the front-end set goes to the next group: everything from here on lives in the long-lived set
      CUQ(cudaEventRecord(fe.ev_free[sf], st)); fe.used[sf] = 1;
      const size_t hmax = std::max(std::max(tree_cap_for(gn, boost), cpay_cap_for(gn, boost)), detail ? diff_cap_for(gn) : 0);
      mark(g, "leaves", st);
      LAUNCH("hist_kernel", hist_kernel<<<dim3((unsigned)((hmax + 16383) / 16384), detail ? 5 : 3, gf), 256, 0, st>>>(dg));
      CUQ(hop(sp, ss)); st = ss;
      if (detail) LAUNCH("rc_encode_int_kernel", rc_encode_int_kernel<<<gf, 32, 0, st>>>(dg));
      if (c->lps_enc) LAUNCH("rc_encode_lps_kernel", rc_encode_lps_kernel<<<dim3((unsigned)((gf + 31) / 32), detail ? 5 : 3), 32, c->enc_reserve ? c->lps_smem_enc : 0, st>>>(dg, gf, cen, color, detail));
      else LAUNCH("rc_encode_kernel", rc_encode_kernel<<<gf, detail ? 160 : 96, serial_smem_enc, st>>>(dg, cen, color, detail));
      CUQ(hop(ss, sp)); st = sp;
      LAUNCH("assemble_kernel", assemble_kernel<<<dim3(64, gf), 256, 0, st>>>(dg, H));
      if (out) LAUNCH("export_kernel", export_kernel<<<dim3(2, gf), 256, 0, st>>>(dg));
      mark(g, "encoded", st);
    }
    if (do_dec) {
      DecFrame *dg = dd + f0;
      size_t pmax = 1;
      for (int i = 0; i < gf; i++) {
        const int k = f0 + i;
        pmax = std::max(pmax, rt ? std::max<size_t>(npts[k], 1) : std::min(ncap_max, std::max<size_t>(pts_cap[k], 1)));
        if (!rt && x.in_kind[k] != PK_DEVICE && in_len[k])
          CUQ(cudaMemcpyAsync((void *)hd[k].in, in[k], in_len[k], cudaMemcpyHostToDevice, c->copy_stream));
      }
      LAUNCH("zero_region_kernel", zero_region_kernel<DecFrame><<<dim3(64, gf), 256, 0, st>>>(dg));
      if (!rt && any_h2d) { CUQ(cudaEventRecord(x.ev_h2d[g], c->copy_stream)); CUQ(cudaStreamWaitEvent(st, x.ev_h2d[g], 0)); }
      if (rt) LAUNCH("link_kernel", link_kernel<<<(gf + 63) / 64, 64, 0, st>>>(df + f0, dg, gf));
      // Two entropy stages.  dec_entropy_kernel (a CTA per frame) has the shorter latency; the lane-per-stream decoder
      // (8 frames to a warp) executes a third of the instructions and wins when the SMs' issue slots are the limit.
      CUQ(hop(sp, ss)); st = ss;
      if (use_lps_dec) {
        // tree layers on the group's stream, speculated colour layers on a side stream at the same time
        cudaStream_t s2 = c->profiling ? st : c->side_streams[sl % c->n_side];
        const unsigned lps_ctas = (unsigned)((gf + LPS_DEC_FRAMES - 1) / LPS_DEC_FRAMES);
        LAUNCH("dec_head_kernel", dec_head_kernel<<<gf, 32, 0, st>>>(dg));
        if (s2 != st) { CUQ(cudaEventRecord(x.ev_side[2 * g], st)); CUQ(cudaStreamWaitEvent(s2, x.ev_side[2 * g], 0)); }
        LAUNCH("rc_decode_lps_kernel<tree>", rc_decode_lps_kernel<true><<<lps_ctas, 32 * (1 + LPS_DEC_FRAMES), sizeof(LpsSmem), st>>>(dg, gf, c->use_ring));
        LAUNCH_S(s2, "rc_decode_lps_kernel<colour>", rc_decode_lps_kernel<false><<<lps_ctas, 32, offsetof(LpsSmem, rg), s2>>>(dg, gf, 0));
        mark(g, "tree", st);
        mark(g, "colour-rc", s2);
        LAUNCH_S(s2, "dec_jpeg_kernel", dec_jpeg_kernel<<<gf, 32, 0, s2>>>(dg));
        mark(g, "jpeg", s2);
        if (s2 != st) { CUQ(cudaEventRecord(x.ev_side[2 * g + 1], s2)); CUQ(cudaStreamWaitEvent(st, x.ev_side[2 * g + 1], 0)); }
        LAUNCH("dec_finish_kernel", dec_finish_kernel<<<gf, 32, 0, st>>>(dg));
      } else
      LAUNCH("dec_entropy_kernel", dec_entropy_kernel<<<gf, 96, serial_smem_dec, st>>>(dg, c->use_ring));
      mark(g, "entropy", st);
      CUQ(hop(ss, sp)); st = sp;
      LAUNCH("jpeg_destuff_kernel", jpeg_destuff_kernel<<<gf, 1024, 0, st>>>(dg));
      LAUNCH("dec_serial_kernel", dec_serial_kernel<<<steered_grid, 64, 0, st>>>(dg, f0, gf));
      if (lines) {
        const unsigned lines_max = (unsigned)(pmax / LINE_PX + 2);
        LAUNCH("lines_index_kernel", lines_index_kernel<<<gf, 32, 0, st>>>(dg));
        LAUNCH("lines_decode_kernel", lines_decode_kernel<<<dim3(lines_max, gf), 32, 0, st>>>(dg));
        LAUNCH("lines_idct_kernel", lines_idct_kernel<<<dim3((unsigned)(((size_t)lines_max * LINE_MCU_STRIDE + LINE_MCU_STRIDE) * 6 / 32 + 1), gf), 256, 0, st>>>(dg, c->d_tables));
      }
      const size_t img_h = pmax / 256 + 2, mcu_h = (img_h + 15) / 16, nblocks = mcu_h * 16 * 6;
      LAUNCH("jpeg_idct_kernel", jpeg_idct_kernel<<<dim3((unsigned)((nblocks + 31) / 32), gf), 256, 0, st>>>(dg, c->d_tables));
      LAUNCH("dec_leaves_kernel", dec_leaves_kernel<<<dim3((unsigned)((pmax + 255) / 256), gf), 256, 0, st>>>(dg));        // frames walked by the pipelined walkers
      LAUNCH("dec_points_kernel", dec_points_kernel<<<dim3((unsigned)((pmax + NODE_THREADS - 1) / NODE_THREADS), gf), NODE_THREADS, 0, st>>>(dg));
      LAUNCH("detail_points_kernel", detail_points_kernel<<<dim3((unsigned)((pmax + 255) / 256), gf), 256, 0, st>>>(dg));   // detail-mode frames only
      mark(g, "decoded", st);
    }
    CUQ(cudaEventRecord(x.ev_done[g], st));
    // decoded clouds for pinned host destinations: one copy-engine transfer per frame out of the set's staging area, on
    // the D2H stream in group order; the set is free once they have left
    bool any_d2h = false;
    if (do_dec) for (int i = 0; i < gf; i++) any_d2h |= x.pts_kind[f0 + i] == PK_PINNED && dcount[f0 + i] > 0;
    if (any_d2h) {
      CUQ(cudaStreamWaitEvent(c->d2h_stream, x.ev_done[g], 0));
      for (int i = 0; i < gf; i++) {
        const int k = f0 + i;
        if (x.pts_kind[k] == PK_PINNED && dcount[k]) CUQ(cudaMemcpyAsync(pts_out[k], hd[k].out_pts, 32 * dcount[k], cudaMemcpyDeviceToHost, c->d2h_stream));
      }
      mark(g, "copied", c->d2h_stream);
      CUQ(cudaEventRecord(ll.ev_free[sl], c->d2h_stream));
      CUQ(cudaEventRecord(x.ev_fin[g], c->d2h_stream));
    } else {
      CUQ(cudaEventRecord(ll.ev_free[sl], st));
      CUQ(cudaEventRecord(x.ev_fin[g], st));
    }
    ll.used[sl] = 1;
    CUQ(cudaGetLastError());
  }
  // the call's end is gathered on the end stream: on the control stream it would hold back the set-up of the next call
  for (int g = 0; g < ngroups; g++) CUQ(cudaStreamWaitEvent(c->end_stream, x.ev_fin[g], 0));
  // The frame records come home by a KERNEL storing into the pinned host copy, not by a copy-engine transfer: the engine's
  // queue is in order, and a small D2H copy issued when the call is collected sits behind every decoded-cloud transfer of the
  // LATER calls that is already queued (each waiting for its group's event).  Measured with pinned host buffers: ccv2_wait of
  // call k returned when call k+2's last cloud had left, the next submit came 600 ms late and the upload engine idled a third
  // of the time (1100 Mpoints/s end to end).
  if (do_enc) { void *hp = nullptr; CUQ(cudaHostGetDevicePointer(&hp, hf, 0)); records_home_kernel<<<32, 256, 0, c->end_stream>>>((uint64_t *)hp, (const uint64_t *)df, sizeof(EncFrame) * (size_t)nframes / 8); launches++; }
  if (do_dec) { void *hp = nullptr; CUQ(cudaHostGetDevicePointer(&hp, hd, 0)); records_home_kernel<<<32, 256, 0, c->end_stream>>>((uint64_t *)hp, (const uint64_t *)dd, sizeof(DecFrame) * (size_t)nframes / 8); launches++; }
  CUQ(cudaGetLastError());
  CUQ(cudaEventRecord(x.ev_end, c->end_stream));
#undef CUQ
  x.launches = launches;
  return CCV2_OK;
}

// ================================================================================================ collect

static int retry_frame(ccv2_codec *c, CallCtx &x, int k, uint32_t fixed_id);

static int finish_call(ccv2_codec *c, CallCtx &x) {
  if (!x.busy) return CCV2_OK;
  x.busy = false;                                             // first: a retry below may come back here through ensure_ring
  const int mode = x.mode, nframes = x.nframes;
  const bool do_enc = mode != 1, do_dec = mode != 0, rt = mode == 2;
  int rc = CCV2_OK;
  auto store = [&](int r) { auto &v = c->done; v.push_back({x.ticket, r}); if (v.size() > 16) v.erase(v.begin()); return r; };
  cudaSetDevice(c->device);
  cudaError_t e = cudaEventSynchronize(x.ev_end);             // the frame records are in the pinned host copies by then (records_home_kernel)
  EncFrame *hf = (EncFrame *)x.h_frames.p; DecFrame *hd = (DecFrame *)x.h_dframes.p;
  if (e != cudaSuccess) { c->err = std::string("collect: ") + cudaGetErrorString(e); drain(c); cudaGetLastError(); return store(CCV2_ERR_CUDA); }
  cudaEventElapsedTime(&x.device_ms, x.ev_start, x.ev_end);
  c->device_ms = x.device_ms; c->launches = x.launches;
  const uint32_t retry_bits = FERR_TREE_CAP | FERR_STREAM_CAP | FERR_JPEG_CAP;
  std::vector<int> retry;
  for (int k = 0; k < nframes; k++) {
    bool enc_ok = true;
    uint64_t slen = 0;
    if (do_enc) {
      const EncFrame &f = hf[k];
      if (x.out_len) x.out_len[k] = 0;
      slen = f.out_len;
      const uint32_t eb = f.error;
      if (eb & ~FERR_CALLER_CAP) {
        enc_ok = false;
        if ((eb & retry_bits) && !(eb & (FERR_DEPTH | FERR_UNSUPPORTED)) && !x.boost) { retry.push_back(k); if (do_dec) x.npts_out[k] = 0; continue; }
        if (rc == CCV2_OK) {
          rc = (eb & FERR_DEPTH) ? CCV2_ERR_DEPTH : (eb & FERR_UNSUPPORTED) ? CCV2_ERR_UNSUPPORTED : CCV2_ERR_WORKSPACE;
          char b[96]; snprintf(b, sizeof b, "frame %d: device error bits 0x%x (encode)", k, eb); c->err = b;
        }
      } else if (slen && x.out && x.out[k]) {
        if ((eb & FERR_CALLER_CAP) || slen > x.out_cap[k]) { if (rc == CCV2_OK) { rc = CCV2_ERR_CAPACITY; c->err = "output buffer too small"; } }
        else {
          if (x.out_kind[k] == PK_PAGEABLE) {
            e = cudaMemcpy(x.out[k], (uint8_t *)x.stage.p + x.stage_off_stream[k], slen, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { c->err = std::string("cudaMemcpy (stream): ") + cudaGetErrorString(e); return store(CCV2_ERR_CUDA); }
          }
          x.out_len[k] = slen;
        }
      } else if (slen && !rt) { if (rc == CCV2_OK) { rc = CCV2_ERR_CAPACITY; c->err = "output buffer too small"; } }
      else if (slen && x.out_len) x.out_len[k] = slen;        // round trip without a stream buffer: report the size only
    }
    if (do_dec) {
      x.npts_out[k] = 0;
      if (rt && (!enc_ok || slen == 0)) continue;              // empty frame: nothing was written, nothing to decode
      const DecFrame &f = hd[k];
      if (f.error) {
        if ((f.error & (FERR_TREE_CAP | FERR_JPEG_CAP)) && !(f.error & (FERR_OUT_CAP | FERR_DEPTH | FERR_UNSUPPORTED)) && !x.boost) { retry.push_back(k); continue; }
        if (rc == CCV2_OK) {
          rc = (f.error & FERR_OUT_CAP) ? CCV2_ERR_CAPACITY : (f.error & FERR_DEPTH) ? CCV2_ERR_DEPTH : (f.error & FERR_UNSUPPORTED) ? CCV2_ERR_UNSUPPORTED
             : (f.error & (FERR_TREE_CAP | FERR_JPEG_CAP)) ? CCV2_ERR_WORKSPACE : CCV2_ERR_STREAM;
          char b[96]; snprintf(b, sizeof b, "frame %d: device error bits 0x%x (decode)", k, f.error); c->err = b;
        }
        continue;
      }
      const uint32_t nout = f.detail ? f.npoints_out : f.V;
      x.npts_out[k] = nout;
      if (x.pts_kind[k] == PK_PAGEABLE && nout) {
        e = cudaMemcpy(x.pts_out[k], (uint8_t *)x.stage.p + x.stage_off_pts[k], 32ull * nout, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { c->err = std::string("cudaMemcpy (points): ") + cudaGetErrorString(e); return store(CCV2_ERR_CUDA); }
      }
    }
  }
  if (c->trace) {
    fprintf(stderr, "ccv2 trace ticket=%d mode=%d frames=%d groups=%d (G %d, fe sets %d, ll sets %d) total %.1f ms\n", x.ticket, mode, nframes, x.ngroups, x.G, x.fe_nsets, x.ll_nsets, x.device_ms);
    for (int g = 0; g < x.ngroups; g++) {
      fprintf(stderr, "  group %2d:", g);
      for (size_t k = 0; k < x.trace_marks.size(); k++) if (x.trace_marks[k].group == g) {
        float t = -1; if (cudaEventElapsedTime(&t, c->timer_on ? c->ev_t0 : x.ev_start, x.ev_trace[k]) != cudaSuccess) { cudaGetLastError(); cudaEventElapsedTime(&t, x.ev_start, x.ev_trace[k]); }
        fprintf(stderr, " %s %.1f |", x.trace_marks[k].label, t);
      }
      fprintf(stderr, "\n");
    }
    for (int dir = 0; dir < 2; dir++) {                       // how evenly did the serial CTAs spread?
      if (dir == 0 ? !do_enc : !do_dec) continue;
      std::vector<int> per_sm(256, 0);
      for (int i = 0; i < nframes; i++) per_sm[(dir == 0 ? hf[i].serial_sm : hd[i].serial_sm) & 255]++;
      std::vector<int> hist(64, 0); int mx = 0;
      for (int s2 = 0; s2 < c->n_sm; s2++) { hist[std::min(63, per_sm[s2])]++; mx = std::max(mx, std::min(63, per_sm[s2])); }
      fprintf(stderr, "  %s serial CTAs per SM:", dir == 0 ? "encode" : "decode");
      for (int k = 0; k <= mx; k++) fprintf(stderr, " %dx%d", hist[k], k);
      fprintf(stderr, "\n");
    }
  }
  prof_collect(c);
  c->last_mode = mode;
  if (!x.boost) {
    if (do_enc) {
      c->enc_host.assign(hf, hf + nframes);
      c->enc_host_valid.assign(nframes, 0);
      // a frame's intermediates are intact while no later group took its ring sets
      const uint64_t fe_end = x.fe_seq0 + x.ngroups, ll_end = x.ll_seq0 + x.ngroups;
      for (int i = 0; i < nframes; i++) {
        const uint64_t g = (uint64_t)(i / x.G);
        c->enc_host_valid[i] = mode == 0 && c->fe.seq == fe_end && c->ll.seq == ll_end && c->fe.nsets == x.fe_nsets && c->ll.nsets == x.ll_nsets &&
                               x.fe_seq0 + g + x.fe_nsets >= fe_end && x.ll_seq0 + g + x.ll_nsets >= ll_end;
      }
      for (int i = nframes - 1; i >= 0; i--) if (hf[i].out_len) { for (int k = 0; k < 3; k++) c->metrics[k] = hf[i].coded[k]; c->frame_id = hf[i].frame_id; break; }
    } else {
      for (int i = nframes - 1; i >= 0; i--) if (!hd[i].error) { for (int k = 0; k < 3; k++) c->metrics[k] = hd[i].coded[k]; c->frame_id = hd[i].frame_id; break; }
    }
  }
  // frames whose workspace bound was exceeded: once more, alone, with bounds that cannot be (sparse deep octrees)
  for (int k : retry) {
    const int r = retry_frame(c, x, k, do_enc ? hf[k].frame_id : 0);
    if (r != CCV2_OK && rc == CCV2_OK) rc = r;
  }
  return store(rc);
}

static int retry_frame(ccv2_codec *c, CallCtx &x, int k, uint32_t fixed_id) {
  // x's arrays are the caller's; the sub-call writes frame k's results straight into them
  const void *p1 = x.pts ? x.pts[k] : nullptr; size_t n1 = x.npts ? x.npts[k] : 0;
  void *o1 = (x.out && x.out[k]) ? x.out[k] : nullptr; size_t oc1 = x.out_cap ? x.out_cap[k] : 0, ol1 = 0;
  const void *i1 = x.in ? x.in[k] : nullptr; size_t il1 = x.in_len ? x.in_len[k] : 0;
  void *po1 = x.pts_out ? x.pts_out[k] : nullptr; size_t pc1 = x.pts_cap ? x.pts_cap[k] : 0, np1 = 0;
  const bool want_out = x.out != nullptr;
  const uint32_t saved_id = c->frame_id; const uint64_t m0 = c->metrics[0], m1 = c->metrics[1], m2 = c->metrics[2];
  int rc = submit_call(c, x.mode, 1, x.pts ? &p1 : nullptr, x.npts ? &n1 : nullptr, want_out ? &o1 : nullptr, &oc1, &ol1,
                       x.in ? &i1 : nullptr, x.in_len ? &il1 : nullptr, x.pts_out ? &po1 : nullptr, x.pts_cap ? &pc1 : nullptr, &np1, true, fixed_id ? fixed_id : 0, nullptr);
  if (rc == CCV2_OK) rc = finish_call(c, c->calls[N_CALLS - 1]);
  if (x.out_len) x.out_len[k] = ol1;
  if (x.npts_out) x.npts_out[k] = np1;
  if (k != x.nframes - 1) { c->frame_id = saved_id; c->metrics[0] = m0; c->metrics[1] = m1; c->metrics[2] = m2; }
  else if (rc == CCV2_OK) {                                    // the last frame of the call defines the codec's metrics
    CallCtx &y = c->calls[N_CALLS - 1];
    if (x.mode != 1) { const EncFrame &f = ((EncFrame *)y.h_frames.p)[0]; if (f.out_len) { for (int q = 0; q < 3; q++) c->metrics[q] = f.coded[q]; c->frame_id = f.frame_id; } }
    else { const DecFrame &f = ((DecFrame *)y.h_dframes.p)[0]; if (!f.error) { for (int q = 0; q < 3; q++) c->metrics[q] = f.coded[q]; c->frame_id = f.frame_id; } }
  }
  return rc;
}

static int wait_ticket(ccv2_codec *c, int ticket) {
  for (auto &x : c->calls) if (x.busy && x.ticket == ticket) return finish_call(c, x);
  for (auto &d : c->done) if (d.ticket == ticket) return d.rc;
  return CCV2_OK;
}

extern "C" {

int ccv2_submit_encode(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                       void *const *out, const size_t *out_cap, size_t *out_len, int *ticket) {
  if (!c || !ticket || nframes < 0 || (nframes && (!pts || !npts || !out || !out_cap || !out_len))) return CCV2_ERR_ARG;
  return submit_call(c, 0, nframes, pts, npts, out, out_cap, out_len, nullptr, nullptr, nullptr, nullptr, nullptr, false, 0, ticket);
}
int ccv2_submit_decode(ccv2_codec *c, int nframes, const void *const *in, const size_t *in_len,
                       void *const *pts_out, const size_t *pts_cap, size_t *npts_out, int *ticket) {
  if (!c || !ticket || nframes < 0 || (nframes && (!in || !in_len || !pts_out || !pts_cap || !npts_out))) return CCV2_ERR_ARG;
  return submit_call(c, 1, nframes, nullptr, nullptr, nullptr, nullptr, nullptr, in, in_len, pts_out, pts_cap, npts_out, false, 0, ticket);
}
int ccv2_submit_roundtrip(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                          void *const *out, const size_t *out_cap, size_t *out_len,
                          void *const *pts_out, const size_t *pts_cap, size_t *npts_out, int *ticket) {
  if (!c || !ticket || nframes < 0 || (nframes && (!pts || !npts || !pts_out || !pts_cap || !npts_out))) return CCV2_ERR_ARG;
  if (out && (!out_cap || !out_len)) return CCV2_ERR_ARG;
  return submit_call(c, 2, nframes, pts, npts, out, out_cap, out_len, nullptr, nullptr, pts_out, pts_cap, npts_out, false, 0, ticket);
}
int ccv2_wait(ccv2_codec *c, int ticket) { if (!c) return CCV2_ERR_ARG; return wait_ticket(c, ticket); }

int ccv2_encode_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                      void *const *out, const size_t *out_cap, size_t *out_len) {
  int t = 0;
  const int rc = ccv2_submit_encode(c, nframes, pts, npts, out, out_cap, out_len, &t);
  return rc ? rc : wait_ticket(c, t);
}
int ccv2_decode_batch(ccv2_codec *c, int nframes, const void *const *in, const size_t *in_len,
                      void *const *pts_out, const size_t *pts_cap, size_t *npts_out) {
  int t = 0;
  const int rc = ccv2_submit_decode(c, nframes, in, in_len, pts_out, pts_cap, npts_out, &t);
  return rc ? rc : wait_ticket(c, t);
}
int ccv2_roundtrip_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                         void *const *out, const size_t *out_cap, size_t *out_len,
                         void *const *pts_out, const size_t *pts_cap, size_t *npts_out) {
  int t = 0;
  const int rc = ccv2_submit_roundtrip(c, nframes, pts, npts, out, out_cap, out_len, pts_out, pts_cap, npts_out, &t);
  return rc ? rc : wait_ticket(c, t);
}

// Device-side stopwatch over any number of calls: start marks the codec's control stream (everything submitted earlier is
// collected first), stop collects what is in flight and returns the time between the mark and the end of the last call.
int ccv2_timer_start(ccv2_codec *c) {
  if (!c) return CCV2_ERR_ARG;
  CU(cudaSetDevice(c->device));
  finish_all(c); drain(c);
  CU(cudaEventRecord(c->ev_t0, c->main_stream));
  c->timer_on = true;
  return CCV2_OK;
}
int ccv2_timer_stop(ccv2_codec *c, float *ms) {
  if (!c || !ms) return CCV2_ERR_ARG;
  CU(cudaSetDevice(c->device));
  const int rc = finish_all(c);                              // host-synchronous: everything submitted has completed
  CU(cudaEventRecord(c->ev_t1, c->main_stream));
  CU(cudaEventSynchronize(c->ev_t1));
  CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
  c->timer_on = false;
  return rc;
}

int ccv2_get_output_cloud(ccv2_codec *c, int frame, void *points_out, size_t cap_points, size_t *npoints) {
  if (!c || frame < 0 || frame >= (int)c->enc_host.size() || !npoints) return CCV2_ERR_ARG;
  finish_all(c);
  if (c->last_mode != 0 || !c->enc_host_valid[frame]) { c->err = "the output cloud is only available right after ccv2_encode_batch (and while the frame's workspace has not been handed on)"; return CCV2_ERR_UNSUPPORTED; }
  CU(cudaSetDevice(c->device));
  const EncFrame &f = c->enc_host[frame];
  if (f.error || c->last_enc_params.detail) { *npoints = 0; return CCV2_OK; }   // detail mode: the reference's callback leaves output_ empty (impl.hpp:1525-1541)
  *npoints = f.V;
  if (f.V == 0) return CCV2_OK;
  if (f.V > cap_points || !points_out) return CCV2_ERR_CAPACITY;
  const bool dev = is_device_ptr(points_out);
  uint8_t *dst = (uint8_t *)points_out;
  if (!dev) { CU(c->out_cloud.ensure(32ull * f.V)); dst = (uint8_t *)c->out_cloud.p; }
  output_cloud_kernel<<<(f.V + 255) / 256, 256, 0, c->fin_stream>>>(f, c->last_enc_params, dst);
  CU(cudaGetLastError());
  if (!dev) CU(cudaMemcpyAsync(points_out, dst, 32ull * f.V, cudaMemcpyDeviceToHost, c->fin_stream));
  CU(cudaStreamSynchronize(c->fin_stream));
  return CCV2_OK;
}

// ---- tile mode (tile_kernels.cuh) -----------------------------------------------------------------------------------
// Stable partition of one frame into 2^tile_bits spatial tiles.  pts_out (device or host, n records) receives the points
// grouped by tile, original order kept inside a tile; tile_offsets[t] .. tile_offsets[t + 1] is tile t.
int ccv2_split_tiles(ccv2_codec *c, const void *pts, size_t n, int tile_bits, void *pts_out, size_t *tile_offsets) {
  if (!c || !tile_offsets || (tile_bits != 3 && tile_bits != 6) || n >= (1u << 28) || (n && (!pts || !pts_out))) return CCV2_ERR_ARG;
  CU(cudaSetDevice(c->device));
  finish_all(c);
  const int k = tile_bits / 3; const uint32_t nt = 1u << tile_bits;
  for (uint32_t t = 0; t <= nt; t++) tile_offsets[t] = 0;
  if (n == 0) return CCV2_OK;
  const uint32_t nblocks = (uint32_t)((n + TILE_BLOCK - 1) / TILE_BLOCK);
  const bool din = is_device_ptr(pts), dout = is_device_ptr(pts_out);
  const size_t stage_in = din ? 0 : (32 * n + 255) & ~size_t(255), stage_out = dout ? 0 : (32 * n + 255) & ~size_t(255);
  const size_t cnt_bytes = ((size_t)nt * nblocks * 4 + 255) & ~size_t(255);
  CU(c->tile_aux.ensure(stage_in + stage_out + cnt_bytes + (nt + 1) * 8 + 256));
  uint8_t *w = (uint8_t *)c->tile_aux.p;
  const uint8_t *src = din ? (const uint8_t *)pts : w;
  uint8_t *dst = dout ? (uint8_t *)pts_out : w + stage_in;
  uint32_t *counts = (uint32_t *)(w + stage_in + stage_out);
  uint64_t *offs = (uint64_t *)(w + stage_in + stage_out + cnt_bytes);
  cudaStream_t st = c->fin_stream;
  if (!din) CU(cudaMemcpyAsync(w, pts, 32 * n, cudaMemcpyHostToDevice, st));
  tile_count_kernel<<<nblocks, TILE_BLOCK, 0, st>>>(src, (uint32_t)n, k, counts, nblocks);
  tile_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, nt, offs);
  tile_scatter_kernel<<<nblocks, TILE_BLOCK, 0, st>>>(src, (uint32_t)n, k, counts, nblocks, dst);
  CU(cudaGetLastError());
  uint64_t h_offs[TILE_MAX + 1];
  CU(cudaMemcpyAsync(h_offs, offs, (nt + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (!dout) CU(cudaMemcpyAsync(pts_out, dst, 32 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (uint32_t t = 0; t <= nt; t++) tile_offsets[t] = (size_t)h_offs[t];
  return CCV2_OK;
}

// One frame -> one reference-format stream per non-empty tile.  Encodes tiles first_tile, first_tile + tile_step, ...
// (a rank of a node takes first_tile = rank, tile_step = world size; 0 and 1 for everything).  out / out_cap / out_len are
// indexed by TILE (2^tile_bits entries; tiles this call does not own are left untouched, empty tiles report 0 bytes);
// tile_npts (may be NULL) receives the number of points of every tile.  Frame ids continue the codec's counter in tile order.
int ccv2_encode_tiles(ccv2_codec *c, const void *pts, size_t n, int tile_bits, int first_tile, int tile_step,
                      void *const *out, const size_t *out_cap, size_t *out_len, size_t *tile_npts) {
  if (!c || !out || !out_cap || !out_len || (tile_bits != 3 && tile_bits != 6) || first_tile < 0 || tile_step < 1) return CCV2_ERR_ARG;
  const int nt = 1 << tile_bits;
  CU(cudaSetDevice(c->device));
  finish_all(c);
  CU(c->tile_pts.ensure(32 * n + 256));
  size_t offs[TILE_MAX + 1];
  int rc = ccv2_split_tiles(c, pts, n, tile_bits, c->tile_pts.p, offs);
  if (rc != CCV2_OK) return rc;
  std::vector<const void *> tp; std::vector<size_t> tn, tcap, tlen; std::vector<void *> to; std::vector<int> owner;
  for (int t = 0; t < nt; t++) {
    if (tile_npts) tile_npts[t] = offs[t + 1] - offs[t];
    if (t < first_tile || (t - first_tile) % tile_step) continue;
    out_len[t] = 0;
    if (offs[t + 1] == offs[t]) continue;                                     // empty tile: no frame (the reference writes nothing for an empty cloud)
    tp.push_back((const uint8_t *)c->tile_pts.p + 32 * offs[t]); tn.push_back(offs[t + 1] - offs[t]); to.push_back(out[t]); tcap.push_back(out_cap[t]); owner.push_back(t);
  }
  tlen.assign(tp.size(), 0);
  if (tp.empty()) return CCV2_OK;
  rc = ccv2_encode_batch(c, (int)tp.size(), tp.data(), tn.data(), to.data(), tcap.data(), tlen.data());
  for (size_t i = 0; i < owner.size(); i++) out_len[owner[i]] = tlen[i];
  return rc;
}

// computeQualityMetric (quality_metrics_impl.hpp:82-239): exact nearest neighbours both ways, see quality_kernels.cuh
int ccv2_quality_metrics(ccv2_codec *c, const void *cloud_a, size_t na, const void *cloud_b, size_t nb, ccv2_quality *out) {
  if (!c || !out || (na && !cloud_a) || (nb && !cloud_b) || na >= (1u << 28) || nb >= (1u << 28)) return CCV2_ERR_ARG;
  CU(cudaSetDevice(c->device));
  finish_all(c);
  memset(out, 0, sizeof *out);
  out->in_point_count = na; out->out_point_count = nb;
  if (na == 0 || nb == 0) return CCV2_OK;
  const bool da = is_device_ptr(cloud_a), db = is_device_ptr(cloud_b);
  const size_t sa = da ? 0 : (32 * na + 255) & ~size_t(255), sb = db ? 0 : (32 * nb + 255) & ~size_t(255);
  CU(c->out_cloud.ensure(sa + sb + 2 * sizeof(QualityAccum) + 256));
  uint8_t *w = (uint8_t *)c->out_cloud.p;
  const uint8_t *pa = da ? (const uint8_t *)cloud_a : w, *pb = db ? (const uint8_t *)cloud_b : w + sa;
  QualityAccum *acc = (QualityAccum *)(w + sa + sb);
  cudaStream_t st = c->fin_stream;
  if (!da) CU(cudaMemcpyAsync(w, cloud_a, 32 * na, cudaMemcpyHostToDevice, st));
  if (!db) CU(cudaMemcpyAsync(w + sa, cloud_b, 32 * nb, cudaMemcpyHostToDevice, st));
  QualityAccum h[2]; memset(h, 0, sizeof h);
  for (int d = 0; d < 2; d++) for (int k = 0; k < 3; k++) h[d].max_xyz[k] = -3.0e38f;
  CU(cudaMemcpyAsync(acc, h, sizeof h, cudaMemcpyHostToDevice, st));
  quality_nn_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(pa, (uint32_t)na, pb, (uint32_t)nb, acc, 1);
  quality_nn_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(pb, (uint32_t)nb, pa, (uint32_t)na, acc + 1, 0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // quality_metrics_impl.hpp:160-238
  float max_a, max_b; memcpy(&max_a, &h[0].max_d2_bits, 4); memcpy(&max_b, &h[1].max_d2_bits, 4);
  max_a = std::sqrt(max_a); max_b = std::sqrt(max_b);
  const double rms_a = std::sqrt(h[0].sum_d2 / (double)na), rms_b = std::sqrt(h[1].sum_d2 / (double)nb);
  const float dist_h = std::max(max_a, max_b);
  const float dist_rms = (float)std::max(rms_a, rms_b);
  const float energy = h[0].max_xyz[0] * h[0].max_xyz[0] + h[0].max_xyz[1] * h[0].max_xyz[1] + h[0].max_xyz[2] * h[0].max_xyz[2];
  out->left_hausdorff = max_a; out->right_hausdorff = max_b; out->symm_hausdorff = dist_h;
  out->left_rms = (float)rms_a; out->right_rms = (float)rms_b; out->symm_rms = dist_rms;
  out->psnr_db = (float)(10 * std::log10(energy / (dist_rms * dist_rms)));
  for (int k = 0; k < 3; k++) out->psnr_yuv[k] = 10 * std::log10(1.0 / (h[0].mse_yuv[k] / (double)na));
  return CCV2_OK;
}

int ccv2_debug_fetch(ccv2_codec *c, int frame, int what, void *host_buf, size_t cap, size_t *len) {
  if (!c || frame < 0 || frame >= (int)c->enc_host.size() || !len) return CCV2_ERR_ARG;
  finish_all(c);
  CU(cudaSetDevice(c->device));
  const EncFrame &f = c->enc_host[frame];
  const void *src = nullptr; size_t n = 0;
  if (what != 5 && (c->last_mode != 0 || !c->enc_host_valid[frame])) { c->err = "encode intermediates are only available right after ccv2_encode_batch"; return CCV2_ERR_UNSUPPORTED; }
  ccv2_frame_info info;
  switch (what) {
    case 0: src = f.leaf_key; n = (size_t)f.V * 8; break;
    case 1: src = f.tree; n = f.B; break;
    case 2: src = f.avg; n = (size_t)f.V * 3; break;
    case 3: src = f.cpay; n = f.ncolor; break;
    case 4: if (f.packed) { c->err = "this frame was sorted as packed (code, colour) words: there are no point indices (CCV2_PACKED=0 keeps them)"; return CCV2_ERR_UNSUPPORTED; }
            src = f.vals[f.npasses & 1]; n = (size_t)f.n_finite * 4; break;
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

#include "ccv2_inter.cuh"

// common.cuh -- shared device structs and primitives for libccv2 (sm_100a).
//
// Layout rule: one EncFrame / DecFrame record per frame lives in device memory; every kernel is launched
// over a *group* of frames with blockIdx.y = frame-in-group and reads its sizes from the record, so the
// host never has to synchronise between stages (all sizes -- depth, V, B, J -- are device-side values).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CCV2_MAX_DEPTH 21
#define FULL_MASK 0xFFFFFFFFu
#define NONE_U32 0xFFFFFFFFu

// error bits written by kernels into frame.error
enum : uint32_t {
  FERR_DEPTH = 1u << 0,       // realised depth > CCV2_MAX_DEPTH
  FERR_TREE_CAP = 1u << 1,    // tree bytes exceed workspace cap
  FERR_STREAM_CAP = 1u << 2,  // compressed stream exceeds arena/caller cap
  FERR_JPEG_CAP = 1u << 3,    // jpeg buffers exceeded
  FERR_BAD_STREAM = 1u << 4,  // decoder: malformed input
  FERR_OUT_CAP = 1u << 5,     // decoder: caller's point buffer too small
  FERR_UNSUPPORTED = 1u << 6, // stream / cloud needs something outside the implemented scope (e.g. a SNAKE image higher than libjpeg's 65500 rows)
  FERR_CALLER_CAP = 1u << 7,  // encoder: the caller's stream buffer is too small
};

enum : int { TK_SORT0 = 0, TK_LEAF = 8, TK_HUFF = 9, TK_STUFF = 10, TK_NODES = 11, TK_EXPAND = 12, TK_DETAIL = 13, TK_COUNT = 16 };

struct __align__(16) JpegTables {   // per codec, device resident
  uint16_t q[2][64];           // quant tables, natural order (lum, chroma)
  uint16_t dc_code[2][12]; uint8_t dc_len[2][12];
  uint16_t ac_code[2][256]; uint8_t ac_len[2][256];
  uint8_t header[623];         // JFIF header template for this quality (height patched per frame)
  uint8_t zz[64];              // zigzag index -> natural index
};

struct EncParams {             // by value to kernels
  double res, inv_res;         // octree resolution; inv_res valid iff res_pow2
  int res_pow2;
  int do_color, color_type, color_reduction, do_centroid;
  int prefix_len;              // points handled exactly by the single-CTA bbox kernel
  int allow_packed;            // voxel-grid mode without centroids: a frame of depth <= 13 sorts ONE 64-bit word per point (40-bit code << 24 | b,g,r)
  int detail;                  // doVoxelGridDownDownSampling = false: per-point residuals and colour differences (impl.hpp:1525-1541)
  float point_res_f;           // [PCL] PointCoding::pointCompressionResolution_ (setPrecision(float))
};

// One entry per change of the bounding box while the cloud is added in input order ([PCL] adoptBoundingBoxToPoint):
// entry 0 is the definition by the first finite point, every later one a growth step (re-rooting).  keygen needs them
// to give a point the key PCL gave it: computed against the box in force when the point was added, plus 1 << depth_before
// on every axis whose minimum moved in a later growth step (oracle/ccv2_oracle.c orc_bbox_keys).
#define CCV2_MAX_EVENTS 32
struct BoxEvent { double mn[3]; uint32_t idx, depth_before, mask, _pad; };   // mn: box minimum after the event; mask bit a: axis a moved

struct EncFrame {
  // region of the group-slot workspace that must be zero before the pipeline starts (zero_region_kernel)
  uint8_t *zero_ptr; uint64_t zero_bytes;
  // input
  const uint8_t *pts; uint32_t n; uint32_t serial_sm;   // serial_sm: SM the range-coder CTA ran on (CCV2_TRACE)
  // bbox / keys
  double bmin[3], bmax[3];
  uint32_t depth, defined, n_finite, violator, rekey, npasses;
  uint32_t n_events, packed; BoxEvent ev[CCV2_MAX_EVENTS];   // packed: sort elements are (code << 24 | bgr), no index array (see keygen_kernel)
  // leaves
  uint32_t V, B;
  // jpeg geometry / sizes
  uint32_t img_h, mcu_h, jbits, J;
  uint32_t ncolor;             // bytes handed to the range coder for the colour layer
  uint32_t ncen;               // centroid bytes (3V) when enabled
  uint32_t ticket[TK_COUNT];
  // results
  uint32_t error, frame_id;
  uint32_t frame_id_fixed, _padf;   // != 0: a retried frame keeps the id it was given the first time (frame_setup_kernel)
  uint8_t *out_ptr; uint64_t out_cap;   // where the finished stream goes (caller's device buffer, the device alias of a pinned host buffer, or a staging area); may be null
  uint64_t out_len; uint64_t coded[3];
  uint32_t rc_len[5];          // coded bytes per layer incl. table (tree, centroid, colour, point differences, colour differences)
  uint32_t rc_int_len;         // coded bytes of the int-coded point counts (detail mode)
  uint32_t npd, ncd;           // detail mode: bytes of point differences (3 per point) / colour differences (3 per point of a multi-point voxel)
  uint32_t itsize, _pad1;      // detail mode: frequency table size of the int coder (after the final increment)
  // group-slot workspace
  uint64_t *keys[2]; uint32_t *vals[2];
  uint32_t *ghist;             // [8][256]
  uint32_t *sort_status;       // [8][tiles_max][256]
  uint32_t tiles_max, _pad2;
  uint64_t *leaf_key; uint32_t *leaf_start; uint32_t *leaf_off; uint8_t *first_new;
  uint64_t *scan_status;       // chained-scan status words (leaf / huff / stuff), 3 x scan_tiles_max
  uint32_t scan_tiles_max, _pad3;
  uint8_t *avg;                // 3 bytes per leaf (B,G,R)
  int16_t *coef;               // jpeg coefficients [mcu][6][64] zigzag order
  uint32_t *jbits_buf;         // unstuffed entropy-coded bits
  uint32_t jbits_cap_words, _pad4;
  uint8_t *line_slots; uint32_t *line_len, *line_off; uint32_t lines_cap, _padl;   // LINES colour mode staging
  // per-frame persistent buffers (live until the frame's stream is assembled)
  uint8_t *tree; uint32_t tree_cap; uint32_t _pad5;
  uint8_t *cen;                // centroid residual bytes
  uint8_t *cpay; uint32_t cpay_cap; uint32_t _pad6;   // colour payload (jpeg file or raw averages)
  uint32_t *hist;              // [5][256]
  uint8_t *stream; uint64_t stream_cap;
  uint8_t *rc_tmp[4]; uint32_t rc_tmp_cap[4];         // range-coded centroid / colour / point-difference / colour-difference layers before assembly
  // detail mode (impl.hpp:1525-1541, 1728-1757)
  uint32_t *counts;            // points per voxel (point_count_data_vector_)
  uint32_t *cd_off;            // per voxel: index of its first colour-difference triple (front-end ring)
  uint8_t *pdiff, *cdiff;      // [PCL] PointCoding / ColorCoding differential vectors
  uint64_t *itab; uint32_t itab_cap, _pad7;           // cumulative frequency table of the int coder
  uint8_t *rc_int; uint32_t rc_int_cap, _pad9;        // int-coded counts before assembly
};

struct DecFrame {
  uint8_t *zero_ptr; uint64_t zero_bytes;      // scan status + JPEG coefficients: zero before the pipeline starts
  const uint8_t *in; uint64_t in_len;
  uint8_t *out_pts; uint64_t out_cap;       // points
  // header
  double res, bmin[3], bmax[3];
  uint64_t point_count;
  uint32_t depth, cct, frame_id;
  uint32_t data_with_color, do_centroid, color_bits;
  uint32_t B, ncen, ncol;
  uint32_t n_bottom, V;
  uint32_t walk_done, huff_done;                 // DFS walk / JPEG Huffman decode already done inside the entropy stage
  // lane-per-stream entropy stage (dec_lps_kernels.cuh): header result, tree layer span, colour speculation
  uint32_t head_ok, tree_n, tree_ok, spec_found, spec_state, spec_ncol, spec_jerr, _pad8;
  uint64_t tree_pos, tree_end, spec_pos, spec_coded;
  uint32_t img_w, img_h, mcu_w, mcu_h, n_blocks;
  uint32_t ticket[TK_COUNT];
  uint32_t error, serial_sm;
  uint64_t coded[3];
  // buffers
  uint8_t *tree; uint32_t tree_cap, _pad1;
  uint8_t *cen; uint32_t cen_cap, _pad2;
  uint8_t *col; uint32_t col_cap, _pad3;
  uint64_t *node_prefix; uint8_t *node_byte; uint32_t node_cap, _pad4;   // bottom-level branches (level depth-1)
  uint64_t *l2_prefix; uint8_t *l2_mask; uint32_t *l2_off;                 // level depth-2 branches recorded by the pipelined walker
  uint32_t n_l2, l2_valid;
  int16_t *coef; uint32_t coef_cap_blocks, _pad5;
  uint8_t *planes; uint32_t planes_cap, _pad6;   // Y | Cb | Cr
  uint16_t *qt;                                  // [2][64] zigzag order, written by the jpeg header parse
  uint8_t *scan; uint32_t scan_start, scan_len;  // de-stuffed entropy-coded segment of the jpeg
  uint32_t dht_off[4], dht_n[4];                 // offsets of the DHT bits[16] (values follow) inside col: dc0 dc1 ac0 ac1
  uint32_t n_lines, lines_cap;                   // LINES colour mode: per-line offset/length/width, quant tables, row planes
  uint32_t *line_off, *line_len, *line_w; uint16_t *line_qt; uint8_t *line_planes;
  uint64_t *scan_status; uint32_t scan_tiles_max, _pad7;
  // detail mode (entropyDecoding impl.hpp:1802-1832, deserializeTreeCallback :1592-1613)
  uint32_t detail, ncounts; uint64_t npdiff, ncdiff;
  float point_res_f; uint32_t _pad11;            // [PCL] readFrameHeader: point_coder_.setPrecision(float(point_resolution))
  uint32_t *counts; uint32_t counts_cap, _pad9;
  uint8_t *pdiff, *cdiff; uint64_t pdiff_cap;
  uint64_t *itab; uint32_t itab_cap, _pad10;
  uint64_t *dleaf_key;                           // Morton code of every voxel in stream order (detail_points_kernel expands them)
  uint32_t npoints_out, _pad12;                  // detail mode: points written (V stays the voxel count)
};

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) { return *(const volatile uint64_t *)p; }
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *(const volatile uint32_t *)p; }

// Shared memory through 32-bit window addresses: with generic pointers the compiler rebuilds the CTA's shared window base
// (S2R SR_CgaCtaId + LEA) in front of every access, which is a third of the instructions of the serial one-lane loops.
__device__ __forceinline__ uint32_t smem_addr(const volatile void *p) { return (uint32_t)__cvta_generic_to_shared(const_cast<const void *>(p)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_volatile_u32(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// ---- TMA 1-D bulk copies (cp.async.bulk, sm_90+): a contiguous global span lands in shared memory as ONE asynchronous
// transfer issued by one thread and signalled through an mbarrier, instead of one load instruction per thread and element.
// src, dst and bytes must be multiples of 16.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t a = smem_addr(bar);
  while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(a), "r"(phase) : "memory");
}

// spread the low 21 bits of v so that bit i lands at bit 3i
__device__ __forceinline__ uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0x1FFFFFull;
  x = (x | (x << 32)) & 0x1F00000000FFFFull;
  x = (x | (x << 16)) & 0x1F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ uint32_t compact3(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x | (x >> 2)) & 0x10C30C30C30C30C3ull;
  x = (x | (x >> 4)) & 0x100F00F00F00F00Full;
  x = (x | (x >> 8)) & 0x1F0000FF0000FFull;
  x = (x | (x >> 16)) & 0x1F00000000FFFFull;
  x = (x | (x >> 32)) & 0x1FFFFFull;
  return (uint32_t)x;
}
// Morton code with x as the most significant bit of each triple (child index = x<<2 | y<<1 | z, [PCL] OctreeKey)
__device__ __forceinline__ uint64_t morton_xyz(uint32_t kx, uint32_t ky, uint32_t kz) {
  return (spread3(kx) << 2) | (spread3(ky) << 1) | spread3(kz);
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(FULL_MASK, v, o); if (lane_id() >= (uint32_t)o) v += t; }
  return v;
}

// Block-wide exclusive scan of one u64 per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block aggregate (valid in every thread).
__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t *total, uint64_t *s_warp /* >= 33 */) {
  uint32_t lane = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint64_t inc = warp_incl_scan_u64(v);
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint64_t x = lane < nw ? s_warp[lane] : 0;
    uint64_t xi = warp_incl_scan_u64(x);
    s_warp[lane] = xi - x;
    if (lane == 31) s_warp[32] = xi;
  }
  __syncthreads();
  uint64_t r = s_warp[w] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

// Decoupled look-back (single-pass chained scan).  status[tile]: bits 63:62 = 0 empty / 1 aggregate / 2 inclusive
// prefix; low 62 bits = value.  Must be called by all 32 lanes of ONE warp of the block; returns the exclusive
// prefix of `tile` in every lane.  Tiles must be handed out by an atomic ticket so predecessors are running.
#define SCAN_FLAG_A (1ull << 62)
#define SCAN_FLAG_P (2ull << 62)
#define SCAN_VMASK ((1ull << 62) - 1)
__device__ __forceinline__ uint64_t scan_lookback(uint64_t *status, uint32_t tile, uint64_t aggregate) {
  uint32_t lane = lane_id();
  if (tile == 0) {
    if (lane == 0) { *(volatile uint64_t *)&status[0] = SCAN_FLAG_P | aggregate; }
    return 0;
  }
  if (lane == 0) *(volatile uint64_t *)&status[tile] = SCAN_FLAG_A | aggregate;
  uint64_t excl = 0;
  int look = (int)tile - 1;
  for (;;) {
    int idx = look - (int)lane;
    uint64_t s;
    do {
      s = idx >= 0 ? ld_volatile_u64(&status[idx]) : SCAN_FLAG_P;
    } while (__any_sync(FULL_MASK, (s >> 62) == 0));
    uint32_t pm = __ballot_sync(FULL_MASK, (s >> 62) == 2);
    if (pm) {
      uint32_t first = __ffs(pm) - 1;
      excl += warp_sum_u64(lane <= first ? (s & SCAN_VMASK) : 0);
      break;
    }
    excl += warp_sum_u64(s & SCAN_VMASK);
    look -= 32;
  }
  if (lane == 0) *(volatile uint64_t *)&status[tile] = SCAN_FLAG_P | (excl + aggregate);
  return excl;
}

// dec_serial_kernel (the fallback path) still maps frames to blocks by index: the block that serves frame f of a group
// is (first_slot + f) mod gridDim.x, so the few frames that need it do not all start on the same SMs.
__device__ __forceinline__ int steered_frame(uint32_t first_slot, uint32_t group_frames) {
  const uint32_t g = gridDim.x;
  const uint32_t f = (blockIdx.x + g - first_slot % g) % g;
  return f < group_frames ? (int)f : -1;
}

// ---- serial CTAs: at most `cap` per SM, enforced by the hardware ------------------------------------------------------
// The range-coder kernels run one CTA per frame and are latency bound, so what matters is how many of them share an
// SM.  Which SM a CTA lands on is the block scheduler's choice, and with eight groups launching serial kernels between
// each other's parallel kernels it is far from even (measured with CCV2_TRACE, which records %smid per frame: up to
// 10 serial CTAs on one SM while 9 SMs had none; one group ran at the solo time of 189 ms, the others took 400-600 ms,
// against 230 ms when every SM holds exactly four).  Steering by block index or letting surplus CTAs exit when their
// SM is full (both tried) depend on the scheduler's policy.  What does not: a serial CTA asks for so much dynamic
// shared memory that only `cap` of them fit on an SM (cap = ceil(frames in the call / SMs)); the scheduler then has
// to put the next one on another SM, whatever else is running.  The memory itself is not used.
__device__ __forceinline__ uint32_t sm_id() { uint32_t v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

// Zeroing as a kernel on the group's own stream.  cudaMemsetAsync goes through a shared in-order engine: a memset queued
// behind a long kernel on one stream held back the memsets (and so the start) of every other group (measured with
// CCV2_TRACE: groups 1..7 of a round trip started only when group 0's encode had finished).
template <typename Frame>
__global__ void __launch_bounds__(256) zero_region_kernel(Frame *frames) {
  Frame &f = frames[blockIdx.y];
  const uint64_t n16 = f.zero_bytes / 16;                   // regions are 256-byte aligned multiples of 256
  uint4 *p = (uint4 *)f.zero_ptr;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) p[i] = z;
}

// Device records -> their pinned host copy through zero-copy stores (see submit_call: the copy engine's in-order queue is the wrong
// road for a small transfer that must not wait for large ones queued earlier).
__global__ void __launch_bounds__(256) records_home_kernel(uint64_t *dst_host, const uint64_t *src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst_host[i] = src[i];
}

// Granlund-Montgomery division of a 32-bit value by an invariant d (2 <= d < 2^31): q = n / d exactly.
struct FastDiv { uint32_t m, sh; };
__host__ __device__ inline FastDiv fastdiv_make(uint32_t d) {
  uint32_t l = 0; while ((1ull << l) < d) l++;
  FastDiv f; f.m = (uint32_t)((((1ull << 32) * ((1ull << l) - d)) / d) + 1); f.sh = l - 1; return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
  uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> 1)) >> f.sh;
}

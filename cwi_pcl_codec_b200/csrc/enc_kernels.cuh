// enc_kernels.cuh -- encode-side kernels: bbox growth, Morton keys, radix sort, leaf segmentation,
// occupancy bytes in DFS order, colour averages, centroid residuals.
// Reference behaviour restated (see oracle/ccv2_oracle.c for the CPU statement of the same steps):
//   [PCL] OctreePointCloud::addPointsFromInputCloud / adoptBoundingBoxToPoint / genOctreeKeyforPoint (impl.hpp:99)
//   [PCL] Octree2BufBase::serializeTree (impl.hpp:166) and serializeTreeCallback (impl.hpp:1509-1578)
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// bbox growth with the reference's sequential (input-order) semantics.  One CTA per frame walks the
// range [start, end) in chunks of blockDim.x points; every thread carries the same box state.
// mode 0: exact prefix [0, prefix_len).  mode 1: slow path, continues from frame.violator to n (only
// when the full-grid verification in keygen found a point outside the prefix box).
// ------------------------------------------------------------------------------------------------
struct Box { double mn[3], mx[3]; uint32_t depth; uint32_t defined; };

__device__ __forceinline__ bool box_violates(const Box &b, const double p[3]) {
  if (!b.defined) return true;
  return p[0] < b.mn[0] || p[1] < b.mn[1] || p[2] < b.mn[2] || p[0] >= b.mx[0] || p[1] >= b.mx[1] || p[2] >= b.mx[2];
}
// adoptBoundingBoxToPoint for one point; loops until the point fits. Returns false on depth overflow.
// Every change of the box is logged (log != nullptr only in the one thread that writes the frame record; nev is
// carried by all threads): keygen replays the log to key every point against the box of its own time.
__device__ inline bool box_adopt(Box &b, const double p[3], double res, uint32_t idx, uint32_t &nev, BoxEvent *log) {
  const double eps = 1.1920928955078125e-07;   // (double) numeric_limits<float>::epsilon()
  for (int guard = 0; guard < 64; guard++) {
    if (!b.defined) {
      for (int a = 0; a < 3; a++) { b.mn[a] = __dsub_rn(p[a], res / 2); b.mx[a] = __dadd_rn(p[a], res / 2); }
      // getKeyBitSize() with leaf_count_ == 0
      uint32_t mk = 2;
      for (int a = 0; a < 3; a++) {
        double t = ceil(__ddiv_rn(__dsub_rn(__dsub_rn(b.mx[a], b.mn[a]), eps), res));
        uint32_t k = t >= 4294967295.0 ? 0xFFFFFFFFu : (t > 0 ? (uint32_t)t : 0u);
        if (k > mk) mk = k;
      }
      uint32_t d = 0; while ((1ull << d) < mk) d++;          // ceil(log2(mk) - eps) for integer mk < 2^23
      if (d > CCV2_MAX_DEPTH) return false;
      b.depth = d;
      double side = __dmul_rn((double)(1u << d), res);
      for (int a = 0; a < 3; a++) {
        double over = __dmul_rn(__dsub_rn(side, __dsub_rn(b.mx[a], b.mn[a])), 0.5);
        if (over > eps) { b.mn[a] = __dsub_rn(b.mn[a], over); b.mx[a] = __dadd_rn(b.mx[a], over); }
      }
      b.defined = 1;
      if (log && nev < CCV2_MAX_EVENTS) { BoxEvent &e = log[nev]; e.idx = idx; e.depth_before = 0; e.mask = 0; e._pad = 0; for (int a = 0; a < 3; a++) e.mn[a] = b.mn[a]; }
      nev++;
      continue;
    }
    bool up[3], any = false;
    for (int a = 0; a < 3; a++) { up[a] = p[a] >= b.mx[a]; any |= up[a] | (p[a] < b.mn[a]); }
    if (!any) return true;
    if (b.depth >= CCV2_MAX_DEPTH) return false;
    double side = __dmul_rn((double)(1u << b.depth), res);
    uint32_t mask = 0;
    for (int a = 0; a < 3; a++) if (!up[a]) { b.mn[a] = __dsub_rn(b.mn[a], side); mask |= 1u << a; }
    if (log && nev < CCV2_MAX_EVENTS) { BoxEvent &e = log[nev]; e.idx = idx; e.depth_before = b.depth; e.mask = mask; e._pad = 0; for (int a = 0; a < 3; a++) e.mn[a] = b.mn[a]; }
    nev++;
    b.depth++;
    side = __dsub_rn(__dmul_rn((double)(1u << b.depth), res), eps);
    for (int a = 0; a < 3; a++) b.mx[a] = __dadd_rn(b.mn[a], side);
  }
  return false;
}

__global__ void __launch_bounds__(1024) bbox_kernel(EncFrame *frames, EncParams P, int mode) {
  EncFrame &f = frames[blockIdx.x];
  __shared__ uint32_t s_first[2][32];
  __shared__ float s_pt[2][3];
  uint32_t start, end, nev;
  Box b;
  if (mode == 0) {
    start = 0; end = min(f.n, (uint32_t)P.prefix_len); nev = 0;
    b.defined = 0; b.depth = 0;
    for (int a = 0; a < 3; a++) { b.mn[a] = 0; b.mx[a] = 0; }
    // a box the host defined before the first point ([PCL] defineBoundingBox, as simplifyPCloud and the macroblock trees
    // do, impl.hpp:336,427): the record carries it together with its entry 0 of the event log; points outside still grow it
    if (f.defined) { b.defined = 1; b.depth = f.depth; nev = f.n_events; for (int a = 0; a < 3; a++) { b.mn[a] = f.bmin[a]; b.mx[a] = f.bmax[a]; } }
  } else {
    if (f.violator == NONE_U32 || (f.error & FERR_DEPTH)) return;
    start = f.violator; end = f.n; nev = f.n_events;
    b.defined = f.defined; b.depth = f.depth;
    for (int a = 0; a < 3; a++) { b.mn[a] = f.bmin[a]; b.mx[a] = f.bmax[a]; }
  }
  bool fail = false;
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  int ph = 0;
  for (uint32_t base = start; base < end && !fail; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    double p[3] = {0, 0, 0};
    bool fin = false;
    if (i < end) {
      float4 q = __ldg((const float4 *)(f.pts + 32ull * i));
      fin = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
      p[0] = q.x; p[1] = q.y; p[2] = q.z;
    }
    for (;;) {
      bool viol = fin && box_violates(b, p);
      uint32_t bal = __ballot_sync(FULL_MASK, viol);
      if (lane == 0) s_first[ph][w] = bal ? (w * 32 + __ffs(bal) - 1) : NONE_U32;
      __syncthreads();
      uint32_t m = lane < (blockDim.x >> 5) ? s_first[ph][lane] : NONE_U32;
      m = __reduce_min_sync(FULL_MASK, m);
      if (m == NONE_U32) { ph ^= 1; break; }
      if (threadIdx.x == m) { s_pt[ph][0] = (float)p[0]; s_pt[ph][1] = (float)p[1]; s_pt[ph][2] = (float)p[2]; }
      __syncthreads();
      double pv[3] = { (double)s_pt[ph][0], (double)s_pt[ph][1], (double)s_pt[ph][2] };
      if (!box_adopt(b, pv, P.res, base + m, nev, threadIdx.x == 0 ? f.ev : nullptr)) { fail = true; ph ^= 1; break; }
      ph ^= 1;
    }
  }
  if (threadIdx.x == 0) {
    f.defined = b.defined; f.depth = b.depth; f.n_events = nev;
    for (int a = 0; a < 3; a++) { f.bmin[a] = b.mn[a]; f.bmax[a] = b.mx[a]; }
    if (fail || nev > CCV2_MAX_EVENTS) f.error |= FERR_DEPTH;
    if (mode == 1) { f.rekey = 1; f.n_finite = f.n; f.violator = NONE_U32; }
  }
}

// ------------------------------------------------------------------------------------------------
// Morton keys.  One thread per point: 16-byte load of (x,y,z,_), FP64 key per axis exactly as
// genOctreeKeyforPoint: (unsigned)((double(p) - min) / resolution).  Also verifies that every point
// beyond the exact prefix fits the box (atomicMin of the first violator) and counts finite points.
// rekey_only: second launch, acts only on frames the slow bbox path touched.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keygen_kernel(EncFrame *frames, EncParams P, int rekey_only) {
  EncFrame &f = frames[blockIdx.y];
  if (rekey_only && !f.rekey) return;
  const uint32_t n = f.n;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  if (!f.defined || (f.error & FERR_DEPTH)) return;     // no finite point in the prefix: handled by the slow path
  const uint32_t depth = f.depth, ne = f.n_events;
  // Box log.  Points added after the last change (all but a handful) see the final box; an earlier point is keyed
  // against the box of its own time and then re-rooted: + 1 << depth_before on every axis whose minimum moved later.
  // For power-of-two resolutions both give the same integers; for any other resolution only this order does.
  __shared__ BoxEvent s_ev[CCV2_MAX_EVENTS];
  const uint32_t last_change = f.ev[ne - 1].idx;
  const bool old_box = blockIdx.x * blockDim.x <= last_change;           // block-uniform
  if (old_box) {
    for (uint32_t k = threadIdx.x; k < ne * (sizeof(BoxEvent) / 8); k += blockDim.x) ((uint64_t *)s_ev)[k] = ((const uint64_t *)f.ev)[k];
    __syncthreads();
  }
  // the sequential walk covered [0, prefix) in the first launch and everything in the slow path; only points beyond it
  // have to be checked against the box (a point inside the walked range fits the box of its own time by construction)
  const uint32_t walked = rekey_only ? n : min(n, (uint32_t)P.prefix_len);
  // Packed sort element: the colour sum of a voxel does not depend on the order of its points, so without centroids (and
  // outside detail mode) nothing needs the point index after the sort.  A Morton code of depth <= 13 has 39 bits (+1 for
  // the non-finite sentinel): code << 24 | b,g,r is ONE 64-bit word per point -- 16 instead of 24 bytes per point and pass
  // through the sort, and leaf_emit reads its colours sequentially instead of gathering them from the cloud.
  const bool packed = P.allow_packed && depth <= 13;
  bool fin = false, viol = false;
  uint64_t key = 1ull << (3 * depth);                    // sorts after every valid code
  if (i < n) {
    float4 q = __ldg((const float4 *)(f.pts + 32ull * i));
    fin = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
    if (fin) {
      double p[3] = { (double)q.x, (double)q.y, (double)q.z };
      uint32_t k[3];
      if (!old_box || i >= last_change) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (i >= walked) viol |= (p[a] < f.bmin[a]) | (p[a] >= f.bmax[a]);
          double t = __dsub_rn(p[a], f.bmin[a]);
          t = P.res_pow2 ? __dmul_rn(t, P.inv_res) : __ddiv_rn(t, P.res);
          k[a] = viol ? 0u : __double2uint_rz(t);
        }
      } else {
        int e = (int)ne - 1;
        while (e > 0 && s_ev[e].idx > i) e--;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          double t = __dsub_rn(p[a], s_ev[e].mn[a]);
          t = P.res_pow2 ? __dmul_rn(t, P.inv_res) : __ddiv_rn(t, P.res);
          uint32_t kk = __double2uint_rz(t);
          for (int j = e + 1; j < (int)ne; j++) kk += ((s_ev[j].mask >> a) & 1u) << s_ev[j].depth_before;
          k[a] = kk;
        }
      }
      key = morton_xyz(k[0], k[1], k[2]);
    }
    if (packed) f.keys[0][i] = (key << 24) | (__ldg((const uint32_t *)(f.pts + 32ull * i + 16)) & 0xFFFFFFu);   // same 32-byte sector as x,y,z
    else { f.keys[0][i] = key; f.vals[0][i] = i; }
  }
  if (i == 0) f.packed = packed;
  // n_finite starts at n (set by the host / the slow bbox path) and only non-finite points touch it: no atomics at
  // all for the usual all-finite cloud (one same-address atomic per warp cost ~20 us per 1M-point frame)
  const uint32_t bad = __popc(__ballot_sync(FULL_MASK, i < n && !fin));
  if (lane_id() == 0 && bad) atomicSub(&f.n_finite, bad);
  if (viol) atomicMin(&f.violator, i);
}

// If the prefix held no finite point at all the box is still undefined: route the frame through the slow path.
__global__ void bbox_fixup_kernel(EncFrame *frames, EncParams P) {
  EncFrame &f = frames[blockIdx.x];
  if (threadIdx.x == 0 && !f.defined && f.n > (uint32_t)P.prefix_len && f.violator == NONE_U32) f.violator = P.prefix_len;
}

__device__ __forceinline__ uint32_t frame_sort_bits(const EncFrame &f) { return 3 * f.depth + (f.n_finite < f.n ? 1u : 0u); }

// Assigns frame ids (impl.hpp:133: pre-increment for every non-empty frame) and the sort pass count.
__global__ void frame_setup_kernel(EncFrame *frames, int nframes, uint32_t *frame_counter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t id = *frame_counter;
  for (int k = 0; k < nframes; k++) {
    EncFrame &f = frames[k];
    if ((f.error & FERR_DEPTH) || !f.defined) { f.n_finite = 0; }   // no finite point at all: empty frame
    if (f.n_finite > 0) { if (f.frame_id_fixed) f.frame_id = f.frame_id_fixed; else { id++; f.frame_id = id; } f.npasses = (frame_sort_bits(f) + 7) / 8; }
    else { f.npasses = 0; f.V = 0; f.B = 0; }
  }
  *frame_counter = id;
}

// ------------------------------------------------------------------------------------------------
// Radix sort of (code, index) pairs, 8-bit digits, LSD, stable: one global-histogram kernel for all
// passes, then one kernel per pass with per-digit decoupled look-back between tiles (tickets give the
// tile order).  Pass p reads buffer p&1 and writes buffer (p+1)&1; the result is in buffer npasses&1.
// ------------------------------------------------------------------------------------------------
#define SORT_THREADS 256
#define SORT_ITEMS 16
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

__global__ void __launch_bounds__(256) sort_hist_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t n = f.n, np = f.npasses;
  uint32_t base = blockIdx.x * SORT_TILE;
  if (base >= n || np == 0) return;
  __shared__ uint32_t sh[8][256];
  for (uint32_t k = threadIdx.x; k < 8 * 256; k += blockDim.x) (&sh[0][0])[k] = 0;
  __syncthreads();
  const uint64_t *keys = f.keys[0];
  const uint32_t kshift = f.packed ? 24 : 0;
  for (uint32_t k = 0; k < SORT_ITEMS; k++) {
    uint32_t i = base + k * SORT_THREADS + threadIdx.x;
    if (i < n) {
      uint64_t key = keys[i] >> kshift;
      for (uint32_t p = 0; p < np; p++) atomicAdd(&sh[p][(key >> (8 * p)) & 255], 1u);
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < np * 256; k += blockDim.x) {
    uint32_t v = (&sh[0][0])[k];
    if (v) atomicAdd(&f.ghist[k], v);
  }
}

#define SORT_SMEM_BYTES (SORT_TILE * 12)     // dynamic shared memory: the tile's keys (8 B) and values (4 B) in digit order
__global__ void __launch_bounds__(SORT_THREADS) sort_pass_kernel(EncFrame *frames, int pass) {
  EncFrame &f = frames[blockIdx.y];
  if ((uint32_t)pass >= f.npasses) return;
  const uint32_t n = f.n;
  const uint32_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
  if (blockIdx.x >= ntiles) return;
  extern __shared__ __align__(16) uint8_t sort_dyn[];
  uint64_t *skey = (uint64_t *)sort_dyn;
  uint32_t *sval = (uint32_t *)(sort_dyn + SORT_TILE * 8);
  __shared__ uint32_t s_tile;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t whist[SORT_THREADS / 32][256];
  __shared__ uint32_t tile_off[256], dbase[256];
  __shared__ uint64_t s_scan[33];
  const uint64_t *skeys = f.keys[pass & 1]; const uint32_t *svals = f.vals[pass & 1];
  const bool packed = f.packed != 0;
  if (threadIdx.x == 0) {
    // The tile -- 4096 keys (32 KB) and 4096 indices (16 KB), both contiguous -- comes in as two TMA bulk copies issued
    // by this one thread as soon as the ticket is known; everybody else meets the data at the mbarrier.  Byte counts are
    // rounded up to 16 (the buffers carry 8 elements of slack).
    const uint32_t t = atomicAdd(&f.ticket[TK_SORT0 + pass], 1u);
    s_tile = t;
    const uint32_t cnt = min((uint32_t)SORT_TILE, n - t * SORT_TILE);
    const uint32_t kb = ((cnt * 8u) + 15u) & ~15u, vb = ((cnt * 4u) + 15u) & ~15u;
    mbar_init(&s_bar, 1);
    mbar_expect_tx(&s_bar, kb + (packed ? 0u : vb));
    bulk_g2s(skey, skeys + (size_t)t * SORT_TILE, kb, &s_bar);
    if (!packed) bulk_g2s(sval, svals + (size_t)t * SORT_TILE, vb, &s_bar);
  }
  for (uint32_t k = threadIdx.x; k < (SORT_THREADS / 32) * 256; k += blockDim.x) (&whist[0][0])[k] = 0;
  // global digit base: exclusive scan of the pass histogram (256 threads, one digit each)
  uint64_t tot;
  uint32_t gbase = (uint32_t)block_excl_scan_u64(f.ghist[pass * 256 + threadIdx.x], &tot, s_scan);
  const uint32_t tile = s_tile;
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  uint64_t *dkeys = f.keys[(pass + 1) & 1]; uint32_t *dvals = f.vals[(pass + 1) & 1];
  const uint32_t shift = 8 * pass + (packed ? 24 : 0);
  const uint32_t wbase = tile * SORT_TILE + w * (32 * SORT_ITEMS);
  uint64_t key[SORT_ITEMS]; uint32_t val[SORT_ITEMS]; uint16_t rank[SORT_ITEMS];
  mbar_wait(&s_bar, 0);
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) { const uint32_t li = w * (32 * SORT_ITEMS) + k * 32 + lane; const bool in = wbase + k * 32 + lane < n; key[k] = in ? skey[li] : ~0ull; val[k] = (in && !packed) ? sval[li] : 0u; }
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) {
    uint32_t i = wbase + k * 32 + lane;
    bool valid = i < n;
    uint32_t d = valid ? (uint32_t)((key[k] >> shift) & 255) : 256u;
    uint32_t peers = __match_any_sync(FULL_MASK, d);
    uint32_t leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && valid) { old = whist[w][d]; whist[w][d] = old + __popc(peers); }
    old = __shfl_sync(FULL_MASK, old, leader);
    rank[k] = (uint16_t)(old + __popc(peers & lanemask_lt()));
    __syncwarp();
  }
  __syncthreads();
  {
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_THREADS / 32; ww++) { uint32_t c = whist[ww][d]; whist[ww][d] = run; run += c; }
    uint32_t *st = f.sort_status + ((size_t)pass * f.tiles_max) * 256 + d;
    uint32_t excl = 0;
    if (tile == 0) {
      *(volatile uint32_t *)&st[0] = (2u << 30) | run;
    } else {
      *(volatile uint32_t *)&st[(size_t)tile * 256] = (1u << 30) | run;
      for (int t = (int)tile - 1; t >= 0; t--) {
        uint32_t s;
        do { s = ld_volatile_u32(&st[(size_t)t * 256]); } while ((s >> 30) == 0);
        excl += s & 0x3FFFFFFFu;
        if ((s >> 30) == 2) break;
      }
      *(volatile uint32_t *)&st[(size_t)tile * 256] = (2u << 30) | (excl + run);
    }
    tile_off[d] = gbase + excl;
    dbase[d] = (uint32_t)block_excl_scan_u64(run, &tot, s_scan);     // where digit d starts inside this tile
  }
  __syncthreads();
  // stage the tile in digit order in shared memory, then write it out with consecutive threads on consecutive
  // addresses inside every digit run (direct scatter wrote one 32-byte sector per 8-byte key)
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) {
    uint32_t i = wbase + k * 32 + lane;
    if (i < n) {
      uint32_t d = (uint32_t)((key[k] >> shift) & 255);
      uint32_t lp = dbase[d] + whist[w][d] + rank[k];
      skey[lp] = key[k]; if (!packed) sval[lp] = val[k];
    }
  }
  __syncthreads();
  const uint32_t tile_n = min((uint32_t)SORT_TILE, n - tile * SORT_TILE);
  for (uint32_t idx = threadIdx.x; idx < tile_n; idx += SORT_THREADS) {
    const uint64_t kk = skey[idx];
    const uint32_t d = (uint32_t)((kk >> shift) & 255);
    const uint32_t pos = tile_off[d] + (idx - dbase[d]);
    dkeys[pos] = kk; if (!packed) dvals[pos] = sval[idx];
  }
}

// ------------------------------------------------------------------------------------------------
// Leaf segmentation over the sorted codes: head flags, leaf ids, and for every leaf the number of
// branch nodes it newly opens -> byte offsets in DFS order.  One chained scan of
// (heads << 36 | opened) pairs.  Each tile also zeroes its slice of the tree byte buffer.
// ------------------------------------------------------------------------------------------------
#define LEAF_THREADS 256
#define LEAF_ITEMS 4
#define LEAF_TILE (LEAF_THREADS * LEAF_ITEMS)

__global__ void __launch_bounds__(LEAF_THREADS) leaf_scan_kernel(EncFrame *frames, int snake) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t nf = f.n_finite;
  const uint32_t ntiles = (nf + LEAF_TILE - 1) / LEAF_TILE;
  if (blockIdx.x >= ntiles || f.npasses == 0) return;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_LEAF], 1u);
  __syncthreads();
  const uint32_t tile = s_tile, d = f.depth;
  const uint64_t *keys = f.keys[f.npasses & 1];
  const uint32_t kshift = f.packed ? 24 : 0;               // packed elements carry the colour in their low 24 bits
  const uint32_t i0 = tile * LEAF_TILE + threadIdx.x * LEAF_ITEMS;
  uint64_t k[LEAF_ITEMS], prev = 0;
  if (i0 > 0 && i0 < nf) prev = keys[i0 - 1] >> kshift;
  uint64_t v[LEAF_ITEMS], sum = 0; uint8_t fn[LEAF_ITEMS];
#pragma unroll
  for (int j = 0; j < LEAF_ITEMS; j++) {
    uint32_t i = i0 + j;
    v[j] = 0; fn[j] = 0;
    if (i < nf) {
      k[j] = keys[i] >> kshift;
      if (i == 0) { v[j] = (1ull << 36) | d; fn[j] = 0; }
      else if (k[j] != prev) {
        uint32_t msb = 63 - __clzll((long long)(k[j] ^ prev));
        uint32_t first_new = d - msb / 3;
        fn[j] = (uint8_t)first_new;
        v[j] = (1ull << 36) | (uint64_t)(d - first_new);
      }
      prev = k[j];
    }
    sum += v[j];
  }
  uint64_t tile_total;
  uint64_t excl = block_excl_scan_u64(sum, &tile_total, s_scan);
  if (threadIdx.x < 32) {
    uint64_t e = scan_lookback(f.scan_status, tile, tile_total);
    if (threadIdx.x == 0) s_excl = e;
  }
  __syncthreads();
  const uint64_t tile_excl = s_excl;
  excl += tile_excl;
#pragma unroll
  for (int j = 0; j < LEAF_ITEMS; j++) {
    uint32_t i = i0 + j;
    if (i < nf) {
      if (v[j]) {
        uint32_t id = (uint32_t)(excl >> 36), off = (uint32_t)(excl & 0xFFFFFFFFFull);
        f.leaf_key[id] = k[j]; f.leaf_start[id] = i; f.leaf_off[id] = off; f.first_new[id] = fn[j];
      }
      excl += v[j];
      if (i == nf - 1) {
        uint32_t V = (uint32_t)(excl >> 36); uint64_t B = excl & 0xFFFFFFFFFull;
        f.V = V; f.leaf_start[V] = nf; f.leaf_off[V] = (uint32_t)B;
        if (f.tree && B > f.tree_cap) { f.error |= FERR_TREE_CAP; B = 0; f.V = 0; }   // f.tree == nullptr: a grid without occupancy bytes (inter_kernels.cuh)
        f.B = (uint32_t)B;
        f.img_h = V / 256 + 1;                         // cjpeg.h:197-198
        f.mcu_h = (f.img_h + 15) / 16;
        // libjpeg refuses images higher than JPEG_MAX_DIMENSION (65500): the reference's jpeg_io fails there, a 16-bit
        // SOF0 height would silently wrap here.  V >= 16.7 M voxels with SNAKE colour is reported, not encoded.
        if (snake && f.img_h > 65500u) { f.error |= FERR_UNSUPPORTED; f.V = 0; f.B = 0; }
      }
    }
  }
  // zero this tile's slice of the tree bytes (the occupancy kernel ORs into it)
  if (!f.tree) return;
  uint64_t b0 = tile_excl & 0xFFFFFFFFFull, b1 = b0 + (tile_total & 0xFFFFFFFFFull);
  if (b1 > f.tree_cap) b1 = f.tree_cap;
  // word-granular zeroing with byte edges
  uint64_t w0 = (b0 + 3) & ~3ull, w1 = b1 & ~3ull;
  if (w0 >= w1) { for (uint64_t p = b0 + threadIdx.x; p < b1; p += blockDim.x) f.tree[p] = 0; }
  else {
    for (uint64_t p = b0 + threadIdx.x; p < w0; p += blockDim.x) f.tree[p] = 0;
    for (uint64_t p = w0 / 4 + threadIdx.x; p < w1 / 4; p += blockDim.x) ((uint32_t *)f.tree)[p] = 0;
    for (uint64_t p = w1 + threadIdx.x; p < b1; p += blockDim.x) f.tree[p] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Per leaf: average colour ([PCL] ColorCoding::encodeAverageOfPoints, impl.hpp:1548-1550), centroid
// residual (pcv2.h:83-97) and the leaf's contributions to the occupancy bytes.
// Leaf j contributes one child bit to every branch it opened (levels first_new..d-1: its own chain at
// leaf_off[j]) and one bit to the already-open parent at level first_new-1, whose byte lives in the
// chain of that parent's first leaf (found by a galloping lower_bound over the sorted leaf codes).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tree_or(uint8_t *tree, uint32_t addr, uint32_t bit) {
  atomicOr((uint32_t *)(tree + (addr & ~3u)), (1u << bit) << (8 * (addr & 3u)));
}

__global__ void __launch_bounds__(256) leaf_emit_kernel(EncFrame *frames, EncParams P) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= V) return;
  const uint32_t d = f.depth;
  const uint64_t key = f.leaf_key[j];
  const uint32_t fn = f.first_new[j], off = f.leaf_off[j];
  // own chain
  for (uint32_t l = fn; l < d; l++) tree_or(f.tree, off + (l - fn), (uint32_t)(key >> (3 * (d - 1 - l))) & 7u);
  if (j > 0) {
    const uint32_t l = fn - 1;                           // parent level (fn >= 1 for j > 0)
    const uint32_t sh = 3 * (d - l);
    const uint64_t target = sh >= 64 ? 0 : (key >> sh) << sh;   // first possible code under the parent
    // gallop backwards to bracket the parent's first leaf, then binary search
    uint32_t hi = j, step = 1, lo;
    for (;;) {
      if (step >= hi) { lo = 0; break; }
      uint32_t probe = hi - step;
      if (f.leaf_key[probe] < target) { lo = probe + 1; break; }
      hi = probe; step <<= 1;
    }
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (f.leaf_key[mid] < target) lo = mid + 1; else hi = mid; }
    const uint32_t pi = lo;                              // parent's first leaf (opened the parent)
    tree_or(f.tree, f.leaf_off[pi] + (l - f.first_new[pi]), (uint32_t)(key >> (3 * (d - 1 - l))) & 7u);
  }
  const uint32_t s0 = f.leaf_start[j], s1 = f.leaf_start[j + 1];
  const uint32_t *vals = f.vals[f.npasses & 1];
  if (P.do_color) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    if (f.packed) {                                        // colours rode through the sort: sequential reads
      const uint64_t *sk = f.keys[f.npasses & 1];
      for (uint32_t k = s0; k < s1; k++) { const uint32_t c = (uint32_t)__ldg(&sk[k]); c0 += c & 0xFF; c1 += (c >> 8) & 0xFF; c2 += (c >> 16) & 0xFF; }
    } else
    for (uint32_t k = s0; k < s1; k++) {
      uint32_t c = __ldg((const uint32_t *)(f.pts + 32ull * vals[k] + 16));
      c0 += c & 0xFF; c1 += (c >> 8) & 0xFF; c2 += (c >> 16) & 0xFF;
    }
    uint32_t len = s1 - s0;
    if (len > 1) { c0 /= len; c1 /= len; c2 /= len; }
    c0 >>= P.color_reduction; c1 >>= P.color_reduction; c2 >>= P.color_reduction;
    uint8_t *o = f.avg + 3ull * j;
    o[0] = (uint8_t)c0; o[1] = (uint8_t)c1; o[2] = (uint8_t)c2;
  }
  if (P.do_centroid && !P.detail) {                      // detail mode codes per-point residuals instead (detail_emit_kernel)
    float ax = 0.f, ay = 0.f, az = 0.f;                  // pcl::compute3DCentroid: float accumulation in index order
    for (uint32_t k = s0; k < s1; k++) {
      float4 q = __ldg((const float4 *)(f.pts + 32ull * vals[k]));
      ax = __fadd_rn(ax, q.x); ay = __fadd_rn(ay, q.y); az = __fadd_rn(az, q.z);
    }
    float cnt = (float)(s1 - s0);
    float c[3] = { __fdiv_rn(ax, cnt), __fdiv_rn(ay, cnt), __fdiv_rn(az, cnt) };
    uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
    for (int a = 0; a < 3; a++) {
      double corner = __dadd_rn(__dmul_rn((double)k3[a], P.res), f.bmin[a]);
      double q = __ddiv_rn(__dsub_rn((double)c[a], corner), (double)0.001f);
      int qi = (int)q;                                   // C truncation
      qi = max(-127, min(127, qi));
      f.cen[3ull * j + a] = (uint8_t)qi;
    }
  }
  if (j == V - 1) {
    if (P.do_centroid) f.ncen = P.detail ? 0 : 3 * V;   // detail mode: the (empty) centroid vector is still written (impl.hpp:1701-1707)
    if (P.do_color && P.color_type != 1 && P.color_type != 2) f.ncolor = 3 * V;   // raw averages are the payload
  }
}

// ------------------------------------------------------------------------------------------------
// The encoder's simplified cloud output_ (impl.hpp:1549-1576; [PCL] getOutputCloud, eval.hpp:862), produced on demand
// from the leaf arrays of the last encode: one thread per leaf.  grid (ceil(V / 256))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) output_cloud_kernel(EncFrame f, EncParams P, uint8_t *out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f.V) return;
  const uint64_t key = f.leaf_key[j];
  const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
  float xyz[3];
  if (P.do_centroid) {                                   // pcl::compute3DCentroid: float accumulation in index order (impl.hpp:1565-1571)
    const uint32_t s0 = f.leaf_start[j], s1 = f.leaf_start[j + 1];
    const uint32_t *vals = f.vals[f.npasses & 1];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (uint32_t k = s0; k < s1; k++) {
      float4 q = __ldg((const float4 *)(f.pts + 32ull * vals[k]));
      ax = __fadd_rn(ax, q.x); ay = __fadd_rn(ay, q.y); az = __fadd_rn(az, q.z);
    }
    const float cnt = (float)(s1 - s0);
    xyz[0] = __fdiv_rn(ax, cnt); xyz[1] = __fdiv_rn(ay, cnt); xyz[2] = __fdiv_rn(az, cnt);
  } else {
#pragma unroll
    for (int a = 0; a < 3; a++) {                        // impl.hpp:1518-1520, 1559-1563
      const double corner = __dadd_rn(__dmul_rn((double)k3[a], P.res), f.bmin[a]);
      xyz[a] = (float)__dadd_rn(corner, __dmul_rn(0.5, P.res));
    }
  }
  uint32_t rgba = 0xFF000000u;                           // PointXYZRGB(): r = g = b = 0, a = 255
  if (P.do_color) { const uint8_t *c = f.avg + 3ull * j; rgba |= (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16); }
  uint4 *o = (uint4 *)(out + 32ull * j);
  o[0] = make_uint4(__float_as_uint(xyz[0]), __float_as_uint(xyz[1]), __float_as_uint(xyz[2]), 0x3F800000u);
  o[1] = make_uint4(rgba, 0, 0, 0);
}

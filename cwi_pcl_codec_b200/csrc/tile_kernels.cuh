// tile_kernels.cuh -- root-octant tile mode (BASELINE configs[3]): one large frame is cut into 8 (or 64)
// spatial tiles of the unit cube, every tile keeps its points IN THEIR ORIGINAL RELATIVE ORDER and is then an ordinary
// frame of the codec (own bounding box, own header and tables), so each tile stream is bit-exact against the reference
// encoder run on that subset -- and the tiles' serial entropy stages, the frame's latency, run side by side (on one GPU
// or round-robin over the ranks of a node).  The reference has no such mode (it has no parallelism at all,
// CMakeLists.txt:85-87): the tile is defined here, on the normalised coordinates evaluate_compression feeds the codec
// (impl.hpp:1915-1945 maps every cloud into [0,1]^3): tile = Morton index (x most significant, like an octree child)
// of floor(p * 2^k) per axis, clamped, k = tile_bits / 3.
//
// A stable counting sort by tile id: per-block bucket counts -> one scan over (tile, block) -> scatter with ranks that
// preserve the input order inside every tile.
#pragma once
#include "common.cuh"

#define TILE_BLOCK 1024
#define TILE_MAX 64

__device__ __forceinline__ uint32_t tile_of(const float4 &p, int k) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return 0;          // dropped by the encoder anyway (impl.hpp:99: isFinite)
  const float s = (float)(1 << k); const int hi = (1 << k) - 1;
  const int tx = min(hi, max(0, (int)floorf(p.x * s))), ty = min(hi, max(0, (int)floorf(p.y * s))), tz = min(hi, max(0, (int)floorf(p.z * s)));
  uint32_t t = 0;
  for (int b = k - 1; b >= 0; b--) t = (t << 3) | (((tx >> b) & 1) << 2) | (((ty >> b) & 1) << 1) | ((tz >> b) & 1);
  return t;
}

// counts[tile * nblocks + block]
__global__ void __launch_bounds__(TILE_BLOCK) tile_count_kernel(const uint8_t *pts, uint32_t n, int k, uint32_t *counts, uint32_t nblocks) {
  __shared__ uint32_t s_cnt[TILE_MAX];
  if (threadIdx.x < TILE_MAX) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * TILE_BLOCK + threadIdx.x;
  const uint32_t t = i < n ? tile_of(__ldg((const float4 *)(pts + 32ull * i)), k) : 0xFFFFu;   // every lane takes part in the match
  const uint32_t peers = __match_any_sync(FULL_MASK, t);
  if (i < n && lane_id() == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_cnt[t], __popc(peers));
  __syncthreads();
  const uint32_t nt = 1u << (3 * k);
  if (threadIdx.x < nt) counts[threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of counts in (tile-major, block-minor) order, in place; offsets[t] = first record of tile t, offsets[nt] = n.  One CTA.
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t *counts, uint32_t nblocks, uint32_t nt, uint64_t *offsets) {
  __shared__ uint64_t s_scan[33];
  __shared__ uint64_t s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  const uint32_t total = nt * nblocks;
  for (uint32_t base = 0; base < total; base += 1024) {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < total ? counts[idx] : 0;
    uint64_t tot;
    const uint64_t ex = block_excl_scan_u64(v, &tot, s_scan) + s_run;
    if (idx < total) { counts[idx] = (uint32_t)ex; if (idx % nblocks == 0) offsets[idx / nblocks] = ex; }
    __syncthreads();
    if (threadIdx.x == 0) s_run += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[nt] = s_run;
}

__global__ void __launch_bounds__(TILE_BLOCK) tile_scatter_kernel(const uint8_t *pts, uint32_t n, int k, const uint32_t *counts, uint32_t nblocks, uint8_t *out) {
  __shared__ uint32_t s_warp[TILE_BLOCK / 32][TILE_MAX];
  for (uint32_t q = threadIdx.x; q < (TILE_BLOCK / 32) * TILE_MAX; q += TILE_BLOCK) (&s_warp[0][0])[q] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * TILE_BLOCK + threadIdx.x, w = threadIdx.x >> 5, lane = lane_id();
  uint32_t t = 0xFFFFu; uint4 a = make_uint4(0, 0, 0, 0), b = a;
  const bool live = i < n;
  if (live) {
    a = __ldg((const uint4 *)(pts + 32ull * i)); b = __ldg((const uint4 *)(pts + 32ull * i + 16));
    t = tile_of(make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), 0.f), k);
  }
  const uint32_t peers = __match_any_sync(FULL_MASK, t);                      // every lane takes part in the match
  const uint32_t rank = __popc(peers & lanemask_lt());
  if (live && lane == (uint32_t)(__ffs(peers) - 1)) s_warp[w][t] = __popc(peers);
  __syncthreads();
  if (threadIdx.x < TILE_MAX) {                                               // per tile: exclusive prefix over the block's warps
    uint32_t run = 0;
    for (int ww = 0; ww < TILE_BLOCK / 32; ww++) { const uint32_t c = s_warp[ww][threadIdx.x]; s_warp[ww][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  if (live) {
    const uint64_t pos = (uint64_t)counts[t * nblocks + blockIdx.x] + s_warp[w][t] + rank;
    uint4 *o = (uint4 *)(out + 32ull * pos);
    o[0] = a; o[1] = b;
  }
}

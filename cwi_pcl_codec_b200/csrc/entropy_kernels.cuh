// entropy_kernels.cuh -- [PCL] StaticRangeCoder (char vectors) on the GPU + frame assembly.
// Reference call sites: impl.hpp:1682-1760 (entropyEncoding), :1766-1835 (entropyDecoding),
// header impl.hpp:1472-1502.  Arithmetic: 32-bit carry-less range coder, top = 1<<24, bottom = 1<<16,
// cumulative table of 257 u32 written raw, totals rescaled below 1<<16, 4 flush bytes (see oracle/ccv2_oracle.c
// rc_encode_to / orc_range_decode and DESIGN.md "range coder word size").
//
// The coder is serial per vector by construction (state = low/range, no stored lengths), so the unit of
// parallelism is the stream: ONE WARP PER STREAM, executed warp-uniformly -- every lane carries the same
// (low, range) state, lanes differ only in which input symbols / output bytes they stage, so symbol loads and
// byte stores are coalesced 32-wide and the serial recurrence never diverges.
#pragma once
#include "common.cuh"

#define RC_TOP (1u << 24)
#define RC_BOTTOM (1u << 16)
#define FRAME_HDR_BYTES 140

// ---- 256-bin histograms of the three layers of each frame: grid (blocks, 3, frames)
__device__ __forceinline__ void enc_layer(const EncFrame &f, int which, const uint8_t *&src, uint32_t &n) {
  if (which == 0) { src = f.tree; n = f.B; }
  else if (which == 1) { src = f.cen; n = f.ncen; }
  else { src = f.cpay; n = f.ncolor; }
}
__global__ void __launch_bounds__(256) hist_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.z];
  if (f.V == 0) return;
  const uint8_t *src; uint32_t n;
  enc_layer(f, blockIdx.y, src, n);
  const uint32_t per_block = 256 * 64;
  uint32_t b0 = blockIdx.x * per_block;
  if (b0 >= n) return;
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  uint32_t b1 = min(n, b0 + per_block);
  // 16-byte loads over the aligned middle, bytes at the edges (src is 16-byte aligned at offset 0)
  for (uint32_t p = b0 + threadIdx.x * 16; p < b1; p += 256 * 16) {
    if (p + 16 <= b1) {
      uint4 v = *(const uint4 *)(src + p);
      uint32_t wv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int k = 0; k < 4; k++) { atomicAdd(&sh[wv[k] & 255], 1u); atomicAdd(&sh[(wv[k] >> 8) & 255], 1u); atomicAdd(&sh[(wv[k] >> 16) & 255], 1u); atomicAdd(&sh[wv[k] >> 24], 1u); }
    } else for (uint32_t q = p; q < b1; q++) atomicAdd(&sh[src[q]], 1u);
  }
  __syncthreads();
  uint32_t v = sh[threadIdx.x];
  if (v) atomicAdd(&f.hist[blockIdx.y * 256 + threadIdx.x], v);
}

// cumulative table with PCL's "+1 if empty" rule and the halving rescale; executed by lane 0
__device__ inline void rc_build_table(const uint32_t *hist, uint32_t *freq /* smem[257] */) {
  freq[0] = 0;
  for (int s = 1; s <= 256; s++) { uint32_t v = freq[s - 1] + hist[s - 1]; if (v <= freq[s - 1]) v = freq[s - 1] + 1; freq[s] = v; }
  while (freq[256] >= RC_BOTTOM) {
    for (int s = 1; s <= 256; s++) { uint32_t v = freq[s] >> 1; if (v <= freq[s - 1]) v = freq[s - 1] + 1; freq[s] = v; }
  }
}

// ---- range encoder: grid (3 layers, frames), one warp each
__global__ void __launch_bounds__(32) rc_encode_kernel(EncFrame *frames, int do_centroid, int do_color) {
  EncFrame &f = frames[blockIdx.y];
  if (f.V == 0) return;
  const int which = blockIdx.x;
  if ((which == 1 && !do_centroid) || (which == 2 && !do_color)) return;
  const uint8_t *src; uint32_t n;
  enc_layer(f, which, src, n);
  uint8_t *dst; uint64_t cap;
  if (which == 0) { dst = f.stream + FRAME_HDR_BYTES + 8; cap = f.stream_cap > FRAME_HDR_BYTES + 8 ? f.stream_cap - FRAME_HDR_BYTES - 8 : 0; }
  else { dst = f.rc_tmp[which - 1]; cap = f.rc_tmp_cap[which - 1]; }
  __shared__ uint32_t freq[257];
  __shared__ uint32_t packed[256];
  const uint32_t lane = lane_id();
  if (lane == 0) rc_build_table(f.hist + which * 256, freq);
  __syncwarp();
  if (cap < 1028 + 8) { if (lane == 0) atomicOr(&f.error, FERR_STREAM_CAP); return; }
  for (uint32_t s = lane; s < 257; s += 32) ((uint32_t *)dst)[s] = freq[s];       // dst is 4-byte aligned
  for (uint32_t s = lane; s < 256; s += 32) packed[s] = (freq[s] << 16) | (freq[s + 1] - freq[s]);
  __syncwarp();
  const FastDiv fd = fastdiv_make(freq[256]);
  uint8_t *out = dst + 1028;
  const uint64_t out_cap = cap - 1028;
  uint32_t low = 0, range = 0xFFFFFFFFu;
  uint64_t cnt = 0; uint32_t mybyte = 0;
  bool overflow = false;
  for (uint32_t base = 0; base < n; base += 32) {
    uint32_t i = base + lane;
    uint32_t pk = i < n ? packed[src[i]] : 0;
    uint32_t m = min(32u, n - base);
    if (cnt + 4 * 32 + 8 > out_cap) { overflow = true; break; }     // a symbol emits at most 4 bytes
    for (uint32_t k = 0; k < m; k++) {
      uint32_t p = __shfl_sync(FULL_MASK, pk, k);
      uint32_t r = fastdiv(range, fd);
      low += (p >> 16) * r;
      range = r * (p & 0xFFFFu);
      for (;;) {
        if ((low ^ (low + range)) >= RC_TOP) {
          if (range >= RC_BOTTOM) break;
          range = (0u - low) & (RC_BOTTOM - 1);
        }
        if (lane == (uint32_t)(cnt & 31)) mybyte = low >> 24;
        cnt++;
        if ((cnt & 31) == 0) out[cnt - 32 + lane] = (uint8_t)mybyte;
        range <<= 8; low <<= 8;
      }
    }
  }
  if (!overflow) {
    for (int k = 0; k < 4; k++) {                        // flush
      if (lane == (uint32_t)(cnt & 31)) mybyte = low >> 24;
      cnt++;
      if ((cnt & 31) == 0) out[cnt - 32 + lane] = (uint8_t)mybyte;
      low <<= 8;
    }
    if (lane < (cnt & 31)) out[(cnt & ~31ull) + lane] = (uint8_t)mybyte;
  }
  if (lane == 0) {
    if (overflow) atomicOr(&f.error, FERR_STREAM_CAP);
    f.rc_len[which] = (uint32_t)(1028 + cnt);
  }
}

// ---- frame assembly: header + size words + layers: grid (blocks, frames)
struct HeaderParams { double octree_res, point_res; uint8_t do_voxel_grid, with_color, color_bits, do_centroid, connectivity, scalable, icp_offset, _p; uint32_t color_type; int32_t macroblock; };

__global__ void __launch_bounds__(256) assemble_kernel(EncFrame *frames, HeaderParams H) {
  EncFrame &f = frames[blockIdx.y];
  if (f.V == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) f.out_len = 0; return; }
  uint8_t *s = f.stream;
  const uint64_t l0 = f.rc_len[0], l1 = H.do_centroid ? f.rc_len[1] : 0, l2 = H.with_color ? f.rc_len[2] : 0;
  const uint64_t off_cen = FRAME_HDR_BYTES + 8 + l0;
  const uint64_t off_col = off_cen + (H.do_centroid ? 4 + l1 : 0);
  const uint64_t total = off_col + (H.with_color ? 8 + l2 : 0);
  if (total > f.stream_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&f.error, FERR_STREAM_CAP); f.out_len = 0; } return; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const char id2[] = "<PCL-OCT-CODECV2-COMPRESSED>", id1[] = "<PCL-OCT-COMPRESSED>";
    for (int k = 0; k < 28; k++) s[k] = id2[k];
    for (int k = 0; k < 20; k++) s[28 + k] = id1[k];
    auto put = [&](uint64_t off, const void *p, int nb) { const uint8_t *q = (const uint8_t *)p; for (int k = 0; k < nb; k++) s[off + k] = q[k]; };
    uint32_t fid = f.frame_id; put(48, &fid, 4);
    s[52] = 1; s[53] = H.do_voxel_grid; s[54] = H.with_color;
    uint64_t pc = f.V; put(55, &pc, 8);
    put(63, &H.octree_res, 8); s[71] = H.color_bits; put(72, &H.point_res, 8);
    put(80, f.bmin, 24); put(104, f.bmax, 24);
    s[128] = H.do_centroid; s[129] = H.connectivity; s[130] = H.scalable;
    put(131, &H.color_type, 4); put(135, &H.macroblock, 4); s[139] = H.icp_offset;
    uint64_t B = f.B; put(140, &B, 8);
    if (H.do_centroid) { uint32_t c = f.ncen; put(off_cen, &c, 4); }
    if (H.with_color) { uint64_t c = f.ncolor; put(off_col, &c, 8); }
    f.out_len = total;
    f.coded[0] = l0; f.coded[1] = l1; f.coded[2] = l2;
  }
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (uint64_t)gridDim.x * blockDim.x;
  if (H.do_centroid) for (uint64_t k = gtid; k < l1; k += gsz) s[off_cen + 4 + k] = f.rc_tmp[0][k];
  if (H.with_color) for (uint64_t k = gtid; k < l2; k += gsz) s[off_col + 8 + k] = f.rc_tmp[1][k];
}

// ================================================================================================
// decode side
// ================================================================================================
struct ByteFeed {             // warp-uniform sequential byte reader with a 32-byte coalesced window
  const uint8_t *p; uint64_t len, pos; uint32_t window; uint64_t wbase;
  __device__ __forceinline__ void init(const uint8_t *ptr, uint64_t l, uint64_t start) { p = ptr; len = l; pos = start; wbase = ~0ull; window = 0; }
  __device__ __forceinline__ uint32_t next() {
    uint64_t b = pos & ~31ull;
    if (b != wbase) { uint64_t i = b + lane_id(); window = i < len ? p[i] : 0; wbase = b; }
    uint32_t v = __shfl_sync(FULL_MASK, window, (int)(pos & 31));
    pos++;
    return v;
  }
};

// decodeStreamToCharVector, warp-uniform. Symbol search without a division: lane l owns the cumulative
// boundaries of symbols 8l..8l+7 and tests freq[s]*r <= code-low; a ballot over the first boundary of each lane
// picks the lane, that lane's local count picks the symbol (== PCL's binary descent over freq[1..255]).
__device__ inline bool rc_decode_layer(ByteFeed &in, uint8_t *out, uint32_t n, uint32_t *freq_s /* smem 257 */, uint64_t *coded) {
  const uint32_t lane = lane_id();
  const uint64_t start = in.pos;
  if (in.pos + 1028 + 4 > in.len) return false;
  // table (raw u32 little-endian, possibly unaligned)
  for (uint32_t s = lane; s < 257; s += 32) {
    const uint8_t *q = in.p + in.pos + 4ull * s;
    freq_s[s] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
  }
  __syncwarp();
  in.pos += 1028;
  const uint32_t total = freq_s[256];
  uint32_t c[9], pkv[8];
#pragma unroll
  for (int k = 0; k < 9; k++) c[k] = freq_s[8 * lane + k];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 8; k++) { pkv[k] = c[k + 1] - c[k]; bad |= c[k + 1] <= c[k]; }
  // a table PCL's encoder can write: freq[0] = 0, strictly increasing, total below 1<<16.  Anything else could
  // drive range to 0 (an endless renormalisation loop in the reference as well), so it is rejected.
  if (__any_sync(FULL_MASK, bad) || freq_s[0] != 0 || total >= RC_BOTTOM) return false;
  const FastDiv fd = fastdiv_make(total);
  uint32_t code = 0, low = 0, range = 0xFFFFFFFFu;
  for (int k = 0; k < 4; k++) code = (code << 8) | in.next();
  uint32_t mysym = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t r = fastdiv(range, fd);
    const uint32_t v = code - low;
    uint32_t cnt = 0, cum = 0, wid = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { if (c[k] * r <= v) { cnt = k + 1; cum = c[k]; wid = pkv[k]; } }
    const uint32_t bal = __ballot_sync(FULL_MASK, cnt > 0);
    const uint32_t L = 31 - __clz(bal | 1u);             // highest lane whose first boundary is <= v (lane 0 always is)
    const uint32_t sym = 8 * L + __shfl_sync(FULL_MASK, cnt, L) - 1;
    cum = __shfl_sync(FULL_MASK, cum, L);
    wid = __shfl_sync(FULL_MASK, wid, L);
    if (lane == (i & 31)) mysym = sym;
    if ((i & 31) == 31) out[i - 31 + lane] = (uint8_t)mysym;
    low += cum * r;
    range = r * wid;
    for (;;) {
      if ((low ^ (low + range)) >= RC_TOP) {
        if (range >= RC_BOTTOM) break;
        range = (0u - low) & (RC_BOTTOM - 1);
      }
      code = (code << 8) | in.next();
      range <<= 8; low <<= 8;
    }
  }
  if (lane < (n & 31)) out[(n & ~31u) + lane] = (uint8_t)mysym;
  if (in.pos > in.len) return false;
  *coded = in.pos - start;
  return true;
}

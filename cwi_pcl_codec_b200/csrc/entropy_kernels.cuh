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
  else if (which == 2) { src = f.cpay; n = f.ncolor; }
  else if (which == 3) { src = f.pdiff; n = f.npd; }       // detail mode: point differences
  else { src = f.cdiff; n = f.ncd; }                       // detail mode: colour differences
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

// number of leading bits (a multiple of 8, 0..24) on which low and low+range agree, from x = low ^ (low+range) != 0
// (compare/select form: FLO sits on the slow XU pipe, ~34 cycles on B200; three ISETP + SEL + IADD3 take ~12)
__device__ __forceinline__ uint32_t rc_equal_bits(uint32_t x) {
  return ((x < (1u << 24)) ? 8u : 0u) + ((x < (1u << 16)) ? 8u : 0u) + ((x < (1u << 8)) ? 8u : 0u);
}

// ---- range encoder: one block per frame (steered over the SMs), one warp per layer (tree, centroid, colour)
__global__ void __launch_bounds__(160) rc_encode_kernel(EncFrame *frames, int do_centroid, int do_color, int detail) {
  EncFrame &f = frames[blockIdx.x];
  if (threadIdx.x == 0) f.serial_sm = sm_id();
  if (f.V == 0) return;
  const int which = threadIdx.x >> 5;
  if ((which == 1 && !do_centroid) || (which == 2 && !do_color) || (which == 3 && !detail) || (which == 4 && !(detail && do_color))) return;
  const uint8_t *src; uint32_t n;
  enc_layer(f, which, src, n);
  uint8_t *dst; uint64_t cap;
  if (which == 0) { dst = f.stream + FRAME_HDR_BYTES + 8; cap = f.stream_cap > FRAME_HDR_BYTES + 8 ? f.stream_cap - FRAME_HDR_BYTES - 8 : 0; }
  else { dst = f.rc_tmp[which - 1]; cap = f.rc_tmp_cap[which - 1]; }
  __shared__ uint32_t freq_all[5][257];
  __shared__ uint32_t packed_all[5][256];
  uint32_t *freq = freq_all[which], *packed = packed_all[which];
  const uint32_t lane = lane_id();
  if (lane == 0) rc_build_table(f.hist + which * 256, freq);
  __syncwarp();
  if (cap < 1028 + 8) { if (lane == 0) atomicOr(&f.error, FERR_STREAM_CAP); return; }
  for (uint32_t s = lane; s < 257; s += 32) ((uint32_t *)dst)[s] = freq[s];       // dst is 4-byte aligned
  for (uint32_t s = lane; s < 256; s += 32) packed[s] = (freq[s] << 16) | (freq[s + 1] - freq[s]);
  __syncwarp();
  const FastDiv fd = fastdiv_make(freq[256]);
  // Output: bytes are collected big-endian in a 64-bit accumulator (warp-uniform) and leave as aligned 32-bit words.
  uint32_t *out32 = (uint32_t *)(dst + 1028);
  const uint32_t cap_words = (uint32_t)min((uint64_t)0x7FFFFFFFu, (cap - 1028) / 4);
  uint32_t low = 0, range = 0xFFFFFFFFu;
  uint64_t acc = 0; uint32_t nacc = 0; uint32_t wp = 0;
  bool overflow = false;
  // one symbol: map the range, then renormalise.  Closed form of PCL's while-loop: the bytes on which low and
  // low+range agree leave first (x = low ^ (low+range) shifts along with both); the rare underflow case (range <
  // bottom while the top bytes differ) loops.  The word flush is branch-free: a taken branch costs ~19 cycles
  // on a lone warp, so the store is predicated and the counters move by selects.
#define RC_FLUSH() do { const bool fl_ = nacc >= 4; const uint32_t nn_ = fl_ ? nacc - 4 : nacc; \
    if (fl_) out32[wp] = __byte_perm((uint32_t)(acc >> (8 * nn_)), 0, 0x0123); wp += fl_ ? 1u : 0u; nacc = nn_; } while (0)
#define RC_ENC_SYMBOL(P) do { const uint32_t p_ = (P); const uint32_t r_ = fastdiv(range, fd); \
    low += (p_ >> 16) * r_; range = r_ * (p_ & 0xFFFFu); \
    const uint32_t sh_ = rc_equal_bits(low ^ (low + range)); \
    acc = (acc << sh_) | __funnelshift_l(low, 0, sh_); nacc += sh_ >> 3; low <<= sh_; range <<= sh_; \
    while (__builtin_expect(range < RC_BOTTOM, 0)) { \
      RC_FLUSH(); \
      range = (0u - low) & (RC_BOTTOM - 1); acc = (acc << 8) | (low >> 24); nacc++; low <<= 8; range <<= 8; \
      const uint32_t s2_ = rc_equal_bits(low ^ (low + range)); \
      acc = (acc << s2_) | __funnelshift_l(low, 0, s2_); nacc += s2_ >> 3; low <<= s2_; range <<= s2_; } \
    RC_FLUSH(); } while (0)
  // fast path: straight-line code for 8 symbols at a time (their table entries fetched by 8 independent shuffles
  // up front); the only branches on it are never-taken forward jumps to the underflow handler below.
#define RC_ENC_FAST(P, K) do { const uint32_t p_ = (P); const uint32_t r_ = fastdiv(range, fd); \
    low += (p_ >> 16) * r_; range = r_ * (p_ & 0xFFFFu); \
    const uint32_t sh_ = rc_equal_bits(low ^ (low + range)); \
    acc = (acc << sh_) | __funnelshift_l(low, 0, sh_); nacc += sh_ >> 3; low <<= sh_; range <<= sh_; \
    if (__builtin_expect(range < RC_BOTTOM, 0)) { kk = (K); goto slow_path; } \
    RC_FLUSH(); } while (0)
  const uint32_t nfull = n & ~31u;
  uint32_t pk_next = nfull ? packed[src[lane]] : 0;
  for (uint32_t base = 0; base < nfull; base += 32) {
    const uint32_t pk = pk_next;
    if (base + 32 < nfull) pk_next = packed[src[base + 32 + lane]];     // software prefetch of the next batch
    if (wp + 40 > cap_words) { overflow = true; break; }            // a symbol emits at most 4 bytes
    uint32_t kk;
    for (uint32_t g = 0; g < 32; g += 8) {
      const uint32_t q0 = __shfl_sync(FULL_MASK, pk, g), q1 = __shfl_sync(FULL_MASK, pk, g + 1), q2 = __shfl_sync(FULL_MASK, pk, g + 2),
                     q3 = __shfl_sync(FULL_MASK, pk, g + 3), q4 = __shfl_sync(FULL_MASK, pk, g + 4), q5 = __shfl_sync(FULL_MASK, pk, g + 5),
                     q6 = __shfl_sync(FULL_MASK, pk, g + 6), q7 = __shfl_sync(FULL_MASK, pk, g + 7);
      RC_ENC_FAST(q0, g); RC_ENC_FAST(q1, g + 1); RC_ENC_FAST(q2, g + 2); RC_ENC_FAST(q3, g + 3);
      RC_ENC_FAST(q4, g + 4); RC_ENC_FAST(q5, g + 5); RC_ENC_FAST(q6, g + 6); RC_ENC_FAST(q7, g + 7);
    }
    continue;
  slow_path:                                                         // rare: finish symbol kk's underflow, then the rest of the batch
    do {
      RC_FLUSH();
      range = (0u - low) & (RC_BOTTOM - 1); acc = (acc << 8) | (low >> 24); nacc++; low <<= 8; range <<= 8;
      const uint32_t s2 = rc_equal_bits(low ^ (low + range));
      acc = (acc << s2) | __funnelshift_l(low, 0, s2); nacc += s2 >> 3; low <<= s2; range <<= s2;
    } while (range < RC_BOTTOM);
    RC_FLUSH();
    for (uint32_t k = kk + 1; k < 32; k++) RC_ENC_SYMBOL(__shfl_sync(FULL_MASK, pk, k));
  }
#undef RC_ENC_FAST
  if (!overflow && nfull < n) {
    const uint32_t i = nfull + lane;
    const uint32_t pk = i < n ? packed[src[i]] : 0;
    if (wp + 40 > cap_words) overflow = true;
    else for (uint32_t k = 0; k < n - nfull; k++) RC_ENC_SYMBOL(__shfl_sync(FULL_MASK, pk, k));
  }
  uint64_t cnt = 0;
  if (!overflow) {
    for (int k = 0; k < 4; k++) {                                    // flush
      acc = (acc << 8) | (low >> 24); nacc++; low <<= 8;
      RC_FLUSH();
    }
    uint8_t *tail = (uint8_t *)(out32 + wp);
    for (uint32_t k = 0; k < nacc; k++) tail[k] = (uint8_t)(acc >> (8 * (nacc - 1 - k)));
    cnt = 4ull * wp + nacc;
  }
#undef RC_ENC_SYMBOL
#undef RC_FLUSH
  if (lane == 0) {
    if (overflow) atomicOr(&f.error, FERR_STREAM_CAP);
    f.rc_len[which] = (uint32_t)(1028 + cnt);
  }
}

// ---- range encoder, lane per stream ---------------------------------------------------------------------------------
// The coder's state machine is scalar, so in rc_encode_kernel 31 of a warp's 32 lanes repeat lane 0's arithmetic and a
// frame costs a whole warp's issue slots; with hundreds of frames in flight the SMs run out of them (7 coder CTAs per
// SM at 1024 frames).  Here every LANE is a coder of its own: a warp encodes the same layer of 32 frames in lock step
// (same instruction stream, 32 different (low, range) states, tables interleaved in shared memory so that lane l only
// ever touches bank l), so 1024 frames need 32 warps per layer and each of them runs at the latency of a lone warp.
// The arithmetic per symbol is the one above (closed-form renormalisation, branch-free word flush); only the rare
// underflow loop diverges.  grid (ceil(frames / 32), 3 layers), one warp per CTA.
#define LPS_SYMS 16
__global__ void __launch_bounds__(32) rc_encode_lps_kernel(EncFrame *frames, int nframes, int do_centroid, int do_color, int detail) {
  const int which = blockIdx.y;
  if ((which == 1 && !do_centroid) || (which == 2 && !do_color) || (which == 3 && !detail) || (which == 4 && !(detail && do_color))) return;
  __shared__ uint32_t tab[257 * 32];                       // [symbol][lane]: cumulative table, then (cum << 16 | width)
  const uint32_t lane = lane_id();
  const int fi = blockIdx.x * 32 + lane;
  const bool have = fi < nframes && frames[fi].V != 0;
  EncFrame &f = frames[have ? fi : blockIdx.x * 32];
  const uint8_t *src = nullptr; uint32_t n = 0;
  uint8_t *dst = nullptr; uint64_t cap = 0;
  if (have) {
    enc_layer(f, which, src, n);
    if (which == 0) { dst = f.stream + FRAME_HDR_BYTES + 8; cap = f.stream_cap > FRAME_HDR_BYTES + 8 ? f.stream_cap - FRAME_HDR_BYTES - 8 : 0; }
    else { dst = f.rc_tmp[which - 1]; cap = f.rc_tmp_cap[which - 1]; }
    if (which == 0) f.serial_sm = sm_id();
  }
  bool overflow = false;
  if (have && cap < 1028 + 8) { overflow = true; n = 0; }
  // cumulative table with PCL's "+1 if empty" rule and the halving rescale (rc_build_table), one column per lane
  uint32_t *col = tab + lane;
  {
    const uint32_t *hist = f.hist + which * 256;
    uint32_t prev = 0; col[0] = 0;
    for (int s2 = 1; s2 <= 256; s2++) { uint32_t v = prev + (have ? hist[s2 - 1] : 1u); if (v <= prev) v = prev + 1; col[s2 * 32] = v; prev = v; }
    while (col[256 * 32] >= RC_BOTTOM) {
      prev = 0;
      for (int s2 = 1; s2 <= 256; s2++) { uint32_t v = col[s2 * 32] >> 1; if (v <= prev) v = prev + 1; col[s2 * 32] = v; prev = v; }
    }
  }
  const FastDiv fd = fastdiv_make(col[256 * 32]);
  if (have && !overflow) for (int s2 = 0; s2 <= 256; s2++) ((uint32_t *)dst)[s2] = col[s2 * 32];     // dst is 4-byte aligned
  { uint32_t c0 = col[0]; for (int s2 = 0; s2 < 256; s2++) { const uint32_t c1 = col[(s2 + 1) * 32]; col[s2 * 32] = (c0 << 16) | (c1 - c0); c0 = c1; } }
  __syncwarp();
  uint32_t *out32 = (uint32_t *)(dst + 1028);
  const uint32_t cap_words = (have && !overflow) ? (uint32_t)min((uint64_t)0x7FFFFFFFu, (cap - 1028) / 4) : 0;
  uint32_t low = 0, range = 0xFFFFFFFFu;
  uint64_t acc = 0; uint32_t nacc = 0, wp = 0;
#define LPS_FLUSH() do { const bool fl_ = nacc >= 4; const uint32_t nn_ = fl_ ? nacc - 4 : nacc; \
    if (fl_) out32[wp] = __byte_perm((uint32_t)(acc >> (8 * nn_)), 0, 0x0123); wp += fl_ ? 1u : 0u; nacc = nn_; } while (0)
#define LPS_MAP(P) const uint32_t p_ = (P); const uint32_t r_ = fastdiv(range, fd); \
    low += (p_ >> 16) * r_; range = r_ * (p_ & 0xFFFFu); \
    const uint32_t sh_ = rc_equal_bits(low ^ (low + range)); \
    acc = (acc << sh_) | __funnelshift_l(low, 0, sh_); nacc += sh_ >> 3; low <<= sh_; range <<= sh_;
#define LPS_UNDERFLOW() do { \
      LPS_FLUSH(); \
      range = (0u - low) & (RC_BOTTOM - 1); acc = (acc << 8) | (low >> 24); nacc++; low <<= 8; range <<= 8; \
      const uint32_t s2_ = rc_equal_bits(low ^ (low + range)); \
      acc = (acc << s2_) | __funnelshift_l(low, 0, s2_); nacc += s2_ >> 3; low <<= s2_; range <<= s2_; } while (range < RC_BOTTOM)
#define LPS_SYMBOL(P) do { LPS_MAP(P) if (__builtin_expect(range < RC_BOTTOM, 0)) LPS_UNDERFLOW(); LPS_FLUSH(); } while (0)
  // fast path: 16 symbols of straight-line code, their table entries fetched up front; the only branches on it are
  // never-taken forward jumps to the underflow handler placed after the loop body (a lane that takes one finishes
  // the batch on its own and meets the others at the loop end).
#define LPS_FAST(P, K) do { LPS_MAP(P) if (__builtin_expect(range < RC_BOTTOM, 0)) { kk = (K); goto slow_path; } LPS_FLUSH(); } while (0)
  const uint32_t nfull = n & ~(uint32_t)(LPS_SYMS - 1);
  uint32_t nmax = nfull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL_MASK, nmax, o));
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (nfull) nxt = *(const uint4 *)src;                                // layers start 16-byte aligned
  for (uint32_t base = 0; base < nmax; base += LPS_SYMS) {
    if (base >= nfull) continue;                                       // lanes whose vector is shorter idle (layers of one kind have similar lengths)
    if (wp + LPS_SYMS + 8 > cap_words) { overflow = true; break; }     // a symbol emits at most 4 bytes
    uint32_t pk[LPS_SYMS];
    {
      const uint4 cur = nxt;
      if (base + LPS_SYMS < nfull) nxt = *(const uint4 *)(src + base + LPS_SYMS);   // next batch in flight while this one is coded
      const uint32_t wv[4] = { cur.x, cur.y, cur.z, cur.w };
#pragma unroll
      for (int k = 0; k < 4; k++) {
        pk[4 * k] = col[(wv[k] & 255u) * 32]; pk[4 * k + 1] = col[((wv[k] >> 8) & 255u) * 32];
        pk[4 * k + 2] = col[((wv[k] >> 16) & 255u) * 32]; pk[4 * k + 3] = col[(wv[k] >> 24) * 32];
      }
    }
    uint32_t kk;
    LPS_FAST(pk[0], 0); LPS_FAST(pk[1], 1); LPS_FAST(pk[2], 2); LPS_FAST(pk[3], 3);
    LPS_FAST(pk[4], 4); LPS_FAST(pk[5], 5); LPS_FAST(pk[6], 6); LPS_FAST(pk[7], 7);
    LPS_FAST(pk[8], 8); LPS_FAST(pk[9], 9); LPS_FAST(pk[10], 10); LPS_FAST(pk[11], 11);
    LPS_FAST(pk[12], 12); LPS_FAST(pk[13], 13); LPS_FAST(pk[14], 14); LPS_FAST(pk[15], 15);
    continue;
  slow_path:                                                           // rare: finish symbol kk's underflow, then the rest of the batch
    LPS_UNDERFLOW();
    LPS_FLUSH();
    for (uint32_t k = kk + 1; k < LPS_SYMS; k++) LPS_SYMBOL(col[(uint32_t)src[base + k] * 32]);
  }
#undef LPS_FAST
  if (!overflow && nfull < n) {
    if (wp + LPS_SYMS + 8 > cap_words) overflow = true;
    else for (uint32_t i = nfull; i < n; i++) LPS_SYMBOL(col[(uint32_t)src[i] * 32]);
  }
  if (!have) return;
  uint64_t cnt = 0;
  if (!overflow) {
    for (int k = 0; k < 4; k++) { acc = (acc << 8) | (low >> 24); nacc++; low <<= 8; LPS_FLUSH(); }   // flush
    uint8_t *tail = (uint8_t *)(out32 + wp);
    for (uint32_t k = 0; k < nacc; k++) tail[k] = (uint8_t)(acc >> (8 * (nacc - 1 - k)));
    cnt = 4ull * wp + nacc;
  }
#undef LPS_SYMBOL
#undef LPS_UNDERFLOW
#undef LPS_MAP
#undef LPS_FLUSH
  if (overflow) atomicOr(&f.error, FERR_STREAM_CAP);
  f.rc_len[which] = (uint32_t)(1028 + cnt);
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
  const uint64_t off_cnt = off_col + (H.with_color ? 8 + l2 : 0);               // detail mode (impl.hpp:1728-1757): counts, point differences, colour differences
  const uint64_t li = H.do_voxel_grid ? 0 : f.rc_int_len, l3 = H.do_voxel_grid ? 0 : f.rc_len[3], l4 = (!H.do_voxel_grid && H.with_color) ? f.rc_len[4] : 0;
  const uint64_t off_pd = off_cnt + (H.do_voxel_grid ? 0 : 8 + li);
  const uint64_t off_cd = off_pd + (H.do_voxel_grid ? 0 : 8 + l3);
  const uint64_t total = off_cd + ((!H.do_voxel_grid && H.with_color) ? 8 + l4 : 0);
  if (total > f.stream_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&f.error, FERR_STREAM_CAP); f.out_len = 0; } return; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const char id2[] = "<PCL-OCT-CODECV2-COMPRESSED>", id1[] = "<PCL-OCT-COMPRESSED>";
    for (int k = 0; k < 28; k++) s[k] = id2[k];
    for (int k = 0; k < 20; k++) s[28 + k] = id1[k];
    auto put = [&](uint64_t off, const void *p, int nb) { const uint8_t *q = (const uint8_t *)p; for (int k = 0; k < nb; k++) s[off + k] = q[k]; };
    uint32_t fid = f.frame_id; put(48, &fid, 4);
    s[52] = 1; s[53] = H.do_voxel_grid; s[54] = H.with_color;
    uint64_t pc = H.do_voxel_grid ? f.V : f.n_finite; put(55, &pc, 8);   // [PCL] writeFrameHeader: leaf_count_ or object_count_
    put(63, &H.octree_res, 8); s[71] = H.color_bits; put(72, &H.point_res, 8);
    put(80, f.bmin, 24); put(104, f.bmax, 24);
    s[128] = H.do_centroid; s[129] = H.connectivity; s[130] = H.scalable;
    put(131, &H.color_type, 4); put(135, &H.macroblock, 4); s[139] = H.icp_offset;
    uint64_t B = f.B; put(140, &B, 8);
    if (H.do_centroid) { uint32_t c = f.ncen; put(off_cen, &c, 4); }
    if (H.with_color) { uint64_t c = f.ncolor; put(off_col, &c, 8); }
    if (!H.do_voxel_grid) {
      uint64_t c = f.V; put(off_cnt, &c, 8);
      c = f.npd; put(off_pd, &c, 8);
      if (H.with_color) { c = f.ncd; put(off_cd, &c, 8); }
    }
    f.out_len = total;
    f.coded[0] = l0; f.coded[1] = l1; f.coded[2] = l2;
  }
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (uint64_t)gridDim.x * blockDim.x;
  if (H.do_centroid) for (uint64_t k = gtid; k < l1; k += gsz) s[off_cen + 4 + k] = f.rc_tmp[0][k];
  if (H.with_color) for (uint64_t k = gtid; k < l2; k += gsz) s[off_col + 8 + k] = f.rc_tmp[1][k];
  if (!H.do_voxel_grid) {
    for (uint64_t k = gtid; k < li; k += gsz) s[off_cnt + 8 + k] = f.rc_int[k];
    for (uint64_t k = gtid; k < l3; k += gsz) s[off_pd + 8 + k] = f.rc_tmp[2][k];
    for (uint64_t k = gtid; k < l4; k += gsz) s[off_cd + 8 + k] = f.rc_tmp[3][k];
  }
}

// ---- stream export: the assembled frame leaves the codec's slot for the caller's buffer (device memory, or pinned
// host memory through its device alias: zero-copy stores over PCIe -- few CTAs on purpose, measured 52 GB/s with
// 8-32 CTAs and 13 GB/s with 592, profiles/mb_pcie_r2.txt).  grid (blocks, frames)
__global__ void __launch_bounds__(256) export_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.y];
  const uint64_t n = f.out_len;
  if (n == 0 || f.error || !f.out_ptr) return;
  if (n > f.out_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&f.error, FERR_CALLER_CAP); return; }
  const uint8_t *s = f.stream; uint8_t *d = f.out_ptr;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (uint64_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
    const uint64_t n16 = n / 16;
    for (uint64_t k = gtid; k < n16; k += gsz) ((uint4 *)d)[k] = ((const uint4 *)s)[k];
    for (uint64_t k = n16 * 16 + gtid; k < n; k += gsz) d[k] = s[k];
  } else for (uint64_t k = gtid; k < n; k += gsz) d[k] = s[k];
}

// ================================================================================================
// decode side
// ================================================================================================
// Sequential input for the warp-uniform range decoder: a 64-bit big-endian window W over the stream, rebuilt from
// three cached aligned 64-bit words (A, B and the prefetched C) whenever whole bytes have been consumed.  The fast
// path consumes at most one byte per symbol, so the window is rebuilt once per 8 symbols, off the per-symbol path.
struct WindowFeed {
  const uint64_t *w; uint32_t last, widx, sbyte, ubits, consumed; uint64_t A, B, Craw, W;
  __device__ __forceinline__ uint64_t raw(uint32_t i) const { return w[min(i, last)]; }
  __device__ __forceinline__ static uint64_t be(uint64_t v) {           // the stream is consumed most significant byte first
    return ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
  }
  __device__ __forceinline__ void rebuild() { W = sbyte ? (A << (8 * sbyte)) | (B >> (64 - 8 * sbyte)) : A; }
  __device__ __forceinline__ void init(const uint8_t *base, uint64_t len, uint64_t pos) {
    const uintptr_t a = (uintptr_t)(base + pos);
    sbyte = (uint32_t)(a & 7);
    w = (const uint64_t *)(a - sbyte);
    const uintptr_t end = ((uintptr_t)(base + len) + 7) & ~(uintptr_t)7;
    last = (uint32_t)((end - (a - sbyte)) / 8) - 1;        // reads past the end repeat the last word (PCL would read EOF garbage)
    widx = 0; A = be(raw(0)); B = be(raw(1)); Craw = raw(2); ubits = 0; consumed = 0;
    rebuild();
  }
  // whole bytes consumed since the last rebuild are in ubits (multiple of 8, <= 64).  The word loaded here is not
  // touched (not even byte-swapped) until the NEXT shift, so its latency never lands on the decode chain.
  __device__ __forceinline__ void advance() {
    consumed += ubits >> 3; sbyte += ubits >> 3; ubits = 0;
    while (sbyte >= 8) { sbyte -= 8; A = B; B = be(Craw); widx++; Craw = raw(widx + 2); }
    rebuild();
  }
  __device__ __forceinline__ uint32_t take_byte() {        // generic path: any number of bytes
    if (ubits == 64) advance();                             // the window may have been drained by the fast path
    const uint32_t b = (uint32_t)(W >> 56);
    W <<= 8; ubits += 8;
    return b;
  }
};

// Shared-memory ring through which the range decoder (warp 0) hands the tree bytes to the DFS walker (warp 1) while it
// is still decoding: the walk is serial too, and pipelining it behind the decoder takes it off the frame's latency.
#define RING_WORDS 512
struct WalkRing { volatile uint32_t ring[RING_WORDS]; volatile uint32_t prod, cons, done, dead; uint32_t B, depth, go; };

// decodeStreamToCharVector, warp-uniform. Symbol search without a division: lane l owns the cumulative
// boundaries of symbols 8l..8l+7 and tests freq[s]*r <= code-low; a ballot over the first boundary of each lane
// picks the lane, a select tree over that lane's (monotone) predicates picks the symbol -- the same symbol as
// PCL's binary descent over freq[1..255].  out must be 4-byte aligned (symbols leave as 32-bit words).
template <bool RING>
__device__ inline bool rc_decode_layer(const uint8_t *base, uint64_t len, uint64_t &pos, uint8_t *out, uint32_t n,
                                       uint32_t *freq_s /* smem 257 */, uint64_t *coded, WalkRing *rg = nullptr) {
  const uint32_t lane = lane_id();
  const uint64_t start = pos;
  if (pos + 1028 + 4 > len) return false;
  for (uint32_t s = lane; s < 257; s += 32) {               // table: raw u32 little-endian, possibly unaligned
    const uint8_t *q = base + pos + 4ull * s;
    freq_s[s] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
  }
  __syncwarp();
  const uint32_t total = freq_s[256];
  uint32_t c[9], pkv[8];
#pragma unroll
  for (int k = 0; k < 9; k++) c[k] = freq_s[8 * lane + k];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 8; k++) { pkv[k] = (c[k] << 16) | ((c[k + 1] - c[k]) & 0xFFFFu); bad |= c[k + 1] <= c[k]; }
  // a table PCL's encoder can write: freq[0] = 0, strictly increasing, total below 1<<16.  Anything else could
  // drive range to 0 (an endless renormalisation loop in the reference as well), so it is rejected.
  if (__any_sync(FULL_MASK, bad) || freq_s[0] != 0 || total >= RC_BOTTOM) return false;
  __syncwarp();
  const FastDiv fd = fastdiv_make(total);
  WindowFeed in; in.init(base, len, pos + 1028);
  uint32_t code = 0, low = 0, range = 0xFFFFFFFFu;
  for (int k = 0; k < 4; k++) code = (code << 8) | in.take_byte();
  in.advance();
  uint32_t osym = 0;
  uint32_t *out32 = (uint32_t *)out;
  const uint32_t sym_base = 8 * lane;
  const bool last_lane = lane == 31;
  // One symbol.  r = range / total; lane-local boundaries t_k = freq[8*lane+k] * r against v = code - low.  The
  // predicates are monotone (true...true false...false) so a 3-level select tree finds the last true one; the
  // owner lane (first boundary <= v < first boundary of the next lane) broadcasts its (cum,width) and symbol through
  // REDUX.MAX (every other lane contributes 0) -- ~20 cycles instead of VOTE + FLO + SHFL (~85).
#define RC_DEC_MAP() \
    const uint32_t r_ = fastdiv(range, fd); const uint32_t v_ = code - low; \
    const bool p0 = c[0] * r_ <= v_, p1 = c[1] * r_ <= v_, p2 = c[2] * r_ <= v_, p3 = c[3] * r_ <= v_; \
    const bool p4 = c[4] * r_ <= v_, p5 = c[5] * r_ <= v_, p6 = c[6] * r_ <= v_, p7 = c[7] * r_ <= v_; \
    const bool own_ = p0 && (last_lane || !(c[8] * r_ <= v_)); \
    const uint32_t a01 = p1 ? pkv[1] : pkv[0], a23 = p3 ? pkv[3] : pkv[2], a45 = p5 ? pkv[5] : pkv[4], a67 = p7 ? pkv[7] : pkv[6]; \
    const uint32_t a03 = p2 ? a23 : a01, a47 = p6 ? a67 : a45; \
    const uint32_t i01 = p1 ? 1u : 0u, i23 = p3 ? 3u : 2u, i45 = p5 ? 5u : 4u, i67 = p7 ? 7u : 6u; \
    const uint32_t i03 = p2 ? i23 : i01, i47 = p6 ? i67 : i45; \
    const uint32_t pk_ = __reduce_max_sync(FULL_MASK, own_ ? (p4 ? a47 : a03) : 0u); \
    const uint32_t sym_ = __reduce_max_sync(FULL_MASK, own_ ? sym_base + (p4 ? i47 : i03) : 0u); \
    osym = (osym >> 8) | (sym_ << 24); \
    low += (pk_ >> 16) * r_; range = r_ * (pk_ & 0xFFFFu);
  // PCL's literal renormalisation loop (any number of bytes, underflow included)
#define RC_DEC_RENORM_GENERIC() do { for (;;) { \
      if ((low ^ (low + range)) >= RC_TOP) { if (range >= RC_BOTTOM) break; range = (0u - low) & (RC_BOTTOM - 1); } \
      code = (code << 8) | in.take_byte(); low <<= 8; range <<= 8; } } while (0)
#define RC_DEC_STORE(I) do { if (((I) & 3) == 3) { out32[(I) >> 2] = osym; if (RING) rg->ring[((I) >> 2) & (RING_WORDS - 1)] = osym; } } while (0)
  // fast path: the symbol shifts in 0, 1 or 2 bytes (99.9 % of symbols), does not underflow, and the window still
  // holds the bytes; anything else takes PCL's literal loop and then resumes the fast path at the next symbol.
#define RC_DEC_FAST(K) do { RC_DEC_MAP() \
    const uint32_t x_ = low ^ (low + range); \
    const uint32_t sh_ = x_ < RC_BOTTOM ? 16u : (x_ < RC_TOP ? 8u : 0u); \
    const uint32_t rs_ = range << sh_; \
    if (__builtin_expect((x_ < 256u) | (rs_ < RC_BOTTOM) | (in.ubits > 48u), 0)) { kk = (K); goto slow_path; } \
    code = __funnelshift_l((uint32_t)(in.W >> 32), code, sh_); in.W <<= sh_; in.ubits += sh_; low <<= sh_; range = rs_; \
    RC_DEC_STORE(i0 + (K)); } while (0)
#define RC_DEC_SYMBOL(I) do { RC_DEC_MAP() RC_DEC_RENORM_GENERIC(); RC_DEC_STORE(I); } while (0)
  const uint32_t n8 = n & ~7u;
  for (uint32_t i0 = 0; i0 < n8; i0 += 8) {
    uint32_t kk;
    if (RING && (i0 & 63) == 0 && i0) {                               // every 16 words: publish, and wait if the walker lags a ring behind
      __threadfence_block();
      rg->prod = i0 >> 2;
      while (!rg->dead && (i0 >> 2) - rg->cons > RING_WORDS - 64) { }
    }
    RC_DEC_FAST(0); resume1: RC_DEC_FAST(1); resume2: RC_DEC_FAST(2); resume3: RC_DEC_FAST(3);
    resume4: RC_DEC_FAST(4); resume5: RC_DEC_FAST(5); resume6: RC_DEC_FAST(6); resume7: RC_DEC_FAST(7);
    resume8: in.advance();
    continue;
  slow_path:                                                         // symbol kk is mapped but not renormalised yet
    RC_DEC_RENORM_GENERIC();
    RC_DEC_STORE(i0 + kk);
    switch (kk) { case 0: goto resume1; case 1: goto resume2; case 2: goto resume3; case 3: goto resume4;
                  case 4: goto resume5; case 5: goto resume6; case 6: goto resume7; default: goto resume8; }
  }
  for (uint32_t i = n8; i < n; i++) RC_DEC_SYMBOL(i);
  in.advance();
#undef RC_DEC_FAST
#undef RC_DEC_SYMBOL
#undef RC_DEC_STORE
#undef RC_DEC_RENORM_GENERIC
#undef RC_DEC_MAP
  if (n & 3) {
    const uint32_t rem = n & 3; osym >>= 8 * (4 - rem);
    for (uint32_t k = 0; k < rem; k++) out[(n & ~3u) + k] = (uint8_t)(osym >> (8 * k));
    if (RING) rg->ring[(n >> 2) & (RING_WORDS - 1)] = osym;
  }
  if (RING) { __threadfence_block(); rg->prod = (n + 3) >> 2; rg->done = 1; }
  pos = pos + 1028 + in.consumed;
  if (pos > len) return false;
  *coded = pos - start;
  return true;
}

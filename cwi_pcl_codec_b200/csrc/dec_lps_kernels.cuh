// dec_lps_kernels.cuh -- entropy stage of the decoder with a LANE-PER-STREAM range decoder for the tree layer.
//
// dec_entropy_kernel (dec_kernels.cuh) gives every frame a CTA whose warp 0 decodes the occupancy bytes warp-uniformly:
// 70 instructions per symbol on all 32 lanes for one stream.  With 1024 frames in flight that is 7 such warps per SM and
// the SMs run out of issue slots.  Here the tree layers of 16 frames are decoded by 16 lanes of ONE
// warp in lock step, each lane with its own (low, range, code) and byte window, so 1024 frames need 64 decoder warps,
// each alone on its SM and running at the latency of a lone warp.  A lane cannot use the other lanes to search the
// cumulative table, so the search is a table lookup: q ~ (code - low) / r from a float reciprocal (a lower bound within
// 2 of the true quotient), a 4096-entry lookup table per stream maps q >> 4 to the first candidate symbol, and up to two
// further boundaries are tested with exact integer products -- the symbol is the same one PCL's binary descent finds.
// More than two steps (0.1 % of symbols) or a renormalisation that is not 0/1/2 bytes takes a generic out-of-line path.
//
//   dec_head_kernel            warp per frame   header parse (syncToHeader, readFrameHeader, tree size word); locates the
//                                               colour layer by a backward scan for its signature (speculation: compressed
//                                               lengths are not stored, so its start is only known once the tree layer is decoded)
//   rc_decode_lps_kernel<true>  CTA per 16 frames: warp 0 = 16 decoder lanes for the tree layers; warps 1..16 = the DFS
//                               walkers of those frames, fed through shared-memory rings while the decoder runs
//   rc_decode_lps_kernel<false> the same decoder on the speculated colour layers, on a second stream at the same time
//   dec_jpeg_kernel             warp per frame (second stream): JPEG marker parse, de-stuffing, Huffman decode
//   dec_finish_kernel           warp per frame: centroid layer, accepts the colour speculation only if the layer starts
//                               exactly where the preceding layers end (otherwise decodes it itself), trailing-byte check
#pragma once
#include "dec_kernels.cuh"

#ifndef LPS_DEC_FRAMES
#define LPS_DEC_FRAMES 8
#endif
#define LPS_LUT_SHIFT 4
#define LPS_LUT_ENTRIES (65536 >> LPS_LUT_SHIFT)
#define LPS_Q_SLACK 3                                      // the float quotient estimate is within [-3, +0] of ... see LPS_DEC_LOOKUP

struct LpsSmem {
  uint32_t pair[(256 + 2) * LPS_DEC_FRAMES];               // [symbol][lane]: cum[s] | cum[s+1] << 16 (two pad rows: the search reads s0 + 1, s0 + 2 unguarded)
  uint8_t lut[LPS_LUT_ENTRIES * LPS_DEC_FRAMES];           // [bucket][lane]: largest s with cum[s] <= (bucket << LPS_LUT_SHIFT) - LPS_Q_SLACK
  WalkRing rg[LPS_DEC_FRAMES];
  uint32_t wstack[LPS_DEC_FRAMES][32];
  uint32_t wlut[256];
};

// ---- header: grid (frames), 32 threads.  Follows dec_entropy_kernel's first stage (impl.hpp:1660-1676, 1489-1502).
__device__ inline void dec_parse_header(DecFrame &f);
__global__ void __launch_bounds__(32) dec_head_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.x];
  if (threadIdx.x == 0) dec_parse_header(f);
  __syncwarp();
  if (!f.head_ok || !f.data_with_color) return;
  const uint32_t lane = lane_id();
  const uint8_t *in = f.in; const uint64_t len = f.in_len;
  if (len <= FRAME_HDR_BYTES + 8 + 1032 + 8 + 1032) return;
  // backward scan for the layer signature (a size word whose upper bytes are zero followed by a plausible cumulative
  // table); candidates o = offset of the u64 size word
  const uint64_t lo_lim = FRAME_HDR_BYTES + 8 + 1032;
  uint64_t found = 0; bool have = false;
  for (uint64_t hi = len - 1040; !have && hi >= lo_lim; hi = hi >= 32 + lo_lim ? hi - 32 : 0) {
    const uint64_t o = hi >= lane ? hi - lane : 0;
    bool q = o >= lo_lim && (in[o + 4] | in[o + 5] | in[o + 6] | in[o + 7] | in[o + 8] | in[o + 9] | in[o + 10] | in[o + 11] | in[o + 14] | in[o + 15]) == 0;
    uint32_t cand = __ballot_sync(FULL_MASK, q);
    while (cand && !have) {                               // validate candidates from the highest offset down
      const uint32_t l = __ffs(cand) - 1; cand &= cand - 1;
      const uint64_t oc = hi - l;
      bool okk = true;
      for (uint32_t s2 = lane; s2 < 256; s2 += 32) {
        const uint32_t a = ld_u32_unaligned(in + oc + 8 + 4 * s2), b2 = ld_u32_unaligned(in + oc + 12 + 4 * s2);
        okk &= (b2 > a) & (b2 < RC_BOTTOM);
      }
      if (__all_sync(FULL_MASK, okk)) { found = oc; have = true; }
    }
    if (hi < 32 + lo_lim) break;
  }
  if (have && lane == 0) {
    const uint64_t nc = ld_u64_unaligned(in + found);
    if (nc > 0 && nc <= f.col_cap) { f.spec_pos = found; f.spec_ncol = (uint32_t)nc; f.spec_found = 1; }
  }
}
__device__ inline void dec_parse_header(DecFrame &f) {
  f.serial_sm = sm_id();
  const uint8_t *in = f.in;
  const uint64_t len = f.in_len;
  uint32_t err = f.error;
  if (err) { f.V = 0; f.B = 0; f.point_count = 0; return; }
  uint64_t pos = 0;
  const char id2[] = "<PCL-OCT-CODECV2-COMPRESSED>", id1[] = "<PCL-OCT-COMPRESSED>";
  uint32_t hp = 0; bool ok = true;
  while (hp < 28) {
    if (pos >= len) { ok = false; break; }
    uint8_t c = in[pos++];
    if (c == 0xFF) { ok = false; break; }                 // (char)0xFF == EOF quirk, SURVEY App. C-9
    if (c != (uint8_t)id2[hp++]) hp = ((uint8_t)id2[0] == c) ? 1 : 0;
  }
  hp = 0;
  while (ok && hp < 20) {
    if (pos >= len) { ok = false; break; }
    uint8_t c = in[pos++];
    if (c != (uint8_t)id1[hp++]) hp = ((uint8_t)id1[0] == c) ? 1 : 0;
  }
  if (!ok || pos + 92 + 8 > len) err = FERR_BAD_STREAM;
  if (!err) {
    const uint8_t *h = in + pos;
    f.frame_id = ld_u32_unaligned(h);
    f.data_with_color = h[6];
    const uint64_t point_count = ld_u64_unaligned(h + 7);
    const double res = ld_f64_unaligned(h + 15);
    f.color_bits = h[23];
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; a++) { bmin[a] = ld_f64_unaligned(h + 32 + 8 * a); bmax[a] = ld_f64_unaligned(h + 56 + 8 * a); f.bmin[a] = bmin[a]; f.bmax[a] = bmax[a]; }
    f.do_centroid = h[80];
    f.cct = ld_u32_unaligned(h + 83);
    f.point_res_f = (float)ld_f64_unaligned(h + 24);
    f.point_count = point_count; f.res = res;
    pos += 92;
    // [PCL] readFrameHeader -> defineBoundingBox -> getKeyBitSize
    const double eps = 1.1920928955078125e-07;
    uint32_t mk = 2, depth = 0;
    for (int a = 0; a < 3; a++) {
      double t = ceil(__ddiv_rn(__dsub_rn(__dsub_rn(bmax[a], bmin[a]), eps), res));
      uint32_t k = (t >= 4294967295.0 || !(t == t)) ? 0xFFFFFFFFu : (t > 0 ? (uint32_t)t : 0u);
      if (k > mk) mk = k;
    }
    while ((1ull << depth) < mk) depth++;
    f.depth = depth;
    if (depth > CCV2_MAX_DEPTH || !(res > 0)) err |= FERR_DEPTH;
    if (point_count > f.out_cap) err |= FERR_OUT_CAP;
    const uint64_t B = ld_u64_unaligned(in + pos); pos += 8;
    if (B > f.tree_cap) err |= FERR_TREE_CAP;
    f.tree_pos = pos; f.tree_n = (uint32_t)min(B, (uint64_t)0xFFFFFFFFu);
  }
  if (err) { f.error |= err; f.V = 0; f.B = 0; f.point_count = 0; }
  else f.head_ok = 1;
}

// ---- tree layer, lane per stream + pipelined walkers.  grid (ceil(frames / 16)), 32 * 17 threads, dynamic shared memory
template <bool TREE>
__global__ void __launch_bounds__(TREE ? 32 * (1 + LPS_DEC_FRAMES) : 32) rc_decode_lps_kernel(DecFrame *frames, int nframes, int use_ring) {
  extern __shared__ __align__(16) uint8_t lps_raw[];
  LpsSmem &S = *reinterpret_cast<LpsSmem *>(lps_raw);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  if (TREE) for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) S.wlut[i] = ((uint32_t)__popc(i) << 16) | ((i ? (uint32_t)(__ffs(i) - 1) : 0u) << 8) | (i & (i - 1));
  if (TREE && threadIdx.x < LPS_DEC_FRAMES) { WalkRing &r = S.rg[threadIdx.x]; r.prod = 0; r.cons = 0; r.done = 0; r.dead = 0; r.go = 0; r.B = 0; r.depth = 0; }
  __syncthreads();

  // ------------------------------------------------------------------ set-up of the decoder lanes (warp 0, lanes 0..15)
  const int fi = blockIdx.x * LPS_DEC_FRAMES + (int)lane;
  const bool dec_lane = warp == 0 && lane < LPS_DEC_FRAMES && fi < nframes;
  DecFrame &f = frames[dec_lane ? fi : blockIdx.x * LPS_DEC_FRAMES];
  bool live = dec_lane && f.head_ok && !f.error && (TREE || f.spec_found);
  uint32_t n = 0;
  uint32_t *pair = S.pair + lane; uint8_t *lut = S.lut + lane;
  FastDiv fd = fastdiv_make(256);
  WindowFeed in;
  uint32_t code = 0, low = 0, range = 0xFFFFFFFFu;
  uint32_t *out32 = nullptr;
  WalkRing *rg = &S.rg[lane & (LPS_DEC_FRAMES - 1)];
  if (warp == 0) {
    const uint8_t *base = f.in; const uint64_t len = f.in_len, pos = TREE ? f.tree_pos : f.spec_pos + 8;
    if (live) { n = TREE ? f.tree_n : f.spec_ncol; if (pos + 1028 + 4 > len) live = false; }
    if (live) {
      // table: raw u32 little-endian, possibly unaligned.  A table PCL's encoder can write: freq[0] = 0, strictly
      // increasing, total below 1 << 16; anything else could drive range to 0, so it is rejected.
      uint32_t prev = ld_u32_unaligned(base + pos);
      bool bad = prev != 0;
      for (int s = 0; s < 256; s++) {
        const uint32_t nx = ld_u32_unaligned(base + pos + 4ull * (s + 1));
        bad |= nx <= prev || nx >= RC_BOTTOM;
        pair[s * LPS_DEC_FRAMES] = (prev & 0xFFFFu) | (nx << 16);
        prev = nx;
      }
      pair[256 * LPS_DEC_FRAMES] = pair[257 * LPS_DEC_FRAMES] = 0xFFFF0000u;
      if (bad) live = false;
      else {
        fd = fastdiv_make(prev);
        uint32_t s = 0;
        for (uint32_t b = 0; b < LPS_LUT_ENTRIES; b++) {
          const uint32_t target = (b << LPS_LUT_SHIFT) > LPS_Q_SLACK ? (b << LPS_LUT_SHIFT) - LPS_Q_SLACK : 0u;
          while (s < 255 && (pair[s * LPS_DEC_FRAMES] >> 16) <= target) s++;
          lut[b * LPS_DEC_FRAMES] = (uint8_t)s;
        }
        in.init(base, len, pos + 1028);
        for (int k = 0; k < 4; k++) code = (code << 8) | in.take_byte();
        in.advance();
        out32 = (uint32_t *)(TREE ? f.tree : f.col);
      }
    }
    if (TREE && dec_lane && !live && !f.error && f.head_ok) { atomicOr(&f.error, FERR_BAD_STREAM); f.B = 0; f.V = 0; }   // a failed colour speculation is not an error
    if (!live) n = 0;
    if (TREE && lane < LPS_DEC_FRAMES) {
      rg->B = n; rg->depth = live ? f.depth : 0;
      rg->go = (use_ring && live && n > 0 && f.depth >= 1 && f.depth <= 17) ? 1u : 0u;
      if (!live) { rg->dead = 1; rg->done = 1; }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ walkers: warp w serves frame slot w - 1
  if (TREE && warp > 0) {
    const int wf = blockIdx.x * LPS_DEC_FRAMES + (int)warp - 1;
    if (lane == 0 && wf < nframes && S.rg[warp - 1].go) dfs_walk_ring(frames[wf], &S.rg[warp - 1], S.wlut, S.wstack[warp - 1]);
    return;
  }

  // ------------------------------------------------------------------ decoder lanes
  const uint32_t pair_a = (uint32_t)__cvta_generic_to_shared(pair), lut_a = (uint32_t)__cvta_generic_to_shared(lut);
  const bool ring = TREE && rg->go != 0;
  uint32_t osym = 0;
  uint8_t *const out8 = (uint8_t *)out32;
  // candidate search: s0 from the lookup table, then exact tests of the next boundaries
  // q = (code - low) / r estimated in float: conversions round toward zero, rcp.approx and the product are within a few
  // ulp, and the float -> integer step is the 2^23 trick (round to nearest), so the estimate is within [-2, +1] of the
  // true quotient (< 65536); the table is built for estimate - LPS_Q_SLACK, which makes s0 a lower bound of the symbol.
#define LPS_DEC_LOOKUP() \
    const uint32_t r_ = fastdiv(range, fd); const uint32_t v_ = code - low; \
    float rc_; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc_) : "f"(__uint2float_rz(r_))); \
    const uint32_t qa_ = __float_as_uint(fminf(__uint2float_rz(v_) * rc_, 65535.0f) + 8388608.0f) & 0xFFFFu; \
    const uint32_t s0_ = lds_u8(lut_a + (qa_ >> LPS_LUT_SHIFT) * LPS_DEC_FRAMES);
  // generic symbol: any number of search steps, PCL's literal renormalisation loop (underflow included)
#define LPS_DEC_GENERIC(I) do { LPS_DEC_LOOKUP() \
    uint32_t s_ = s0_, p_ = lds_u32(pair_a + s_ * (4 * LPS_DEC_FRAMES)); \
    while (s_ < 255u && (p_ >> 16) * r_ <= v_) { s_++; p_ = lds_u32(pair_a + s_ * (4 * LPS_DEC_FRAMES)); } \
    osym = (osym >> 8) | (s_ << 24); \
    low += (p_ & 0xFFFFu) * r_; range = r_ * ((p_ >> 16) - (p_ & 0xFFFFu)); \
    for (;;) { \
      if ((low ^ (low + range)) >= RC_TOP) { if (range >= RC_BOTTOM) break; range = (0u - low) & (RC_BOTTOM - 1); } \
      code = (code << 8) | in.take_byte(); low <<= 8; range <<= 8; } \
    LPS_DEC_STORE(I); } while (0)
#define LPS_DEC_STORE(I) do { if (((I) & 3) == 3) { out32[(I) >> 2] = osym; if (ring) rg->ring[((I) >> 2) & (RING_WORDS - 1)] = osym; } } while (0)
  // fast path: at most two search steps, 0/1/2 renormalisation bytes, window not exhausted; nothing is committed
  // before the symbol is known to qualify, so the generic path can redo it from the same state
#define LPS_DEC_FAST(K) do { LPS_DEC_LOOKUP() \
    const uint32_t pad_ = pair_a + s0_ * (4 * LPS_DEC_FRAMES); \
    const uint32_t pa_ = lds_u32(pad_), pb_ = lds_u32(pad_ + 4 * LPS_DEC_FRAMES), pc_ = lds_u32(pad_ + 8 * LPS_DEC_FRAMES); \
    const bool t1_ = s0_ < 255u && (pa_ >> 16) * r_ <= v_; \
    const bool t2_ = t1_ && s0_ < 254u && (pb_ >> 16) * r_ <= v_; \
    const bool t3_ = t2_ && s0_ < 253u && (pc_ >> 16) * r_ <= v_; \
    const uint32_t p_ = t2_ ? pc_ : (t1_ ? pb_ : pa_); \
    const uint32_t s_ = s0_ + (t1_ ? 1u : 0u) + (t2_ ? 1u : 0u); \
    const uint32_t lo2_ = low + (p_ & 0xFFFFu) * r_, rg2_ = r_ * ((p_ >> 16) - (p_ & 0xFFFFu)); \
    const uint32_t x_ = lo2_ ^ (lo2_ + rg2_); \
    const uint32_t sh_ = x_ < RC_BOTTOM ? 16u : (x_ < RC_TOP ? 8u : 0u); \
    const uint32_t rs_ = rg2_ << sh_; \
    if (__builtin_expect(t3_ | (x_ < 256u) | (rs_ < RC_BOTTOM) | (in.ubits > 48u), 0)) { kk = (K); goto slow_path; } \
    osym = (osym >> 8) | (s_ << 24); \
    code = __funnelshift_l((uint32_t)(in.W >> 32), code, sh_); in.W <<= sh_; in.ubits += sh_; low = lo2_ << sh_; range = rs_; \
    LPS_DEC_STORE(i0 + (K)); } while (0)
  const uint32_t n4 = n & ~3u;
  uint32_t nmax = n4;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL_MASK, nmax, o));
  for (uint32_t i0 = 0; i0 < nmax; i0 += 4) {
    if (i0 >= n4) continue;                                            // this lane's vector is done: idle
    uint32_t kk;
    if (ring && (i0 & 63) == 0 && i0) {                                // every 16 words: publish, and wait if the walker lags a ring behind
      __threadfence_block();
      rg->prod = i0 >> 2;
      while (!rg->dead && (i0 >> 2) - rg->cons > RING_WORDS - 64) { }
    }
    LPS_DEC_FAST(0); LPS_DEC_FAST(1); LPS_DEC_FAST(2); LPS_DEC_FAST(3);
    in.advance();
    continue;
  slow_path:                                                           // this lane finishes the batch on the generic path and meets the others at the loop end
    for (uint32_t k = kk; k < 4; k++) LPS_DEC_GENERIC(i0 + k);
    in.advance();
  }
  if (live) {
    for (uint32_t i = n4; i < n; i++) LPS_DEC_GENERIC(i);
    in.advance();
    if (n & 3) {
      const uint32_t rem = n & 3; osym >>= 8 * (4 - rem);
      for (uint32_t k = 0; k < rem; k++) out8[(n & ~3u) + k] = (uint8_t)(osym >> (8 * k));
      if (ring) rg->ring[(n >> 2) & (RING_WORDS - 1)] = osym;
    }
    if (TREE) {
      __threadfence_block();
      rg->prod = (n + 3) >> 2; rg->done = 1;
      const uint64_t end = f.tree_pos + 1028 + in.consumed;
      if (end > f.in_len) { atomicOr(&f.error, FERR_BAD_STREAM); f.B = 0; f.V = 0; rg->dead = 1; }
      else { f.tree_end = end; f.tree_ok = 1; }
    } else {
      const uint64_t end = f.spec_pos + 8 + 1028 + in.consumed;
      if (end == f.in_len) { f.spec_coded = 1028 + in.consumed; f.spec_state = 1; }    // the colour layer is the last one: it must end with the stream
    }
  }
#undef LPS_DEC_FAST
#undef LPS_DEC_STORE
#undef LPS_DEC_GENERIC
#undef LPS_DEC_LOOKUP
}

// ---- JPEG entropy stage of the speculated colour layer: grid (frames), 32 threads
__global__ void __launch_bounds__(32) dec_jpeg_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.x];
  if (!f.head_ok || f.error || f.spec_state != 1 || f.cct != 1) return;
  __shared__ HuffDec hd[4];
  const uint32_t lane = lane_id();
  uint32_t jerr = 0;
  if (lane == 0) { f.ncol = f.spec_ncol; jerr = jpeg_parse_header(f); }       // the JPEG stages read ncol
  jerr = __shfl_sync(FULL_MASK, jerr, 0);
  __syncwarp();
  if (!jerr) { warp_destuff(f); __syncwarp(); if (lane == 0) { jpeg_huff_decode(f, hd); f.huff_done = 1; } }
  if (lane == 0) f.spec_jerr = jerr;
}

// ---- centroid layer, colour verification / fallback, totals: grid (frames), 32 threads
__global__ void __launch_bounds__(32) dec_finish_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.x];
  if (!f.head_ok || f.error) return;
  __shared__ uint32_t freq[257];
  const uint32_t lane = lane_id();
  const uint8_t *in = f.in; const uint64_t len = f.in_len;
  uint32_t err = 0;
  bool ok = f.tree_ok != 0;
  uint64_t pos = f.tree_end;
  uint64_t coded[3] = { ok ? pos - f.tree_pos : 0, 0, 0 };
  uint32_t ncen = 0, ncol = 0;
  const uint32_t data_with_color = f.data_with_color, cct = f.cct;
  if (ok && f.do_centroid) {
    if (pos + 4 > len) ok = false;
    else {
      ncen = ld_u32_unaligned(in + pos); pos += 4;
      if (ncen > f.cen_cap) { ok = false; err |= FERR_TREE_CAP; }
      else ok = rc_decode_layer<false>(in, len, pos, f.cen, ncen, freq, &coded[1]);
    }
  }
  bool spec_used = false;
  if (ok && data_with_color) {
    if (f.spec_state == 1 && f.spec_pos == pos) {          // speculation confirmed: the layer starts exactly where we are
      ncol = f.spec_ncol; coded[2] = f.spec_coded; pos = len; spec_used = true;
      if (f.spec_jerr) { ok = false; err |= f.spec_jerr; }
    } else {
      if (f.spec_state == 1) {                              // mis-speculation (never observed): undo its side effects
        for (uint32_t k = lane; k < f.coef_cap_blocks * 32; k += 32) ((uint32_t *)f.coef)[k] = 0;
        if (lane == 0) f.huff_done = 0;
        __syncwarp();
      }
      if (pos + 8 > len) ok = false;
      else {
        uint64_t nc = ld_u64_unaligned(in + pos); pos += 8;
        if (nc > f.col_cap) { ok = false; err |= FERR_JPEG_CAP; }
        else { ncol = (uint32_t)nc; ok = rc_decode_layer<false>(in, len, pos, f.col, ncol, freq, &coded[2]); }
      }
    }
  }
  // trailing bytes switch the reference into detail mode (impl.hpp:1802-1806): the enhancement vectors follow
  __syncwarp();
  if (ok && pos != len) { ok = decode_detail_layers(f, in, len, pos, freq, err); if (ok && pos != len) ok = false; }
  if (lane == 0) {
    if (!ok) { atomicOr(&f.error, err ? err : FERR_BAD_STREAM); f.B = 0; f.V = 0; }
    else {
      f.B = f.tree_n; f.ncen = ncen; f.ncol = ncol; f.coded[0] = coded[0]; f.coded[1] = coded[1]; f.coded[2] = coded[2];
      if (data_with_color && cct == 1 && !spec_used) { const uint32_t je = jpeg_parse_header(f); if (je) { atomicOr(&f.error, je); f.B = 0; f.V = 0; } }
    }
  }
}

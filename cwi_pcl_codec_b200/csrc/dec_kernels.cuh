// dec_kernels.cuh -- decode side: header parse + range decode (serial, one warp per frame), DFS walk of the
// occupancy bytes ([PCL] Octree2BufBase::deserializeTree, impl.hpp:278), JPEG Huffman decode + IDCT + fancy
// upsampling + colour ([libjpeg] defaults, jpeg_io.hpp:140-162), inverse snake (snake.h:123-137) and point
// materialisation (deserializeTreeCallback impl.hpp:1584-1653, [PCL] ColorCoding::decodePoints).
#pragma once
#include "common.cuh"
#include "entropy_kernels.cuh"
#include "jpeg_enc_kernels.cuh"

__device__ __forceinline__ double ld_f64_unaligned(const uint8_t *p) { uint64_t v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | p[k]; return __longlong_as_double((long long)v); }
__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t *p) { uint64_t v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | p[k]; return v; }
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

__device__ inline uint32_t jpeg_parse_header(DecFrame &f);
struct HuffDec { int mincode[17]; int maxcode[17]; int valptr[17]; uint8_t vals[256]; uint16_t look[512]; };   // look: (len << 8) | sym for codes <= 9 bits
__device__ inline void jpeg_huff_decode(DecFrame &f, HuffDec *hd);
__device__ inline void warp_destuff(DecFrame &f);

__device__ inline void dfs_walk_ring(DecFrame &f, WalkRing *rg, const uint32_t *lut, uint32_t *stack);
__device__ inline bool decode_detail_layers(DecFrame &f, const uint8_t *in, uint64_t len, uint64_t &pos, uint32_t *freq_s, uint32_t &err);

// ---- stage 1: header + entropy decoding of the layers (impl.hpp:231-261, 1766-1835), one block per frame (steered):
// warp 0 parses the header and range-decodes tree -> [centroid] -> colour (serially dependent: no stored lengths);
// warp 1 walks the occupancy bytes out of a shared-memory ring while warp 0 is still producing them.
// warp 2 decodes the colour layer speculatively while warp 0 is still busy with the tree layer: the colour layer is the
// last layer of the frame and starts with an unmistakable signature (u64 size below 2^32, then 257 strictly increasing
// u32 below 2^16 starting at 0), so its offset can be found by scanning backwards from the end; warp 0 accepts the
// result only if its own position after the preceding layers lands exactly there, otherwise it decodes the layer
// itself as the reference would.
struct SpecColour { volatile uint32_t state; uint32_t pos, ncol, coded, jerr; uint32_t with_color, cct; };
__global__ void __launch_bounds__(96) dec_entropy_kernel(DecFrame *frames, int use_ring) {
  DecFrame &f = frames[blockIdx.x];
  if (threadIdx.x == 0) f.serial_sm = sm_id();
  __shared__ uint32_t freq[257], freq2[257];
  __shared__ WalkRing rg;
  __shared__ SpecColour sc;
  __shared__ HuffDec hd[4];
  __shared__ uint32_t lut[256];                            // per 8-bit child mask: popcount << 16 | index of the lowest bit << 8 | mask without its lowest bit
  __shared__ uint32_t wstack[32];                          // walker: remaining-children mask per open level
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = ((uint32_t)__popc(i) << 16) | ((i ? (uint32_t)(__ffs(i) - 1) : 0u) << 8) | (i & (i - 1));
  const uint32_t lane = lane_id();
  const bool decoder = threadIdx.x < 32;
  const uint32_t role = threadIdx.x >> 5;                  // 0 range decoder, 1 DFS walker, 2 speculative colour layer
  const uint8_t *in = f.in;
  const uint64_t len = f.in_len;
  uint32_t err = f.error;
  uint64_t pos = 0, B = 0;
  uint32_t do_centroid = 0, data_with_color = 0, cct = 0, depth = 0;
  if (threadIdx.x == 0) { rg.prod = 0; rg.cons = 0; rg.done = 0; rg.dead = 0; rg.go = 0; sc.state = 0; sc.with_color = 0; sc.cct = 0; sc.jerr = 0; }
  if (decoder && !err) {
    // syncToHeader: scan for the two magics like the reference (impl.hpp:1660-1676)
    const char id2[] = "<PCL-OCT-CODECV2-COMPRESSED>", id1[] = "<PCL-OCT-COMPRESSED>";
    uint32_t hp = 0; bool ok = true;
    while (hp < 28) {
      if (pos >= len) { ok = false; break; }
      uint8_t c = in[pos++];
      if (c == 0xFF) { ok = false; break; }               // (char)0xFF == EOF quirk, SURVEY App. C-9
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
      const uint32_t frame_id = ld_u32_unaligned(h);
      data_with_color = h[6];
      const uint64_t point_count = ld_u64_unaligned(h + 7);
      const double res = ld_f64_unaligned(h + 15);
      const uint8_t color_bits = h[23];
      const double point_res = ld_f64_unaligned(h + 24);
      double bmin[3], bmax[3];
      for (int a = 0; a < 3; a++) { bmin[a] = ld_f64_unaligned(h + 32 + 8 * a); bmax[a] = ld_f64_unaligned(h + 56 + 8 * a); }
      do_centroid = h[80];
      cct = ld_u32_unaligned(h + 83);
      pos += 92;
      // [PCL] readFrameHeader -> defineBoundingBox -> getKeyBitSize
      const double eps = 1.1920928955078125e-07;
      uint32_t mk = 2;
      for (int a = 0; a < 3; a++) {
        double t = ceil(__ddiv_rn(__dsub_rn(__dsub_rn(bmax[a], bmin[a]), eps), res));
        uint32_t k = (t >= 4294967295.0 || !(t == t)) ? 0xFFFFFFFFu : (t > 0 ? (uint32_t)t : 0u);
        if (k > mk) mk = k;
      }
      while ((1ull << depth) < mk) depth++;
      if (depth > CCV2_MAX_DEPTH || !(res > 0)) err |= FERR_DEPTH;
      if (point_count > f.out_cap) err |= FERR_OUT_CAP;
      B = ld_u64_unaligned(in + pos); pos += 8;
      if (B > f.tree_cap) err |= FERR_TREE_CAP;
      if (lane == 0) {
        f.frame_id = frame_id; f.data_with_color = data_with_color; f.point_count = point_count; f.res = res; f.color_bits = color_bits;
        for (int a = 0; a < 3; a++) { f.bmin[a] = bmin[a]; f.bmax[a] = bmax[a]; }
        f.do_centroid = do_centroid; f.cct = cct; f.depth = depth; f.point_res_f = (float)point_res;
      }
    }
    if (lane == 0) { rg.B = (uint32_t)B; rg.depth = depth; rg.go = (use_ring && !err && B > 0 && depth >= 1 && depth <= 17) ? 1u : 0u;
                     sc.with_color = (use_ring && !err) ? data_with_color : 0u; sc.cct = cct; }
  }
  __syncthreads();
  const bool ring = rg.go != 0;
  if (role == 1) {                                          // walker warp: lane 0 walks, the other lanes retire
    if (lane == 0 && ring) dfs_walk_ring(f, &rg, lut, wstack);
    return;
  }
  if (role == 2) {                                          // speculative colour layer (+ JPEG entropy decode)
    uint32_t st = 2;
    if (sc.with_color && len > FRAME_HDR_BYTES + 8 + 1032 + 8 + 1032) {
      // backward scan for the layer signature; candidates o = offset of the u64 size word
      const uint64_t lo_lim = FRAME_HDR_BYTES + 8 + 1032;
      uint64_t found = 0; bool have = false;
      for (uint64_t hi = len - 1040; !have && hi >= lo_lim; hi = hi >= 32 + lo_lim ? hi - 32 : 0) {
        const uint64_t o = hi >= lane ? hi - lane : 0;
        bool q = o >= lo_lim && (in[o + 4] | in[o + 5] | in[o + 6] | in[o + 7] | in[o + 8] | in[o + 9] | in[o + 10] | in[o + 11] | in[o + 14] | in[o + 15]) == 0;
        uint32_t cand = __ballot_sync(FULL_MASK, q);
        while (cand && !have) {                             // validate candidates from the highest offset down
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
      if (have) {
        const uint64_t nc = ld_u64_unaligned(in + found);
        uint64_t p2 = found + 8, cd = 0;
        if (nc > 0 && nc <= f.col_cap && rc_decode_layer<false>(in, len, p2, f.col, (uint32_t)nc, freq2, &cd) && p2 == len) {
          uint32_t jerr = 0;
          if (sc.cct == 1) {
            f.ncol = (uint32_t)nc;                          // the JPEG stages read it
            if (lane == 0) jerr = jpeg_parse_header(f);
            jerr = __shfl_sync(FULL_MASK, jerr, 0);
            __syncwarp();
            if (!jerr) { warp_destuff(f); __syncwarp(); if (lane == 0) { jpeg_huff_decode(f, hd); f.huff_done = 1; } }
          }
          if (lane == 0) { sc.pos = (uint32_t)found; sc.ncol = (uint32_t)nc; sc.coded = (uint32_t)cd; sc.jerr = jerr; }
          st = 1;
        }
      }
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); sc.state = st; }
    return;
  }
  if (err) { if (lane == 0) { f.error |= err; f.V = 0; f.B = 0; f.point_count = 0; rg.dead = 1; rg.done = 1; } return; }
  uint64_t coded[3] = { 0, 0, 0 };
  bool ok = ring ? rc_decode_layer<true>(in, len, pos, f.tree, (uint32_t)B, freq, &coded[0], &rg)
                 : rc_decode_layer<false>(in, len, pos, f.tree, (uint32_t)B, freq, &coded[0]);
  if (!ok && lane == 0) { rg.dead = 1; rg.done = 1; }      // releases a waiting walker
  uint32_t ncen = 0, ncol = 0;
  if (ok && do_centroid) {
    if (pos + 4 > len) ok = false;
    else {
      ncen = ld_u32_unaligned(in + pos); pos += 4;
      if (ncen > f.cen_cap) { ok = false; err |= FERR_TREE_CAP; }
      else ok = rc_decode_layer<false>(in, len, pos, f.cen, ncen, freq, &coded[1]);
    }
  }
  bool spec_used = false;
  if (ok && data_with_color) {
    while (sc.with_color && sc.state == 0) { }             // warp 2 is normally long done
    if (sc.with_color && sc.state == 1 && sc.pos == pos) {  // speculation confirmed: the layer starts exactly where we are
      ncol = sc.ncol; coded[2] = sc.coded; pos = len; spec_used = true;
      if (sc.jerr) { ok = false; err |= sc.jerr; }
    } else {
      if (sc.with_color && sc.state == 1) {                 // mis-speculation (never observed): undo its side effects
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
      f.B = (uint32_t)B; f.ncen = ncen; f.ncol = ncol; f.coded[0] = coded[0]; f.coded[1] = coded[1]; f.coded[2] = coded[2];
      if (data_with_color && cct == 1 && !spec_used) { const uint32_t je = jpeg_parse_header(f); if (je) { atomicOr(&f.error, je); f.B = 0; f.V = 0; } }
    }
  }
}

// ---- stage 2a: DFS walk of the occupancy bytes. Emits one (prefix, byte) record per bottom-level branch
// (level depth-1); leaves are expanded in parallel afterwards.  Serial by nature (a node's position is known
// only after its left siblings' subtrees are consumed).  One thread; 16-byte loads.
struct SeqBytes {
  const uint8_t *p; uint32_t n, pos; uint4 cur; uint32_t cbase;
  __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t len) { p = ptr; n = len; pos = 0; cbase = NONE_U32; }
  __device__ __forceinline__ uint32_t next() {
    uint32_t b = pos & ~15u;
    if (b != cbase) { cur = *(const uint4 *)(p + b); cbase = b; }
    uint32_t w = (pos & 8) ? ((pos & 4) ? cur.w : cur.z) : ((pos & 4) ? cur.y : cur.x);
    uint32_t v = (w >> (8 * (pos & 3))) & 255;
    pos++;
    return v;
  }
};
// Fast walk for depth <= 17: the child masks of the open branches live in two 64-bit registers (8 bits per level),
// the stream is consumed through a 64-bit window, and the children of a level depth-2 branch (all bottom-level
// nodes, contiguous in the stream) are drained in a tight loop.
__device__ inline void dfs_walk_fast(DecFrame &f) {
  const uint32_t B = f.B, d = f.depth;
  const uint64_t *src = (const uint64_t *)f.tree;          // 8-byte aligned, padded
  uint32_t pos = 0; uint64_t win = src[0]; uint32_t wleft = 8; uint64_t nextw = src[1]; uint32_t wi = 2;
  const uint32_t nwords = (B + 7) / 8 + 1;
#define NEXT_BYTE(dst) do { dst = (uint32_t)win & 255u; win >>= 8; pos++; if (--wleft == 0) { win = nextw; wleft = 8; nextw = wi < nwords ? src[wi] : 0; wi++; } } while (0)
  uint32_t nb = 0; const uint32_t cap = f.node_cap;
  uint64_t *np = f.node_prefix; uint8_t *nby = f.node_byte;
  bool bad = false;
  uint32_t m; NEXT_BYTE(m);
  if (d == 1) { np[0] = 0; nby[0] = (uint8_t)m; nb = 1; }
  else {
    uint64_t s0 = 0, s1 = 0, prefix = 0; uint32_t L = 0;
    for (;;) {
      if (m == 0) {
        if (L == 0) break;
        L--; prefix >>= 3;
        m = (uint32_t)((L < 8 ? s0 >> (8 * L) : s1 >> (8 * (L - 8))) & 255u);
        continue;
      }
      if (L + 2 == d) {                                   // children are bottom-level branches: one byte each
        const uint32_t k = __popc(m);
        if (pos + k > B || nb + k > cap) { bad = true; break; }
        const uint64_t pre = prefix << 3;
        do {
          const uint32_t c = __ffs(m) - 1; m &= m - 1;
          uint32_t byte; NEXT_BYTE(byte);
          np[nb] = pre | c; nby[nb] = (uint8_t)byte; nb++;
        } while (m);
        continue;                                          // m == 0: pop
      }
      const uint32_t c = __ffs(m) - 1; m &= m - 1;
      if (pos >= B) { bad = true; break; }
      if (L < 8) { const uint32_t sh = 8 * L; s0 = (s0 & ~(255ull << sh)) | ((uint64_t)m << sh); }
      else { const uint32_t sh = 8 * (L - 8); s1 = (s1 & ~(255ull << sh)) | ((uint64_t)m << sh); }
      prefix = (prefix << 3) | c; L++;
      NEXT_BYTE(m);
    }
  }
#undef NEXT_BYTE
  if (pos != B) bad = true;
  if (bad) { f.error |= FERR_BAD_STREAM; nb = 0; }
  f.n_bottom = nb;
}

// The pipelined walker (warp 1 of dec_entropy_kernel): reads the occupancy bytes out of the shared-memory ring
// while the range decoder is still producing them.  It only has to visit the branches above the bottom level: for
// a branch at level depth-2 it records (prefix, child mask, stream offset of the first child) and skips the child
// bytes -- dec_leaves_kernel turns those records into points in parallel afterwards.
#ifndef WALK_NAP_NS
#define WALK_NAP_NS 2048
#endif
__device__ inline void dfs_walk_ring(DecFrame &f, WalkRing *rg, const uint32_t *lut, uint32_t *stack) {
  const uint32_t B = rg->B, d = rg->depth;
  const uint32_t rg_a = smem_addr(rg), lut_a = smem_addr(lut), st_a = smem_addr(stack);
  const uint32_t ring_a = rg_a + (uint32_t)offsetof(WalkRing, ring), prod_a = rg_a + (uint32_t)offsetof(WalkRing, prod),
                 cons_a = rg_a + (uint32_t)offsetof(WalkRing, cons), dead_a = rg_a + (uint32_t)offsetof(WalkRing, dead);
  uint32_t pos = 0, cw = NONE_U32, win = 0, avail = 0, pub = 0;
  bool bad = false;
  auto read_byte = [&](uint32_t &dst) -> bool {             // byte at stream offset pos; false when the producer died
    const uint32_t wi = pos >> 2;
    if (wi != cw) {
      // the decoder publishes every 64 symbols (8-15 us) into a ring 2048 symbols deep: sleep through most of that rather than
      // take issue slots from the decoders on this SM (at 256 ns per nap the wake-ups alone were ~11 M instructions per frame)
      while (wi >= avail) { avail = lds_volatile_u32(prod_a); if (lds_volatile_u32(dead_a)) return false; if (wi >= avail) __nanosleep(WALK_NAP_NS); }
      win = lds_volatile_u32(ring_a + ((wi & (RING_WORDS - 1)) << 2)); cw = wi;
      if ((wi >> 4) != pub) { pub = wi >> 4; sts_volatile_u32(cons_a, wi); }
    }
    dst = __byte_perm(win, 0, 0x4440u | (pos & 3u)); pos++;
    return true;
  };
  uint32_t m = 0, n2 = 0, nb = 0;
  const uint32_t cap = f.node_cap;
  uint64_t *const l2p = f.l2_prefix; uint8_t *const l2m = f.l2_mask; uint32_t *const l2o = f.l2_off;
  if (!read_byte(m)) bad = true;
  else if (d == 1) { f.node_prefix[0] = 0; f.node_byte[0] = (uint8_t)m; nb = 1; }
  else {
    uint64_t prefix = 0; uint32_t L = 0;
    const uint32_t Lb = d - 2;                              // branches at this level have bottom-level children
    for (;;) {
      if (L == Lb) {                                        // record (prefix, mask, offset of the first child) and skip the children
        const uint32_t k = lds_u32(lut_a + (m << 2)) >> 16;
        if (pos + k > B || n2 >= cap) { bad = true; break; }
        l2p[n2] = prefix; l2m[n2] = (uint8_t)m; l2o[n2] = pos; n2++;
        pos += k; m = 0;
      }
      while (m == 0 && L) { L--; prefix >>= 3; m = lds_u32(st_a + (L << 2)); }   // pop exhausted branches
      if (m == 0) break;                                     // the root is exhausted: done
      const uint32_t e = lds_u32(lut_a + (m << 2));          // descend into the next child
      sts_u32(st_a + (L << 2), e & 255u); prefix = (prefix << 3) | ((e >> 8) & 7u); L++;
      if (pos >= B) { bad = true; break; }
      if (!read_byte(m)) { bad = true; break; }
    }
  }
  sts_volatile_u32(dead_a, 1u);                             // the decoder must never wait for a walker that has left
  if (pos != B) bad = true;
  if (bad) { atomicOr(&f.error, FERR_BAD_STREAM); nb = 0; n2 = 0; }
  f.n_bottom = nb; f.n_l2 = n2; f.l2_valid = (d >= 2 && !bad) ? 1u : 0u;
  f.walk_done = 1;
}

__device__ inline void dfs_walk(DecFrame &f) {
  const uint32_t B = f.B, d = f.depth;
  if (B == 0 || d == 0) { f.n_bottom = 0; return; }
  SeqBytes sb; sb.init(f.tree, B);
  uint32_t nb = 0; const uint32_t cap = f.node_cap;
  uint64_t m0 = 0, m1 = 0, m2 = 0;                       // child masks of the open branches, 8 bits per level
  auto getm = [&](uint32_t l) -> uint32_t { uint64_t w = l < 8 ? m0 : (l < 16 ? m1 : m2); return (uint32_t)(w >> (8 * (l & 7))) & 255u; };
  auto setm = [&](uint32_t l, uint32_t v) { uint64_t sh = 8 * (l & 7), msk = ~(255ull << sh), nv = (uint64_t)v << sh;
    if (l < 8) m0 = (m0 & msk) | nv; else if (l < 16) m1 = (m1 & msk) | nv; else m2 = (m2 & msk) | nv; };
  bool bad = false;
  if (d == 1) {
    if (nb < cap) { f.node_prefix[0] = 0; f.node_byte[0] = (uint8_t)sb.next(); nb = 1; }
  } else {
    uint32_t level = 0; uint64_t prefix = 0;
    setm(0, sb.next());
    for (;;) {
      uint32_t m = getm(level);
      if (m == 0) { if (level == 0) break; level--; prefix >>= 3; continue; }
      uint32_t c = __ffs(m) - 1;
      setm(level, m & (m - 1));
      if (sb.pos >= B) { bad = true; break; }
      uint32_t byte = sb.next();
      uint64_t child = (prefix << 3) | c;
      if (level + 2 < d) { level++; prefix = child; setm(level, byte); }
      else {                                             // child is a bottom-level branch: record it, do not descend
        if (nb >= cap) { bad = true; break; }
        f.node_prefix[nb] = child; f.node_byte[nb] = (uint8_t)byte; nb++;
      }
    }
  }
  if (sb.pos != B) bad = true;                            // [PCL] would simply stop; a well-formed frame consumes every byte
  if (bad) { f.error |= FERR_BAD_STREAM; nb = 0; }
  f.n_bottom = nb;
}

// ---- stage 2b: JPEG of the single SNAKE image: marker parse (serial, short), parallel de-stuffing of the
// entropy-coded segment, serial Huffman decode (no restart markers => one dependent bit stream per image)
struct JpegInfo { uint32_t w, h, scan, dht_off[4], dht_n[4]; };
// marker parse of one baseline JFIF file as libjpeg writes it for jpeg_io's settings (3 components, 2x2 / 1x1 / 1x1,
// no restart markers).  qt receives the two quantisation tables in zigzag order.  Returns false on anything else.
__device__ inline bool jpeg_parse(const uint8_t *in, uint32_t len, JpegInfo &o, uint16_t *qt) {
  bool bad = false;
  uint32_t w = 0, h = 0, scan = 0, have = 0;
  if (len < 4 || in[0] != 0xFF || in[1] != 0xD8) bad = true;
  uint32_t pos = 2;
  while (!bad && pos + 4 <= len) {
    if (in[pos] != 0xFF) { bad = true; break; }
    uint32_t m = in[pos + 1], L = ((uint32_t)in[pos + 2] << 8) | in[pos + 3];
    const uint8_t *s = in + pos + 4;
    if (L < 2 || pos + 2 + L > len) { bad = true; break; }
    if (m == 0xDB) {
      uint32_t q = 0;
      while (q + 65 <= L - 2) { uint32_t t = s[q] & 15; if ((s[q] >> 4) || t > 1) { bad = true; break; } for (int i = 0; i < 64; i++) qt[t * 64 + i] = s[q + 1 + i]; have |= 1u << t; q += 65; }
    } else if (m == 0xC0) {
      if (L < 17) { bad = true; break; }
      h = ((uint32_t)s[1] << 8) | s[2]; w = ((uint32_t)s[3] << 8) | s[4];
      if (s[0] != 8 || s[5] != 3 || s[7] != 0x22 || s[10] != 0x11 || s[13] != 0x11 || s[8] != 0 || s[11] != 1 || s[14] != 1) bad = true;
    } else if (m == 0xC4) {
      uint32_t q = 0;
      while (q + 17 <= L - 2) {
        uint32_t tc = s[q] >> 4, th = s[q] & 15, nv = 0;
        if (tc > 1 || th > 1) { bad = true; break; }
        for (int i = 0; i < 16; i++) nv += s[q + 1 + i];
        if (nv > 256 || q + 17 + nv > L - 2) { bad = true; break; }
        o.dht_off[tc * 2 + th] = pos + 4 + q + 1; o.dht_n[tc * 2 + th] = nv;
        have |= 4u << (tc * 2 + th);
        q += 17 + nv;
      }
    } else if (m == 0xDA) { scan = pos + 2 + L; break; }
    else if (m == 0xC2 || m == 0xDD) { bad = true; break; }      // progressive / restart intervals: libjpeg as driven by jpeg_io never emits them
    pos += 2 + L;
  }
  if (!scan || w == 0 || h == 0 || have != 0x3F) bad = true;
  o.w = w; o.h = h; o.scan = scan;
  return !bad;
}
__device__ inline uint32_t jpeg_parse_header(DecFrame &f) {   // SNAKE image of the frame; returns FERR_* bits (0 = ok)
  JpegInfo ji;
  bool ok = jpeg_parse(f.col, f.ncol, ji, f.qt);
  if (ok && ji.w != 256) ok = false;                             // SNAKE images are 256 wide (cjpeg.h:197)
  const uint32_t mcu_w = 16, mcu_h = ok ? (ji.h + 15) / 16 : 0, nblocks = mcu_w * mcu_h * 6;
  uint32_t eb = 0;
  if (ok && nblocks > f.coef_cap_blocks) { ok = false; eb |= FERR_JPEG_CAP; }
  if (!ok) { f.img_w = f.img_h = f.mcu_w = f.mcu_h = f.n_blocks = 0; f.scan_start = f.scan_len = 0; return eb | FERR_BAD_STREAM; }
  for (int t = 0; t < 4; t++) { f.dht_off[t] = ji.dht_off[t]; f.dht_n[t] = ji.dht_n[t]; }
  f.img_w = ji.w; f.img_h = ji.h; f.mcu_w = mcu_w; f.mcu_h = mcu_h; f.n_blocks = nblocks; f.scan_start = ji.scan;
  return 0;
}

// removes the 0x00 stuffed after every 0xFF of the entropy-coded segment; one CTA per frame, chunks in order
__global__ void __launch_bounds__(1024) jpeg_destuff_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.x];
  if (f.error || f.huff_done || !f.data_with_color || f.cct != 1 || f.n_blocks == 0) return;
  const uint8_t *in = f.col + f.scan_start;
  uint32_t n = f.ncol - f.scan_start;
  if (n >= 2 && in[n - 2] == 0xFF && in[n - 1] == 0xD9) n -= 2;     // EOI
  __shared__ uint64_t s_scan[33];
  __shared__ uint8_t s_last[1024];
  uint32_t obase = 0; uint32_t carry = 0;                            // last byte of the previous chunk
  for (uint32_t c0 = 0; c0 < n; c0 += 1024 * 16) {
    const uint32_t b0 = c0 + threadIdx.x * 16;
    uint8_t by[16]; uint32_t nv = 0;
    if (b0 < n) { nv = min(16u, n - b0); for (uint32_t k = 0; k < nv; k++) by[k] = in[b0 + k]; }
    s_last[threadIdx.x] = nv ? by[nv - 1] : 0;
    __syncthreads();
    uint32_t prev = threadIdx.x ? s_last[threadIdx.x - 1] : carry;
    const uint32_t chunk_last = s_last[1023];
    uint32_t keep = 0, cnt = 0;
    for (uint32_t k = 0; k < nv; k++) { bool kp = !(by[k] == 0 && prev == 0xFF); keep |= (uint32_t)kp << k; cnt += kp; prev = by[k]; }
    uint64_t tot;
    uint64_t excl = block_excl_scan_u64(cnt, &tot, s_scan);
    uint32_t o = obase + (uint32_t)excl;
    for (uint32_t k = 0; k < nv; k++) if (keep >> k & 1) f.scan[o++] = by[k];
    obase += (uint32_t)tot; carry = chunk_last;
    __syncthreads();
  }
  if (threadIdx.x == 0) f.scan_len = obase;
}

// de-stuffing by one warp (speculative colour path inside dec_entropy_kernel): 4 bytes per lane per step
__device__ inline uint32_t warp_destuff_span(const uint8_t *in, uint32_t n, uint8_t *out);
__device__ inline void warp_destuff(DecFrame &f) {
  const uint32_t len = warp_destuff_span(f.col + f.scan_start, f.ncol - f.scan_start, f.scan);
  if (lane_id() == 0) f.scan_len = len;
}
__device__ inline uint32_t warp_destuff_span(const uint8_t *in, uint32_t n, uint8_t *out) {
  const uint32_t lane = lane_id();
  if (n >= 2 && in[n - 2] == 0xFF && in[n - 1] == 0xD9) n -= 2;     // EOI
  uint32_t obase = 0, carry = 0;
  for (uint32_t c0 = 0; c0 < n; c0 += 128) {
    const uint32_t b0 = c0 + 4 * lane;
    uint32_t by[4] = { 0, 0, 0, 0 }, nv = 0;
    if (b0 < n) { nv = min(4u, n - b0); for (uint32_t k = 0; k < nv; k++) by[k] = in[b0 + k]; }
    uint32_t prev = __shfl_up_sync(FULL_MASK, by[3], 1);
    if (lane == 0) prev = carry;
    uint32_t keep = 0, cnt = 0;
    for (uint32_t k = 0; k < nv; k++) { const bool kp = !(by[k] == 0 && prev == 0xFF); keep |= (uint32_t)kp << k; cnt += kp; prev = by[k]; }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= (uint32_t)o) inc += t; }
    uint32_t o = obase + inc - cnt;
    for (uint32_t k = 0; k < nv; k++) if (keep >> k & 1) out[o++] = (uint8_t)by[k];
    obase += __shfl_sync(FULL_MASK, inc, 31);
    carry = __shfl_sync(FULL_MASK, by[3], 31);
  }
  return obase;
}

struct JBits {               // MSB-first bit reader over the de-stuffed segment, 32-bit big-endian refills, next word prefetched
  const uint32_t *w; uint32_t nwords, wi; uint64_t acc; uint32_t nb; uint32_t nxt;
  __device__ __forceinline__ uint32_t fetch(uint32_t i) const { return i < nwords ? __byte_perm(w[i], 0, 0x0123) : 0u; }
  __device__ __forceinline__ void init(const uint8_t *p, uint32_t nbytes) {
    w = (const uint32_t *)p; nwords = (nbytes + 3) / 4;
    acc = ((uint64_t)fetch(0) << 32) | fetch(1); nb = 64; nxt = fetch(2); wi = 3;
  }
  __device__ __forceinline__ uint32_t peek(uint32_t k) const { return (uint32_t)(acc >> (64 - k)); }          // 1 <= k <= 32
  __device__ __forceinline__ void skip(uint32_t k) { acc <<= k; nb -= k; if (nb <= 32) { acc |= (uint64_t)nxt << (32 - nb); nb += 32; nxt = fetch(wi); wi++; } }
};
__device__ inline int huff_sym(JBits &b, const HuffDec &h) {
  uint32_t e = h.look[b.peek(9)];
  if (e) { b.skip(e >> 8); return e & 255; }
  for (int l = 10; l <= 16; l++) {
    int code = (int)b.peek(l);
    if (code <= h.maxcode[l]) { b.skip(l); return h.vals[(h.valptr[l] + code - h.mincode[l]) & 255]; }
  }
  b.skip(16);
  return 0;
}
__device__ inline void huff_build(HuffDec &h, const uint8_t *bits, const uint8_t *vals, int nvals) {
  for (int k = 0; k < 256; k++) h.vals[k] = k < nvals ? vals[k] : 0;
  for (int k = 0; k < 512; k++) h.look[k] = 0;
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    int cnt = bits[l - 1];
    h.mincode[l] = code; h.valptr[l] = k;
    h.maxcode[l] = cnt ? code + cnt - 1 : -1;
    if (l <= 9) for (int c = 0; c < cnt; c++) {
      int cc = code + c;
      for (int fv = 0; fv < (1 << (9 - l)); fv++) h.look[((cc << (9 - l)) | fv) & 511] = (uint16_t)((l << 8) | vals[k + c]);
    }
    code += cnt; k += cnt; code <<= 1;
  }
}
__device__ __forceinline__ int jextend(int v, int n) { return n == 0 ? 0 : (v < (1 << (n - 1)) ? v - (1 << n) + 1 : v); }

__device__ inline void huff_decode_blocks(const uint8_t *scan, uint32_t scan_len, uint32_t nblocks, short *coef, const HuffDec *hd);
__device__ inline void jpeg_huff_decode(DecFrame &f, HuffDec *hd /* smem[4]: dc0 dc1 ac0 ac1 */) {
  const uint32_t nblocks = f.n_blocks;
  if (nblocks == 0) return;
  for (int t = 0; t < 4; t++) huff_build(hd[t], f.col + f.dht_off[t], f.col + f.dht_off[t] + 16, (int)f.dht_n[t]);
  huff_decode_blocks(f.scan, f.scan_len, nblocks, f.coef, hd);
}
// serial Huffman decode of `nblocks` blocks in MCU order (Y0 Y1 Y2 Y3 Cb Cr); coef is pre-zeroed, zigzag order, quantised
__device__ inline void huff_decode_blocks(const uint8_t *scan, uint32_t scan_len, uint32_t nblocks, short *coef, const HuffDec *hd) {
  JBits br; br.init(scan, scan_len);
  int pred[3] = { 0, 0, 0 };
  uint32_t blk = 0;
  for (uint32_t g = 0; g < nblocks; g++) {
    const int comp = blk < 4 ? 0 : (int)blk - 3; const int ts = comp ? 1 : 0;
    short *c = coef + (size_t)g * 64;
    int n = huff_sym(br, hd[ts]);
    int diff = 0;
    if (n) { diff = jextend((int)br.peek(n), n); br.skip(n); }
    pred[comp] += diff; c[0] = (short)pred[comp];
    for (int k = 1; k < 64; k++) {
      int rs = huff_sym(br, hd[2 + ts]), r = rs >> 4, s = rs & 15;
      if (s == 0) { if (r == 15) { k += 15; continue; } break; }
      k += r;
      int v = jextend((int)br.peek(s), s); br.skip(s);
      if (k > 63) break;
      c[k] = (short)v;
    }
    blk = blk == 5 ? 0 : blk + 1;
  }
}

__global__ void __launch_bounds__(64) dec_serial_kernel(DecFrame *frames, int first_slot, int group_frames) {
  const int fi = steered_frame(first_slot, group_frames);
  if (fi < 0) return;
  DecFrame &f = frames[fi];
  __shared__ HuffDec hd[4];
  if (lane_id() != 0) return;
  const int role = threadIdx.x >> 5;                       // warp 0: DFS walk, warp 1: JPEG Huffman decode
  if (f.error) { if (role == 0) f.n_bottom = 0; else f.n_blocks = 0; return; }
  if (role == 0) { if (f.walk_done) return; if (f.depth >= 1 && f.depth <= 17 && f.B > 0) dfs_walk_fast(f); else dfs_walk(f); }
  else if (f.data_with_color && f.cct == 1 && !f.huff_done) jpeg_huff_decode(f, hd);
}

// ---- stage 3: dequantise + ISLOW IDCT. 8 threads per block, 32 blocks per CTA; writes Y / Cb / Cr planes
__device__ __forceinline__ uint8_t jpeg_range_limit(int x) {
  int v = x & 1023;
  return v < 128 ? (uint8_t)(v + 128) : (v < 512 ? 255 : (v < 896 ? 0 : (uint8_t)(v - 896)));
}
__device__ __forceinline__ void idct8(const int *v, int *o, bool first) {
  int z2 = v[2], z3 = v[6], z1 = (z2 + z3) * 4433;
  int t2 = z1 - z3 * 15137, t3 = z1 + z2 * 6270;
  int t0 = (v[0] + v[4]) * 8192, t1 = (v[0] - v[4]) * 8192;
  int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  int a0 = v[7], a1 = v[5], a2 = v[3], a3 = v[1];
  z1 = a0 + a3; z2 = a1 + a2; z3 = a0 + a2; int z4 = a1 + a3, z5 = (z3 + z4) * 9633;
  a0 *= 2446; a1 *= 16819; a2 *= 25172; a3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
  a0 += z1 + z3; a1 += z2 + z4; a2 += z2 + z3; a3 += z1 + z4;
  const int n = first ? 11 : 18;
  o[0] = JDESCALE(t10 + a3, n); o[1] = JDESCALE(t11 + a2, n); o[2] = JDESCALE(t12 + a1, n); o[3] = JDESCALE(t13 + a0, n);
  o[4] = JDESCALE(t13 - a0, n); o[5] = JDESCALE(t12 - a1, n); o[6] = JDESCALE(t11 - a2, n); o[7] = JDESCALE(t10 - a3, n);
}
__device__ __forceinline__ void dec_plane_ptrs(const DecFrame &f, uint8_t *&Y, uint8_t *&Cb, uint8_t *&Cr, uint32_t &YW, uint32_t &CW) {
  YW = f.mcu_w * 16; CW = f.mcu_w * 8;
  const size_t ysz = (size_t)YW * f.mcu_h * 16, csz = (size_t)CW * f.mcu_h * 8;
  Y = f.planes; Cb = f.planes + ysz; Cr = Cb + csz;
}
__global__ void __launch_bounds__(256) jpeg_idct_kernel(DecFrame *frames, const JpegTables *T) {
  DecFrame &f = frames[blockIdx.y];
  const uint32_t nblocks = f.n_blocks;
  if (f.error || blockIdx.x * 32 >= nblocks) return;
  __shared__ int ws[32][64];
  __shared__ uint8_t zz[64];
  if (threadIdx.x < 64) zz[threadIdx.x] = T->zz[threadIdx.x];
  __syncthreads();
  const uint32_t lb = threadIdx.x >> 3, k = threadIdx.x & 7, g = blockIdx.x * 32 + lb;
  const bool act = g < nblocks;
  const uint32_t blk = g % 6, mcu = g / 6;
  if (act) {
    const short *c = f.coef + (size_t)g * 64;
    const uint16_t *q = f.qt + (blk >= 4 ? 64 : 0);
    // natural-order position (row r, col k) <- zigzag index: build inverse on the fly
    int v[8], o[8];
    // gather column k: natural index r*8 + k; find zigzag position by scanning the small table
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = 0;
    for (int z = 0; z < 64; z++) { uint32_t nat = zz[z]; if ((nat & 7) == k) v[nat >> 3] = (int)c[z] * (int)q[z]; }
    idct8(v, o, true);
#pragma unroll
    for (int r = 0; r < 8; r++) ws[lb][r * 8 + k] = o[r];
  }
  __syncthreads();
  if (act) {
    int v[8], o[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = ws[lb][k * 8 + c];
    idct8(v, o, false);
    uint8_t *Y, *Cb, *Cr; uint32_t YW, CW;
    dec_plane_ptrs(f, Y, Cb, Cr, YW, CW);
    const uint32_t mx = mcu % f.mcu_w, my = mcu / f.mcu_w;
    uint8_t *dst;
    if (blk < 4) dst = Y + (size_t)(my * 16 + (blk >> 1) * 8 + k) * YW + mx * 16 + (blk & 1) * 8;
    else dst = (blk == 4 ? Cb : Cr) + (size_t)(my * 8 + k) * CW + mx * 8;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) { lo |= (uint32_t)jpeg_range_limit(o[c]) << (8 * c); hi |= (uint32_t)jpeg_range_limit(o[4 + c]) << (8 * c); }
    *(uint2 *)dst = make_uint2(lo, hi);
  }
}

// ---- stage 4: expand bottom-level branches into leaves (chained scan of popcounts) and write the points
#define NODE_THREADS 256
__device__ __forceinline__ uint32_t dec_color_lines(const DecFrame &f, uint32_t i);
__device__ __forceinline__ uint32_t dec_color(const DecFrame &f, uint32_t i) {
  if (!f.data_with_color) return 0x00FFFFFFu;                 // [PCL] setDefaultColor (white, alpha 0)
  if (f.cct == 2) return dec_color_lines(f, i);
  if (f.cct != 1) {
    if (3ull * i + 2 >= f.ncol) return 0;
    const uint32_t red = f.cct == 0 ? 8 - f.color_bits : 0;
    const uint8_t *c = f.col + 3ull * i;
    return (((uint32_t)c[0] << red) & 255) | ((((uint32_t)c[1] << red) & 255) << 8) | ((((uint32_t)c[2] << red) & 255) << 16);
  }
  const uint32_t w = f.img_w, h = f.img_h;
  if (i >= w * h) return 0;
  const uint32_t off = snake_forward_256(i, h), x = off & 255, y = off >> 8;
  uint8_t *Y, *Cb, *Cr; uint32_t YW, CW;
  dec_plane_ptrs(f, Y, Cb, Cr, YW, CW);
  const uint32_t cw = (w + 1) >> 1, ch = (h + 1) >> 1, cx = x >> 1, cy = y >> 1;
  const uint32_t oy = (y & 1) ? min(cy + 1, ch - 1) : (cy > 0 ? cy - 1 : 0);
  int cc[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const uint8_t *P = c ? Cr : Cb;
    int cs = 3 * P[(size_t)cy * CW + cx] + P[(size_t)oy * CW + cx], o;
    if (!(x & 1)) { if (cx == 0) o = (4 * cs + 8) >> 4; else { int l = 3 * P[(size_t)cy * CW + cx - 1] + P[(size_t)oy * CW + cx - 1]; o = (3 * cs + l + 8) >> 4; } }
    else { if (cx == cw - 1) o = (4 * cs + 7) >> 4; else { int r = 3 * P[(size_t)cy * CW + cx + 1] + P[(size_t)oy * CW + cx + 1]; o = (3 * cs + r + 7) >> 4; } }
    cc[c] = o - 128;
  }
  const int yy = Y[(size_t)y * YW + x];
  int R = yy + ((91881 * cc[1] + 32768) >> 16), B = yy + ((116130 * cc[0] + 32768) >> 16), G = yy + ((-22554 * cc[0] - 46802 * cc[1] + 32768) >> 16);
  R = min(255, max(0, R)); G = min(255, max(0, G)); B = min(255, max(0, B));
  return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
}

__global__ void __launch_bounds__(NODE_THREADS) dec_points_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.y];
  const uint32_t nb = f.n_bottom;
  const uint32_t ntiles = (nb + NODE_THREADS - 1) / NODE_THREADS;
  if (f.error || f.l2_valid || blockIdx.x >= ntiles) return;       // l2_valid: dec_leaves_kernel has written the points
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  __shared__ uint64_t s_prefix[NODE_THREADS];
  __shared__ uint16_t s_leaf[NODE_THREADS * 8];             // node-in-tile << 3 | child, in output order
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_NODES], 1u);
  __syncthreads();
  const uint32_t tile = s_tile, g = tile * NODE_THREADS + threadIdx.x;
  uint32_t byte = 0; uint64_t prefix = 0;
  if (g < nb) { byte = f.node_byte[g]; prefix = f.node_prefix[g]; }
  s_prefix[threadIdx.x] = prefix;
  const uint32_t cnt = __popc(byte);
  uint64_t tot;
  const uint64_t excl_local = block_excl_scan_u64(cnt, &tot, s_scan);
  if (threadIdx.x < 32) { uint64_t e = scan_lookback(f.scan_status, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  {                                                          // a node lists its leaves; afterwards one thread per LEAF
    uint32_t w = (uint32_t)excl_local, m = byte;
    while (m) { s_leaf[w++] = (uint16_t)((threadIdx.x << 3) | (__ffs(m) - 1)); m &= m - 1; }
  }
  __syncthreads();
  const uint64_t base = s_excl;
  const uint32_t detail = f.detail;
  if (g == nb - 1) { uint64_t V = base + excl_local + cnt; f.V = (uint32_t)V; if (detail ? (V != f.ncounts || V > f.counts_cap) : (V != f.point_count || V > f.out_cap)) atomicOr(&f.error, FERR_BAD_STREAM); }
  const double res = f.res;
  const uint32_t do_centroid = f.do_centroid;
  const uint64_t out_cap = f.out_cap;
  for (uint32_t j = threadIdx.x; j < (uint32_t)tot; j += NODE_THREADS) {     // consecutive threads write consecutive 32-byte records
    const uint64_t i64 = base + j;
    if (i64 >= (detail ? (uint64_t)f.counts_cap : out_cap)) break;
    const uint32_t i = (uint32_t)i64;
    const uint32_t e = s_leaf[j];
    const uint64_t key = (s_prefix[e >> 3] << 3) | (e & 7u);
    if (detail) { if (i < f.counts_cap) f.dleaf_key[i] = key; continue; }      // detail mode: detail_points_kernel expands the voxels
    const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
    float xyz[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (do_centroid) {                                    // pcv2.h:103-118
        double corner = __dadd_rn(__dmul_rn((double)k3[a], res), f.bmin[a]);
        uint32_t q = (3ull * i + a) < f.ncen ? f.cen[3ull * i + a] : 0;
        xyz[a] = (float)__dadd_rn(corner, (double)__fmul_rn((float)q, 0.001f));
      } else xyz[a] = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k3[a], 0.5), res), f.bmin[a]);   // impl.hpp:1630-1632
    }
    const uint32_t rgba = dec_color(f, i);
    uint4 *o = (uint4 *)(f.out_pts + 32ull * i);
    o[0] = make_uint4(__float_as_uint(xyz[0]), __float_as_uint(xyz[1]), __float_as_uint(xyz[2]), 0x3F800000u);
    o[1] = make_uint4(rgba, 0, 0, 0);
  }
}

// The pipelined walker's level-(d-2) records straight to points: a record's children are the bottom-level bytes at its
// stream offset, their set bits are the leaves.  One chained scan over the records' leaf counts gives every record its
// place in the output; the records of a tile list their leaves in shared memory and then one thread per LEAF writes a
// 32-byte point (consecutive threads, consecutive records).  Does the work of dec_points_kernel (and of a separate expansion pass) on this
// path: no bottom-node arrays, one scan instead of two.  grid (ceil(node_cap / 256), frames)
__global__ void __launch_bounds__(256) dec_leaves_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.y];
  if (f.error || !f.l2_valid) return;
  const uint32_t n2 = f.n_l2;
  const uint32_t ntiles = (n2 + 255) / 256;
  if (blockIdx.x >= ntiles) return;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  __shared__ uint64_t s_prefix[256];
  __shared__ uint16_t s_leaf[256 * 64];                      // record-in-tile << 6 | child << 3 | leaf bit, in output order
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_EXPAND], 1u);
  __syncthreads();
  const uint32_t tile = s_tile, g = tile * 256 + threadIdx.x;
  uint32_t mask = 0, off = 0; uint64_t prefix = 0;
  if (g < n2) { mask = f.l2_mask[g]; off = f.l2_off[g]; prefix = f.l2_prefix[g]; }
  s_prefix[threadIdx.x] = prefix;
  uint32_t bytes[8], cnt = 0;
  const uint32_t k = __popc(mask);
#pragma unroll
  for (int i = 0; i < 8; i++) { bytes[i] = (uint32_t)i < k ? f.tree[off + i] : 0u; cnt += __popc(bytes[i]); }
  uint64_t tot;
  const uint64_t excl_local = block_excl_scan_u64(cnt, &tot, s_scan);
  if (threadIdx.x < 32) { uint64_t e = scan_lookback(f.scan_status + f.scan_tiles_max, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  {
    uint32_t w = (uint32_t)excl_local, m = mask;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (!m) break;
      const uint32_t c = __ffs(m) - 1; m &= m - 1;
      uint32_t b = bytes[i];
      while (b) { s_leaf[w++] = (uint16_t)((threadIdx.x << 6) | (c << 3) | (__ffs(b) - 1)); b &= b - 1; }
    }
  }
  __syncthreads();
  const uint64_t base = s_excl;
  const uint32_t detail = f.detail;
  if (g == n2 - 1) { uint64_t V = base + excl_local + cnt; f.V = (uint32_t)V; if (detail ? (V != f.ncounts || V > f.counts_cap) : (V != f.point_count || V > f.out_cap)) atomicOr(&f.error, FERR_BAD_STREAM); }
  const double res = f.res;
  const uint32_t do_centroid = f.do_centroid;
  const uint64_t out_cap = f.out_cap;
  for (uint32_t j = threadIdx.x; j < (uint32_t)tot; j += 256) {
    const uint64_t i64 = base + j;
    if (i64 >= (detail ? (uint64_t)f.counts_cap : out_cap)) break;
    const uint32_t i = (uint32_t)i64;
    const uint32_t e = s_leaf[j];
    const uint64_t key = (((s_prefix[e >> 6] << 3) | ((e >> 3) & 7u)) << 3) | (e & 7u);
    if (detail) { if (i < f.counts_cap) f.dleaf_key[i] = key; continue; }      // detail mode: detail_points_kernel expands the voxels
    const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
    float xyz[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (do_centroid) {                                    // pcv2.h:103-118
        double corner = __dadd_rn(__dmul_rn((double)k3[a], res), f.bmin[a]);
        uint32_t q = (3ull * i + a) < f.ncen ? f.cen[3ull * i + a] : 0;
        xyz[a] = (float)__dadd_rn(corner, (double)__fmul_rn((float)q, 0.001f));
      } else xyz[a] = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k3[a], 0.5), res), f.bmin[a]);   // impl.hpp:1630-1632
    }
    const uint32_t rgba = dec_color(f, i);
    uint4 *o = (uint4 *)(f.out_pts + 32ull * i);
    o[0] = make_uint4(__float_as_uint(xyz[0]), __float_as_uint(xyz[1]), __float_as_uint(xyz[2]), 0x3F800000u);
    o[1] = make_uint4(rgba, 0, 0, 0);
  }
}

// Round trip: hands the encoder's device-resident stream of every frame of a group to the decoder's frame record.
__global__ void link_kernel(const EncFrame *enc, DecFrame *dec, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dec[i].in = enc[i].stream;
  dec[i].in_len = enc[i].out_len;
  if (enc[i].error || enc[i].out_len == 0) dec[i].error = FERR_BAD_STREAM;   // empty / failed frame: the host skips it
}

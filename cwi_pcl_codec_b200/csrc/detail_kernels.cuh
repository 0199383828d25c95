// detail_kernels.cuh -- detail mode (doVoxelGridDownDownSampling = false, the class default: codec.h:108-143).
// Reference behaviour restated (oracle/ccv2_oracle.c carries the CPU statement of the same steps):
//   serializeTreeCallback impl.hpp:1525-1541: per voxel the point count, [PCL] PointCoding::encodePoints (3 residual bytes
//     per point against the voxel's lower corner at point_resolution), [PCL] ColorCoding::encodePoints (average colour +
//     XOR differences of every point of a multi-point voxel)
//   entropyEncoding impl.hpp:1728-1757: u64 + int-coded counts ([PCL] StaticRangeCoder::encodeIntVectorToStream, the
//     64-bit coder), u64 + char-coded point differences, u64 + char-coded colour differences
//   entropyDecoding impl.hpp:1802-1832 and deserializeTreeCallback impl.hpp:1592-1613 for the inverse.
#pragma once
#include "common.cuh"
#include "entropy_kernels.cuh"

// ---- encode: counts and colour-difference offsets.  One chained scan over the voxels of (len > 1 ? len : 0).
// grid (ceil(n / 1024), frames), 256 threads x 4 voxels; status words: 4th region of scan_status
__global__ void __launch_bounds__(256) detail_scan_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  const uint32_t ntiles = (V + 1023) / 1024;
  if (blockIdx.x >= ntiles) return;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_DETAIL], 1u);
  __syncthreads();
  const uint32_t tile = s_tile, j0 = tile * 1024 + threadIdx.x * 4;
  uint32_t len[4]; uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t j = j0 + k;
    len[k] = j < V ? f.leaf_start[j + 1] - f.leaf_start[j] : 0u;
    sum += len[k] > 1 ? len[k] : 0u;
  }
  uint64_t tot;
  uint64_t ex = block_excl_scan_u64(sum, &tot, s_scan);
  if (threadIdx.x < 32) { const uint64_t e = scan_lookback(f.scan_status + 3 * (size_t)f.scan_tiles_max, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  __syncthreads();
  ex += s_excl;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t j = j0 + k;
    if (j < V) {
      f.counts[j] = len[k]; f.cd_off[j] = (uint32_t)ex;
      ex += len[k] > 1 ? len[k] : 0u;
      if (j == V - 1) { f.ncd = 3u * (uint32_t)ex; f.npd = 3u * f.leaf_start[V]; }
    }
  }
}

// ---- encode: one thread per voxel writes the residuals of its points (sorted order = index order inside a voxel) and,
// for a multi-point voxel, the XOR colour differences against the voxel average.  grid (ceil(n / 256), frames)
__global__ void __launch_bounds__(256) detail_emit_kernel(EncFrame *frames, EncParams P) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f.V) return;
  const uint64_t key = f.leaf_key[j];
  const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
  double corner[3];
#pragma unroll
  for (int a = 0; a < 3; a++) corner[a] = __dadd_rn(__dmul_rn((double)k3[a], P.res), f.bmin[a]);   // impl.hpp:1518-1520
  const uint32_t s0 = f.leaf_start[j], s1 = f.leaf_start[j + 1], len = s1 - s0;
  const uint32_t *vals = f.vals[f.npasses & 1];
  const double pres = (double)P.point_res_f;
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  for (uint32_t k = s0; k < s1; k++) {
    const uint8_t *rec = f.pts + 32ull * vals[k];
    const float4 q = __ldg((const float4 *)rec);
    const float pf[3] = { q.x, q.y, q.z };
#pragma unroll
    for (int a = 0; a < 3; a++) {                        // [PCL] PointCoding::encodePoints: (p - corner) / precision in double, C truncation, clamp +-127
      const double t = __ddiv_rn(__dsub_rn((double)pf[a], corner[a]), pres);
      int qi = (int)t;
      qi = max(-127, min(127, qi));
      f.pdiff[3ull * k + a] = (uint8_t)qi;
    }
    if (P.do_color) { const uint32_t c = __ldg((const uint32_t *)(rec + 16)); c0 += c & 0xFF; c1 += (c >> 8) & 0xFF; c2 += (c >> 16) & 0xFF; }
  }
  if (P.do_color && len > 1) {                           // [PCL] ColorCoding::encodePoints: differences only for multi-point voxels, before the reduction of the average
    c0 /= len; c1 /= len; c2 /= len;
    uint8_t *o = f.cdiff + 3ull * f.cd_off[j];
    for (uint32_t k = s0; k < s1; k++) {
      const uint32_t c = __ldg((const uint32_t *)(f.pts + 32ull * vals[k] + 16));
      o[0] = (uint8_t)(((uint8_t)c0 ^ (uint8_t)(c & 0xFF)) >> P.color_reduction);
      o[1] = (uint8_t)(((uint8_t)c1 ^ (uint8_t)((c >> 8) & 0xFF)) >> P.color_reduction);
      o[2] = (uint8_t)(((uint8_t)c2 ^ (uint8_t)((c >> 16) & 0xFF)) >> P.color_reduction);
      o += 3;
    }
  }
}

// ---- [PCL] StaticRangeCoder::encodeIntVectorToStream for the point counts.  One warp per frame: table size by the
// reference's growth rule (order dependent: found in one pass, a warp-wide search for the next symbol that makes the table
// grow), histogram, cumulative table (+1 for empty symbols), header, then the 64-bit carry-less coder on lane 0.
__global__ void __launch_bounds__(32) rc_encode_int_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.x];
  const uint32_t V = f.V, lane = lane_id();
  if (V == 0 || f.error) return;
  const uint32_t *in = f.counts;
  // table size: for every symbol in order: if (sym + 1 >= size) do size <<= 1 while (sym + 1 > size)
  uint64_t tsize = 1;
  for (uint32_t base = 0; base < V;) {
    const uint32_t i = base + lane;
    const bool grow = i < V && (uint64_t)in[i] + 1 >= tsize;
    const uint32_t m = __ballot_sync(FULL_MASK, grow);
    if (!m) { base += 32; continue; }
    const uint32_t first = __ffs(m) - 1;
    const uint64_t sym = __shfl_sync(FULL_MASK, i < V ? in[i] : 0u, first);
    do { tsize <<= 1; } while (sym + 1 > tsize);
    base += first + 1;                                   // later symbols are tested against the grown table
  }
  tsize++;
  bool bad = tsize > f.itab_cap;
  uint64_t *cf = f.itab;
  if (!bad) {
    for (uint64_t k = lane; k < tsize; k += 32) cf[k] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < V; i += 32) atomicAdd((unsigned long long *)&cf[in[i] + 1], 1ull);
    __syncwarp(); __threadfence_block();
    uint64_t carry = 0;                                  // cf[f] = cf[f-1] + max(hist[f], 1)
    for (uint64_t k0 = 1; k0 < tsize; k0 += 32) {
      const uint64_t k = k0 + lane;
      uint64_t v = k < tsize ? (cf[k] ? cf[k] : 1ull) : 0ull;
      v = warp_incl_scan_u64(v) + carry;
      if (k < tsize) cf[k] = v;
      carry = __shfl_sync(FULL_MASK, v, 31);
    }
    __syncwarp(); __threadfence_block();
  }
  const uint64_t total = bad ? 1 : cf[tsize - 1];          // < 2^48 for any frame this codec accepts: the reference's rescaling never triggers
  const uint32_t bits = 64 - __clzll((long long)total);
  const uint32_t bsz = (bits + 7) / 8;                      // == ceil(log2(total + 1) / 8)
  const uint64_t hdr = 9 + (tsize - 1) * bsz;
  if (!bad && hdr + 8ull * V + 16 > f.rc_int_cap) bad = true;   // a symbol costs at most 48 bits
  if (bad) { if (lane == 0) { atomicOr(&f.error, FERR_TREE_CAP); f.rc_int_len = 0; } return; }
  uint8_t *out = f.rc_int;
  if (lane < 8) out[lane] = (uint8_t)(tsize >> (8 * lane));
  if (lane == 8) out[8] = (uint8_t)bsz;
  for (uint64_t k = 1 + lane; k < tsize; k += 32) { const uint64_t v = cf[k]; uint8_t *o = out + 9 + (k - 1) * bsz; for (uint32_t b = 0; b < bsz; b++) o[b] = (uint8_t)(v >> (8 * b)); }
  __syncwarp();
  if (lane == 0) {
    const uint64_t top = 1ull << 56, bottom = 1ull << 48;
    uint64_t low = 0, range = ~0ull, w = hdr;
    for (uint32_t i = 0; i < V; i++) {
      const uint32_t sym = in[i];
      range /= total;
      low += cf[sym] * range;
      range *= cf[sym + 1] - cf[sym];
      while ((low ^ (low + range)) < top || (range < bottom && ((range = (0ull - low) & (bottom - 1)), true))) {
        out[w++] = (uint8_t)(low >> 56); range <<= 8; low <<= 8;
      }
    }
    for (int k = 0; k < 8; k++) { out[w++] = (uint8_t)(low >> 56); low <<= 8; }
    f.rc_int_len = (uint32_t)w; f.itsize = (uint32_t)tsize;
  }
}

// ---- decodeStreamToIntVector (impl.hpp:1817), by the calling warp; lane 0 runs the coder.  Returns false on a malformed layer.
__device__ inline bool rc_decode_int_layer(const uint8_t *base, uint64_t len, uint64_t &pos, uint32_t *out, uint32_t n, uint64_t *cf, uint32_t cf_cap) {
  const uint32_t lane = lane_id();
  if (pos + 9 > len) return false;
  const uint64_t tsize = ld_u64_unaligned(base + pos); const uint32_t bsz = base[pos + 8];
  if (tsize < 2 || tsize > cf_cap || bsz == 0 || bsz > 8 || pos + 9 + (tsize - 1) * bsz + 8 > len) return false;
  const uint8_t *tab = base + pos + 9;
  if (lane == 0) cf[0] = 0;
  for (uint64_t k = 1 + lane; k < tsize; k += 32) { uint64_t v = 0; const uint8_t *q = tab + (k - 1) * bsz; for (int b = (int)bsz - 1; b >= 0; b--) v = (v << 8) | q[b]; cf[k] = v; }
  __syncwarp(); __threadfence_block();
  uint64_t p = pos + 9 + (tsize - 1) * bsz;
  int ok = 1;
  if (lane == 0) {
    const uint64_t top = 1ull << 56, bottom = 1ull << 48, total = cf[tsize - 1];
    uint64_t code = 0, low = 0, range = ~0ull;
    for (int k = 0; k < 8; k++) code = (code << 8) | base[p++];
    if (total == 0) ok = 0;
    for (uint32_t i = 0; ok && i < n; i++) {
      range /= total;
      if (range == 0) { ok = 0; break; }
      const uint64_t count = (code - low) / range;
      uint64_t sym = 0, ss = (tsize - 1) / 2;                // the reference's descent: sSize = (size - 1) / 2, halved
      while (ss > 0) { if (cf[sym + ss] <= count) sym += ss; ss /= 2; }
      out[i] = (uint32_t)sym;
      low += cf[sym] * range;
      range *= cf[sym + 1] - cf[sym];
      if (range == 0) { ok = 0; break; }
      while ((low ^ (low + range)) < top || (range < bottom && ((range = (0ull - low) & (bottom - 1)), true))) {
        const uint64_t ch = p < len ? base[p] : 0; p++;
        code = (code << 8) | ch; range <<= 8; low <<= 8;
      }
    }
    if (p > len) ok = 0;
  }
  ok = __shfl_sync(FULL_MASK, ok, 0);
  p = __shfl_sync(FULL_MASK, p, 0);
  pos = p;
  return ok != 0;
}

// The enhancement vectors that follow the colour layer (impl.hpp:1808-1832), decoded by the calling warp.  On entry pos is
// the first byte after the colour layer (pos < len: that is what switches the reference into detail mode, :1802-1806).
__device__ inline bool decode_detail_layers(DecFrame &f, const uint8_t *in, uint64_t len, uint64_t &pos, uint32_t *freq_s, uint32_t &err) {
  uint64_t cd;
  if (pos + 8 > len) return false;
  const uint64_t nc = ld_u64_unaligned(in + pos); pos += 8;
  if (nc > f.counts_cap) { err |= FERR_TREE_CAP; return false; }
  if (!rc_decode_int_layer(in, len, pos, f.counts, (uint32_t)nc, f.itab, f.itab_cap)) return false;
  if (pos + 8 > len) return false;
  const uint64_t np = ld_u64_unaligned(in + pos); pos += 8;
  if (np > f.pdiff_cap) { err |= FERR_TREE_CAP; return false; }
  if (!rc_decode_layer<false>(in, len, pos, f.pdiff, (uint32_t)np, freq_s, &cd)) return false;
  uint64_t ncd = 0;
  if (f.data_with_color) {
    if (pos + 8 > len) return false;
    ncd = ld_u64_unaligned(in + pos); pos += 8;
    if (ncd > f.pdiff_cap) { err |= FERR_TREE_CAP; return false; }
    if (!rc_decode_layer<false>(in, len, pos, f.cdiff, (uint32_t)ncd, freq_s, &cd)) return false;
    // The reference reads these differences into color_coder_ (impl.hpp:1828) but decodes JPEG-type colours with
    // jp_color_coder_, whose difference vector is empty: undefined behaviour.  Only type 0 is defined.
    if (f.cct != 0) { err |= FERR_UNSUPPORTED; return false; }
  }
  if (lane_id() == 0) { f.detail = 1; f.ncounts = (uint32_t)nc; f.npdiff = np; f.ncdiff = ncd; }
  return true;
}

// ---- decode: voxels -> points.  One chained scan over the counts; a thread per voxel writes its points
// ([PCL] PointCoding::decodePoints, ColorCoding::decodePoints).  grid (ceil(voxels / 256), frames)
__device__ __forceinline__ uint32_t dec_color(const DecFrame &f, uint32_t i);
__global__ void __launch_bounds__(256) detail_points_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.y];
  if (f.error || !f.detail) return;
  const uint32_t V = f.V;
  const uint32_t ntiles = (V + 255) / 256;
  if (blockIdx.x >= ntiles) return;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_DETAIL], 1u);
  __syncthreads();
  const uint32_t tile = s_tile, j = tile * 256 + threadIdx.x;
  const uint32_t cnt = j < V ? f.counts[j] : 0u;
  // two sums in one scan: points (low 32 bits) and points of multi-point voxels (high 32 bits: where the colour differences start)
  uint64_t tot;
  uint64_t ex = block_excl_scan_u64((uint64_t)cnt | ((uint64_t)(cnt > 1 ? cnt : 0u) << 32), &tot, s_scan);
  if (threadIdx.x < 32) { const uint64_t e = scan_lookback(f.scan_status + 2 * (size_t)f.scan_tiles_max, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  __syncthreads();
  ex += s_excl;
  if (j >= V) return;
  const uint64_t p0 = ex & 0xFFFFFFFFull, c0 = ex >> 32;
  if (j == V - 1) {                                         // totals must match the header and the vectors
    const uint64_t np = p0 + cnt, nc = c0 + (cnt > 1 ? cnt : 0u);
    if (np != f.point_count || 3 * np != f.npdiff || (f.data_with_color && 3 * nc != f.ncdiff) || f.ncounts != V) atomicOr(&f.error, FERR_BAD_STREAM);
    f.npoints_out = (uint32_t)min(np, (uint64_t)0xFFFFFFFFu);   // what the caller gets back: the number of points
  }
  if (p0 + cnt > f.out_cap || 3 * (p0 + cnt) > f.npdiff || (f.data_with_color && cnt > 1 && 3 * (c0 + cnt) > f.ncdiff)) { atomicOr(&f.error, FERR_BAD_STREAM); return; }
  const uint64_t key = f.dleaf_key[j];
  const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
  double corner[3];
#pragma unroll
  for (int a = 0; a < 3; a++) corner[a] = __dadd_rn(__dmul_rn((double)k3[a], f.res), f.bmin[a]);
  const float pres_f = f.point_res_f;
  const uint32_t avg = dec_color(f, j);                     // type 0: the average, already shifted back by the bit reduction
  const uint32_t red = f.cct == 0 ? 8 - f.color_bits : 0;
  for (uint32_t q = 0; q < cnt; q++) {
    const uint8_t *d = f.pdiff + 3 * (p0 + q);
    float xyz[3];
#pragma unroll
    for (int a = 0; a < 3; a++) xyz[a] = (float)__dadd_rn(corner[a], (double)__fmul_rn((float)d[a], pres_f));   // uchar * float precision, added to the double corner
    uint32_t rgba = avg;
    if (f.data_with_color && cnt > 1) {
      const uint8_t *e = f.cdiff + 3 * (c0 + q);
      const uint32_t df = (((uint32_t)e[0] << red) & 255u) | ((((uint32_t)e[1] << red) & 255u) << 8) | ((((uint32_t)e[2] << red) & 255u) << 16);
      rgba = avg ^ df;
    }
    uint4 *o = (uint4 *)(f.out_pts + 32ull * (p0 + q));
    o[0] = make_uint4(__float_as_uint(xyz[0]), __float_as_uint(xyz[1]), __float_as_uint(xyz[2]), 0x3F800000u);
    o[1] = make_uint4(rgba, 0, 0, 0);
  }
}

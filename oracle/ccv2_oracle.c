/*
 * ccv2_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see ccv2_oracle.h for status / parity pinning).
 *
 * Every function cites the reference file:line whose behaviour it restates.  Abbreviations:
 *   impl.hpp  = cloud_codec_v2/include/pcl/cloud_codec_v2/impl/point_cloud_codec_v2_impl.hpp
 *   cjpeg.h   = cloud_codec_v2/include/pcl/cloud_codec_v2/color_coding_jpeg.h
 *   snake.h   = cloud_codec_v2/include/pcl/cloud_codec_v2/snake_grid_mapping.h
 *   pcv2.h    = cloud_codec_v2/include/pcl/cloud_codec_v2/point_coding_v2.h
 *   jpeg_io.hpp = jpeg_io/include/pcl/io/impl/jpeg_io.hpp
 * [PCL] / [libjpeg] = un-vendored third-party code the reference calls (PCL 1.8-1.10 pinned by
 * CMakeLists.txt:65-68 / README.md:13; libjpeg-turbo by CMakeLists.txt:121); restated from the
 * published algorithms.
 */
#include "ccv2_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ utils */
typedef struct { uint8_t *p; size_t n, cap; } bbuf;
static int bb_reserve(bbuf *b, size_t extra) {
  if (b->n + extra <= b->cap) return 0;
  size_t nc = b->cap ? b->cap * 2 : 4096;
  while (nc < b->n + extra) nc *= 2;
  uint8_t *q = (uint8_t *)realloc(b->p, nc);
  if (!q) return -1;
  b->p = q; b->cap = nc; return 0;
}
static inline void bb_push(bbuf *b, uint8_t v) { if (b->n == b->cap) bb_reserve(b, 1); b->p[b->n++] = v; }
static void bb_write(bbuf *b, const void *src, size_t n) { bb_reserve(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

void orc_free(void *p) { free(p); }
void orc_free_debug(orc_debug *d) {
  if (!d) return;
  free(d->leaf_keys); free(d->tree_bytes); free(d->avg_colors); free(d->color_payload); free(d->centroid_bytes); free(d->output_cloud);
  memset(d, 0, sizeof *d);
}
void orc_default_params(orc_params *p) {
  /* evaluate_compression defaults: eval.hpp:377-395 with parameter_config.txt:1-18 */
  memset(p, 0, sizeof *p);
  p->point_resolution = ldexp(1.0, -11); p->octree_resolution = ldexp(1.0, -11);
  p->do_voxel_grid = 1; p->do_color = 1; p->color_bit_resolution = 8; p->color_coding_type = 1;
  p->do_centroid = 0; p->create_scalable = 0; p->code_connectivity = 0; p->jpeg_quality = 85;
  p->macroblock_size = 16; p->do_icp_color_offset = 0;
}

/* ------------------------------------------------------------------ [PCL] StaticRangeCoder (char vectors)
 * pcl/compression/impl/entropy_range_coder.hpp, StaticRangeCoder::encodeCharVectorToStream /
 * decodeStreamToCharVector; called at impl.hpp:1694,1706,1719 and :1778,1789,1798.
 * The char-vector coder is the 32-bit Subbotin carry-less coder (DWord = uint32_t, top = 1<<24,
 * bottom = 1<<16, totals rescaled below 1<<16, 4 flush bytes); the 64-bit variant in the same file is
 * the *Int*-vector coder used only by detail mode (impl.hpp:1738).  SURVEY.md App. B.4 describes the
 * 64-bit arithmetic for both; see DESIGN.md "range coder word size" for why this oracle uses 32 bits. */
static void rc_build_table(const uint8_t *in, size_t n, uint32_t freq[257]) {
  uint64_t hist[257];
  memset(hist, 0, sizeof hist);
  for (size_t i = 0; i < n; i++) hist[in[i] + 1]++;
  freq[0] = 0;
  for (int f = 1; f <= 256; f++) {
    freq[f] = freq[f - 1] + (uint32_t)hist[f];
    if (freq[f] <= freq[f - 1]) freq[f] = freq[f - 1] + 1;
  }
  while (freq[256] >= (1u << 16)) {          /* "rescale if numerical limits are reached" */
    for (int f = 1; f <= 256; f++) {
      freq[f] /= 2;
      if (freq[f] <= freq[f - 1]) freq[f] = freq[f - 1] + 1;
    }
  }
}
static void rc_encode_to(bbuf *os, const uint8_t *in, size_t n, uint64_t *coded_len) {
  const uint32_t top = 1u << 24, bottom = 1u << 16;
  uint32_t freq[257];
  size_t start = os->n;
  rc_build_table(in, n, freq);
  bb_write(os, freq, sizeof freq);             /* raw little-endian u32 x 257 */
  uint32_t low = 0, range = 0xFFFFFFFFu;
  bb_reserve(os, n + n / 2 + 16);
  for (size_t i = 0; i < n; i++) {
    uint8_t ch = in[i];
    range /= freq[256];
    low += freq[ch] * range;
    range *= freq[ch + 1] - freq[ch];
    while ((low ^ (low + range)) < top || (range < bottom && ((range = (0u - low) & (bottom - 1)), 1))) {
      bb_push(os, (uint8_t)(low >> 24));
      range <<= 8; low <<= 8;
    }
  }
  for (int i = 0; i < 4; i++) { bb_push(os, (uint8_t)(low >> 24)); low <<= 8; }
  if (coded_len) *coded_len = os->n - start;
}
int orc_range_encode(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  bbuf b = {0};
  rc_encode_to(&b, in, n, NULL);
  *out = b.p; *out_len = b.n; return 0;
}
int orc_range_decode(const uint8_t *in, size_t in_len, uint8_t *out, size_t n, size_t *consumed) {
  const uint32_t top = 1u << 24, bottom = 1u << 16;
  uint32_t freq[257];
  size_t pos = 0;
  if (in_len < sizeof freq + 4) return -1;
  memcpy(freq, in, sizeof freq); pos = sizeof freq;
  if (freq[256] == 0) return -2;
  uint32_t code = 0, low = 0, range = 0xFFFFFFFFu;
  for (int i = 0; i < 4; i++) code = (code << 8) | in[pos++];
  for (size_t i = 0; i < n; i++) {
    range /= freq[256];
    if (range == 0) return -3;
    uint32_t count = (code - low) / range;
    uint8_t sym = 0;
    for (int s = 128; s > 0; s >>= 1) if (freq[sym + s] <= count) sym = (uint8_t)(sym + s);
    out[i] = sym;
    low += freq[sym] * range;
    range *= freq[sym + 1] - freq[sym];
    while ((low ^ (low + range)) < top || (range < bottom && ((range = (0u - low) & (bottom - 1)), 1))) {
      uint8_t ch = pos < in_len ? in[pos] : 0;   /* PCL reads past EOF as whatever the stream gives; 0 here */
      pos++;
      code = (code << 8) | ch; range <<= 8; low <<= 8;
    }
  }
  if (consumed) *consumed = pos;
  return pos <= in_len ? 0 : -4;
}

/* ------------------------------------------------------------------ [PCL] StaticRangeCoder (int vectors)
 * pcl/compression/impl/entropy_range_coder.hpp, StaticRangeCoder::encodeIntVectorToStream / decodeStreamToIntVector;
 * called only in detail mode (impl.hpp:1738, :1817) for the per-voxel point counts.  64-bit Subbotin coder
 * (top = 1<<56, bottom = 1<<48, 8 flush bytes) over `unsigned int` symbols with a table that grows by doubling while
 * the symbols are scanned; header = u64 table size, u8 bytes per entry, entries 1..size-1 at that width.  Restated from
 * the published PCL source (status [M]: unverifiable offline like the rest of the PCL-inherited parts). */
static void rc_encode_int_to(bbuf *os, const uint32_t *in, size_t n, uint64_t *coded_len) {
  const uint64_t top = 1ull << 56, bottom = 1ull << 48, max_range = 1ull << 48;
  size_t start = os->n;
  uint64_t tsize = 1;                                   /* frequencyTableSize */
  size_t cap = 4;
  uint64_t *cf = (uint64_t *)calloc(cap, 8);
  for (size_t i = 0; i < n; i++) {
    uint64_t sym = in[i];
    if (sym + 1 >= tsize) {                            /* "frequency table is to small -> adaptively extend it" */
      uint64_t old = tsize;
      do { tsize <<= 1; } while (sym + 1 > tsize);
      if (cap < tsize + 1) { size_t nc = (size_t)tsize + 1; cf = (uint64_t *)realloc(cf, nc * 8); memset(cf + cap, 0, (nc - cap) * 8); cap = nc; }
      memset(cf + old + 1, 0, (size_t)(tsize - old) * 8);
    }
    cf[sym + 1]++;
  }
  tsize++;
  if (cap < tsize) { cf = (uint64_t *)realloc(cf, (size_t)tsize * 8); memset(cf + cap, 0, ((size_t)tsize - cap) * 8); cap = (size_t)tsize; }
  for (uint64_t f = 1; f < tsize; f++) { cf[f] = cf[f - 1] + cf[f]; if (cf[f] <= cf[f - 1]) cf[f] = cf[f - 1] + 1; }
  while (cf[tsize - 1] >= max_range)
    for (uint64_t f = 1; f < tsize; f++) { cf[f] /= 2; if (cf[f] <= cf[f - 1]) cf[f] = cf[f - 1] + 1; }
  uint8_t bsz = (uint8_t)ceil(log2((double)(cf[tsize - 1] + 1)) / 8.0);
  bb_write(os, &tsize, 8); bb_write(os, &bsz, 1);
  for (uint64_t f = 1; f < tsize; f++) bb_write(os, &cf[f], bsz);           /* little-endian host: the low bsz bytes */
  uint64_t low = 0, range = ~0ull;
  for (size_t i = 0; i < n; i++) {
    uint32_t sym = in[i];
    range /= cf[tsize - 1];
    low += cf[sym] * range;
    range *= cf[sym + 1] - cf[sym];
    while ((low ^ (low + range)) < top || (range < bottom && ((range = (0ull - low) & (bottom - 1)), 1))) {
      bb_push(os, (uint8_t)(low >> 56));
      range <<= 8; low <<= 8;
    }
  }
  for (int i = 0; i < 8; i++) { bb_push(os, (uint8_t)(low >> 56)); low <<= 8; }
  free(cf);
  if (coded_len) *coded_len = os->n - start;
}
int orc_range_encode_int(const uint32_t *in, size_t n, uint8_t **out, size_t *out_len) {
  bbuf b = {0};
  rc_encode_int_to(&b, in, n, NULL);
  *out = b.p; *out_len = b.n; return 0;
}
int orc_range_decode_int(const uint8_t *in, size_t in_len, uint32_t *out, size_t n, size_t *consumed) {
  const uint64_t top = 1ull << 56, bottom = 1ull << 48;
  size_t pos = 0;
  if (in_len < 9) return -1;
  uint64_t tsize; memcpy(&tsize, in, 8); uint8_t bsz = in[8]; pos = 9;
  if (tsize < 2 || tsize > (1ull << 28) || bsz == 0 || bsz > 8 || pos + (tsize - 1) * bsz + 8 > in_len) return -2;
  uint64_t *cf = (uint64_t *)calloc((size_t)tsize + 1, 8);
  for (uint64_t f = 1; f < tsize; f++) { memcpy(&cf[f], in + pos, bsz); pos += bsz; }
  uint64_t code = 0, low = 0, range = ~0ull;
  for (int i = 0; i < 8; i++) code = (code << 8) | in[pos++];
  int rc = 0;
  for (size_t i = 0; i < n; i++) {
    range /= cf[tsize - 1];
    if (range == 0) { rc = -3; break; }
    uint64_t count = (code - low) / range;
    size_t sym = 0, ss = (size_t)((tsize - 1) / 2);
    while (ss > 0) { if (cf[sym + ss] <= count) sym += ss; ss /= 2; }
    out[i] = (uint32_t)sym;
    low += cf[sym] * range;
    range *= cf[sym + 1] - cf[sym];
    while ((low ^ (low + range)) < top || (range < bottom && ((range = (0ull - low) & (bottom - 1)), 1))) {
      uint8_t ch = pos < in_len ? in[pos] : 0; pos++;
      code = (code << 8) | ch; range <<= 8; low <<= 8;
    }
  }
  free(cf);
  if (consumed) *consumed = pos;
  return rc ? rc : (pos <= in_len ? 0 : -4);
}

/* ------------------------------------------------------------------ snake (snake.h:46-71; closed form SURVEY App. B.7) */
void orc_snake_positions_literal(int w, int h, int32_t *pos) {
  int w_pos = 0, h_pos = 0, mbw = 0, mbh = 0, to_right = 1;
  for (long i = 0; i < (long)w * h; i++) {
    pos[i] = w_pos + (h_pos + 8 * mbh) * w + mbw * 8;          /* operator++ snake.h:67-71 */
    if (to_right) w_pos++; else w_pos--;                         /* updatePos snake.h:46-64 */
    if (((w_pos % 8 == 0) && to_right) || w_pos < 0) {
      h_pos++;
      to_right = !to_right;
      w_pos = to_right ? 0 : 7;
      if (h_pos % 8 == 0 || (h_pos + mbh * 8 == h)) {
        h_pos = 0; mbw++;
        if (mbw % (w / 8) == 0) { mbw = 0; mbh++; }
      }
    }
  }
}
int32_t orc_snake_pos_closed(int w, int h, int64_t i) {
  int64_t nbw = w / 8, F = h / 8, full = F * nbw * 64, mh, j, k, R;
  if (i < full) { mh = i / (nbw * 64); j = (i % (nbw * 64)) / 64; k = i % 64; R = 8; }
  else { mh = F; R = h - 8 * F; int64_t ip = i - full; j = ip / (8 * R); k = ip % (8 * R); }
  int64_t r = k / 8, c = k % 8;
  int64_t col = ((j * R + r) & 1) ? 7 - c : c;
  return (int32_t)(col + (r + 8 * mh) * w + 8 * j);
}

/* ------------------------------------------------------------------ [libjpeg] baseline JPEG */
static const uint8_t ZZ[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
  21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
static const uint8_t QBASE_L[64] = { 16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
  14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
static const uint8_t QBASE_C[64] = { 17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
  47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
static const uint8_t DC_L_BITS[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 };
static const uint8_t DC_C_BITS[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
static const uint8_t DC_VALS[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t AC_L_BITS[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d };
static const uint8_t AC_L_VALS[162] = { 0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
  0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
  0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
  0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
  0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
  0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
  0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8,
  0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
static const uint8_t AC_C_BITS[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };
static const uint8_t AC_C_VALS[162] = { 0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
  0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
  0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46,
  0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
  0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98,
  0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
  0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7,
  0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

typedef struct { uint16_t code[256]; uint8_t len[256]; } huff_enc;
static void huff_build_enc(const uint8_t bits[16], const uint8_t *vals, huff_enc *h) {
  memset(h, 0, sizeof *h);
  unsigned code = 0; int k = 0;
  for (int l = 1; l <= 16; l++) {
    for (int i = 0; i < bits[l - 1]; i++) { h->code[vals[k]] = (uint16_t)code; h->len[vals[k]] = (uint8_t)l; code++; k++; }
    code <<= 1;
  }
}
static void jpeg_quant_table(const uint8_t base[64], int quality, uint16_t q[64] /* natural order */) {
  if (quality <= 0) quality = 1;
  if (quality > 100) quality = 100;
  int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
  for (int i = 0; i < 64; i++) {
    long t = ((long)base[i] * scale + 50L) / 100L;
    if (t <= 0) t = 1;
    if (t > 255) t = 255;                       /* force_baseline = TRUE (jpeg_io.hpp:291-292) */
    q[i] = (uint16_t)t;
  }
}
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static void fdct_islow(int *d /* 64, natural order, in place */) {
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < 8; k++) {
      int *p = pass == 0 ? d + 8 * k : d + k;
      int st = pass == 0 ? 1 : 8;
      int v0 = p[0], v1 = p[st], v2 = p[2 * st], v3 = p[3 * st], v4 = p[4 * st], v5 = p[5 * st], v6 = p[6 * st], v7 = p[7 * st];
      int t0 = v0 + v7, t7 = v0 - v7, t1 = v1 + v6, t6 = v1 - v6, t2 = v2 + v5, t5 = v2 - v5, t3 = v3 + v4, t4 = v3 - v4;
      int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
      int n = pass == 0 ? 11 : 15;
      if (pass == 0) { p[0] = (t10 + t11) << 2; p[4 * st] = (t10 - t11) << 2; }
      else { p[0] = DESCALE(t10 + t11, 2); p[4 * st] = DESCALE(t10 - t11, 2); }
      int z1 = (t12 + t13) * 4433;
      p[2 * st] = DESCALE(z1 + t13 * 6270, n);
      p[6 * st] = DESCALE(z1 - t12 * 15137, n);
      z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * 9633;
      int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
      z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
      p[7 * st] = DESCALE(a4 + z1 + z3, n); p[5 * st] = DESCALE(a5 + z2 + z4, n);
      p[3 * st] = DESCALE(a6 + z2 + z3, n); p[st] = DESCALE(a7 + z1 + z4, n);
    }
  }
}
typedef struct { bbuf *b; uint32_t acc; int nbits; } bitw;
static inline void bw_put(bitw *w, unsigned code, int len) {
  if (!len) return;
  w->acc = (w->acc << len) | (code & ((1u << len) - 1));   /* acc never holds more than 7 + 16 bits */
  w->nbits += len;
  while (w->nbits >= 8) {
    uint8_t c = (uint8_t)(w->acc >> (w->nbits - 8));
    bb_push(w->b, c);
    if (c == 0xFF) bb_push(w->b, 0);
    w->nbits -= 8;
  }
}
static inline int bitlen(int v) { int n = 0; while (v) { n++; v >>= 1; } return n; }
static void huff_encode_block(bitw *w, const int16_t *zz /* zigzag order */, int *pred, const huff_enc *dc, const huff_enc *ac) {
  int diff = zz[0] - *pred; *pred = zz[0];
  int t = diff < 0 ? -diff : diff, v = diff < 0 ? diff - 1 : diff;
  int n = bitlen(t);
  bw_put(w, dc->code[n], dc->len[n]);
  bw_put(w, (unsigned)v, n);
  int r = 0;
  for (int k = 1; k < 64; k++) {
    int c = zz[k];
    if (c == 0) { r++; continue; }
    while (r > 15) { bw_put(w, ac->code[0xF0], ac->len[0xF0]); r -= 16; }
    t = c < 0 ? -c : c; v = c < 0 ? c - 1 : c; n = bitlen(t);
    int s = (r << 4) | n;
    bw_put(w, ac->code[s], ac->len[s]);
    bw_put(w, (unsigned)v, n);
    r = 0;
  }
  if (r > 0) bw_put(w, ac->code[0], ac->len[0]);
}
static void put16(bbuf *b, unsigned v) { bb_push(b, (uint8_t)(v >> 8)); bb_push(b, (uint8_t)v); }
static void jpeg_put_dht(bbuf *b, int tc_th, const uint8_t bits[16], const uint8_t *vals) {
  int n = 0; for (int i = 0; i < 16; i++) n += bits[i];
  put16(b, 0xFFC4); put16(b, 2 + 1 + 16 + n); bb_push(b, (uint8_t)tc_th);
  bb_write(b, bits, 16); bb_write(b, vals, n);
}
/* the 623-byte header libjpeg emits for these settings */
static void jpeg_write_header(bbuf *b, int w, int h, const uint16_t ql[64], const uint16_t qc[64]) {
  static const uint8_t app0[] = { 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 };
  put16(b, 0xFFD8); bb_write(b, app0, sizeof app0);
  for (int t = 0; t < 2; t++) {
    put16(b, 0xFFDB); put16(b, 67); bb_push(b, (uint8_t)t);
    for (int i = 0; i < 64; i++) bb_push(b, (uint8_t)(t ? qc[ZZ[i]] : ql[ZZ[i]]));
  }
  put16(b, 0xFFC0); put16(b, 17); bb_push(b, 8); put16(b, h); put16(b, w); bb_push(b, 3);
  bb_push(b, 1); bb_push(b, 0x22); bb_push(b, 0); bb_push(b, 2); bb_push(b, 0x11); bb_push(b, 1); bb_push(b, 3); bb_push(b, 0x11); bb_push(b, 1);
  jpeg_put_dht(b, 0x00, DC_L_BITS, DC_VALS); jpeg_put_dht(b, 0x10, AC_L_BITS, AC_L_VALS);
  jpeg_put_dht(b, 0x01, DC_C_BITS, DC_VALS); jpeg_put_dht(b, 0x11, AC_C_BITS, AC_C_VALS);
  put16(b, 0xFFDA); put16(b, 12); bb_push(b, 3); bb_push(b, 1); bb_push(b, 0x00); bb_push(b, 2); bb_push(b, 0x11); bb_push(b, 3); bb_push(b, 0x11);
  bb_push(b, 0); bb_push(b, 63); bb_push(b, 0);
}
static void quant_block(const int *d, const uint16_t q[64], int16_t *zz) {
  for (int k = 0; k < 64; k++) {
    int x = d[ZZ[k]], qv = q[ZZ[k]] << 3, neg = x < 0;
    if (neg) x = -x;
    x = (x + (qv >> 1)) / qv;
    zz[k] = (int16_t)(neg ? -x : x);
  }
}
int orc_jpeg_encode(const uint8_t *rgb, int w, int h, int quality, uint8_t **out, size_t *out_len) {
  if (w <= 0 || h <= 0) return -1;
  uint16_t ql[64], qc[64];
  jpeg_quant_table(QBASE_L, quality, ql); jpeg_quant_table(QBASE_C, quality, qc);
  huff_enc dcl, dcc, acl, acc;
  huff_build_enc(DC_L_BITS, DC_VALS, &dcl); huff_build_enc(DC_C_BITS, DC_VALS, &dcc);
  huff_build_enc(AC_L_BITS, AC_L_VALS, &acl); huff_build_enc(AC_C_BITS, AC_C_VALS, &acc);
  int mw = (w + 15) / 16, mh = (h + 15) / 16;
  int ybw = (w + 7) / 8, ybh = (h + 7) / 8;          /* real Y blocks */
  int cw = (w + 1) / 2, ch = (h + 1) / 2;            /* chroma plane, real size */
  int cbw = (cw + 7) / 8;
  int YW = mw * 16, YH = mh * 16, CW = mw * 8, CH = mh * 8;
  /* planes: Y padded by replication; chroma downsampled from the replicated full-res planes */
  uint8_t *Y = (uint8_t *)malloc((size_t)YW * YH);
  uint8_t *Cb = (uint8_t *)malloc((size_t)CW * CH), *Cr = (uint8_t *)malloc((size_t)CW * CH);
  int FW = 2 * 8 * cbw > YW ? 2 * 8 * cbw : YW;       /* full-res chroma width before downsampling */
  int FH = 2 * ch;
  uint8_t *fb = (uint8_t *)malloc((size_t)FW * FH), *fr = (uint8_t *)malloc((size_t)FW * FH);
  if (!Y || !Cb || !Cr || !fb || !fr) return -2;
  for (int y = 0; y < YH || y < FH; y++) {
    int sy = y < h ? y : h - 1;
    for (int x = 0; x < FW; x++) {
      int sx = x < w ? x : w - 1;
      const uint8_t *px = rgb + 3 * ((size_t)sy * w + sx);
      int r = px[0], g = px[1], b = px[2];
      if (y < YH && x < YW) Y[(size_t)y * YW + x] = (uint8_t)((19595 * r + 38470 * g + 7471 * b + 32768) >> 16);
      if (y < FH) {
        fb[(size_t)y * FW + x] = (uint8_t)((-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16);
        fr[(size_t)y * FW + x] = (uint8_t)((32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16);
      }
    }
  }
  for (int y = 0; y < CH; y++) {
    int sy = y < ch ? y : ch - 1;                     /* replicate last downsampled row */
    for (int x = 0; x < CW; x++) {
      int sx = x < 8 * cbw ? x : 8 * cbw - 1;         /* cannot trigger for 4:2:0 (8*cbw == CW or CW-8 only when dummy Y) */
      int bias = (sx & 1) ? 2 : 1;
      const uint8_t *p0 = fb + (size_t)(2 * sy) * FW + 2 * sx, *p1 = p0 + FW;
      Cb[(size_t)y * CW + x] = (uint8_t)((p0[0] + p0[1] + p1[0] + p1[1] + bias) >> 2);
      p0 = fr + (size_t)(2 * sy) * FW + 2 * sx; p1 = p0 + FW;
      Cr[(size_t)y * CW + x] = (uint8_t)((p0[0] + p0[1] + p1[0] + p1[1] + bias) >> 2);
    }
  }
  bbuf b = {0};
  jpeg_write_header(&b, w, h, ql, qc);
  bitw bw = { &b, 0, 0 };
  int pred[3] = { 0, 0, 0 };
  int d[64]; int16_t zz[64]; int16_t last_dc = 0;
  for (int my = 0; my < mh; my++) for (int mx = 0; mx < mw; mx++) {
    for (int blk = 0; blk < 4; blk++) {
      int bx = mx * 2 + (blk & 1), by = my * 2 + (blk >> 1);
      if (bx >= ybw || by >= ybh) {                   /* dummy block: DC of previous block in MCU order */
        memset(zz, 0, sizeof zz); zz[0] = last_dc;
      } else {
        for (int yy = 0; yy < 8; yy++) for (int xx = 0; xx < 8; xx++) d[yy * 8 + xx] = (int)Y[(size_t)(by * 8 + yy) * YW + bx * 8 + xx] - 128;
        fdct_islow(d); quant_block(d, ql, zz);
      }
      last_dc = zz[0];
      huff_encode_block(&bw, zz, &pred[0], &dcl, &acl);
    }
    for (int c = 0; c < 2; c++) {
      const uint8_t *P = c ? Cr : Cb;
      for (int yy = 0; yy < 8; yy++) for (int xx = 0; xx < 8; xx++) d[yy * 8 + xx] = (int)P[(size_t)(my * 8 + yy) * CW + mx * 8 + xx] - 128;
      fdct_islow(d); quant_block(d, qc, zz);
      last_dc = zz[0];
      huff_encode_block(&bw, zz, &pred[1 + c], &dcc, &acc);
    }
  }
  if (bw.nbits > 0) bw_put(&bw, 0x7F, 8 - bw.nbits);  /* pad with 1-bits */
  put16(&b, 0xFFD9);
  free(Y); free(Cb); free(Cr); free(fb); free(fr);
  *out = b.p; *out_len = b.n;
  return 0;
}

/* ---- decoder */
typedef struct { int mincode[17], maxcode[18], valptr[17]; uint8_t vals[256]; } huff_dec;
static void huff_build_dec(const uint8_t bits[16], const uint8_t *vals, int nvals, huff_dec *h) {
  int code = 0, k = 0;
  memcpy(h->vals, vals, (size_t)nvals);
  for (int l = 1; l <= 16; l++) {
    h->valptr[l] = k; h->mincode[l] = code;
    code += bits[l - 1]; k += bits[l - 1];
    h->maxcode[l] = bits[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  h->maxcode[17] = 0x7FFFFFFF;
}
typedef struct { const uint8_t *p; size_t n, pos; uint32_t acc; int nbits; int hit_marker; } bitr;
static inline int br_bit(bitr *r) {
  if (r->nbits == 0) {
    uint8_t c = 0;
    if (!r->hit_marker && r->pos < r->n) {
      c = r->p[r->pos];
      if (c == 0xFF) {
        if (r->pos + 1 < r->n && r->p[r->pos + 1] == 0) r->pos += 2;
        else { r->hit_marker = 1; c = 0; }
      } else r->pos++;
    }
    r->acc = c; r->nbits = 8;
  }
  r->nbits--;
  return (r->acc >> r->nbits) & 1;
}
static inline int br_bits(bitr *r, int n) { int v = 0; while (n--) v = (v << 1) | br_bit(r); return v; }
static int huff_decode_sym(bitr *r, const huff_dec *h) {
  int code = 0;
  for (int l = 1; l <= 16; l++) {
    code = (code << 1) | br_bit(r);
    if (h->maxcode[l] >= 0 && code <= h->maxcode[l] && code >= h->mincode[l]) return h->vals[h->valptr[l] + code - h->mincode[l]];
  }
  return 0;
}
static inline int jext(int v, int n) { return n == 0 ? 0 : (v < (1 << (n - 1)) ? v - (1 << n) + 1 : v); }
static inline uint8_t range_limit(int x) {             /* libjpeg sample_range_limit + CENTERJSAMPLE, index masked with 1023 */
  int v = x & 1023;
  if (v < 128) return (uint8_t)(v + 128);
  if (v < 512) return 255;
  if (v < 896) return 0;
  return (uint8_t)(v - 896);
}
static void idct_islow(const int *c /* dequantised, natural */, uint8_t *out, int stride) {
  int ws[64];
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < 8; k++) {
      const int *p = pass == 0 ? c + k : ws + 8 * k;
      int st = pass == 0 ? 8 : 1;
      int v0 = p[0], v1 = p[st], v2 = p[2 * st], v3 = p[3 * st], v4 = p[4 * st], v5 = p[5 * st], v6 = p[6 * st], v7 = p[7 * st];
      int z2 = v2, z3 = v6, z1 = (z2 + z3) * 4433;
      int t2 = z1 - z3 * 15137, t3 = z1 + z2 * 6270;
      int t0, t1;
      if (pass == 0) { t0 = (v0 + v4) * 8192; t1 = (v0 - v4) * 8192; }
      else { t0 = (v0 + v4) * 8192; t1 = (v0 - v4) * 8192; }
      int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
      int a0 = v7, a1 = v5, a2 = v3, a3 = v1;
      z1 = a0 + a3; z2 = a1 + a2; z3 = a0 + a2; int z4 = a1 + a3, z5 = (z3 + z4) * 9633;
      a0 *= 2446; a1 *= 16819; a2 *= 25172; a3 *= 12299;
      z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
      a0 += z1 + z3; a1 += z2 + z4; a2 += z2 + z3; a3 += z1 + z4;
      int n = pass == 0 ? 11 : 18;
      int o[8] = { DESCALE(t10 + a3, n), DESCALE(t11 + a2, n), DESCALE(t12 + a1, n), DESCALE(t13 + a0, n),
                   DESCALE(t13 - a0, n), DESCALE(t12 - a1, n), DESCALE(t11 - a2, n), DESCALE(t10 - a3, n) };
      if (pass == 0) for (int i = 0; i < 8; i++) ws[i * 8 + k] = o[i];
      else for (int i = 0; i < 8; i++) out[k * stride + i] = range_limit(o[i]);
    }
  }
}
int orc_jpeg_decode(const uint8_t *in, size_t len, uint8_t **rgb_out, int *wo, int *ho) {
  /* parse markers: only what libjpeg's own writer (above) can produce: baseline, 3 comps 2x2/1x1/1x1, no restarts */
  uint16_t q[4][64]; int have_q[4] = { 0 };
  uint8_t hb[2][2][16], hv[2][2][256]; int hn[2][2] = { { 0 } };
  int w = 0, h = 0; size_t pos = 2, scan = 0;
  if (len < 4 || in[0] != 0xFF || in[1] != 0xD8) return -1;
  while (pos + 4 <= len) {
    if (in[pos] != 0xFF) return -2;
    int m = in[pos + 1]; size_t L = ((size_t)in[pos + 2] << 8) | in[pos + 3];
    const uint8_t *s = in + pos + 4;
    if (pos + 2 + L > len) return -3;
    if (m == 0xDB) {
      size_t o = 0;
      while (o + 65 <= L - 2) { int t = s[o] & 15; if (s[o] >> 4) return -4; for (int i = 0; i < 64; i++) q[t][ZZ[i]] = s[o + 1 + i]; have_q[t] = 1; o += 65; }
    } else if (m == 0xC0) {
      h = (s[1] << 8) | s[2]; w = (s[3] << 8) | s[4];
      if (s[0] != 8 || s[5] != 3 || s[7] != 0x22 || s[10] != 0x11 || s[13] != 0x11) return -5;
    } else if (m == 0xC4) {
      size_t o = 0;
      while (o + 17 <= L - 2) {
        int tc = s[o] >> 4, th = s[o] & 15, n = 0;
        if (tc > 1 || th > 1) return -6;
        memcpy(hb[tc][th], s + o + 1, 16);
        for (int i = 0; i < 16; i++) n += s[o + 1 + i];
        memcpy(hv[tc][th], s + o + 17, (size_t)n); hn[tc][th] = n; o += 17 + (size_t)n;
      }
    } else if (m == 0xDA) { scan = pos + 2 + L; break; }
    pos += 2 + L;
  }
  if (!scan || w <= 0 || h <= 0 || !have_q[0] || !have_q[1]) return -7;
  huff_dec dcl, dcc, acl, acc;
  huff_build_dec(hb[0][0], hv[0][0], hn[0][0], &dcl); huff_build_dec(hb[0][1], hv[0][1], hn[0][1], &dcc);
  huff_build_dec(hb[1][0], hv[1][0], hn[1][0], &acl); huff_build_dec(hb[1][1], hv[1][1], hn[1][1], &acc);
  int mw = (w + 15) / 16, mh = (h + 15) / 16, YW = mw * 16, YH = mh * 16, CW = mw * 8, CH = mh * 8;
  int cw = (w + 1) / 2, ch = (h + 1) / 2;
  uint8_t *Y = (uint8_t *)malloc((size_t)YW * YH), *Cb = (uint8_t *)malloc((size_t)CW * CH), *Cr = (uint8_t *)malloc((size_t)CW * CH);
  uint8_t *rgb = (uint8_t *)malloc((size_t)w * h * 3);
  if (!Y || !Cb || !Cr || !rgb) return -8;
  bitr br = { in, len, scan, 0, 0, 0 };
  int pred[3] = { 0, 0, 0 }, coef[64];
  for (int my = 0; my < mh; my++) for (int mx = 0; mx < mw; mx++) {
    for (int blk = 0; blk < 6; blk++) {
      int comp = blk < 4 ? 0 : blk - 3;
      const huff_dec *dc = comp ? &dcc : &dcl, *ac = comp ? &acc : &acl;
      const uint16_t *qt = q[comp ? 1 : 0];
      memset(coef, 0, sizeof coef);
      int n = huff_decode_sym(&br, dc);
      int diff = jext(br_bits(&br, n), n);
      pred[comp] += diff; coef[0] = pred[comp] * qt[0];
      for (int k = 1; k < 64; k++) {
        int rs = huff_decode_sym(&br, ac), r = rs >> 4, s = rs & 15;
        if (s == 0) { if (r == 15) { k += 15; continue; } break; }
        k += r;
        if (k > 63) break;
        coef[ZZ[k]] = jext(br_bits(&br, s), s) * qt[ZZ[k]];
      }
      if (blk < 4) idct_islow(coef, Y + (size_t)(my * 16 + (blk >> 1) * 8) * YW + mx * 16 + (blk & 1) * 8, YW);
      else idct_islow(coef, (comp == 1 ? Cb : Cr) + (size_t)(my * 8) * CW + mx * 8, CW);
    }
  }
  /* upsample (fancy when the downsampled width > 2, else replication) + colour */
  int fancy = cw > 2;
  for (int y = 0; y < h; y++) {
    int cy = y >> 1, v = y & 1;
    int oy = v == 0 ? (cy > 0 ? cy - 1 : 0) : (cy + 1 < ch ? cy + 1 : ch - 1);
    for (int x = 0; x < w; x++) {
      int cx = x >> 1, cbv, crv;
      if (!fancy) { cbv = Cb[(size_t)cy * CW + cx]; crv = Cr[(size_t)cy * CW + cx]; }
      else {
        for (int c = 0; c < 2; c++) {
          const uint8_t *P = c ? Cr : Cb;
          int cs = 3 * P[(size_t)cy * CW + cx] + P[(size_t)oy * CW + cx], o;
          if (!(x & 1)) { if (cx == 0) o = (4 * cs + 8) >> 4; else { int l = 3 * P[(size_t)cy * CW + cx - 1] + P[(size_t)oy * CW + cx - 1]; o = (3 * cs + l + 8) >> 4; } }
          else { if (cx == cw - 1) o = (4 * cs + 7) >> 4; else { int r = 3 * P[(size_t)cy * CW + cx + 1] + P[(size_t)oy * CW + cx + 1]; o = (3 * cs + r + 7) >> 4; } }
          if (c) crv = o; else cbv = o;
        }
      }
      int yy = Y[(size_t)y * YW + x], cb = cbv - 128, cr = crv - 128;
      int R = yy + ((91881 * cr + 32768) >> 16), B = yy + ((116130 * cb + 32768) >> 16), G = yy + ((-22554 * cb - 46802 * cr + 32768) >> 16);
      uint8_t *o = rgb + 3 * ((size_t)y * w + x);
      o[0] = (uint8_t)(R < 0 ? 0 : R > 255 ? 255 : R); o[1] = (uint8_t)(G < 0 ? 0 : G > 255 ? 255 : G); o[2] = (uint8_t)(B < 0 ? 0 : B > 255 ? 255 : B);
    }
  }
  free(Y); free(Cb); free(Cr);
  *rgb_out = rgb; *wo = w; *ho = h;
  return 0;
}

/* ------------------------------------------------------------------ [PCL] OctreePointCloud bbox growth + keys */
typedef struct { double min[3], max[3]; uint32_t depth; int defined; double res; } obox;
static const double EPSF = 1.1920928955078125e-07;    /* (double) std::numeric_limits<float>::epsilon() */
static void get_key_bit_size_first(obox *b) {          /* getKeyBitSize() with leaf_count_ == 0 */
  uint32_t mk = 2;
  for (int a = 0; a < 3; a++) { uint32_t k = (uint32_t)ceil((b->max[a] - b->min[a] - EPSF) / b->res); if (k > mk) mk = k; }
  double d = ceil(log2((double)mk) - EPSF);
  b->depth = d > 32 ? 32 : (uint32_t)d;
  double side = (double)(1u << b->depth) * b->res;
  for (int a = 0; a < 3; a++) { double over = (side - (b->max[a] - b->min[a])) / 2.0; if (over > EPSF) { b->min[a] -= over; b->max[a] += over; } }
}
/* init != NULL: the tree starts with a box the caller defined ([PCL] defineBoundingBox -> getKeyBitSize, as simplifyPCloud and
 * generate_macroblock_tree do, impl.hpp:336,427); points outside still grow it. */
static int bbox_keys_impl(const void *pts, size_t n, double res, const obox *init, double bb_min[3], double bb_max[3], uint32_t *depth, uint32_t *keys, uint8_t *finite) {
  obox b; memset(&b, 0, sizeof b); b.res = res;
  if (init) { b = *init; b.res = res; }
  const uint8_t *base = (const uint8_t *)pts;
  size_t n_seen = 0;
  for (size_t i = 0; i < n; i++) {
    float pf[3]; memcpy(pf, base + 32 * i, 12);
    int fin = isfinite(pf[0]) && isfinite(pf[1]) && isfinite(pf[2]);   /* addPointsFromInputCloud: isFinite() */
    if (finite) finite[i] = (uint8_t)fin;
    if (!fin) continue;
    double p[3] = { pf[0], pf[1], pf[2] };
    for (;;) {                                                          /* adoptBoundingBoxToPoint */
      int lo[3], up[3], any = 0;
      for (int a = 0; a < 3; a++) { lo[a] = p[a] < b.min[a]; up[a] = p[a] >= b.max[a]; any |= lo[a] | up[a]; }
      if (!any && b.defined) break;
      if (b.defined) {
        if (b.depth >= 21) return -1;                                   /* oracle limit: Morton code in 63 bits */
        double side = (double)(1u << b.depth) * res;
        for (int a = 0; a < 3; a++) if (!up[a]) {
          b.min[a] -= side;
          /* re-rooting: old root becomes child with bit (!up) on this axis => earlier keys gain 1<<old_depth */
          for (size_t j = 0; j < i; j++) if (!finite || finite[j]) keys[3 * j + a] += 1u << b.depth;
        }
        b.depth++;
        side = (double)(1u << b.depth) * res - EPSF;
        for (int a = 0; a < 3; a++) b.max[a] = b.min[a] + side;
      } else {
        for (int a = 0; a < 3; a++) { b.min[a] = p[a] - res / 2; b.max[a] = p[a] + res / 2; }
        get_key_bit_size_first(&b);
        b.defined = 1;
      }
    }
    for (int a = 0; a < 3; a++) keys[3 * i + a] = (uint32_t)((p[a] - b.min[a]) / res);   /* genOctreeKeyforPoint */
    n_seen++;
  }
  for (int a = 0; a < 3; a++) { bb_min[a] = b.min[a]; bb_max[a] = b.max[a]; }
  *depth = b.depth;
  return n_seen ? 0 : 1;
}
int orc_bbox_keys(const void *pts, size_t n, double res, double bb_min[3], double bb_max[3], uint32_t *depth, uint32_t *keys, uint8_t *finite) {
  return bbox_keys_impl(pts, n, res, NULL, bb_min, bb_max, depth, keys, finite);
}

static inline uint64_t morton3(uint32_t x, uint32_t y, uint32_t z, uint32_t depth) {
  uint64_t m = 0;
  for (uint32_t l = 0; l < depth; l++) {
    uint32_t bit = depth - 1 - l;
    m = (m << 3) | (uint64_t)((((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1));
  }
  return m;
}
static inline void demorton3(uint64_t m, uint32_t depth, uint32_t k[3]) {
  k[0] = k[1] = k[2] = 0;
  for (uint32_t l = 0; l < depth; l++) {
    uint32_t c = (uint32_t)(m >> (3 * (depth - 1 - l))) & 7;
    k[0] = (k[0] << 1) | (c >> 2); k[1] = (k[1] << 1) | ((c >> 1) & 1); k[2] = (k[2] << 1) | (c & 1);
  }
}

/* recursive DFS = Octree2BufBase::serializeTreeRecursive over the implicit tree of sorted distinct leaf codes */
static void dfs_rec(const uint64_t *codes, size_t lo, size_t hi, uint32_t level, uint32_t depth, bbuf *out) {
  /* node at `level` covering leaves [lo,hi); children split on the 3 bits at shift 3*(depth-1-level) */
  uint32_t sh = 3 * (depth - 1 - level);
  uint8_t pat = 0; size_t i = lo;
  size_t start[9]; int cidx[8], nc = 0;
  while (i < hi) {
    int c = (int)((codes[i] >> sh) & 7); pat |= (uint8_t)(1 << c);
    cidx[nc] = c; start[nc++] = i;
    while (i < hi && (int)((codes[i] >> sh) & 7) == c) i++;
  }
  start[nc] = hi;
  bb_push(out, pat);
  if (level + 1 < depth) for (int k = 0; k < nc; k++) dfs_rec(codes, start[k], start[k + 1], level + 1, depth, out);
  (void)cidx;
}
int orc_dfs_recursive(const uint64_t *codes, size_t v, uint32_t depth, uint8_t **bytes, size_t *nbytes) {
  bbuf b = {0};
  if (v && depth) dfs_rec(codes, 0, v, 0, depth, &b);
  *bytes = b.p; *nbytes = b.n; return 0;
}

/* stable LSD radix sort of (code, idx) pairs; stability keeps point indices inside a leaf in input order,
 * which is what the leaf container's vector<int> holds ([PCL] OctreeContainerPointIndices::addPointIndex) */
static void sort_pairs(uint64_t *k, uint32_t *v, size_t n, uint32_t bits) {
  uint64_t *ka = k, *kb = (uint64_t *)malloc(n * 8 + 8);
  uint32_t *va = v, *vb = (uint32_t *)malloc(n * 4 + 4);
  for (uint32_t sh = 0; sh < bits; sh += 11) {
    size_t cnt[2049]; memset(cnt, 0, sizeof cnt);
    for (size_t i = 0; i < n; i++) cnt[((ka[i] >> sh) & 2047) + 1]++;
    for (int i = 0; i < 2048; i++) cnt[i + 1] += cnt[i];
    for (size_t i = 0; i < n; i++) { size_t d = cnt[(ka[i] >> sh) & 2047]++; kb[d] = ka[i]; vb[d] = va[i]; }
    uint64_t *tk = ka; ka = kb; kb = tk; uint32_t *tv = va; va = vb; vb = tv;
  }
  if (ka != k) { memcpy(k, ka, n * 8); memcpy(v, va, n * 4); free(ka); free(va); }
  else { free(kb); free(vb); }
}

/* ------------------------------------------------------------------ header (impl.hpp:1472-1486 + [PCL] base writeFrameHeader) */
static const char HDR_V2[] = "<PCL-OCT-CODECV2-COMPRESSED>";   /* point_cloud_codec_v2.h:371 */
static const char HDR_V1[] = "<PCL-OCT-COMPRESSED>";
#define HDR_BYTES 140
static void write_header(bbuf *os, const orc_params *p, uint32_t frame_id, int with_color, uint64_t point_count, const obox *b) {
  bb_write(os, HDR_V2, 28); bb_write(os, HDR_V1, 20);
  bb_write(os, &frame_id, 4);
  uint8_t f;
  f = 1; bb_write(os, &f, 1);                                   /* i_frame_: always true (impl.hpp:126-130) */
  f = (uint8_t)(p->do_voxel_grid != 0); bb_write(os, &f, 1);
  f = (uint8_t)(with_color != 0); bb_write(os, &f, 1);
  bb_write(os, &point_count, 8);
  double d = p->octree_resolution; bb_write(os, &d, 8);
  f = (uint8_t)p->color_bit_resolution; bb_write(os, &f, 1);    /* color_coder_.getBitDepth() */
  d = (double)(float)p->point_resolution; bb_write(os, &d, 8);  /* point_coder_.getPrecision() is float */
  bb_write(os, b->min, 24); bb_write(os, b->max, 24);
  f = (uint8_t)(p->do_centroid != 0); bb_write(os, &f, 1);
  f = (uint8_t)(p->code_connectivity != 0); bb_write(os, &f, 1);
  f = (uint8_t)(p->create_scalable != 0); bb_write(os, &f, 1);
  uint32_t cct = (uint32_t)p->color_coding_type; bb_write(os, &cct, 4);
  int32_t mb = p->macroblock_size; bb_write(os, &mb, 4);
  f = (uint8_t)(p->do_icp_color_offset != 0); bb_write(os, &f, 1);
}

/* ------------------------------------------------------------------ colour payloads (cjpeg.h:115-139,187-226,244-317) */
static int color_payload_snake(const uint8_t *avg, size_t v, int quality, bbuf *out) {
  int w = 256, h = (int)(v / 256) + 1;                           /* cjpeg.h:197-198 */
  size_t px = (size_t)w * h;
  uint8_t *lin = (uint8_t *)malloc(px * 3), *img = (uint8_t *)malloc(px * 3);
  memcpy(lin, avg, 3 * v);
  for (size_t i = v; i < px; i++) memcpy(lin + 3 * i, avg + 3 * (v - 1), 3);   /* pad with last colour cjpeg.h:203-213 */
  for (size_t i = 0; i < px; i++) memcpy(img + 3 * (size_t)orc_snake_pos_closed(w, h, (int64_t)i), lin + 3 * i, 3);  /* doMapping snake.h:105-118 */
  uint8_t *j; size_t jl;
  int rc = orc_jpeg_encode(img, w, h, quality, &j, &jl);
  if (!rc) { bb_write(out, j, jl); free(j); }
  free(lin); free(img);
  return rc;
}
static int color_payload_lines(const uint8_t *avg, size_t v, int quality, bbuf *out) {
  long num_lines = (long)(v / 2048);                             /* cjpeg.h:248-249 */
  uint32_t line_count = (uint32_t)(num_lines ? num_lines : 1);
  bb_write(out, &line_count, 4);                                  /* JPEGLineData::serialize cjpeg.h:72-83 */
  for (uint32_t i = 0; i < line_count; i++) {
    size_t off = (size_t)i * 2048, wpx = (i + 1 == line_count) ? v - off : 2048;
    uint8_t *j; size_t jl;
    int rc = orc_jpeg_encode(avg + 3 * off, (int)wpx, 1, quality, &j, &jl);
    if (rc) return rc;
    uint32_t l32 = (uint32_t)jl; bb_write(out, &l32, 4); bb_write(out, j, jl); free(j);
  }
  return 0;
}

/* ------------------------------------------------------------------ encodePointCloud (impl.hpp:80-213) */
int orc_encode(const orc_params *p, uint32_t frame_id, const void *pts, size_t n, uint8_t **out, size_t *out_len, orc_info *info, orc_debug *dbg) {
  *out = NULL; *out_len = 0;
  if (info) memset(info, 0, sizeof *info);
  if (dbg) memset(dbg, 0, sizeof *dbg);
  const int detail = !p->do_voxel_grid;                          /* doVoxelGridDownDownSampling=false: the class default (codec.h:108-143) */
  if (n == 0) return 0;
  double t0 = now_ms();
  const uint8_t *base = (const uint8_t *)pts;
  uint32_t *kxyz = (uint32_t *)malloc(n * 12 + 12); uint8_t *fin = (uint8_t *)malloc(n + 1);
  obox b; memset(&b, 0, sizeof b); b.res = p->octree_resolution;
  int rc = orc_bbox_keys(pts, n, p->octree_resolution, b.min, b.max, &b.depth, kxyz, fin);
  if (rc < 0) { free(kxyz); free(fin); return rc; }
  if (rc == 1) { free(kxyz); free(fin); return 0; }              /* leaf_count_ == 0: frame dropped (impl.hpp:206-212) */
  size_t nf = 0;
  uint64_t *codes = (uint64_t *)malloc(n * 8); uint32_t *idx = (uint32_t *)malloc(n * 4);
  for (size_t i = 0; i < n; i++) if (fin[i]) { codes[nf] = morton3(kxyz[3 * i], kxyz[3 * i + 1], kxyz[3 * i + 2], b.depth); idx[nf++] = (uint32_t)i; }
  free(kxyz); free(fin);
  double t1 = now_ms();
  sort_pairs(codes, idx, nf, 3 * b.depth);
  double t2 = now_ms();
  /* serializeTree (sort-based equivalent of the DFS) + leaf callbacks (impl.hpp:1509-1578) */
  int with_color = p->do_color != 0;                             /* PointXYZRGB always has an rgb field (impl.hpp:105-120) */
  int reduction = p->color_coding_type == 0 ? 8 - p->color_bit_resolution : 0;   /* jp_color_coder_ is never configured (App. C-3) */
  if (reduction < 0) reduction = 0;
  bbuf tree = {0}, avg = {0}, cen = {0}, pdiff = {0}, cdiff = {0};
  uint32_t *counts = detail ? (uint32_t *)malloc((nf ? nf : 1) * 4) : NULL;   /* point_count_data_vector_ (impl.hpp:1528) */
  const float pres_f = (float)p->point_resolution;              /* [PCL] PointCoding::setPrecision(float) */
  uint64_t *leaf_keys = (uint64_t *)malloc(nf * 8); size_t V = 0;
  uint8_t *outc = dbg ? (uint8_t *)calloc(nf ? nf : 1, 32) : NULL;   /* output_: one PointXYZRGB per leaf, DFS order (impl.hpp:1576) */
  uint32_t d = b.depth;
  /* pass 1: leaves, colours */
  for (size_t i = 0; i < nf;) {
    size_t j = i; while (j < nf && codes[j] == codes[i]) j++;
    leaf_keys[V++] = codes[i];
    if (detail) {                                                /* impl.hpp:1525-1541 */
      uint32_t k3d[3]; demorton3(codes[i], d, k3d);
      double corner[3];
      for (int a = 0; a < 3; a++) corner[a] = (double)k3d[a] * p->octree_resolution + b.min[a];
      counts[V - 1] = (uint32_t)(j - i);
      for (size_t k = i; k < j; k++) {                          /* [PCL] PointCoding::encodePoints: 3 bytes per point, index order */
        float pf[3]; memcpy(pf, base + 32 * (size_t)idx[k], 12);
        for (int a = 0; a < 3; a++) {
          int q = (int)(((double)pf[a] - corner[a]) / pres_f);
          if (q > 127) q = 127;
          if (q < -127) q = -127;
          bb_push(&pdiff, (uint8_t)q);
        }
      }
      if (with_color) {                                          /* [PCL] ColorCoding::encodePoints (jp_color_coder_ inherits it unchanged) */
        uint32_t s0 = 0, s1 = 0, s2 = 0, len = (uint32_t)(j - i);
        for (size_t k = i; k < j; k++) { uint32_t c; memcpy(&c, base + 32 * (size_t)idx[k] + 16, 4); s0 += c & 0xFF; s1 += (c >> 8) & 0xFF; s2 += (c >> 16) & 0xFF; }
        if (len > 1) {
          s0 /= len; s1 /= len; s2 /= len;
          for (size_t k = i; k < j; k++) {
            uint32_t c; memcpy(&c, base + 32 * (size_t)idx[k] + 16, 4);
            bb_push(&cdiff, (uint8_t)((uint8_t)((uint8_t)s0 ^ (uint8_t)(c & 0xFF)) >> reduction));
            bb_push(&cdiff, (uint8_t)((uint8_t)((uint8_t)s1 ^ (uint8_t)((c >> 8) & 0xFF)) >> reduction));
            bb_push(&cdiff, (uint8_t)((uint8_t)((uint8_t)s2 ^ (uint8_t)((c >> 16) & 0xFF)) >> reduction));
          }
        }
        s0 >>= reduction; s1 >>= reduction; s2 >>= reduction;
        bb_push(&avg, (uint8_t)s0); bb_push(&avg, (uint8_t)s1); bb_push(&avg, (uint8_t)s2);
      }
      i = j;
      continue;                                                  /* no centroid, no output_ point in detail mode */
    }
    if (with_color) {                                            /* [PCL] ColorCoding::encodeAverageOfPoints */
      uint32_t s0 = 0, s1 = 0, s2 = 0, len = (uint32_t)(j - i);
      for (size_t k = i; k < j; k++) { uint32_t c; memcpy(&c, base + 32 * (size_t)idx[k] + 16, 4); s0 += c & 0xFF; s1 += (c >> 8) & 0xFF; s2 += (c >> 16) & 0xFF; }
      if (len > 1) { s0 /= len; s1 /= len; s2 /= len; }
      s0 >>= reduction; s1 >>= reduction; s2 >>= reduction;
      bb_push(&avg, (uint8_t)s0); bb_push(&avg, (uint8_t)s1); bb_push(&avg, (uint8_t)s2);
    }
    float outp[3];
    {                                                            /* impl.hpp:1518-1520, 1559-1563: corner + 0.5 * resolution, assigned to float */
      uint32_t k3c[3]; demorton3(codes[i], d, k3c);
      for (int a = 0; a < 3; a++) { double corner = (double)k3c[a] * p->octree_resolution + b.min[a]; outp[a] = (float)(corner + 0.5 * p->octree_resolution); }
    }
    if (p->do_centroid) {                                        /* impl.hpp:1565-1573, pcv2.h:83-97, pcl::compute3DCentroid (float accumulation, index order) */
      uint32_t k3[3]; demorton3(codes[i], d, k3);
      float acc[3] = { 0, 0, 0 };
      for (size_t k = i; k < j; k++) { float pf[3]; memcpy(pf, base + 32 * (size_t)idx[k], 12); acc[0] += pf[0]; acc[1] += pf[1]; acc[2] += pf[2]; }
      float cnt = (float)(j - i);
      for (int a = 0; a < 3; a++) {
        float c = acc[a] / cnt;
        outp[a] = c;                                             /* impl.hpp:1569-1571 */
        double corner = (double)k3[a] * p->octree_resolution + b.min[a];
        int q = (int)(((double)c - corner) / 0.001f);          /* centroid_coder_ precision stays 0.001f (App. C-4) */
        if (q > 127) q = 127;
        if (q < -127) q = -127;
        bb_push(&cen, (uint8_t)q);
      }
    }
    if (outc) {                                                  /* PointXYZRGB(): data[3] = 1, r = g = b = 0, a = 255; colours = the bytes just pushed (impl.hpp:1552-1556) */
      uint8_t *o = outc + 32 * (V - 1);
      const float one = 1.0f;
      memcpy(o, outp, 12); memcpy(o + 12, &one, 4);
      if (with_color) { o[16] = avg.p[avg.n - 3]; o[17] = avg.p[avg.n - 2]; o[18] = avg.p[avg.n - 1]; }
      o[19] = 255;
    }
    i = j;
  }
  double t3 = now_ms();
  /* pass 2: occupancy bytes in DFS order: for leaf i emit the bytes of the branches it newly opens */
  {
    /* occupancy of a level-l node = OR of child digits over its leaves; compute per level by a linear sweep */
    /* node_byte[l] is accumulated lazily: we need the byte when the node is *opened*, i.e. before its later
     * children are seen, so first compute for every leaf and level the byte of its level-l ancestor. */
    /* do it level by level into a table indexed by "first leaf of node" */
    uint8_t **lvl = (uint8_t **)calloc(d, sizeof(uint8_t *));
    for (uint32_t l = 0; l < d; l++) lvl[l] = (uint8_t *)calloc(V, 1);
    for (uint32_t l = 0; l < d; l++) {
      uint32_t sh = 3 * (d - 1 - l);
      size_t first = 0;
      for (size_t i = 0; i < V; i++) {
        if (i > 0 && (l == 0 ? 0 : ((leaf_keys[i] >> (sh + 3)) != (leaf_keys[i - 1] >> (sh + 3))))) first = i;
        lvl[l][first] |= (uint8_t)(1u << ((leaf_keys[i] >> sh) & 7));
      }
    }
    for (size_t i = 0; i < V; i++) {
      uint32_t first_new = 0;
      if (i > 0) { uint64_t x = leaf_keys[i] ^ leaf_keys[i - 1]; int msb = 63 - __builtin_clzll(x); first_new = d - (uint32_t)(msb / 3); }
      for (uint32_t l = first_new; l < d; l++) bb_push(&tree, lvl[l][i]);
    }
    for (uint32_t l = 0; l < d; l++) free(lvl[l]);
    free(lvl);
  }
  double t4 = now_ms();
  /* colour payload (entropyEncoding -> getAverageDataVector, impl.hpp:1716, cjpeg.h:115-139) */
  bbuf col = {0};
  if (with_color) {
    if (p->color_coding_type == 1) rc = color_payload_snake(avg.p, V, p->jpeg_quality, &col);
    else if (p->color_coding_type == 2) rc = color_payload_lines(avg.p, V, p->jpeg_quality, &col);
    else bb_write(&col, avg.p, avg.n);                           /* type 0 (PCL) and 3 (GRID): raw averages */
    if (rc) return rc;
  }
  double t5 = now_ms();
  /* header + entropy mux (impl.hpp:175-178, 1682-1760) */
  bbuf os = {0};
  write_header(&os, p, frame_id, with_color, detail ? (uint64_t)nf : (uint64_t)V, &b);   /* [PCL] writeFrameHeader: leaf_count_ or object_count_ */
  uint64_t coded[3] = { 0, 0, 0 };
  uint64_t sz = tree.n; bb_write(&os, &sz, 8);
  rc_encode_to(&os, tree.p, tree.n, &coded[0]);
  if (p->do_centroid) { uint32_t c32 = (uint32_t)cen.n; bb_write(&os, &c32, 4); rc_encode_to(&os, cen.p, cen.n, &coded[1]); }
  if (with_color) { sz = col.n; bb_write(&os, &sz, 8); rc_encode_to(&os, col.p, col.n, &coded[2]); }
  if (detail) {                                                  /* impl.hpp:1728-1757 */
    uint64_t c2;
    sz = V; bb_write(&os, &sz, 8); rc_encode_int_to(&os, counts, V, &c2);
    sz = pdiff.n; bb_write(&os, &sz, 8); rc_encode_to(&os, pdiff.p, pdiff.n, &c2);
    if (with_color) { sz = cdiff.n; bb_write(&os, &sz, 8); rc_encode_to(&os, cdiff.p, cdiff.n, &c2); }
  }
  free(counts); free(pdiff.p); free(cdiff.p);
  double t6 = now_ms();
  if (info) {
    info->depth = d; memcpy(info->bb_min, b.min, 24); memcpy(info->bb_max, b.max, 24);
    info->n_finite = nf; info->n_leaves = V; info->n_tree_bytes = tree.n; info->n_color_bytes = col.n;
    memcpy(info->coded, coded, sizeof coded);
    info->t_ms[0] = t1 - t0; info->t_ms[1] = t2 - t1; info->t_ms[2] = t3 - t2; info->t_ms[3] = t4 - t3; info->t_ms[4] = t5 - t4; info->t_ms[5] = t6 - t5;
    info->t_ms[7] = t6 - t0;
  }
  if (dbg) { dbg->leaf_keys = leaf_keys; leaf_keys = NULL; dbg->tree_bytes = tree.p; tree.p = NULL; dbg->avg_colors = avg.p; avg.p = NULL;
             dbg->color_payload = col.p; col.p = NULL; dbg->centroid_bytes = cen.p; cen.p = NULL; dbg->output_cloud = outc; outc = NULL; }
  free(leaf_keys); free(tree.p); free(avg.p); free(cen.p); free(col.p); free(codes); free(idx); free(outc);
  *out = os.p; *out_len = os.n;
  return 0;
}

/* ------------------------------------------------------------------ decodePointCloud (impl.hpp:224-310) */
static int find_magic(const uint8_t *in, size_t len, size_t *pos) {   /* syncToHeader impl.hpp:1660-1676 */
  size_t hp = 0, i = 0;
  while (hp < 28) {
    if (i >= len) return -1;
    uint8_t c = in[i++];
    if (c == 0xFF) return -1;                                    /* (char)0xFF == EOF quirk, App. C-9 */
    if (c != (uint8_t)HDR_V2[hp++]) hp = ((uint8_t)HDR_V2[0] == c) ? 1 : 0;
  }
  hp = 0;
  while (hp < 20) {                                              /* [PCL] base syncToHeader */
    if (i >= len) return -1;
    uint8_t c = in[i++];
    if (c != (uint8_t)HDR_V1[hp++]) hp = ((uint8_t)HDR_V1[0] == c) ? 1 : 0;
  }
  *pos = i; return 0;
}
int orc_decode(const uint8_t *in, size_t len, void **pts_out, size_t *n_out, orc_info *info) {
  *pts_out = NULL; *n_out = 0;
  if (info) memset(info, 0, sizeof *info);
  double t0 = now_ms();
  size_t pos;
  if (find_magic(in, len, &pos)) return -1;
  if (pos + 92 > len) return -2;
  uint32_t frame_id; memcpy(&frame_id, in + pos, 4); pos += 4;
  uint8_t i_frame = in[pos++], vg = in[pos++], data_with_color = in[pos++];
  uint64_t point_count; memcpy(&point_count, in + pos, 8); pos += 8;
  double res; memcpy(&res, in + pos, 8); pos += 8;
  uint8_t color_bits = in[pos++];
  double pres; memcpy(&pres, in + pos, 8); pos += 8;
  double bmin[3], bmax[3]; memcpy(bmin, in + pos, 24); pos += 24; memcpy(bmax, in + pos, 24); pos += 24;
  uint8_t do_centroid = in[pos++]; pos += 2;                     /* connectivity, scalable: carried, unused */
  uint32_t cct; memcpy(&cct, in + pos, 4); pos += 4;
  pos += 4 + 1;                                                  /* macroblock_size, do_icp_color_offset */
  (void)i_frame; (void)vg; (void)frame_id;
  /* [PCL] readFrameHeader -> defineBoundingBox -> getKeyBitSize (App. B.3) */
  uint32_t mk = 2;
  for (int a = 0; a < 3; a++) { uint32_t k = (uint32_t)ceil((bmax[a] - bmin[a] - EPSF) / res); if (k > mk) mk = k; }
  double dd = ceil(log2((double)mk) - EPSF);
  uint32_t d = dd > 32 ? 32 : (uint32_t)dd;
  if (d > 21) return -3;
  /* entropyDecoding impl.hpp:1766-1835 */
  if (pos + 8 > len) return -4;
  uint64_t B; memcpy(&B, in + pos, 8); pos += 8;
  uint8_t *tree = (uint8_t *)malloc(B + 1); size_t used;
  int rc = orc_range_decode(in + pos, len - pos, tree, B, &used);
  if (rc) { free(tree); return rc; }
  pos += used;
  uint8_t *cen = NULL; uint32_t ncen = 0;
  if (do_centroid) {
    if (pos + 4 > len) return -5;
    memcpy(&ncen, in + pos, 4); pos += 4;
    cen = (uint8_t *)malloc(ncen + 1);
    rc = orc_range_decode(in + pos, len - pos, cen, ncen, &used); if (rc) return rc;
    pos += used;
  }
  uint8_t *col = NULL; uint64_t ncol = 0;
  if (data_with_color) {
    if (pos + 8 > len) return -6;
    memcpy(&ncol, in + pos, 8); pos += 8;
    col = (uint8_t *)malloc(ncol + 1);
    rc = orc_range_decode(in + pos, len - pos, col, ncol, &used); if (rc) return rc;
    pos += used;
  }
  /* trailing data = the enhancement vectors: the reference switches to detail mode on peek() (impl.hpp:1802-1806) */
  const int detail = pos != len;
  uint32_t *counts = NULL; uint64_t ncounts = 0; uint8_t *pdiff = NULL; uint64_t npdiff = 0; uint8_t *cdiff = NULL; uint64_t ncdiff = 0;
  if (detail) {                                                  /* impl.hpp:1808-1832 */
    if (pos + 8 > len) return -12;
    memcpy(&ncounts, in + pos, 8); pos += 8;
    if (ncounts > (1ull << 30)) return -12;
    counts = (uint32_t *)malloc((ncounts + 1) * 4);
    rc = orc_range_decode_int(in + pos, len - pos, counts, ncounts, &used); if (rc) return rc;
    pos += used;
    if (pos + 8 > len) return -13;
    memcpy(&npdiff, in + pos, 8); pos += 8;
    if (npdiff > (1ull << 34)) return -13;
    pdiff = (uint8_t *)malloc(npdiff + 1);
    rc = orc_range_decode(in + pos, len - pos, pdiff, npdiff, &used); if (rc) return rc;
    pos += used;
    if (data_with_color) {
      if (pos + 8 > len) return -14;
      memcpy(&ncdiff, in + pos, 8); pos += 8;
      if (ncdiff > (1ull << 34)) return -14;
      cdiff = (uint8_t *)malloc(ncdiff + 1);
      rc = orc_range_decode(in + pos, len - pos, cdiff, ncdiff, &used); if (rc) return rc;
      pos += used;
      /* The reference reads these diffs into color_coder_ (impl.hpp:1828) but decodes JPEG-type colours with
       * jp_color_coder_, whose diff vector is empty: undefined behaviour.  Only type 0 is defined. */
      if (cct != 0) return -15;
    }
  }
  double t1 = now_ms();
  /* initializeDecoding cjpeg.h:150-172 */
  uint8_t *avg = NULL; size_t navg = 0;
  int reduction = 0;
  if (data_with_color) {
    if (cct == 1) {
      uint8_t *img; int w, h;
      rc = orc_jpeg_decode(col, ncol, &img, &w, &h); if (rc) return rc;
      navg = (size_t)w * h * 3; avg = (uint8_t *)malloc(navg);
      for (size_t i = 0; i < (size_t)w * h; i++) memcpy(avg + 3 * i, img + 3 * (size_t)orc_snake_pos_closed(w, h, (int64_t)i), 3);  /* undoSnakeGridMapping snake.h:123-137 */
      free(img);
    } else if (cct == 2) {
      uint32_t lc; memcpy(&lc, col, 4); size_t o = 4; bbuf acc = {0};
      for (uint32_t i = 0; i < lc; i++) {
        uint32_t l32; memcpy(&l32, col + o, 4); o += 4;
        uint8_t *img; int w, h;
        rc = orc_jpeg_decode(col + o, l32, &img, &w, &h); if (rc) return rc;
        bb_write(&acc, img, (size_t)w * h * 3); free(img); o += l32;
      }
      avg = acc.p; navg = acc.n;
    } else {
      avg = col; navg = ncol; col = NULL;
      if (cct == 0) reduction = 8 - color_bits;
    }
  }
  double t2 = now_ms();
  /* deserializeTree (App. B.3): iterative DFS over the byte stream */
  uint8_t *outp = (uint8_t *)calloc(point_count ? point_count : 1, 32);
  size_t np = 0, bp = 0, nleaf = 0, pd = 0, cd = 0;
  const float pres_f = (float)pres;                              /* [PCL] readFrameHeader: point_coder_.setPrecision (static_cast<float> (point_resolution)) */
  {
    uint8_t mask[32]; uint32_t level = 0; uint64_t code = 0;
    if (B > 0) {
      mask[0] = tree[bp++];
      for (;;) {
        if (mask[level] == 0) { if (level == 0) break; level--; code >>= 3; continue; }
        int c = __builtin_ctz(mask[level]); mask[level] &= (uint8_t)(mask[level] - 1);
        uint64_t child = (code << 3) | (uint64_t)c;
        if (level + 1 < d) {
          if (bp >= B) { free(outp); return -7; }
          level++; code = child; mask[level] = tree[bp++];
        } else if (detail) {                                     /* impl.hpp:1592-1613 + colour :1639-1651 */
          if (nleaf >= ncounts) { free(outp); return -8; }
          uint32_t cnt = counts[nleaf], k3[3]; demorton3(child, d, k3);
          if (np + cnt > point_count || pd + 3ull * cnt > npdiff) { free(outp); return -8; }
          uint8_t av[3] = { 0, 0, 0 };
          if (data_with_color) {
            if (3 * nleaf + 2 >= navg) { free(outp); return -9; }
            for (int k = 0; k < 3; k++) av[k] = (uint8_t)(avg[3 * nleaf + k] << reduction);
            if (cnt > 1 && cd + 3ull * cnt > ncdiff) { free(outp); return -9; }
          }
          for (uint32_t q = 0; q < cnt; q++) {
            uint8_t *o = outp + 32 * np;
            float xyz[3];
            for (int a = 0; a < 3; a++) {                        /* [PCL] PointCoding::decodePoints */
              double corner = (double)k3[a] * res + bmin[a];
              xyz[a] = (float)(corner + pdiff[pd++] * pres_f);
            }
            memcpy(o, xyz, 12);
            float one = 1.0f; memcpy(o + 12, &one, 4);
            uint32_t rgba;
            if (data_with_color) {                               /* [PCL] ColorCoding::decodePoints */
              if (cnt > 1) {
                uint8_t df[3]; for (int k = 0; k < 3; k++) df[k] = (uint8_t)(cdiff[cd++] << reduction);
                rgba = (uint32_t)(av[0] ^ df[0]) | ((uint32_t)(av[1] ^ df[1]) << 8) | ((uint32_t)(av[2] ^ df[2]) << 16);
              } else rgba = (uint32_t)av[0] | ((uint32_t)av[1] << 8) | ((uint32_t)av[2] << 16);
            } else rgba = 0x00FFFFFFu;
            memcpy(o + 16, &rgba, 4);
            np++;
          }
          nleaf++;
        } else {
          if (np >= point_count) { free(outp); return -8; }
          uint32_t k3[3]; demorton3(child, d, k3);
          float xyz[3];
          for (int a = 0; a < 3; a++) {
            if (do_centroid) {                                   /* pcv2.h:103-118 */
              double corner = (double)k3[a] * res + bmin[a];
              xyz[a] = (float)(corner + cen[3 * np + a] * 0.001f);
            } else xyz[a] = (float)(((double)k3[a] + 0.5) * res + bmin[a]);   /* impl.hpp:1630-1632 */
          }
          uint8_t *o = outp + 32 * np;
          memcpy(o, xyz, 12);
          float one = 1.0f; memcpy(o + 12, &one, 4);             /* PointXYZRGB ctor: data[3] = 1.0f */
          uint32_t rgba;
          if (data_with_color) {                                 /* [PCL] ColorCoding::decodePoints */
            if (3 * np + 2 >= navg) { free(outp); return -9; }
            rgba = ((uint32_t)avg[3 * np] << reduction & 0xFF) | (((uint32_t)avg[3 * np + 1] << reduction & 0xFF) << 8) | (((uint32_t)avg[3 * np + 2] << reduction & 0xFF) << 16);
          } else rgba = 0x00FFFFFFu;                             /* setDefaultColor: white, alpha 0 */
          memcpy(o + 16, &rgba, 4);
          np++;
        }
      }
    }
  }
  double t3 = now_ms();
  if (info) {
    info->depth = d; memcpy(info->bb_min, bmin, 24); memcpy(info->bb_max, bmax, 24);
    info->n_leaves = np; info->n_tree_bytes = B; info->n_color_bytes = ncol;
    info->t_ms[0] = t1 - t0; info->t_ms[1] = t2 - t1; info->t_ms[2] = t3 - t2; info->t_ms[7] = t3 - t0;
  }
  free(tree); free(cen); free(col); free(avg); free(counts); free(pdiff); free(cdiff);
  *pts_out = outp; *n_out = np;
  return 0;
}

/* ------------------------------------------------------------------ computeQualityMetric
 * apps/evaluate_compression/include/pcl/apps/evaluate_compression/impl/quality_metrics_impl.hpp:63-70 (YUV), :82-239.
 * The reference asks two kd-trees for nearest neighbours; this restatement searches exhaustively (exact, O(na * nb):
 * test sizes only).  Distances: float sum of float squares in x, y, z order ([PCL] KdTreeFLANN, FLANN L2_Simple). */
static void yuv_of(const uint8_t *rec, float yuv[3]) {
  double b = rec[16], g = rec[17], r = rec[18];
  yuv[0] = (float)((0.299 * r + 0.587 * g + 0.114 * b) / 255.0);
  yuv[1] = (float)((-0.147 * r - 0.289 * g + 0.436 * b) / 255.0);
  yuv[2] = (float)((0.615 * r - 0.515 * g - 0.100 * b) / 255.0);
}
static void nn_pass(const uint8_t *q, size_t nq, const uint8_t *t, size_t nt, double *sum, float *mx, double mse[3]) {
  *sum = 0; *mx = -3.4028235e38f; if (mse) mse[0] = mse[1] = mse[2] = 0;
  for (size_t i = 0; i < nq; i++) {
    float p[3]; memcpy(p, q + 32 * i, 12);
    if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
    float best = 3.0e38f; size_t bj = (size_t)-1;
    for (size_t j = 0; j < nt; j++) {
      float c[3]; memcpy(c, t + 32 * j, 12);
      if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) continue;
      float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
      float d = dx * dx; d += dy * dy; d += dz * dz;
      if (d < best) { best = d; bj = j; }
    }
    if (bj == (size_t)-1) continue;
    if (best > *mx) *mx = best;
    *sum += best;
    if (mse) {
      float a[3], b[3]; yuv_of(q + 32 * i, a); yuv_of(t + 32 * bj, b);
      for (int k = 0; k < 3; k++) mse[k] += (a[k] - b[k]) * (a[k] - b[k]);
    }
  }
}
int orc_quality_metrics(const void *cloud_a, size_t na, const void *cloud_b, size_t nb, orc_quality *out) {
  memset(out, 0, sizeof *out);
  out->in_point_count = na; out->out_point_count = nb;
  if (!na || !nb) return 0;
  double sa, sb, mse[3]; float ma, mb;
  nn_pass((const uint8_t *)cloud_a, na, (const uint8_t *)cloud_b, nb, &sa, &ma, mse);
  nn_pass((const uint8_t *)cloud_b, nb, (const uint8_t *)cloud_a, na, &sb, &mb, NULL);
  ma = sqrtf(ma); mb = sqrtf(mb);
  double ra = sqrt(sa / (double)na), rb = sqrt(sb / (double)nb);
  float dist_h = ma > mb ? ma : mb, dist_rms = (float)(ra > rb ? ra : rb);
  float mxs[3] = { -3.4028235e38f, -3.4028235e38f, -3.4028235e38f };        /* getMinMax3D(cloud_a): finite points */
  for (size_t i = 0; i < na; i++) {
    float p[3]; memcpy(p, (const uint8_t *)cloud_a + 32 * i, 12);
    if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
    for (int k = 0; k < 3; k++) if (p[k] > mxs[k]) mxs[k] = p[k];
  }
  float energy = mxs[0] * mxs[0] + mxs[1] * mxs[1] + mxs[2] * mxs[2];
  out->left_hausdorff = ma; out->right_hausdorff = mb; out->symm_hausdorff = dist_h;
  out->left_rms = (float)ra; out->right_rms = (float)rb; out->symm_rms = dist_rms;
  out->psnr_db = (float)(10 * log10(energy / (dist_rms * dist_rms)));
  for (int k = 0; k < 3; k++) out->psnr_yuv[k] = 10 * log10(1.0 / (mse[k] / (double)na));
  return 0;
}

#include "ccv2_oracle_inter.c"

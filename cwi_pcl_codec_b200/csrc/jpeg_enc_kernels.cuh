// jpeg_enc_kernels.cuh -- colour layer, encode side: snake-grid gather (snake.h:46-71,105-118, closed form
// SURVEY App. B.7) fused with libjpeg's baseline pipeline exactly as jpeg_io drives it (jpeg_io.hpp:259-311):
// RGB->YCbCr, h2v2 downsample, ISLOW FDCT, quantise, standard-table Huffman, byte stuffing.
#pragma once
#include "common.cuh"

#define JPEG_HDR_BYTES 623
#define JDESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

// inverse of the snake mapping: pixel (x, y) of the w=256 image -> linear voxel index
__device__ __forceinline__ uint32_t snake_inverse_256(uint32_t x, uint32_t y, uint32_t h) {
  const uint32_t nbw = 32, F = h >> 3;
  uint32_t mh = y >> 3, j = x >> 3, col = x & 7, r, R, base;
  if (mh < F) { R = 8; r = y & 7; base = mh * nbw * 64 + j * 64; }
  else { R = h - 8 * F; r = y - 8 * F; base = F * nbw * 64 + j * 8 * R; }
  uint32_t c = ((j * R + r) & 1) ? 7 - col : col;
  return base + r * 8 + c;
}
// forward mapping: linear index -> pixel offset (x + y*256)
__device__ __forceinline__ uint32_t snake_forward_256(uint32_t i, uint32_t h) {
  const uint32_t nbw = 32, F = h >> 3, full = F * nbw * 64;
  uint32_t mh, j, k, R;
  if (i < full) { mh = i / (nbw * 64); j = (i % (nbw * 64)) >> 6; k = i & 63; R = 8; }
  else { mh = F; R = h - 8 * F; uint32_t ip = i - full; j = ip / (8 * R); k = ip % (8 * R); }
  uint32_t r = k >> 3, c = k & 7;
  uint32_t col = ((j * R + r) & 1) ? 7 - c : c;
  return col + (r + 8 * mh) * 256 + 8 * j;
}

__device__ __forceinline__ void fdct8(int *v, bool first) {
  int t0 = v[0] + v[7], t7 = v[0] - v[7], t1 = v[1] + v[6], t6 = v[1] - v[6];
  int t2 = v[2] + v[5], t5 = v[2] - v[5], t3 = v[3] + v[4], t4 = v[3] - v[4];
  int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  const int n = first ? 11 : 15;
  if (first) { v[0] = (t10 + t11) * 4; v[4] = (t10 - t11) * 4; }
  else { v[0] = JDESCALE(t10 + t11, 2); v[4] = JDESCALE(t10 - t11, 2); }
  int z1 = (t12 + t13) * 4433;
  v[2] = JDESCALE(z1 + t13 * 6270, n);
  v[6] = JDESCALE(z1 - t12 * 15137, n);
  z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * 9633;
  int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
  v[7] = JDESCALE(a4 + z1 + z3, n); v[5] = JDESCALE(a5 + z2 + z4, n);
  v[3] = JDESCALE(a6 + z2 + z3, n); v[1] = JDESCALE(a7 + z1 + z4, n);
}

// One CTA (256 threads) per 16x16 MCU of the 256-wide snake image.
__global__ void __launch_bounds__(256) jpeg_mcu_kernel(EncFrame *frames, const JpegTables *T) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  if (V == 0) return;
  const uint32_t h = f.img_h, mcu_h = f.mcu_h;
  const uint32_t mcu = blockIdx.x;
  if (mcu >= mcu_h * 16) return;
  const uint32_t mx = mcu & 15, my = mcu >> 4;
  __shared__ int sY[16][17], sCb[16][17], sCr[16][17];
  __shared__ int work[6][64];
  __shared__ short outc[6][64];
  const uint32_t t = threadIdx.x, px = t & 15, py = t >> 4;
  const uint32_t ch = (h + 1) >> 1;
  {
    uint32_t x = mx * 16 + px, y = my * 16 + py;
    uint32_t yy = min(y, h - 1);
    uint32_t cyg = min(y >> 1, ch - 1);
    uint32_t yc = min(2 * cyg + (py & 1), h - 1);
    uint32_t i = min(snake_inverse_256(x, yy, h), V - 1);       // padding pixels take the last colour (cjpeg.h:203-213)
    const uint8_t *c = f.avg + 3ull * i;
    int r = c[0], g = c[1], b = c[2];
    sY[py][px] = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
    if (yc != yy) { i = min(snake_inverse_256(x, yc, h), V - 1); c = f.avg + 3ull * i; r = c[0]; g = c[1]; b = c[2]; }
    sCb[py][px] = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
    sCr[py][px] = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
  }
  __syncthreads();
  // stage samples (minus 128) for the six blocks
  {
    uint32_t blk = (py >> 3) * 2 + (px >> 3);
    work[blk][(py & 7) * 8 + (px & 7)] = sY[py][px] - 128;
    if (t < 128) {
      uint32_t c = t >> 6, k = t & 63, cx = k & 7, cy = k >> 3;
      int (*p)[17] = c ? sCr : sCb;
      int bias = (cx & 1) ? 2 : 1;
      work[4 + c][k] = ((p[2 * cy][2 * cx] + p[2 * cy][2 * cx + 1] + p[2 * cy + 1][2 * cx] + p[2 * cy + 1][2 * cx + 1] + bias) >> 2) - 128;
    }
  }
  __syncthreads();
  if (t < 48) {                                        // row pass
    int *p = &work[t >> 3][(t & 7) * 8], v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = p[k];
    fdct8(v, true);
#pragma unroll
    for (int k = 0; k < 8; k++) p[k] = v[k];
  }
  __syncthreads();
  if (t < 48) {                                        // column pass
    int *p = &work[t >> 3][t & 7], v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = p[8 * k];
    fdct8(v, false);
#pragma unroll
    for (int k = 0; k < 8; k++) p[8 * k] = v[k];
  }
  __syncthreads();
  for (uint32_t e = t; e < 384; e += 256) {            // quantise, zigzag order
    uint32_t blk = e >> 6, k = e & 63, nat = T->zz[k];
    int x = work[blk][nat], q = (int)T->q[blk >= 4][nat] << 3;
    bool neg = x < 0; if (neg) x = -x;
    x = (x + (q >> 1)) / q;
    outc[blk][k] = (short)(neg ? -x : x);
  }
  __syncthreads();
  // dummy Y blocks below the image: all zero, DC copied from the previous block in MCU order (jccoefct.c)
  const uint32_t ybh = (h + 7) >> 3;
  if (my * 2 + 1 >= ybh && t < 128) { uint32_t blk = 2 + (t >> 6), k = t & 63; outc[blk][k] = k == 0 ? outc[1][0] : (short)0; }
  __syncthreads();
  short *dst = f.coef + (size_t)mcu * 384;
  for (uint32_t e = t; e < 384; e += 256) dst[e] = (&outc[0][0])[e];
}

// ---- Huffman: thread per block; length pass -> chained scan -> write pass
struct BitSink {
  uint32_t *buf; uint32_t cap_words; uint64_t acc; uint32_t nb; uint32_t word; bool first; uint32_t *err;
  __device__ __forceinline__ void flush_word(uint32_t w, bool shared_word) {
    if (word < cap_words) {
      uint32_t v = __byte_perm(w, 0, 0x0123);
      if (shared_word) atomicOr(&buf[word], v); else buf[word] = v;
    } else if (err) *err = 1;
    word++;
  }
  __device__ __forceinline__ void put(uint32_t code, uint32_t len) {
    acc = (acc << len) | (code & ((1u << len) - 1)); nb += len;
    if (nb >= 32) { flush_word((uint32_t)(acc >> (nb - 32)), first); first = false; nb -= 32; }
  }
  __device__ __forceinline__ void finish() { if (nb) flush_word((uint32_t)(acc << (32 - nb)), true); }
};
template <bool WRITE>
__device__ __forceinline__ uint32_t huff_block(const short *zz, int pred, const JpegTables *T, int tsel, BitSink *sink) {
  uint32_t bits = 0;
  int diff = zz[0] - pred;
  int tv = diff < 0 ? -diff : diff, v = diff < 0 ? diff - 1 : diff;
  uint32_t n = tv ? 32 - __clz(tv) : 0;
  bits += T->dc_len[tsel][n] + n;
  if (WRITE) { sink->put(T->dc_code[tsel][n], T->dc_len[tsel][n]); if (n) sink->put((uint32_t)v, n); }
  uint32_t r = 0;
  for (int k = 1; k < 64; k++) {
    int c = zz[k];
    if (c == 0) { r++; continue; }
    while (r > 15) { bits += T->ac_len[tsel][0xF0]; if (WRITE) sink->put(T->ac_code[tsel][0xF0], T->ac_len[tsel][0xF0]); r -= 16; }
    tv = c < 0 ? -c : c; v = c < 0 ? c - 1 : c; n = 32 - __clz(tv);
    uint32_t s = (r << 4) | n;
    bits += T->ac_len[tsel][s] + n;
    if (WRITE) { sink->put(T->ac_code[tsel][s], T->ac_len[tsel][s]); sink->put((uint32_t)v, n); }
    r = 0;
  }
  if (r > 0) { bits += T->ac_len[tsel][0]; if (WRITE) sink->put(T->ac_code[tsel][0], T->ac_len[tsel][0]); }
  return bits;
}

#define HUFF_THREADS 128
__global__ void __launch_bounds__(HUFF_THREADS) jpeg_huff_kernel(EncFrame *frames, const JpegTables *Tg) {
  EncFrame &f = frames[blockIdx.y];
  if (f.V == 0) return;
  const uint32_t nblocks = f.mcu_h * 16 * 6;
  const uint32_t ntiles = (nblocks + HUFF_THREADS - 1) / HUFF_THREADS;
  if (blockIdx.x >= ntiles) return;
  __shared__ JpegTables T;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  for (uint32_t k = threadIdx.x; k < sizeof(JpegTables) / 4; k += blockDim.x) ((uint32_t *)&T)[k] = ((const uint32_t *)Tg)[k];
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_HUFF], 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t g = tile * HUFF_THREADS + threadIdx.x;
  __align__(16) short zz[64]; int pred = 0; uint32_t bits = 0; int tsel = 0;
  if (g < nblocks) {
    const uint32_t mcu = g / 6, blk = g % 6;
    const uint4 *src = (const uint4 *)(f.coef + (size_t)g * 64);
#pragma unroll
    for (int k = 0; k < 8; k++) ((uint4 *)zz)[k] = src[k];
    tsel = blk >= 4;
    if (blk >= 4) { if (mcu > 0) pred = f.coef[(size_t)(g - 6) * 64]; }
    else if (blk > 0) pred = f.coef[(size_t)(g - 1) * 64];
    else if (mcu > 0) pred = f.coef[(size_t)(g - 3) * 64];          // Y3 of the previous MCU
    bits = huff_block<false>(zz, pred, &T, tsel, nullptr);
  }
  uint64_t tot;
  uint64_t excl = block_excl_scan_u64(bits, &tot, s_scan);
  if (threadIdx.x < 32) { uint64_t e = scan_lookback(f.scan_status + f.scan_tiles_max, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  __syncthreads();
  excl += s_excl;
  if (g < nblocks) {
    BitSink sk; sk.buf = f.jbits_buf; sk.cap_words = f.jbits_cap_words; sk.acc = 0; sk.nb = (uint32_t)(excl & 31); sk.word = (uint32_t)(excl >> 5);
    sk.first = true; sk.err = nullptr;
    huff_block<true>(zz, pred, &T, tsel, &sk);
    sk.finish();
    if (sk.word > sk.cap_words) atomicOr(&f.error, FERR_JPEG_CAP);
    if (g == nblocks - 1) f.jbits = (uint32_t)(excl + bits);
  }
}

// ---- byte stuffing + file assembly: header | stuffed entropy bytes | EOI
#define STUFF_THREADS 256
#define STUFF_BYTES 16
__global__ void __launch_bounds__(STUFF_THREADS) jpeg_stuff_kernel(EncFrame *frames, const JpegTables *Tg) {
  EncFrame &f = frames[blockIdx.y];
  if (f.V == 0 || (f.error & FERR_JPEG_CAP)) return;
  const uint32_t tbits = f.jbits;
  const uint32_t U = (tbits + 7) >> 3;
  const uint32_t per_tile = STUFF_THREADS * STUFF_BYTES;
  const uint32_t ntiles = max(1u, (U + per_tile - 1) / per_tile);
  if (blockIdx.x >= ntiles) return;
  __shared__ uint32_t s_tile; __shared__ uint64_t s_scan[33]; __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&f.ticket[TK_STUFF], 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t b0 = tile * per_tile + threadIdx.x * STUFF_BYTES;
  uint8_t by[STUFF_BYTES]; uint32_t cnt = 0, nvalid = 0;
  if (b0 < U) {
    uint4 v = *(const uint4 *)((const uint8_t *)f.jbits_buf + b0);
    *(uint4 *)by = v;
    nvalid = min((uint32_t)STUFF_BYTES, U - b0);
    if (b0 + nvalid == U && (tbits & 7)) by[nvalid - 1] |= (uint8_t)((1u << (8 - (tbits & 7))) - 1);   // pad with 1-bits
    for (uint32_t k = 0; k < nvalid; k++) cnt += 1 + (by[k] == 0xFF);
  }
  uint64_t tot;
  uint64_t excl = block_excl_scan_u64(cnt, &tot, s_scan);
  if (threadIdx.x < 32) { uint64_t e = scan_lookback(f.scan_status + 2 * (size_t)f.scan_tiles_max, tile, tot); if (threadIdx.x == 0) s_excl = e; }
  __syncthreads();
  excl += s_excl;
  uint8_t *out = f.cpay;
  const uint64_t total_if_last = JPEG_HDR_BYTES + excl + cnt + 2;
  if (b0 < U) {
    uint64_t o = JPEG_HDR_BYTES + excl;
    if (o + cnt + 2 > f.cpay_cap) { atomicOr(&f.error, FERR_JPEG_CAP); }
    else {
      for (uint32_t k = 0; k < nvalid; k++) { out[o++] = by[k]; if (by[k] == 0xFF) out[o++] = 0; }
      if (b0 + nvalid == U) { out[o] = 0xFF; out[o + 1] = 0xD9; f.J = (uint32_t)total_if_last; f.ncolor = (uint32_t)total_if_last; }
    }
  }
  if (tile == 0) {
    for (uint32_t k = threadIdx.x; k < JPEG_HDR_BYTES; k += blockDim.x) {
      uint8_t v = Tg->header[k];
      if (k == 163) v = (uint8_t)(f.img_h >> 8); else if (k == 164) v = (uint8_t)f.img_h;
      out[k] = v;
    }
    if (U == 0 && threadIdx.x == 0) { out[JPEG_HDR_BYTES] = 0xFF; out[JPEG_HDR_BYTES + 1] = 0xD9; f.J = JPEG_HDR_BYTES + 2; f.ncolor = JPEG_HDR_BYTES + 2; }
  }
}

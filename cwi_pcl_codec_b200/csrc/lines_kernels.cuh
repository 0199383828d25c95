// lines_kernels.cuh -- colour coding type 2 (LINES): ColorCodingJPEG::encodeJPEGLines / decodeJPEGLines
// (cjpeg.h:244-344) with JPEGLineData::serialize (cjpeg.h:66-98).  The averaged colours are cut into V/2048 images of
// 2048x1 pixels (the last one takes the remainder, up to 4095 wide; a cloud below 2048 voxels is one Vx1 image), each a
// stand-alone baseline JPEG; the payload is  u32 line_count, then per line  u32 length + JFIF bytes.
// libjpeg facts that matter for an Nx1 image: one MCU row; the Y plane rows replicate row 0, block
// row 1 (Y2, Y3) lies below the image => dummy blocks (all zero, DC of the previous block in MCU order); the right-most Y1
// is a dummy too when ceil(w/8) is odd; chroma has one real row.  Lines are independent, so -- unlike the snake image --
// the Huffman decode parallelises over lines.
#pragma once
#include "common.cuh"
#include "jpeg_enc_kernels.cuh"

#define LINE_PX 2048
#define LINE_BITS_WORDS 4096            // per-line scratch for the unstuffed entropy bits (16 KiB)
#define LINE_SLOT_BYTES 32768           // per-line staging for the finished JFIF file

__device__ __forceinline__ void lines_geometry(uint32_t V, uint32_t &n_lines, uint32_t &last_w) {
  if (V < LINE_PX) { n_lines = 1; last_w = V; }           // cjpeg.h:255-270
  else { n_lines = V / LINE_PX; last_w = V - LINE_PX * (n_lines - 1); }
}
__device__ __forceinline__ void line_of_mcu(uint32_t mcu, uint32_t n_lines, uint32_t last_w, uint32_t &line, uint32_t &mx, uint32_t &w) {
  line = min(mcu / (LINE_PX / 16), n_lines - 1);
  mx = mcu - line * (LINE_PX / 16);
  w = line == n_lines - 1 ? last_w : LINE_PX;
}

// ---- encode: colour conversion + FDCT + quantisation, one CTA per MCU of any line
__global__ void __launch_bounds__(256) lines_mcu_kernel(EncFrame *frames, const JpegTables *T) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  if (V == 0) return;
  uint32_t n_lines, last_w; lines_geometry(V, n_lines, last_w);
  const uint32_t total = (LINE_PX / 16) * (n_lines - 1) + (last_w + 15) / 16;
  const uint32_t mcu = blockIdx.x;
  if (mcu >= total) return;
  uint32_t line, mx, w; line_of_mcu(mcu, n_lines, last_w, line, mx, w);
  __shared__ int sY[16], sCb[16], sCr[16];
  __shared__ int work[6][64];
  __shared__ short outc[6][64];
  const uint32_t t = threadIdx.x;
  if (t < 16) {                                              // one pixel row; every other row replicates it
    const uint32_t x = min(mx * 16 + t, w - 1);             // right edge replicates the last real column
    const uint8_t *c = f.avg + 3ull * (line * LINE_PX + x);
    const int r = c[0], g = c[1], b = c[2];
    sY[t] = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
    sCb[t] = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
    sCr[t] = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
  }
  __syncthreads();
  {
    const uint32_t px = t & 15, py = t >> 4;
    work[(py >> 3) * 2 + (px >> 3)][(py & 7) * 8 + (px & 7)] = sY[px] - 128;
    if (t < 128) {
      const uint32_t c = t >> 6, k = t & 63, cx = k & 7;
      const int *p = c ? sCr : sCb;
      const int bias = (cx & 1) ? 2 : 1;
      work[4 + c][k] = ((2 * p[2 * cx] + 2 * p[2 * cx + 1] + bias) >> 2) - 128;   // both source rows are the same row
    }
  }
  __syncthreads();
  if (t < 48) { int *p = &work[t >> 3][(t & 7) * 8], v[8]; for (int k = 0; k < 8; k++) v[k] = p[k]; fdct8(v, true); for (int k = 0; k < 8; k++) p[k] = v[k]; }
  __syncthreads();
  if (t < 48) { int *p = &work[t >> 3][t & 7], v[8]; for (int k = 0; k < 8; k++) v[k] = p[8 * k]; fdct8(v, false); for (int k = 0; k < 8; k++) p[8 * k] = v[k]; }
  __syncthreads();
  for (uint32_t e = t; e < 384; e += 256) {
    const uint32_t blk = e >> 6, k = e & 63, nat = T->zz[k];
    int x = work[blk][nat]; const int q = (int)T->q[blk >= 4][nat] << 3;
    const bool neg = x < 0; if (neg) x = -x;
    x = (x + (q >> 1)) / q;
    outc[blk][k] = (short)(neg ? -x : x);
  }
  __syncthreads();
  const bool dummy1 = 2 * mx + 1 >= (w + 7) / 8;             // Y1 right of the last real block column
  if (t < 192) { const uint32_t blk = 1 + (t >> 6), k = t & 63; if (blk >= 2 || dummy1) outc[blk][k] = 0; }
  __syncthreads();
  if (t == 0) { if (dummy1) outc[1][0] = outc[0][0]; outc[2][0] = outc[1][0]; outc[3][0] = outc[2][0]; }
  __syncthreads();
  short *dst = f.coef + (size_t)mcu * 384;
  for (uint32_t e = t; e < 384; e += 256) dst[e] = (&outc[0][0])[e];
}

// ---- encode: Huffman + stuffing + JFIF header of one line, one CTA per line (thread t owns MCU t)
__global__ void __launch_bounds__(256) lines_huff_kernel(EncFrame *frames, const JpegTables *Tg) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  if (V == 0) return;
  uint32_t n_lines, last_w; lines_geometry(V, n_lines, last_w);
  const uint32_t line = blockIdx.x;
  if (line >= n_lines) return;
  if (n_lines > f.lines_cap) { if (line == 0 && threadIdx.x == 0) atomicOr(&f.error, FERR_JPEG_CAP); return; }
  const uint32_t w = line == n_lines - 1 ? last_w : LINE_PX, mcus = (w + 15) / 16, first_mcu = line * (LINE_PX / 16);
  __shared__ JpegTables T;
  __shared__ uint64_t s_scan[33];
  __shared__ uint32_t s_bad;
  for (uint32_t k = threadIdx.x; k < sizeof(JpegTables) / 4; k += blockDim.x) ((uint32_t *)&T)[k] = ((const uint32_t *)Tg)[k];
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const uint32_t t = threadIdx.x;
  uint32_t *bits_buf = f.jbits_buf + (size_t)line * LINE_BITS_WORDS;               // zeroed by zero_region_kernel
  uint8_t *slot = f.line_slots + (size_t)line * LINE_SLOT_BYTES;
  __align__(16) short zz[64];
  uint32_t mybits = 0;
  auto pred_of = [&](uint32_t b) -> int {
    const size_t g = ((size_t)first_mcu + t) * 6 + b;
    if (b >= 4) return t ? f.coef[(g - 6) * 64] : 0;
    if (b > 0) return f.coef[(g - 1) * 64];
    return t ? f.coef[(g - 3) * 64] : 0;                     // Y3 of the previous MCU; 0 at the start of the line
  };
  if (t < mcus) for (uint32_t b = 0; b < 6; b++) {
    const uint4 *src = (const uint4 *)(f.coef + (((size_t)first_mcu + t) * 6 + b) * 64);
    for (int k = 0; k < 8; k++) ((uint4 *)zz)[k] = src[k];
    mybits += huff_block<false>(zz, pred_of(b), &T, b >= 4, nullptr);
  }
  uint64_t tot;
  const uint64_t excl = block_excl_scan_u64(mybits, &tot, s_scan);
  const uint32_t tbits = (uint32_t)tot;
  if (tbits > LINE_BITS_WORDS * 32u - 64u) { if (t == 0) atomicOr(&f.error, FERR_JPEG_CAP); return; }
  if (t < mcus) {
    BitSink sk; sk.buf = bits_buf; sk.cap_words = LINE_BITS_WORDS; sk.acc = 0; sk.nb = (uint32_t)(excl & 31); sk.word = (uint32_t)(excl >> 5); sk.first = true; sk.err = nullptr;
    for (uint32_t b = 0; b < 6; b++) {
      const uint4 *src = (const uint4 *)(f.coef + (((size_t)first_mcu + t) * 6 + b) * 64);
      for (int k = 0; k < 8; k++) ((uint4 *)zz)[k] = src[k];
      huff_block<true>(zz, pred_of(b), &T, b >= 4, &sk);
    }
    sk.finish();
  }
  __threadfence_block();
  __syncthreads();
  // stuffing: 16 bytes per thread per round, running offset across rounds
  const uint32_t U = (tbits + 7) >> 3;
  const uint8_t *ub = (const uint8_t *)bits_buf;
  uint32_t obase = JPEG_HDR_BYTES;
  for (uint32_t c0 = 0; c0 < U; c0 += 256 * 16) {
    const uint32_t b0 = c0 + t * 16;
    __align__(16) uint8_t by[16]; uint32_t nv = 0, cnt = 0;
    if (b0 < U) {
      nv = min(16u, U - b0);
      *(uint4 *)by = __ldcg((const uint4 *)(ub + b0));
      if (b0 + nv == U && (tbits & 7)) by[nv - 1] |= (uint8_t)((1u << (8 - (tbits & 7))) - 1);      // pad with 1-bits
      for (uint32_t k = 0; k < nv; k++) cnt += 1 + (by[k] == 0xFF);
    }
    uint64_t ct;
    const uint64_t ex = block_excl_scan_u64(cnt, &ct, s_scan);
    uint32_t o = obase + (uint32_t)ex;
    if (o + cnt + 2 > LINE_SLOT_BYTES) { if (nv) s_bad = 1; }
    else for (uint32_t k = 0; k < nv; k++) { slot[o++] = by[k]; if (by[k] == 0xFF) slot[o++] = 0; }
    obase += (uint32_t)ct;
  }
  __syncthreads();
  if (s_bad) { if (t == 0) atomicOr(&f.error, FERR_JPEG_CAP); return; }
  for (uint32_t k = t; k < JPEG_HDR_BYTES; k += blockDim.x) {
    uint8_t v = T.header[k];
    if (k == 163) v = 0; else if (k == 164) v = 1; else if (k == 165) v = (uint8_t)(w >> 8); else if (k == 166) v = (uint8_t)w;
    slot[k] = v;
  }
  if (t == 0) { slot[obase] = 0xFF; slot[obase + 1] = 0xD9; f.line_len[line] = obase + 2; }
}

// ---- encode: JPEGLineData::serialize -- offsets (one CTA per frame) and the copy (one CTA per line)
__global__ void __launch_bounds__(1024) lines_offsets_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.x];
  const uint32_t V = f.V;
  if (V == 0 || (f.error & FERR_JPEG_CAP)) return;
  uint32_t n_lines, last_w; lines_geometry(V, n_lines, last_w);
  __shared__ uint64_t s_scan[33];
  if (n_lines > f.lines_cap) return;
  const uint32_t *len = f.line_len;
  uint32_t *off = f.line_off;
  uint32_t base = 4;
  for (uint32_t c0 = 0; c0 < n_lines; c0 += 1024) {
    const uint32_t i = c0 + threadIdx.x;
    const uint32_t v = i < n_lines ? 4 + len[i] : 0;
    uint64_t tot;
    const uint64_t ex = block_excl_scan_u64(v, &tot, s_scan);
    if (i < n_lines) off[i] = base + (uint32_t)ex;
    base += (uint32_t)tot;
  }
  if (threadIdx.x == 0) {
    if (base > f.cpay_cap) { atomicOr(&f.error, FERR_JPEG_CAP); return; }
    for (int k = 0; k < 4; k++) f.cpay[k] = (uint8_t)(n_lines >> (8 * k));
    f.J = base; f.ncolor = base;
  }
}
__global__ void __launch_bounds__(256) lines_copy_kernel(EncFrame *frames) {
  EncFrame &f = frames[blockIdx.y];
  const uint32_t V = f.V;
  if (V == 0 || (f.error & FERR_JPEG_CAP)) return;
  uint32_t n_lines, last_w; lines_geometry(V, n_lines, last_w);
  const uint32_t line = blockIdx.x;
  if (line >= n_lines) return;
  if (n_lines > f.lines_cap) return;
  const uint32_t len = f.line_len[line], off = f.line_off[line];
  const uint8_t *slot = f.line_slots + (size_t)line * LINE_SLOT_BYTES;
  if (threadIdx.x < 4) f.cpay[off + threadIdx.x] = (uint8_t)(len >> (8 * threadIdx.x));
  for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) f.cpay[off + 4 + k] = slot[k];
}

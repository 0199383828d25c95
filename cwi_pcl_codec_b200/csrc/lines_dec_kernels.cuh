// lines_dec_kernels.cuh -- decode side of colour coding type 2 (LINES): ColorCodingJPEG::decodeJPEGLines (cjpeg.h:319-344),
// JPEGLineData::deserialize (cjpeg.h:84-98).  See lines_kernels.cuh for the format.  Lines are independent JPEGs, so the
// Huffman decode runs one warp per line.
#pragma once
#include "dec_kernels.cuh"
#include "lines_kernels.cuh"

#define LINE_MCU_STRIDE (LINE_PX / 16)     // coefficient layout: line i starts at MCU 128*i; only the last line may be longer

// container parse (serial, short): offsets / lengths of the per-line JFIF files.  grid (frames), 32 threads
__global__ void __launch_bounds__(32) lines_index_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.x];
  if (threadIdx.x != 0 || f.error || !f.data_with_color || f.cct != 2) return;
  const uint8_t *c = f.col; const uint32_t n = f.ncol;
  bool bad = n < 4;
  uint32_t cnt = bad ? 0 : ld_u32_unaligned(c), pos = 4;
  if (cnt == 0 || cnt > f.lines_cap) bad = true;
  for (uint32_t i = 0; !bad && i < cnt; i++) {
    if (pos + 4 > n) { bad = true; break; }
    const uint32_t len = ld_u32_unaligned(c + pos);
    if (len > n - pos - 4) { bad = true; break; }
    f.line_off[i] = pos + 4; f.line_len[i] = len; pos += 4 + len;
  }
  if (bad) { atomicOr(&f.error, FERR_BAD_STREAM); f.n_lines = 0; return; }
  f.n_lines = cnt;
}

// one warp per line: marker parse, de-stuffing, Huffman decode -> quantised coefficients.  grid (lines, frames)
__global__ void __launch_bounds__(32) lines_decode_kernel(DecFrame *frames) {
  DecFrame &f = frames[blockIdx.y];
  const uint32_t line = blockIdx.x;
  if (f.error || f.cct != 2 || !f.data_with_color || line >= f.n_lines) return;
  __shared__ HuffDec hd[4];
  __shared__ JpegInfo ji;
  __shared__ uint32_t s_ok;
  const uint32_t lane = lane_id();
  const uint32_t off = f.line_off[line], len = f.line_len[line];
  const uint8_t *in = f.col + off;
  if (lane == 0) {
    bool ok = jpeg_parse(in, len, ji, f.line_qt + (size_t)line * 128);
    // an Nx1 image; every line but the last is 2048 wide in a stream the reference writes (cjpeg.h:277-306)
    if (ok && (ji.h != 1 || ji.w > 2 * LINE_PX - 1 || (line + 1 < f.n_lines && ji.w > LINE_PX))) ok = false;
    s_ok = ok;
    f.line_w[line] = ok ? ji.w : 0;
  }
  __syncwarp();
  if (!s_ok) { if (lane == 0) atomicOr(&f.error, FERR_BAD_STREAM); return; }
  uint8_t *scan = f.scan + (off & ~3u);                      // 4-byte aligned home of the de-stuffed segment
  const uint32_t slen = warp_destuff_span(in + ji.scan, len - ji.scan, scan);
  __syncwarp();
  if (lane == 0) {
    for (int t = 0; t < 4; t++) huff_build(hd[t], in + ji.dht_off[t], in + ji.dht_off[t] + 16, (int)ji.dht_n[t]);
    const uint32_t mcus = (ji.w + 15) / 16;
    if (((size_t)line * LINE_MCU_STRIDE + mcus) * 6 > f.coef_cap_blocks) { atomicOr(&f.error, FERR_JPEG_CAP); return; }
    huff_decode_blocks(scan, slen, mcus * 6, f.coef + (size_t)line * LINE_MCU_STRIDE * 384, hd);
  }
}

// dequantise + ISLOW IDCT of the blocks that carry the one real pixel row (Y0, Y1, Cb, Cr); 8 threads per block
__global__ void __launch_bounds__(256) lines_idct_kernel(DecFrame *frames, const JpegTables *T) {
  DecFrame &f = frames[blockIdx.y];
  if (f.error || f.cct != 2 || !f.data_with_color) return;
  const uint32_t n_lines = f.n_lines;
  if (n_lines == 0) return;
  const uint32_t total_mcus = (n_lines - 1) * LINE_MCU_STRIDE + 2 * LINE_MCU_STRIDE;     // last line may be up to 4095 wide
  if (blockIdx.x * 32 >= total_mcus * 6) return;
  __shared__ int ws[32][64];
  __shared__ uint8_t zz[64];
  if (threadIdx.x < 64) zz[threadIdx.x] = T->zz[threadIdx.x];
  __syncthreads();
  const uint32_t lb = threadIdx.x >> 3, k = threadIdx.x & 7, g = blockIdx.x * 32 + lb;
  const uint32_t mcu = g / 6, blk = g % 6;
  const uint32_t line = min(mcu / LINE_MCU_STRIDE, n_lines - 1), mx = mcu - line * LINE_MCU_STRIDE;
  const uint32_t w = f.line_w[line];
  const bool act = mcu < total_mcus && mx * 16 < w && blk != 2 && blk != 3;
  if (act) {
    const short *c = f.coef + (size_t)g * 64;
    const uint16_t *q = f.line_qt + (size_t)line * 128 + (blk >= 4 ? 64 : 0);
    int v[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = 0;
    for (int z = 0; z < 64; z++) { uint32_t nat = zz[z]; if ((nat & 7) == k) v[nat >> 3] = (int)c[z] * (int)q[z]; }
    idct8(v, o, true);
#pragma unroll
    for (int r = 0; r < 8; r++) ws[lb][r * 8 + k] = o[r];
  }
  __syncthreads();
  if (act && k == 0) {                                       // only pixel row 0 exists
    int v[8], o[8];
#pragma unroll
    for (int cI = 0; cI < 8; cI++) v[cI] = ws[lb][cI];
    idct8(v, o, false);
    uint8_t *P = f.line_planes + (size_t)line * 8192;
    uint8_t *dst = blk < 2 ? P + mx * 16 + blk * 8 : (blk == 4 ? P + 4096 + mx * 8 : P + 6144 + mx * 8);
#pragma unroll
    for (int cI = 0; cI < 8; cI++) dst[cI] = jpeg_range_limit(o[cI]);
  }
}

// colour of voxel i in LINES mode: libjpeg's upsampling for a one-row image (fancy when the chroma row is wider than 2)
__device__ __forceinline__ uint32_t dec_color_lines(const DecFrame &f, uint32_t i) {
  const uint32_t n_lines = f.n_lines;
  if (n_lines == 0) return 0;
  const uint32_t line = min(i / LINE_PX, n_lines - 1), x = i - line * LINE_PX, w = f.line_w[line];
  if (x >= w) return 0;
  const uint8_t *P = f.line_planes + (size_t)line * 8192;
  const uint32_t cw = (w + 1) >> 1, cx = x >> 1;
  int cc[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const uint8_t *C = P + 4096 + 2048 * c;
    int o;
    if (cw <= 2) o = C[cx];
    else {
      const int cs = 4 * C[cx];                               // the row above / below replicate the only row
      if (!(x & 1)) o = cx == 0 ? (4 * cs + 8) >> 4 : (3 * cs + 4 * C[cx - 1] + 8) >> 4;
      else o = cx == cw - 1 ? (4 * cs + 7) >> 4 : (3 * cs + 4 * C[cx + 1] + 7) >> 4;
    }
    cc[c] = o - 128;
  }
  const int yy = P[x];
  int R = yy + ((91881 * cc[1] + 32768) >> 16), B = yy + ((116130 * cc[0] + 32768) >> 16), G = yy + ((-22554 * cc[0] - 46802 * cc[1] + 32768) >> 16);
  R = min(255, max(0, R)); G = min(255, max(0, G)); B = min(255, max(0, B));
  return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
}

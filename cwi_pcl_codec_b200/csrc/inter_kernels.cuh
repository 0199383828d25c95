// inter_kernels.cuh -- inter-frame (predictive) coding on the GPU.
// Reference behaviour restated (oracle/ccv2_oracle_inter.c is the CPU statement of the same steps; rtc / qc =
// cloud_codec_v2/include/pcl/cloud_codec_v2/impl/{rigid_transform_coding,quaternion_coding}_impl.hpp):
//   simplifyPCloud impl.hpp:318-400, generate_macroblock_tree :410-431, do_icp_prediction :443-568,
//   encodePointCloudDeltaFrame :787-1112, decodePointCloudDeltaFrame :1120-1235, RigidTransformCoding rtc:63-203,
//   QuaternionCoding qc:55-222.
// The three octrees the reference builds per delta frame (the P frame's voxel grid, the two macroblock trees) are the
// encoder's own front end run on a box the host defined (bbox -> keys -> radix sort -> leaf scan, enc_kernels.cuh):
// a leaf = a run of equal codes, its point list = the run's index values (stable sort: input order).  What is new here:
//   simplify_kernel      one thread per voxel: centre / centroid + integer colour mean           (HBM bound, 32 B out per voxel)
//   mb_match_kernel      one thread per P macroblock: findLeaf in the I tree (binary search over sorted codes)
//   mb_icp_kernel        one single-warp CTA per shared macroblock: the gates, then point-to-point ICP -- neighbour search by all
//                        lanes over shared-memory tiles of the target block, the closed-form alignment by ONE thread in
//                        double in the oracle's summation order (what makes the P stream bit-exact against it), transform
//                        quantisation (quaternion or two-row mode)                                 (FP32 issue / latency bound)
//   mb_scan_kernel       offsets of every macroblock's chunk, unpredicted points and predicted points (one CTA, chained)
//   mb_write_kernel      P-stream chunks, the cloud of unpredicted points, optionally the predicted frame
//   pchunk_* kernels     decoder: chunk walk (one thread: the chunk sizes chain), per-chunk findLeaf + transform
//                        decompression, offsets, transformPointCloud of the I block
// Arithmetic: the library is compiled with --fmad=false, so every expression below rounds exactly like the same
// expression in the oracle (gcc -ffp-contract=off); double division and sqrt are IEEE on both sides.
#pragma once
#include "common.cuh"
#include <math_constants.h>

struct MbResult {                // one per P macroblock
  int32_t match;                 // leaf index in the I macroblock tree, -1: exclusive block
  uint32_t ok;                   // 1: predicted (chunk written), 0: its points are coded intra
  int16_t words[10]; uint32_t nw;
  int8_t off[3]; uint8_t _p;
  uint32_t iters; float fitness;
};
struct InterCtx {                // by value to the kernels
  const EncFrame *gp, *gi;       // macroblock grids of the P cloud and of the I cloud (device records)
  const uint8_t *P, *I;          // the clouds the grids index (32-byte records)
  MbResult *res;
  float *cur, *tgt; float *d2; uint32_t *nn;     // scratch: per I point (sorted position) / per P point (sorted position)
  uint32_t *p_off, *x_off, *o_off;               // per macroblock (+1): chunk byte offset, unpredicted-point offset, out-cloud offset
  uint8_t *p_stream; uint8_t *intra_pts; uint8_t *out_pts;   // out_pts may be null
  uint32_t *ticket;              // [0] ICP work ticket, [1..] totals: p bytes, unpredicted points, out points, shared, converged, macroblocks
  int color_offset, max_iter;
  double point_res, tf_eps, fit_eps;
};

// ------------------------------------------------------------------------------------------------ grid plumbing
// npasses / empty-frame rule for a grid record (frame_setup_kernel without the frame ids); `chain`: this grid's input is the
// simplified cloud another grid produced, so its point count is only known on the device.
__global__ void grid_setup_kernel(EncFrame *f, const EncFrame *chain_from) {
  if (threadIdx.x || blockIdx.x) return;
  if ((f->error & FERR_DEPTH) || !f->defined) f->n_finite = 0;
  if (f->n_finite > 0) f->npasses = (frame_sort_bits(*f) + 7) / 8; else { f->npasses = 0; f->V = 0; f->B = 0; }
  (void)chain_from;
}
__global__ void grid_chain_kernel(EncFrame *f, const EncFrame *from) {
  if (threadIdx.x || blockIdx.x) return;
  f->n = from->V; f->n_finite = from->V;
}

// simplifyPCloud (impl.hpp:343-397): one thread per voxel of grid g; out = V x 32-byte records in DFS order
__global__ void __launch_bounds__(256) simplify_kernel(const EncFrame *g, EncParams P, uint8_t *out) {
  const EncFrame &f = *g;
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f.V) return;
  const uint64_t key = f.leaf_key[j];
  const uint32_t k3[3] = { compact3(key >> 2), compact3(key >> 1), compact3(key) };
  const uint32_t s0 = f.leaf_start[j], s1 = f.leaf_start[j + 1];
  const uint32_t *vals = f.vals[f.npasses & 1];
  float xyz[3];
  if (P.do_centroid) {                                   // pcl::compute3DCentroid: float sums in index order, one division
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (uint32_t k = s0; k < s1; k++) { const float4 q = __ldg((const float4 *)(f.pts + 32ull * vals[k])); ax = ax + q.x; ay = ay + q.y; az = az + q.z; }
    const float cnt = (float)(s1 - s0);
    xyz[0] = ax / cnt; xyz[1] = ay / cnt; xyz[2] = az / cnt;
  } else {
#pragma unroll
    for (int a = 0; a < 3; a++) xyz[a] = (float)(((double)k3[a] + 0.5) * P.res + f.bmin[a]);   // [PCL] genLeafNodeCenterFromOctreeKey
  }
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  if (f.packed) { const uint64_t *sk = f.keys[f.npasses & 1]; for (uint32_t k = s0; k < s1; k++) { const uint32_t c = (uint32_t)__ldg(&sk[k]); c0 += c & 0xFF; c1 += (c >> 8) & 0xFF; c2 += (c >> 16) & 0xFF; } }
  else for (uint32_t k = s0; k < s1; k++) { const uint32_t c = __ldg((const uint32_t *)(f.pts + 32ull * vals[k] + 16)); c0 += c & 0xFF; c1 += (c >> 8) & 0xFF; c2 += (c >> 16) & 0xFF; }
  const uint32_t len = s1 - s0;
  c0 /= len; c1 /= len; c2 /= len;                         // impl.hpp:383-397: (char)(sum / size)
  uint4 *o = (uint4 *)(out + 32ull * j);
  o[0] = make_uint4(__float_as_uint(xyz[0]), __float_as_uint(xyz[1]), __float_as_uint(xyz[2]), 0x3F800000u);
  o[1] = make_uint4(0xFF000000u | c0 | (c1 << 8) | (c2 << 16), 0, 0, 0);
}

// [PCL] findLeaf(x, y, z) of the I macroblock tree for a key of the P tree: the integer key coordinates are taken as they
// are (both trees start from the same unit box); a key outside the I tree's range finds nothing.
__device__ __forceinline__ int32_t grid_find(const EncFrame &g, uint32_t kx, uint32_t ky, uint32_t kz) {
  const uint32_t d = g.depth;
  if (d < 32 && ((kx >> d) | (ky >> d) | (kz >> d))) return -1;
  const uint64_t code = morton_xyz(kx, ky, kz);
  uint32_t lo = 0, hi = g.V;
  while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (g.leaf_key[m] < code) lo = m + 1; else hi = m; }
  return (lo < g.V && g.leaf_key[lo] == code) ? (int32_t)lo : -1;
}
__global__ void __launch_bounds__(256) mb_match_kernel(InterCtx X) {
  const uint32_t L = blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= X.gp->V) return;
  const uint64_t key = X.gp->leaf_key[L];
  MbResult &r = X.res[L];
  r.match = grid_find(*X.gi, compact3(key >> 2), compact3(key >> 1), compact3(key));
  r.ok = 0; r.nw = 0; r.off[0] = r.off[1] = r.off[2] = 0; r.iters = 0; r.fitness = 0.f;
}

// ------------------------------------------------------------------------------------------------ transform coders
// QuaternionCoding (qc:55-222); q = (w, x, y, z)
__device__ __forceinline__ float qclampf(float v) { if (v < -1) v = -1; else if (v > 1) v = 1; return v; }
__device__ inline void quat_compress_d(const float q[4], int16_t s[3]) {
  const float scale = 1.41421f;
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  if (w > x && w > y && w > z) {
    float rx = x * scale, ry = y * scale, rz = z * scale;
    if (w < 0) { rx = -rx; ry = -ry; rz = -rz; }
    rx = qclampf(rx); ry = qclampf(ry); rz = qclampf(rz);
    s[0] = (int16_t)(int)(rx * 32767); s[1] = (int16_t)(((int)(ry * 32767) & 0xfffe) | 1); s[2] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 1);
  } else if (z > x && z > y) {
    float rx = x * scale, ry = y * scale, rw = w * scale;
    if (z < 0) { rx = -rx; ry = -ry; rw = -rw; }
    rx = qclampf(rx); ry = qclampf(ry); rw = qclampf(rw);
    s[0] = (int16_t)(int)(rx * 32767); s[1] = (int16_t)(((int)(ry * 32767) & 0xfffe) | 1); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 0);
  } else if (y > x) {
    float rx = x * scale, rz = z * scale, rw = w * scale;
    if (y < 0) { rx = -rx; rz = -rz; rw = -rw; }
    rx = qclampf(rx); rz = qclampf(rz); rw = qclampf(rw);
    s[0] = (int16_t)(int)(rx * 32767); s[1] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 0); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 1);
  } else {
    float ry = y * scale, rz = z * scale, rw = w * scale;
    if (x < 0) { ry = -ry; rz = -rz; rw = -rw; }
    ry = qclampf(ry); rz = qclampf(rz); rw = qclampf(rw);
    s[0] = (int16_t)(int)(ry * 32767); s[1] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 0); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 0);
  }
}
__device__ inline void quat_decompress_d(const int16_t sin[3], float q[4]) {
  int16_t s0 = sin[0], s1 = sin[1], s2 = sin[2];
  const int which = ((s1 & 1) << 1) | (s2 & 1);
  s1 = (int16_t)(s1 & ~1); s2 = (int16_t)(s2 & ~1);
  const float scale = 1.0f / 32767.0f / 1.41421f;
  const float FE = 1.1920928955078125e-07f;
  float w, x, y, z;
  if (which == 3) { x = s0 * scale; y = s1 * scale; z = s2 * scale; w = 1 - (x * x) - (y * y) - (z * z); if (w > FE) w = sqrtf(w); }
  else if (which == 2) { x = s0 * scale; y = s1 * scale; w = s2 * scale; z = 1 - (x * x) - (y * y) - (w * w); if (z > FE) z = sqrtf(z); }
  else if (which == 1) { x = s0 * scale; z = s1 * scale; w = s2 * scale; y = 1 - (x * x) - (z * z) - (w * w); if (y > FE) y = sqrtf(y); }
  else { y = s0 * scale; z = s1 * scale; w = s2 * scale; x = 1 - (y * y) - (z * z) - (w * w); if (x > FE) x = sqrtf(x); }
  q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
// Eigen::Quaternion<float>(Matrix3f) / toRotationMatrix(); m row-major 3x3
__device__ inline void mat_to_quat_d(const float m[9], float q[4]) {
  float t = m[0] + m[4] + m[8];
  if (t > 0.0f) {
    t = sqrtf(t + 1.0f); q[0] = 0.5f * t; t = 0.5f / t;
    q[1] = (m[7] - m[5]) * t; q[2] = (m[2] - m[6]) * t; q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(m[4 * i] - m[4 * j] - m[4 * k] + 1.0f);
    q[1 + i] = 0.5f * t; t = 0.5f / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t; q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t; q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
__device__ inline void quat_to_mat_d(const float q[4], float r[9]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  r[0] = 1 - (tyy + tzz); r[1] = txy - twz; r[2] = txz + twy;
  r[3] = txy + twz; r[4] = 1 - (txx + tzz); r[5] = tyz - twx;
  r[6] = txz - twy; r[7] = tyz + twx; r[8] = 1 - (txx + tyy);
}
// RigidTransformCoding::compressRigidTransform (rtc:63-148); m row-major 4x4; returns the word count (6 or 10)
__device__ inline int rigid_compress_d(const float *m, int16_t *out) {
  const float scaling_factor = (float)((float)32767 / 2.5);
  float rot[9]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rot[3 * r + c] = m[4 * r + c];
  float q[4], qt[4], res[9]; int16_t s[3];
  mat_to_quat_d(rot, q);
  quat_compress_d(q, s); quat_decompress_d(s, qt); quat_to_mat_d(qt, res);
  bool stable = true;
  for (int i = 0; i < 9 && stable; i++) if ((double)fabsf(res[i] - rot[i]) > 0.001) stable = false;
  int n = 0;
  if (!stable) {
    int16_t w6 = 0;
    for (int l = 0; l < 3; l++) {
      out[l] = (int16_t)(int)(rot[l] * (float)(32767 - 1)); out[l + 3] = (int16_t)(int)(rot[3 + l] * (float)(32767 - 1));
      w6 = (int16_t)(w6 + (rot[6 + l] < 0 ? 1 << l : 0));
    }
    out[6] = w6; n = 7;
  } else { quat_compress_d(q, out); n = 3; }
  for (int a = 0; a < 3; a++) {
    float t = m[4 * a + 3];
    if ((double)t > 2.5) t = 2.5f;
    if ((double)t < -2.5) t = -2.5f;
    out[n++] = (int16_t)(int)(t * (scaling_factor - 1));
  }
  return n;
}
// RigidTransformCoding::deCompressRigidTransform (rtc:158-203)
__device__ inline void rigid_decompress_d(const int16_t *in, int nwords, float *m) {
  const float scaling_factor = (float)((float)32767 / 2.5);
  float r[9];
  if (nwords == 6) { float q[4]; quat_decompress_d(in, q); quat_to_mat_d(q, r); }
  else {
    for (int l = 0; l < 3; l++) {
      r[l] = ((float)in[l]) / (float)(32767 - 1); r[3 + l] = ((float)in[l + 3]) / (float)(32767 - 1);
      r[6 + l] = sqrtf(1 - r[l] * r[l] - r[3 + l] * r[3 + l]);
      if (((1 << l) & ((int)in[6])) == 1 << l) r[6 + l] = -r[6 + l];
    }
  }
  for (int a = 0; a < 3; a++) { for (int c = 0; c < 3; c++) m[4 * a + c] = r[3 * a + c]; m[4 * a + 3] = ((float)in[nwords - 3 + a]) / ((float)(scaling_factor - 1)); }
  m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
}
// pcl::transformPointCloud, SSE path: x*c0 + (y*c1 + (z*c2 + c3)); o[3] is the record's data[3]
__device__ __forceinline__ void xform_pcl_d(const float *m, float x, float y, float z, float o[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) o[r] = x * m[4 * r] + (y * m[4 * r + 1] + (z * m[4 * r + 2] + m[4 * r + 3]));
}

// ------------------------------------------------------------------------------------------------ registration
// largest eigenvector of a symmetric 4x4 matrix: cyclic Jacobi, at most 24 sweeps (oracle: jacobi4_max)
__device__ inline void jacobi4_max_d(double A[4][4], double v[4]) {
  double E[4][4] = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
  for (int sweep = 0; sweep < 24; sweep++) {
    double off = 0, dn = 0;
    for (int p = 0; p < 4; p++) { dn += A[p][p] * A[p][p]; for (int q = p + 1; q < 4; q++) off += A[p][q] * A[p][q]; }
    if (off <= 1e-32 * dn || off == 0.0) break;
    for (int p = 0; p < 3; p++) for (int q = p + 1; q < 4; q++) {
      const double apq = A[p][q];
      if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 4; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
      for (int k = 0; k < 4; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      for (int k = 0; k < 4; k++) { const double ekp = E[k][p], ekq = E[k][q]; E[k][p] = c * ekp - s * ekq; E[k][q] = s * ekp + c * ekq; }
    }
  }
  int best = 0;
  for (int k = 1; k < 4; k++) if (A[k][k] > A[best][best]) best = k;
  for (int k = 0; k < 4; k++) v[k] = E[k][best];
}
// closed-form rigid alignment of s[i] onto t[nn[i]], ONE thread, index order (oracle: estimate_rigid)
__device__ inline void estimate_rigid_d(const float *s, const float *t, const uint32_t *nn, uint32_t n, float *T) {
  double ms[3] = { 0, 0, 0 }, mt[3] = { 0, 0, 0 };
  for (uint32_t i = 0; i < n; i++) { const float *a = s + 3 * i, *b = t + 3 * (size_t)nn[i]; for (int k = 0; k < 3; k++) { ms[k] += (double)a[k]; mt[k] += (double)b[k]; } }
  for (int k = 0; k < 3; k++) { ms[k] /= (double)n; mt[k] /= (double)n; }
  double H[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
  for (uint32_t i = 0; i < n; i++) {
    const float *a = s + 3 * i, *b = t + 3 * (size_t)nn[i];
    double ds[3], dt[3];
    for (int k = 0; k < 3; k++) { ds[k] = (double)a[k] - ms[k]; dt[k] = (double)b[k] - mt[k]; }
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) H[k][l] += ds[k] * dt[l];
  }
  double N[4][4];
  N[0][0] = (H[0][0] + H[1][1]) + H[2][2]; N[0][1] = H[1][2] - H[2][1]; N[0][2] = H[2][0] - H[0][2]; N[0][3] = H[0][1] - H[1][0];
  N[1][1] = (H[0][0] - H[1][1]) - H[2][2]; N[1][2] = H[0][1] + H[1][0]; N[1][3] = H[2][0] + H[0][2];
  N[2][2] = (H[1][1] - H[0][0]) - H[2][2]; N[2][3] = H[1][2] + H[2][1];
  N[3][3] = (H[2][2] - H[0][0]) - H[1][1];
  for (int p = 0; p < 4; p++) for (int q = 0; q < p; q++) N[p][q] = N[q][p];
  double q[4]; jacobi4_max_d(N, q);
  double nrm = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
  if (!(nrm > 0)) { q[0] = 1; q[1] = q[2] = q[3] = 0; nrm = 1; }
  if (q[0] < 0) nrm = -nrm;
  const double w = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, z = q[3] / nrm;
  double R[3][3];
  R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z); R[0][2] = 2 * (x * z + w * y);
  R[1][0] = 2 * (x * y + w * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
  R[2][0] = 2 * (x * z - w * y); R[2][1] = 2 * (y * z + w * x); R[2][2] = 1 - 2 * (x * x + y * y);
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) T[4 * a + b] = (float)R[a][b];
    T[4 * a + 3] = (float)(mt[a] - ((R[a][0] * ms[0] + R[a][1] * ms[1]) + R[a][2] * ms[2]));
  }
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// One WARP per macroblock: a macroblock holds a patch of a surface -- a few dozen voxels (22 on average for the 1M-point frames at
// 11 bits, at most 16^3) -- so 32 lanes cover the neighbour search and 32 single-warp CTAs per SM keep 4736 registrations in flight.
#define ICP_THREADS 32
#define ICP_TILE 512
// nearest neighbour of every cur[i] among tgt[0..nt): float ((dx*dx + dy*dy) + dz*dz), first minimum.  All threads of the CTA.
__device__ inline void icp_nn(const float *cur, uint32_t ns, const float *tgt, uint32_t nt, uint32_t *nn, float *d2, float *s_tile, bool tile_resident) {
  for (uint32_t i0 = 0; i0 < ns; i0 += ICP_THREADS) {
    const uint32_t i = i0 + threadIdx.x;
    float cx = 0, cy = 0, cz = 0;
    if (i < ns) { cx = cur[3 * i]; cy = cur[3 * i + 1]; cz = cur[3 * i + 2]; }
    float best = CUDART_INF_F; uint32_t bj = 0;
    for (uint32_t t0 = 0; t0 < nt; t0 += ICP_TILE) {
      const uint32_t tn = min((uint32_t)ICP_TILE, nt - t0);
      if (!tile_resident) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < 3 * tn; k += ICP_THREADS) s_tile[k] = tgt[3 * (size_t)t0 + k];
        __syncthreads();
      }
      if (i < ns) {
        for (uint32_t j = 0; j < tn; j++) {
          const float dx = cx - s_tile[3 * j], dy = cy - s_tile[3 * j + 1], dz = cz - s_tile[3 * j + 2];
          const float d = (dx * dx + dy * dy) + dz * dz;
          if (d < best) { best = d; bj = t0 + j; }
        }
      }
    }
    if (i < ns) { nn[i] = bj; d2[i] = best; }
  }
  __syncthreads();
}

// do_icp_prediction (impl.hpp:443-568) + compressRigidTransform for every shared macroblock; persistent CTAs take
// macroblocks from a ticket.
__global__ void __launch_bounds__(ICP_THREADS) mb_icp_kernel(InterCtx X) {   // 128 registers: 16 warps per SM; capping at 64 (32 warps) spills the Jacobi arrays and is slower (4.2 against 3.5 ms per 1M-point frame)
  __shared__ float s_tile[3 * ICP_TILE];
  __shared__ float s_T[16], s_F[16];
  __shared__ uint32_t s_L; __shared__ int s_go, s_conv;
  const EncFrame &gp = *X.gp, &gi = *X.gi;
  const uint32_t nmb = gp.V;
  const uint32_t *pv = gp.vals[gp.npasses & 1], *iv = gi.vals[gi.npasses & 1];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_L = atomicAdd(&X.ticket[0], 1u);
    __syncthreads();
    const uint32_t L = s_L;
    if (L >= nmb) return;
    MbResult &r = X.res[L];
    if (r.match < 0) continue;
    const uint32_t ps0 = gp.leaf_start[L], np = gp.leaf_start[L + 1] - ps0;
    const uint32_t is0 = gi.leaf_start[r.match], ni = gi.leaf_start[r.match + 1] - is0;
    if (threadIdx.x == 0) {
      atomicAdd(&X.ticket[4], 1u);                         // shared_macroblock_count
      bool do_icp = np > 6 ? ((np < ni * 2) && ((double)np >= (double)ni * 0.5)) : false;
      if ((uint64_t)np * (uint64_t)ni > (1ull << 24)) do_icp = false;   // not in the reference: see oracle/ccv2_oracle_inter.c icp_prediction (macroblocks far beyond 16 voxels)
      if (do_icp) {                                        // colour variance gate, offsets (impl.hpp:463-535); index 0,1,2 = r,g,b = record bytes 18,17,16
        double in_av[3] = { 0, 0, 0 }, out_av[3] = { 0, 0, 0 }, in_var = 0, out_var = 0;
        for (uint32_t i = 0; i < ni; i++) { const uint32_t c = *(const uint32_t *)(X.I + 32ull * iv[is0 + i] + 16); in_av[0] += (double)((c >> 16) & 0xFF); in_av[1] += (double)((c >> 8) & 0xFF); in_av[2] += (double)(c & 0xFF); }
        for (int a = 0; a < 3; a++) in_av[a] /= (double)ni;
        for (uint32_t i = 0; i < ni; i++) {
          const uint32_t c = *(const uint32_t *)(X.I + 32ull * iv[is0 + i] + 16);
          const double rr = (double)((c >> 16) & 0xFF) - in_av[0], gg = (double)((c >> 8) & 0xFF) - in_av[1], bb = (double)(c & 0xFF) - in_av[2];
          in_var += rr * rr + gg * gg + bb * bb;
        }
        in_var /= (double)(3 * (size_t)ni);
        for (uint32_t i = 0; i < np; i++) { const uint32_t c = *(const uint32_t *)(X.P + 32ull * pv[ps0 + i] + 16); out_av[0] += (double)((c >> 16) & 0xFF); out_av[1] += (double)((c >> 8) & 0xFF); out_av[2] += (double)(c & 0xFF); }
        for (int a = 0; a < 3; a++) out_av[a] /= (double)np;
        for (uint32_t i = 0; i < np; i++) {
          const uint32_t c = *(const uint32_t *)(X.P + 32ull * pv[ps0 + i] + 16);
          const double rr = (double)((c >> 16) & 0xFF) - out_av[0], gg = (double)((c >> 8) & 0xFF) - out_av[1], bb = (double)(c & 0xFF) - out_av[2];
          out_var += rr * rr + gg * gg + bb * bb;
        }
        out_var /= (double)(3 * (size_t)np);
        if (in_var > 100.0 || out_var > 100.0) do_icp = false;
        if (X.color_offset) for (int a = 0; a < 3; a++) if (fabs(out_av[a] - in_av[a]) < 32) r.off[a] = (int8_t)(int)(out_av[a] - in_av[a]);
      }
      s_go = do_icp;
      for (int k = 0; k < 16; k++) s_F[k] = (k % 5 == 0) ? 1.0f : 0.0f;
      s_conv = 0;
    }
    __syncthreads();
    if (!s_go) continue;
    float *cur = X.cur + 3ull * is0, *tgt = X.tgt + 3ull * ps0, *d2 = X.d2 + is0; uint32_t *nn = X.nn + is0;
    for (uint32_t i = threadIdx.x; i < ni; i += ICP_THREADS) { const float4 q = *(const float4 *)(X.I + 32ull * iv[is0 + i]); cur[3 * i] = q.x; cur[3 * i + 1] = q.y; cur[3 * i + 2] = q.z; }
    for (uint32_t i = threadIdx.x; i < np; i += ICP_THREADS) { const float4 q = *(const float4 *)(X.P + 32ull * pv[ps0 + i]); tgt[3 * i] = q.x; tgt[3 * i + 1] = q.y; tgt[3 * i + 2] = q.z; }
    __syncthreads();
    const bool resident = np <= ICP_TILE;                  // the target block stays in shared memory over all iterations
    if (resident) { for (uint32_t k = threadIdx.x; k < 3 * np; k += ICP_THREADS) s_tile[k] = tgt[k]; __syncthreads(); }
    double prev_mse = 1.7976931348623157e308;
    uint32_t iters = 0;
    // ns >= 3 holds: np > 6 and np < 2 ni
    for (;;) {
      icp_nn(cur, ni, tgt, np, nn, d2, s_tile, resident);
      if (threadIdx.x == 0) estimate_rigid_d(cur, tgt, nn, ni, s_T);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < ni; i += ICP_THREADS) {   // IterativeClosestPoint::transformCloud: ((m0*x + m1*y) + m2*z) + m3
        const float x = cur[3 * i], y = cur[3 * i + 1], z = cur[3 * i + 2];
#pragma unroll
        for (int a = 0; a < 3; a++) cur[3 * i + a] = ((s_T[4 * a] * x + s_T[4 * a + 1] * y) + s_T[4 * a + 2] * z) + s_T[4 * a + 3];
      }
      iters++;
      if (threadIdx.x == 0) {
        float G[16];
        for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) G[4 * a + c] = ((s_T[4 * a] * s_F[c] + s_T[4 * a + 1] * s_F[4 + c]) + s_T[4 * a + 2] * s_F[8 + c]) + s_T[4 * a + 3] * s_F[12 + c];
        for (int k = 0; k < 16; k++) s_F[k] = G[k];
        // DefaultConvergenceCriteria::hasConverged (PCL 1.10)
        int stop = 0;
        if ((int)iters >= X.max_iter) stop = 1;
        else {
          const double cos_angle = 0.5 * (double)(s_T[0] + s_T[5] + s_T[10] - 1);
          const double tr2 = (double)(s_T[3] * s_T[3] + s_T[7] * s_T[7] + s_T[11] * s_T[11]);
          if (cos_angle >= 1.0 - X.tf_eps && tr2 <= X.tf_eps) stop = 1;
          else {
            double mse = 0; for (uint32_t i = 0; i < ni; i++) mse += (double)d2[i];
            mse /= (double)ni;
            if (fabs(mse - prev_mse) < 1e-12) stop = 1;
            else if (fabs(mse - prev_mse) / prev_mse < X.fit_eps) stop = 1;
            prev_mse = mse;
          }
        }
        s_conv = stop;
      }
      __syncthreads();
      if (s_conv) break;
    }
    // getFitnessScore: the ORIGINAL source block under the final transform
    for (uint32_t i = threadIdx.x; i < ni; i += ICP_THREADS) {
      const float4 q = *(const float4 *)(X.I + 32ull * iv[is0 + i]);
      float o[4]; xform_pcl_d(s_F, q.x, q.y, q.z, o);
      cur[3 * i] = o[0]; cur[3 * i + 1] = o[1]; cur[3 * i + 2] = o[2];
    }
    __syncthreads();
    icp_nn(cur, ni, tgt, np, nn, d2, s_tile, resident);
    if (threadIdx.x == 0) {
      double fs = 0; for (uint32_t i = 0; i < ni; i++) fs += (double)d2[i];
      fs /= (double)ni;
      r.iters = iters; r.fitness = (float)fs;
      if (fs < X.point_res * 2) {
        r.ok = 1; r.nw = (uint32_t)rigid_compress_d(s_F, r.words);
        atomicAdd(&X.ticket[5], 1u);                       // convergence_count
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ output layout
// Exclusive offsets over the P macroblocks in DFS order: chunk bytes, unpredicted points, predicted-frame points.
// One CTA; chunks of 1024 macroblocks with a carried base.
__global__ void __launch_bounds__(1024) mb_scan_kernel(InterCtx X, int want_out) {
  __shared__ uint64_t s_scan[33];
  const EncFrame &gp = *X.gp, &gi = *X.gi;
  const uint32_t nmb = gp.V;
  uint64_t base_p = 0, base_x = 0, base_o = 0;
  for (uint32_t L0 = 0; L0 < nmb; L0 += 1024) {
    const uint32_t L = L0 + threadIdx.x;
    uint64_t vp = 0, vx = 0, vo = 0;
    if (L < nmb) {
      const MbResult &r = X.res[L];
      const uint32_t np = gp.leaf_start[L + 1] - gp.leaf_start[L];
      if (r.ok) { vp = 1 + 6 + 2 * r.nw + (X.color_offset ? 3 : 0); vo = gi.leaf_start[r.match + 1] - gi.leaf_start[r.match]; }
      else { vx = np; vo = np; }
    }
    uint64_t tp, tx, to;
    const uint64_t ep = block_excl_scan_u64(vp, &tp, s_scan), ex = block_excl_scan_u64(vx, &tx, s_scan), eo = block_excl_scan_u64(vo, &to, s_scan);
    if (L < nmb) { X.p_off[L] = (uint32_t)(base_p + ep); X.x_off[L] = (uint32_t)(base_x + ex); X.o_off[L] = (uint32_t)(base_o + eo); }
    base_p += tp; base_x += tx; base_o += to;
  }
  if (threadIdx.x == 0) {
    X.p_off[nmb] = (uint32_t)base_p; X.x_off[nmb] = (uint32_t)base_x; X.o_off[nmb] = (uint32_t)base_o;
    X.ticket[1] = (uint32_t)base_p; X.ticket[2] = (uint32_t)base_x; X.ticket[3] = want_out ? (uint32_t)base_o : 0u; X.ticket[6] = nmb;
  }
}

// One CTA per P macroblock: its chunk [u8 size][3 x i16 key][nw x i16][3 x i8] (impl.hpp:877-883), or its points appended
// to the cloud that is coded intra (impl.hpp:916-937); with write_out_cloud also the predicted frame (impl.hpp:893-913).
__global__ void __launch_bounds__(128) mb_write_kernel(InterCtx X) {
  const EncFrame &gp = *X.gp, &gi = *X.gi;
  const uint32_t *pv = gp.vals[gp.npasses & 1], *iv = gi.vals[gi.npasses & 1];
  __shared__ float s_m[16];
  for (uint32_t L = blockIdx.x; L < gp.V; L += gridDim.x) {
  const MbResult &r = X.res[L];
  const uint32_t ps0 = gp.leaf_start[L], np = gp.leaf_start[L + 1] - ps0;
  __syncthreads();
  if (r.ok) {
    if (threadIdx.x == 0) {
      uint8_t *o = X.p_stream + X.p_off[L];
      const uint64_t key = gp.leaf_key[L];
      const int16_t k3[3] = { (int16_t)(int)compact3(key >> 2), (int16_t)(int)compact3(key >> 1), (int16_t)(int)compact3(key) };
      *o++ = (uint8_t)(6 + 2 * r.nw + (X.color_offset ? 3 : 0));
      for (int a = 0; a < 3; a++) { *o++ = (uint8_t)(k3[a] & 0xFF); *o++ = (uint8_t)((uint16_t)k3[a] >> 8); }
      for (uint32_t k = 0; k < r.nw; k++) { *o++ = (uint8_t)(r.words[k] & 0xFF); *o++ = (uint8_t)((uint16_t)r.words[k] >> 8); }
      if (X.color_offset) for (int a = 0; a < 3; a++) *o++ = (uint8_t)r.off[a];
    }
    if (X.out_pts) {
      if (threadIdx.x == 0) rigid_decompress_d(r.words, (int)r.nw, s_m);
      __syncthreads();
      const uint32_t is0 = gi.leaf_start[r.match], ni = gi.leaf_start[r.match + 1] - is0;
      for (uint32_t i = threadIdx.x; i < ni; i += blockDim.x) {
        const uint4 *src = (const uint4 *)(X.I + 32ull * iv[is0 + i]);
        uint4 a = src[0], b = src[1];
        float o[4]; xform_pcl_d(s_m, __uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), o);
        a = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
        if (X.color_offset) {                              // pt.r += offsets[0] ... (uint8 arithmetic)
          const uint32_t rr = ((b.x >> 16) + (uint32_t)(int)r.off[0]) & 0xFF, gg = ((b.x >> 8) + (uint32_t)(int)r.off[1]) & 0xFF, bl = (b.x + (uint32_t)(int)r.off[2]) & 0xFF;
          b.x = (b.x & 0xFF000000u) | (rr << 16) | (gg << 8) | bl;
        }
        uint4 *dst = (uint4 *)(X.out_pts + 32ull * (X.o_off[L] + i));
        dst[0] = a; dst[1] = b;
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
      const uint4 *src = (const uint4 *)(X.P + 32ull * pv[ps0 + i]);
      const uint4 a = src[0], b = src[1];
      uint4 *dst = (uint4 *)(X.intra_pts + 32ull * (X.x_off[L] + i));
      dst[0] = a; dst[1] = b;
      if (X.out_pts) { uint4 *d2 = (uint4 *)(X.out_pts + 32ull * (X.o_off[L] + i)); d2[0] = a; d2[1] = b; }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------ decoder
struct PChunk { uint32_t pos; int32_t match; uint32_t count; uint32_t nw; float m[16]; int8_t off[3]; uint8_t _p; };
struct DeltaDecCtx {
  const EncFrame *gi; const uint8_t *I;
  const uint8_t *p_stream; uint32_t p_len;
  PChunk *chunks; uint32_t chunk_cap;
  uint32_t *c_off;               // per chunk (+1): first output point
  uint32_t *totals;              // [0] chunks, [1] predicted points, [2] decoded macroblocks
  uint8_t *out; uint32_t out_cap;
  int color_offset;
};
// The chunk sizes chain (impl.hpp:1136-1141): one thread walks them, out of shared-memory windows of the stream that the
// whole CTA loads (a walk through global memory pays an L2 round trip per chunk: 4.1 ms for the 22 k chunks of a 1M-point
// frame, measured; through the windows 0.5 ms).  A zero size or a chunk that runs past the end stops the walk.
#define WALK_WIN 8192
__global__ void __launch_bounds__(256) pchunk_walk_kernel(DeltaDecCtx X) {
  __shared__ __align__(16) uint8_t s_win[WALK_WIN];
  __shared__ uint32_t s_pos, s_n, s_stop;
  const uint32_t extra = X.color_offset ? 3 : 0;
  if (threadIdx.x == 0) { s_pos = 0; s_n = 0; s_stop = 0; }
  __syncthreads();
  while (!s_stop) {
    const uint32_t w0 = s_pos & ~15u, wn = min((uint32_t)WALK_WIN, X.p_len - w0);   // s_pos < p_len holds here
    for (uint32_t k = threadIdx.x; k < wn; k += blockDim.x) s_win[k] = X.p_stream[w0 + k];
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t pos = s_pos, n = s_n; bool stop = false;
      while (pos < w0 + wn) {
        if (n >= X.chunk_cap) { stop = true; break; }
        const uint32_t chunk = s_win[pos - w0];
        if (chunk == 0 || chunk < 6 + extra || pos + 1 + chunk > X.p_len) { stop = true; break; }
        X.chunks[n++].pos = pos + 1;
        pos += 1 + chunk;
      }
      if (pos >= X.p_len) stop = true;
      s_pos = pos; s_n = n; s_stop = stop ? 1u : 0u;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) X.totals[0] = s_n;
}
// one thread per chunk: key -> findLeaf, transform decompression (impl.hpp:1143-1166)
__global__ void __launch_bounds__(128) pchunk_prepare_kernel(DeltaDecCtx X) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= X.totals[0]) return;
  PChunk &c = X.chunks[k];
  const uint8_t *b = X.p_stream + c.pos;
  const uint32_t extra = X.color_offset ? 3 : 0;
  const uint32_t chunk = b[-1];
  const int nw = (int)((chunk - 6 - extra) / 2);
  int16_t key[3], w[128];
  for (int a = 0; a < 3; a++) key[a] = (int16_t)((uint16_t)b[2 * a] | ((uint16_t)b[2 * a + 1] << 8));
  for (int a = 0; a < nw; a++) w[a] = (int16_t)((uint16_t)b[6 + 2 * a] | ((uint16_t)b[6 + 2 * a + 1] << 8));
  for (int a = 0; a < 3; a++) c.off[a] = extra ? (int8_t)b[6 + 2 * nw + a] : (int8_t)0;
  c.nw = (uint32_t)nw; c.match = -1; c.count = 0;
  if (nw < 6) return;                                      // fewer words than any transform: skipped (undefined in the reference)
  const EncFrame &gi = *X.gi;
  c.match = grid_find(gi, (uint32_t)(int)key[0], (uint32_t)(int)key[1], (uint32_t)(int)key[2]);
  if (c.match < 0) return;                                 // "no corresponding i block"
  c.count = gi.leaf_start[c.match + 1] - gi.leaf_start[c.match];
  rigid_decompress_d(w, nw, c.m);
}
__global__ void __launch_bounds__(1024) pchunk_scan_kernel(DeltaDecCtx X) {
  __shared__ uint64_t s_scan[33];
  const uint32_t n = X.totals[0];
  uint64_t base = 0, blocks = 0;
  for (uint32_t k0 = 0; k0 < n; k0 += 1024) {
    const uint32_t k = k0 + threadIdx.x;
    const uint64_t v = k < n ? X.chunks[k].count : 0, hit = (k < n && X.chunks[k].match >= 0) ? 1 : 0;
    uint64_t t, tb;
    const uint64_t e = block_excl_scan_u64(v, &t, s_scan);
    block_excl_scan_u64(hit, &tb, s_scan);
    if (k < n) X.c_off[k] = (uint32_t)(base + e);
    base += t; blocks += tb;
  }
  if (threadIdx.x == 0) { X.c_off[n] = (uint32_t)base; X.totals[1] = (uint32_t)base; X.totals[2] = (uint32_t)blocks; }
}
// one CTA per chunk: transformPointCloud of the I block + the colour offsets as the reference's decoder applies them
// (p.r += p.r + offset, impl.hpp:1187-1189)
__global__ void __launch_bounds__(128) pchunk_apply_kernel(DeltaDecCtx X) {
  const EncFrame &gi = *X.gi;
  const uint32_t *iv = gi.vals[gi.npasses & 1];
  __shared__ float s_m[16];
  if (X.totals[1] > X.out_cap) return;
  for (uint32_t k = blockIdx.x; k < X.totals[0]; k += gridDim.x) {
  const PChunk &c = X.chunks[k];
  if (c.match < 0) continue;
  const uint32_t is0 = gi.leaf_start[c.match];
  __syncthreads();
  if (threadIdx.x < 16) s_m[threadIdx.x] = c.m[threadIdx.x];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < c.count; i += blockDim.x) {
    const uint4 *src = (const uint4 *)(X.I + 32ull * iv[is0 + i]);
    uint4 a = src[0], b = src[1];
    float o[4]; xform_pcl_d(s_m, __uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), o);
    a = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
    if (X.color_offset) {
      const uint32_t r0 = (b.x >> 16) & 0xFF, g0 = (b.x >> 8) & 0xFF, b0 = b.x & 0xFF;
      const uint32_t rr = (r0 + (r0 + (uint32_t)(int)c.off[0])) & 0xFF, gg = (g0 + (g0 + (uint32_t)(int)c.off[1])) & 0xFF, bl = (b0 + (b0 + (uint32_t)(int)c.off[2])) & 0xFF;
      b.x = (b.x & 0xFF000000u) | (rr << 16) | (gg << 8) | bl;
    }
    uint4 *dst = (uint4 *)(X.out + 32ull * (X.c_off[k] + i));
    dst[0] = a; dst[1] = b;
  }
  }
}

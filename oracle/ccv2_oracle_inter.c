/*
 * ccv2_oracle_inter.c -- CPU oracle, inter-frame (predictive) path.  TEST INFRASTRUCTURE ONLY (see ccv2_oracle.h).
 * Textually included at the end of ccv2_oracle.c (it uses that file's static helpers).
 *
 * Restates (paths as in ccv2_oracle.h; rtc = impl/rigid_transform_coding_impl.hpp, qc = impl/quaternion_coding_impl.hpp):
 *   simplifyPCloud                    impl.hpp:318-400
 *   generate_macroblock_tree          impl.hpp:410-431
 *   do_icp_prediction                 impl.hpp:443-568
 *   encodePointCloudDeltaFrame        impl.hpp:787-1112   (the sequential branch, num_threads_ == 0; the OpenMP branch
 *                                                          computes the same thing in the same output order)
 *   decodePointCloudDeltaFrame        impl.hpp:1120-1235
 *   RigidTransformCoding              rtc:63-203
 *   QuaternionCoding                  qc:55-222
 *
 * PARITY STATUS: UNPINNED, and for the ICP step unpinnable bit for bit.  The reference delegates the registration to
 * pcl::IterativeClosestPoint (FLANN kd-tree neighbours, Eigen::umeyama on an Eigen JacobiSVD, float), whose rounding
 * depends on the Eigen version, its vectorisation and the kd-tree's tie order; two builds of the reference do not
 * produce the same bits either.  What is restated exactly is everything AROUND the registration -- block structure,
 * gates, transform quantisation, chunk format, decoder arithmetic -- and the registration's ALGORITHM as PCL 1.10 runs
 * it (point-to-point ICP, nearest neighbours without distance limit, closed-form rigid alignment per iteration,
 * DefaultConvergenceCriteria with PCL's thresholds, getFitnessScore), with an arithmetic this file defines:
 *   - neighbour distances in float, ((dx*dx + dy*dy) + dz*dz), the first minimum wins (FLANN L2_Simple accumulates in
 *     this order);
 *   - means and the 3x3 cross-covariance in double, summed in index order; the optimal rotation by Horn's quaternion
 *     method (largest eigenvector of the 4x4 symmetric matrix, cyclic Jacobi in double) -- the same rotation
 *     Eigen::umeyama's SVD gives whenever it is unique; result rounded to float like PCL's Matrix4f;
 *   - cloud transforms in float in the order of the PCL routine that does them (IterativeClosestPoint::transformCloud:
 *     ((m0*x + m1*y) + m2*z) + m3; pcl::transformPointCloud, SSE path of 1.10: x*c0 + (y*c1 + (z*c2 + c3))).
 * The CUDA path follows the same definition operation for operation, so GPU and oracle streams are compared bit for bit.
 */

#include <float.h>

/* ------------------------------------------------------------------ a voxel grid over the unit box: the octree of
 * simplifyPCloud / generate_macroblock_tree seen through its leaves ([PCL] defineBoundingBox(0,0,0,1,1,1) +
 * addPointsFromInputCloud: points outside still grow the box) */
typedef struct {
  obox b;
  size_t nf, V;
  uint64_t *codes; uint32_t *idx;       /* nf finite points sorted by Morton code, stable: indices ascend inside a leaf */
  size_t *start;                         /* V + 1 */
  uint64_t *leaf;                        /* V distinct codes, ascending = DFS leaf order */
} vgrid;
static void vgrid_free(vgrid *g) { free(g->codes); free(g->idx); free(g->start); free(g->leaf); memset(g, 0, sizeof *g); }
static int vgrid_build(const void *pts, size_t n, double res, vgrid *g) {
  memset(g, 0, sizeof *g);
  obox init; memset(&init, 0, sizeof init); init.res = res;
  for (int a = 0; a < 3; a++) { init.min[a] = 0.0; init.max[a] = 1.0; }
  get_key_bit_size_first(&init); init.defined = 1;
  uint32_t *kxyz = (uint32_t *)malloc(n * 12 + 12); uint8_t *fin = (uint8_t *)malloc(n + 1);
  g->b = init;
  int rc = bbox_keys_impl(pts, n, res, &init, g->b.min, g->b.max, &g->b.depth, kxyz, fin);
  if (rc < 0) { free(kxyz); free(fin); return rc; }
  g->codes = (uint64_t *)malloc(n * 8 + 8); g->idx = (uint32_t *)malloc(n * 4 + 4);
  for (size_t i = 0; i < n; i++) if (fin[i]) { g->codes[g->nf] = morton3(kxyz[3 * i], kxyz[3 * i + 1], kxyz[3 * i + 2], g->b.depth); g->idx[g->nf++] = (uint32_t)i; }
  free(kxyz); free(fin);
  sort_pairs(g->codes, g->idx, g->nf, 3 * g->b.depth);
  g->start = (size_t *)malloc((g->nf + 2) * sizeof(size_t)); g->leaf = (uint64_t *)malloc((g->nf + 1) * 8);
  for (size_t i = 0; i < g->nf;) { size_t j = i; while (j < g->nf && g->codes[j] == g->codes[i]) j++; g->leaf[g->V] = g->codes[i]; g->start[g->V++] = i; i = j; }
  g->start[g->V] = g->nf;
  return 0;
}
/* [PCL] findLeaf(x, y, z): a key outside the tree's range finds nothing */
static long vgrid_find(const vgrid *g, const uint32_t k[3]) {
  for (int a = 0; a < 3; a++) if (g->b.depth < 32 && (k[a] >> g->b.depth)) return -1;
  uint64_t code = morton3(k[0], k[1], k[2], g->b.depth);
  size_t lo = 0, hi = g->V;
  while (lo < hi) { size_t m = (lo + hi) / 2; if (g->leaf[m] < code) lo = m + 1; else hi = m; }
  return (lo < g->V && g->leaf[lo] == code) ? (long)lo : -1;
}

/* ------------------------------------------------------------------ simplifyPCloud (impl.hpp:318-400) */
int orc_simplify(const orc_params *p, const void *pts, size_t n, void **out, size_t *nout) {
  *out = NULL; *nout = 0;
  vgrid g; int rc = vgrid_build(pts, n, p->octree_resolution, &g);
  if (rc < 0) return rc;
  const uint8_t *base = (const uint8_t *)pts;
  uint8_t *o = (uint8_t *)calloc(g.V ? g.V : 1, 32);
  for (size_t L = 0; L < g.V; L++) {
    uint8_t *q = o + 32 * L;
    uint32_t k3[3]; demorton3(g.leaf[L], g.b.depth, k3);
    float xyz[3];
    const size_t s0 = g.start[L], s1 = g.start[L + 1];
    if (!p->do_centroid) {                                       /* [PCL] genLeafNodeCenterFromOctreeKey */
      for (int a = 0; a < 3; a++) xyz[a] = (float)(((double)k3[a] + 0.5f) * p->octree_resolution + g.b.min[a]);
    } else {                                                     /* pcl::compute3DCentroid: float sums in index order, one division */
      float acc[3] = { 0, 0, 0 };
      for (size_t k = s0; k < s1; k++) { float pf[3]; memcpy(pf, base + 32 * (size_t)g.idx[k], 12); acc[0] += pf[0]; acc[1] += pf[1]; acc[2] += pf[2]; }
      const float cnt = (float)(s1 - s0);
      for (int a = 0; a < 3; a++) xyz[a] = acc[a] / cnt;
    }
    long cs[3] = { 0, 0, 0 };                                    /* impl.hpp:383-397: long sums, (char)(sum / size) */
    for (size_t k = s0; k < s1; k++) { const uint8_t *c = base + 32 * (size_t)g.idx[k] + 16; cs[0] += c[0]; cs[1] += c[1]; cs[2] += c[2]; }
    const float one = 1.0f;                                      /* PointXYZRGB(): data[3] = 1, a = 255 */
    memcpy(q, xyz, 12); memcpy(q + 12, &one, 4);
    for (int a = 0; a < 3; a++) q[16 + a] = (uint8_t)(cs[a] / (long)(s1 - s0));
    q[19] = 255;
  }
  *out = o; *nout = g.V;
  vgrid_free(&g);
  return 0;
}

/* ------------------------------------------------------------------ QuaternionCoding (qc:55-222); q = (w, x, y, z) */
static inline float qclamp(float v) { if (v < -1) v = -1; else if (v > 1) v = 1; return v; }
static void quat_compress(const float q[4], int16_t s[3]) {
  const float scale = 1.41421f;
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  if (w > x && w > y && w > z) {
    float rx = x * scale, ry = y * scale, rz = z * scale;
    if (w < 0) { rx = -rx; ry = -ry; rz = -rz; }
    rx = qclamp(rx); ry = qclamp(ry); rz = qclamp(rz);
    s[0] = (int16_t)(rx * 32767); s[1] = (int16_t)(((int)(ry * 32767) & 0xfffe) | 1); s[2] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 1);
  } else if (z > x && z > y) {
    float rx = x * scale, ry = y * scale, rw = w * scale;
    if (z < 0) { rx = -rx; ry = -ry; rw = -rw; }
    rx = qclamp(rx); ry = qclamp(ry); rw = qclamp(rw);
    s[0] = (int16_t)(rx * 32767); s[1] = (int16_t)(((int)(ry * 32767) & 0xfffe) | 1); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 0);
  } else if (y > x) {
    float rx = x * scale, rz = z * scale, rw = w * scale;
    if (y < 0) { rx = -rx; rz = -rz; rw = -rw; }
    rx = qclamp(rx); rz = qclamp(rz); rw = qclamp(rw);
    s[0] = (int16_t)(rx * 32767); s[1] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 0); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 1);
  } else {
    float ry = y * scale, rz = z * scale, rw = w * scale;
    if (x < 0) { ry = -ry; rz = -rz; rw = -rw; }
    ry = qclamp(ry); rz = qclamp(rz); rw = qclamp(rw);
    s[0] = (int16_t)(ry * 32767); s[1] = (int16_t)(((int)(rz * 32767) & 0xfffe) | 0); s[2] = (int16_t)(((int)(rw * 32767) & 0xfffe) | 0);
  }
}
static void quat_decompress(const int16_t sin[3], float q[4]) {
  int16_t s0 = sin[0], s1 = sin[1], s2 = sin[2];
  const int which = ((s1 & 1) << 1) | (s2 & 1);
  s1 &= (int16_t)0xfffe; s2 &= (int16_t)0xfffe;
  const float scale = 1.0f / 32767.0f / 1.41421f;
  const float FE = 1.1920928955078125e-07f;
  float w, x, y, z;
  if (which == 3) { x = s0 * scale; y = s1 * scale; z = s2 * scale; w = 1 - (x * x) - (y * y) - (z * z); if (w > FE) w = sqrtf(w); }
  else if (which == 2) { x = s0 * scale; y = s1 * scale; w = s2 * scale; z = 1 - (x * x) - (y * y) - (w * w); if (z > FE) z = sqrtf(z); }
  else if (which == 1) { x = s0 * scale; z = s1 * scale; w = s2 * scale; y = 1 - (x * x) - (z * z) - (w * w); if (y > FE) y = sqrtf(y); }
  else { y = s0 * scale; z = s1 * scale; w = s2 * scale; x = 1 - (y * y) - (z * z) - (w * w); if (x > FE) x = sqrtf(x); }
  q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
/* Eigen::Quaternion<float>(Matrix3f) and ::toRotationMatrix(), float */
static void mat_to_quat(const float m[3][3], float q[4]) {
  float t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0f) {
    t = sqrtf(t + 1.0f); q[0] = 0.5f * t; t = 0.5f / t;
    q[1] = (m[2][1] - m[1][2]) * t; q[2] = (m[0][2] - m[2][0]) * t; q[3] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0; if (m[1][1] > m[0][0]) i = 1; if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(m[i][i] - m[j][j] - m[k][k] + 1.0f);
    q[1 + i] = 0.5f * t; t = 0.5f / t;
    q[0] = (m[k][j] - m[j][k]) * t; q[1 + j] = (m[j][i] + m[i][j]) * t; q[1 + k] = (m[k][i] + m[i][k]) * t;
  }
}
static void quat_to_mat(const float q[4], float r[3][3]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  r[0][0] = 1 - (tyy + tzz); r[0][1] = txy - twz; r[0][2] = txz + twy;
  r[1][0] = txy + twz; r[1][1] = 1 - (txx + tzz); r[1][2] = tyz - twx;
  r[2][0] = txz - twy; r[2][1] = tyz + twx; r[2][2] = 1 - (txx + tyy);
}

/* ------------------------------------------------------------------ RigidTransformCoding (rtc:63-203); m row-major 4x4 */
int orc_compress_rigid_transform(const float *m, int16_t *out, int *nwords) {
  const float scaling_factor = (float)((float)32767 / 2.5);
  float rot[3][3]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rot[r][c] = m[4 * r + c];
  float q[4], qt[4], res[3][3]; int16_t s[3];
  mat_to_quat(rot, q);
  quat_compress(q, s); quat_decompress(s, qt); quat_to_mat(qt, res);
  int stable = 1;
  for (int i = 0; i < 9 && stable; i++) if (fabsf(res[i / 3][i % 3] - rot[i / 3][i % 3]) > 0.001) stable = 0;   /* float difference against the double 0.001 */
  int n = 0;
  if (!stable) {                                                 /* two rows + a word of signs */
    int16_t w6 = 0;
    for (int l = 0; l < 3; l++) {
      out[l] = (int16_t)(int)(rot[0][l] * (32767 - 1)); out[l + 3] = (int16_t)(int)(rot[1][l] * (32767 - 1));
      w6 = (int16_t)(w6 + (rot[2][l] < 0 ? 1 << l : 0));
    }
    out[6] = w6; n = 7;
  } else { quat_compress(q, out); n = 3; }
  for (int a = 0; a < 3; a++) {
    float t = m[4 * a + 3];
    if (t > 2.5) t = 2.5f;
    if (t < -2.5) t = -2.5f;
    out[n++] = (int16_t)(int)(t * (scaling_factor - 1));
  }
  *nwords = n;
  return 0;
}
int orc_decompress_rigid_transform(const int16_t *in, int nwords, float *m) {
  const float scaling_factor = (float)((float)32767 / 2.5);
  float r[3][3];
  if (nwords == 6) { float q[4]; quat_decompress(in, q); quat_to_mat(q, r); }
  else {
    for (int l = 0; l < 3; l++) {
      r[0][l] = ((float)in[l]) / (32767 - 1); r[1][l] = ((float)in[l + 3]) / (32767 - 1);
      r[2][l] = sqrtf(1 - r[0][l] * r[0][l] - r[1][l] * r[1][l]);
      if (((1 << l) & ((int)in[6])) == 1 << l) r[2][l] = -r[2][l];
    }
  }
  for (int a = 0; a < 3; a++) { for (int c = 0; c < 3; c++) m[4 * a + c] = r[a][c]; m[4 * a + 3] = ((float)in[nwords - 3 + a]) / ((float)(scaling_factor - 1)); }
  m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
  return 0;
}

/* ------------------------------------------------------------------ registration */
/* pcl::transformPointCloud (SSE path): column-wise x*c0 + (y*c1 + (z*c2 + c3)) */
static inline void xform_pcl(const float *m, const float *p, float *o) {
  for (int r = 0; r < 3; r++) o[r] = p[0] * m[4 * r] + (p[1] * m[4 * r + 1] + (p[2] * m[4 * r + 2] + m[4 * r + 3]));
}
/* IterativeClosestPoint::transformCloud: Eigen Matrix4f * Vector4f(x, y, z, 1) */
static inline void xform_icp(const float *m, const float *p, float *o) {
  for (int r = 0; r < 3; r++) o[r] = ((m[4 * r] * p[0] + m[4 * r + 1] * p[1]) + m[4 * r + 2] * p[2]) + m[4 * r + 3];
}
static void nn_search(const float *q, size_t nq, const float *t, size_t nt, uint32_t *nn, float *d2) {
  for (size_t i = 0; i < nq; i++) {
    float best = INFINITY; uint32_t bj = 0;
    for (size_t j = 0; j < nt; j++) {
      const float dx = q[3 * i] - t[3 * j], dy = q[3 * i + 1] - t[3 * j + 1], dz = q[3 * i + 2] - t[3 * j + 2];
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (d < best) { best = d; bj = (uint32_t)j; }
    }
    nn[i] = bj; d2[i] = best;
  }
}
/* largest eigenvector of a symmetric 4x4 matrix, cyclic Jacobi (rows/columns p < q in order), at most 24 sweeps */
static void jacobi4_max(double A[4][4], double v[4]) {
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
/* closed-form rigid alignment of s[i] onto t[nn[i]] (TransformationEstimationSVD's optimum); T row-major 4x4 float */
static void estimate_rigid(const float *s, const float *t, const uint32_t *nn, size_t n, float *T) {
  double ms[3] = { 0, 0, 0 }, mt[3] = { 0, 0, 0 };
  for (size_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) { ms[a] += (double)s[3 * i + a]; mt[a] += (double)t[3 * (size_t)nn[i] + a]; }
  for (int a = 0; a < 3; a++) { ms[a] /= (double)n; mt[a] /= (double)n; }
  double H[3][3] = { { 0 } };
  for (size_t i = 0; i < n; i++) {
    double ds[3], dt[3];
    for (int a = 0; a < 3; a++) { ds[a] = (double)s[3 * i + a] - ms[a]; dt[a] = (double)t[3 * (size_t)nn[i] + a] - mt[a]; }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) H[a][b] += ds[a] * dt[b];
  }
  double N[4][4];
  N[0][0] = (H[0][0] + H[1][1]) + H[2][2]; N[0][1] = H[1][2] - H[2][1]; N[0][2] = H[2][0] - H[0][2]; N[0][3] = H[0][1] - H[1][0];
  N[1][1] = (H[0][0] - H[1][1]) - H[2][2]; N[1][2] = H[0][1] + H[1][0]; N[1][3] = H[2][0] + H[0][2];
  N[2][2] = (H[1][1] - H[0][0]) - H[2][2]; N[2][3] = H[1][2] + H[2][1];
  N[3][3] = (H[2][2] - H[0][0]) - H[1][1];
  for (int p = 0; p < 4; p++) for (int q = 0; q < p; q++) N[p][q] = N[q][p];
  double q[4]; jacobi4_max(N, q);
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
/* pcl::IterativeClosestPoint::align + hasConverged + getFitnessScore as do_icp_prediction drives them (impl.hpp:544-560).
 * src, tgt: packed xyz floats.  Returns the iteration count; *converged as Registration::hasConverged. */
int orc_icp(const float *src, size_t ns, const float *tgt, size_t nt, int max_iter, double tf_eps, double fit_eps, float *F, int *converged, double *fitness) {
  float *cur = (float *)malloc(ns * 12 + 12), *d2 = (float *)malloc(ns * 4 + 4); uint32_t *nn = (uint32_t *)malloc(ns * 4 + 4);
  memcpy(cur, src, ns * 12);
  for (int i = 0; i < 16; i++) F[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  double prev_mse = DBL_MAX; int iters = 0; *converged = 0;
  const double rot_thr = 1.0 - tf_eps, tr_thr = tf_eps;
  if (ns >= 3 && nt > 0) for (;;) {
    nn_search(cur, ns, tgt, nt, nn, d2);
    float T[16]; estimate_rigid(cur, tgt, nn, ns, T);
    for (size_t i = 0; i < ns; i++) { float o[3]; xform_icp(T, cur + 3 * i, o); memcpy(cur + 3 * i, o, 12); }
    float G[16];                                                 /* final_transformation_ = transformation_ * final_transformation_ */
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) G[4 * r + c] = ((T[4 * r] * F[c] + T[4 * r + 1] * F[4 + c]) + T[4 * r + 2] * F[8 + c]) + T[4 * r + 3] * F[12 + c];
    memcpy(F, G, sizeof G);
    iters++;
    /* DefaultConvergenceCriteria::hasConverged (PCL 1.10; max_iterations_similar_transforms_ = 0, absolute MSE 1e-12) */
    if (iters >= max_iter) { *converged = 1; break; }
    const double cos_angle = 0.5 * (double)(T[0] + T[5] + T[10] - 1);
    const double tr2 = (double)(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
    if (cos_angle >= rot_thr && tr2 <= tr_thr) { *converged = 1; break; }
    double mse = 0; for (size_t i = 0; i < ns; i++) mse += (double)d2[i];
    mse /= (double)ns;
    if (fabs(mse - prev_mse) < 1e-12) { *converged = 1; break; }
    if (fabs(mse - prev_mse) / prev_mse < fit_eps) { *converged = 1; break; }
    prev_mse = mse;
  }
  /* getFitnessScore: the ORIGINAL source under the final transform */
  double fs = 0;
  if (ns && nt) {
    for (size_t i = 0; i < ns; i++) xform_pcl(F, src + 3 * i, cur + 3 * i);
    nn_search(cur, ns, tgt, nt, nn, d2);
    for (size_t i = 0; i < ns; i++) fs += (double)d2[i];
    fs /= (double)ns;
  } else fs = DBL_MAX;
  *fitness = fs;
  free(cur); free(d2); free(nn);
  return iters;
}

/* do_icp_prediction (impl.hpp:443-568).  ic / pc: 32-byte records of the two blocks.  Returns 1 when a transform was found. */
static int icp_prediction(const orc_params *p, const uint8_t *ic, size_t ni, const uint8_t *pc, size_t np, float *rt, int8_t rgb_off[3]) {
  int do_icp = np > 6 ? ((np < ni * 2) && ((double)np >= (double)ni * 0.5)) : 0;
  /* NOT in the reference: a macroblock pair with more than 2^24 point pairs (only possible with macroblocks far larger than the
   * reference's 16 voxels -- 16^3 x 16^3 is exactly 2^24) is not registered; its points are coded intra.  The exhaustive
   * neighbour search of this restatement and of the CUDA kernel (one warp per macroblock) would take seconds per block there. */
  if ((uint64_t)np * (uint64_t)ni > (1ull << 24)) do_icp = 0;
  if (!do_icp) return 0;
  double in_av[3] = { 0, 0, 0 }, out_av[3] = { 0, 0, 0 }, in_var = 0, out_var = 0;   /* index 0,1,2 = r,g,b (record bytes 18,17,16) */
  for (size_t i = 0; i < ni; i++) for (int a = 0; a < 3; a++) in_av[a] += (double)ic[32 * i + 18 - a];
  for (int a = 0; a < 3; a++) in_av[a] /= (double)ni;
  for (size_t i = 0; i < ni; i++) {
    const double r = ic[32 * i + 18] - in_av[0], g = ic[32 * i + 17] - in_av[1], b = ic[32 * i + 16] - in_av[2];
    in_var += r * r + g * g + b * b;
  }
  in_var /= (double)(3 * ni);
  for (size_t i = 0; i < np; i++) for (int a = 0; a < 3; a++) out_av[a] += (double)pc[32 * i + 18 - a];
  for (int a = 0; a < 3; a++) out_av[a] /= (double)np;
  for (size_t i = 0; i < np; i++) {
    const double r = pc[32 * i + 18] - out_av[0], g = pc[32 * i + 17] - out_av[1], b = pc[32 * i + 16] - out_av[2];
    out_var += r * r + g * g + b * b;
  }
  out_var /= (double)(3 * np);
  if (in_var > 100.0f || out_var > 100.0f) do_icp = 0;           /* icp_var_threshold_ (codec.h:139) */
  if (p->do_icp_color_offset) for (int a = 0; a < 3; a++) if (fabs(out_av[a] - in_av[a]) < 32) rgb_off[a] = (int8_t)(out_av[a] - in_av[a]);
  if (!do_icp) return 0;
  float *s = (float *)malloc(ni * 12 + 12), *t = (float *)malloc(np * 12 + 12);
  for (size_t i = 0; i < ni; i++) memcpy(s + 3 * i, ic + 32 * i, 12);
  for (size_t i = 0; i < np; i++) memcpy(t + 3 * i, pc + 32 * i, 12);
  const float eps = 1e-8f;                                       /* transformationepsilon_ is a float (codec.h:142,309) */
  int conv; double fit;
  orc_icp(s, ni, t, np, 50, (double)eps, (double)(3 * eps), rt, &conv, &fit);
  free(s); free(t);
  return conv && fit < p->point_resolution * 2;
}

/* the intra coder both directions construct for what cannot be predicted (impl.hpp:1089-1101, 1208-1220): ten explicit
 * constructor arguments, the rest are the class defaults (scalable stream on, JPEG quality 75, codec.h:108-143) */
static void delta_intra_params(const orc_params *p, orc_params *q) {
  *q = *p;
  q->do_voxel_grid = 1; q->do_color = 1; q->create_scalable = 1; q->code_connectivity = 0; q->jpeg_quality = 75;
  q->macroblock_size = 16; q->do_icp_color_offset = 0;
}

/* encodePointCloudDeltaFrame (impl.hpp:787-1112) */
int orc_encode_delta(const orc_params *p, const void *icloud, size_t ni, const void *pcloud, size_t np, int icp_on_original,
                     uint8_t **i_out, size_t *i_len, uint8_t **p_out, size_t *p_len, void **out_cloud, size_t *n_out_cloud, orc_delta_info *info) {
  *i_out = NULL; *i_len = 0; *p_out = NULL; *p_len = 0;
  if (out_cloud) { *out_cloud = NULL; *n_out_cloud = 0; }
  if (info) memset(info, 0, sizeof *info);
  void *simp = NULL; size_t nsimp = 0;
  const uint8_t *P = (const uint8_t *)pcloud; size_t nP = np;
  int rc;
  if (!icp_on_original) { if ((rc = orc_simplify(p, pcloud, np, &simp, &nsimp)) < 0) return rc; P = (const uint8_t *)simp; nP = nsimp; }
  const uint8_t *I = (const uint8_t *)icloud;
  const double mres = p->octree_resolution * p->macroblock_size;
  vgrid gi, gp;
  if ((rc = vgrid_build(I, ni, mres, &gi)) < 0) { free(simp); return rc; }
  if ((rc = vgrid_build(P, nP, mres, &gp)) < 0) { vgrid_free(&gi); free(simp); return rc; }
  bbuf ps = {0}, intra = {0}, oc = {0};
  uint64_t shared = 0, conv = 0;
  for (size_t L = 0; L < gp.V; L++) {
    uint32_t k3[3]; demorton3(gp.leaf[L], gp.b.depth, k3);
    const size_t s0 = gp.start[L], cnt = gp.start[L + 1] - s0;
    uint8_t *pc = (uint8_t *)malloc(cnt * 32);
    for (size_t k = 0; k < cnt; k++) memcpy(pc + 32 * k, P + 32 * (size_t)gp.idx[s0 + k], 32);
    const long Li = vgrid_find(&gi, k3);
    int ok = 0;
    if (Li >= 0) {
      shared++;
      const size_t i0 = gi.start[Li], icnt = gi.start[Li + 1] - i0;
      uint8_t *ic = (uint8_t *)malloc(icnt * 32);
      for (size_t k = 0; k < icnt; k++) memcpy(ic + 32 * k, I + 32 * (size_t)gi.idx[i0 + k], 32);
      float rt[16]; int8_t off[3] = { 0, 0, 0 };
      ok = icp_prediction(p, ic, icnt, pc, cnt, rt, off);
      if (ok) {
        conv++;
        int16_t comp[10]; int nw;
        orc_compress_rigid_transform(rt, comp, &nw);
        int16_t key[3] = { (int16_t)(int)k3[0], (int16_t)(int)k3[1], (int16_t)(int)k3[2] };
        const uint8_t chunk = (uint8_t)(6 + 2 * nw + (p->do_icp_color_offset ? 3 : 0));
        bb_push(&ps, chunk); bb_write(&ps, key, 6); bb_write(&ps, comp, 2 * (size_t)nw);
        if (p->do_icp_color_offset) bb_write(&ps, off, 3);
        if (out_cloud) {
          float mdec[16]; orc_decompress_rigid_transform(comp, nw, mdec);
          for (size_t k = 0; k < icnt; k++) {
            uint8_t rec[32]; memcpy(rec, ic + 32 * k, 32);
            float xi[3], xo[4]; memcpy(xi, rec, 12); xform_pcl(mdec, xi, xo);
            xo[3] = xi[0] * mdec[12] + (xi[1] * mdec[13] + (xi[2] * mdec[14] + mdec[15]));
            memcpy(rec, xo, 16);
            if (p->do_icp_color_offset) { rec[18] = (uint8_t)(rec[18] + off[0]); rec[17] = (uint8_t)(rec[17] + off[1]); rec[16] = (uint8_t)(rec[16] + off[2]); }
            bb_write(&oc, rec, 32);
          }
        }
      }
      free(ic);
    }
    if (!ok) { bb_write(&intra, pc, cnt * 32); if (out_cloud) bb_write(&oc, pc, cnt * 32); }
    free(pc);
  }
  orc_params q; delta_intra_params(p, &q);
  rc = orc_encode(&q, 1, intra.p, intra.n / 32, i_out, i_len, NULL, NULL);
  if (info) {
    info->macro_blocks = gp.V; info->shared_blocks = shared; info->converged_blocks = conv; info->n_intra_points = intra.n / 32; info->n_p_points = nP;
    info->shared_percentage = (float)shared / (float)gp.V; info->convergence_percentage = (float)conv / (float)shared;
  }
  *p_out = ps.p; *p_len = ps.n;
  if (out_cloud) { *out_cloud = oc.p; *n_out_cloud = oc.n / 32; } else free(oc.p);
  free(intra.p); free(simp); vgrid_free(&gi); vgrid_free(&gp);
  return rc;
}

/* decodePointCloudDeltaFrame (impl.hpp:1120-1235) */
int orc_decode_delta(const orc_params *p, const void *icloud, size_t ni, const uint8_t *i_in, size_t i_len, const uint8_t *p_in, size_t p_len,
                     void **out, size_t *nout, uint64_t *decoded_blocks) {
  *out = NULL; *nout = 0;
  const uint8_t *I = (const uint8_t *)icloud;
  vgrid gi; int rc = vgrid_build(I, ni, p->octree_resolution * p->macroblock_size, &gi);
  if (rc < 0) return rc;
  bbuf oc = {0}; uint64_t nblocks = 0;
  size_t pos = 0;
  const size_t extra = p->do_icp_color_offset ? 3 : 0;
  while (pos < p_len) {
    const uint8_t chunk = p_in[pos++];
    if (chunk == 0) break;
    if (chunk < 6 + extra || pos + chunk > p_len) break;         /* a truncated stream: the reference's reads fail and its loop ends */
    int16_t key[3]; memcpy(key, p_in + pos, 6);
    const int nw = (int)((chunk - 6 - extra) / 2);
    int16_t comp[128]; memcpy(comp, p_in + pos + 6, 2 * (size_t)nw);
    int8_t off[3] = { 0, 0, 0 }; if (extra) memcpy(off, p_in + pos + 6 + 2 * nw, 3);
    pos += chunk;
    if (nw < 6) continue;                                        /* fewer words than any transform: undefined in the reference, skipped here */
    const uint32_t k3[3] = { (uint32_t)(int)key[0], (uint32_t)(int)key[1], (uint32_t)(int)key[2] };
    const long Li = vgrid_find(&gi, k3);
    if (Li < 0) continue;                                        /* "no corresponding i block" */
    float mdec[16]; orc_decompress_rigid_transform(comp, nw, mdec);
    nblocks++;
    for (size_t k = gi.start[Li]; k < gi.start[Li + 1]; k++) {
      uint8_t rec[32]; memcpy(rec, I + 32 * (size_t)gi.idx[k], 32);
      float xi[3], xo[4]; memcpy(xi, rec, 12); xform_pcl(mdec, xi, xo);
      xo[3] = xi[0] * mdec[12] + (xi[1] * mdec[13] + (xi[2] * mdec[14] + mdec[15]));
      memcpy(rec, xo, 16);
      if (extra) {                                               /* impl.hpp:1187-1189: p.r += p.r + offset (the doubling is the reference's) */
        rec[18] = (uint8_t)(rec[18] + (rec[18] + off[0])); rec[17] = (uint8_t)(rec[17] + (rec[17] + off[1])); rec[16] = (uint8_t)(rec[16] + (rec[16] + off[2]));
      }
      bb_write(&oc, rec, 32);
    }
  }
  if (i_len) {
    void *ip = NULL; size_t n = 0;
    rc = orc_decode(i_in, i_len, &ip, &n, NULL);
    if (rc == 0 && n) bb_write(&oc, ip, n * 32);
    free(ip);
    if (rc == -1) rc = 0;                                        /* no header found: decodePointCloud returns silently (impl.hpp:231) */
  }
  if (decoded_blocks) *decoded_blocks = nblocks;
  *out = oc.p; *nout = oc.n / 32;
  vgrid_free(&gi);
  return rc;
}

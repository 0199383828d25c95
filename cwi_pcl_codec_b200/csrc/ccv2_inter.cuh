// ccv2_inter.cuh -- host side of the inter-frame (predictive) path; textually included at the end of ccv2_api.cu.
// encodePointCloudDeltaFrame / decodePointCloudDeltaFrame (impl.hpp:787-1112, 1120-1235) as three front-end runs on a
// host-defined unit box (the P frame's voxel grid, the two macroblock trees), the macroblock kernels of inter_kernels.cuh,
// and one call into a child codec for the points no macroblock predicted (the reference constructs a fresh intra coder
// per delta frame, impl.hpp:1089-1101 / 1208-1220).  The calls are synchronous: one host wait for the counts the intra
// coder's call needs, one at the end.
#pragma once

// like CU(), but nothing stays queued behind the caller's back when a call fails half way: the inter path runs on fin_stream
#define CUI(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e__); cudaStreamSynchronize(c->fin_stream); cudaGetLastError(); return CCV2_ERR_CUDA; } } while (0)

namespace {

// workspace of one grid (the front-end arrays a leaf scan needs; no JPEG or tree buffers)
size_t carve_grid(uint8_t *base, size_t n, EncFrame *f, size_t *zero_bytes) {
  Carver cv(base);
  const uint32_t tiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE) + 1;
  const uint32_t scan_tiles = (uint32_t)(n / 1024) + 8;
  uint32_t *ghist = cv.take<uint32_t>(8 * 256);
  uint32_t *sort_status = cv.take<uint32_t>((size_t)8 * tiles * 256);
  uint64_t *scan_status = cv.take<uint64_t>((size_t)4 * scan_tiles);
  const size_t z1 = cv.end();
  uint64_t *k0 = cv.take<uint64_t>(n + 8), *k1 = cv.take<uint64_t>(n + 8);
  uint32_t *v0 = cv.take<uint32_t>(n + 8), *v1 = cv.take<uint32_t>(n + 8);
  uint64_t *leaf_key = cv.take<uint64_t>(n + 8);
  uint32_t *leaf_start = cv.take<uint32_t>(n + 8), *leaf_off = cv.take<uint32_t>(n + 8);
  uint8_t *first_new = cv.take<uint8_t>(n + 8);
  if (f) {
    f->ghist = ghist; f->sort_status = sort_status; f->tiles_max = tiles; f->scan_status = scan_status; f->scan_tiles_max = scan_tiles;
    f->keys[0] = k0; f->keys[1] = k1; f->vals[0] = v0; f->vals[1] = v1;
    f->leaf_key = leaf_key; f->leaf_start = leaf_start; f->leaf_off = leaf_off; f->first_new = first_new;
    f->zero_ptr = base; f->zero_bytes = z1;
  }
  if (zero_bytes) *zero_bytes = z1;
  return cv.end();
}

// [PCL] defineBoundingBox(0, 0, 0, 1, 1, 1) -> getKeyBitSize() with no leaves: depth from the box, the box centred in the
// octree's cube (oracle: get_key_bit_size_first).  Also entry 0 of the record's box log.
int define_unit_box(EncFrame &f, double res) {
  const double eps = 1.1920928955078125e-07;
  double mn[3] = { 0, 0, 0 }, mx[3] = { 1, 1, 1 };
  uint32_t mk = 2;
  for (int a = 0; a < 3; a++) {
    const double t = std::ceil((mx[a] - mn[a] - eps) / res);
    const uint32_t k = t >= 4294967295.0 ? 0xFFFFFFFFu : (t > 0 ? (uint32_t)t : 0u);
    mk = std::max(mk, k);
  }
  uint32_t d = 0; while ((1ull << d) < mk) d++;
  if (d > CCV2_MAX_DEPTH) return CCV2_ERR_DEPTH;
  const double side = (double)(1u << d) * res;
  for (int a = 0; a < 3; a++) { const double over = (side - (mx[a] - mn[a])) / 2.0; if (over > eps) { mn[a] -= over; mx[a] += over; } }
  f.defined = 1; f.depth = d; f.n_events = 1;
  for (int a = 0; a < 3; a++) { f.bmin[a] = mn[a]; f.bmax[a] = mx[a]; f.ev[0].mn[a] = mn[a]; }
  f.ev[0].idx = 0; f.ev[0].depth_before = 0; f.ev[0].mask = 0; f.ev[0]._pad = 0;
  return CCV2_OK;
}

EncParams grid_params(double res, bool centroid, bool packed) {
  EncParams P; memset(&P, 0, sizeof P);
  P.res = res;
  { int ex; const double m = frexp(res, &ex); P.res_pow2 = (m == 0.5); P.inv_res = P.res_pow2 ? 1.0 / res : 0.0; }
  P.do_color = 1; P.color_type = 3; P.do_centroid = centroid ? 1 : 0; P.prefix_len = 16384; P.allow_packed = packed ? 1 : 0;
  return P;
}

// bbox -> keys -> radix sort -> leaf scan of ONE record (device pointer d) holding at most n_upper points
void launch_grid(cudaStream_t st, EncFrame *d, const EncParams &P, size_t n_upper, uint64_t &launches) {
  const size_t gn = std::max<size_t>(n_upper, 1);
  const unsigned gx256 = (unsigned)((gn + 255) / 256), gtiles = (unsigned)((gn + SORT_TILE - 1) / SORT_TILE);
  zero_region_kernel<EncFrame><<<dim3(64, 1), 256, 0, st>>>(d);
  bbox_kernel<<<1, 1024, 0, st>>>(d, P, 0);
  keygen_kernel<<<dim3(gx256, 1), 256, 0, st>>>(d, P, 0);
  bbox_kernel<<<1, 1024, 0, st>>>(d, P, 1);
  keygen_kernel<<<dim3(gx256, 1), 256, 0, st>>>(d, P, 1);
  grid_setup_kernel<<<1, 32, 0, st>>>(d, nullptr);
  sort_hist_kernel<<<dim3(gtiles, 1), 256, 0, st>>>(d);
  for (int p = 0; p < 8; p++) sort_pass_kernel<<<dim3(gtiles, 1), SORT_THREADS, SORT_SMEM_BYTES, st>>>(d, p);
  leaf_scan_kernel<<<dim3((unsigned)((gn + LEAF_TILE - 1) / LEAF_TILE), 1), LEAF_THREADS, 0, st>>>(d, 0);
  launches += 15;
}

int grid_error(ccv2_codec *c, const EncFrame &f, const char *what) {
  if (f.error & FERR_DEPTH) { c->err = std::string(what) + ": octree depth > 21"; return CCV2_ERR_DEPTH; }
  if (f.error) { char b[96]; snprintf(b, sizeof b, "%s: device error bits 0x%x", what, f.error); c->err = b; return CCV2_ERR_WORKSPACE; }
  return CCV2_OK;
}

// the codec both directions use for what cannot be predicted: ten explicit constructor arguments, the rest class defaults
int get_intra_child(ccv2_codec *c, ccv2_codec **out) {
  if (!c->intra_child) {
    ccv2_params q = c->prm;
    q.profile = CCV2_MANUAL_CONFIGURATION; q.show_statistics = 0; q.do_voxel_grid_downsampling = 1; q.i_frame_rate = 0; q.do_color_encoding = 1;
    q.create_scalable_stream = 1; q.code_connectivity = 0; q.jpeg_quality = 75; q.num_threads = 0; q.macroblock_size = 16; q.do_icp_color_offset = 0;
    const int rc = ccv2_create(&q, c->device, &c->intra_child);
    if (rc != CCV2_OK) { c->err = std::string("intra coder of the delta frame: ") + ccv2_last_error(nullptr); return rc; }
  }
  *out = c->intra_child;
  return CCV2_OK;
}

}  // namespace

extern "C" {

size_t ccv2_max_p_stream_size(size_t np) { return 30 * np + 64; }

int ccv2_simplify(ccv2_codec *c, const void *pts, size_t n, void *pts_out, size_t cap_points, size_t *npts) {
  if (!c || !npts || (n && !pts) || n >= (1u << 28)) return CCV2_ERR_ARG;
  CUI(cudaSetDevice(c->device));
  finish_all(c);
  *npts = 0;
  if (n == 0) return CCV2_OK;
  const bool cen = c->prm.do_voxel_grid_centroid != 0;
  const bool din = is_device_ptr(pts), dout = pts_out && is_device_ptr(pts_out);
  Carver m(nullptr);
  m.take<EncFrame>(1); if (!din) m.take<uint8_t>(32 * n); m.take<uint8_t>(32 * n); const size_t goff = m.end();
  const size_t gbytes = carve_grid(nullptr, n, nullptr, nullptr);
  CUI(c->inter_ws.ensure(goff + gbytes + 256));
  Carver cv(c->inter_ws.p);
  EncFrame *d_rec = cv.take<EncFrame>(1);
  uint8_t *d_in = din ? (uint8_t *)pts : cv.take<uint8_t>(32 * n);
  uint8_t *d_s = cv.take<uint8_t>(32 * n);
  EncFrame h; memset(&h, 0, sizeof h);
  carve_grid((uint8_t *)c->inter_ws.p + goff, n, &h, nullptr);
  h.pts = d_in; h.n = (uint32_t)n; h.n_finite = (uint32_t)n; h.violator = NONE_U32;
  int rc = define_unit_box(h, c->prm.octree_resolution);
  if (rc) { c->err = "octree depth > 21"; return rc; }
  cudaStream_t st = c->fin_stream;
  uint64_t launches = 0;
  if (!din) CUI(cudaMemcpyAsync(d_in, pts, 32 * n, cudaMemcpyHostToDevice, st));
  CUI(cudaMemcpyAsync(d_rec, &h, sizeof h, cudaMemcpyHostToDevice, st));
  const EncParams P = grid_params(c->prm.octree_resolution, cen, !cen && c->allow_packed);
  launch_grid(st, d_rec, P, n, launches);
  simplify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rec, P, d_s);
  CUI(cudaGetLastError());
  CUI(cudaMemcpyAsync(&h, d_rec, sizeof h, cudaMemcpyDeviceToHost, st));
  CUI(cudaStreamSynchronize(st));
  if ((rc = grid_error(c, h, "simplify")) != CCV2_OK) return rc;
  *npts = h.V;
  c->launches = launches + 1;
  if (h.V == 0) return CCV2_OK;
  if (h.V > cap_points || !pts_out) return CCV2_ERR_CAPACITY;
  CUI(cudaMemcpy(pts_out, d_s, 32ull * h.V, dout ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  return CCV2_OK;
}

}  // extern "C"

// The prediction stage of one delta frame: grids, matching, ICP, P stream (copied to the caller), predicted frame; the points
// no macroblock predicted go to intra_dst (device memory, room for np records; nullptr: a buffer in the workspace, returned
// through *intra_ptr).  One host wait at the end (the counts).
static int delta_predict(ccv2_codec *c, const void *icloud, size_t ni, const void *pcloud, size_t np, int icp_on_original, uint8_t *intra_dst, const uint8_t **intra_ptr, size_t *n_intra,
                         void *p_out, size_t p_cap, size_t *p_len, void *out_cloud, size_t out_cap_points, size_t *n_out, ccv2_delta_info *info, uint64_t &launches) {
  const ccv2_params &prm = c->prm;
  const bool cen = prm.do_voxel_grid_centroid != 0, orig = icp_on_original != 0, want_out = out_cloud != nullptr;
  const double res = prm.octree_resolution, mres = prm.octree_resolution * prm.macroblock_size;
  const bool di = !ni || is_device_ptr(icloud), dp = !np || is_device_ptr(pcloud);
  const size_t ncI = std::max<size_t>(ni, 1), ncP = std::max<size_t>(np, 1);
  const size_t g_p = carve_grid(nullptr, ncP, nullptr, nullptr), g_i = carve_grid(nullptr, ncI, nullptr, nullptr);
  // ---- carve (first pass measures)
  struct Lay { EncFrame *rec; uint8_t *dI, *dP, *S, *g0, *g1, *g2; MbResult *res; float *cur, *tgt, *d2; uint32_t *nn, *p_off, *x_off, *o_off, *ticket; uint8_t *pstr, *intra, *outp; } L;
  auto carve = [&](void *base) {
    Carver cv(base);
    L.rec = cv.take<EncFrame>(3);
    L.ticket = cv.take<uint32_t>(16);
    L.dI = di ? nullptr : cv.take<uint8_t>(32 * ncI); L.dP = dp ? nullptr : cv.take<uint8_t>(32 * ncP);
    L.S = orig ? nullptr : cv.take<uint8_t>(32 * ncP);
    L.g0 = orig ? nullptr : cv.take<uint8_t>(g_p); L.g1 = cv.take<uint8_t>(g_p); L.g2 = cv.take<uint8_t>(g_i);
    L.res = cv.take<MbResult>(ncP + 1);
    L.cur = cv.take<float>(3 * ncI + 4); L.tgt = cv.take<float>(3 * ncP + 4); L.d2 = cv.take<float>(ncI + 4); L.nn = cv.take<uint32_t>(ncI + 4);
    L.p_off = cv.take<uint32_t>(ncP + 2); L.x_off = cv.take<uint32_t>(ncP + 2); L.o_off = cv.take<uint32_t>(ncP + 2);
    L.pstr = cv.take<uint8_t>(ccv2_max_p_stream_size(ncP)); L.intra = intra_dst ? intra_dst : cv.take<uint8_t>(32 * ncP);
    L.outp = want_out ? cv.take<uint8_t>(32 * (ncI + ncP)) : nullptr;
    return cv.end();
  };
  const size_t total = carve(nullptr);
  CUI(c->inter_ws.ensure(total + 256));
  carve(c->inter_ws.p);
  const uint8_t *I = di ? (const uint8_t *)icloud : L.dI, *Praw = dp ? (const uint8_t *)pcloud : L.dP;
  const uint8_t *Pc = orig ? Praw : L.S;                      // the cloud the P macroblock tree indexes
  // ---- records
  EncFrame h[3]; memset(h, 0, sizeof h);
  int rc;
  if (!orig) {
    carve_grid(L.g0, ncP, &h[0], nullptr);
    h[0].pts = Praw; h[0].n = (uint32_t)np; h[0].n_finite = (uint32_t)np; h[0].violator = NONE_U32;
    if ((rc = define_unit_box(h[0], res)) != CCV2_OK) { c->err = "octree depth > 21"; return rc; }
  }
  carve_grid(L.g1, ncP, &h[1], nullptr);
  h[1].pts = Pc; h[1].n = (uint32_t)(orig ? np : 0); h[1].n_finite = h[1].n; h[1].violator = NONE_U32;     // simplified: the count arrives on the device
  carve_grid(L.g2, ncI, &h[2], nullptr);
  h[2].pts = I; h[2].n = (uint32_t)ni; h[2].n_finite = (uint32_t)ni; h[2].violator = NONE_U32;
  if ((rc = define_unit_box(h[1], mres)) != CCV2_OK || (rc = define_unit_box(h[2], mres)) != CCV2_OK) { c->err = "octree depth > 21"; return rc; }
  // ---- enqueue
  cudaStream_t st = c->fin_stream;
  if (!c->inter_ev0) { CUI(cudaEventCreate(&c->inter_ev0)); CUI(cudaEventCreate(&c->inter_ev1)); }
  CUI(cudaEventRecord(c->inter_ev0, st));
  if (!di && ni) CUI(cudaMemcpyAsync(L.dI, icloud, 32 * ni, cudaMemcpyHostToDevice, st));
  if (!dp && np) CUI(cudaMemcpyAsync(L.dP, pcloud, 32 * np, cudaMemcpyHostToDevice, st));
  CUI(cudaMemcpyAsync(L.rec, h, sizeof h, cudaMemcpyHostToDevice, st));
  CUI(cudaMemsetAsync(L.ticket, 0, 64, st));
  if (!orig) {
    const EncParams P0 = grid_params(res, cen, !cen && c->allow_packed);
    launch_grid(st, L.rec + 0, P0, np, launches);
    simplify_kernel<<<(unsigned)((ncP + 255) / 256), 256, 0, st>>>(L.rec + 0, P0, L.S);
    grid_chain_kernel<<<1, 32, 0, st>>>(L.rec + 1, L.rec + 0);
    launches += 2;
  }
  const EncParams PM = grid_params(mres, false, false);
  launch_grid(st, L.rec + 1, PM, np, launches);
  launch_grid(st, L.rec + 2, PM, ni, launches);
  InterCtx X; memset(&X, 0, sizeof X);
  X.gp = L.rec + 1; X.gi = L.rec + 2; X.P = Pc; X.I = I; X.res = L.res;
  X.cur = L.cur; X.tgt = L.tgt; X.d2 = L.d2; X.nn = L.nn; X.p_off = L.p_off; X.x_off = L.x_off; X.o_off = L.o_off;
  X.p_stream = L.pstr; X.intra_pts = L.intra; X.out_pts = L.outp; X.ticket = L.ticket;
  X.color_offset = prm.do_icp_color_offset != 0; X.max_iter = 50;                        // icp_max_iterations_ (codec.h:140)
  const float tfe = 1e-8f;                                                                // transformationepsilon_ is a float (codec.h:142,309)
  X.point_res = prm.point_resolution; X.tf_eps = (double)tfe; X.fit_eps = (double)(3 * tfe);
  mb_match_kernel<<<(unsigned)((ncP + 255) / 256), 256, 0, st>>>(X);
  mb_icp_kernel<<<c->n_sm * 32, ICP_THREADS, 0, st>>>(X);
  mb_scan_kernel<<<1, 1024, 0, st>>>(X, want_out ? 1 : 0);
  mb_write_kernel<<<c->n_sm * 8, 128, 0, st>>>(X);
  launches += 4;
  CUI(cudaGetLastError());
  CUI(cudaEventRecord(c->inter_ev1, st));
  uint32_t tot[16];
  CUI(cudaMemcpyAsync(tot, L.ticket, 64, cudaMemcpyDeviceToHost, st));
  CUI(cudaMemcpyAsync(h, L.rec, sizeof h, cudaMemcpyDeviceToHost, st));
  CUI(cudaStreamSynchronize(st));
  for (int g = orig ? 1 : 0; g < 3; g++) if ((rc = grid_error(c, h[g], g == 0 ? "P voxel grid" : g == 1 ? "P macroblock tree" : "I macroblock tree")) != CCV2_OK) return rc;
  const size_t plen = tot[1], nx = tot[2], nout = tot[3];
  float pms = 0; cudaEventElapsedTime(&pms, c->inter_ev0, c->inter_ev1);
  if (info) {
    info->macro_blocks = tot[6]; info->shared_blocks = tot[4]; info->converged_blocks = tot[5]; info->n_intra_points = nx; info->n_p_points = h[1].n;
    info->shared_percentage = (float)tot[4] / (float)tot[6]; info->convergence_percentage = (float)tot[5] / (float)tot[4];
    info->predict_ms = pms;
  }
  c->mb_percentage = (float)tot[4] / (float)tot[6]; c->mb_convergence = (float)tot[5] / (float)tot[4];
  // ---- results
  *p_len = plen;
  if (plen) {
    if (!p_out || plen > p_cap) { c->err = "P stream buffer too small"; return CCV2_ERR_CAPACITY; }
    CUI(cudaMemcpy(p_out, L.pstr, plen, is_device_ptr(p_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  }
  if (want_out) {
    *n_out = nout;
    if (nout > out_cap_points) { c->err = "predicted-frame buffer too small"; return CCV2_ERR_CAPACITY; }
    if (nout) CUI(cudaMemcpy(out_cloud, L.outp, 32 * nout, is_device_ptr(out_cloud) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  }
  *n_intra = nx; if (intra_ptr) *intra_ptr = L.intra;
  return CCV2_OK;
}

// the unpredicted points of n delta frames through the child codec in ONE pipelined call; every frame is written by "a fresh
// intra coder" (impl.hpp:1089-1101): frame id 1 each, an empty cloud writes nothing (impl.hpp:206-212)
static int delta_intra_encode(ccv2_codec *c, int n, const void *const *pts, const size_t *npts, void *const *out, const size_t *cap, size_t *len, float *ms, uint64_t &launches) {
  ccv2_codec *ch = nullptr;
  int rc = get_intra_child(c, &ch);
  if (rc != CCV2_OK) return rc;
  int ticket = 0;
  rc = submit_call(ch, 0, n, pts, npts, out, cap, len, nullptr, nullptr, nullptr, nullptr, nullptr, false, /*fixed frame id*/ 1, &ticket);
  if (rc == CCV2_OK) rc = wait_ticket(ch, ticket);
  if (rc != CCV2_OK) { c->err = std::string("intra coder of the delta frame: ") + ccv2_last_error(ch); return rc; }
  launches += ccv2_last_launch_count(ch);
  if (ms) *ms = ccv2_last_device_ms(ch);
  return CCV2_OK;
}

extern "C" {

int ccv2_encode_delta(ccv2_codec *c, const void *icloud, size_t ni, const void *pcloud, size_t np, int icp_on_original,
                      void *i_out, size_t i_cap, size_t *i_len, void *p_out, size_t p_cap, size_t *p_len,
                      void *out_cloud, size_t out_cap_points, size_t *n_out, ccv2_delta_info *info) {
  if (!c || !i_len || !p_len || (ni && !icloud) || (np && !pcloud) || ni >= (1u << 28) || np >= (1u << 27) || (out_cloud && !n_out)) return CCV2_ERR_ARG;   // np < 2^27: the P stream's byte offsets are 32-bit (30 bytes per macroblock at most)
  if (c->prm.macroblock_size < 1) { c->err = "macroblock_size must be >= 1"; return CCV2_ERR_ARG; }
  CUI(cudaSetDevice(c->device));
  c->err.clear();                                            // no finish_all: the delta path shares no workspace with submitted intra calls, which may stay in flight
  *i_len = 0; *p_len = 0; if (n_out) *n_out = 0;
  ccv2_delta_info li; memset(&li, 0, sizeof li);
  uint64_t launches = 0;
  const uint8_t *ip = nullptr; size_t nx = 0;
  int rc = delta_predict(c, icloud, ni, pcloud, np, icp_on_original, nullptr, &ip, &nx, p_out, p_cap, p_len, out_cloud, out_cap_points, n_out, &li, launches);
  if (rc == CCV2_OK && nx) {
    if (!i_out) { c->err = "I stream buffer missing"; rc = CCV2_ERR_CAPACITY; }
    else { const void *p1 = ip; void *o1 = i_out; rc = delta_intra_encode(c, 1, &p1, &nx, &o1, &i_cap, i_len, &li.intra_ms, launches); }
  }
  if (info) *info = li;
  c->launches = launches; c->device_ms = li.predict_ms + li.intra_ms;
  return rc;
}

// n delta frames in one call: the prediction stages one after the other, then ALL the intra parts as one batch through the
// child codec, whose serial range-coder stage is latency bound -- 29 frames take about as long as one.
int ccv2_encode_delta_batch(ccv2_codec *c, int nframes, const void *const *icloud, const size_t *ni, const void *const *pcloud, const size_t *np, int icp_on_original,
                            void *const *i_out, const size_t *i_cap, size_t *i_len, void *const *p_out, const size_t *p_cap, size_t *p_len, ccv2_delta_info *info) {
  if (!c || nframes < 0 || (nframes && (!icloud || !ni || !pcloud || !np || !i_out || !i_cap || !i_len || !p_out || !p_cap || !p_len))) return CCV2_ERR_ARG;
  if (c->prm.macroblock_size < 1) { c->err = "macroblock_size must be >= 1"; return CCV2_ERR_ARG; }
  CUI(cudaSetDevice(c->device));
  c->err.clear();                                            // no finish_all: the delta path shares no workspace with submitted intra calls, which may stay in flight
  size_t tot = 0; std::vector<size_t> off(nframes + 1, 0);
  for (int k = 0; k < nframes; k++) {
    if ((ni[k] && !icloud[k]) || (np[k] && !pcloud[k]) || ni[k] >= (1u << 28) || np[k] >= (1u << 27)) return CCV2_ERR_ARG;
    off[k] = tot; tot += (32 * std::max<size_t>(np[k], 1) + 255) & ~size_t(255);
    i_len[k] = 0; p_len[k] = 0;
  }
  CUI(c->inter_batch.ensure(tot + 256));
  uint64_t launches = 0;
  std::vector<const void *> ip; std::vector<size_t> nx(nframes, 0), icap, ilen; std::vector<void *> io; std::vector<int> who;
  float pms = 0;
  for (int k = 0; k < nframes; k++) {
    ccv2_delta_info li; memset(&li, 0, sizeof li);
    const int rc = delta_predict(c, icloud[k], ni[k], pcloud[k], np[k], icp_on_original, (uint8_t *)c->inter_batch.p + off[k], nullptr, &nx[k], p_out[k], p_cap[k], &p_len[k], nullptr, 0, nullptr, &li, launches);
    if (info) info[k] = li;
    if (rc != CCV2_OK) return rc;
    pms += li.predict_ms;
    if (nx[k]) {
      if (!i_out[k]) { c->err = "I stream buffer missing"; return CCV2_ERR_CAPACITY; }
      ip.push_back((const uint8_t *)c->inter_batch.p + off[k]); io.push_back(i_out[k]); icap.push_back(i_cap[k]); who.push_back(k);
    }
  }
  float ims = 0;
  if (!who.empty()) {
    std::vector<size_t> nn; for (int k : who) nn.push_back(nx[k]);
    ilen.assign(who.size(), 0);
    const int rc = delta_intra_encode(c, (int)who.size(), ip.data(), nn.data(), io.data(), icap.data(), ilen.data(), &ims, launches);
    for (size_t q = 0; q < who.size(); q++) i_len[who[q]] = ilen[q];
    if (rc != CCV2_OK) return rc;
    if (info) for (int k : who) info[k].intra_ms = ims / (float)who.size();
  }
  c->launches = launches; c->device_ms = pms + ims;
  return CCV2_OK;
}

}  // extern "C"

// The prediction half of decodePointCloudDeltaFrame (impl.hpp:1120-1203): the I macroblock tree, the chunk walk, the
// predicted points into pts_out (host or device).  One host wait (the counts).
static int delta_apply(ccv2_codec *c, const void *icloud, size_t ni, const void *p_in, size_t p_len, void *pts_out, size_t cap_points, size_t *npts, uint64_t *decoded_blocks, uint64_t &launches) {
  *npts = 0; if (decoded_blocks) *decoded_blocks = 0;
  const ccv2_params &prm = c->prm;
  const double mres = prm.octree_resolution * prm.macroblock_size;
  const bool di = !ni || is_device_ptr(icloud), dps = !p_len || is_device_ptr(p_in), dout = !cap_points || is_device_ptr(pts_out);
  const size_t ncI = std::max<size_t>(ni, 1);
  const uint32_t extra = prm.do_icp_color_offset ? 3 : 0;
  const size_t chunk_cap = p_len / (7 + extra) + 1;
  const size_t g_i = carve_grid(nullptr, ncI, nullptr, nullptr);
  struct Lay { EncFrame *rec; uint32_t *totals, *c_off; uint8_t *dI, *dps, *g; PChunk *chunks; uint8_t *stage; } L;
  auto carve = [&](void *base) {
    Carver cv(base);
    L.rec = cv.take<EncFrame>(1); L.totals = cv.take<uint32_t>(16);
    L.dI = di ? nullptr : cv.take<uint8_t>(32 * ncI); L.dps = dps ? nullptr : cv.take<uint8_t>(p_len + 16);
    L.g = cv.take<uint8_t>(g_i); L.chunks = cv.take<PChunk>(chunk_cap + 1); L.c_off = cv.take<uint32_t>(chunk_cap + 2);
    L.stage = dout ? nullptr : cv.take<uint8_t>(32 * ncI);      // predicted points are I points moved: at most ni... per chunk; see the capacity check below
    return cv.end();
  };
  const size_t total = carve(nullptr);
  CUI(c->inter_ws.ensure(total + 256));
  carve(c->inter_ws.p);
  const uint8_t *I = di ? (const uint8_t *)icloud : L.dI;
  EncFrame h; memset(&h, 0, sizeof h);
  carve_grid(L.g, ncI, &h, nullptr);
  h.pts = I; h.n = (uint32_t)ni; h.n_finite = (uint32_t)ni; h.violator = NONE_U32;
  int rc;
  if ((rc = define_unit_box(h, mres)) != CCV2_OK) { c->err = "octree depth > 21"; return rc; }
  cudaStream_t st = c->fin_stream;
  if (!di && ni) CUI(cudaMemcpyAsync(L.dI, icloud, 32 * ni, cudaMemcpyHostToDevice, st));
  if (!dps && p_len) CUI(cudaMemcpyAsync(L.dps, p_in, p_len, cudaMemcpyHostToDevice, st));
  CUI(cudaMemcpyAsync(L.rec, &h, sizeof h, cudaMemcpyHostToDevice, st));
  CUI(cudaMemsetAsync(L.totals, 0, 64, st));
  launch_grid(st, L.rec, grid_params(mres, false, false), ni, launches);
  DeltaDecCtx X; memset(&X, 0, sizeof X);
  X.gi = L.rec; X.I = I; X.p_stream = dps ? (const uint8_t *)p_in : L.dps; X.p_len = (uint32_t)p_len;
  X.chunks = L.chunks; X.chunk_cap = (uint32_t)chunk_cap; X.c_off = L.c_off; X.totals = L.totals;
  // a host destination is staged; a chunk list may name a block twice, so the staging area holds at most ncI points and a
  // larger prediction is redone below into a buffer of the right size
  X.out = dout ? (uint8_t *)pts_out : L.stage; X.out_cap = (uint32_t)std::min<size_t>(dout ? cap_points : std::min(cap_points, ncI), 0xFFFFFFFFu);
  X.color_offset = extra != 0;
  pchunk_walk_kernel<<<1, 256, 0, st>>>(X);
  pchunk_prepare_kernel<<<(unsigned)((chunk_cap + 127) / 128), 128, 0, st>>>(X);
  pchunk_scan_kernel<<<1, 1024, 0, st>>>(X);
  pchunk_apply_kernel<<<c->n_sm * 8, 128, 0, st>>>(X);
  launches += 4;
  CUI(cudaGetLastError());
  uint32_t tot[16];
  CUI(cudaMemcpyAsync(tot, L.totals, 64, cudaMemcpyDeviceToHost, st));
  CUI(cudaMemcpyAsync(&h, L.rec, sizeof h, cudaMemcpyDeviceToHost, st));
  CUI(cudaStreamSynchronize(st));
  if ((rc = grid_error(c, h, "I macroblock tree")) != CCV2_OK) return rc;
  const size_t npred = tot[1];
  if (decoded_blocks) *decoded_blocks = tot[2];
  *npts = npred;
  if (npred > cap_points) { c->err = "point buffer too small"; return CCV2_ERR_CAPACITY; }
  if (!dout && npred) {
    if (npred > ncI) {                                       // repeated blocks: stage again with room for all of them
      DevBuf big; CUI(big.ensure(32 * npred));
      X.out = (uint8_t *)big.p; X.out_cap = (uint32_t)npred;
      pchunk_apply_kernel<<<c->n_sm * 8, 128, 0, st>>>(X);
      cudaError_t e = cudaMemcpyAsync(pts_out, big.p, 32 * npred, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      big.release();
      CUI(e);
    } else CUI(cudaMemcpy(pts_out, L.stage, 32 * npred, cudaMemcpyDeviceToHost));
  }
  return CCV2_OK;
}

extern "C" {

int ccv2_decode_delta(ccv2_codec *c, const void *icloud, size_t ni, const void *i_in, size_t i_len, const void *p_in, size_t p_len,
                      void *pts_out, size_t cap_points, size_t *npts, uint64_t *decoded_blocks) {
  const void *ic1 = icloud, *ii1 = i_in, *pi1 = p_in; void *o1 = pts_out;
  return ccv2_decode_delta_batch(c, 1, &ic1, &ni, &ii1, &i_len, &pi1, &p_len, &o1, &cap_points, npts, decoded_blocks);
}

// n delta frames: the predicted macroblocks of every frame first, then the intra-coded rest of ALL frames as one batch through
// the child codec (decodePointCloud of the I streams, appended behind the predicted points, impl.hpp:1226-1230).
int ccv2_decode_delta_batch(ccv2_codec *c, int nframes, const void *const *icloud, const size_t *ni, const void *const *i_in, const size_t *i_len,
                            const void *const *p_in, const size_t *p_len, void *const *pts_out, const size_t *cap_points, size_t *npts, uint64_t *decoded_blocks) {
  if (!c || nframes < 0 || (nframes && (!icloud || !ni || !i_in || !i_len || !p_in || !p_len || !pts_out || !cap_points || !npts))) return CCV2_ERR_ARG;
  for (int k = 0; k < nframes; k++)
    if ((ni[k] && !icloud[k]) || (i_len[k] && !i_in[k]) || (p_len[k] && !p_in[k]) || ni[k] >= (1u << 28) || p_len[k] >= (1ull << 32) || (cap_points[k] && !pts_out[k])) return CCV2_ERR_ARG;
  CUI(cudaSetDevice(c->device));
  c->err.clear();                                            // no finish_all: the delta path shares no workspace with submitted intra calls, which may stay in flight
  uint64_t launches = 0;
  std::vector<const void *> ii; std::vector<size_t> il, cap, nin; std::vector<void *> oo; std::vector<int> who;
  for (int k = 0; k < nframes; k++) {
    const int rc = delta_apply(c, icloud[k], ni[k], p_in[k], p_len[k], pts_out[k], cap_points[k], &npts[k], decoded_blocks ? &decoded_blocks[k] : nullptr, launches);
    if (rc != CCV2_OK) return rc;
    if (i_len[k]) { ii.push_back(i_in[k]); il.push_back(i_len[k]); oo.push_back((uint8_t *)pts_out[k] + 32 * npts[k]); cap.push_back(cap_points[k] - npts[k]); who.push_back(k); }
  }
  if (!who.empty()) {
    ccv2_codec *ch = nullptr;
    int rc = get_intra_child(c, &ch);
    if (rc != CCV2_OK) return rc;
    nin.assign(who.size(), 0);
    rc = ccv2_decode_batch(ch, (int)who.size(), ii.data(), il.data(), oo.data(), cap.data(), nin.data());
    for (size_t q = 0; q < who.size(); q++) npts[who[q]] += nin[q];
    if (rc != CCV2_OK) { c->err = std::string("intra coder of the delta frame: ") + ccv2_last_error(ch); return rc; }
    launches += ccv2_last_launch_count(ch);
  }
  c->launches = launches;
  return CCV2_OK;
}

}  // extern "C"

/*
 * ccv2.h -- C ABI of the B200-native cloud_codec_v2 intra encode/decode hot path (libccv2.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point names the
 * reference interface it replaces (paths relative to the cwi-dis/cwi-pcl-codec tree):
 *   codec.h  = cloud_codec_v2/include/pcl/cloud_codec_v2/point_cloud_codec_v2.h
 *   impl.hpp = cloud_codec_v2/include/pcl/cloud_codec_v2/impl/point_cloud_codec_v2_impl.hpp
 *   eval.hpp = apps/evaluate_compression/include/pcl/apps/evaluate_compression/impl/evaluate_compression_impl.hpp
 *
 * Points are PCL's 32-byte PointXYZRGB records: x,y,z float32 at 0/4/8, data[3] at 12, b,g,r,a uint8 at
 * 16..19, padding to 32 (the struct the reference instantiates, cloud_codec_v2/src/point_cloud_codec_v2.cpp:45).
 * Every data pointer may be host (pageable or pinned) or device memory of the codec's device; the library
 * detects which.  DEVICE cloud buffers (input and output) must be 16-byte aligned (kernels move records as 16-byte
 * words); host buffers are staged by the copy engines and need no alignment, nor do streams.  All functions return 0 on success or a negative ccv2_status.  No exceptions cross the ABI.
 * A codec handle owns its CUDA streams and workspaces and is not thread-safe (like the reference object).
 * There is no CPU fallback: without a CUDA device ccv2_create fails with CCV2_ERR_CUDA.
 *
 * Memory kinds.  Device pointers are read and written in place.  Pinned host memory (ccv2_host_alloc / cudaMallocHost /
 * cudaHostRegister) takes the fast path: clouds go up and decoded clouds come down through the copy engines, streams
 * leave by zero-copy stores, nothing waits on the host inside a call.  Pageable host memory works everywhere but is
 * staged (uploads block the calling thread, results are copied when the call is collected).
 * A decoded-cloud buffer in pinned memory receives ONE transfer of min(capacity, points that can come out) records:
 * records beyond the reported count are unspecified.
 *
 * Process-wide CUDA settings are the host's business: the library keeps up to 28 streams busy and runs best with
 * CUDA_DEVICE_MAX_CONNECTIONS=32 exported before the process creates its CUDA context (with the default of 8 the
 * streams share hardware queues and the long serial kernels serialise, measured 2x on a round trip); it does not touch
 * the environment itself.
 *
 * Limits: frames below 2^28 points; realised octree depth <= 21; SNAKE colour images at most 65500 rows (V < 16.7 M
 * voxels: libjpeg's own limit, CCV2_ERR_UNSUPPORTED beyond); LINES colour mode stages at most 16 KiB of entropy-coded
 * bits per line (the last line may hold 4095 voxels: ~4 bytes per voxel, CCV2_ERR_WORKSPACE beyond).
 */
#ifndef CCV2_H
#define CCV2_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum ccv2_status {
  CCV2_OK = 0,
  CCV2_ERR_ARG = -1,          /* bad argument */
  CCV2_ERR_CUDA = -2,         /* CUDA runtime error / no device (see ccv2_last_error) */
  CCV2_ERR_UNSUPPORTED = -3,  /* outside the implemented scope: profiles != MANUAL; decoding a detail-mode frame with a JPEG colour type (undefined in the reference) */
  CCV2_ERR_CAPACITY = -4,     /* caller's output buffer too small */
  CCV2_ERR_WORKSPACE = -5,    /* internal workspace bound exceeded (tree bytes / stream arena) */
  CCV2_ERR_STREAM = -6,       /* malformed compressed stream */
  CCV2_ERR_DEPTH = -7         /* realised octree depth > 21 (Morton code does not fit 63 bits) */
} ccv2_status;

/* The constructor surface of OctreePointCloudCodecV2 (codec.h:108-143), argument for argument, plus the two
 * setters evaluate_compression calls right after construction (eval.hpp:415-417). */
typedef struct ccv2_params {
  int32_t profile;                 /* compression_Profiles_e; only MANUAL_CONFIGURATION (= 13) is implemented */
  int32_t show_statistics;         /* showStatistics_arg (accepted, ignored: PCL_INFO printing is not reproduced) */
  double point_resolution;         /* pointResolution_arg   (eval.hpp:381: 2^-(octree_bits+enh_bits)) */
  double octree_resolution;        /* octreeResolution_arg  (eval.hpp:383: 2^-octree_bits) */
  int32_t do_voxel_grid_downsampling; /* doVoxelGridDownDownSampling_arg (eval.hpp:385 passes true); false = detail mode, the class default: per-point
                                         residuals at point_resolution and colour differences (impl.hpp:1525-1541, 1728-1757) */
  uint32_t i_frame_rate;           /* iFrameRate_arg (eval.hpp:386 passes 0: every frame is an I frame) */
  int32_t do_color_encoding;       /* doColorEncoding_arg */
  uint8_t color_bit_resolution;    /* colorBitResolution_arg */
  uint8_t color_coding_type;       /* colorCodingType_arg: 0 PCL average, 1 JPEG snake, 2 JPEG lines, 3 raw grid */
  uint8_t _pad0[2];
  int32_t do_voxel_grid_centroid;  /* doVoxelGridCentroid_arg (keep_centroid) */
  int32_t create_scalable_stream;  /* createScalableStream_arg (header only) */
  int32_t code_connectivity;       /* codeConnectivity_arg (header only) */
  int32_t jpeg_quality;            /* jpeg_quality_arg */
  int32_t num_threads;             /* num_threads_arg (only used by the inter-frame predictor; ignored) */
  int32_t macroblock_size;         /* setMacroblockSize (codec.h:149-152); header field, default 16 */
  int32_t do_icp_color_offset;     /* setDoICPColorOffset (codec.h:164-167); header field, default 0 */
} ccv2_params;

#define CCV2_MANUAL_CONFIGURATION 13   /* pcl::io::MANUAL_CONFIGURATION (12 profiles, COMPRESSION_PROFILE_COUNT, then MANUAL) */

typedef struct ccv2_codec ccv2_codec;

/* Fills the values evaluate_compression uses with parameter_config.txt (octree_bits 11, colour type 1, jpeg 85). */
void ccv2_default_params(ccv2_params *p);

/* Replaces: OctreePointCloudCodecV2 constructor (codec.h:108-143). device = CUDA ordinal. */
int ccv2_create(const ccv2_params *p, int device, ccv2_codec **out);
/* Replaces: ~OctreePointCloudCodecV2 (codec.h:146). */
void ccv2_destroy(ccv2_codec *c);

/* Upper bound of the compressed size of a frame of npts points (for sizing `out` buffers). */
size_t ccv2_max_compressed_size(size_t npts);

/* Replaces: encodePointCloud(const PointCloudConstPtr&, std::ostream&) (codec.h:174-175, impl.hpp:80-213) for
 * `nframes` independent frames in one call (the throughput path; nframes = 1 is the reference's call).
 * pts[i]: npts[i] x 32-byte records.  out[i]: buffer of out_cap[i] bytes; out_len[i] receives the stream
 * length, 0 for an empty / all-non-finite cloud (the reference writes nothing, impl.hpp:206-212).
 * Frame ids continue the codec's counter exactly as repeated encodePointCloud calls would (impl.hpp:133). */
int ccv2_encode_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                      void *const *out, const size_t *out_cap, size_t *out_len);

/* Replaces: decodePointCloud(std::istream&, PointCloudPtr&) (codec.h:177-178, impl.hpp:224-310).
 * in[i]: one compressed frame of in_len[i] bytes (one frame per stream, like the reference: impl.hpp:1802-1806).
 * pts_out[i]: buffer for pts_cap[i] points (32 bytes each); npts_out[i] receives the decoded count. */
int ccv2_decode_batch(ccv2_codec *c, int nframes, const void *const *in, const size_t *in_len,
                      void *const *pts_out, const size_t *pts_cap, size_t *npts_out);

/* Encode then decode every frame in one pipelined call -- what evaluate_compression does per frame (eval.hpp:818-843:
 * do_encoding then do_decoding on the stream just produced).  The decoder reads the encoder's device-resident stream,
 * so the host->device copies of later frames overlap the device->host copies of earlier ones.  out may be NULL (or
 * hold NULL entries) when the caller does not want the compressed bytes back; out_len still receives the sizes. */
int ccv2_roundtrip_batch(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                         void *const *out, const size_t *out_cap, size_t *out_len,
                         void *const *pts_out, const size_t *pts_cap, size_t *npts_out);

/* Asynchronous forms of the three calls above: enqueue the whole batch and return a ticket; ccv2_wait(ticket) blocks until
 * the results are in the caller's buffers and returns the call's status.  Three calls may be in flight on one handle (a fourth
 * submit collects the oldest first): they share the codec's workspace rings, so the uploads and parallel kernels of the
 * next call overlap the serial entropy stages of the previous one -- a continuous pipeline over consecutive batches,
 * which the synchronous calls cannot give (each pays the pipeline's fill and drain).  All arrays passed to a submit call
 * (pointer tables, sizes, out_len / npts_out) must stay alive until the ticket has been waited for.  A call that consumes
 * another call's output (decode of streams an in-flight encode is still writing) must wait for that ticket first. */
int ccv2_submit_encode(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                       void *const *out, const size_t *out_cap, size_t *out_len, int *ticket);
int ccv2_submit_decode(ccv2_codec *c, int nframes, const void *const *in, const size_t *in_len,
                       void *const *pts_out, const size_t *pts_cap, size_t *npts_out, int *ticket);
int ccv2_submit_roundtrip(ccv2_codec *c, int nframes, const void *const *pts, const size_t *npts,
                          void *const *out, const size_t *out_cap, size_t *out_len,
                          void *const *pts_out, const size_t *pts_cap, size_t *npts_out, int *ticket);
int ccv2_wait(ccv2_codec *c, int ticket);

/* Device-side stopwatch across calls (bench.py): start collects everything in flight and marks the codec's control stream;
 * stop collects again and returns the milliseconds between the mark and the completion of the last call (CUDA events). */
int ccv2_timer_start(ccv2_codec *c);
int ccv2_timer_stop(ccv2_codec *c, float *ms);

/* Reads point_count from a frame header in HOST memory (SURVEY App. A offset 55) so a caller can size
 * pts_out before decoding (scans for the magic like syncToHeader; a count of 2^28 or more -- the codec's frame limit --
 * is reported as CCV2_ERR_STREAM, so a forged header cannot size an allocation).  Replaces nothing in the reference
 * (its decoder grows a std::vector). */
int ccv2_peek_point_count(const void *in_host, size_t len, uint64_t *npts);

/* Replaces: getPerformanceMetrics() (codec.h:193-197): coded bytes of {octree, centroid, colour} layers of
 * the LAST frame encoded or decoded (impl.hpp:1697,1710,1723). */
int ccv2_get_metrics(ccv2_codec *c, uint64_t m[3]);

/* frame_ID_ accessors (state that reaches the bitstream, impl.hpp:133). */
int ccv2_set_frame_id(ccv2_codec *c, uint32_t next_minus_one);
uint32_t ccv2_get_frame_id(const ccv2_codec *c);

/* Number of kernel launches issued by the last encode/decode batch call (bench.py's gpu_launches). */
uint64_t ccv2_last_launch_count(const ccv2_codec *c);
/* Device time of the batch call collected last (CUDA events on the codec's control stream, submit to completion), milliseconds. */
float ccv2_last_device_ms(const ccv2_codec *c);

/* Profiling hook used by bench.py's roofline leg: when on, a batch call runs all its frames as one group on ONE
 * stream and the library brackets every kernel launch with CUDA events on that stream; afterwards
 * ccv2_get_profile(idx) returns per kernel name the summed device time and launch count of the last call
 * (CCV2_ERR_ARG once idx runs past the last kernel). */
int ccv2_set_profiling(ccv2_codec *c, int on);
int ccv2_get_profile(const ccv2_codec *c, int idx, const char **name, float *total_ms, int *launches);

/* Text of the last error on this codec (or of ccv2_create when c == NULL). */
const char *ccv2_last_error(const ccv2_codec *c);
const char *ccv2_status_string(int status);

/* [PCL] OctreePointCloudCompression::getOutputCloud() as evaluate_compression uses it after encodePointCloud
 * (eval.hpp:862): the encoder's simplified cloud output_ (impl.hpp:96, filled at impl.hpp:1549-1576) -- one 32-byte
 * PointXYZRGB per occupied voxel in stream (DFS) order, at the voxel centre `corner + 0.5 * resolution` (or the float
 * centroid of the voxel's points when doVoxelGridCentroid is set), carrying the voxel's average colour before JPEG.
 * Valid for frame `frame` of the LAST ccv2_encode_batch call, until the next call on the handle and only while the frame's
 * workspace has not been handed on to a later group of the same call (always true for the last 128 frames of a call; centroid mode
 * re-reads the input cloud, which must still be alive if it was passed as a device pointer).  points_out: host or
 * device memory for cap_points records; *npoints receives the voxel count (also when cap_points is too small:
 * CCV2_ERR_CAPACITY). */
int ccv2_get_output_cloud(ccv2_codec *c, int frame, void *points_out, size_t cap_points, size_t *npoints);

/* Tile mode (BASELINE configs[3]; no counterpart in the reference, which has no parallelism: CMakeLists.txt:85-87).  One
 * large frame is cut into 2^tile_bits (8 or 64) spatial tiles of the unit cube evaluate_compression normalises clouds
 * into (impl.hpp:1915-1945): tile = Morton index, x most significant, of floor(p * 2^(tile_bits/3)) per axis, clamped.
 * Every tile keeps its points in their original relative order and is encoded as an ordinary frame, so each tile stream
 * is what encodePointCloud (codec.h:174-175) writes for that subset and any reference decoder reads it; the union of the
 * decoded tiles is the decoded frame.  The tiles' serial entropy stages run side by side: that is what shortens the
 * latency of a single frame, on one GPU or over the ranks of a node (first_tile = rank, tile_step = world size).
 * ccv2_split_tiles: stable partition only (pts_out: n records, host or device; tile_offsets: 2^tile_bits + 1 entries).
 * ccv2_encode_tiles: partition + encode; out / out_cap / out_len (and tile_npts, may be NULL) are indexed by tile. */
int ccv2_split_tiles(ccv2_codec *c, const void *pts, size_t n, int tile_bits, void *pts_out, size_t *tile_offsets);
int ccv2_encode_tiles(ccv2_codec *c, const void *pts, size_t n, int tile_bits, int first_tile, int tile_step,
                      void *const *out, const size_t *out_cap, size_t *out_len, size_t *tile_npts);

/* Replaces: computeQualityMetric(cloud_a, cloud_b, QualityMetric&) of evaluate_compression
 * (apps/evaluate_compression/include/pcl/apps/evaluate_compression/impl/quality_metrics_impl.hpp:82-239; struct
 * quality_metrics.h:53-75): cloud_a = original, cloud_b = decoded, both 32-byte PointXYZRGB records in host or device
 * memory.  Nearest neighbours are exact (exhaustive search on the GPU instead of two kd-trees); geometry is symmetric,
 * colour PSNR is A -> B on a 0..1 YUV scale like the reference's.  Non-finite points take no part. */
typedef struct ccv2_quality {
  uint64_t in_point_count, out_point_count;
  float symm_rms, symm_hausdorff, left_hausdorff, right_hausdorff, left_rms, right_rms;
  double psnr_db;
  double psnr_yuv[3];
} ccv2_quality;
int ccv2_quality_metrics(ccv2_codec *c, const void *cloud_a, size_t na, const void *cloud_b, size_t nb, ccv2_quality *out);

/* ---- Inter-frame (predictive) coding: BASELINE configs[2], SURVEY 8 rows a14 / f-1 ---------------------------------------
 * Replaces: encodePointCloudDeltaFrame / decodePointCloudDeltaFrame (codec.h:180-190, impl.hpp:787-1112, 1120-1235) with
 * their helpers simplifyPCloud (impl.hpp:318-400), generate_macroblock_tree (:410-431), do_icp_prediction (:443-568) and
 * the RigidTransformCoding / QuaternionCoding classes (rigid_transform_coding_impl.hpp:63-203,
 * quaternion_coding_impl.hpp:55-222).  icloud = the frame predicted FROM (evaluate_compression passes the encoder's
 * simplified cloud of the previous frame, eval.hpp:862 -> ccv2_get_output_cloud; the decoder passes its decoded previous
 * frame), pcloud = the frame to code.  Both streams are the reference's: the P stream is the chunk list
 * [u8 size][3 x i16 macroblock key][6 | 10 x i16 transform][3 x i8 colour offsets] (impl.hpp:877-883), the I stream an
 * ordinary intra frame of the points no macroblock predicted, written by a fresh codec with the reference's ten explicit
 * constructor arguments (impl.hpp:1089-1101: scalable stream on, JPEG quality 75 whatever this codec uses).
 * The registration follows PCL 1.10's IterativeClosestPoint as the reference configures it (50 iterations, transformation
 * epsilon 1e-8f, fitness epsilon 3e-8f, accepted when the fitness is below 2 x point_resolution) with the arithmetic
 * oracle/ccv2_oracle_inter.c states; no build of the reference reproduces another build's ICP bits, so parity for this path
 * is: bit-exact against the oracle, format + quality against the reference.
 * The delta calls are synchronous, but they share no workspace with ccv2_submit_* calls of the same handle, which may stay in
 * flight across them (bench.py --mode inter decodes a group's I frames while its P frames are being predicted).
 * Limits: pcloud below 2^27 points, icloud below 2^28.
 * All cloud and stream pointers may be host or device memory.  out_cloud (may be NULL) receives the predicted frame the
 * reference writes when write_out_cloud is set.  Macroblock size and colour offsets come from ccv2_params. */
typedef struct ccv2_delta_info {
  uint64_t macro_blocks, shared_blocks, converged_blocks;   /* macro_block_count, shared_macroblock_count, convergence_count (impl.hpp:803-805) */
  uint64_t n_intra_points, n_p_points;                      /* points coded intra; points of the (simplified) P cloud */
  float shared_percentage, convergence_percentage;          /* getMacroBlockPercentage / getMacroBlockConvergencePercentage (codec.h:200-210) */
  float predict_ms, intra_ms;                               /* device time of the prediction stage / of the intra coder's call */
} ccv2_delta_info;
int ccv2_encode_delta(ccv2_codec *c, const void *icloud, size_t ni, const void *pcloud, size_t np, int icp_on_original,
                      void *i_out, size_t i_cap, size_t *i_len, void *p_out, size_t p_cap, size_t *p_len,
                      void *out_cloud, size_t out_cap_points, size_t *n_out, ccv2_delta_info *info);
/* pts_out: cap_points records; the predicted macroblocks come first (chunk order), then the intra-coded points.
 * decoded_blocks (may be NULL): macroblocks that found their I block. */
int ccv2_decode_delta(ccv2_codec *c, const void *icloud, size_t ni, const void *i_in, size_t i_len, const void *p_in, size_t p_len,
                      void *pts_out, size_t cap_points, size_t *npts, uint64_t *decoded_blocks);
/* Several delta frames in one call (frame k: pcloud[k] against icloud[k]).  The prediction stages run one after the other;
 * the intra-coded parts of ALL frames go through the intra coder as ONE pipelined batch -- its serial range-coder stage is
 * latency bound, so a batch of 29 frames takes about as long as one.  Same streams as frame-by-frame calls. */
int ccv2_encode_delta_batch(ccv2_codec *c, int nframes, const void *const *icloud, const size_t *ni, const void *const *pcloud, const size_t *np,
                            int icp_on_original, void *const *i_out, const size_t *i_cap, size_t *i_len,
                            void *const *p_out, const size_t *p_cap, size_t *p_len, ccv2_delta_info *info /* nframes entries or NULL */);
int ccv2_decode_delta_batch(ccv2_codec *c, int nframes, const void *const *icloud, const size_t *ni, const void *const *i_in, const size_t *i_len,
                            const void *const *p_in, const size_t *p_len, void *const *pts_out, const size_t *cap_points, size_t *npts,
                            uint64_t *decoded_blocks /* nframes entries or NULL */);
/* simplifyPCloud alone (impl.hpp:318-400): one point per occupied voxel of the unit-box octree, DFS order. */
int ccv2_simplify(ccv2_codec *c, const void *pts, size_t n, void *pts_out, size_t cap_points, size_t *npts);
/* Upper bound of a P stream for a P cloud of np points. */
size_t ccv2_max_p_stream_size(size_t np);

/* Pinned host memory helpers (cudaMallocHost / cudaFreeHost) for callers that want full-speed PCIe copies. */
void *ccv2_host_alloc(size_t bytes);
void ccv2_host_free(void *p);

/* Test hook: copies an intermediate of frame `frame` of the last encode batch to host.
 * what: 0 leaf Morton codes (u64 x V), 1 tree bytes (B), 2 average colours (3V), 3 colour payload (J),
 *       4 sorted point indices (u32 x n_finite), 5 frame info (ccv2_frame_info). */
typedef struct ccv2_frame_info {
  uint32_t depth, n_finite, n_leaves, n_tree_bytes, n_color_bytes, error;
  double bb_min[3], bb_max[3];
  uint64_t coded[3];
} ccv2_frame_info;
int ccv2_debug_fetch(ccv2_codec *c, int frame, int what, void *host_buf, size_t cap, size_t *len);

#ifdef __cplusplus
}
#endif
#endif

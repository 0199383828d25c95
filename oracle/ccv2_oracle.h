/*
 * ccv2_oracle.h -- CPU oracle for the cloud_codec_v2 intra encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker / the CPU side of a comparison.
 *
 * What it is: a dependency-free C restatement of the algorithm the reference
 * (cwi-dis/cwi-pcl-codec, /root/reference) runs for
 *   OctreePointCloudCodecV2<PointXYZRGB>::encodePointCloud / decodePointCloud
 *   (cloud_codec_v2/include/pcl/cloud_codec_v2/impl/point_cloud_codec_v2_impl.hpp:80-310)
 * including the arithmetic the reference inherits from two un-vendored third
 * parties: PCL 1.8-1.10 (octree bbox growth / key generation / serializeTree /
 * StaticRangeCoder / ColorCoding) and libjpeg-turbo (baseline JPEG, via jpeg_io).
 *
 * PARITY STATUS
 *   - JPEG encode bytes / decode pixels: PINNED against libjpeg-turbo through the
 *     golden vectors in tests/golden/ (generated with Pillow by
 *     tests/golden/make_jpeg_golden.py).
 *   - PCL-inherited parts (bbox growth, keys, DFS order, range coder, header):
 *     PARITY UNPINNED -- the reference ships no tests/golden vectors and PCL is
 *     not available offline, so these follow SURVEY.md Appendix B only.  Frozen
 *     inputs + SHA-256 of the oracle's streams live in tests/golden/ so a later
 *     run of real PCL can be diffed in one command.
 *   - The reference itself cannot be compiled here (needs PCL, Boost, Eigen,
 *     jpeglib.h) => no oracle/_ref.
 */
#ifndef CCV2_ORACLE_H
#define CCV2_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ctor surface of OctreePointCloudCodecV2 (point_cloud_codec_v2.h:108-143), MANUAL_CONFIGURATION */
typedef struct orc_params {
  double point_resolution;      /* pointResolution_arg  */
  double octree_resolution;     /* octreeResolution_arg */
  int do_voxel_grid;            /* doVoxelGridDownDownSampling_arg; 0 = detail mode (per-point residuals, impl.hpp:1525-1541, 1728-1757) */
  int do_color;                 /* doColorEncoding_arg */
  int color_bit_resolution;     /* colorBitResolution_arg */
  int color_coding_type;        /* 0 = PCL avg (bit-reduced), 1 = SNAKE jpeg, 2 = LINES jpeg, 3 = GRID (raw) */
  int do_centroid;              /* doVoxelGridCentroid_arg */
  int create_scalable;          /* createScalableStream_arg */
  int code_connectivity;        /* codeConnectivity_arg */
  int jpeg_quality;             /* jpeg_quality_arg */
  int macroblock_size;          /* header only (16) */
  int do_icp_color_offset;      /* header only (0) */
} orc_params;

typedef struct orc_info {
  uint32_t depth;               /* realised octree depth */
  double bb_min[3], bb_max[3];
  uint64_t n_finite;            /* points that passed isFinite */
  uint64_t n_leaves;            /* V */
  uint64_t n_tree_bytes;        /* B */
  uint64_t n_color_bytes;       /* bytes handed to the range coder for colour (J for jpeg modes) */
  uint64_t coded[3];            /* compression_performance_metrics (impl.hpp:1697,1710,1723) */
  double t_ms[8];               /* stage timings of the oracle run (bbox/keys, sort, serialise, colour, jpeg, entropy, ...) */
} orc_info;

/* intermediates, for stage-by-stage parity tests; every pointer is malloc'd by the oracle, free with orc_free */
typedef struct orc_debug {
  uint64_t *leaf_keys;          /* V Morton codes (x = MSB of each triple), ascending */
  uint8_t *tree_bytes;          /* B */
  uint8_t *avg_colors;          /* 3V bytes, B,G,R per leaf, before jpeg */
  uint8_t *color_payload;       /* n_color_bytes */
  uint8_t *centroid_bytes;      /* 3V or NULL */
  uint8_t *output_cloud;        /* V x 32 bytes: the encoder's simplified cloud output_ (impl.hpp:96,1549-1576; [PCL] getOutputCloud, eval.hpp:862) */
} orc_debug;

void orc_default_params(orc_params *p);

/* encodePointCloud (impl.hpp:80-213). pts: n x 32-byte PointXYZRGB (x,y,z f32 @0,4,8; b,g,r,a @16..19).
 * frame_id is the value written to the header (the codec pre-increments, first frame = 1).
 * Returns 0, or <0 on error; an empty / all-non-finite cloud yields *out_len = 0 (impl.hpp:206-212). */
int orc_encode(const orc_params *p, uint32_t frame_id, const void *pts, size_t n,
               uint8_t **out, size_t *out_len, orc_info *info, orc_debug *dbg);

/* decodePointCloud (impl.hpp:224-310). Output: *n x 32-byte points, malloc'd. */
int orc_decode(const uint8_t *in, size_t len, void **pts, size_t *n, orc_info *info);

void orc_free(void *p);
void orc_free_debug(orc_debug *d);

/* ---- building blocks, exported for unit tests ---- */
/* PCL StaticRangeCoder::encodeCharVectorToStream: out gets 1028 + k + 8 bytes */
int orc_range_encode(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
/* decodeStreamToCharVector: consumes exactly *consumed bytes of in */
int orc_range_decode(const uint8_t *in, size_t in_len, uint8_t *out, size_t n, size_t *consumed);
/* PCL StaticRangeCoder::encodeIntVectorToStream / decodeStreamToIntVector (64-bit coder; detail-mode point counts) */
int orc_range_encode_int(const uint32_t *in, size_t n, uint8_t **out, size_t *out_len);
int orc_range_decode_int(const uint8_t *in, size_t in_len, uint32_t *out, size_t n, size_t *consumed);
/* libjpeg baseline encoder as jpeg_io drives it (jpeg_io.hpp:259-311); rgb interleaved 3 bytes/pixel */
int orc_jpeg_encode(const uint8_t *rgb, int w, int h, int quality, uint8_t **out, size_t *out_len);
/* libjpeg decoder defaults (jpeg_io.hpp:140-162): ISLOW + fancy upsampling */
int orc_jpeg_decode(const uint8_t *in, size_t len, uint8_t **rgb, int *w, int *h);
/* snake_grid_mapping.h:46-71 (literal iterator) and SURVEY App. B.7 (closed form): position of linear index i */
void orc_snake_positions_literal(int w, int h, int32_t *pos /* w*h */);
int32_t orc_snake_pos_closed(int w, int h, int64_t i);
/* sequential bbox growth + keys. keys_xyz: 3*n u32 (undefined for non-finite), finite: n bytes */
int orc_bbox_keys(const void *pts, size_t n, double resolution, double bb_min[3], double bb_max[3],
                  uint32_t *depth, uint32_t *keys_xyz, uint8_t *finite);
/* recursive DFS reference of Octree2BufBase::serializeTreeRecursive over distinct Morton leaf codes
 * (used to cross-check the sort-based emission) */
int orc_dfs_recursive(const uint64_t *leaf_codes, size_t v, uint32_t depth, uint8_t **bytes, size_t *nbytes);

/* computeQualityMetric (quality_metrics_impl.hpp:82-239), exhaustive nearest neighbours: test sizes only */
typedef struct orc_quality {
  uint64_t in_point_count, out_point_count;
  float symm_rms, symm_hausdorff, left_hausdorff, right_hausdorff, left_rms, right_rms;
  double psnr_db;
  double psnr_yuv[3];
} orc_quality;
int orc_quality_metrics(const void *cloud_a, size_t na, const void *cloud_b, size_t nb, orc_quality *out);


/* ---- inter-frame (predictive) path: oracle/ccv2_oracle_inter.c -- PARITY UNPINNED (see that file's header) ---- */
typedef struct orc_delta_info {
  uint64_t macro_blocks, shared_blocks, converged_blocks;   /* macro_block_count, shared_macroblock_count, convergence_count (impl.hpp:803-805) */
  uint64_t n_intra_points, n_p_points;                      /* points coded intra; points of the (simplified) P cloud */
  float shared_percentage, convergence_percentage;          /* getMacroBlockPercentage / getMacroBlockConvergencePercentage (impl.hpp:1105-1106) */
} orc_delta_info;
/* simplifyPCloud (impl.hpp:318-400): one point per occupied voxel of the unit-box octree, DFS order */
int orc_simplify(const orc_params *p, const void *pts, size_t n, void **out, size_t *nout);
/* encodePointCloudDeltaFrame (impl.hpp:787-1112).  out_cloud may be NULL (write_out_cloud = false, as eval.hpp:506 passes). */
int orc_encode_delta(const orc_params *p, const void *icloud, size_t ni, const void *pcloud, size_t np, int icp_on_original,
                     uint8_t **i_out, size_t *i_len, uint8_t **p_out, size_t *p_len, void **out_cloud, size_t *n_out_cloud, orc_delta_info *info);
/* decodePointCloudDeltaFrame (impl.hpp:1120-1235) */
int orc_decode_delta(const orc_params *p, const void *icloud, size_t ni, const uint8_t *i_in, size_t i_len, const uint8_t *p_in, size_t p_len,
                     void **out, size_t *nout, uint64_t *decoded_blocks);
/* RigidTransformCoding::compressRigidTransform / deCompressRigidTransform (rigid_transform_coding_impl.hpp:63-203); m: row-major 4x4 */
int orc_compress_rigid_transform(const float *m, int16_t *out /* 10 */, int *nwords /* 6 or 10 */);
int orc_decompress_rigid_transform(const int16_t *in, int nwords, float *m);
/* pcl::IterativeClosestPoint as do_icp_prediction drives it (impl.hpp:544-560); packed xyz floats; F: row-major 4x4 */
int orc_icp(const float *src, size_t ns, const float *tgt, size_t nt, int max_iter, double tf_eps, double fit_eps, float *F, int *converged, double *fitness);

#ifdef __cplusplus
}
#endif
#endif

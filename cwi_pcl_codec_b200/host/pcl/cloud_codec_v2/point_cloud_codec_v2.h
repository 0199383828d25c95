// point_cloud_codec_v2.h (B200 facade) -- C++ host side of the drop-in: the reference's class name, constructor
// argument order and method names (cloud_codec_v2/include/pcl/cloud_codec_v2/point_cloud_codec_v2.h:70-197 of
// cwi-dis/cwi-pcl-codec), implemented as a thin forwarder to the C ABI in include/ccv2.h.  Nothing here computes:
// every byte of the stream is produced by the CUDA library.  Only PointXYZRGB is supported, like the reference's one
// explicit instantiation (cloud_codec_v2/src/point_cloud_codec_v2.cpp:45).
#pragma once
#include "../pcl_shim.h"
#include "../../../../include/ccv2.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <unordered_map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace pcl { namespace io {

template <typename PointT> class OctreePointCloudCodecV2;

// pcl::io::BoundingBox of the reference (point_cloud_codec_v2.h:62-66: two Eigen::Vector4f); plain floats here
struct BoundingBox { float min_xyz[4]; float max_xyz[4]; };

template <>
class OctreePointCloudCodecV2<pcl::PointXYZRGB> {
public:
  typedef pcl::PointCloud<pcl::PointXYZRGB> PointCloud;
  typedef PointCloud::Ptr PointCloudPtr;
  typedef PointCloud::ConstPtr PointCloudConstPtr;

  // codec.h:108-143 -- same 14 arguments, same defaults
  OctreePointCloudCodecV2(compression_Profiles_e compressionProfile_arg = MED_RES_ONLINE_COMPRESSION_WITH_COLOR,
                          bool showStatistics_arg = false, const double pointResolution_arg = 0.001,
                          const double octreeResolution_arg = 0.01, bool doVoxelGridDownDownSampling_arg = false,
                          const unsigned int iFrameRate_arg = 0, bool doColorEncoding_arg = true,
                          const unsigned char colorBitResolution_arg = 6, const unsigned char colorCodingType_arg = 0,
                          bool doVoxelGridCentroid_arg = true, bool createScalableStream_arg = true,
                          bool codeConnectivity_arg = false, int jpeg_quality_arg = 75, int num_threads = 0, int device = 0)
      : h_(nullptr) {
    ccv2_default_params(&p_);
    p_.profile = compressionProfile_arg; p_.show_statistics = showStatistics_arg;
    p_.point_resolution = pointResolution_arg; p_.octree_resolution = octreeResolution_arg;
    p_.do_voxel_grid_downsampling = doVoxelGridDownDownSampling_arg; p_.i_frame_rate = iFrameRate_arg;
    p_.do_color_encoding = doColorEncoding_arg; p_.color_bit_resolution = colorBitResolution_arg;
    p_.color_coding_type = colorCodingType_arg; p_.do_voxel_grid_centroid = doVoxelGridCentroid_arg;
    p_.create_scalable_stream = createScalableStream_arg; p_.code_connectivity = codeConnectivity_arg;
    p_.jpeg_quality = jpeg_quality_arg; p_.num_threads = num_threads;
    p_.macroblock_size = 16; p_.do_icp_color_offset = 0;                     // codec.h:138,141
    device_ = device;
    metrics_[0] = metrics_[1] = metrics_[2] = 0;
  }
  ~OctreePointCloudCodecV2() { ccv2_destroy(h_); }
  OctreePointCloudCodecV2(const OctreePointCloudCodecV2 &) = delete;
  OctreePointCloudCodecV2 &operator=(const OctreePointCloudCodecV2 &) = delete;

  // codec.h:149-167: header fields; must be called before the first encode (the handle is created lazily)
  void setMacroblockSize(int size) { p_.macroblock_size = size; }
  void setDoICPColorOffset(bool doit) { p_.do_icp_color_offset = doit; }

  // codec.h:174-175, impl.hpp:80-213.  void like the reference; an empty cloud writes nothing (impl.hpp:206-212).
  void encodePointCloud(const PointCloudConstPtr &cloud_arg, std::ostream &compressed_tree_data_out_arg) {
    ensure();
    const void *pts = cloud_arg->points.data();
    size_t n = cloud_arg->points.size();
    buf_.resize(ccv2_max_compressed_size(n));
    void *o = buf_.data();
    size_t cap = buf_.size(), len = 0;
    if (ccv2_encode_batch(h_, 1, &pts, &n, &o, &cap, &len) != CCV2_OK) { last_error_ = ccv2_last_error(h_); return; }
    compressed_tree_data_out_arg.write(reinterpret_cast<const char *>(buf_.data()), static_cast<std::streamsize>(len));
    ccv2_get_metrics(h_, metrics_);
  }

  // codec.h:177-178, impl.hpp:224-310.  One frame per stream (impl.hpp:1802-1806); a stream without a frame header
  // returns silently (impl.hpp:231).
  void decodePointCloud(std::istream &compressed_tree_data_in_arg, PointCloudPtr &cloud_arg) {
    ensure();
    std::string s((std::istreambuf_iterator<char>(compressed_tree_data_in_arg)), std::istreambuf_iterator<char>());
    uint64_t cnt = 0;
    if (ccv2_peek_point_count(s.data(), s.size(), &cnt) != CCV2_OK) return;
    try { cloud_arg->points.resize(cnt ? cnt : 1); }
    catch (const std::bad_alloc &) { last_error_ = "decodePointCloud: cannot allocate the decoded cloud"; cloud_arg->points.clear(); return; }
    const void *i = s.data();
    size_t il = s.size(), cap = cloud_arg->points.size(), n = 0;
    void *o = cloud_arg->points.data();
    if (ccv2_decode_batch(h_, 1, &i, &il, &o, &cap, &n) != CCV2_OK) { last_error_ = ccv2_last_error(h_); cloud_arg->points.clear(); return; }
    cloud_arg->points.resize(n);
    cloud_arg->width = static_cast<std::uint32_t>(n); cloud_arg->height = 1; cloud_arg->is_dense = false;   // impl.hpp:284-286
    ccv2_get_metrics(h_, metrics_);
  }

  // [PCL] OctreePointCloudCompression::getOutputCloud (eval.hpp:862): the simplified cloud of the last encodePointCloud
  // (impl.hpp:96, 1549-1576), fetched from the device on demand.
  PointCloudPtr getOutputCloud() {
    PointCloudPtr out(new PointCloud());
    if (!h_) return out;
    size_t n = 0;
    int rc = ccv2_get_output_cloud(h_, 0, nullptr, 0, &n);
    if ((rc != CCV2_OK && rc != CCV2_ERR_CAPACITY) || n == 0) return out;
    out->points.resize(n);
    if (ccv2_get_output_cloud(h_, 0, out->points.data(), n, &n) != CCV2_OK) { last_error_ = ccv2_last_error(h_); out->points.clear(); return out; }
    out->width = static_cast<std::uint32_t>(n); out->height = 1;
    return out;
  }

  // codec.h:184-186, impl.hpp:787-1112: pcloud_arg coded against icloud_arg -- the P stream (macroblock chunks) and the I
  // stream (an intra frame of what no macroblock predicted); out_cloud_arg receives the predicted frame when write_out_cloud.
  void encodePointCloudDeltaFrame(const PointCloudConstPtr &icloud_arg, const PointCloudConstPtr &pcloud_arg, PointCloudPtr &out_cloud_arg,
                                  std::ostream &i_coded_data, std::ostream &p_coded_data, bool icp_on_original = false, bool write_out_cloud = false) {
    ensure();
    const size_t ni = icloud_arg->points.size(), np = pcloud_arg->points.size();
    buf_.resize(ccv2_max_compressed_size(np));
    std::vector<unsigned char> pbuf(ccv2_max_p_stream_size(np));
    std::vector<pcl::PointXYZRGB> oc(write_out_cloud ? ni + np + 1 : 0);
    size_t il = 0, pl = 0, no = 0;
    ccv2_delta_info info;
    out_cloud_arg->height = 1; out_cloud_arg->width = 0;                     // impl.hpp:814-815
    if (ccv2_encode_delta(h_, icloud_arg->points.data(), ni, pcloud_arg->points.data(), np, icp_on_original ? 1 : 0, buf_.data(), buf_.size(), &il,
                          pbuf.data(), pbuf.size(), &pl, write_out_cloud ? (void *)oc.data() : nullptr, oc.size(), write_out_cloud ? &no : nullptr, &info) != CCV2_OK) {
      last_error_ = ccv2_last_error(h_); return;
    }
    i_coded_data.write(reinterpret_cast<const char *>(buf_.data()), static_cast<std::streamsize>(il));
    p_coded_data.write(reinterpret_cast<const char *>(pbuf.data()), static_cast<std::streamsize>(pl));
    if (write_out_cloud) { oc.resize(no); out_cloud_arg->points.insert(out_cloud_arg->points.end(), oc.begin(), oc.end()); out_cloud_arg->width = (std::uint32_t)out_cloud_arg->points.size(); }
    shared_macroblock_percentage_ = info.shared_percentage; shared_macroblock_convergence_percentage_ = info.convergence_percentage;
    delta_info_ = info;
  }
  // codec.h:180-182, impl.hpp:577-786: the older form of the delta encoder.  Same prediction; its P stream is the chunk list
  // WITHOUT the size bytes (impl.hpp:650-660), which no decoder of the reference reads -- produced by re-framing the stream.
  void generatePointCloudDeltaFrame(const PointCloudConstPtr &icloud_arg, const PointCloudConstPtr &pcloud_arg, PointCloudPtr &out_cloud_arg,
                                    std::ostream &i_coded_data, std::ostream &p_coded_data, bool icp_on_original = false, bool write_out_cloud = true) {
    std::stringstream p_sized;
    encodePointCloudDeltaFrame(icloud_arg, pcloud_arg, out_cloud_arg, i_coded_data, p_sized, icp_on_original, write_out_cloud);
    const std::string s = p_sized.str();
    for (size_t pos = 0; pos < s.size();) {
      const size_t size = (unsigned char)s[pos];
      if (size == 0 || pos + 1 + size > s.size()) break;
      p_coded_data.write(s.data() + pos + 1, (std::streamsize)size);
      pos += 1 + size;
    }
  }
  // codec.h:188-190, impl.hpp:1120-1235: predicted macroblocks first (chunk order), then the intra-coded rest, appended to out_cloud_arg
  void decodePointCloudDeltaFrame(const PointCloudConstPtr &icloud_arg, PointCloudPtr &out_cloud_arg, std::istream &i_coded_data, std::istream &p_coded_data) {
    ensure();
    const std::string is((std::istreambuf_iterator<char>(i_coded_data)), std::istreambuf_iterator<char>());
    const std::string ps((std::istreambuf_iterator<char>(p_coded_data)), std::istreambuf_iterator<char>());
    uint64_t cnt = 0;
    if (!is.empty() && ccv2_peek_point_count(is.data(), is.size(), &cnt) != CCV2_OK) cnt = 0;
    const size_t ni = icloud_arg->points.size();
    size_t cap = ni + cnt + 1, n = 0;
    std::vector<pcl::PointXYZRGB> out;
    for (int attempt = 0; attempt < 2; attempt++) {                          // a chunk list may name a block more than once: grow once
      out.resize(cap);
      const int rc = ccv2_decode_delta(h_, icloud_arg->points.data(), ni, is.data(), is.size(), ps.data(), ps.size(), out.data(), cap, &n, nullptr);
      if (rc == CCV2_OK) break;
      if (rc == CCV2_ERR_CAPACITY && attempt == 0) { cap = n + cnt + 1; continue; }
      last_error_ = ccv2_last_error(h_); return;
    }
    out.resize(n);
    out_cloud_arg->points.insert(out_cloud_arg->points.end(), out.begin(), out.end());
    out_cloud_arg->width = (std::uint32_t)out_cloud_arg->points.size(); out_cloud_arg->height = 1;
  }
  // codec.h:200-210
  float getMacroBlockPercentage() { return shared_macroblock_percentage_; }
  float getMacroBlockConvergencePercentage() { return shared_macroblock_convergence_percentage_; }
  const ccv2_delta_info &lastDeltaInfo() const { return delta_info_; }

  // codec.h:193-197
  uint64_t *getPerformanceMetrics() { return metrics_; }

  // ---- the three statics evaluate_compression calls around the codec (codec.h:216-227, impl.hpp:1840-1986).  Host-side
  // float arithmetic in the reference's order: they run once per group on the caller's clouds, outside the hot path.
  // impl.hpp:1871-1966: the first frame's bounding box, expanded by bb_expand_factor on every side, is kept until a frame
  // does not fit strictly inside it; every frame is mapped to (p - min) / (max - min) in float; the LAST box is returned.
  static BoundingBox normalize_pointclouds(std::vector<PointCloudPtr> &point_clouds, std::vector<BoundingBox> &bounding_boxes,
                                           double bb_expand_factor, unsigned int debug_level = 0) {
    float mn_bb[3] = {1000.f, 1000.f, 1000.f}, mx_bb[3] = {-1000.f, -1000.f, -1000.f};
    bool is_bb_init = false;
    bounding_boxes.resize(point_clouds.size());
    for (size_t k = 0; k < point_clouds.size(); k++) {
      float mn[3], mx[3];
      min_max_3d(*point_clouds[k], mn, mx);
      if (!((mn[0] > mn_bb[0]) && (mn[1] > mn_bb[1]) && (mn[2] > mn_bb[2]))) is_bb_init = false;
      if (!((mx[0] < mx_bb[0]) && (mx[1] < mx_bb[1]) && (mx[2] < mx_bb[2]))) is_bb_init = false;
      if (!is_bb_init) {
        for (int a = 0; a < 3; a++) {                                       // float - double * float, stored as float (impl.hpp:1917-1923)
          const float ext = std::fabs(mx[a] - mn[a]);
          mn_bb[a] = (float)((double)mn[a] - bb_expand_factor * (double)ext);
          mx_bb[a] = (float)((double)mx[a] + bb_expand_factor * (double)ext);
        }
        is_bb_init = true;
        if (debug_level > 0) std::cout << "re-intialized bounding box !!! " << std::endl;
      }
      float dyn[3];
      for (int a = 0; a < 3; a++) { dyn[a] = mx_bb[a] - mn_bb[a]; bounding_boxes[k].min_xyz[a] = mn_bb[a]; bounding_boxes[k].max_xyz[a] = mx_bb[a]; }
      bounding_boxes[k].min_xyz[3] = bounding_boxes[k].max_xyz[3] = 0.f;
      for (auto &p : point_clouds[k]->points) {                            // impl.hpp:1936-1946: offset, then dynamic range
        p.x -= mn_bb[0]; p.y -= mn_bb[1]; p.z -= mn_bb[2];
        p.x /= dyn[0]; p.y /= dyn[1]; p.z /= dyn[2];
      }
    }
    BoundingBox bb;
    for (int a = 0; a < 3; a++) { bb.min_xyz[a] = mn_bb[a]; bb.max_xyz[a] = mx_bb[a]; }
    bb.min_xyz[3] = 0.f; bb.max_xyz[3] = 0.f;
    return bb;
  }
  // impl.hpp:1968-1986
  static void restore_scaling(PointCloudPtr &point_cloud, const BoundingBox &bb) {
    const float dyn[3] = {bb.max_xyz[0] - bb.min_xyz[0], bb.max_xyz[1] - bb.min_xyz[1], bb.max_xyz[2] - bb.min_xyz[2]};
    for (auto &p : point_cloud->points) {
      p.x *= dyn[0]; p.y *= dyn[1]; p.z *= dyn[2];
      p.x += bb.min_xyz[0]; p.y += bb.min_xyz[1]; p.z += bb.min_xyz[2];
    }
  }
  // impl.hpp:1840-1866 -> [PCL] RadiusOutlierRemoval(setRadiusSearch(radius), setMinNeighborsInRadius(min_points)): a point
  // stays when MORE than min_points points (itself included) lie within `radius`.  Exact, through a uniform grid of
  // radius-sized cells; non-finite points are dropped like PCL's filter does.
  static void remove_outliers(std::vector<PointCloudPtr> &point_clouds, int min_points, double radius, unsigned int debug_level = 0) {
    if (min_points <= 0) return;
    for (auto &pc : point_clouds) {
      const auto &pts = pc->points;
      std::unordered_map<uint64_t, std::vector<uint32_t> > grid;
      auto cell = [&](float v) { return (int64_t)std::floor((double)v / radius); };
      auto key = [](int64_t x, int64_t y, int64_t z) { return ((uint64_t)(x & 0x1FFFFF) << 42) | ((uint64_t)(y & 0x1FFFFF) << 21) | (uint64_t)(z & 0x1FFFFF); };
      for (uint32_t i = 0; i < pts.size(); i++) if (std::isfinite(pts[i].x) && std::isfinite(pts[i].y) && std::isfinite(pts[i].z)) grid[key(cell(pts[i].x), cell(pts[i].y), cell(pts[i].z))].push_back(i);
      PointCloudPtr out(new PointCloud());
      const float r2 = (float)(radius * radius);
      for (uint32_t i = 0; i < pts.size(); i++) {
        const auto &p = pts[i];
        if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
        int k = 0;
        const int64_t cx = cell(p.x), cy = cell(p.y), cz = cell(p.z);
        for (int64_t dx = -1; dx <= 1 && k <= min_points; dx++) for (int64_t dy = -1; dy <= 1; dy++) for (int64_t dz = -1; dz <= 1; dz++) {
          auto it = grid.find(key(cx + dx, cy + dy, cz + dz));
          if (it == grid.end()) continue;
          for (uint32_t j : it->second) { const float ex = p.x - pts[j].x, ey = p.y - pts[j].y, ez = p.z - pts[j].z; if (ex * ex + ey * ey + ez * ez <= r2) k++; }
        }
        if (k > min_points) out->points.push_back(p);
      }
      if (debug_level > 2) std::cout << "filtered out a total of: " << pts.size() - out->points.size() << " outliers" << std::endl;
      out->width = (uint32_t)out->points.size(); out->height = 1; out->is_dense = false;
      pc = out;
    }
  }
  // computeQualityMetric (quality_metrics_impl.hpp:82-239) on this codec's device
  bool computeQuality(const PointCloud &cloud_a, const PointCloud &cloud_b, ccv2_quality &q) {
    ensure();
    return ccv2_quality_metrics(h_, cloud_a.points.data(), cloud_a.points.size(), cloud_b.points.data(), cloud_b.points.size(), &q) == CCV2_OK;
  }
  const std::string &lastError() const { return last_error_; }

private:
  static void min_max_3d(const PointCloud &c, float mn[3], float mx[3]) {   // [PCL] getMinMax3D: finite points only
    mn[0] = mn[1] = mn[2] = 3.4028235e38f; mx[0] = mx[1] = mx[2] = -3.4028235e38f;
    for (const auto &p : c.points) {
      if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
  }
  void ensure() {
    if (h_) return;
    if (ccv2_create(&p_, device_, &h_) != CCV2_OK) throw std::runtime_error(std::string("ccv2_create: ") + ccv2_last_error(nullptr));
  }
  ccv2_params p_;
  ccv2_codec *h_;
  int device_;
  uint64_t metrics_[3];
  std::vector<unsigned char> buf_;
  std::string last_error_;
  float shared_macroblock_percentage_ = 0.f, shared_macroblock_convergence_percentage_ = 0.f;
  ccv2_delta_info delta_info_{};
};

}}  // namespace pcl::io

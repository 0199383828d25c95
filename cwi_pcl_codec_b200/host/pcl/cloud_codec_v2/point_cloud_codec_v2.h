// point_cloud_codec_v2.h (B200 facade) -- C++ host side of the drop-in: the reference's class name, constructor
// argument order and method names (cloud_codec_v2/include/pcl/cloud_codec_v2/point_cloud_codec_v2.h:70-197 of
// cwi-dis/cwi-pcl-codec), implemented as a thin forwarder to the C ABI in include/ccv2.h.  Nothing here computes:
// every byte of the stream is produced by the CUDA library.  Only PointXYZRGB is supported, like the reference's one
// explicit instantiation (cloud_codec_v2/src/point_cloud_codec_v2.cpp:45).
#pragma once
#include "../pcl_shim.h"
#include "../../../../include/ccv2.h"

#include <iostream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

namespace pcl { namespace io {

template <typename PointT> class OctreePointCloudCodecV2;

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
    cloud_arg->points.resize(cnt ? cnt : 1);
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

  // codec.h:193-197
  uint64_t *getPerformanceMetrics() { return metrics_; }
  const std::string &lastError() const { return last_error_; }

private:
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
};

}}  // namespace pcl::io

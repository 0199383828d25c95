// pcl_shim.h -- the few PCL types the cloud_codec_v2 boundary needs, for hosts WITHOUT PCL (this repository's build
// box has none).  With real PCL present, include <pcl/point_types.h> / <pcl/point_cloud.h> instead and define
// CCV2_HAVE_PCL: the facade in cloud_codec_v2/point_cloud_codec_v2.h only relies on the members declared here.
// Layout facts mirrored (not copied) from PCL: PointXYZRGB is 32 bytes, x,y,z,data[3] floats at 0..15, b,g,r,a bytes at
// 16..19 (also readable as the packed float `rgb` / uint32 `rgba`), padding to 32; PointCloud<T>::points is contiguous.
#pragma once
#ifndef CCV2_HAVE_PCL
#include <cstdint>
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZRGB {
  union { float data[4]; struct { float x, y, z; }; };
  union {
    struct { std::uint8_t b, g, r, a; };
    float rgb;
    std::uint32_t rgba;
    float data_c[4];
  };
  PointXYZRGB() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} { r = g = b = 0; a = 255; }
};
static_assert(sizeof(PointXYZRGB) == 32, "PointXYZRGB must be PCL's 32-byte record");

template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
};

namespace io {
// pcl::io::compression_Profiles_e (only the value evaluate_compression uses is meaningful here)
enum compression_Profiles_e {
  LOW_RES_ONLINE_COMPRESSION_WITHOUT_COLOR, LOW_RES_ONLINE_COMPRESSION_WITH_COLOR, MED_RES_ONLINE_COMPRESSION_WITHOUT_COLOR,
  MED_RES_ONLINE_COMPRESSION_WITH_COLOR, HIGH_RES_ONLINE_COMPRESSION_WITHOUT_COLOR, HIGH_RES_ONLINE_COMPRESSION_WITH_COLOR,
  LOW_RES_OFFLINE_COMPRESSION_WITHOUT_COLOR, LOW_RES_OFFLINE_COMPRESSION_WITH_COLOR, MED_RES_OFFLINE_COMPRESSION_WITHOUT_COLOR,
  MED_RES_OFFLINE_COMPRESSION_WITH_COLOR, HIGH_RES_OFFLINE_COMPRESSION_WITHOUT_COLOR, HIGH_RES_OFFLINE_COMPRESSION_WITH_COLOR,
  COMPRESSION_PROFILE_COUNT, MANUAL_CONFIGURATION
};
}  // namespace io
}  // namespace pcl
#endif

// real_pcl_dump.cpp -- pins the oracle against the REAL reference.  Needs what the reference needs (PCL 1.9/1.10, Boost,
// Eigen, libjpeg-turbo) and the unmodified cwi-pcl-codec tree; it cannot be built in the offline container this repository
// was developed in, which is why the oracle's PCL-inherited parts are "parity unpinned" (DESIGN.md section 2).
//
//   g++ -O2 -std=c++14 real_pcl_dump.cpp -I<cwi-pcl-codec>/cloud_codec_v2/include -I<cwi-pcl-codec>/jpeg_io/include \
//       $(pkg-config --cflags --libs pcl_io-1.10 pcl_octree-1.10 pcl_common-1.10) -lpcl_jpeg_io -lturbojpeg -ljpeg -o real_pcl_dump
//   python tools/export_golden_inputs.py /tmp/golden      # writes <case>.bin (n x 32-byte PointXYZRGB) and <case>.args
//   for f in /tmp/golden/*.bin; do ./real_pcl_dump "${f%.bin}"; done          # writes <case>.stream and <case>.decoded
//   python tools/compare_real_pcl.py /tmp/golden            # SHA-256 against tests/golden/stream_hashes.json
//
// <case>.args holds the 14 constructor arguments of OctreePointCloudCodecV2 (point_cloud_codec_v2.h:108-143) in order,
// exactly as evaluate_compression passes them (evaluate_compression_impl.hpp:377-395), followed by macroblock_size and
// do_icp_color_offset for the two setters it calls (:415-417).
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/cloud_codec_v2/point_cloud_codec_v2.h>
#include <pcl/cloud_codec_v2/impl/point_cloud_codec_v2_impl.hpp>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

typedef pcl::PointXYZRGB PointT;
typedef pcl::io::OctreePointCloudCodecV2<PointT> Codec;

static Codec *make_codec(std::istream &a, int &macroblock, int &icp_offset) {
  int profile, stats, vg, iframe, do_color, color_bits, cct, centroid, scalable, connectivity, jq, threads;
  double pres, ores;
  a >> profile >> stats >> pres >> ores >> vg >> iframe >> do_color >> color_bits >> cct >> centroid >> scalable >> connectivity >> jq >> threads >> macroblock >> icp_offset;
  return new Codec((pcl::io::compression_Profiles_e)profile, stats != 0, pres, ores, vg != 0, (unsigned)iframe, do_color != 0,
                   (unsigned char)color_bits, (unsigned char)cct, centroid != 0, scalable != 0, connectivity != 0, jq, threads);
}

int main(int argc, char **argv) {
  if (argc < 2) { std::cerr << "usage: real_pcl_dump <case-prefix>\n"; return 2; }
  const std::string base = argv[1];
  std::ifstream args((base + ".args").c_str()), bin((base + ".bin").c_str(), std::ios::binary);
  if (!args || !bin) { std::cerr << "cannot open " << base << ".args / .bin\n"; return 2; }
  std::stringstream argtext; argtext << args.rdbuf();
  std::vector<char> raw((std::istreambuf_iterator<char>(bin)), std::istreambuf_iterator<char>());
  static_assert(sizeof(PointT) == 32, "PointXYZRGB is 32 bytes");
  pcl::PointCloud<PointT>::Ptr cloud(new pcl::PointCloud<PointT>());
  cloud->points.resize(raw.size() / 32);
  std::memcpy(cloud->points.data(), raw.data(), cloud->points.size() * 32);     // the frozen records are PCL's own layout
  cloud->width = (uint32_t)cloud->points.size(); cloud->height = 1; cloud->is_dense = false;

  int mb = 16, icp = 0;
  Codec *enc = make_codec(argtext, mb, icp);
  enc->setMacroblockSize(mb); enc->setDoICPColorOffset(icp != 0);
  std::stringstream coded;
  enc->encodePointCloud(cloud, coded);                                            // first frame of a fresh codec: frame_ID_ = 1
  const std::string s = coded.str();
  std::ofstream((base + ".stream").c_str(), std::ios::binary).write(s.data(), (std::streamsize)s.size());

  argtext.clear(); argtext.seekg(0);
  Codec *dec = make_codec(argtext, mb, icp);
  pcl::PointCloud<PointT>::Ptr out(new pcl::PointCloud<PointT>());
  std::stringstream in(s);
  dec->decodePointCloud(in, out);
  std::ofstream((base + ".decoded").c_str(), std::ios::binary).write((const char *)out->points.data(), (std::streamsize)(out->points.size() * 32));
  const uint64_t *m = enc->getPerformanceMetrics();
  std::cout << base << ": " << cloud->points.size() << " points -> " << s.size() << " bytes -> " << out->points.size()
            << " points; coded bytes tree/centroid/colour " << m[0] << "/" << m[1] << "/" << m[2] << "\n";
  delete enc; delete dec;
  return 0;
}

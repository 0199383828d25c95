// real_pcl_delta_dump.cpp -- the inter-frame half of tools/real_pcl_dump.cpp: runs the UNMODIFIED reference's delta-frame
// encoder and decoder on the frozen frame pair of tests/golden/delta_inputs.npz, and the reference DECODER on this
// repository's streams.  Like real_pcl_dump.cpp it needs PCL 1.9/1.10 + the cwi-pcl-codec tree and cannot be built in the
// offline container (same compile line, see tools/diff_against_real_pcl.md).
//
//   python tools/compare_real_pcl_delta.py export /tmp/delta     # <case>.f0.bin, <case>.f1.bin, <case>.args, <case>.our_istream, <case>.our_pstream
//   for a in /tmp/delta/*.args; do ./real_pcl_delta_dump "${a%.args}"; done
//   python tools/compare_real_pcl_delta.py check /tmp/delta
//
// What it writes per case:
//   .icloud            getOutputCloud() of the intra encode of frame 0 (what evaluate_compression predicts from, eval.hpp:862)
//   .ref_istream/.ref_pstream   encodePointCloudDeltaFrame(icloud, frame 1) of the reference (codec.h:184-186)
//   .ref_decoded       decodePointCloudDeltaFrame of the reference on ITS streams
//   .ref_decoded_ours  decodePointCloudDeltaFrame of the reference on OUR streams (must equal our own decode bit for bit:
//                      the decoder has no registration in it)
//   .stats             shared / convergence percentages (codec.h:200-210)
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
typedef pcl::PointCloud<PointT> Cloud;
typedef pcl::io::OctreePointCloudCodecV2<PointT> Codec;

static Codec *make_codec(std::istream &a, int &macroblock, int &icp_offset, int &icp_on_original) {
  int profile, stats, vg, iframe, do_color, color_bits, cct, centroid, scalable, connectivity, jq, threads;
  double pres, ores;
  a >> profile >> stats >> pres >> ores >> vg >> iframe >> do_color >> color_bits >> cct >> centroid >> scalable >> connectivity >> jq >> threads >> macroblock >> icp_offset >> icp_on_original;
  return new Codec((pcl::io::compression_Profiles_e)profile, stats != 0, pres, ores, vg != 0, (unsigned)iframe, do_color != 0,
                   (unsigned char)color_bits, (unsigned char)cct, centroid != 0, scalable != 0, connectivity != 0, jq, threads);
}
static Cloud::Ptr load(const std::string &path) {
  std::ifstream bin(path.c_str(), std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(bin)), std::istreambuf_iterator<char>());
  Cloud::Ptr c(new Cloud());
  c->points.resize(raw.size() / 32);
  std::memcpy(c->points.data(), raw.data(), c->points.size() * 32);
  c->width = (uint32_t)c->points.size(); c->height = 1; c->is_dense = false;
  return c;
}
static std::string slurp(const std::string &path) { std::ifstream f(path.c_str(), std::ios::binary); std::stringstream s; s << f.rdbuf(); return s.str(); }
static void dump(const std::string &path, const void *p, size_t n) { std::ofstream(path.c_str(), std::ios::binary).write((const char *)p, (std::streamsize)n); }

int main(int argc, char **argv) {
  if (argc < 2) { std::cerr << "usage: real_pcl_delta_dump <case-prefix>\n"; return 2; }
  const std::string base = argv[1];
  std::stringstream argtext; argtext << slurp(base + ".args");
  static_assert(sizeof(PointT) == 32, "PointXYZRGB is 32 bytes");
  Cloud::Ptr f0 = load(base + ".f0.bin"), f1 = load(base + ".f1.bin");
  int mb = 16, off = 0, orig = 0;
  Codec *enc = make_codec(argtext, mb, off, orig);
  enc->setMacroblockSize(mb); enc->setDoICPColorOffset(off != 0);
  std::stringstream intra;
  enc->encodePointCloud(f0, intra);
  Cloud::Ptr icloud = orig ? f0 : enc->getOutputCloud();
  dump(base + ".icloud", icloud->points.data(), icloud->points.size() * 32);
  Cloud::Ptr predicted(new Cloud());
  std::stringstream is, ps;
  enc->encodePointCloudDeltaFrame(icloud, f1, predicted, is, ps, orig != 0, false);       // as eval.hpp:506 calls it
  const std::string i_s = is.str(), p_s = ps.str();
  dump(base + ".ref_istream", i_s.data(), i_s.size()); dump(base + ".ref_pstream", p_s.data(), p_s.size());
  std::ofstream((base + ".stats").c_str()) << enc->getMacroBlockPercentage() << " " << enc->getMacroBlockConvergencePercentage() << "\n";
  {
    Cloud::Ptr out(new Cloud());
    std::stringstream i2(i_s), p2(p_s);
    enc->decodePointCloudDeltaFrame(icloud, out, i2, p2);
    dump(base + ".ref_decoded", out->points.data(), out->points.size() * 32);
  }
  const std::string oi = slurp(base + ".our_istream"), op = slurp(base + ".our_pstream");
  if (!oi.empty() || !op.empty()) {
    Cloud::Ptr out(new Cloud());
    std::stringstream i2(oi), p2(op);
    enc->decodePointCloudDeltaFrame(icloud, out, i2, p2);
    dump(base + ".ref_decoded_ours", out->points.data(), out->points.size() * 32);
  }
  std::cout << base << ": I cloud " << icloud->points.size() << " points, reference delta frame " << i_s.size() << " + " << p_s.size() << " bytes\n";
  delete enc;
  return 0;
}

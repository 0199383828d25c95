// evaluate_compression_mini -- a minimal C++ harness over the facade, shaped like the reference's
// evaluate_compression round trip (apps/evaluate_compression/.../evaluate_compression_impl.hpp:377-414 codec construction,
// :448-494 do_encoding / do_decoding with a timer around exactly the codec call, :818-843 encode -> decode per frame).
// It keeps the option names of the reference's parameter surface that matter for the intra path (eval.hpp:137-169).
// Input: a raw file of 32-byte PointXYZRGB records (--input) or a generated test cloud (--synthetic N).
#include "pcl/cloud_codec_v2/point_cloud_codec_v2.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

static void usage() {
  std::printf("usage: evaluate_compression_mini [--input file.xyzrgb32 | --synthetic N] [--octree_bits 11] [--enh_bits 0]\n"
              "       [--color_bits 8] [--color_coding_type 1] [--keep_centroid 0] [--jpeg_quality 85] [--frames 1]\n"
              "       [--dump_stream out.bin] [--dump_cloud out.xyzrgb32] [--device 0]\n");
}

int main(int argc, char **argv) {
  std::string input, dump_stream, dump_cloud;
  long synthetic = 0; int octree_bits = 11, enh_bits = 0, color_bits = 8, color_coding_type = 1, keep_centroid = 0, jpeg_quality = 85, frames = 1, device = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&](const char *name) -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", name); std::exit(2); } return argv[++i]; };
    if (a == "--help" || a == "-h") { usage(); return 0; }
    else if (a == "--input") input = next("--input");
    else if (a == "--synthetic") synthetic = std::atol(next("--synthetic"));
    else if (a == "--octree_bits") octree_bits = std::atoi(next(a.c_str()));
    else if (a == "--enh_bits") enh_bits = std::atoi(next(a.c_str()));
    else if (a == "--color_bits") color_bits = std::atoi(next(a.c_str()));
    else if (a == "--color_coding_type") color_coding_type = std::atoi(next(a.c_str()));
    else if (a == "--keep_centroid") keep_centroid = std::atoi(next(a.c_str()));
    else if (a == "--jpeg_quality" || a == "-j") jpeg_quality = std::atoi(next(a.c_str()));
    else if (a == "--frames") frames = std::atoi(next(a.c_str()));
    else if (a == "--dump_stream") dump_stream = next(a.c_str());
    else if (a == "--dump_cloud") dump_cloud = next(a.c_str());
    else if (a == "--device") device = std::atoi(next(a.c_str()));
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 2; }
  }
  typedef pcl::PointCloud<pcl::PointXYZRGB> Cloud;
  Cloud::Ptr cloud(new Cloud());
  if (!input.empty()) {
    std::ifstream f(input, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", input.c_str()); return 1; }
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    cloud->points.resize(s.size() / 32);
    std::memcpy(cloud->points.data(), s.data(), cloud->points.size() * 32);
  } else {
    if (synthetic <= 0) { usage(); return 2; }
    cloud->points.resize(synthetic);
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    for (auto &p : cloud->points) {
      const double u = rnd() * 6.283185307179586, v = std::acos(2 * rnd() - 1);
      p.x = (float)(0.5 + 0.3 * std::sin(v) * std::cos(u)); p.y = (float)(0.5 + 0.25 * std::sin(v) * std::sin(u)); p.z = (float)(0.5 + 0.35 * std::cos(v));
      p.r = (std::uint8_t)(128 + 100 * std::sin(20 * p.x)); p.g = (std::uint8_t)(128 + 90 * std::cos(17 * p.y)); p.b = (std::uint8_t)(128 + 110 * std::sin(13 * p.z)); p.a = 255;
    }
  }
  cloud->width = (std::uint32_t)cloud->points.size(); cloud->height = 1;
  // eval.hpp:377-395: MANUAL_CONFIGURATION, resolutions from the bit counts, voxel-grid mode, I frames only
  const double point_res = std::pow(2.0, -(octree_bits + enh_bits)), octree_res = std::pow(2.0, -octree_bits);
  try {
    pcl::io::OctreePointCloudCodecV2<pcl::PointXYZRGB> encoder(pcl::io::MANUAL_CONFIGURATION, false, point_res, octree_res, true, 0, color_bits > 0,
                                                               (unsigned char)color_bits, (unsigned char)color_coding_type, keep_centroid != 0, false, false, jpeg_quality, 1, device);
    pcl::io::OctreePointCloudCodecV2<pcl::PointXYZRGB> decoder(pcl::io::MANUAL_CONFIGURATION, false, point_res, octree_res, true, 0, color_bits > 0,
                                                               (unsigned char)color_bits, (unsigned char)color_coding_type, keep_centroid != 0, false, false, jpeg_quality, 1, device);
    std::printf("frame;points;compressed_byte_size;octree_bytes;centroid_bytes;color_bytes;decoded_points;encoding_time_ms;decoding_time_ms\n");
    for (int fr = 0; fr < frames; fr++) {
      std::stringstream ss;
      auto t0 = std::chrono::steady_clock::now();
      encoder.encodePointCloud(cloud, ss);
      auto t1 = std::chrono::steady_clock::now();
      const std::string data = ss.str();
      uint64_t *m = encoder.getPerformanceMetrics();
      Cloud::Ptr dec(new Cloud());
      std::stringstream in(data);
      auto t2 = std::chrono::steady_clock::now();
      decoder.decodePointCloud(in, dec);
      auto t3 = std::chrono::steady_clock::now();
      std::printf("%d;%zu;%zu;%llu;%llu;%llu;%zu;%.3f;%.3f\n", fr, cloud->points.size(), data.size(), (unsigned long long)m[0], (unsigned long long)m[1],
                  (unsigned long long)m[2], dec->points.size(), std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t3 - t2).count());
      if (fr == 0 && !dump_stream.empty()) { std::ofstream o(dump_stream, std::ios::binary); o.write(data.data(), (std::streamsize)data.size()); }
      if (fr == 0 && !dump_cloud.empty()) { std::ofstream o(dump_cloud, std::ios::binary); o.write((const char *)dec->points.data(), (std::streamsize)(dec->points.size() * 32)); }
      if (!encoder.lastError().empty() || !decoder.lastError().empty()) { std::fprintf(stderr, "codec error: %s %s\n", encoder.lastError().c_str(), decoder.lastError().c_str()); return 1; }
    }
  } catch (const std::exception &e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}

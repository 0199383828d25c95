// evaluate_compression (B200 harness) -- the reference's evaluation tool for the intra path, over the facade:
//   apps/evaluate_compression/include/pcl/apps/evaluate_compression/impl/evaluate_compression_impl.hpp
//     :137-169  the 29 options (same names, short flags and defaults -- note jpeg_quality defaults to 0, -j)
//     :250-305  optional ../parameter_config.txt or ./parameter_config.txt (key=value); the command line wins
//     :377-417  codec construction from the options
//     :683-792  evaluate(): directory of .ply / .pcd files, groups of group_size files
//     :795-897  evaluate_group(): outlier filter, bounding-box normalisation, per frame encode -> decode -> quality ->
//               restore_scaling -> pointcloud_<n>.ply, CSV line
//   .../impl/quality_metrics_impl.hpp:242-285  CSV header / line
//     :498-527, 854-889  do_delta_coding: frame i+1 coded against frame i (encoder's simplified cloud, or the input with
//               icp_on_original), decoded against the decoded frame i, predictive_quality_csv, delta_decoded_pc_<n>.ply
// Out of scope (ignored with a note, like the reference's own "not implemented" options): visualization, algorithm V1, num_threads.
#include "pcl/cloud_codec_v2/point_cloud_codec_v2.h"
#include "pcl/io/cloud_io_lite.h"

#include <chrono>
#include <dirent.h>
#include <map>
#include <sys/stat.h>

typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> Cloud;
typedef pcl::io::OctreePointCloudCodecV2<PointT> Codec;

struct Opt { const char *name; char shortf; const char *def; bool is_bool; const char *help; };
static const Opt OPTS[] = {                                                   // eval.hpp:137-169
  {"help", 'h', "", true, "produce help message"},
  {"K_outlier_filter", 'K', "0", false, "K neighbours for radius outlier filter"},
  {"radius", 0, "0.01", false, "radius outlier filter, maximum radius"},
  {"group_size", 'g', "0", false, "maximum number of files to be compressed together (0=read all files, then en(de)code 1 by 1)"},
  {"bb_expand_factor", 'f', "0.20", false, "bounding box expansion to keep bounding box accross frames"},
  {"algorithm", 'a', "V2", false, "compression algorithm ('V1' or 'V2')"},
  {"input_directories", 'i', "", false, "Directory containing supported files (.pcd or .ply)"},
  {"output_directory", 'o', "", false, "Directory to store decompressed pointclouds (.ply)"},
  {"show_statistics", 's', "0", true, "gather and show a bunch of releavant statistical data"},
  {"visualization", 'v', "0", true, "show both original and decoded PointClouds graphically"},
  {"point_resolution", 'p', "0.20", false, "XYZ resolution of point coordinates"},
  {"octree_resolution", 'r', "0.20", false, "voxel size"},
  {"octree_bits", 'b', "11", false, "octree resolution (bits)"},
  {"color_bits", 'c', "8", false, "color resolution (bits)"},
  {"enh_bits", 'e', "0", false, "bits to code the points towards the center"},
  {"color_coding_type", 't', "1", false, "pcl=0,jpeg=1 or graph transform"},
  {"macroblock_size", 'm', "16", false, "size of macroblocks used for predictive frame (must be power of 2)"},
  {"keep_centroid", 0, "0", false, "keep voxel grid positions or not"},
  {"create_scalable", 0, "0", false, "create scalable bitstream (not yet implemented)"},
  {"do_connectivity_coding", 0, "0", false, "connectivity coding (not yet implemented)"},
  {"icp_on_original", 0, "0", false, "icp_on_original"},
  {"jpeg_quality", 'j', "0", false, "jpeg quality parameter"},
  {"do_delta_coding", 'd', "0", false, "use delta (predictive) en(de)coding"},
  {"do_quality_computation", 'q', "0", false, "compute quality of en(de)coding"},
  {"do_icp_color_offset", 0, "0", false, "do color offset coding on predictive frames"},
  {"num_threads", 'n', "1", false, "number of omp cores (1=default, 1 thread, no parallel execution)"},
  {"intra_frame_quality_csv", 0, "intra_frame_quality.csv", false, "intra frame coding quality results file name (.csv file)"},
  {"predictive_quality_csv", 0, "predictive_quality.csv", false, "predictive coding quality results file name (.csv file)"},
  {"debug_level", 0, "0", false, "debug print level (0=no debug print, 3=all debug print)"},
  {"device", 0, "0", false, "CUDA device ordinal (this harness only)"},
  {"skip_coding", 0, "0", true, "load, filter, normalise, restore and write only: no codec call, no GPU (this harness only)"},
};
static const int NOPTS = sizeof OPTS / sizeof OPTS[0];

struct Options {
  std::map<std::string, std::string> v; std::map<std::string, bool> from_cmdline;
  std::vector<std::string> input_directories;
  long i(const char *k) const { return std::atol(v.at(k).c_str()); }
  double d(const char *k) const { return std::atof(v.at(k).c_str()); }
  bool b(const char *k) const { const std::string &s = v.at(k); return s == "1" || s == "true" || s == "on" || s == "yes"; }
  const std::string &s(const char *k) const { return v.at(k); }
};
static const Opt *find_opt(const std::string &name) { for (int k = 0; k < NOPTS; k++) if (name == OPTS[k].name) return &OPTS[k]; return nullptr; }
static const Opt *find_short(char c) { for (int k = 0; k < NOPTS; k++) if (OPTS[k].shortf == c) return &OPTS[k]; return nullptr; }

static bool parse_options(int argc, char **argv, Options &o) {
  for (int k = 0; k < NOPTS; k++) o.v[OPTS[k].name] = OPTS[k].def;
  bool ok = true;
  for (int i = 1; i < argc; i++) {                                            // command line first: it takes precedence (eval.hpp:268-270)
    std::string a = argv[i], val; const Opt *op = nullptr; bool has_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      const size_t eq = a.find('=');
      op = find_opt(a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2));
      if (eq != std::string::npos) { val = a.substr(eq + 1); has_val = true; }
    } else if (a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
      op = find_short(a[1]);
      if (a.size() > 2) { val = a.substr(2); has_val = true; }
    } else { o.input_directories.push_back(a); continue; }                    // positional: input_directories (eval.hpp:169)
    if (!op) { std::cerr << "Unrecognized options on command line:\n" << a << "\n"; ok = false; continue; }
    if (!has_val) {
      if (op->is_bool && (i + 1 >= argc || argv[i + 1][0] == '-')) val = "1";   // implicit_value(true)
      else if (i + 1 < argc) val = argv[++i];
      else { std::cerr << "option " << a << " needs a value\n"; return false; }
    }
    if (std::string(op->name) == "input_directories") o.input_directories.push_back(val);
    else { o.v[op->name] = val; o.from_cmdline[op->name] = true; }
  }
  if (!ok) return false;
  std::ifstream cfg("../parameter_config.txt");                              // eval.hpp:256-266
  if (!cfg) cfg.open("parameter_config.txt");
  if (!cfg) std::cerr << " Optional file 'parameter_config.txt' not found in the working directory or its parent.\n";
  std::string line;
  while (cfg && std::getline(cfg, line)) {
    const size_t h = line.find('#'); if (h != std::string::npos) line.erase(h);
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    auto trim = [](std::string s) { const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
    const std::string k = trim(line.substr(0, eq)), v = trim(line.substr(eq + 1));
    const Opt *op = find_opt(k);
    if (!op) { std::cerr << "Unrecognized options in configuration file:\n" << k << "\n"; return false; }
    if (k == "input_directories") { if (o.input_directories.empty()) o.input_directories.push_back(v); }
    else if (!o.from_cmdline.count(k)) o.v[k] = v;
  }
  return true;
}

static void print_csv_header(std::ostream &o) {                                // quality_metrics_impl.hpp:265-285
  o << "compression setting; " << "in point count;" << "out point count;" << "compressed_byte_size;" << "compressed_byte_size_per_output_point;"
    << "octree_byte_size_per_voxel;" << "centroid_byte_size_per_voxel;" << "color_byte_size_per_voxel;" << "symm_rms;" << "symm_haussdorff;" << "psnr_db;"
    << "psnr_colors_y;" << "psnr_colors_u;" << "psnr_colors_v;" << "encoding_time_ms;" << "decoding_time_ms;" << std::endl;
}
struct FrameStats { size_t compressed_size; uint64_t bytes[3]; double enc_ms, dec_ms; };
static void print_csv_line(std::ostream &o, const std::string &setting, const ccv2_quality &q, const FrameStats &f) {   // quality_metrics_impl.hpp:242-262
  const double n = 1.0 * q.out_point_count;
  o << setting << ";" << q.in_point_count << ";" << q.out_point_count << ";" << f.compressed_size << ";" << f.compressed_size / n << ";"
    << f.bytes[0] / n << ";" << f.bytes[1] / n << ";" << f.bytes[2] / n << ";" << q.symm_rms << ";" << q.symm_hausdorff << ";" << q.psnr_db << ";"
    << q.psnr_yuv[0] << ";" << q.psnr_yuv[1] << ";" << q.psnr_yuv[2] << ";" << f.enc_ms << ";" << f.dec_ms << ";" << std::endl;
}

static bool ends_with(const std::string &s, const char *suf) { const size_t n = std::strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

int main(int argc, char **argv) {
  Options o;
  if (!parse_options(argc, argv, o)) return 1;
  if (o.b("help")) { for (int k = 0; k < NOPTS; k++) std::printf("  --%s%s%c (=%s)\t%s\n", OPTS[k].name, OPTS[k].shortf ? " -" : "", OPTS[k].shortf ? OPTS[k].shortf : ' ', OPTS[k].def, OPTS[k].help); return 0; }
  const int debug_level = (int)o.i("debug_level");
  if (debug_level > 0) { std::cout << "debug_level=" << debug_level << "\n"; for (auto &kv : o.v) std::cout << "\t " << kv.first << "=" << kv.second << "\n"; }
  if (o.s("algorithm") != "V2") { std::cerr << "only algorithm V2 (cloud_codec_v2) is implemented here\n"; return 1; }
  if (o.b("visualization")) std::cerr << "note: No visualization configured\n";
  if (o.input_directories.size() > 1) { std::cout << "Fusing multiple directories not implemented.\n"; return 1; }
  if (o.input_directories.empty()) { std::cout << "Need to specify a directory containing Point Cloud files (.pcd or .ply).\n"; return 1; }
  const int octree_bits = (int)o.i("octree_bits"), enh_bits = (int)o.i("enh_bits"), color_bits = (int)o.i("color_bits");
  const int group_size = (int)o.i("group_size"), device = (int)o.i("device");
  const double bb_expand = o.d("bb_expand_factor");
  // eval.hpp:381-384: resolutions come from the bit counts; point_resolution / octree_resolution are only used when octree_bits <= 0
  const double point_res = octree_bits > 0 ? std::pow(2.0, -1.0 * (octree_bits + enh_bits)) : o.d("point_resolution");
  const double octree_res = octree_bits > 0 ? std::pow(2.0, -1.0 * octree_bits) : o.d("point_resolution");
  std::ostringstream setting;                                                 // eval.hpp:740
  setting << "octree_bits=" << octree_bits << " color_bits=" << color_bits << " enh._bits=" << enh_bits << "_colortype=" << o.i("color_coding_type") << " centroid=" << o.i("keep_centroid");
  std::ofstream csv;
  if (!o.s("intra_frame_quality_csv").empty()) { csv.open(o.s("intra_frame_quality_csv").c_str()); print_csv_header(csv); }
  std::ofstream pcsv;
  if (!o.s("predictive_quality_csv").empty()) { pcsv.open(o.s("predictive_quality_csv").c_str()); print_csv_header(pcsv); }
  std::vector<std::string> filenames;
  {
    DIR *d = opendir(o.input_directories[0].c_str());
    if (!d) { std::cerr << "'" << o.input_directories[0] << "' is not a directory.\n"; return 1; }
    while (dirent *e = readdir(d)) { std::string n = e->d_name; if (n != "." && n != "..") filenames.push_back(o.input_directories[0] + "/" + n); }
    closedir(d);
    std::sort(filenames.begin(), filenames.end());
  }
  int output_index = 0;
  { std::stringstream ss(filenames.empty() ? std::string() : filenames[0]); int v = -1; ss >> v; if (v > 0) output_index = v; }   // eval.hpp:762-768
  auto make_codec = [&](bool stats) {                                         // eval.hpp:377-414
    return std::unique_ptr<Codec>(new Codec(pcl::io::MANUAL_CONFIGURATION, stats, point_res, octree_res, true, 0, color_bits > 0, (unsigned char)color_bits,
                                            (unsigned char)o.i("color_coding_type"), o.i("keep_centroid") != 0, o.b("create_scalable"), false, (int)o.i("jpeg_quality"), (int)o.i("num_threads"), device));
  };
  std::unique_ptr<Codec> encoder, decoder;
  auto complete_initialization = [&]() { encoder = make_codec(o.b("show_statistics")); decoder = make_codec(false); encoder->setMacroblockSize((int)o.i("macroblock_size")); encoder->setDoICPColorOffset(o.b("do_icp_color_offset")); };
  try {
    complete_initialization();
    std::vector<Cloud::Ptr> group;
    size_t count = 0;
    auto evaluate_group = [&]() -> bool {                                      // eval.hpp:795-897
      std::vector<Cloud::Ptr> working;
      for (auto &g : group) working.push_back(Cloud::Ptr(new Cloud(*g)));
      if (o.i("K_outlier_filter") > 0) Codec::remove_outliers(working, (int)o.i("K_outlier_filter"), o.d("radius"), (unsigned)debug_level);
      pcl::io::BoundingBox bb{};
      std::vector<pcl::io::BoundingBox> boxes;
      if (bb_expand > 0.0) bb = Codec::normalize_pointclouds(working, boxes, bb_expand, (unsigned)debug_level);
      for (size_t i = 0; i < working.size(); i++) {
        Cloud::Ptr pc = working[i];
        if (o.b("skip_coding")) {
          Cloud::Ptr rescaled(new Cloud(*pc));
          if (bb_expand > 0.0) Codec::restore_scaling(rescaled, bb);
          if (!o.s("output_directory").empty()) { mkdir(o.s("output_directory").c_str(), 0777); pcl::io_lite::save_ply_ascii(o.s("output_directory") + "/pointcloud_" + std::to_string(output_index++) + ".ply", *rescaled); }
          continue;
        }
        std::stringstream ss;
        FrameStats fs{};
        auto t0 = std::chrono::steady_clock::now();
        encoder->encodePointCloud(pc, ss);
        auto t1 = std::chrono::steady_clock::now();
        fs.enc_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        uint64_t *m = encoder->getPerformanceMetrics();
        fs.bytes[0] = m[0]; fs.bytes[1] = m[1]; fs.bytes[2] = m[2];
        const std::string s = ss.str();
        fs.compressed_size = s.size();
        std::cout << " octreeCoding " << fs.compressed_size << " bytes  base layer  " << std::endl;
        std::stringstream coded(s);
        Cloud::Ptr out(new Cloud());
        auto t2 = std::chrono::steady_clock::now();
        decoder->decodePointCloud(coded, out);
        auto t3 = std::chrono::steady_clock::now();
        fs.dec_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
        if (!encoder->lastError().empty() || !decoder->lastError().empty()) { std::cerr << "codec error: " << encoder->lastError() << " " << decoder->lastError() << "\n"; return false; }
        if (o.b("do_quality_computation")) {                                  // on the normalised clouds, like the reference (eval.hpp:838-846)
          ccv2_quality q;
          if (!decoder->computeQuality(*pc, *out, q)) { std::cerr << "quality computation failed\n"; return false; }
          if (csv.is_open()) print_csv_line(csv, setting.str(), q, fs);
        }
        Cloud::Ptr rescaled(new Cloud(*out));
        if (bb_expand > 0.0) Codec::restore_scaling(rescaled, bb);
        if (!o.s("output_directory").empty()) {
          mkdir(o.s("output_directory").c_str(), 0777);
          pcl::io_lite::save_ply_ascii(o.s("output_directory") + "/pointcloud_" + std::to_string(output_index++) + ".ply", *rescaled);
        }
        // iterative closest point predictive coding of the NEXT frame against this one (eval.hpp:854-889)
        if (o.b("do_delta_coding") && bb_expand >= 0 && i + 1 < working.size()) {
          Cloud::Ptr predicted(new Cloud());
          std::cout << " delta coding frame nr " << i + 1 << std::endl;
          std::stringstream p_pdat, p_idat;
          FrameStats ps{};
          const bool icp_on_original = o.b("icp_on_original");
          Cloud::Ptr icl = icp_on_original ? pc : encoder->getOutputCloud();
          auto d0 = std::chrono::steady_clock::now();
          encoder->encodePointCloudDeltaFrame(icl, working[i + 1], predicted, p_idat, p_pdat, icp_on_original, false);   // eval.hpp:506
          auto d1 = std::chrono::steady_clock::now();
          ps.enc_ms = std::chrono::duration<double, std::milli>(d1 - d0).count();
          const size_t ib = p_idat.str().size(), pb = p_pdat.str().size();
          ps.bytes[0] = ib; ps.bytes[1] = pb; ps.bytes[2] = 0; ps.compressed_size = ib + pb;                         // eval.hpp:508-511
          std::cout << " encoded a predictive frame: coded " << ib << " bytes intra and " << pb << " inter frame encoded " << std::endl;
          auto d2 = std::chrono::steady_clock::now();
          encoder->decodePointCloudDeltaFrame(out, predicted, p_idat, p_pdat);                                          // eval.hpp:525: predicted from the DECODED frame i
          auto d3 = std::chrono::steady_clock::now();
          ps.dec_ms = std::chrono::duration<double, std::milli>(d3 - d2).count();
          if (!encoder->lastError().empty()) { std::cerr << "codec error: " << encoder->lastError() << "\n"; return false; }
          std::cout << " shared macroblocks " << encoder->getMacroBlockPercentage() << ", of which predicted " << encoder->getMacroBlockConvergencePercentage() << std::endl;
          if (o.b("do_quality_computation")) {
            ccv2_quality q;
            if (!decoder->computeQuality(*working[i + 1], *predicted, q)) { std::cerr << "quality computation failed\n"; return false; }
            if (pcsv.is_open()) print_csv_line(pcsv, setting.str(), q, ps);
          }
          Codec::restore_scaling(predicted, bb);
          if (!o.s("output_directory").empty()) pcl::io_lite::save_ply_ascii(o.s("output_directory") + "/delta_decoded_pc_" + std::to_string(output_index) + ".ply", *predicted);
        }
      }
      return true;
    };
    for (auto &fn : filenames) {                                               // eval.hpp:756-792
      Cloud::Ptr pc(new Cloud());
      bool loaded = false;
      if (ends_with(fn, ".ply")) loaded = pcl::io_lite::load_ply(fn, *pc);
      else if (ends_with(fn, ".pcd")) loaded = pcl::io_lite::load_pcd(fn, *pc);
      if (!loaded) continue;
      group.push_back(pc);
      count++;
      if (group_size == 0 && count < filenames.size()) continue;
      if (group_size == 0 || count == filenames.size() || count % (size_t)group_size == 0) {
        if (!evaluate_group()) return 1;
        complete_initialization();                                             // fresh codecs per group: frame ids restart at 1
        group.clear(); count = 0;
      }
    }
    if (!group.empty() && !evaluate_group()) return 1;
  } catch (const std::exception &e) { std::cerr << "error: " << e.what() << "\n"; return 1; }
  return 0;
}

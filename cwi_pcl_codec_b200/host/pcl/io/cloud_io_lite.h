// cloud_io_lite.h -- the file formats evaluate_compression reads and writes, without PCL: PLY (ascii, binary_little_endian)
// and PCD (ascii, binary) in, PLY (ascii) out.  Replaces, for the harness only, [PCL] PLYReader / PCDReader / PLYWriter as
// used at apps/evaluate_compression/.../evaluate_compression_impl.hpp:599-694 (load_ply_file, load_pcd_file,
// load_input_cloud) and :532-540 (do_output).  Vertex properties x, y, z (+ red, green, blue[, alpha] or a packed rgb /
// rgba field) are kept; every other property is skipped.
#pragma once
#include "../pcl_shim.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pcl { namespace io_lite {

struct Prop { std::string name; int size; char kind; };                    // kind: 'i' signed, 'u' unsigned, 'f' float

inline bool ply_type(const std::string &t, int &size, char &kind) {
  static const struct { const char *n; int s; char k; } T[] = {
    {"char", 1, 'i'}, {"int8", 1, 'i'}, {"uchar", 1, 'u'}, {"uint8", 1, 'u'}, {"short", 2, 'i'}, {"int16", 2, 'i'}, {"ushort", 2, 'u'}, {"uint16", 2, 'u'},
    {"int", 4, 'i'}, {"int32", 4, 'i'}, {"uint", 4, 'u'}, {"uint32", 4, 'u'}, {"float", 4, 'f'}, {"float32", 4, 'f'}, {"double", 8, 'f'}, {"float64", 8, 'f'}};
  for (auto &e : T) if (t == e.n) { size = e.s; kind = e.k; return true; }
  return false;
}
inline double read_scalar(const unsigned char *p, int size, char kind) {
  switch (kind) {
    case 'f': if (size == 4) { float v; std::memcpy(&v, p, 4); return v; } else { double v; std::memcpy(&v, p, 8); return v; }
    case 'u': if (size == 1) return p[0]; if (size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; } { uint32_t v; std::memcpy(&v, p, 4); return v; }
    default: if (size == 1) return (int8_t)p[0]; if (size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; } { int32_t v; std::memcpy(&v, p, 4); return v; }
  }
}
inline void assign_field(pcl::PointXYZRGB &pt, const std::string &name, double v, const unsigned char *raw, int size) {
  if (name == "x") pt.x = (float)v; else if (name == "y") pt.y = (float)v; else if (name == "z") pt.z = (float)v;
  else if (name == "red" || name == "r" || name == "diffuse_red") pt.r = (uint8_t)v;
  else if (name == "green" || name == "g" || name == "diffuse_green") pt.g = (uint8_t)v;
  else if (name == "blue" || name == "b" || name == "diffuse_blue") pt.b = (uint8_t)v;
  else if (name == "alpha") pt.a = (uint8_t)v;
  else if ((name == "rgb" || name == "rgba") && raw && size == 4) { uint32_t c; std::memcpy(&c, raw, 4); pt.b = c & 255; pt.g = (c >> 8) & 255; pt.r = (c >> 16) & 255; if (name == "rgba") pt.a = c >> 24; }
}

inline bool load_ply(const std::string &path, pcl::PointCloud<pcl::PointXYZRGB> &cloud) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::string line;
  if (!std::getline(f, line) || line.substr(0, 3) != "ply") return false;
  bool ascii = true, in_vertex = false, seen_vertex = false; size_t nvert = 0;
  std::vector<Prop> props;
  std::vector<std::pair<size_t, std::vector<Prop> > > before;             // elements stored ahead of the vertices (fixed-size properties only)
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line); std::string w; ls >> w;
    if (w == "format") { std::string fm; ls >> fm; if (fm == "ascii") ascii = true; else if (fm == "binary_little_endian") ascii = false; else return false; }
    else if (w == "element") { std::string name; size_t n; ls >> name >> n; in_vertex = name == "vertex"; if (in_vertex) { nvert = n; seen_vertex = true; } else if (!seen_vertex) before.push_back({n, {}}); }
    else if (w == "property") {
      std::string t; ls >> t;
      if (t == "list") { if (in_vertex || !seen_vertex) return false; continue; }
      Prop p; ls >> p.name; if (!ply_type(t, p.size, p.kind)) return false;
      if (in_vertex) props.push_back(p); else if (!seen_vertex && !before.empty()) before.back().second.push_back(p);
    } else if (w == "end_header") break;
  }
  if (!seen_vertex) return false;
  cloud.points.assign(nvert, pcl::PointXYZRGB());
  if (ascii) {
    for (auto &e : before) for (size_t i = 0; i < e.first; i++) std::getline(f, line);
    for (size_t i = 0; i < nvert; i++) {
      if (!std::getline(f, line)) return false;
      std::istringstream ls(line);
      for (auto &p : props) { double v; if (!(ls >> v)) return false; assign_field(cloud.points[i], p.name, v, nullptr, 0); }
    }
  } else {
    for (auto &e : before) { size_t rec = 0; for (auto &p : e.second) rec += p.size; f.seekg((std::streamoff)(rec * e.first), std::ios::cur); }
    size_t rec = 0; for (auto &p : props) rec += p.size;
    std::vector<unsigned char> buf(rec * nvert);
    f.read((char *)buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) return false;
    for (size_t i = 0; i < nvert; i++) {
      const unsigned char *q = buf.data() + i * rec;
      for (auto &p : props) { assign_field(cloud.points[i], p.name, read_scalar(q, p.size, p.kind), q, p.size); q += p.size; }
    }
  }
  cloud.width = (uint32_t)nvert; cloud.height = 1; cloud.is_dense = false;
  return true;
}

inline bool load_pcd(const std::string &path, pcl::PointCloud<pcl::PointXYZRGB> &cloud) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<std::string> fields, types; std::vector<int> sizes, counts; size_t npts = 0; std::string data, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line); std::string w; ls >> w;
    std::string t;
    if (w == "FIELDS") while (ls >> t) fields.push_back(t);
    else if (w == "SIZE") { int v; while (ls >> v) sizes.push_back(v); }
    else if (w == "TYPE") while (ls >> t) types.push_back(t);
    else if (w == "COUNT") { int v; while (ls >> v) counts.push_back(v); }
    else if (w == "POINTS") ls >> npts;
    else if (w == "DATA") { ls >> data; break; }
  }
  if (fields.empty() || sizes.size() != fields.size() || types.size() != fields.size()) return false;
  if (counts.empty()) counts.assign(fields.size(), 1);
  cloud.points.assign(npts, pcl::PointXYZRGB());
  if (data == "ascii") {
    for (size_t i = 0; i < npts; i++) {
      if (!std::getline(f, line)) return false;
      std::istringstream ls(line);
      for (size_t k = 0; k < fields.size(); k++) for (int c = 0; c < counts[k]; c++) {
        std::string tok; if (!(ls >> tok)) return false;
        if (c) continue;
        if ((fields[k] == "rgb" || fields[k] == "rgba") && sizes[k] == 4) {
          uint32_t bits;
          if (types[k] == "F") { float v = std::stof(tok); std::memcpy(&bits, &v, 4); } else bits = (uint32_t)std::stoul(tok);
          assign_field(cloud.points[i], fields[k], 0, (const unsigned char *)&bits, 4);
        } else assign_field(cloud.points[i], fields[k], std::stod(tok), nullptr, 0);
      }
    }
  } else if (data == "binary") {
    size_t rec = 0; for (size_t k = 0; k < fields.size(); k++) rec += (size_t)sizes[k] * counts[k];
    std::vector<unsigned char> buf(rec * npts);
    f.read((char *)buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) return false;
    for (size_t i = 0; i < npts; i++) {
      const unsigned char *q = buf.data() + i * rec;
      for (size_t k = 0; k < fields.size(); k++) {
        const char kind = types[k] == "F" ? 'f' : (types[k] == "U" ? 'u' : 'i');
        assign_field(cloud.points[i], fields[k], read_scalar(q, sizes[k], kind), q, sizes[k]);
        q += (size_t)sizes[k] * counts[k];
      }
    }
  } else return false;                                                    // binary_compressed (LZF) is not read
  cloud.width = (uint32_t)npts; cloud.height = 1; cloud.is_dense = false;
  return true;
}

inline bool save_ply_ascii(const std::string &path, const pcl::PointCloud<pcl::PointXYZRGB> &cloud) {
  FILE *o = std::fopen(path.c_str(), "w");
  if (!o) return false;
  std::fprintf(o, "ply\nformat ascii 1.0\ncomment cloud_codec_v2 B200 harness\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n"
                  "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n", cloud.points.size());
  for (const auto &p : cloud.points) std::fprintf(o, "%.9g %.9g %.9g %u %u %u\n", p.x, p.y, p.z, (unsigned)p.r, (unsigned)p.g, (unsigned)p.b);
  std::fclose(o);
  return true;
}

}}  // namespace pcl::io_lite

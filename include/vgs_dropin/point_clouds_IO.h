// point_clouds_IO.h (drop-in) — the I/O entry points the reference's driver uses (reference
// point_clouds_IO.h:64-108, point_clouds_IO.cpp:22-76, 147-169), without PCL:
//   inputTaskTxtFile      : every line of the task file, comments and blanks included; parameter k = line k
//   inputPointCloudData   : PCD v0.7 reader (DATA ascii | binary | binary_compressed (LZF); x y z as F 4, other fields skipped)
//   saveColoredClusters   : PCD writer (binary, x y z rgb), one deterministic colour per cluster
// The reference's PCLVisualizer windows (showColoredClusters) are out of scope (GUI).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "pcl_shim.h"

// IO.cpp:147-169 — getline per line; a trailing '\r' (the shipped task files are CRLF) stays in the string,
// atof/atoi ignore it (test:25-37)
inline std::vector<std::string> inputTaskTxtFile(const std::string& pathname_file) {
  std::vector<std::string> task_vector;
  std::ifstream f(pathname_file.c_str());
  std::string line;
  while (std::getline(f, line)) task_vector.push_back(line);
  return task_vector;
}

inline std::string vgs_rstrip(const std::string& s) {
  size_t e = s.size();
  while (e > 0 && (s[e - 1] == '\r' || s[e - 1] == '\n' || s[e - 1] == ' ' || s[e - 1] == '\t')) e--;
  return s.substr(0, e);
}

// LZF decompression (the codec of PCD "binary_compressed"): literal runs (ctrl < 32) and back references
// (length ctrl>>5 (+ extension byte when 7) + 2, offset ((ctrl & 31) << 8 | next byte) + 1)
inline bool vgs_lzf_decompress(const unsigned char* ip, size_t in_len, unsigned char* out, size_t out_len) {
  const unsigned char* const ie = ip + in_len;
  size_t op = 0;
  while (ip < ie) {
    unsigned ctrl = *ip++;
    if (ctrl < 32) {
      ctrl++;
      if (op + ctrl > out_len || ip + ctrl > ie) return false;
      std::memcpy(out + op, ip, ctrl);
      op += ctrl; ip += ctrl;
    } else {
      size_t len = ctrl >> 5;
      if (len == 7) { if (ip >= ie) return false; len += *ip++; }
      if (ip >= ie) return false;
      size_t off = ((size_t)(ctrl & 0x1f) << 8) + *ip++ + 1;
      len += 2;
      if (off > op || op + len > out_len) return false;
      for (size_t i = 0; i < len; i++, op++) out[op] = out[op - off];   // may overlap
    }
  }
  return op == out_len;
}

// IO.h:64-80 (pcl::io::loadPCDFile into PointXYZ: only x, y, z survive)
inline int inputPointCloudData(const std::string& name, PCXYZPtr cloud) {
  std::ifstream f(name.c_str(), std::ios::binary);
  if (!f) return -1;
  std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
  long npoints = -1; std::string data, line;
  while (std::getline(f, line)) {
    line = vgs_rstrip(line);
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line); std::string key; ss >> key;
    if (key == "FIELDS") { std::string t; while (ss >> t) fields.push_back(t); }
    else if (key == "SIZE") { int t; while (ss >> t) sizes.push_back(t); }
    else if (key == "TYPE") { char t; while (ss >> t) types.push_back(t); }
    else if (key == "COUNT") { int t; while (ss >> t) counts.push_back(t); }
    else if (key == "POINTS") ss >> npoints;
    else if (key == "DATA") { ss >> data; break; }
  }
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (npoints < 0 || fields.size() != sizes.size() || fields.size() != types.size()) return -2;
  int ix = -1, iy = -1, iz = -1; std::vector<int> offs(fields.size()); int rec = 0;
  for (size_t i = 0; i < fields.size(); i++) {
    offs[i] = rec; rec += sizes[i] * counts[i];
    if (fields[i] == "x") ix = (int)i;
    if (fields[i] == "y") iy = (int)i;
    if (fields[i] == "z") iz = (int)i;
  }
  if (ix < 0 || iy < 0 || iz < 0) return -3;
  cloud->points.clear();
  cloud->points.reserve((size_t)npoints);
  if (data == "ascii") {
    for (long p = 0; p < npoints; p++) {
      if (!std::getline(f, line)) break;
      // strtod per token: accepts the "nan" / "inf" tokens PCL writes for non-finite points (operator>> does not,
      // which would turn such a row into a ghost point at the origin)
      std::istringstream ss(line); std::vector<double> v; std::string tok;
      while (ss >> tok) {
        char* endp = nullptr;
        double t = std::strtod(tok.c_str(), &endp);
        if (endp == tok.c_str()) t = std::numeric_limits<double>::quiet_NaN();
        v.push_back(t);
      }
      int col = 0; float xyz[3] = {0, 0, 0};
      for (size_t i = 0; i < fields.size(); i++) {
        if ((int)i == ix && col < (int)v.size()) xyz[0] = (float)v[col];
        if ((int)i == iy && col < (int)v.size()) xyz[1] = (float)v[col];
        if ((int)i == iz && col < (int)v.size()) xyz[2] = (float)v[col];
        col += counts[i];
      }
      cloud->points.push_back(pcl::PointXYZ(xyz[0], xyz[1], xyz[2]));
    }
  } else if (data == "binary") {
    if (types[ix] != 'F' || sizes[ix] != 4 || types[iy] != 'F' || sizes[iy] != 4 || types[iz] != 'F' || sizes[iz] != 4) return -4;
    std::vector<char> buf((size_t)rec * (size_t)npoints);
    f.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) return -5;
    for (long p = 0; p < npoints; p++) {
      float x, y, z;
      std::memcpy(&x, &buf[(size_t)p * rec + offs[ix]], 4);
      std::memcpy(&y, &buf[(size_t)p * rec + offs[iy]], 4);
      std::memcpy(&z, &buf[(size_t)p * rec + offs[iz]], 4);
      cloud->points.push_back(pcl::PointXYZ(x, y, z));
    }
  } else if (data == "binary_compressed") {
    // u32 compressed size, u32 uncompressed size, LZF stream; the payload is field-major (all x, all y, ...)
    if (types[ix] != 'F' || sizes[ix] != 4 || types[iy] != 'F' || sizes[iy] != 4 || types[iz] != 'F' || sizes[iz] != 4) return -4;
    std::uint32_t csz = 0, usz = 0;
    f.read(reinterpret_cast<char*>(&csz), 4);
    f.read(reinterpret_cast<char*>(&usz), 4);
    if (!f || (size_t)usz != (size_t)rec * (size_t)npoints) return -7;
    std::vector<unsigned char> cbuf(csz), ubuf(usz);
    f.read(reinterpret_cast<char*>(cbuf.data()), csz);
    if ((size_t)f.gcount() != (size_t)csz) return -5;
    if (!vgs_lzf_decompress(cbuf.data(), csz, ubuf.data(), usz)) return -8;
    auto field_base = [&](int fi) { return (size_t)offs[fi] * (size_t)npoints; };   // field-major layout
    for (long p = 0; p < npoints; p++) {
      float x, y, z;
      std::memcpy(&x, &ubuf[field_base(ix) + (size_t)p * 4], 4);
      std::memcpy(&y, &ubuf[field_base(iy) + (size_t)p * 4], 4);
      std::memcpy(&z, &ubuf[field_base(iz) + (size_t)p * 4], 4);
      cloud->points.push_back(pcl::PointXYZ(x, y, z));
    }
  } else return -6;
  cloud->width = (std::uint32_t)cloud->points.size(); cloud->height = 1;
  return 0;
}

// IO.h:83-97 inputPointCloudData2 — PLY point cloud (pcl::io::loadPLYFile): header-driven reader for `format ascii`,
// `binary_little_endian` and `binary_big_endian`; the vertex element's x, y, z may be float or double and may be
// surrounded by any scalar properties (colours, normals, intensity ...); list properties inside the vertex element
// and elements before it are skipped by size; everything after the vertex element (faces) is ignored.
// Returns 0, or a negative code (-1 cannot open, -2 not a PLY / bad header, -3 no x/y/z, -4 truncated data).
inline int inputPointCloudData2(const std::string& name, PCXYZPtr cloud) {
  std::ifstream f(name, std::ios::binary);
  if (!f) return -1;
  struct Prop { std::string type, list_count; bool is_list = false; int role = -1; };
  struct Elem { std::string name; long count = 0; std::vector<Prop> props; };
  auto type_size = [](const std::string& t) -> int {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
  };
  std::string line;
  if (!std::getline(f, line) || vgs_rstrip(line) != "ply") return -2;
  int fmt = -1;   // 0 ascii, 1 little endian, 2 big endian
  std::vector<Elem> elems;
  bool ended = false;
  while (std::getline(f, line)) {
    line = vgs_rstrip(line);
    std::istringstream ss(line);
    std::string tok;
    ss >> tok;
    if (tok == "format") {
      std::string kind; ss >> kind;
      fmt = kind == "ascii" ? 0 : kind == "binary_little_endian" ? 1 : kind == "binary_big_endian" ? 2 : -1;
    } else if (tok == "element") {
      Elem e; ss >> e.name >> e.count; elems.push_back(e);
    } else if (tok == "property") {
      if (elems.empty()) return -2;
      Prop p; std::string t; ss >> t;
      if (t == "list") { p.is_list = true; ss >> p.list_count >> p.type; }
      else p.type = t;
      std::string pname; ss >> pname;
      if (!p.is_list) p.role = pname == "x" ? 0 : pname == "y" ? 1 : pname == "z" ? 2 : -1;
      elems.back().props.push_back(p);
    } else if (tok == "end_header") { ended = true; break; }
  }
  if (!ended || fmt < 0) return -2;
  auto read_scalar = [&](const std::string& t, double& v) -> bool {   // binary scalar -> double
    unsigned char b[8];
    const int sz = type_size(t);
    if (sz == 0) return false;
    f.read(reinterpret_cast<char*>(b), sz);
    if (f.gcount() != sz) return false;
    if (fmt == 2) for (int i = 0; i < sz / 2; i++) std::swap(b[i], b[sz - 1 - i]);
    if (t == "float" || t == "float32") { float x; std::memcpy(&x, b, 4); v = x; }
    else if (t == "double" || t == "float64") { double x; std::memcpy(&x, b, 8); v = x; }
    else if (t == "char" || t == "int8") { signed char x; std::memcpy(&x, b, 1); v = x; }
    else if (t == "uchar" || t == "uint8") { v = b[0]; }
    else if (t == "short" || t == "int16") { std::int16_t x; std::memcpy(&x, b, 2); v = x; }
    else if (t == "ushort" || t == "uint16") { std::uint16_t x; std::memcpy(&x, b, 2); v = x; }
    else if (t == "int" || t == "int32") { std::int32_t x; std::memcpy(&x, b, 4); v = x; }
    else { std::uint32_t x; std::memcpy(&x, b, 4); v = x; }
    return true;
  };
  cloud->points.clear();
  for (const Elem& e : elems) {
    const bool is_vertex = e.name == "vertex";
    if (is_vertex) {
      bool have[3] = {false, false, false};
      for (const Prop& p : e.props) if (p.role >= 0) have[p.role] = true;
      if (!(have[0] && have[1] && have[2])) return -3;
      cloud->points.reserve((size_t)e.count);
    }
    for (long i = 0; i < e.count; i++) {
      float xyz[3] = {0, 0, 0};
      if (fmt == 0) {
        if (!std::getline(f, line)) return -4;
        std::istringstream ss(line);
        for (const Prop& p : e.props) {
          double v = 0;
          if (p.is_list) { double c = 0; if (!(ss >> c)) return -4; for (long k = 0; k < (long)c; k++) if (!(ss >> v)) return -4; }
          else { if (!(ss >> v)) return -4; if (is_vertex && p.role >= 0) xyz[p.role] = (float)v; }
        }
      } else {
        for (const Prop& p : e.props) {
          double v = 0;
          if (p.is_list) { double c = 0; if (!read_scalar(p.list_count, c)) return -4; for (long k = 0; k < (long)c; k++) if (!read_scalar(p.type, v)) return -4; }
          else { if (!read_scalar(p.type, v)) return -4; if (is_vertex && p.role >= 0) xyz[p.role] = (float)v; }
        }
      }
      if (is_vertex) cloud->points.push_back(pcl::PointXYZ(xyz[0], xyz[1], xyz[2]));
    }
    if (is_vertex) break;   // faces and later elements are not needed
  }
  cloud->width = (std::uint32_t)cloud->points.size(); cloud->height = 1;
  return 0;
}

// IO.cpp:22-71 — coloured copy of the clustered points (points outside every cluster are not written)
inline void saveColoredClusters(const std::string& fileoutpath_name, PCXYZPtr input_cloud,
                                const std::vector<std::vector<int>>& clusters_points_idx) {
  size_t total = 0;
  for (auto& c : clusters_points_idx) total += c.size();
  FILE* f = std::fopen(fileoutpath_name.c_str(), "wb");
  if (!f) throw std::runtime_error("saveColoredClusters: cannot open " + fileoutpath_name);
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\n"
                  "COUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", total, total);
  for (size_t c = 0; c < clusters_points_idx.size(); c++) {
    std::uint32_t hsh = (std::uint32_t)c * 2654435761u;
    std::uint32_t rgb = (((hsh >> 8) & 255u) << 16) | (((hsh >> 16) & 255u) << 8) | ((hsh >> 24) & 255u);
    for (int p : clusters_points_idx[c]) {
      float xyz[3] = {input_cloud->points[p].x, input_cloud->points[p].y, input_cloud->points[p].z};
      std::fwrite(xyz, 4, 3, f);
      std::fwrite(&rgb, 4, 1, f);
    }
  }
  std::fclose(f);
}

// IO.cpp:73-76 — an already coloured cloud (the draw* members' output) as a binary XYZRGB PCD
inline void saveColoredClusters(const std::string& fileoutpath_name, pcl::PointCloud<pcl::PointXYZRGB>::Ptr colored_cloud) {
  FILE* f = std::fopen(fileoutpath_name.c_str(), "wb");
  if (!f) throw std::runtime_error("saveColoredClusters: cannot open " + fileoutpath_name);
  const size_t total = colored_cloud->points.size();
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\n"
                  "COUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", total, total);
  for (const auto& p : colored_cloud->points) {
    float xyz[3] = {p.x, p.y, p.z};
    std::uint32_t rgb = ((std::uint32_t)p.r << 16) | ((std::uint32_t)p.g << 8) | (std::uint32_t)p.b;
    std::fwrite(xyz, 4, 3, f);
    std::fwrite(&rgb, 4, 1, f);
  }
  std::fclose(f);
}

// IO.h:100-108 — plain XYZ cloud as a binary PCD; 0, or -1 when the file cannot be written (as the reference)
inline int outputPointCloudData(const std::string& outName, PCXYZPtr dataCloud) {
  FILE* f = std::fopen(outName.c_str(), "wb");
  if (!f) return -1;
  const size_t total = dataCloud->points.size();
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
                  "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", total, total);
  for (const auto& p : dataCloud->points) {
    float xyz[3] = {p.x, p.y, p.z};
    std::fwrite(xyz, 4, 3, f);
  }
  return std::fclose(f) == 0 ? 0 : -1;
}

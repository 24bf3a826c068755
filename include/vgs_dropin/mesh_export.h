// mesh_export.h — shared helpers of the draw* members of the two drop-in classes (reference
// voxel_segmentation.h:424-1104, supervoxel_segmentation.h:424-664): cube / wire-frame / normal-stick
// geometry with the reference's vertex numbering and polygon order, deterministic colours, typed
// fetches of the stage outputs through the C ABI, and an ASCII PLY writer for pcl::PolygonMesh.
// Host-side formatting only; nothing here computes segmentation results.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vgs_b200.h"
#include "pcl_shim.h"

namespace vgs_dropin {

// One colour per index.  The reference draws rand()%256 triples after srand(time(0)) (VS.h:436, 524,
// 672, 808, 960, 1033): not reproducible, so the drop-in hashes the voxel / cluster index instead.
inline void color_of(uint32_t index, uint8_t& r, uint8_t& g, uint8_t& b) {
  uint32_t h = index * 2654435761u;
  r = (uint8_t)(h >> 8); g = (uint8_t)(h >> 16); b = (uint8_t)(h >> 24);
}

template <class T>
std::vector<T> fetch(vgs_handle h, vgs_blob_kind kind) {
  size_t bytes = 0;
  if (vgs_debug_get(h, kind, nullptr, &bytes) != VGS_OK) throw std::runtime_error(std::string("libvgs_b200: ") + vgs_last_error(h));
  std::vector<T> v(bytes / sizeof(T));
  if (bytes && vgs_debug_get(h, kind, v.data(), &bytes) != VGS_OK)
    throw std::runtime_error(std::string("libvgs_b200: ") + vgs_last_error(h));
  return v;
}

// PCL's OctreePointCloud::genLeafNodeCenterFromOctreeKey, which the draw* members call through
// getVoxelCenterFromOctreeKey (VS.h:463, 551, 836, 1060): double resolution and double box minimum —
// not the float-narrowed copies that voxel_centers_ is built from (VS.h:2106).
inline pcl::PointXYZ leaf_center(const uint32_t key[3], double resolution, const double box_min[3]) {
  pcl::PointXYZ c;
  c.x = (float)(((double)key[0] + 0.5f) * resolution + box_min[0]);
  c.y = (float)(((double)key[1] + 0.5f) * resolution + box_min[1]);
  c.z = (float)(((double)key[2] + 0.5f) * resolution + box_min[2]);
  return c;
}

inline pcl::PointXYZRGB vertex(float x, float y, float z, uint8_t r, uint8_t g, uint8_t b) {
  pcl::PointXYZRGB v;
  v.x = x; v.y = y; v.z = z; v.r = r; v.g = g; v.b = b;
  return v;
}

inline void push_poly(pcl::PolygonMesh& m, uint32_t a, uint32_t b, uint32_t c) {
  pcl::Vertices t;
  t.vertices.push_back(a); t.vertices.push_back(b); t.vertices.push_back(c);
  m.polygons.push_back(t);
}

// Eight corners in the reference's numbering (VS.h:560-583): 0-3 bottom ring (-x-y, +x-y, +x+y, -x+y), 4-7 the top ring.
// float centre +- 0.5 * float resolution is evaluated in double and narrowed on assignment, as in the reference.
inline void push_corners(pcl::PointCloud<pcl::PointXYZRGB>& verts, const pcl::PointXYZ& c, float res, uint8_t r, uint8_t g, uint8_t b) {
  const float xl = (float)(c.x - 0.5 * res), xh = (float)(c.x + 0.5 * res);
  const float yl = (float)(c.y - 0.5 * res), yh = (float)(c.y + 0.5 * res);
  const float zl = (float)(c.z - 0.5 * res), zh = (float)(c.z + 0.5 * res);
  const float xs[8] = {xl, xh, xh, xl, xl, xh, xh, xl};
  const float ys[8] = {yl, yl, yh, yh, yl, yl, yh, yh};
  for (int j = 0; j < 8; j++) verts.points.push_back(vertex(xs[j], ys[j], j < 4 ? zl : zh, r, g, b));
}

// Twelve triangles of box i in the reference's order (VS.h:599-645)
inline void push_box_faces(pcl::PolygonMesh& m, uint32_t i) {
  static const uint8_t T[12][3] = {{0, 1, 2}, {0, 2, 3}, {4, 5, 6}, {4, 6, 7}, {0, 1, 5}, {0, 5, 4},
                                   {1, 2, 5}, {2, 6, 5}, {0, 3, 7}, {0, 7, 4}, {2, 3, 7}, {2, 7, 6}};
  for (auto& t : T) push_poly(m, i * 8 + t[0], i * 8 + t[1], i * 8 + t[2]);
}

// Twelve edges of box i as degenerate triangles (a, b, a), reference order (VS.h:893-934)
inline void push_box_edges(pcl::PolygonMesh& m, uint32_t i) {
  static const uint8_t E[12][2] = {{0, 1}, {0, 3}, {1, 2}, {2, 3}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  for (auto& e : E) push_poly(m, i * 8 + e[0], i * 8 + e[1], i * 8 + e[0]);
}

// ASCII PLY (vertex x y z red green blue; face list) — what pcl::io::savePLYFile would be used for.
inline int savePolygonMeshPLY(const std::string& path, const pcl::PolygonMesh& m) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return -1;
  std::fprintf(f, "ply\nformat ascii 1.0\ncomment vgs_b200 drop-in mesh export\nelement vertex %zu\n", m.cloud.points.size());
  std::fprintf(f, "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n");
  std::fprintf(f, "element face %zu\nproperty list uchar int vertex_indices\nend_header\n", m.polygons.size());
  for (const auto& v : m.cloud.points) std::fprintf(f, "%.9g %.9g %.9g %u %u %u\n", v.x, v.y, v.z, (unsigned)v.r, (unsigned)v.g, (unsigned)v.b);
  for (const auto& p : m.polygons) {
    std::fprintf(f, "%zu", p.vertices.size());
    for (uint32_t i : p.vertices) std::fprintf(f, " %u", i);
    std::fprintf(f, "\n");
  }
  return std::fclose(f) == 0 ? 0 : -1;
}

}  // namespace vgs_dropin

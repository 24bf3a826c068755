// pcl_shim.h — the handful of pcl:: / Eigen:: names the reference's public interface mentions
// (voxel_segmentation.h:44-52, 191; test:10-60), so the drop-in headers compile without PCL.
// With a real PCL installation define VGS_DROPIN_USE_PCL and include the PCL headers instead.
#pragma once
#ifndef VGS_DROPIN_USE_PCL
#include <cstdint>
#include <memory>
#include <vector>

namespace Eigen {
template <class T> using aligned_allocator = std::allocator<T>;
}

namespace pcl {
struct PointXYZ {       // 16 bytes like PCL's (float data[4])
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct PointXYZRGB {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  std::uint8_t b = 0, g = 0, r = 0, a = 255;
  float pad2_[3] = {0, 0, 0};
};
struct PointXYZRGBA : PointXYZRGB {};
struct Normal {
  float normal_x = 0, normal_y = 0, normal_z = 0, pad_ = 0;
  float curvature = 0;
  float pad2_[3] = {0, 0, 0};
};
template <class PointT> struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  std::size_t size() const { return points.size(); }
  void clear() { points.clear(); width = height = 0; }
};
struct Vertices { std::vector<std::uint32_t> vertices; };
// pcl::PolygonMesh keeps its vertices as a PCLPointCloud2 blob; the shim keeps the typed cloud that
// toPCLPointCloud2 would have serialised (the draw* members only ever put XYZRGB vertices in it).
struct PolygonMesh {
  typedef std::shared_ptr<PolygonMesh> Ptr;
  PointCloud<PointXYZRGB> cloud;
  std::vector<Vertices> polygons;
};
inline void toPCLPointCloud2(const PointCloud<PointXYZRGB>& src, PointCloud<PointXYZRGB>& dst) { dst = src; }
}  // namespace pcl
#endif

typedef pcl::PointCloud<pcl::PointXYZRGB>::Ptr PCXYZRGBPtr;
typedef pcl::PointCloud<pcl::PointXYZRGB> PCXYZRGB;
typedef pcl::PointCloud<pcl::PointXYZ>::Ptr PCXYZPtr;
typedef pcl::PointCloud<pcl::PointXYZ> PCXYZ;

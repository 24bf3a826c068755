// voxel_segmentation.h (drop-in) — pcl::VoxelBasedSegmentation<PointT> with the reference's public
// member names and call order (reference voxel_segmentation.h:84-421, 947-1014; driven as in
// test:51-76), every stage forwarded to the C ABI of libvgs_b200.so (include/vgs_b200.h).
// The reference signals no errors (all stage methods return void); failures of the CUDA path
// throw std::runtime_error with vgs_last_error() — there is no CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vgs_b200.h"
#include "pcl_shim.h"

namespace pcl {

template <typename PointT>
class VoxelBasedSegmentation {
 public:
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;

  // VS.h:84 — the resolution arrives as double(float voxel_size) (test:26,51)
  explicit VoxelBasedSegmentation(double input_resolution, int device = 0) : resolution_(input_resolution) {
    vgs_config cfg{};
    cfg.mode = VGS_MODE_VGS;
    cfg.device = device;
    cfg.leaf_order = VGS_LEAF_DESCENDING;  // PCL 1.8.1 leaf iterator (README.md:9 of the reference)
    if (vgs_create(&h_, &cfg) != VGS_OK) throw std::runtime_error(std::string("vgs_create: ") + vgs_last_error(nullptr));
  }
  ~VoxelBasedSegmentation() { vgs_destroy(h_); }
  VoxelBasedSegmentation(const VoxelBasedSegmentation&) = delete;
  VoxelBasedSegmentation& operator=(const VoxelBasedSegmentation&) = delete;

  // inherited from pcl::octree::OctreePointCloud in the reference (test:52,54,56)
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void addPointsFromInputCloud() {
    if (!input_ || input_->points.empty()) throw std::runtime_error("addPointsFromInputCloud: no input cloud");
    static_assert(sizeof(PointT) % 4 == 0, "point type must be float-aligned");
    ck(vgs_set_points(h_, &input_->points[0].x, (int64_t)input_->points.size(), (int)sizeof(PointT), 0));
    ck(vgs_voxelize(h_, (float)resolution_));
  }
  void getBoundingBox(double& min_x, double& min_y, double& min_z, double& max_x, double& max_y, double& max_z) {
    double b[6];
    ck(vgs_get_bounding_box(h_, b));
    min_x = b[0]; min_y = b[1]; min_z = b[2]; max_x = b[3]; max_y = b[4]; max_z = b[5];
  }

  int getCloudPointNum(PCXYZPtr input_data) {  // VS.h:94
    points_num_ = (int)input_data->points.size();
    points_cloud_ = input_data;
    return points_num_;
  }
  int getVoxelNum() {  // VS.h:104
    int64_t v = 0;
    ck(vgs_voxel_count(h_, &v));
    voxels_num_ = (int)v;
    return voxels_num_;
  }
  int getClusterNum() { return clusters_num_; }                            // VS.h:111
  std::vector<std::vector<int>> getClusterIdx() { return clusters_point_idx_; }  // VS.h:117

  void setVoxelSize(double input_resolution, int points_num_min, int voxels_num_min, int voxels_adj_min) {  // VS.h:124
    voxel_resolution_ = (float)input_resolution;
    voxel_points_min_ = points_num_min;
    cluster_voxels_min_ = voxels_num_min;
    voxel_adjacency_min_ = voxels_adj_min;
  }
  void setBoundingBox(double min_x, double min_y, double min_z, double max_x, double max_y, double max_z) {  // VS.h:133
    double b[6] = {min_x, min_y, min_z, max_x, max_y, max_z};
    ck(vgs_set_bounding_box(h_, b));
  }
  void setVoxelCenters() {}  // VS.h:146: the voxel table already exists after addPointsFromInputCloud
  std::vector<pcl::PointXYZ, Eigen::aligned_allocator<pcl::PointXYZ>> getVoxelCenters() {  // VS.h:191
    int64_t v = 0;
    ck(vgs_voxel_count(h_, &v));
    std::vector<float> c((size_t)v * 3);
    if (v) ck(vgs_get_voxel_centers(h_, c.data()));
    std::vector<pcl::PointXYZ, Eigen::aligned_allocator<pcl::PointXYZ>> out((size_t)v);
    for (int64_t i = 0; i < v; i++) { out[i].x = c[3 * i]; out[i].y = c[3 * i + 1]; out[i].z = c[3 * i + 2]; }
    return out;
  }

  void calcualteVoxelCloudAttributes(PCXYZPtr /*input_cloud*/) { ck(vgs_compute_features(h_, voxel_points_min_)); }  // VS.h:290 [sic]
  void findAllVoxelAdjacency(float graph_size) { ck(vgs_find_adjacency(h_, graph_size)); }                            // VS.h:223
  void segmentVoxelCloudWithGraphModel(float cut_thred, float sig_p, float sig_n, float sig_o, float sig_e, float sig_c,
                                       float sig_w) {  // VS.h:372
    vgs_sigmas s{sig_p, sig_n, sig_o, sig_e, sig_c, sig_w};
    ck(vgs_segment(h_, &s, cut_thred, voxel_adjacency_min_));
    int64_t all = 0, exported = 0;
    ck(vgs_cluster_count(h_, cluster_voxels_min_, &all, &exported));
    clusters_num_ = (int)all;  // VS.h:2084 counts every cluster, singletons included
  }

  // VS.h:947 "This is obligatory!": builds the per-cluster point-index lists (clusters with more than
  // voxels_min voxels) and the coloured cloud.  Colours are a deterministic hash of the cluster
  // index (the reference uses rand() seeded with time(0)).
  void drawColorMapofPointsinClusters(pcl::PointCloud<pcl::PointXYZRGB>::Ptr output_cloud) {
    int64_t nc = 0, nt = 0;
    ck(vgs_get_clusters_csr(h_, cluster_voxels_min_, &nc, &nt, nullptr, nullptr));
    std::vector<int64_t> off((size_t)nc + 1);
    std::vector<int32_t> idx((size_t)(nt > 0 ? nt : 1));
    ck(vgs_get_clusters_csr(h_, cluster_voxels_min_, &nc, &nt, off.data(), idx.data()));
    clusters_point_idx_.assign((size_t)nc, std::vector<int>());
    for (int64_t c = 0; c < nc; c++) {
      clusters_point_idx_[c].assign(idx.begin() + off[c], idx.begin() + off[c + 1]);
      if (output_cloud && points_cloud_) {
        uint32_t hsh = (uint32_t)c * 2654435761u;
        for (int p : clusters_point_idx_[c]) {
          pcl::PointXYZRGB q;
          q.x = points_cloud_->points[p].x; q.y = points_cloud_->points[p].y; q.z = points_cloud_->points[p].z;
          q.r = (uint8_t)(hsh >> 8); q.g = (uint8_t)(hsh >> 16); q.b = (uint8_t)(hsh >> 24);
          output_cloud->push_back(q);
        }
      }
    }
  }

  // canonical per-point labels (not in the reference: smallest point index of the point's cluster)
  std::vector<int> getPointLabels() {
    std::vector<int> lab((size_t)points_num_);
    ck(vgs_get_point_labels(h_, cluster_voxels_min_, lab.data(), 0));
    return lab;
  }
  vgs_handle handle() { return h_; }

 private:
  void ck(vgs_status s) { if (s != VGS_OK) throw std::runtime_error(std::string("libvgs_b200: ") + vgs_last_error(h_)); }
  vgs_handle h_ = nullptr;
  double resolution_;
  PointCloudConstPtr input_;
  PCXYZPtr points_cloud_;
  int points_num_ = 0, voxels_num_ = 0, clusters_num_ = 0;
  int voxel_points_min_ = 0, voxel_adjacency_min_ = 0, cluster_voxels_min_ = 0;
  float voxel_resolution_ = 0;
  std::vector<std::vector<int>> clusters_point_idx_;
};

}  // namespace pcl

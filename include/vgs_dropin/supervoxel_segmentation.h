// supervoxel_segmentation.h (drop-in) — pcl::SuperVoxelBasedSegmentation<PointT> with the reference's
// public member names and call order (reference supervoxel_segmentation.h:85-421, 613, driven as in
// test:138-160), forwarded to the C ABI of libvgs_b200.so in SVGS mode.
//
// The reference obtains its supervoxels from PCL's VCCS (pcl::SupervoxelClustering, SV.h:265-284 — third-party
// code).  Here createSupervoxels runs the library's own restatement of that generator
// (vgs_make_supervoxels_vccs) unless per-point labels were supplied (setSupervoxelLabels: what
// getLabeledCloud()/getMaxLabel() returned, SV.h:283-284) or the plain seed grid was asked for
// (useSeedGridSupervoxels).  Everything downstream of the labels is the reference's algorithm on the GPU.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vgs_b200.h"
#include <algorithm>
#include "host_parallel.h"
#include "mesh_export.h"
#include "pcl_shim.h"

namespace pcl {

template <typename PointT>
class SuperVoxelBasedSegmentation {
 public:
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;

  explicit SuperVoxelBasedSegmentation(double input_resolution, int device = 0) : resolution_(input_resolution) {  // SV.h:85
    vgs_config cfg{};
    cfg.mode = VGS_MODE_SVGS;
    cfg.device = device;
    cfg.leaf_order = VGS_LEAF_DESCENDING;
    if (vgs_acquire(&h_, &cfg) != VGS_OK) throw std::runtime_error(std::string("vgs_acquire: ") + vgs_last_error(nullptr));
  }
  ~SuperVoxelBasedSegmentation() { vgs_release(h_); }
  SuperVoxelBasedSegmentation(const SuperVoxelBasedSegmentation&) = delete;
  SuperVoxelBasedSegmentation& operator=(const SuperVoxelBasedSegmentation&) = delete;

  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }            // PCL, test:139
  void addPointsFromInputCloud() {                                                   // PCL, test:141
    if (!input_ || input_->points.empty()) throw std::runtime_error("addPointsFromInputCloud: no input cloud");
    ck(vgs_set_points(h_, &input_->points[0].x, (int64_t)input_->points.size(), (int)sizeof(PointT), 0));
    ck(vgs_voxelize(h_, (float)resolution_));
  }
  void getBoundingBox(double& min_x, double& min_y, double& min_z, double& max_x, double& max_y, double& max_z) {
    double b[6];
    ck(vgs_get_bounding_box(h_, b));
    min_x = b[0]; min_y = b[1]; min_z = b[2]; max_x = b[3]; max_y = b[4]; max_z = b[5];
  }
  void getTaskVector(std::vector<std::string> input_vector) { task_vector_ = input_vector; }   // SV.h:96 (stored, never read)
  int getCloudPointNum(PCXYZPtr input_data) {                                        // SV.h:101
    points_num_ = (int)input_data->points.size();
    points_cloud_ = input_data;
    return points_num_;
  }
  int getVoxelNum() { int64_t v = 0; ck(vgs_voxel_count(h_, &v)); voxels_num_ = (int)v; return voxels_num_; }  // SV.h:111
  int getSuperVoxelNum() { return supervoxels_num_; }                                // SV.h:118
  int getClusterNum() { return clusters_num_; }                                      // SV.h:124
  std::vector<std::vector<int>> getClusterIdx() { return vgs_dropin::copy_lists(clusters_point_idx_); }   // SV.h:130 (by value)

  void setVoxelSize(double input_resolution, int points_num_min) {                   // SV.h:143
    voxel_resolution_ = (float)input_resolution; voxel_points_min_ = points_num_min;
  }
  void setSupervoxelSize(double input_resolution, int voxels_num_min, int points_num_min, int adjacency_num_min) {  // SV.h:150
    seed_resolution_ = (float)input_resolution; supervoxel_voxel_min_ = voxels_num_min;
    supervoxel_point_min_ = points_num_min; supervoxel_adjacency_min_ = adjacency_num_min;
  }
  void setGraphSize(double small_resolution, double large_resolution) {              // SV.h:159
    graph_resolution_ = (float)large_resolution; adjacent_resolution_ = (float)small_resolution;
  }
  void setBoundingBox(double min_x, double min_y, double min_z, double max_x, double max_y, double max_z) {  // SV.h:166
    double b[6] = {min_x, min_y, min_z, max_x, max_y, max_z};
    ck(vgs_set_bounding_box(h_, b));
  }
  void setSupervoxelCentersCentroids() {}   // SV.h:178: its results are never consumed by the SVGS path

  // extension: the labels pcl::SupervoxelClustering::getLabeledCloud() / getMaxLabel() gave (SV.h:283-284)
  void setSupervoxelLabels(const std::vector<int>& label_per_point, int max_label) {
    labels_ = label_per_point; max_label_ = max_label; have_labels_ = true;
  }

  // extension: one supervoxel per occupied seed_resolution grid cell instead of VCCS
  void useSeedGridSupervoxels(bool on) { seed_grid_ = on; }

  // SV.h:245 createSupervoxels: VCCS with the colour / spatial / normal importances set by
  // segmentSupervoxelCloudWithGraphModel (SV.h:366-369) and refineSupervoxels(5) (SV.h:278)
  void createSupervoxels() {
    if (have_labels_) {
      if ((int)labels_.size() != points_num_) throw std::runtime_error("setSupervoxelLabels: one label per input point expected");
      ck(vgs_set_supervoxel_labels(h_, labels_.data(), max_label_, 0));
    } else if (seed_grid_) {
      ck(vgs_make_supervoxels_grid(h_, seed_resolution_));
    } else {
      ck(vgs_make_supervoxels_vccs(h_, seed_resolution_, color_impt_, spatial_impt_, normal_impt_, 5));
    }
  }
  // labels of the supervoxel generator (getLabeledCloud / getMaxLabel, SV.h:283-284); not available for the seed grid
  std::vector<int> getSupervoxelLabels(int* max_label = nullptr) {
    std::vector<int> lab((size_t)points_num_);
    int32_t ml = 0;
    ck(vgs_get_supervoxel_labels(h_, lab.data(), &ml, 0));
    if (max_label) *max_label = ml;
    return lab;
  }

  // SV.h:362.  sig_a/sig_b/sig_l are the VCCS colour/spatial/normal importances (SV.h:366-369).
  void segmentSupervoxelCloudWithGraphModel(float sig_a, float sig_b, float sig_l, float cut_thred, float sig_p,
                                            float sig_n, float sig_o, float sig_e, float sig_c, float sig_w) {
    color_impt_ = sig_a; spatial_impt_ = sig_b; normal_impt_ = sig_l;
    createSupervoxels();                                       // SV.h:372
    ck(vgs_compute_features(h_, supervoxel_point_min_));       // calcualteSupervoxelCloudAttributes SV.h:1238
    int64_t nsv = 0;
    ck(vgs_unit_count(h_, &nsv));
    supervoxels_num_ = (int)nsv;
    ck(vgs_find_adjacency(h_, graph_resolution_));             // findAllSupervoxelNeighbors SV.h:1477
    vgs_sigmas s{sig_p, sig_n, sig_o, sig_e, sig_c, sig_w};
    ck(vgs_segment(h_, &s, cut_thred, supervoxel_adjacency_min_));
    // clusteringSupervoxels builds the point lists right away (SV.h:2109-2126), all clusters, no filter
    int64_t nc = 0, nt = 0;
    ck(vgs_get_clusters_csr(h_, 0, &nc, &nt, nullptr, nullptr));
    std::vector<int64_t> off((size_t)nc + 1);
    std::unique_ptr<int32_t[]> idx(new int32_t[(size_t)(nt > 0 ? nt : 1)]);
    ck(vgs_get_clusters_csr(h_, 0, &nc, &nt, off.data(), idx.get()));
    vgs_dropin::csr_to_lists(off, idx.get(), clusters_point_idx_);
    clusters_num_ = (int)nc;
  }

  void drawColorMapofPointsinClusters(pcl::PointCloud<pcl::PointXYZRGB>::Ptr output_cloud) {   // SV.h:613
    if (!output_cloud || !points_cloud_) return;
    std::vector<size_t> first(clusters_point_idx_.size() + 1, 0);
    for (size_t c = 0; c < clusters_point_idx_.size(); c++) first[c + 1] = first[c] + clusters_point_idx_[c].size();
    const size_t at = output_cloud->points.size();
    output_cloud->points.reserve(at + first.back());
    vgs_dropin::advise_huge(output_cloud->points.data(), (at + first.back()) * sizeof(pcl::PointXYZRGB));
    output_cloud->points.resize(at + first.back());
    pcl::PointXYZRGB* dst = output_cloud->points.data() + at;
    const auto& src = points_cloud_->points;
    vgs_dropin::parallel_blocks(clusters_point_idx_.size(), 64, [&](size_t b, size_t e) {
      for (size_t c = b; c < e; c++) {
        const uint32_t hsh = (uint32_t)c * 2654435761u;
        size_t k = first[c];
        for (int p : clusters_point_idx_[c]) {
          pcl::PointXYZRGB& q = dst[k++];
          q.x = src[(size_t)p].x; q.y = src[(size_t)p].y; q.z = src[(size_t)p].z;
          q.r = (uint8_t)(hsh >> 8); q.g = (uint8_t)(hsh >> 16); q.b = (uint8_t)(hsh >> 24);
        }
      }
    });
    output_cloud->width = (std::uint32_t)output_cloud->points.size(); output_cloud->height = 1;
  }
  // ---- display exports (SV.h:424-611); colours are vgs_dropin::color_of(index), see mesh_export.h ----

  // SV.h:424: one stick centroid -> centroid + seed_resolution * normal per supervoxel with more than points_min points
  void drawNormofVoxels(pcl::PolygonMesh::Ptr output_mesh) {
    const std::vector<float> rec = vgs_dropin::fetch<float>(h_, VGS_BLOB_RECORDS);
    const std::vector<int64_t> off = vgs_dropin::fetch<int64_t>(h_, VGS_BLOB_UNIT_OFFSETS);
    pcl::PointCloud<pcl::PointXYZRGB> verts;
    uint8_t r, g, b;
    vgs_dropin::color_of(0u, r, g, b);
    for (int64_t i = 0; i + 1 < (int64_t)off.size(); i++) {
      if (!((int)(off[i + 1] - off[i]) > supervoxel_point_min_)) continue;
      const float* c = &rec[16 * i];
      const float* nrm = c + 3;
      verts.points.push_back(vgs_dropin::vertex(c[0], c[1], c[2], r, g, b));
      verts.points.push_back(vgs_dropin::vertex(c[0] + seed_resolution_ * nrm[0], c[1] + seed_resolution_ * nrm[1],
                                                c[2] + seed_resolution_ * nrm[2], r, g, b));
      // the reference indexes the two vertices with the SUPERVOXEL id i (SV.h:489), which points past the
      // vertex array as soon as one supervoxel was skipped; the drop-in indexes the stick it just wrote
      const uint32_t k = (uint32_t)(verts.points.size() / 2 - 1);
      vgs_dropin::push_poly(*output_mesh, k * 2, k * 2 + 1, k * 2);
    }
    pcl::toPCLPointCloud2(verts, output_mesh->cloud);
  }
  // SV.h:500: every point, one colour per octree voxel, voxels in leaf-iterator order (descending x-major Morton key)
  void drawColorMapofPointsinVoxels(pcl::PointCloud<pcl::PointXYZRGB>::Ptr output_cloud) {
    const std::vector<uint32_t> key = vgs_dropin::fetch<uint32_t>(h_, VGS_BLOB_POINT_KEY);
    std::vector<std::pair<uint64_t, int32_t>> order;
    order.reserve(key.size() / 3);
    for (size_t p = 0; p < key.size() / 3; p++) {
      if (key[3 * p] == 0xFFFFFFFFu) continue;   // non-finite point: never inserted into the octree
      uint64_t code = 0;
      for (int bit = 20; bit >= 0; bit--)
        code = (code << 3) | (uint64_t)(((key[3 * p] >> bit) & 1u) << 2 | ((key[3 * p + 1] >> bit) & 1u) << 1 | ((key[3 * p + 2] >> bit) & 1u));
      order.emplace_back(~code, (int32_t)p);
    }
    std::sort(order.begin(), order.end());
    uint32_t voxel = 0;
    for (size_t j = 0; j < order.size(); j++) {
      if (j && order[j].first != order[j - 1].first) voxel++;
      uint8_t r, g, b;
      vgs_dropin::color_of(voxel, r, g, b);
      const pcl::PointXYZ& q = points_cloud_->points[order[j].second];
      output_cloud->points.push_back(vgs_dropin::vertex(q.x, q.y, q.z, r, g, b));
    }
    output_cloud->width = (uint32_t)output_cloud->points.size();
    output_cloud->height = 1;
  }
  // SV.h:560: every labelled point, one colour per supervoxel, supervoxels in id order
  void drawColorMapofPointsinSupervoxels(pcl::PointCloud<pcl::PointXYZRGB>::Ptr output_cloud) {
    const std::vector<int64_t> off = vgs_dropin::fetch<int64_t>(h_, VGS_BLOB_UNIT_OFFSETS);
    const std::vector<int32_t> pts = vgs_dropin::fetch<int32_t>(h_, VGS_BLOB_UNIT_POINTS);
    for (int64_t i = 0; i + 1 < (int64_t)off.size(); i++) {
      uint8_t r, g, b;
      vgs_dropin::color_of((uint32_t)i, r, g, b);
      for (int64_t j = off[i]; j < off[i + 1]; j++) {
        const pcl::PointXYZ& q = points_cloud_->points[pts[j]];
        output_cloud->points.push_back(vgs_dropin::vertex(q.x, q.y, q.z, r, g, b));
      }
    }
    output_cloud->width = (uint32_t)output_cloud->points.size();
    output_cloud->height = 1;
  }

  std::vector<int> getPointLabels() {
    std::vector<int> lab((size_t)points_num_);
    ck(vgs_get_point_labels(h_, 0, lab.data(), 0));
    return lab;
  }
  vgs_handle handle() { return h_; }

 private:
  void ck(vgs_status s) { if (s != VGS_OK) throw std::runtime_error(std::string("libvgs_b200: ") + vgs_last_error(h_)); }
  vgs_handle h_ = nullptr;
  double resolution_;
  PointCloudConstPtr input_;
  PCXYZPtr points_cloud_;
  std::vector<int> labels_;
  std::vector<std::string> task_vector_;
  int max_label_ = 0;
  bool have_labels_ = false, seed_grid_ = false;
  float color_impt_ = 0, spatial_impt_ = 0, normal_impt_ = 0;
  int points_num_ = 0, voxels_num_ = 0, supervoxels_num_ = 0, clusters_num_ = 0;
  int voxel_points_min_ = 0, supervoxel_voxel_min_ = 0, supervoxel_point_min_ = 0, supervoxel_adjacency_min_ = 0;
  float voxel_resolution_ = 0, seed_resolution_ = 0, graph_resolution_ = 0, adjacent_resolution_ = 0;
  std::vector<std::vector<int>> clusters_point_idx_;
};

}  // namespace pcl

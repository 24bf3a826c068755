// voxel_segmentation.h (drop-in) — pcl::VoxelBasedSegmentation<PointT> with the reference's public
// member names and call order (reference voxel_segmentation.h:84-421, 947-1014; driven as in
// test:51-76), every stage forwarded to the C ABI of libvgs_b200.so (include/vgs_b200.h).
// The reference signals no errors (all stage methods return void); failures of the CUDA path
// throw std::runtime_error with vgs_last_error() — there is no CPU fallback.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vgs_b200.h"
#include "host_parallel.h"
#include "mesh_export.h"
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
    // pooled: the reference makes one object per cloud (test:51); the device working set is kept between objects
    if (vgs_acquire(&h_, &cfg) != VGS_OK) throw std::runtime_error(std::string("vgs_acquire: ") + vgs_last_error(nullptr));
  }
  ~VoxelBasedSegmentation() { vgs_release(h_); }
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
  std::vector<std::vector<int>> getClusterIdx() { return vgs_dropin::copy_lists(clusters_point_idx_); }  // VS.h:117 (by value)

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

  // VS.h:198, 211: internal bookkeeping of calcualteVoxelCloudAttributes in the reference (they APPEND one flag per call,
  // VS.h:351-361); the used / clustered state lives on the device here, so calling them from outside has no effect
  void setVoxelChosen(int /*voxel_idx*/, int /*chosen_ornot*/) {}
  void setVoxelClustered(int /*voxel_idx*/, int /*clustered_ornot*/) {}
  // VS.h:269: ids of the voxels within graph_size of voxel_id, nearest first, the voxel itself first (FLANN radius order)
  std::vector<int> getOneVoxelAdjacency(int voxel_id) {
    int n = 0;
    std::vector<int> ids(256);                       // a voxel has at most 255 neighbours (vgs_find_adjacency fails beyond)
    ck(vgs_get_unit_adjacency(h_, voxel_id, ids.data(), (int)ids.size(), &n));
    ids.resize((size_t)n);
    return ids;
  }

  void calcualteVoxelCloudAttributes(PCXYZPtr /*input_cloud*/) { ck(vgs_compute_features(h_, voxel_points_min_)); }  // VS.h:290 [sic]
  void findAllVoxelAdjacency(float graph_size) { ck(vgs_find_adjacency(h_, graph_size)); }                           // VS.h:223
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
    std::unique_ptr<int32_t[]> idx(new int32_t[(size_t)(nt > 0 ? nt : 1)]);     // not zero-filled: the copy writes every entry
    ck(vgs_get_clusters_csr(h_, cluster_voxels_min_, &nc, &nt, off.data(), idx.get()));
    vgs_dropin::csr_to_lists(off, idx.get(), clusters_point_idx_);
    if (!(output_cloud && points_cloud_)) return;
    // the coloured cloud: clustered point k of the CSR -> output point at + k (cluster after cluster, as the reference
    // pushes them); the gather from the input cloud runs on a few host threads
    const size_t at = output_cloud->points.size();
    output_cloud->points.reserve(at + (size_t)nt);    // one allocation for all clustered points
    vgs_dropin::advise_huge(output_cloud->points.data(), (at + (size_t)nt) * sizeof(pcl::PointXYZRGB));
    output_cloud->points.resize(at + (size_t)nt);
    pcl::PointXYZRGB* dst = output_cloud->points.data() + at;
    const auto& src = points_cloud_->points;
    vgs_dropin::parallel_blocks((size_t)nt, (size_t)1 << 16, [&](size_t b, size_t e) {
      size_t c = (size_t)(std::upper_bound(off.begin(), off.end(), (int64_t)b) - off.begin()) - 1;
      for (size_t k = b; k < e;) {
        while ((size_t)off[c + 1] <= k) c++;
        const size_t stop = std::min(e, (size_t)off[c + 1]);
        const uint32_t hsh = (uint32_t)c * 2654435761u;
        const uint8_t r = (uint8_t)(hsh >> 8), g = (uint8_t)(hsh >> 16), bl = (uint8_t)(hsh >> 24);
        for (; k < stop; k++) {
          const auto& p = src[(size_t)idx[(std::ptrdiff_t)k]];
          pcl::PointXYZRGB& q = dst[k];
          q.x = p.x; q.y = p.y; q.z = p.z; q.r = r; q.g = g; q.b = bl;
        }
      }
    });
    output_cloud->width = (std::uint32_t)output_cloud->points.size(); output_cloud->height = 1;
  }

  // ---- display exports (VS.h:424-945, 1016-1104).  Voxels are visited in voxel-id order (= the leaf
  // iterator's order); colours are vgs_dropin::color_of(voxel or cluster index), see mesh_export.h. ----

  // VS.h:424: the points of every used voxel (more than points_min points), one colour per voxel
  void drawColorMapofPointsinVoxels(pcl::PointCloud<pcl::PointXYZRGB>::Ptr output_cloud) {
    const VoxelTable t = voxelTable();
    for (int64_t v = 0; v < t.n; v++) {
      if (!(t.size(v) > voxel_points_min_)) continue;
      uint8_t r, g, b;
      vgs_dropin::color_of((uint32_t)v, r, g, b);
      for (int64_t j = t.off[v]; j < t.off[v + 1]; j++) {
        const pcl::PointXYZ& q = points_cloud_->points[t.pts[j]];
        output_cloud->points.push_back(vgs_dropin::vertex(q.x, q.y, q.z, r, g, b));
      }
    }
    output_cloud->width = (uint32_t)output_cloud->points.size();
    output_cloud->height = 1;
  }
  // VS.h:510: one coloured box (8 vertices, 12 triangles) per used voxel
  void drawColorMapofVoxels(pcl::PolygonMesh::Ptr output_mesh) { boxesOfUsedVoxels(*output_mesh, false); }
  // VS.h:794: the same boxes as wire frames (12 degenerate triangles a-b-a per box)
  void drawFrameMapofVoxels(pcl::PolygonMesh::Ptr output_mesh) { boxesOfUsedVoxels(*output_mesh, true); }
  // VS.h:653: boxes of the used voxels of every cluster, one colour per cluster, clusters in cluster order
  void drawColorMapofClusteredVoxels(pcl::PolygonMesh::Ptr output_mesh) {
    const VoxelTable t = voxelTable();
    const std::vector<int32_t> root = vgs_dropin::fetch<int32_t>(h_, VGS_BLOB_UNIT_ROOT);
    const std::vector<float> ctr = vgs_dropin::fetch<float>(h_, VGS_BLOB_UNIT_CENTER);   // voxel_centers_ (VS.h:690)
    // cluster m = m-th distinct root in ascending order; members in ascending voxel id
    std::vector<int32_t> cluster_of_root((size_t)t.n, -1);
    int32_t nc = 0;
    for (int64_t v = 0; v < t.n; v++) if (root[v] == v) cluster_of_root[v] = nc++;
    std::vector<int64_t> start((size_t)nc + 1, 0);
    for (int64_t v = 0; v < t.n; v++) start[cluster_of_root[root[v]] + 1]++;
    for (int32_t c = 0; c < nc; c++) start[c + 1] += start[c];
    std::vector<int32_t> member((size_t)t.n);
    { std::vector<int64_t> cur(start.begin(), start.end() - 1);
      for (int64_t v = 0; v < t.n; v++) member[cur[cluster_of_root[root[v]]]++] = (int32_t)v; }
    pcl::PointCloud<pcl::PointXYZRGB> verts;
    uint32_t box = 0;
    for (int32_t c = 0; c < nc; c++) {
      uint8_t r, g, b;
      vgs_dropin::color_of((uint32_t)c, r, g, b);
      for (int64_t j = start[c]; j < start[c + 1]; j++) {
        const int32_t v = member[j];
        if (!(t.size(v) > voxel_points_min_)) continue;   // voxel_used_ (VS.h:693)
        vgs_dropin::push_corners(verts, pcl::PointXYZ(ctr[3 * v], ctr[3 * v + 1], ctr[3 * v + 2]), voxel_resolution_, r, g, b);
        vgs_dropin::push_box_faces(*output_mesh, box++);
      }
    }
    pcl::toPCLPointCloud2(verts, output_mesh->cloud);
  }
  // VS.h:1016: one stick centre -> centre + voxel_resolution * normal per used voxel (degenerate triangle 0-1-0)
  void drawNormofVoxels(pcl::PolygonMesh::Ptr output_mesh) {
    const VoxelTable t = voxelTable();
    const std::vector<float> rec = vgs_dropin::fetch<float>(h_, VGS_BLOB_RECORDS);
    pcl::PointCloud<pcl::PointXYZRGB> verts;
    uint8_t r, g, b;
    vgs_dropin::color_of(0u, r, g, b);
    uint32_t i = 0;
    for (int64_t v = 0; v < t.n; v++) {
      if (!(t.size(v) > voxel_points_min_)) continue;
      const pcl::PointXYZ c = vgs_dropin::leaf_center(&t.key[3 * v], resolution_, t.box);
      const float* nrm = &rec[16 * v + 3];
      verts.points.push_back(vgs_dropin::vertex(c.x, c.y, c.z, r, g, b));
      verts.points.push_back(vgs_dropin::vertex(c.x + voxel_resolution_ * nrm[0], c.y + voxel_resolution_ * nrm[1],
                                                c.z + voxel_resolution_ * nrm[2], r, g, b));
      vgs_dropin::push_poly(*output_mesh, i * 2, i * 2 + 1, i * 2);
      i++;
    }
    pcl::toPCLPointCloud2(verts, output_mesh->cloud);
  }

  // canonical per-point labels (not in the reference: smallest point index of the point's cluster)
  std::vector<int> getPointLabels() {
    std::vector<int> lab((size_t)points_num_);
    ck(vgs_get_point_labels(h_, cluster_voxels_min_, lab.data(), 0));
    return lab;
  }
  vgs_handle handle() { return h_; }

 private:
  struct VoxelTable {   // voxel-id order: keys, point lists, PCL's bounding box
    int64_t n = 0;
    std::vector<uint32_t> key;
    std::vector<int64_t> off;
    std::vector<int32_t> pts;
    double box[6];
    int size(int64_t v) const { return (int)(off[v + 1] - off[v]); }
  };
  VoxelTable voxelTable() {
    VoxelTable t;
    ck(vgs_voxel_count(h_, &t.n));
    t.key = vgs_dropin::fetch<uint32_t>(h_, VGS_BLOB_UNIT_KEY);
    t.off = vgs_dropin::fetch<int64_t>(h_, VGS_BLOB_UNIT_OFFSETS);
    t.pts = vgs_dropin::fetch<int32_t>(h_, VGS_BLOB_UNIT_POINTS);
    ck(vgs_get_bounding_box(h_, t.box));
    return t;
  }
  void boxesOfUsedVoxels(pcl::PolygonMesh& mesh, bool frame) {
    const VoxelTable t = voxelTable();
    pcl::PointCloud<pcl::PointXYZRGB> verts;
    uint32_t i = 0;
    for (int64_t v = 0; v < t.n; v++) {
      if (!(t.size(v) > voxel_points_min_)) continue;
      uint8_t r, g, b;
      vgs_dropin::color_of((uint32_t)v, r, g, b);
      vgs_dropin::push_corners(verts, vgs_dropin::leaf_center(&t.key[3 * v], resolution_, t.box), voxel_resolution_, r, g, b);
      if (frame) vgs_dropin::push_box_edges(mesh, i); else vgs_dropin::push_box_faces(mesh, i);
      i++;
    }
    pcl::toPCLPointCloud2(verts, mesh.cloud);
  }
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

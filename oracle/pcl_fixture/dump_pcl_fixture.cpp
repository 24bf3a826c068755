// dump_pcl_fixture.cpp — PINNING KIT for the CPU oracle (test infrastructure; not part of the product).
//
// The oracle restates what PCL 1.8.1 does behind the reference's calls from memory: dynamic
// bounding box + octree keys (voxel_segmentation.h:84,146-189), leaf-iterator order (voxel ids), pcl::eigen33
// (voxel_segmentation.h:1436), FLANN radius search order (voxel_segmentation.h:223-265) and VCCS
// (supervoxel_segmentation.h:245-284).  PCL is not installable where the oracle was written, so parity is UNPINNED.
// Whoever has PCL 1.8.1 can pin it: build this program against PCL 1.8.1 (CMakeLists.txt next to it), run
//      ./dump_pcl_fixture tests/golden/pcl_fixture.bin
// and run `python -m pytest tests/test_pcl_fixture.py`: the test regenerates the same cloud (fixture_cloud.py, identical
// integer recurrence), runs the oracle and compares every section bit for bit.
//
// File layout (little endian): magic "VGSPCL01", then sections  { u32 tag, u64 bytes, payload }:
//   1 cloud        f32[N*3]            the generated points (cross-check of the generator)
//   2 bbox         f64[6]              getBoundingBox after addPointsFromInputCloud
//   3 point_key    u32[N*3]            OctreeKey of every point (genOctreeKeyforPoint against the FINAL box)
//   4 leaf_keys    u32[V*3]            leaf keys in leaf-iterator order (= voxel ids of the reference)
//   5 leaf_off     u32[V+1]  6 leaf_pts i32[...]   point indices of every leaf in container order
//   7 centers      f32[V*3]            getOccupiedVoxelCenters order / values
//   8 eigen_in     f32[M*9]  9 eigen_val f32[M*3]  10 eigen_vec f32[M*9]   pcl::eigen33 (Matrix3f, evecs column major)
//  11 adj_off      u32[V+1] 12 adj_idx  i32[...]  13 adj_d2 f32[...]       KdTreeFLANN::radiusSearch over the centres, r = 0.5
//  14 vccs_label   u32[N]   15 vccs_max u32[1]    SupervoxelClustering(0.05, 0.25), importances 0 / 0.25 / 0.75, refine 5
#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/common/eigen.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/octree/octree_pointcloud.h>
#include <pcl/octree/octree_iterator.h>
#include <pcl/segmentation/supervoxel_clustering.h>

#include <cstdint>
#include <cstdio>
#include <vector>

static FILE* g_out;
static void section(uint32_t tag, const void* p, uint64_t bytes) {
  fwrite(&tag, 4, 1, g_out); fwrite(&bytes, 8, 1, g_out); fwrite(p, 1, bytes, g_out);
}

int main(int argc, char** argv) {
  const int N = 20000;
  g_out = fopen(argc > 1 ? argv[1] : "pcl_fixture.bin", "wb");
  if (!g_out) return 1;
  fwrite("VGSPCL01", 1, 8, g_out);

  // ---- the cloud (same recurrence as fixture_cloud.py; doubles, then narrowed) ----
  std::vector<double> u((size_t)N * 5);
  uint32_t x = 12345u;
  for (size_t i = 0; i < u.size(); i++) { x = 1664525u * x + 1013904223u; u[i] = (double)x / 4294967296.0; }
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>);
  std::vector<float> raw((size_t)N * 3);
  for (int i = 0; i < N; i++) {
    const double a = u[5 * i], b = u[5 * i + 1], n1 = u[5 * i + 2], n2 = u[5 * i + 3], n3 = u[5 * i + 4];
    const double nx = 0.004 * (n1 + n2 - 1.0), ny = 0.004 * (n2 + n3 - 1.0), nz = 0.004 * (n3 + n1 - 1.0);
    double p[3];
    if (i % 3 == 0) { p[0] = 0.3 + 4.0 * a; p[1] = 0.2 + 3.0 * b; p[2] = 0.1; }
    else if (i % 3 == 1) { p[0] = 0.3 + 4.0 * a; p[1] = 3.2; p[2] = 0.1 + 2.5 * b; }
    else { p[0] = 4.3 + 2.0 * a; p[1] = 0.2 + 3.0 * b; p[2] = 0.1 + 1.1547 * a; }
    pcl::PointXYZ q((float)(p[0] + nx), (float)(p[1] + ny), (float)(p[2] + nz));
    cloud->push_back(q);
    raw[3 * i] = q.x; raw[3 * i + 1] = q.y; raw[3 * i + 2] = q.z;
  }
  section(1, raw.data(), raw.size() * 4);

  // ---- octree exactly as the reference builds it: OctreePointCloud(res), setInputCloud, addPointsFromInputCloud ----
  const double res = 0.15f;   // the reference passes a float (Task_File_VGS.txt: 0.15)
  typedef pcl::octree::OctreePointCloud<pcl::PointXYZ> Octree;
  Octree oct(res);
  oct.setInputCloud(cloud);
  oct.addPointsFromInputCloud();
  double bb[6];
  oct.getBoundingBox(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5]);
  section(2, bb, sizeof(bb));
  {
    std::vector<uint32_t> keys((size_t)N * 3);
    for (int i = 0; i < N; i++) {
      const pcl::PointXYZ& p = cloud->points[i];
      keys[3 * i] = (uint32_t)((p.x - bb[0]) / res);       // genOctreeKeyforPoint
      keys[3 * i + 1] = (uint32_t)((p.y - bb[1]) / res);
      keys[3 * i + 2] = (uint32_t)((p.z - bb[2]) / res);
    }
    section(3, keys.data(), keys.size() * 4);
  }
  std::vector<uint32_t> leaf_keys, leaf_off(1, 0);
  std::vector<int32_t> leaf_pts;
  for (Octree::LeafNodeIterator it = oct.leaf_begin(); it != oct.leaf_end(); ++it) {
    const pcl::octree::OctreeKey& k = it.getCurrentOctreeKey();
    leaf_keys.push_back(k.x); leaf_keys.push_back(k.y); leaf_keys.push_back(k.z);
    std::vector<int> idx;
    it.getLeafContainer().getPointIndices(idx);
    for (int p : idx) leaf_pts.push_back(p);
    leaf_off.push_back((uint32_t)leaf_pts.size());
  }
  section(4, leaf_keys.data(), leaf_keys.size() * 4);
  section(5, leaf_off.data(), leaf_off.size() * 4);
  section(6, leaf_pts.data(), leaf_pts.size() * 4);
  std::vector<pcl::PointXYZ, Eigen::aligned_allocator<pcl::PointXYZ> > centers;
  oct.getOccupiedVoxelCenters(centers);
  {
    std::vector<float> c(centers.size() * 3);
    for (size_t i = 0; i < centers.size(); i++) { c[3 * i] = centers[i].x; c[3 * i + 1] = centers[i].y; c[3 * i + 2] = centers[i].z; }
    section(7, c.data(), c.size() * 4);
  }

  // ---- pcl::eigen33 on scatter matrices of the first 64 leaves with > 3 points (float, the reference's call) ----
  {
    std::vector<float> in, val, vec;
    int made = 0;
    for (size_t l = 0; l + 1 < leaf_off.size() && made < 64; l++) {
      const int cnt = (int)(leaf_off[l + 1] - leaf_off[l]);
      if (cnt <= 3) continue;
      float mx = 0, my = 0, mz = 0;
      for (uint32_t t = leaf_off[l]; t < leaf_off[l + 1]; t++) { const pcl::PointXYZ& p = cloud->points[leaf_pts[t]]; mx += p.x; my += p.y; mz += p.z; }
      mx /= cnt; my /= cnt; mz /= cnt;
      Eigen::Matrix3f m = Eigen::Matrix3f::Zero();
      for (uint32_t t = leaf_off[l]; t < leaf_off[l + 1]; t++) {
        const pcl::PointXYZ& p = cloud->points[leaf_pts[t]];
        const float dx = p.x - mx, dy = p.y - my, dz = p.z - mz;
        m(0, 0) += dx * dx; m(0, 1) += dx * dy; m(0, 2) += dx * dz; m(1, 1) += dy * dy; m(1, 2) += dy * dz; m(2, 2) += dz * dz;
      }
      m(1, 0) = m(0, 1); m(2, 0) = m(0, 2); m(2, 1) = m(1, 2);
      Eigen::Matrix3f evecs; Eigen::Vector3f evals;
      pcl::eigen33(m, evecs, evals);
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) in.push_back(m(r, c));
      for (int r = 0; r < 3; r++) val.push_back(evals(r));
      for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) vec.push_back(evecs(r, c));
      made++;
    }
    section(8, in.data(), in.size() * 4); section(9, val.data(), val.size() * 4); section(10, vec.data(), vec.size() * 4);
  }

  // ---- FLANN radius search over the voxel centres (findAllVoxelAdjacency: KdTreeFLANN, radiusSearch(center, 0.5)) ----
  {
    pcl::PointCloud<pcl::PointXYZ>::Ptr cc(new pcl::PointCloud<pcl::PointXYZ>);
    for (size_t i = 0; i < centers.size(); i++) cc->push_back(centers[i]);
    pcl::KdTreeFLANN<pcl::PointXYZ> tree;
    tree.setInputCloud(cc);
    std::vector<uint32_t> off(1, 0);
    std::vector<int32_t> idx_all;
    std::vector<float> d2_all;
    for (size_t i = 0; i < cc->size(); i++) {
      std::vector<int> idx; std::vector<float> d2;
      tree.radiusSearch(cc->points[i], 0.5f, idx, d2);
      for (size_t t = 0; t < idx.size(); t++) { idx_all.push_back(idx[t]); d2_all.push_back(d2[t]); }
      off.push_back((uint32_t)idx_all.size());
    }
    section(11, off.data(), off.size() * 4); section(12, idx_all.data(), idx_all.size() * 4); section(13, d2_all.data(), d2_all.size() * 4);
  }

  // ---- VCCS as createSupervoxels calls it (supervoxel_segmentation.h:245-284) ----
  {
    pcl::SupervoxelClustering<pcl::PointXYZ> super(0.05f, 0.25f);
    super.setInputCloud(cloud);
    super.setColorImportance(0.0f);
    super.setSpatialImportance(0.25f);
    super.setNormalImportance(0.75f);
    std::map<uint32_t, pcl::Supervoxel<pcl::PointXYZ>::Ptr> clusters;
    super.extract(clusters);
    super.refineSupervoxels(5, clusters);
    pcl::PointCloud<pcl::PointXYZL>::Ptr labeled = super.getLabeledCloud();
    std::vector<uint32_t> lab(N, 0);
    for (int i = 0; i < N && i < (int)labeled->size(); i++) lab[i] = labeled->points[i].label;
    const uint32_t mx = (uint32_t)super.getMaxLabel();
    section(14, lab.data(), lab.size() * 4); section(15, &mx, 4);
  }
  fclose(g_out);
  printf("wrote %s: %d points, %zu leaves\n", argc > 1 ? argv[1] : "pcl_fixture.bin", N, leaf_off.size() - 1);
  return 0;
}

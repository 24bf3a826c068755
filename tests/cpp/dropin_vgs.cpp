// dropin_vgs.cpp — test program: drives pcl::VoxelBasedSegmentation through exactly the call
// sequence of the reference's usage snippet (reference file `test`, lines 51-76) using the drop-in
// header, on a raw float32 xyz file, and writes canonical labels for comparison with the oracle.
//   dropin_vgs <xyz.f32> <n> <labels.i32> voxel graph sig_p sig_n sig_o sig_e sig_c sig_w cut points_min adjacency_min voxels_min
// VGS_DROPIN_REPEAT=k: the whole sequence k times (a fresh object each time, destroyed inside the timed region); prints the
// wall-clock milliseconds of every run after the first and its phases (bench.py's e2e_dropin: pageable 16-byte-stride cloud
// in, coloured cloud + getClusterIdx() out).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vgs_dropin/voxel_segmentation.h"

int main(int argc, char** argv) {
  if (argc < 16) { fprintf(stderr, "usage\n"); return 2; }
  const char* in = argv[1];
  long n = atol(argv[2]);
  const char* out = argv[3];
  float voxel_size = (float)atof(argv[4]), graph_size = (float)atof(argv[5]);
  float sig_p = (float)atof(argv[6]), sig_n = (float)atof(argv[7]), sig_o = (float)atof(argv[8]), sig_e = (float)atof(argv[9]),
        sig_c = (float)atof(argv[10]), sig_w = (float)atof(argv[11]), cut_thred = (float)atof(argv[12]);
  int points_min = atoi(argv[13]), adjacency_min = atoi(argv[14]), voxels_min = atoi(argv[15]);

  PCXYZPtr input_cloud(new PCXYZ);
  {
    std::vector<float> buf((size_t)n * 3);
    FILE* f = fopen(in, "rb");
    if (!f || fread(buf.data(), 4, (size_t)n * 3, f) != (size_t)n * 3) { fprintf(stderr, "read failed\n"); return 3; }
    fclose(f);
    for (long i = 0; i < n; i++) input_cloud->push_back(pcl::PointXYZ(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]));
  }
  const char* rep_env = getenv("VGS_DROPIN_REPEAT");
  const int repeat = rep_env ? atoi(rep_env) : 1;
  try {
   for (int rep = 0; rep < repeat; rep++) {
    const auto t0 = std::chrono::steady_clock::now();
    auto last = t0;
    double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // voxelise (incl. H2D), centres, features, adjacency, segment, clusters + cloud, getClusterIdx, destruction
    auto lap = [&](int i) { const auto now = std::chrono::steady_clock::now(); ph[i] += std::chrono::duration<double, std::milli>(now - last).count(); last = now; };
    std::vector<int> lab;
    int nvox = 0, ncl = 0;
    size_t n_centers = 0, n_exported = 0, n_coloured = 0;
    {
      double min_x = 0, min_y = 0, min_z = 0, max_x = 0, max_y = 0, max_z = 0;
      pcl::PointCloud<pcl::PointXYZRGB>::Ptr clustered_cloud(new PCXYZRGB);
      std::vector<pcl::PointXYZ, Eigen::aligned_allocator<pcl::PointXYZ>> voxel_centers;

      // Voxelization (test:51-57)
      pcl::VoxelBasedSegmentation<pcl::PointXYZ> voxel_structure(voxel_size);
      voxel_structure.setInputCloud(input_cloud);
      voxel_structure.getCloudPointNum(input_cloud);
      voxel_structure.addPointsFromInputCloud();
      voxel_structure.setVoxelSize(voxel_size, points_min, voxels_min, adjacency_min);
      voxel_structure.getBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
      voxel_structure.setBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
      lap(0);
      // centres (test:60-62)
      voxel_structure.setVoxelCenters();
      voxel_centers = voxel_structure.getVoxelCenters();
      nvox = voxel_structure.getVoxelNum();
      lap(1);
      // features, adjacency, segmentation (test:65-71)
      voxel_structure.calcualteVoxelCloudAttributes(input_cloud);
      lap(2);
      voxel_structure.findAllVoxelAdjacency(graph_size);
      if (rep == 0) {   // VS.h:269: the radius neighbours of a voxel, itself first (not part of test:51-76: checked once, untimed runs only)
        std::vector<int> a0 = voxel_structure.getOneVoxelAdjacency(0), al = voxel_structure.getOneVoxelAdjacency(nvox - 1);
        if (a0.empty() || a0[0] != 0 || al.empty() || al[0] != nvox - 1) { fprintf(stderr, "getOneVoxelAdjacency: self is not first\n"); return 4; }
      }
      lap(3);
      voxel_structure.segmentVoxelCloudWithGraphModel(cut_thred, sig_p, sig_n, sig_o, sig_e, sig_c, sig_w);
      lap(4);
      // output (test:74-76)
      voxel_structure.drawColorMapofPointsinClusters(clustered_cloud);
      lap(5);
      std::vector<std::vector<int>> clusters_points_idx = voxel_structure.getClusterIdx();
      lap(6);
      n_centers = voxel_centers.size(); ncl = voxel_structure.getClusterNum();
      n_exported = clusters_points_idx.size(); n_coloured = clustered_cloud->size();
      if (rep + 1 == repeat) {
        lab.assign((size_t)n, -1);
        for (auto& cl : clusters_points_idx) {
          int mn = cl.empty() ? -1 : cl[0];
          for (int p : cl) mn = p < mn ? p : mn;
          for (int p : cl) lab[p] = mn;
        }
        last = std::chrono::steady_clock::now();
      }
    }   // the object, the coloured cloud and the lists are destroyed here, as at the end of segmentationVGS (test:86)
    lap(7);
    if (repeat > 1 && rep > 0) {
      double ms = 0;
      for (double v : ph) ms += v;
      printf("dropin_ms %.3f\n", ms);
      printf("dropin_phases %.3f %.3f %.3f %.3f %.3f %.3f %.3f %.3f\n", ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
    }
    if (rep + 1 < repeat) continue;
    FILE* f = fopen(out, "wb");
    fwrite(lab.data(), 4, (size_t)n, f);
    fclose(f);
    printf("voxels %d centers %zu clusters_all %d exported %zu coloured_points %zu\n", nvox, n_centers, ncl, n_exported, n_coloured);
   }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

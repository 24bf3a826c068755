// vgs_driver.cpp — the main() the reference never shipped: reads a Task_File_*.txt (parameter k = line k,
// IO.cpp:147-169), loads the PCD it names, dispatches on the "Method" line ([24]: 2 = VGS, 3 = SVGS,
// Task_File_VGS.txt:24-25) and runs segmentationVGS / segmentationSVGS exactly as the reference's usage
// snippet does (test:9-86, 91-170) on the drop-in classes.  Usage:
//   vgs_driver <task_file> [input.pcd] [output.pcd] [supervoxel_labels.i32]
// With VGS_DRIVER_EXPORT_DIR=<dir> the display exports the snippet declares but never saves (test:42-47, 131-135:
// coloured voxels, frames, normals, clustered voxels) are written there as PLY / PCD files.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "vgs_dropin/point_clouds_IO.h"
#include "vgs_dropin/supervoxel_segmentation.h"
#include "vgs_dropin/voxel_segmentation.h"

using std::string;

static int segmentationVGS(const string& out_name, PCXYZPtr input_cloud, const std::vector<string>& input_vector) {
  float voxel_size = (float)std::atof(input_vector[28].c_str());   // test:26-37
  float graph_size = (float)std::atof(input_vector[30].c_str());
  float sig_p = (float)std::atof(input_vector[32].c_str()), sig_n = (float)std::atof(input_vector[34].c_str());
  float sig_o = (float)std::atof(input_vector[36].c_str()), sig_e = (float)std::atof(input_vector[38].c_str());
  float sig_c = (float)std::atof(input_vector[40].c_str()), sig_w = (float)std::atof(input_vector[42].c_str());
  float cut_thred = (float)std::atof(input_vector[44].c_str());
  int points_min = std::atoi(input_vector[46].c_str()), adjacency_min = std::atoi(input_vector[48].c_str());
  int voxels_min = std::atoi(input_vector[50].c_str());
  double min_x = 0, min_y = 0, min_z = 0, max_x = 0, max_y = 0, max_z = 0;
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr clustered_cloud(new PCXYZRGB);

  pcl::VoxelBasedSegmentation<pcl::PointXYZ> voxel_structure(voxel_size);   // test:51-57
  voxel_structure.setInputCloud(input_cloud);
  voxel_structure.getCloudPointNum(input_cloud);
  voxel_structure.addPointsFromInputCloud();
  voxel_structure.setVoxelSize(voxel_size, points_min, voxels_min, adjacency_min);
  voxel_structure.getBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
  voxel_structure.setBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
  voxel_structure.setVoxelCenters();                                         // test:60-62
  voxel_structure.getVoxelCenters();
  int nv = voxel_structure.getVoxelNum();
  voxel_structure.calcualteVoxelCloudAttributes(input_cloud);               // test:65
  voxel_structure.findAllVoxelAdjacency(graph_size);                        // test:68
  voxel_structure.segmentVoxelCloudWithGraphModel(cut_thred, sig_p, sig_n, sig_o, sig_e, sig_c, sig_w);  // test:71
  voxel_structure.drawColorMapofPointsinClusters(clustered_cloud);          // test:74
  std::vector<std::vector<int>> clusters_points_idx = voxel_structure.getClusterIdx();
  std::printf(" In total %d\n", voxel_structure.getClusterNum());            // VS.h:2086
  std::printf("VGS: %zu points, %d voxels, %zu segments written\n", input_cloud->size(), nv, clusters_points_idx.size());
  saveColoredClusters(out_name, input_cloud, clusters_points_idx);          // test:80
  if (const char* dir = std::getenv("VGS_DRIVER_EXPORT_DIR")) {
    const string d = string(dir) + "/";
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr colored_cloud(new PCXYZRGB);                                   // test:42
    pcl::PolygonMesh::Ptr colored_voxels(new pcl::PolygonMesh), frames_voxels(new pcl::PolygonMesh);     // test:44-47
    pcl::PolygonMesh::Ptr Normes_voxels(new pcl::PolygonMesh), clustered_voxels(new pcl::PolygonMesh);
    voxel_structure.drawColorMapofPointsinVoxels(colored_cloud);
    voxel_structure.drawColorMapofVoxels(colored_voxels);
    voxel_structure.drawFrameMapofVoxels(frames_voxels);
    voxel_structure.drawNormofVoxels(Normes_voxels);
    voxel_structure.drawColorMapofClusteredVoxels(clustered_voxels);
    saveColoredClusters(d + "points_in_voxels.pcd", colored_cloud);
    if (vgs_dropin::savePolygonMeshPLY(d + "colored_voxels.ply", *colored_voxels) || vgs_dropin::savePolygonMeshPLY(d + "frames_voxels.ply", *frames_voxels) ||
        vgs_dropin::savePolygonMeshPLY(d + "normals_voxels.ply", *Normes_voxels) ||
        vgs_dropin::savePolygonMeshPLY(d + "clustered_voxels.ply", *clustered_voxels))
      throw std::runtime_error("cannot write the display exports to " + d);
    std::printf("display exports written to %s\n", d.c_str());
  }
  return 0;
}

static int segmentationSVGS(const string& out_name, PCXYZPtr input_cloud, const std::vector<string>& input_vector,
                            const std::vector<int>* labels) {
  float voxel_size = (float)std::atof(input_vector[28].c_str());   // test:108-125
  float seed_size = (float)std::atof(input_vector[30].c_str());
  float graph_size = (float)std::atof(input_vector[32].c_str());
  float sig_p = (float)std::atof(input_vector[34].c_str()), sig_n = (float)std::atof(input_vector[36].c_str());
  float sig_o = (float)std::atof(input_vector[38].c_str()), sig_e = (float)std::atof(input_vector[40].c_str());
  float sig_c = (float)std::atof(input_vector[42].c_str()), sig_w = (float)std::atof(input_vector[44].c_str());
  float sig_a = (float)std::atof(input_vector[46].c_str()), sig_b = (float)std::atof(input_vector[48].c_str());
  sig_c = (float)std::atof(input_vector[50].c_str());              // test:102-103,122: the second sig_c overwrites the first
  float cut_thred = (float)std::atof(input_vector[52].c_str());
  int points_min = std::atoi(input_vector[54].c_str());
  int voxels_min = std::atoi(input_vector[58].c_str()), adjacency_min = std::atoi(input_vector[60].c_str());
  double min_x = 0, min_y = 0, min_z = 0, max_x = 0, max_y = 0, max_z = 0;
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr clustered_cloud(new PCXYZRGB);

  pcl::SuperVoxelBasedSegmentation<pcl::PointXYZ> supervoxel_structure(voxel_size);   // test:138-142
  supervoxel_structure.setInputCloud(input_cloud);
  supervoxel_structure.getCloudPointNum(input_cloud);
  supervoxel_structure.addPointsFromInputCloud();
  supervoxel_structure.setVoxelSize(voxel_size, points_min);                          // test:144-146
  supervoxel_structure.setSupervoxelSize(seed_size, voxels_min, points_min, adjacency_min);
  supervoxel_structure.setGraphSize(seed_size * 2, graph_size);
  supervoxel_structure.getBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);      // test:148-153
  supervoxel_structure.setBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
  supervoxel_structure.setSupervoxelCentersCentroids();
  int nv = supervoxel_structure.getVoxelNum();
  if (labels) {
    int ml = 0;
    for (int l : *labels) ml = l > ml ? l : ml;
    supervoxel_structure.setSupervoxelLabels(*labels, ml + 1);
  }
  supervoxel_structure.segmentSupervoxelCloudWithGraphModel(sig_a, sig_b, sig_c, cut_thred, sig_p, sig_n, sig_o, sig_e, sig_c, sig_w);  // test:156
  supervoxel_structure.drawColorMapofPointsinClusters(clustered_cloud);               // test:159
  std::vector<std::vector<int>> clusters_points_idx = supervoxel_structure.getClusterIdx();
  std::printf("In total %d segments\n", supervoxel_structure.getClusterNum());         // SV.h:2129
  std::printf("SVGS: %zu points, %d voxels, %d supervoxels, %zu segments written\n", input_cloud->size(), nv,
              supervoxel_structure.getSuperVoxelNum(), clusters_points_idx.size());
  saveColoredClusters(out_name, input_cloud, clusters_points_idx);
  if (const char* dir = std::getenv("VGS_DRIVER_EXPORT_DIR")) {
    const string d = string(dir) + "/";
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr colored_voxels(new PCXYZRGB), colored_supervoxels(new PCXYZRGB);   // test:131-133
    pcl::PolygonMesh::Ptr normes_spvoxels(new pcl::PolygonMesh);                                              // test:135
    supervoxel_structure.drawColorMapofPointsinVoxels(colored_voxels);
    supervoxel_structure.drawColorMapofPointsinSupervoxels(colored_supervoxels);
    supervoxel_structure.drawNormofVoxels(normes_spvoxels);
    saveColoredClusters(d + "points_in_voxels.pcd", colored_voxels);
    saveColoredClusters(d + "points_in_supervoxels.pcd", colored_supervoxels);
    if (vgs_dropin::savePolygonMeshPLY(d + "normals_supervoxels.ply", *normes_spvoxels)) throw std::runtime_error("cannot write the display exports to " + d);
    std::printf("display exports written to %s\n", d.c_str());
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: vgs_driver <task_file> [input.pcd] [output.pcd] [labels.i32]\n"); return 2; }
  std::vector<string> input_vector = inputTaskTxtFile(argv[1]);
  if (input_vector.size() < 51) { std::fprintf(stderr, "task file too short (%zu lines)\n", input_vector.size()); return 2; }
  string in_name = argc > 2 ? argv[2] : vgs_rstrip(input_vector[12]) + vgs_rstrip(input_vector[15]);
  string out_name = argc > 3 ? argv[3] : vgs_rstrip(input_vector[18]) + vgs_rstrip(input_vector[21]);
  int method = std::atoi(input_vector[24].c_str());
  PCXYZPtr input_cloud(new PCXYZ);
  const bool is_ply = in_name.size() >= 4 && (in_name.compare(in_name.size() - 4, 4, ".ply") == 0 || in_name.compare(in_name.size() - 4, 4, ".PLY") == 0);
  int rc = is_ply ? inputPointCloudData2(in_name, input_cloud) : inputPointCloudData(in_name, input_cloud);   // IO.h:64 / IO.h:83
  if (rc != 0) { std::fprintf(stderr, "cannot read %s (rc %d)\n", in_name.c_str(), rc); return 3; }
  try {
    if (method == 2) return segmentationVGS(out_name, input_cloud, input_vector);
    if (method == 3) {
      std::vector<int> labels;
      if (argc > 4) {
        FILE* f = std::fopen(argv[4], "rb");
        if (!f) { std::fprintf(stderr, "cannot read labels\n"); return 3; }
        labels.resize(input_cloud->size());
        size_t got = std::fread(labels.data(), 4, labels.size(), f);
        std::fclose(f);
        if (got != labels.size()) { std::fprintf(stderr, "label file too short\n"); return 3; }
      }
      if (input_vector.size() < 61) { std::fprintf(stderr, "SVGS task file too short\n"); return 2; }
      return segmentationSVGS(out_name, input_cloud, input_vector, argc > 4 ? &labels : nullptr);
    }
    std::fprintf(stderr, "unknown method %d (2 = VGS, 3 = SVGS)\n", method);
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}

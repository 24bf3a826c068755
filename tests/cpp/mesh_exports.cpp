// mesh_exports.cpp — test program for the display exports of the two drop-in classes
// (reference voxel_segmentation.h:424-1104, supervoxel_segmentation.h:424-611): runs VGS and SVGS on
// a raw float32 xyz file and writes every draw* result (PLY meshes, raw coloured clouds).
//   mesh_exports <xyz.f32> <n> <outdir>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "vgs_dropin/supervoxel_segmentation.h"
#include "vgs_dropin/voxel_segmentation.h"

static void write_cloud(const std::string& path, const pcl::PointCloud<pcl::PointXYZRGB>& c) {
  FILE* f = fopen(path.c_str(), "wb");
  for (const auto& p : c.points) {
    float xyz[3] = {p.x, p.y, p.z};
    unsigned char rgb[4] = {p.r, p.g, p.b, 0};
    fwrite(xyz, 4, 3, f);
    fwrite(rgb, 1, 4, f);
  }
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
  long n = atol(argv[2]);
  std::string out = argv[3];
  PCXYZPtr cloud(new PCXYZ);
  {
    std::vector<float> buf((size_t)n * 3);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), 4, (size_t)n * 3, f) != (size_t)n * 3) { fprintf(stderr, "read failed\n"); return 3; }
    fclose(f);
    for (long i = 0; i < n; i++) cloud->push_back(pcl::PointXYZ(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]));
  }
  try {
    double b[6];
    {  // VGS, Task_File_VGS.txt values
      pcl::VoxelBasedSegmentation<pcl::PointXYZ> vs(0.15f);
      vs.setInputCloud(cloud);
      vs.getCloudPointNum(cloud);
      vs.addPointsFromInputCloud();
      vs.setVoxelSize(0.15f, 10, 3, 3);
      vs.getBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]);
      vs.setBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]);
      vs.setVoxelCenters();
      vs.getVoxelNum();
      vs.calcualteVoxelCloudAttributes(cloud);
      vs.findAllVoxelAdjacency(0.5f);
      vs.segmentVoxelCloudWithGraphModel(0.3f, 0.2f, 0.2f, 0.2f, 0.2f, 0.2f, 2.0f);
      pcl::PointCloud<pcl::PointXYZRGB>::Ptr pv(new PCXYZRGB);
      pcl::PolygonMesh::Ptr boxes(new pcl::PolygonMesh), frames(new pcl::PolygonMesh), normals(new pcl::PolygonMesh),
          clustered(new pcl::PolygonMesh);
      vs.drawColorMapofPointsinVoxels(pv);
      vs.drawColorMapofVoxels(boxes);
      vs.drawFrameMapofVoxels(frames);
      vs.drawNormofVoxels(normals);
      vs.drawColorMapofClusteredVoxels(clustered);
      write_cloud(out + "/vgs_points_in_voxels.bin", *pv);
      if (vgs_dropin::savePolygonMeshPLY(out + "/vgs_boxes.ply", *boxes) || vgs_dropin::savePolygonMeshPLY(out + "/vgs_frames.ply", *frames) ||
          vgs_dropin::savePolygonMeshPLY(out + "/vgs_normals.ply", *normals) || vgs_dropin::savePolygonMeshPLY(out + "/vgs_clustered.ply", *clustered))
        throw std::runtime_error("PLY write failed");
      printf("vgs boxes %zu frames %zu normals %zu clustered %zu points %zu\n", boxes->polygons.size(), frames->polygons.size(),
             normals->polygons.size(), clustered->polygons.size(), pv->size());
    }
    {  // SVGS, Task_File_SVGS.txt values, built-in seed-grid supervoxels
      pcl::SuperVoxelBasedSegmentation<pcl::PointXYZ> ss(0.05f);
      ss.setInputCloud(cloud);
      ss.getCloudPointNum(cloud);
      ss.addPointsFromInputCloud();
      ss.setVoxelSize(0.05f, 10);
      ss.setSupervoxelSize(0.25f, 3, 10, 3);
      ss.setGraphSize(0.05f, 0.5f);
      ss.getBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]);
      ss.setBoundingBox(b[0], b[1], b[2], b[3], b[4], b[5]);
      ss.segmentSupervoxelCloudWithGraphModel(0.0f, 0.25f, 0.75f, 0.5f, 0.2f, 0.2f, 0.2f, 0.2f, 0.2f, 1.0f);
      pcl::PointCloud<pcl::PointXYZRGB>::Ptr pv(new PCXYZRGB), ps(new PCXYZRGB);
      pcl::PolygonMesh::Ptr normals(new pcl::PolygonMesh);
      ss.drawColorMapofPointsinVoxels(pv);
      ss.drawColorMapofPointsinSupervoxels(ps);
      ss.drawNormofVoxels(normals);
      write_cloud(out + "/svgs_points_in_voxels.bin", *pv);
      write_cloud(out + "/svgs_points_in_supervoxels.bin", *ps);
      if (vgs_dropin::savePolygonMeshPLY(out + "/svgs_normals.ply", *normals)) throw std::runtime_error("PLY write failed");
      // supervoxel id per point of the built-in seed grid + the canonical labels, so the test can replay the oracle on them
      std::vector<int32_t> unit = vgs_dropin::fetch<int32_t>(ss.handle(), VGS_BLOB_POINT_UNIT);
      std::vector<int> lab = ss.getPointLabels();
      FILE* f = fopen((out + "/svgs_point_unit.i32").c_str(), "wb"); fwrite(unit.data(), 4, unit.size(), f); fclose(f);
      f = fopen((out + "/svgs_labels.i32").c_str(), "wb"); fwrite(lab.data(), 4, lab.size(), f); fclose(f);
      printf("svgs voxels %d supervoxels %d normals %zu points %zu %zu\n", ss.getVoxelNum(), ss.getSuperVoxelNum(), normals->polygons.size(),
             pv->size(), ps->size());
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

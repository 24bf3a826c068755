// pcd_roundtrip.cpp — test program for the drop-in I/O layer: reads a PCD (ascii / binary /
// binary_compressed) with inputPointCloudData — or, with a leading "ply", a PLY with inputPointCloudData2 — and dumps
// x y z as raw float32; "xyzpcd" re-writes the cloud with outputPointCloudData; "task" lists a task file.
#include <cstdio>
#include "vgs_dropin/point_clouds_IO.h"
int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "task") {
    std::vector<std::string> v = inputTaskTxtFile(argv[2]);
    std::printf("%zu\n", v.size());
    for (auto& l : v) std::printf("[%s]\n", vgs_rstrip(l).c_str());
    return 0;
  }
  if (argc < 3) return 2;
  PCXYZPtr cloud(new PCXYZ);
  if (argc >= 4 && std::string(argv[1]) == "ply") {
    int rc = inputPointCloudData2(argv[2], cloud);
    if (rc != 0) { std::fprintf(stderr, "rc %d\n", rc); return 3; }
    FILE* f = std::fopen(argv[3], "wb");
    for (auto& p : cloud->points) { float xyz[3] = {p.x, p.y, p.z}; std::fwrite(xyz, 4, 3, f); }
    std::fclose(f);
    std::printf("%zu\n", cloud->size());
    return 0;
  }
  if (argc >= 4 && std::string(argv[1]) == "xyzpcd") {
    int rc = inputPointCloudData(argv[2], cloud);
    if (rc != 0) { std::fprintf(stderr, "rc %d\n", rc); return 3; }
    return outputPointCloudData(argv[3], cloud) == 0 ? 0 : 4;
  }
  int rc = inputPointCloudData(argv[1], cloud);
  if (rc != 0) { std::fprintf(stderr, "rc %d\n", rc); return 3; }
  FILE* f = std::fopen(argv[2], "wb");
  for (auto& p : cloud->points) { float xyz[3] = {p.x, p.y, p.z}; std::fwrite(xyz, 4, 3, f); }
  std::fclose(f);
  std::printf("%zu\n", cloud->size());
  return 0;
}

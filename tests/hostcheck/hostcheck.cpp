// hostcheck.cpp — TEST-ONLY host build of the per-thread device functions in
// vgs_svgs_segmentation_b200/csrc/vgs_math.cuh, so their arithmetic can be compared with the
// oracle without a GPU (pytest -m "not gpu").  Never loaded by the product library.
#include "../../vgs_svgs_segmentation_b200/csrc/vgs_math.cuh"

extern "C" {
void hc_pair(const float* a, const float* b, const float* sig6, int svgs, float* out2) {
  vgs::PairParams P{sig6[0], sig6[1], sig6[2], sig6[3], sig6[4], sig6[5], svgs};
  vgs::pair_weights(a, b, P, out2[0], out2[1]);
}
void hc_unit_record(const float* xyz, int cnt, int used, int svgs, float* rec16) {
  vgs::unit_record([&](int j, float& x, float& y, float& z) { x = xyz[3 * j]; y = xyz[3 * j + 1]; z = xyz[3 * j + 2]; },
                   cnt, used != 0, svgs, rec16);
}
uint64_t hc_morton(uint32_t x, uint32_t y, uint32_t z) { return vgs::morton_encode(x, y, z); }
void hc_demorton(uint64_t m, uint32_t* xyz) { vgs::morton_decode(m, xyz[0], xyz[1], xyz[2]); }
}

/*
 * vgs_oracle.cpp — CPU ORACLE (restatement) of the VGS / SVGS segmentation hot path.
 *
 * TEST INFRASTRUCTURE ONLY — see vgs_oracle.h.  PARITY UNPINNED (no reference tests / golden
 * vectors exist; PCL/FLANN/Eigen are absent and the reference does not compile).
 *
 * Every function cites the reference lines it restates ("VS.h" = voxel_segmentation.h,
 * "SV.h" = supervoxel_segmentation.h, "test" = the file /root/reference/test).  Third-party
 * behaviour (PCL 1.8.1 octree / eigen33, FLANN radius search, Eigen reductions) is restated from
 * the published algorithms and isolated in pcl_* / flann_* functions.
 *
 * Single-threaded like the reference.  Build: g++ -O2 -std=c++17 -ffp-contract=off (no fast-math).
 *
 * Libm policy (vgso_params.math): the reference calls acos/sin/cos/atan2/log/pow on float
 * arguments, which resolve to the float overloads; those are not bit-reproducible between libms
 * (MSVC CRT of the original, glibc 2.39 here).  math=0 calls glibc's float functions (what a GCC
 * build of the reference would do), math=1 evaluates them correctly rounded (double evaluation,
 * one rounding) which is the libm-independent definition the CUDA path implements.
 */
#include "vgs_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#include <array>

static int g_threads = 1;   // vgso_set_threads

namespace {

// ------------------------------------------------------------------------------------------
// libm policies
// ------------------------------------------------------------------------------------------
template <int MATH> struct Lm;
template <> struct Lm<0> {
  static float acos_(float x) { return std::acos(x); }
  static float sin_(float x) { return std::sin(x); }
  static float cos_(float x) { return std::cos(x); }
  static float atan2_(float y, float x) { return std::atan2(y, x); }
  static float log_(float x) { return std::log(x); }
  static float pow_(float x, float y) { return std::pow(x, y); }
};
template <> struct Lm<1> {
  static float acos_(float x) { return (float)std::acos((double)x); }
  static float sin_(float x) { return (float)std::sin((double)x); }
  static float cos_(float x) { return (float)std::cos((double)x); }
  static float atan2_(float y, float x) { return (float)std::atan2((double)y, (double)x); }
  static float log_(float x) { return (float)std::log((double)x); }
  static float pow_(float x, float y) { return (float)std::pow((double)x, (double)y); }
};

struct V3 {
  float v[3];
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};

// Eigen: a.cross(b)  (Eigen/src/Geometry/OrthoMethods.h) — each product rounded, then subtracted
inline V3 cross3(const V3& a, const V3& b) {
  V3 r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}
// Eigen: squaredNorm() of a fixed 3-vector = redux_novec_unroller<.,.,0,3>: x^2 + (y^2 + z^2)
inline float sqnorm3(const V3& a) { return a[0] * a[0] + (a[1] * a[1] + a[2] * a[2]); }
inline V3 div3(const V3& a, float s) {
  V3 r;
  r[0] = a[0] / s; r[1] = a[1] / s; r[2] = a[2] / s;
  return r;
}
// Eigen: v.normalized()
inline V3 normalized3(const V3& a) {
  float z = sqnorm3(a);
  if (z > 0.f) return div3(a, std::sqrt(z));
  return a;
}
// Eigen: v.unitOrthogonal() for 3-vectors (Eigen/src/Geometry/OrthoMethods.h)
inline V3 unit_orthogonal3(const V3& s) {
  const float prec = 1e-5f;  // NumTraits<float>::dummy_precision()
  V3 p;
  auto much_smaller = [&](float a, float b) { return std::fabs(a) <= std::fabs(b) * prec; };
  if (!much_smaller(s[0], s[2]) || !much_smaller(s[1], s[2])) {
    float invnm = 1.0f / std::sqrt(s[0] * s[0] + s[1] * s[1]);
    p[0] = -s[1] * invnm; p[1] = s[0] * invnm; p[2] = 0.f;
  } else {
    float invnm = 1.0f / std::sqrt(s[1] * s[1] + s[2] * s[2]);
    p[0] = 0.f; p[1] = -s[2] * invnm; p[2] = s[1] * invnm;
  }
  return p;
}

// ------------------------------------------------------------------------------------------
// pcl::eigen33 (PCL 1.8.1 common/impl/eigen.hpp), Scalar = float.  Call sites VS.h:1166, 1403;
// SV.h:788, 1012.  m is row-major symmetric.
// ------------------------------------------------------------------------------------------
inline void pcl_compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

template <int MATH>
inline void pcl_compute_roots(const float m[3][3], float roots[3]) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] -
             m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] +
             m[1][1] * m[2][2] - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < FLT_EPSILON) {
    pcl_compute_roots2(c2, c1, roots);
  } else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.f) a_over_3 = 0.f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.f) q = 0.f;
    float rho = std::sqrt(-a_over_3);
    float theta = Lm<MATH>::atan2_(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = Lm<MATH>::cos_(theta);
    float sin_theta = Lm<MATH>::sin_(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      std::swap(roots[1], roots[2]);
      if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.f) pcl_compute_roots2(c2, c1, roots);
  }
}

// largest of the three row cross products of (A - ev I); returns its squared length
inline float best_cross(const float a[3][3], float ev, V3& out) {
  V3 r0{{a[0][0] - ev, a[0][1], a[0][2]}};
  V3 r1{{a[1][0], a[1][1] - ev, a[1][2]}};
  V3 r2{{a[2][0], a[2][1], a[2][2] - ev}};
  V3 v1 = cross3(r0, r1), v2 = cross3(r0, r2), v3 = cross3(r1, r2);
  float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
  if (l1 >= l2 && l1 >= l3) { out = div3(v1, std::sqrt(l1)); return l1; }
  if (l2 >= l1 && l2 >= l3) { out = div3(v2, std::sqrt(l2)); return l2; }
  out = div3(v3, std::sqrt(l3));
  return l3;
}

// evecs[r][c]: column c is the eigenvector of evals[c]
template <int MATH>
void pcl_eigen33(const float mat[3][3], float evecs[3][3], float evals[3]) {
  float scale = 0.f;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) scale = std::max(scale, std::fabs(mat[i][j]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) a[i][j] = mat[i][j] / scale;
  pcl_compute_roots<MATH>(a, evals);
  V3 col[3];
  if ((evals[2] - evals[0]) <= FLT_EPSILON) {
    col[0] = V3{{1, 0, 0}}; col[1] = V3{{0, 1, 0}}; col[2] = V3{{0, 0, 1}};
  } else if ((evals[1] - evals[0]) <= FLT_EPSILON) {
    best_cross(a, evals[2], col[2]);
    col[1] = unit_orthogonal3(col[2]);
    col[0] = cross3(col[1], col[2]);
  } else if ((evals[2] - evals[1]) <= FLT_EPSILON) {
    best_cross(a, evals[0], col[0]);
    col[1] = unit_orthogonal3(col[0]);
    col[2] = cross3(col[0], col[1]);
  } else {
    float mmax[3];
    unsigned min_el = 2, max_el = 2;
    mmax[2] = best_cross(a, evals[2], col[2]);
    // column 1: min/max bookkeeping uses the winning length
    mmax[1] = best_cross(a, evals[1], col[1]);
    min_el = mmax[1] <= mmax[min_el] ? 1 : min_el;
    max_el = mmax[1] > mmax[max_el] ? 1 : max_el;
    // column 0: PCL 1.8.1 compares len3 (the row1 x row2 length) in all three branches
    {
      V3 r0{{a[0][0] - evals[0], a[0][1], a[0][2]}};
      V3 r1{{a[1][0], a[1][1] - evals[0], a[1][2]}};
      V3 r2{{a[2][0], a[2][1], a[2][2] - evals[0]}};
      V3 v1 = cross3(r0, r1), v2 = cross3(r0, r2), v3 = cross3(r1, r2);
      float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
      if (l1 >= l2 && l1 >= l3) { mmax[0] = l1; col[0] = div3(v1, std::sqrt(l1)); }
      else if (l2 >= l1 && l2 >= l3) { mmax[0] = l2; col[0] = div3(v2, std::sqrt(l2)); }
      else { mmax[0] = l3; col[0] = div3(v3, std::sqrt(l3)); }
      min_el = l3 <= mmax[min_el] ? 0 : min_el;
      max_el = l3 > mmax[max_el] ? 0 : max_el;
    }
    unsigned mid_el = 3 - min_el - max_el;
    col[min_el] = normalized3(cross3(col[(min_el + 1) % 3], col[(min_el + 2) % 3]));
    col[mid_el] = normalized3(cross3(col[(mid_el + 1) % 3], col[(mid_el + 2) % 3]));
  }
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) evecs[r][c] = col[c][r];
  for (int i = 0; i < 3; i++) evals[i] *= scale;
}

// ------------------------------------------------------------------------------------------
// unit attributes
// ------------------------------------------------------------------------------------------
struct Attr {
  float c[3];   // centroid          VS.h:1126 voxel_centroids_
  float n[3];   // normal            VS.h:1131 voxel_norms_
  float e[8];   // eigen features    VS.h:1130 voxel_eigens_
  int eig_len;  // 1 (empty_attribute, VS.h:1453) or 8
};
inline bool pos_valid(const Attr& a) { return a.c[0] != 0.f && a.c[1] != 0.f && a.c[2] != 0.f; }  // VS.h:1829
inline bool nrm_valid(const Attr& a) { return a.n[0] != 0.f && a.n[1] != 0.f && a.n[2] != 0.f; }  // VS.h:1840
inline bool eig_valid(const Attr& a) { return a.eig_len > 1; }                                    // VS.h:1695

// calculateVoxelCentroid VS.h:1358-1378 / SV.h:745-765
inline void centroid_of(const float* const* pts, int cnt, float out[3]) {
  float xs = 0, ys = 0, zs = 0;
  for (int i = 0; i < cnt; i++) {
    xs = xs + pts[i][0]; ys = ys + pts[i][1]; zs = zs + pts[i][2];
  }
  out[0] = xs / cnt; out[1] = ys / cnt; out[2] = zs / cnt;
}

// calculateCorvariance VS.h:1533-1594 (no /N) ; SV.h:1372-1435 (/N at 1425)
inline void covariance_of(const float* const* pts, int cnt, bool divide_n, float C[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i][j] = 0.f;
  if (cnt > 3) {
    float s[3] = {0, 0, 0};
    for (int i = 0; i < cnt; i++) {
      s[0] = s[0] + pts[i][0]; s[1] = s[1] + pts[i][1]; s[2] = s[2] + pts[i][2];
    }
    float k[3] = {s[0] / cnt, s[1] / cnt, s[2] / cnt};
    for (int j = 0; j < cnt; j++) {
      float d[3] = {pts[j][0] - k[0], pts[j][1] - k[1], pts[j][2] - k[2]};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          float single = d[r] * d[c];
          C[r][c] = C[r][c] + single;
        }
    }
    if (divide_n)
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[r][c] = C[r][c] / cnt;
  }
}

// calculateVoxelNorms VS.h:1380-1429 / calculateSupervoxelNorms SV.h:988-1040
// + calculateEigenFeatures VS.h:1147-1228 / SV.h:768-847 (the reference runs covariance+eigen33
// twice with identical inputs; once is the same result)
template <int MATH>
void unit_features(const float* const* pts, int cnt, bool svgs, Attr& out) {
  centroid_of(pts, cnt, out.c);
  float C[3][3], evec[3][3], ev[3];
  covariance_of(pts, cnt, svgs, C);
  pcl_eigen33<MATH>(C, evec, ev);
  // normal
  float vx = 0 - pts[0][0];
  float vy = 0 - pts[0][1];
  float vz = (float)(1.5 - (double)pts[0][2]);
  float nx = evec[0][0], ny = evec[1][0], nz = evec[2][0];
  if ((nx * vx + ny * vy + nz * vz) < 0) {
    nx = nx * -1; ny = ny * -1; nz = nz * -1;
  }
  out.n[0] = nx; out.n[1] = ny; out.n[2] = nz;
  // eigen features
  out.eig_len = 8;
  float* f = out.e;
  if (ev[0] == 0 && ev[1] == 0 && ev[2] == 0) {
    for (int i = 0; i < 8; i++) f[i] = 0.f;
    return;
  }
  double nrm = std::sqrt((double)ev[0] * (double)ev[0] + (double)ev[1] * (double)ev[1] +
                         (double)ev[2] * (double)ev[2]);
  float e3 = (float)((double)ev[0] / nrm);
  float e2 = (float)((double)ev[1] / nrm);
  float e1 = (float)((double)ev[2] / nrm);
  int k = 0;
  if (!svgs) {
    if (e1 == 0) { f[k++] = 0.f; f[k++] = 1.f; f[k++] = 0.f; }
    else { f[k++] = (e1 - e2) / e1; f[k++] = (e2 - e3) / e1; f[k++] = e3 / e1; }
    f[k++] = e3 / (e1 + e2 + e3);                       // change of curvature VS.h:1197
    if (e2 == 0) f[k++] = 0.f; else f[k++] = (e1 - e3) / e1;  // anisotropy VS.h:1199-1206
  } else {
    if (e1 == 0) { f[k++] = 0.f; f[k++] = 1.f; f[k++] = 0.f; f[k++] = 0.f; }
    else { f[k++] = (e1 - e2) / e1; f[k++] = (e2 - e3) / e1; f[k++] = e3 / e1; f[k++] = (e1 - e3) / e1; }
    f[k++] = e3 / (e1 + e2 + e3);                       // SV.h:824
  }
  if (e1 * e2 * e3 == 0) f[k++] = 0.f;
  else f[k++] = -1 * (e1 * Lm<MATH>::log_(e1) + e2 * Lm<MATH>::log_(e2) + e3 * Lm<MATH>::log_(e3));
  f[k++] = e1 + e2 + e3;
  f[k++] = Lm<MATH>::pow_(e1 * e2 * e3, (float)(1.0 / 3));
}

// ------------------------------------------------------------------------------------------
// measuringDistance VS.h:1597-1720 / SV.h:1756-1878 ; distanceWeight VS.h:1722-1740 / SV.h:1880-1905
// ------------------------------------------------------------------------------------------
struct Sig { float p, n, o, e, c, w; };

template <int MATH>
void measuring_distance(const Attr& v1, const Attr& v2, bool svgs, float d5[5], int64_t* ub_corner) {
  float dist_space = 100, dist_angle = 100, dist_stair = 100, dist_eigen = 100, dist_convx = 100;
  float cos_v1_dist = 0, cos_v2_dist = 0, dist_v1_v2 = 0, cos_v1_v2 = 0, cos_d_s = 0;
  float thred_singular = 0;
  float dist_v1 = 0, dist_v2 = 0, dist_o1 = 0, dist_o2 = 0;
  double a_1 = 0, a_2 = 0, a_1_2 = 0, a_d_s = 0, a_d_s1 = 0, a_d_s2 = 0, PI = 3.1415926;
  float u[3] = {0, 0, 0}, pr[3] = {0, 0, 0};
  bool have_u = false;
  if (pos_valid(v1) && pos_valid(v2)) {
    float dx = v1.c[0] - v2.c[0], dy = v1.c[1] - v2.c[1], dz = v1.c[2] - v2.c[2];
    dist_v1_v2 = (float)std::sqrt((double)dx * (double)dx + (double)dy * (double)dy + (double)dz * (double)dz);
    dist_space = dist_v1_v2;
    if (dist_v1_v2 != 0) {
      u[0] = dx / dist_v1_v2; u[1] = dy / dist_v1_v2; u[2] = dz / dist_v1_v2;
      pr[0] = v1.c[1] * v2.c[2] - v1.c[2] * v2.c[1];
      pr[1] = v1.c[2] * v2.c[0] - v1.c[0] * v2.c[2];
      pr[2] = v1.c[0] * v2.c[1] - v1.c[1] * v2.c[0];
      have_u = true;
    }
  }
  if (nrm_valid(v1) && nrm_valid(v2)) {
    // VGS guards with dist_space!=0 (VS.h:1642), SVGS with dist_v1_v2!=0 (SV.h:1804).  With an
    // empty position and valid normals VGS reads never-filled vectors (undefined behaviour);
    // defined here as SVGS behaves: skip the block, stair = 0.
    bool enter = svgs ? (dist_v1_v2 != 0) : (dist_space != 0);
    if (enter && !have_u) {
      if (ub_corner) (*ub_corner)++;
      enter = false;
      dist_stair = 0;
    } else if (!enter && svgs) {
      dist_stair = 0;  // SV.h:1827-1830
    }
    if (enter) {
      const float* n1 = v1.n; const float* n2 = v2.n;
      cos_v1_v2 = (n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2]);
      cos_v1_dist = (n1[0] * u[0] + n1[1] * u[1] + n1[2] * u[2]);
      cos_v2_dist = (n2[0] * u[0] + n2[1] * u[1] + n2[2] * u[2]);
      cos_d_s = (pr[0] * u[0] + pr[1] * u[1] + pr[2] * u[2]);
      a_1 = Lm<MATH>::acos_(cos_v1_dist);
      a_2 = Lm<MATH>::acos_(cos_v2_dist);
      a_1_2 = Lm<MATH>::acos_(cos_v1_v2);
      a_d_s1 = Lm<MATH>::acos_(cos_d_s);
      a_d_s2 = PI - a_d_s1;
      dist_angle = Lm<MATH>::acos_(cos_v1_v2);
      dist_v1 = n1[0] * v1.c[0] + n1[1] * v1.c[1] + n1[2] * v1.c[2];
      dist_v2 = n2[0] * v2.c[0] + n2[1] * v2.c[1] + n2[2] * v2.c[2];
      dist_o1 = n1[0] * v2.c[0] + n1[1] * v2.c[1] + n1[2] * v2.c[2];
      dist_o2 = n2[0] * v1.c[0] + n2[1] * v1.c[1] + n2[2] * v1.c[2];
      float t1 = dist_o1 - dist_v1, t2 = dist_o2 - dist_v2;
      dist_stair = (float)std::sqrt((double)t1 * (double)t1 + (double)t2 * (double)t2);
    }
    double temp_a = 0.5, temp_off = PI / 6, max_singular = PI / 2;
    thred_singular = (float)((double)(float)max_singular / (1 + std::exp(-1 * temp_a * (a_1_2 - temp_off))));
    a_d_s = a_d_s1;
    if (a_d_s1 > a_d_s2) a_d_s = a_d_s2;
    if (a_d_s > (double)thred_singular) dist_convx = (float)std::fabs(a_1 - a_2);
    else dist_convx = (float)PI;
  }
  if (eig_valid(v1) && eig_valid(v2)) {
    float eigen_cos = 0, eigen_abs1 = 0, eigen_abs2 = 0;
    for (int i = svgs ? 0 : 4; i < 8; i++) {  // VS.h:1699 starts at 4, SV.h:1857 at 0
      eigen_cos = eigen_cos + v1.e[i] * v2.e[i];
      eigen_abs1 = eigen_abs1 + v1.e[i] * v1.e[i];
      eigen_abs2 = eigen_abs2 + v2.e[i] * v2.e[i];
    }
    if (eigen_abs1 != 0 && eigen_abs2 != 0)
      dist_eigen = 1.0f - eigen_cos / (std::sqrt(eigen_abs1) * std::sqrt(eigen_abs2));
  }
  d5[0] = dist_space; d5[1] = dist_angle; d5[2] = dist_stair; d5[3] = dist_eigen; d5[4] = dist_convx;
}

inline float distance_weight(const float d[5], const Sig& s, bool svgs) {
  float sd;
  if (!svgs) {
    float qs = d[0] / s.p, qa = d[1] / s.n, qt = d[2] / s.o, qc = d[4] / s.c, qe = d[3] / s.e;
    sd = (float)std::sqrt((double)qs * (double)qs + (double)qa * (double)qa + (double)qt * (double)qt +
                          (double)qc * (double)qc + (double)qe * (double)qe);
  } else {
    sd = (float)std::sqrt((double)d[0] * (double)d[0] / (double)s.p + (double)d[1] * (double)d[1] / (double)s.n +
                          (double)d[3] * (double)d[3] / (double)s.e + (double)d[2] * (double)d[2] / (double)s.o);
  }
  return (float)std::exp(-0.5 * (double)sd / ((double)s.w * (double)s.w));
}

// ------------------------------------------------------------------------------------------
// cutGraphSegmentation VS.h:1913-2029 / SV.h:1908-2054.  W row-major: W[row*n+col] =
// weight(v1 = idx[row], v2 = idx[col]) (VS.h:1888-1892).  Flat index f = col*n+row (Eigen
// column-major resize, VS.h:1922).  Sort: (weight desc, f asc), NaN last (std::sort's tie order
// is unspecified in the reference; this total order is the project's definition).
// ------------------------------------------------------------------------------------------
struct WI { float w; int f; };
struct NearRec { int centre, a, b; };

int cut_graph(float k, const float* W, int n, std::vector<int>& members_local,
              float near_tol, int64_t* near_count, std::vector<NearRec>* near_out, int centre_id,
              const int* gid) {
  std::vector<WI> arr((size_t)n * n);
  for (int f = 0; f < n * n; f++) {
    int col = f / n, row = f - col * n;
    arr[f].w = W[(size_t)row * n + col];
    arr[f].f = f;
  }
  std::sort(arr.begin(), arr.end(), [](const WI& a, const WI& b) {
    bool an = std::isnan(a.w), bn = std::isnan(b.w);
    if (an != bn) return bn;        // non-NaN first
    if (!an && a.w != b.w) return a.w > b.w;
    return a.f < b.f;
  });
  std::vector<float> seg_int(n, 1.0f);
  std::vector<std::vector<int>> seg_ver(n);
  std::vector<int> seg_size(n, 1), ver_seg(n);
  for (int i = 0; i < n; i++) { seg_ver[i].push_back(i); ver_seg[i] = i; }
  for (size_t i = 0; i < arr.size(); i++) {
    int v1 = arr[i].f / n;
    int v2 = arr[i].f - v1 * n;
    float w = arr[i].w;
    if (ver_seg[v1] != ver_seg[v2]) {
      int s1 = seg_size[ver_seg[v1]], s2 = seg_size[ver_seg[v2]];
      float m1 = seg_int[ver_seg[v1]] - k / s1;
      float m2 = seg_int[ver_seg[v2]] - k / s2;
      int mi, mn; float thr;
      if (m1 >= m2) { mi = ver_seg[v1]; mn = ver_seg[v2]; thr = m1; }
      else { mi = ver_seg[v2]; mn = ver_seg[v1]; thr = m2; }
      if (near_count && !std::isnan(w)) {
        float den = std::max(std::fabs(w), std::fabs(thr));
        if (den > 0 && std::fabs(w - thr) <= near_tol * den) {
          (*near_count)++;
          if (near_out && near_out->size() < (1u << 20))
            near_out->push_back(NearRec{centre_id, gid ? gid[v1] : v1, gid ? gid[v2] : v2});
        }
      }
      if (w > thr) {
        seg_int[mi] = w;
        for (int j = 0; j < seg_size[mn]; j++) {
          seg_ver[mi].push_back(seg_ver[mn][j]);
          ver_seg[seg_ver[mn][j]] = mi;
        }
        seg_size[mi] = seg_size[mi] + seg_size[mn];
        seg_size[mn] = 0;
        seg_ver[mn].clear();
      }
    }
  }
  int seg_con = ver_seg[0];
  members_local = seg_ver[seg_con];  // reference member order (merge order)
  return (int)members_local.size();
}

// ------------------------------------------------------------------------------------------
// PCL 1.8.1 OctreePointCloud: dynamic bounding box + keys (octree_pointcloud.hpp:
// addPointIdx / adoptBoundingBoxToPoint / getKeyBitSize / genOctreeKeyforPoint).
// Call sites test:51-56.
// ------------------------------------------------------------------------------------------
struct PclOctree {
  double mn[3], mx[3], res;
  unsigned depth = 0;
  bool defined = false;
  struct Ev { unsigned lowered; unsigned depth_old; };
  std::vector<Ev> events;

  void key_bit_size_first() {
    const float minValue = std::numeric_limits<float>::epsilon();
    unsigned mk[3];
    for (int a = 0; a < 3; a++) mk[a] = (unsigned)((mx[a] - mn[a]) / res);
    unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
    double lg = std::log((double)max_voxels) / std::log(2.0);
    depth = std::max(std::min(32u, (unsigned)std::ceil(lg - minValue)), 0u);
    double side = (double)(1 << depth) * res - minValue;
    for (int a = 0; a < 3; a++) {
      double over = (side - (mx[a] - mn[a])) / 2.0;
      mn[a] -= over; mx[a] += over;
    }
  }
  void adopt(const float p[3]) {
    const float minValue = std::numeric_limits<float>::epsilon();
    while (true) {
      bool lo[3], up[3], any = false;
      for (int a = 0; a < 3; a++) {
        lo[a] = (p[a] < mn[a]); up[a] = (p[a] >= mx[a]);
        any = any || lo[a] || up[a];
      }
      if (any || !defined) {
        if (defined) {
          double side = (double)(1 << depth) * res;
          unsigned lowered = 0;
          for (int a = 0; a < 3; a++)
            if (!up[a]) { mn[a] -= side; lowered |= 1u << a; }
          events.push_back(Ev{lowered, depth});
          depth++;
          side = (double)(1 << depth) * res - minValue;
          for (int a = 0; a < 3; a++) mx[a] = mn[a] + side;
        } else {
          for (int a = 0; a < 3; a++) { mn[a] = p[a] - res / 2; mx[a] = p[a] + res / 2; }
          key_bit_size_first();
          defined = true;
        }
      } else break;
    }
  }
  void gen_key(const float p[3], unsigned key[3]) const {
    for (int a = 0; a < 3; a++) key[a] = (unsigned)((p[a] - mn[a]) / res);
  }
};

// x-major Morton code (child index = x<<2 | y<<1 | z, octree_key.h getChildIdxWithDepthMask)
inline uint64_t morton_xmajor(const unsigned k[3], unsigned depth) {
  uint64_t m = 0;
  for (int b = (int)depth - 1; b >= 0; b--) {
    m = (m << 3) | (uint64_t)((((k[0] >> b) & 1u) << 2) | (((k[1] >> b) & 1u) << 1) | ((k[2] >> b) & 1u));
  }
  return m;
}

struct CSR {
  std::vector<int64_t> off;
  std::vector<int32_t> idx;
  void from(const std::vector<std::vector<int>>& l, bool sorted) {
    off.assign(l.size() + 1, 0);
    idx.clear();
    for (size_t i = 0; i < l.size(); i++) {
      size_t b = idx.size();
      idx.insert(idx.end(), l[i].begin(), l[i].end());
      if (sorted) std::sort(idx.begin() + b, idx.end());
      off[i + 1] = (int64_t)idx.size();
    }
  }
};

}  // namespace

struct vgso_handle_s {
  vgso_params p;
  int64_t n = 0;
  double bbox[6] = {0, 0, 0, 0, 0, 0};
  std::vector<uint32_t> point_key;
  std::vector<int32_t> point_unit;
  std::vector<uint32_t> unit_key;
  std::vector<float> unit_center;
  std::vector<std::vector<int>> unit_points;
  std::vector<Attr> attr;
  std::vector<uint8_t> used;
  std::vector<std::vector<int>> adj;  // [count, ids...] as the reference stores it (VS.h:249-260)
  std::vector<std::vector<int>> conn;
  std::vector<int32_t> attach;
  CSR c_points, c_adj, c_conn0, c_conn1, c_conn2, c_clusters;
  std::vector<float> f_centroid, f_normal, f_eigen;
  std::vector<int32_t> unit_cluster, point_label;
  std::vector<NearRec> near;
  int64_t stats[16] = {0};

  template <int MATH> int run(const float* xyz, int64_t n, int stride, const int32_t* labels, int32_t max_label);
};

template <int MATH>
int vgso_handle_s::run(const float* xyz, int64_t N, int stride, const int32_t* labels, int32_t max_label) {
  const bool svgs = p.mode == 1;
  n = N;
  std::memset(stats, 0, sizeof(stats));
  near.clear();
  stats[0] = N;
  auto P = [&](int64_t i) { return xyz + i * stride; };

  // ---- (a1) octree voxelisation: test:51-56 (PCL addPointsFromInputCloud) ----
  PclOctree oc;
  oc.res = (double)p.voxel_size;  // VS.h:84 ctor takes double(float voxel_size), test:26,51
  point_key.assign((size_t)N * 3, 0xFFFFFFFFu);
  std::vector<uint32_t> epoch((size_t)N, 0);
  int64_t n_finite = 0;
  for (int64_t i = 0; i < N; i++) {
    const float* q = P(i);
    if (!(std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))) continue;
    oc.adopt(q);
    unsigned k[3];
    oc.gen_key(q, k);
    for (int a = 0; a < 3; a++) point_key[(size_t)i * 3 + a] = k[a];
    epoch[i] = (uint32_t)oc.events.size();
    n_finite++;
  }
  stats[1] = n_finite;
  stats[11] = oc.depth;
  stats[14] = (int64_t)oc.events.size();
  if (oc.depth > 21) return 2;
  {  // keys shift when a new root is added above the old one (child slot = !upper violation)
    size_t ne = oc.events.size();
    std::vector<uint32_t> suffix((ne + 1) * 3, 0);
    for (size_t e = ne; e-- > 0;)
      for (int a = 0; a < 3; a++)
        suffix[e * 3 + a] = suffix[(e + 1) * 3 + a] + (((oc.events[e].lowered >> a) & 1u) ? (1u << oc.events[e].depth_old) : 0u);
    for (int64_t i = 0; i < N; i++) {
      if (point_key[(size_t)i * 3] == 0xFFFFFFFFu) continue;
      for (int a = 0; a < 3; a++) point_key[(size_t)i * 3 + a] += suffix[(size_t)epoch[i] * 3 + a];
    }
  }
  for (int a = 0; a < 3; a++) { bbox[a] = oc.mn[a]; bbox[3 + a] = oc.mx[a]; }

  // ---- (a2) leaf order = voxel ids: VS.h:146-167 (LeafNodeIterator, descending x-major Morton
  //      for PCL 1.8.1), per-leaf indices ascending ----
  std::vector<std::pair<uint64_t, int32_t>> mi;
  mi.reserve((size_t)n_finite);
  for (int64_t i = 0; i < N; i++) {
    if (point_key[(size_t)i * 3] == 0xFFFFFFFFu) continue;
    unsigned k[3] = {point_key[(size_t)i * 3], point_key[(size_t)i * 3 + 1], point_key[(size_t)i * 3 + 2]};
    uint64_t m = morton_xmajor(k, oc.depth);
    if (p.leaf_order == 0) m = ~m;
    mi.emplace_back(m, (int32_t)i);
  }
  std::sort(mi.begin(), mi.end());
  std::vector<std::vector<int>> vox_points;
  std::vector<uint32_t> vox_key;
  std::vector<int32_t> point_voxel((size_t)N, -1);
  for (size_t j = 0; j < mi.size(); j++) {
    if (j == 0 || mi[j].first != mi[j - 1].first) {
      vox_points.emplace_back();
      int32_t i = mi[j].second;
      for (int a = 0; a < 3; a++) vox_key.push_back(point_key[(size_t)i * 3 + a]);
    }
    vox_points.back().push_back(mi[j].second);
    point_voxel[mi[j].second] = (int32_t)vox_points.size() - 1;
  }
  const int64_t NV = (int64_t)vox_points.size();
  // centres: VS.h:2102-2109 with the float-narrowed members set by setVoxelSize / setBoundingBox
  // (VS.h:127, 136-142, 1121-1123; test:55-57)
  std::vector<float> vox_center((size_t)NV * 3);
  {
    float res_f = (float)(double)p.voxel_size;
    float mn_f[3] = {(float)oc.mn[0], (float)oc.mn[1], (float)oc.mn[2]};
    for (int64_t v = 0; v < NV; v++)
      for (int a = 0; a < 3; a++)
        vox_center[(size_t)v * 3 + a] =
            (float)(((double)vox_key[(size_t)v * 3 + a] + 0.5f) * res_f + mn_f[a]);
  }

  // ---- units ----
  if (!svgs) {
    unit_points = std::move(vox_points);
    unit_key = std::move(vox_key);
    unit_center = std::move(vox_center);
    point_unit = std::move(point_voxel);
  } else {
    // createSupervoxels SV.h:288-323: labels 1..max_label-1 with >=1 point, ascending label,
    // points ascending (the label loop at SV.h:313 stops before max_label)
    int32_t ml = max_label;
    if (ml <= 0) { ml = 0; for (int64_t i = 0; i < N; i++) ml = std::max(ml, labels[i]); ml += 1; }
    std::vector<std::vector<int>> map((size_t)ml + 1);
    for (int64_t i = 0; i < N; i++) {
      int32_t l = labels[i];
      if (l > 0 && l <= ml) map[l].push_back((int)i);
    }
    unit_points.clear();
    point_unit.assign((size_t)N, -1);
    for (int32_t k = 0; k < ml; k++)
      if (!map[k].empty()) {
        for (int i : map[k]) point_unit[i] = (int32_t)unit_points.size();
        unit_points.push_back(std::move(map[k]));
      }
    unit_key.clear();
    unit_center.clear();
    stats[15] = NV;  // getVoxelNum() of the 0.05 m octree (SV.h:111-116)
  }
  const int64_t V = (int64_t)unit_points.size();
  stats[2] = V;
  c_points.from(unit_points, false);

  // ---- (a3-a8 / a20) attributes: VS.h:290-369 ; SV.h:1238-1303 ----
  attr.assign((size_t)V, Attr{{0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}, 1});
  used.assign((size_t)V, 0);
  {
    std::vector<const float*> pts;
    for (int64_t v = 0; v < V; v++) {
      int cnt = (int)unit_points[v].size();
      bool u = svgs ? true : (cnt > p.points_min);  // VS.h:322 ; SV.h:1288
      used[v] = u;
      if (!u) continue;
      pts.resize(cnt);
      for (int j = 0; j < cnt; j++) pts[j] = P(unit_points[v][j]);
      unit_features<MATH>(pts.data(), cnt, svgs, attr[v]);
      stats[3]++;
    }
  }
  f_centroid.resize((size_t)V * 3); f_normal.resize((size_t)V * 3); f_eigen.assign((size_t)V * 8, 0.f);
  for (int64_t v = 0; v < V; v++) {
    for (int a = 0; a < 3; a++) { f_centroid[v * 3 + a] = attr[v].c[a]; f_normal[v * 3 + a] = attr[v].n[a]; }
    if (attr[v].eig_len == 8) for (int a = 0; a < 8; a++) f_eigen[v * 8 + a] = attr[v].e[a];
  }

  // ---- (a9 / a21) FLANN radius search: VS.h:223-265 over voxel centres ; SV.h:1477-1521 over
  //      supervoxel centroids.  Exact, strict dist < (float)(r*r), L2_Simple float accumulation,
  //      sorted by (dist, index).  A uniform grid only prunes candidates. ----
  adj.assign((size_t)V, {});
  {
    const float* sp = svgs ? f_centroid.data() : unit_center.data();
    double r = (double)p.graph_size;
    float r2 = (float)(r * r);
    double cell = r * 1.001 + 1e-9;
    double g0[3] = {1e300, 1e300, 1e300};
    for (int64_t v = 0; v < V; v++) for (int a = 0; a < 3; a++) g0[a] = std::min(g0[a], (double)sp[v * 3 + a]);
    auto cellof = [&](const float* q, int64_t c[3]) { for (int a = 0; a < 3; a++) c[a] = (int64_t)std::floor(((double)q[a] - g0[a]) / cell); };
    auto ckey = [](const int64_t c[3]) { return ((uint64_t)(c[0] & 0x1FFFFF) << 42) | ((uint64_t)(c[1] & 0x1FFFFF) << 21) | (uint64_t)(c[2] & 0x1FFFFF); };
    std::unordered_map<uint64_t, std::vector<int>> grid;
    grid.reserve((size_t)V);
    for (int64_t v = 0; v < V; v++) { int64_t c[3]; cellof(sp + v * 3, c); grid[ckey(c)].push_back((int)v); }
    std::vector<std::pair<float, int>> res;
    int64_t E = 0, maxn = 0;
    for (int64_t v = 0; v < V; v++) {
      const float* q = sp + v * 3;
      int64_t c[3]; cellof(q, c);
      res.clear();
      for (int64_t dx = -1; dx <= 1; dx++) for (int64_t dy = -1; dy <= 1; dy++) for (int64_t dz = -1; dz <= 1; dz++) {
        int64_t cc[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
        if (cc[0] < 0 || cc[1] < 0 || cc[2] < 0) continue;
        auto it = grid.find(ckey(cc));
        if (it == grid.end()) continue;
        for (int u : it->second) {
          const float* t = sp + (size_t)u * 3;
          float acc = 0.f;
          for (int a = 0; a < 3; a++) { float diff = q[a] - t[a]; acc += diff * diff; }
          if (acc < r2) res.emplace_back(acc, u);
        }
      }
      std::sort(res.begin(), res.end());
      std::vector<int>& l = adj[v];
      l.assign(res.size() + 1, 0);
      l[0] = (int)res.size();
      for (size_t j = 0; j < res.size(); j++) l[j + 1] = res[j].second;
      E += (int64_t)res.size();
      maxn = std::max<int64_t>(maxn, (int64_t)res.size());
    }
    stats[4] = E;
    if (!svgs) stats[15] = maxn;
    std::vector<std::vector<int>> tmp((size_t)V);
    for (int64_t v = 0; v < V; v++) tmp[v].assign(adj[v].begin() + 1, adj[v].end());
    c_adj.from(tmp, false);
  }

  // ---- (a10-a14) local graph + cut per used unit: VS.h:372-412 ; SV.h:383-413 ----
  Sig sg{p.sig_p, p.sig_n, p.sig_o, p.sig_e, p.sig_c, p.sig_w};
  float near_tol = p.near_tol > 0 ? p.near_tol : 1e-5f;
  conn.assign((size_t)V, {});
  // The units are independent of each other; the reference runs them one after the other on one thread.  With
  // vgso_set_threads(n > 1) (bench.py --impl reference: "all the host threads it can use") the loop is split over
  // OpenMP threads — same arithmetic per unit, counters and near-threshold records merged afterwards in unit order.
  {
    const int nthreads = std::max(1, g_threads);
    std::vector<std::array<int64_t, 4>> tstats((size_t)nthreads, std::array<int64_t, 4>{0, 0, 0, 0});
    std::vector<std::vector<NearRec>> tnear((size_t)nthreads);
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      std::vector<float> W;
      std::vector<int> members;
      int64_t* st = tstats[(size_t)tid].data();   // [0] pair evals  [1] NaN weights  [2] ub corner  [3] near threshold
#pragma omp for schedule(dynamic, 64)
      for (int64_t i = 0; i < V; i++) {
        if (!used[i]) continue;
        const int nn = adj[i][0];
        const int* gid = adj[i].data() + 1;
        W.assign((size_t)nn * nn, 0.f);
        for (int a = 0; a < nn; a++)
          for (int b = 0; b < nn; b++) {
            if (a != b) {
              float d5[5];
              measuring_distance<MATH>(attr[gid[a]], attr[gid[b]], svgs, d5, &st[2]);
              float w = distance_weight(d5, sg, svgs);
              if (std::isnan(w)) st[1]++;
              W[(size_t)a * nn + b] = w;
              st[0]++;
            } else W[(size_t)a * nn + b] = 1.f;
          }
        cut_graph(p.cut_thred, W.data(), nn, members, near_tol, &st[3], &tnear[(size_t)tid], (int)i, gid);
        conn[i].resize(members.size());
        for (size_t j = 0; j < members.size(); j++) conn[i][j] = gid[members[j]];
      }
    }
    for (int t = 0; t < nthreads; t++) {
      stats[5] += tstats[(size_t)t][0]; stats[6] += tstats[(size_t)t][1]; stats[7] += tstats[(size_t)t][2]; stats[8] += tstats[(size_t)t][3];
      near.insert(near.end(), tnear[(size_t)t].begin(), tnear[(size_t)t].end());
    }
    if (nthreads > 1)
      std::stable_sort(near.begin(), near.end(), [](const NearRec& a, const NearRec& b) { return a.centre < b.centre; });
  }
  c_conn0.from(conn, true);

  // ---- (a15) crossValidation VS.h:2111-2179 / SV.h:2142-2184 (literal, in place) ----
  for (int64_t i = 0; i < V; i++) {
    int inthis = (int)conn[i].size();
    if (inthis > 1) {
      std::vector<int> nw;
      for (int j = 0; j < inthis; j++) {
        int s = conn[i][j];
        bool found = false;
        for (int x : conn[s]) if (x == (int)i) found = true;
        if (found) nw.push_back(s);
      }
      conn[i] = nw;
    }
  }
  c_conn1.from(conn, true);

  // ---- (a16) closestCheck VS.h:2181-2303 / SV.h:2186-2305 (literal; the candidate loop starts
  //      at slot 0 = the COUNT, VS.h:2243; distanceProbability := distanceWeight) ----
  attach.assign((size_t)V, -1);
  for (int64_t i = 0; i < V; i++) {
    int inthis = (int)conn[i].size();
    if (inthis > 0 && inthis < 2) {
      stats[12]++;
      if ((int)adj[i].size() > p.adjacency_min) {
        float min_dis = 0; int min_idx = -1;
        for (size_t j = 0; j < adj[i].size(); j++) {
          int c = adj[i][j];
          if (c < 0 || c >= V) continue;  // slot 0 = count may equal V (out of range in the reference)
          if (conn[c].size() > 1) {
            float d5[5];
            measuring_distance<MATH>(attr[i], attr[c], svgs, d5, &stats[7]);
            float t = distance_weight(d5, sg, svgs);
            if (t >= min_dis) { min_dis = t; min_idx = c; }
          }
        }
        if (min_idx != -1) {
          conn[i].push_back(min_idx);
          conn[min_idx].push_back((int)i);
          attach[i] = min_idx;
          stats[13]++;
        }
      }
    }
  }
  c_conn2.from(conn, true);

  // ---- (a17) clusteringVoxels / recursionSearch VS.h:2032-2099 ; SV.h:2057-2107 (iterative DFS
  //      reproducing the recursion's discovery order; seed appended last) ----
  std::vector<std::vector<int>> clusters;
  unit_cluster.assign((size_t)V, -1);
  {
    std::vector<uint8_t> clustered((size_t)V, 0);
    std::vector<std::pair<int, size_t>> stack;
    for (int64_t i = 0; i < V; i++) {
      if (clustered[i]) continue;
      std::vector<int> in;
      clustered[i] = 1;
      stack.clear();
      stack.emplace_back((int)i, 0);
      while (!stack.empty()) {
        auto& top = stack.back();
        const std::vector<int>& l = conn[top.first];
        if (top.second >= l.size()) { stack.pop_back(); continue; }
        int c = l[top.second++];
        if (!clustered[c]) {
          in.push_back(c);
          clustered[c] = 1;
          stack.emplace_back(c, 0);
        }
      }
      in.push_back((int)i);
      for (int v : in) unit_cluster[v] = (int32_t)clusters.size();
      clusters.push_back(std::move(in));
    }
  }
  stats[9] = (int64_t)clusters.size();

  // ---- (a18) export: VS.h:947-1014 (clusters with > voxels_min voxels) ; SV.h:2109-2126 (all) ----
  point_label.assign((size_t)N, -1);
  {
    std::vector<std::vector<int>> out;
    for (auto& cl : clusters) {
      if (!svgs && !((int)cl.size() > p.voxels_min)) continue;
      if (svgs && cl.empty()) continue;
      std::vector<int> pts;
      for (int v : cl) pts.insert(pts.end(), unit_points[v].begin(), unit_points[v].end());
      int mnp = std::numeric_limits<int>::max();
      for (int q : pts) mnp = std::min(mnp, q);
      for (int q : pts) point_label[q] = mnp;
      out.push_back(std::move(pts));
    }
    stats[10] = (int64_t)out.size();
    c_clusters.from(out, false);
  }
  return 0;
}

extern "C" {

int vgso_set_threads(int n) {
#ifdef _OPENMP
  g_threads = n > 0 ? n : 1;
#else
  (void)n;
  g_threads = 1;   // built without OpenMP
#endif
  return g_threads;
}

vgso_handle vgso_create(const vgso_params* p) {
  vgso_handle h = new vgso_handle_s();
  h->p = *p;
  return h;
}
void vgso_destroy(vgso_handle h) { delete h; }

int vgso_run(vgso_handle h, const float* xyz, int64_t n, int stride, const int32_t* labels, int32_t max_label) {
  if (h->p.mode == 1 && !labels) return 1;
  if (h->p.math == 0) return h->run<0>(xyz, n, stride, labels, max_label);
  return h->run<1>(xyz, n, stride, labels, max_label);
}

const void* vgso_get(vgso_handle h, int kind, int64_t* count) {
  auto ret = [&](const void* p, int64_t c) { if (count) *count = c; return p; };
  switch (kind) {
    case VGSO_BBOX: return ret(h->bbox, 6);
    case VGSO_POINT_KEY: return ret(h->point_key.data(), (int64_t)h->point_key.size());
    case VGSO_POINT_UNIT: return ret(h->point_unit.data(), (int64_t)h->point_unit.size());
    case VGSO_UNIT_KEY: return ret(h->unit_key.data(), (int64_t)h->unit_key.size());
    case VGSO_UNIT_CENTER: return ret(h->unit_center.data(), (int64_t)h->unit_center.size());
    case VGSO_UNIT_OFFSETS: return ret(h->c_points.off.data(), (int64_t)h->c_points.off.size());
    case VGSO_UNIT_POINTS: return ret(h->c_points.idx.data(), (int64_t)h->c_points.idx.size());
    case VGSO_CENTROID: return ret(h->f_centroid.data(), (int64_t)h->f_centroid.size());
    case VGSO_NORMAL: return ret(h->f_normal.data(), (int64_t)h->f_normal.size());
    case VGSO_EIGEN: return ret(h->f_eigen.data(), (int64_t)h->f_eigen.size());
    case VGSO_USED: return ret(h->used.data(), (int64_t)h->used.size());
    case VGSO_ADJ_OFFSETS: return ret(h->c_adj.off.data(), (int64_t)h->c_adj.off.size());
    case VGSO_ADJ_IDX: return ret(h->c_adj.idx.data(), (int64_t)h->c_adj.idx.size());
    case VGSO_CONN0_OFFSETS: return ret(h->c_conn0.off.data(), (int64_t)h->c_conn0.off.size());
    case VGSO_CONN0_IDX: return ret(h->c_conn0.idx.data(), (int64_t)h->c_conn0.idx.size());
    case VGSO_CONN1_OFFSETS: return ret(h->c_conn1.off.data(), (int64_t)h->c_conn1.off.size());
    case VGSO_CONN1_IDX: return ret(h->c_conn1.idx.data(), (int64_t)h->c_conn1.idx.size());
    case VGSO_CONN2_OFFSETS: return ret(h->c_conn2.off.data(), (int64_t)h->c_conn2.off.size());
    case VGSO_CONN2_IDX: return ret(h->c_conn2.idx.data(), (int64_t)h->c_conn2.idx.size());
    case VGSO_UNIT_CLUSTER: return ret(h->unit_cluster.data(), (int64_t)h->unit_cluster.size());
    case VGSO_POINT_LABEL: return ret(h->point_label.data(), (int64_t)h->point_label.size());
    case VGSO_CLUSTER_OFFSETS: return ret(h->c_clusters.off.data(), (int64_t)h->c_clusters.off.size());
    case VGSO_CLUSTER_POINTS: return ret(h->c_clusters.idx.data(), (int64_t)h->c_clusters.idx.size());
    case VGSO_NEAR_EDGES: return ret(h->near.data(), (int64_t)h->near.size() * 3);
    case VGSO_STATS: return ret(h->stats, 16);
    case VGSO_ATTACH: return ret(h->attach.data(), (int64_t)h->attach.size());
  }
  if (count) *count = 0;
  return nullptr;
}

static Attr mk_attr(const float* c, const float* n, const float* e, int flags) {
  Attr a{{0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}, 1};
  if (flags & 1) for (int i = 0; i < 3; i++) a.c[i] = c[i];
  if (flags & 2) for (int i = 0; i < 3; i++) a.n[i] = n[i];
  if (flags & 4) { a.eig_len = 8; for (int i = 0; i < 8; i++) a.e[i] = e[i]; }
  return a;
}

void vgso_pair(const vgso_params* p, const float* c1, const float* n1, const float* e1, int flags1,
               const float* c2, const float* n2, const float* e2, int flags2, float* out6) {
  Attr a = mk_attr(c1, n1, e1, flags1), b = mk_attr(c2, n2, e2, flags2);
  Sig sg{p->sig_p, p->sig_n, p->sig_o, p->sig_e, p->sig_c, p->sig_w};
  bool svgs = p->mode == 1;
  if (p->math == 0) measuring_distance<0>(a, b, svgs, out6, nullptr);
  else measuring_distance<1>(a, b, svgs, out6, nullptr);
  out6[5] = distance_weight(out6, sg, svgs);
}

void vgso_eigen33(const float* mat9, float* evals3, float* evecs9, int math) {
  float m[3][3], ev[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = mat9[i * 3 + j];
  if (math == 0) pcl_eigen33<0>(m, ev, evals3); else pcl_eigen33<1>(m, ev, evals3);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) evecs9[i * 3 + j] = ev[i][j];
}

void vgso_features(const vgso_params* p, const float* xyz, int64_t n, float* out14) {
  std::vector<const float*> pts((size_t)n);
  for (int64_t i = 0; i < n; i++) pts[i] = xyz + i * 3;
  Attr a{{0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}, 1};
  if (p->math == 0) unit_features<0>(pts.data(), (int)n, p->mode == 1, a);
  else unit_features<1>(pts.data(), (int)n, p->mode == 1, a);
  for (int i = 0; i < 3; i++) { out14[i] = a.c[i]; out14[3 + i] = a.n[i]; }
  for (int i = 0; i < 8; i++) out14[6 + i] = a.e[i];
}

int vgso_cut(float cut_thred, const float* w, int n, int32_t* out) {
  std::vector<int> members;
  cut_graph(cut_thred, w, n, members, 0.f, nullptr, nullptr, 0, nullptr);
  std::sort(members.begin(), members.end());
  for (size_t i = 0; i < members.size(); i++) out[i] = members[i];
  return (int)members.size();
}

}  // extern "C"

// vgs_math.cuh — per-thread arithmetic of the VGS/SVGS hot path (sm_100a device code; the same
// functions compile as host code so tests/hostcheck can unit-test them without a GPU).
//
// Parity rules: every float/double operation is an IEEE basic operation in
// the order the reference performs it (this file is compiled with -fmad=false / -ffp-contract=off),
// and every libm call the reference makes on a float (acos, sin, cos, atan2, log, pow) is
// evaluated CORRECTLY ROUNDED: in double, rounded once to float.
//
// Reference: voxel_segmentation.h (VS.h) / supervoxel_segmentation.h (SV.h) lines cited per function.
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>

#ifdef __CUDACC__
#define VGS_HD __host__ __device__ __forceinline__
#else
#define VGS_HD inline
#endif

namespace vgs {

// double acos / exp behind a call on the device: pair_weights uses them seven times, and inlining every copy makes
// the pair kernels exceed the instruction cache (ncu: no_instruction stalls)
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ double vgs_dacos(double x) { return acos(x); }
static __device__ __noinline__ double vgs_dexp(double x) { return exp(x); }
#else
VGS_HD double vgs_dacos(double x) { return acos(x); }
VGS_HD double vgs_dexp(double x) { return exp(x); }
#endif

#ifdef VGS_FAST_ROUNDED_MATH
// ---- the float-rounded acos / exp of pair_weights: a short evaluation + a rounding-boundary test ----
// OFF by default (VGS_FAST_ROUNDED_MATH): measured on the 10 M-point site, k_rows_fill 2.29 -> 2.43 ms and k_local_graph2
// 2.76 -> 2.90 ms WITH it — the two-branch evaluation + boundary test + the library fallback that has to stay make the pair
// kernels longer (2 784 -> 3 184 SASS instructions against a 32 KB instruction cache) and the saved fp64 work does not pay for
// it.  Kept, with its proof harness, for kernels that are not instruction-cache bound.
// pair_weights needs (float)acos((double)x) four times and (float)exp(y) up to three times per pair; the library double
// functions behind them are 131 and 60 SASS instructions, a third of the pair kernels.  Only the FLOAT rounding of the result
// is used, so a polynomial with a relative error below 2^-49 suffices whenever the value is not within VGS_FM_REL of a
// float rounding boundary — in that case (about one call in 2^19) the library function decides, as before.  The fast
// path uses IEEE operations only (fma, +, *, sqrt, rint, conversions), so this header computes the same bits on the host:
// tools/fastmath_check.cpp compares vgs_acosf_cr with (float)acos((double)x) for EVERY float in [-1, 1] (and the pi - acos
// pair, and exp over 4e9 arguments of the pair kernels' range): no difference, i.e. results are unchanged.
constexpr double VGS_FM_REL = 1.1368683772161603e-13;   // 2^-43: > 30x the error bound of the evaluations below
constexpr double VGS_FM_ABS = 2.2737367544323206e-13;   // 2^-42, absolute margin for pi - acos (the error of acos is < 2^-48)
// Polynomial coefficients, highest degree first.  On the device they live in constant memory: an FP64 instruction takes a
// constant-bank operand directly, an immediate double costs two extra (uniform-datapath) instructions per coefficient.
#if defined(__CUDA_ARCH__)
#define VGS_COEF_TABLE static __constant__ double
#else
#define VGS_COEF_TABLE static const double
#endif
// (asin(sqrt t) / sqrt t - 1) / t on [0, 1/4], degree 10 (interpolation at Chebyshev nodes; relative error of asin 2^-50)
VGS_COEF_TABLE vgs_c_asin[11] = {0x1.c8a4a8d5d7026p-6, -0x1.bf16e7c9f283cp-8, 0x1.fa1b2b4831188p-7, 0x1.512bc40e88a9ep-7,
                                 0x1.cf5ed14c7cb7ep-7, 0x1.1c0d74beb3610p-6,  0x1.6e8f34a32a3ecp-6, 0x1.f1c6ff7f5507fp-6,
                                 0x1.6db6dba99e56dp-5, 0x1.33333333030cfp-4,  0x1.55555555555bbp-3};
// exp(r) on |r| <= ln2 / 2, degree 10 (relative error 2^-51); the constant term 1 is added by the last fma
VGS_COEF_TABLE vgs_c_exp[10] = {0x1.28a2ca617b969p-22, 0x1.72fafebdaf273p-19, 0x1.a019a6611cad5p-16, 0x1.a01978b8b3d18p-13,
                                0x1.6c16c17f46982p-10, 0x1.1111112ddae8bp-7,  0x1.55555555520a4p-5,  0x1.555555554b736p-3,
                                0x1.0000000000005p-1,  0x1.000000000001ep+0};
VGS_HD double vgs_asin_poly(double t) {
  double p = vgs_c_asin[0];
#pragma unroll
  for (int i = 1; i < 11; i++) p = fma(p, t, vgs_c_asin[i]);
  return p;
}
// acos(x), |x| <= 1, relative error < 2^-49 (fdlibm's case split, without its division)
VGS_HD double vgs_acos_poly(double x) {
  const double a = fabs(x);
  if (a <= 0.5) {
    const double t = x * x;
    const double r = fma(x * t, vgs_asin_poly(t), x);                      // asin(x)
    return 0x1.921fb54442d18p+0 - (r - 0x1.1a62633145c07p-54);              // pi/2 - asin(x)
  }
  const double z = (1.0 - a) * 0.5, sq = sqrt(z);
  const double r = fma(sq * z, vgs_asin_poly(z), sq);                       // asin(sqrt((1 - |x|) / 2))
  return x > 0 ? 2.0 * r : 0x1.921fb54442d18p+1 - (2.0 * r - 0x1.1a62633145c07p-53);
}
// exp(y), -60 <= y <= 2, relative error < 2^-46: y = k ln2 + r, degree-10 polynomial on |r| <= ln2 / 2, scaled by 2^k
VGS_HD double vgs_exp_poly(double y) {
  const double k = rint(y * 0x1.71547652b82fep+0);
  double r = fma(-k, 6.93147180369123816490e-01, y);
  r = fma(-k, 1.90821492927058770002e-10, r);
  double p = vgs_c_exp[0];
#pragma unroll
  for (int i = 1; i < 10; i++) p = fma(p, r, vgs_c_exp[i]);
  p = fma(p, r, 1.0);
  union { long long i; double d; } sc;
  sc.i = (long long)(1023 + (int)k) << 52;
  return p * sc.d;
}
// the float both v (1 - rel) and v (1 + rel) round to, if they agree (v >= 0)
VGS_HD bool vgs_float_if_sure(double v, double rel, float& out) {
  const float lo = (float)(v * (1.0 - rel)), hi = (float)(v * (1.0 + rel));
  out = lo;
  return lo == hi;
}
#if defined(__CUDA_ARCH__)
#define VGS_MATH_FN static __device__ __noinline__
#else
#define VGS_MATH_FN VGS_HD
#endif
// (float)acos((double)x)
VGS_MATH_FN float vgs_acosf_cr(float x) {
  const double xd = (double)x;
  if (fabs(xd) <= 1.0) {
    float f;
    if (vgs_float_if_sure(vgs_acos_poly(xd), VGS_FM_REL, f)) return f;
  }
  return (float)vgs_dacos(xd);
}
// a = (float)acos((double)x) and pia = (float)(pi - acos((double)x)) (the subtraction in double, as pair_weights has it)
VGS_MATH_FN void vgs_acosf_pair_cr(float x, float& a, float& pia) {
  const double xd = (double)x, PI_D = 3.14159265358979323846;
  if (fabs(xd) <= 1.0) {
    const double v = vgs_acos_poly(xd), w = PI_D - v;
    const float wlo = (float)(w - VGS_FM_ABS), whi = (float)(w + VGS_FM_ABS);
    pia = wlo;
    if (vgs_float_if_sure(v, VGS_FM_REL, a) && wlo == whi) return;
  }
  const double ad = vgs_dacos(xd);
  a = (float)ad; pia = (float)(PI_D - ad);
}
// (float)exp(y)
VGS_MATH_FN float vgs_expf_cr(double y) {
  if (y >= -60.0 && y <= 2.0) {
    float f;
    if (vgs_float_if_sure(vgs_exp_poly(y), VGS_FM_REL, f)) return f;
  }
  return (float)vgs_dexp(y);
}
// (float)(c / (1 + exp(u))), c > 0
VGS_MATH_FN float vgs_logistic_cr(double c, double u) {
  if (u >= -60.0 && u <= 2.0) {
    float f;
    if (vgs_float_if_sure(c / (1 + vgs_exp_poly(u)), VGS_FM_REL, f)) return f;
  }
  return (float)(c / (1 + vgs_dexp(u)));
}

#endif   // VGS_FAST_ROUNDED_MATH

// ---- correctly rounded float libm (double evaluation, one rounding) ----
#ifdef VGS_FAST_ROUNDED_MATH
VGS_HD float cr_acosf(float x) { return vgs_acosf_cr(x); }
#else
VGS_HD float cr_acosf(float x) { return (float)vgs_dacos((double)x); }
#endif
VGS_HD float cr_sinf(float x) { return (float)sin((double)x); }
VGS_HD float cr_cosf(float x) { return (float)cos((double)x); }
VGS_HD float cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
VGS_HD float cr_logf(float x) { return (float)log((double)x); }
VGS_HD float cr_powf(float x, float y) { return (float)pow((double)x, (double)y); }

// ---- unit record: 16 floats = 64 B, one per voxel / supervoxel ----
// [0..2] centroid  [3..5] normal  [6..13] eigen features  [14] point count (int bits)  [15] flags (int bits)
enum : int { REC_FLOATS = 16, REC_COUNT = 14, REC_FLAGS = 15 };
enum : int { F_POS = 1, F_NRM = 2, F_EIG = 4, F_USED = 8 };

VGS_HD int f2i(float f) { union { float f; int i; } u; u.f = f; return u.i; }
VGS_HD float i2f(int i) { union { float f; int i; } u; u.i = i; return u.f; }
VGS_HD uint32_t f2u(float f) { union { float f; uint32_t i; } u; u.f = f; return u.i; }

struct PairParams {
  float sig_p, sig_n, sig_o, sig_e, sig_c, sig_w;
  int svgs;
};

// ---- 3-vectors as Eigen evaluates them (pcl::eigen33 uses Eigen::Vector3f) ----
struct F3 { float x, y, z; };
VGS_HD F3 cross3(F3 a, F3 b) {  // Eigen OrthoMethods.h: products rounded, then subtracted
  F3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}
VGS_HD float sqn3(F3 a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }  // Eigen redux order for size 3
VGS_HD F3 div3(F3 a, float s) { F3 r; r.x = a.x / s; r.y = a.y / s; r.z = a.z / s; return r; }
VGS_HD F3 normalized3(F3 a) { float z = sqn3(a); if (z > 0.f) return div3(a, sqrtf(z)); return a; }
VGS_HD F3 unit_orthogonal3(F3 s) {  // Eigen unitOrthogonal(), dummy_precision<float>() = 1e-5
  F3 p;
  bool sx = fabsf(s.x) <= fabsf(s.z) * 1e-5f, sy = fabsf(s.y) <= fabsf(s.z) * 1e-5f;
  if (!sx || !sy) {
    float inv = 1.0f / sqrtf(s.x * s.x + s.y * s.y);
    p.x = -s.y * inv; p.y = s.x * inv; p.z = 0.f;
  } else {
    float inv = 1.0f / sqrtf(s.y * s.y + s.z * s.z);
    p.x = 0.f; p.y = -s.z * inv; p.z = s.y * inv;
  }
  return p;
}

// symmetric 3x3 as 6 floats: a00 a01 a02 a11 a12 a22
struct Sym3 { float a00, a01, a02, a11, a12, a22; };

// pcl::computeRoots2 / computeRoots (PCL 1.8.1 common/impl/eigen.hpp), Scalar=float.
VGS_HD void roots2(float b, float c, float* r) {
  r[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
VGS_HD void roots3(const Sym3& m, float* r) {
  float c0 = m.a00 * m.a11 * m.a22 + 2.0f * m.a01 * m.a02 * m.a12 - m.a00 * m.a12 * m.a12 -
             m.a11 * m.a02 * m.a02 - m.a22 * m.a01 * m.a01;
  float c1 = m.a00 * m.a11 - m.a01 * m.a01 + m.a00 * m.a22 - m.a02 * m.a02 + m.a11 * m.a22 - m.a12 * m.a12;
  float c2 = m.a00 + m.a11 + m.a22;
  if (fabsf(c0) < FLT_EPSILON) { roots2(c2, c1, r); return; }
  const float inv3 = (float)(1.0 / 3.0);
  const float sqrt3 = sqrtf(3.0f);
  float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.f) a_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.f) q = 0.f;
  float rho = sqrtf(-a_3);
  float theta = cr_atan2f(sqrtf(-q), half_b) * inv3;
  float ct = cr_cosf(theta), st = cr_sinf(theta);
  r[0] = c2_3 + 2.0f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  float t;
  if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  if (r[1] >= r[2]) {
    t = r[1]; r[1] = r[2]; r[2] = t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  }
  if (r[0] <= 0.f) roots2(c2, c1, r);
}

// the three row cross products of (A - ev I): lengths and vectors
VGS_HD void row_crosses(const Sym3& a, float ev, F3& v1, F3& v2, F3& v3, float& l1, float& l2, float& l3) {
  F3 r0{a.a00 - ev, a.a01, a.a02}, r1{a.a01, a.a11 - ev, a.a12}, r2{a.a02, a.a12, a.a22 - ev};
  v1 = cross3(r0, r1); v2 = cross3(r0, r2); v3 = cross3(r1, r2);
  l1 = sqn3(v1); l2 = sqn3(v2); l3 = sqn3(v3);
}
VGS_HD float pick_cross(F3 v1, F3 v2, F3 v3, float l1, float l2, float l3, F3& out) {
  if (l1 >= l2 && l1 >= l3) { out = div3(v1, sqrtf(l1)); return l1; }
  if (l2 >= l1 && l2 >= l3) { out = div3(v2, sqrtf(l2)); return l2; }
  out = div3(v3, sqrtf(l3));
  return l3;
}

// pcl::eigen33(mat, evecs, evals): ascending eigenvalues, col[k] = eigenvector of ev[k].
// Call sites VS.h:1166,1403 ; SV.h:788,1012.
VGS_HD void eigen33(const Sym3& mat, float* ev, F3* col) {
  float scale = fmaxf(fmaxf(fmaxf(fabsf(mat.a00), fabsf(mat.a01)), fmaxf(fabsf(mat.a02), fabsf(mat.a11))),
                      fmaxf(fabsf(mat.a12), fabsf(mat.a22)));
  if (scale <= FLT_MIN) scale = 1.0f;
  Sym3 a{mat.a00 / scale, mat.a01 / scale, mat.a02 / scale, mat.a11 / scale, mat.a12 / scale, mat.a22 / scale};
  roots3(a, ev);
  F3 v1, v2, v3; float l1, l2, l3;
  if ((ev[2] - ev[0]) <= FLT_EPSILON) {
    col[0] = F3{1, 0, 0}; col[1] = F3{0, 1, 0}; col[2] = F3{0, 0, 1};
  } else if ((ev[1] - ev[0]) <= FLT_EPSILON) {
    row_crosses(a, ev[2], v1, v2, v3, l1, l2, l3);
    pick_cross(v1, v2, v3, l1, l2, l3, col[2]);
    col[1] = unit_orthogonal3(col[2]);
    col[0] = cross3(col[1], col[2]);
  } else if ((ev[2] - ev[1]) <= FLT_EPSILON) {
    row_crosses(a, ev[0], v1, v2, v3, l1, l2, l3);
    pick_cross(v1, v2, v3, l1, l2, l3, col[0]);
    col[1] = unit_orthogonal3(col[0]);
    col[2] = cross3(col[0], col[1]);
  } else {
    float mmax[3];
    int min_el = 2, max_el = 2;
    row_crosses(a, ev[2], v1, v2, v3, l1, l2, l3);
    mmax[2] = pick_cross(v1, v2, v3, l1, l2, l3, col[2]);
    row_crosses(a, ev[1], v1, v2, v3, l1, l2, l3);
    mmax[1] = pick_cross(v1, v2, v3, l1, l2, l3, col[1]);
    min_el = mmax[1] <= mmax[min_el] ? 1 : min_el;
    max_el = mmax[1] > mmax[max_el] ? 1 : max_el;
    row_crosses(a, ev[0], v1, v2, v3, l1, l2, l3);
    mmax[0] = pick_cross(v1, v2, v3, l1, l2, l3, col[0]);
    // PCL 1.8.1 compares len3 here whichever cross product won
    min_el = l3 <= mmax[min_el] ? 0 : min_el;
    max_el = l3 > mmax[max_el] ? 0 : max_el;
    int mid_el = 3 - min_el - max_el;
    col[min_el] = normalized3(cross3(col[(min_el + 1) % 3], col[(min_el + 2) % 3]));
    col[mid_el] = normalized3(cross3(col[(mid_el + 1) % 3], col[(mid_el + 2) % 3]));
  }
  ev[0] *= scale; ev[1] *= scale; ev[2] *= scale;
}

// Normal + eight eigen features from a finished scatter matrix.
// VS.h:1380-1429 + 1147-1228 (VGS) ; SV.h:988-1040 + 768-847 (SVGS: feature order differs).
// p0 = first point of the unit (lowest index).  rec[3..13] are written.
VGS_HD void finish_features(const Sym3& C, float p0x, float p0y, float p0z, int svgs, float* rec) {
  float ev[3]; F3 col[3];
  eigen33(C, ev, col);
  float vx = 0 - p0x, vy = 0 - p0y, vz = (float)(1.5 - (double)p0z);
  float nx = col[0].x, ny = col[0].y, nz = col[0].z;
  if ((nx * vx + ny * vy + nz * vz) < 0) { nx = nx * -1; ny = ny * -1; nz = nz * -1; }
  rec[3] = nx; rec[4] = ny; rec[5] = nz;
  float* f = rec + 6;
  if (ev[0] == 0 && ev[1] == 0 && ev[2] == 0) {
    for (int i = 0; i < 8; i++) f[i] = 0.f;
    return;
  }
  double nrm = sqrt((double)ev[0] * (double)ev[0] + (double)ev[1] * (double)ev[1] + (double)ev[2] * (double)ev[2]);
  float e3 = (float)((double)ev[0] / nrm), e2 = (float)((double)ev[1] / nrm), e1 = (float)((double)ev[2] / nrm);
  float L, P, S, aniso;
  float curv = e3 / (e1 + e2 + e3);
  if (e1 == 0) { L = 0.f; P = 1.f; S = 0.f; }
  else { L = (e1 - e2) / e1; P = (e2 - e3) / e1; S = e3 / e1; }
  if (!svgs) aniso = (e2 == 0) ? 0.f : (e1 - e3) / e1;   // VS.h:1199-1206
  else aniso = (e1 == 0) ? 0.f : (e1 - e3) / e1;         // SV.h:808-821
  float ent;
  if (e1 * e2 * e3 == 0) ent = 0.f;
  else ent = -1 * (e1 * cr_logf(e1) + e2 * cr_logf(e2) + e3 * cr_logf(e3));
  float sum = e1 + e2 + e3;
  float omni = cr_powf(e1 * e2 * e3, (float)(1.0 / 3));
  f[0] = L; f[1] = P; f[2] = S;
  if (!svgs) { f[3] = curv; f[4] = aniso; } else { f[3] = aniso; f[4] = curv; }
  f[5] = ent; f[6] = sum; f[7] = omni;
}

// Attribute flags of a finished record (VS.h:1829, 1840, 1695).
VGS_HD int attr_flags(const float* rec, bool used) {
  int fl = 0;
  if (rec[0] != 0.f && rec[1] != 0.f && rec[2] != 0.f) fl |= F_POS;
  if (rec[3] != 0.f && rec[4] != 0.f && rec[5] != 0.f) fl |= F_NRM;
  if (used) fl |= F_EIG | F_USED;
  return fl;
}

#ifdef VGS_FAST_ROUNDED_MATH
VGS_HD float pw_expf(double y) { return vgs_expf_cr(y); }
#else
VGS_HD float pw_expf(double y) { return (float)vgs_dexp(y); }
#endif

// ---- perceptual-grouping weight of the UNORDERED pair {a,b}: returns w(a->b) and w(b->a) ----
// measuringDistance VS.h:1597-1720 / SV.h:1756-1878 ; distanceWeight VS.h:1722-1740 / SV.h:1880-1905.
// The two orders share S, A, T, E; only the convexity cue C differs (acosf(-x) != pi - acosf(x)
// in floating point), and SVGS ignores C (SV.h:1900), so one evaluation serves both orders.
VGS_HD void pair_weights(const float* A_, const float* B_, const PairParams& P, float& w_ab, float& w_ba) {
  const int fa = f2i(A_[REC_FLAGS]), fb = f2i(B_[REC_FLAGS]);
  const double PI = 3.1415926;
  float S = 100.f, A = 100.f, T = 100.f, E = 100.f, C_ab = 100.f, C_ba = 100.f;
  float d = 0.f, ux = 0.f, uy = 0.f, uz = 0.f, px = 0.f, py = 0.f, pz = 0.f;
  bool have_u = false;
  if ((fa & F_POS) && (fb & F_POS)) {
    float dx = A_[0] - B_[0], dy = A_[1] - B_[1], dz = A_[2] - B_[2];
    d = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy + (double)dz * (double)dz);
    S = d;
    if (d != 0.f) {
      ux = dx / d; uy = dy / d; uz = dz / d;
      px = A_[1] * B_[2] - A_[2] * B_[1];
      py = A_[2] * B_[0] - A_[0] * B_[2];
      pz = A_[0] * B_[1] - A_[1] * B_[0];
      have_u = true;
    }
  }
  if ((fa & F_NRM) && (fb & F_NRM)) {
    bool enter = P.svgs ? (d != 0.f) : (S != 0.f);
    if (enter && !have_u) { enter = false; T = 0.f; }   // VGS undefined-behaviour corner, defined as SVGS
    else if (!enter && P.svgs) T = 0.f;                 // SV.h:1827-1830
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, a12 = 0, ads1 = 0, ads2 = 0;
    if (enter) {
      const float* n1 = A_ + 3; const float* n2 = B_ + 3;
      float c12 = (n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2]);
      float c1d = (n1[0] * ux + n1[1] * uy + n1[2] * uz);
      float c2d = (n2[0] * ux + n2[1] * uy + n2[2] * uz);
      A = cr_acosf(c12);
      a12 = (double)A;
      float D1 = n1[0] * A_[0] + n1[1] * A_[1] + n1[2] * A_[2];
      float D2 = n2[0] * B_[0] + n2[1] * B_[1] + n2[2] * B_[2];
      float O1 = n1[0] * B_[0] + n1[1] * B_[1] + n1[2] * B_[2];
      float O2 = n2[0] * A_[0] + n2[1] * A_[1] + n2[2] * A_[2];
      float t1 = O1 - D1, t2 = O2 - D2;
      T = (float)sqrt((double)t1 * (double)t1 + (double)t2 * (double)t2);
      if (!P.svgs) {
        float cds = (px * ux + py * uy + pz * uz);
        // order (a,b): acosf(c1d), acosf(c2d); order (b,a): u -> -u exactly, acosf(-c2d), acosf(-c1d).
        // acos(-x) = pi - acos(x) is evaluated in double and rounded once, i.e. the same correctly
        // rounded float as (float)acos(-(double)x) (the 1-ulp double error of the subtraction moves a float
        // with probability ~2^-28) — two double acos per pair instead of four.
#ifdef VGS_FAST_ROUNDED_MATH
        float a1f, a2f, b1f, b2f;
        vgs_acosf_pair_cr(c1d, a1f, b2f);      // a1 = (float)acos(c1d), b2 = (float)(pi - acos(c1d))
        vgs_acosf_pair_cr(c2d, a2f, b1f);
        a1 = (double)a1f; a2 = (double)a2f; b1 = (double)b1f; b2 = (double)b2f;
#else
        const double ad1 = vgs_dacos((double)c1d), ad2 = vgs_dacos((double)c2d);
        a1 = (double)(float)ad1; a2 = (double)(float)ad2;
        b1 = (double)(float)(3.14159265358979323846 - ad2); b2 = (double)(float)(3.14159265358979323846 - ad1);
#endif
        ads1 = (double)cr_acosf(cds);
        ads2 = PI - ads1;
      }
    }
    if (!P.svgs) {
      double max_singular = PI / 2;
#ifdef VGS_FAST_ROUNDED_MATH
      float thr = vgs_logistic_cr((double)(float)max_singular, -1 * 0.5 * (a12 - PI / 6));   // (float)(c / (1 + exp(u)))
#else
      float thr = (float)((double)(float)max_singular / (1 + vgs_dexp(-1 * 0.5 * (a12 - PI / 6))));
#endif
      double ads = ads1;
      if (ads1 > ads2) ads = ads2;
      if (ads > (double)thr) { C_ab = (float)fabs(a1 - a2); C_ba = (float)fabs(b1 - b2); }
      else { C_ab = (float)PI; C_ba = (float)PI; }
    }
  }
  if ((fa & F_EIG) && (fb & F_EIG)) {
    float ec = 0, ea = 0, eb = 0;
    for (int i = P.svgs ? 0 : 4; i < 8; i++) {
      ec = ec + A_[6 + i] * B_[6 + i];
      ea = ea + A_[6 + i] * A_[6 + i];
      eb = eb + B_[6 + i] * B_[6 + i];
    }
    if (ea != 0 && eb != 0) E = 1.0f - ec / (sqrtf(ea) * sqrtf(eb));
  }
  double w2 = (double)P.sig_w * (double)P.sig_w;
  if (!P.svgs) {
    float qs = S / P.sig_p, qa = A / P.sig_n, qt = T / P.sig_o, qe = E / P.sig_e;
    double base = (double)qs * (double)qs + (double)qa * (double)qa + (double)qt * (double)qt;
    double ee = (double)qe * (double)qe;
    float qc = C_ab / P.sig_c;
    float sd = (float)sqrt(base + (double)qc * (double)qc + ee);
    w_ab = pw_expf(-0.5 * (double)sd / w2);
    if (f2u(C_ba) == f2u(C_ab)) { w_ba = w_ab; }
    else {
      qc = C_ba / P.sig_c;
      sd = (float)sqrt(base + (double)qc * (double)qc + ee);
      w_ba = pw_expf(-0.5 * (double)sd / w2);
    }
  } else {
    float sd = (float)sqrt((double)S * (double)S / (double)P.sig_p + (double)A * (double)A / (double)P.sig_n +
                           (double)E * (double)E / (double)P.sig_e + (double)T * (double)T / (double)P.sig_o);
    w_ab = pw_expf(-0.5 * (double)sd / w2);
    w_ba = w_ab;
  }
}

// Full record of one unit from its points in ascending point-index order (sequential fp32 sums
// exactly as calculateVoxelCentroid VS.h:1358-1378 and calculateCorvariance VS.h:1533-1594 /
// SV.h:1372-1435 accumulate them).  fetch(j, x, y, z) loads the j-th point of the unit.
template <class Fetch>
VGS_HD void unit_record(Fetch fetch, int cnt, bool used, int svgs, float* rec) {
  for (int i = 0; i < REC_FLOATS; i++) rec[i] = 0.f;
  rec[REC_COUNT] = i2f(cnt);
  if (!used) { rec[REC_FLAGS] = i2f(0); return; }
  float sx = 0, sy = 0, sz = 0, p0x = 0, p0y = 0, p0z = 0;
  for (int j = 0; j < cnt; j++) {
    float x, y, z;
    fetch(j, x, y, z);
    if (j == 0) { p0x = x; p0y = y; p0z = z; }
    sx = sx + x; sy = sy + y; sz = sz + z;
  }
  float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
  rec[0] = mx; rec[1] = my; rec[2] = mz;
  Sym3 C{0, 0, 0, 0, 0, 0};
  if (cnt > 3) {
    for (int j = 0; j < cnt; j++) {
      float x, y, z;
      fetch(j, x, y, z);
      float dx = x - mx, dy = y - my, dz = z - mz;
      C.a00 = C.a00 + dx * dx; C.a01 = C.a01 + dx * dy; C.a02 = C.a02 + dx * dz;
      C.a11 = C.a11 + dy * dy; C.a12 = C.a12 + dy * dz; C.a22 = C.a22 + dz * dz;
    }
    if (svgs) {  // SV.h:1425
      C.a00 = C.a00 / cnt; C.a01 = C.a01 / cnt; C.a02 = C.a02 / cnt;
      C.a11 = C.a11 / cnt; C.a12 = C.a12 / cnt; C.a22 = C.a22 / cnt;
    }
  }
  finish_features(C, p0x, p0y, p0z, svgs, rec);
  rec[REC_FLAGS] = i2f(attr_flags(rec, true));
}

// ---- x-major Morton codes (PCL child index = x<<2|y<<1|z) ----
VGS_HD uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
VGS_HD uint32_t compact3(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
  x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
  x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
  x = (x ^ (x >> 32)) & 0x1fffffull;
  return (uint32_t)x;
}
VGS_HD uint64_t morton_encode(uint32_t kx, uint32_t ky, uint32_t kz) {
  return (spread3(kx) << 2) | (spread3(ky) << 1) | spread3(kz);
}
VGS_HD void morton_decode(uint64_t m, uint32_t& kx, uint32_t& ky, uint32_t& kz) {
  kx = compact3(m >> 2); ky = compact3(m >> 1); kz = compact3(m);
}
VGS_HD uint64_t hash64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

}  // namespace vgs

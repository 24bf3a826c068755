// fastmath_check.cpp — dev tool / proof harness for the short float-rounded acos / exp of csrc/vgs_math.cuh (the
// VGS_FAST_ROUNDED_MATH option; off in the product build: measured slower in the instruction-cache-bound pair kernels).
// The fast paths use IEEE operations only (fma, +, *, /, sqrt, rint, conversions), so this host build computes the bits
// the device computes.  Checked here:
//   1. vgs_acosf_cr(x) == (float)acos((double)x) and vgs_acosf_pair_cr(x) == ((float)acos, (float)(pi - acos)) for EVERY
//      float x in [-1, 1] (2 x 1 065 353 217 values);
//   2. vgs_expf_cr(y) == (float)exp(y) for y = -0.5 (double)sd / w2 over EVERY float sd in [0, 480] with w2 = 4 and 1 (the
//      task files' sigma_w = 2 and 1), and over 10^9 random doubles in [-60, 2];
//   3. vgs_logistic_cr(c, u) == (float)(c / (1 + exp(u))) for u = -0.5 ((double)A - PI / 6) over every float A in [0, pi].
// It also reports how often the rounding-boundary test hands the decision to the library function.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -fopenmp tools/fastmath_check.cpp -o /tmp/fastmath_check && /tmp/fastmath_check
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VGS_FAST_ROUNDED_MATH 1
#include "../vgs_svgs_segmentation_b200/csrc/vgs_math.cuh"

static float f_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t u_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
  const bool quick = argc > 1;      // any argument: every 64th value only
  const uint32_t step = quick ? 64 : 1;
  long long bad = 0, n = 0, fb = 0;
  // 1. acos over all floats in [-1, 1]
  for (int sign = 0; sign < 2; sign++) {
    long long b = 0, c = 0, f = 0;
#pragma omp parallel for reduction(+ : b, c, f) schedule(static, 1 << 16)
    for (int64_t u = 0; u <= 0x3f800000ll; u += step) {
      const float x = f_of((uint32_t)u | (sign ? 0x80000000u : 0u));
      const double ad = acos((double)x);
      const float ra = (float)ad, rp = (float)(3.14159265358979323846 - ad);
      float a, pia;
      vgs::vgs_acosf_pair_cr(x, a, pia);
      if (u_of(vgs::vgs_acosf_cr(x)) != u_of(ra) || u_of(a) != u_of(ra) || u_of(pia) != u_of(rp)) b++;
      float t;
      if (!vgs::vgs_float_if_sure(vgs::vgs_acos_poly((double)x), vgs::VGS_FM_REL, t)) f++;
      c++;
    }
    bad += b; n += c; fb += f;
  }
  printf("acos: %lld floats, %lld mismatches, %lld decided by the library (%.2e)\n", n, bad, fb, (double)fb / (double)n);
  // 2. exp over the weights' arguments
  long long badx = 0, nx = 0, fbx = 0;
  for (double w2 : {4.0, 1.0}) {
    long long b = 0, c = 0, f = 0;
#pragma omp parallel for reduction(+ : b, c, f) schedule(static, 1 << 16)
    for (int64_t u = 0; u <= (int64_t)0x43f00000; u += step) {      // 0 .. 480
      const float sd = f_of((uint32_t)u);
      const double y = -0.5 * (double)sd / w2;
      if (u_of(vgs::vgs_expf_cr(y)) != u_of((float)exp(y))) b++;
      float t;
      if (!(y >= -60.0 && y <= 2.0 && vgs::vgs_float_if_sure(vgs::vgs_exp_poly(y), vgs::VGS_FM_REL, t))) f++;
      c++;
    }
    badx += b; nx += c; fbx += f;
  }
  {
    long long b = 0, c = 0;
    const int64_t N = quick ? 20000000 : 1000000000;
#pragma omp parallel reduction(+ : b, c)
    {
      std::mt19937_64 rng(12345 + 977 * (uint64_t)
#ifdef _OPENMP
                                        omp_get_thread_num()
#else
                                        0
#endif
      );
      std::uniform_real_distribution<double> d(-60.0, 2.0);
#pragma omp for schedule(static)
      for (int64_t i = 0; i < N; i++) {
        const double y = d(rng);
        if (u_of(vgs::vgs_expf_cr(y)) != u_of((float)exp(y))) b++;
        c++;
      }
    }
    badx += b; nx += c;
  }
  printf("exp: %lld arguments, %lld mismatches, %lld of the weight arguments decided by the library or out of range\n", nx, badx, fbx);
  // 3. the singular-threshold logistic over all angles
  long long badl = 0, nl = 0;
  {
    const double PI = 3.1415926, cmax = (double)(float)(PI / 2);
    long long b = 0, c = 0;
#pragma omp parallel for reduction(+ : b, c) schedule(static, 1 << 16)
    for (int64_t u = 0; u <= (int64_t)u_of(3.1415927f); u += step) {
      const double a12 = (double)f_of((uint32_t)u);
      const double uu = -1 * 0.5 * (a12 - PI / 6);
      if (u_of(vgs::vgs_logistic_cr(cmax, uu)) != u_of((float)(cmax / (1 + exp(uu))))) b++;
      c++;
    }
    badl += b; nl += c;
  }
  printf("logistic: %lld angles, %lld mismatches\n", nl, badl);
  const bool ok = bad == 0 && badx == 0 && badl == 0;
  printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}

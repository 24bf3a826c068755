// host_membench.cpp — dev tool: what the host side of the drop-in output costs on THIS box (10 M points): zero-initialised
// and raw allocations, the coloured-cloud resize (with and without the huge-page hint / a parallel pre-fault), the threaded
// gather, list building and copying.  g++ -O2 -std=c++17 -pthread -Iinclude tools/host_membench.cpp -o /tmp/host_membench
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <memory>
#include <numeric>
#include <random>

#include "vgs_dropin/host_parallel.h"
#include "vgs_dropin/pcl_shim.h"
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); }
int main() {
  const size_t n = 10000000;
  PCXYZPtr in(new PCXYZ);
  in->points.resize(n);
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::mt19937 rng(1);
  std::shuffle(idx.begin(), idx.end(), rng);
  std::vector<int64_t> off = {0, (int64_t)n / 2};
  while ((size_t)off.back() < n) off.push_back(std::min<int64_t>((int64_t)n, off.back() + 3845));
  printf("threads %u\n", vgs_dropin::host_threads());
  for (int rep = 0; rep < 3; rep++) {
    auto t = clk::now();
    { std::vector<int32_t> z(n); printf("vector<int>(n) %.1f | ", ms(t)); }
    t = clk::now();
    std::vector<std::vector<int>> lists;
    vgs_dropin::csr_to_lists(off, idx.data(), lists);
    printf("csr_to_lists %.1f | ", ms(t));
    t = clk::now();
    auto cp = vgs_dropin::copy_lists(lists);
    printf("copy_lists %.1f | ", ms(t));
    t = clk::now();
    PCXYZRGBPtr out(new PCXYZRGB);
    out->points.resize(n);
    printf("resize %.1f | ", ms(t));
    t = clk::now();
    pcl::PointXYZRGB* dst = out->points.data();
    vgs_dropin::parallel_blocks(n, 1 << 16, [&](size_t b, size_t e) {
      for (size_t k = b; k < e; k++) { const auto& p = in->points[idx[k]]; auto& q = dst[k]; q.x = p.x; q.y = p.y; q.z = p.z; q.r = 1; q.g = 2; q.b = 3; }
    });
    printf("paint %.1f | ", ms(t));
    t = clk::now();
    {
      PCXYZRGBPtr o2(new PCXYZRGB);
      o2->points.reserve(n);
      vgs_dropin::advise_huge(o2->points.data(), n * 32);
      o2->points.resize(n);
      printf("hint+resize %.1f | ", ms(t));
    }
    t = clk::now();
    {
      PCXYZRGBPtr o2(new PCXYZRGB);
      o2->points.reserve(n);
      char* raw = (char*)o2->points.data();
      vgs_dropin::parallel_blocks(n * 32 / 4096, 256, [&](size_t b, size_t e) { for (size_t pg = b; pg < e; pg++) raw[pg * 4096] = 0; });
      printf("prefault %.1f ", ms(t));
      o2->points.resize(n);
      printf("+resize %.1f | ", ms(t));
    }
    t = clk::now();
    out.reset(); cp.clear(); cp.shrink_to_fit(); lists.clear(); lists.shrink_to_fit();
    printf("destroy %.1f\n", ms(t));
  }
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  if (f) { char b[128] = {0}; if (fgets(b, 127, f)) printf("THP: %s", b); fclose(f); }
}

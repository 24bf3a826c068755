// host_parallel.h — the host-side loops of the drop-in classes that touch every point (the reference's
// vector<vector<int>> cluster lists and its coloured output cloud, voxel_segmentation.h:947-1014, 117-121) run on a few
// threads: at 10 M points they are memory-bound gathers and first-touch page faults that one core does at ~2 GB/s.
// Formatting only; nothing here computes segmentation results.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <thread>
#include <vector>

namespace vgs_dropin {

inline unsigned host_threads() {
  static const unsigned n = [] {
    const char* e = std::getenv("VGS_DROPIN_THREADS");
    long v = e ? std::atol(e) : (long)std::thread::hardware_concurrency();
    if (v < 1) v = 1;
    if (v > 16) v = 16;
    return (unsigned)v;
  }();
  return n;
}

// f(begin, end) over [0, n) in contiguous blocks, one per thread; runs inline when the range is small
template <class F>
void parallel_blocks(size_t n, size_t min_per_thread, F f) {
  unsigned t = host_threads();
  if (min_per_thread > 0) t = (unsigned)std::min<size_t>(t, std::max<size_t>(1, n / min_per_thread));
  if (t <= 1) { f((size_t)0, n); return; }
  std::vector<std::thread> th;
  th.reserve(t - 1);
  const size_t per = (n + t - 1) / t;
  for (unsigned i = 1; i < t; i++) {
    const size_t b = std::min(n, per * i), e = std::min(n, per * (i + 1));
    if (b < e) th.emplace_back([=, &f] { f(b, e); });
  }
  f((size_t)0, std::min(n, per));
  for (auto& x : th) x.join();
}

// contiguous cluster ranges of about equal element weight, one per thread: f(first_cluster, last_cluster)
template <class SizeAt, class F>
void parallel_by_weight(size_t nc, SizeAt size_at, F f) {
  std::vector<size_t> pre(nc + 1, 0);
  for (size_t c = 0; c < nc; c++) pre[c + 1] = pre[c] + size_at(c);
  const size_t total = pre[nc];
  const unsigned t = total < ((size_t)1 << 20) ? 1u : host_threads();
  if (t <= 1) { f((size_t)0, nc); return; }
  std::vector<std::thread> th;
  for (unsigned i = 0; i < t; i++) {
    const size_t lo = total * i / t, hi = total * (i + 1) / t;
    const size_t cb = (size_t)(std::lower_bound(pre.begin(), pre.end() - 1, lo) - pre.begin());
    const size_t ce = i + 1 == t ? nc : (size_t)(std::lower_bound(pre.begin(), pre.end() - 1, hi) - pre.begin());
    if (cb < ce) th.emplace_back([=, &f] { f(cb, ce); });
  }
  for (auto& x : th) x.join();
}

// the reference's vector<vector<int>> from a CSR (offsets, indices): cluster c = idx[off[c] .. off[c+1]).  One pass per
// list (assign: allocate + copy, no zero fill first), lists shared out by element count.
template <class OffT>
inline void csr_to_lists(const std::vector<OffT>& off, const int* idx, std::vector<std::vector<int>>& out) {
  const size_t nc = off.empty() ? 0 : off.size() - 1;
  out.clear();
  out.resize(nc);
  parallel_by_weight(nc, [&](size_t c) { return (size_t)(off[c + 1] - off[c]); },
                     [&](size_t b, size_t e) { for (size_t c = b; c < e; c++) out[c].assign(idx + off[c], idx + off[c + 1]); });
}

// a hint for big, freshly reserved host buffers (the coloured output cloud is 32 B per point): back them with huge
// pages where the kernel allows it — 2 MB instead of 4 KB per first-touch fault.  Harmless where unsupported.
void advise_huge(const void* p, size_t bytes);

// deep copy of the lists (getClusterIdx returns them by value, VS.h:117-121)
inline std::vector<std::vector<int>> copy_lists(const std::vector<std::vector<int>>& src) {
  std::vector<std::vector<int>> out(src.size());
  parallel_by_weight(src.size(), [&](size_t c) { return src[c].size(); },
                     [&](size_t b, size_t e) { for (size_t c = b; c < e; c++) out[c] = src[c]; });
  return out;
}

}  // namespace vgs_dropin

#if defined(__linux__)
#include <sys/mman.h>
#include <cstdint>
inline void vgs_dropin::advise_huge(const void* p, size_t bytes) {
  const uintptr_t huge = (uintptr_t)2 << 20;
  const uintptr_t a = ((uintptr_t)p + huge - 1) & ~(huge - 1), b = ((uintptr_t)p + bytes) & ~(huge - 1);
  if (b > a) (void)madvise((void*)a, (size_t)(b - a), MADV_HUGEPAGE);
}
#else
inline void vgs_dropin::advise_huge(const void*, size_t) {}
#endif

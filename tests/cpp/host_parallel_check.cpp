// host_parallel_check.cpp — the host-side helpers of the drop-in classes (include/vgs_dropin/host_parallel.h): the threaded
// CSR -> vector<vector<int>> conversion, the threaded deep copy and parallel_blocks must equal their serial definitions for
// every thread count, incl. empty clusters, one giant cluster and tiny inputs.  Prints "ok" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>

#include "vgs_dropin/host_parallel.h"

static bool check(size_t n, const std::vector<long>& off) {
  std::vector<int> idx(n ? n : 1);
  std::iota(idx.begin(), idx.end(), 7);
  std::vector<std::vector<int>> lists;
  vgs_dropin::csr_to_lists(off, idx.data(), lists);
  if (lists.size() != (off.empty() ? 0 : off.size() - 1)) return false;
  for (size_t c = 0; c + 1 < off.size(); c++)
    if (lists[c] != std::vector<int>(idx.begin() + off[c], idx.begin() + off[c + 1])) return false;
  if (vgs_dropin::copy_lists(lists) != lists) return false;
  // parallel_blocks covers [0, n) exactly once
  std::vector<unsigned char> hit(n ? n : 1, 0);
  vgs_dropin::parallel_blocks(n, 1000, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) hit[i]++; });
  for (size_t i = 0; i < n; i++) if (hit[i] != 1) return false;
  return true;
}

int main() {
  bool ok = true;
  {   // one giant cluster, empty clusters, many small ones
    const size_t n = 3000000;
    std::vector<long> off = {0, 0, (long)(n / 2), (long)(n / 2)};
    while ((size_t)off.back() < n) off.push_back(std::min<long>((long)n, off.back() + 3845));
    ok = ok && check(n, off);
  }
  ok = ok && check(1000, {0, 10, 10, 1000});
  ok = ok && check(5, {0, 5});
  ok = ok && check(0, {0});
  ok = ok && check(0, {});
  char big[1 << 16];
  vgs_dropin::advise_huge(big, sizeof(big));     // a hint: must be harmless on any memory
  printf(ok ? "ok threads=%u\n" : "MISMATCH threads=%u\n", vgs_dropin::host_threads());
  return ok ? 0 : 1;
}

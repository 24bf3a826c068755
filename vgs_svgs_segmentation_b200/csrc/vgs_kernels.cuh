// vgs_kernels.cuh — the stage kernels of the VGS/SVGS hot path for sm_100a.
// One kernel per row of SURVEY.md §2's kernel table; algorithmic bytes per unit in DESIGN.md.
#pragma once
#include "vgs_prims.cuh"

namespace vgs {

constexpr int MAX_EPOCHS = 40;
constexpr int N_CLASSES = 13;
// class c holds local graphs whose ENUMERATED vertex count (used neighbours; all neighbours when the
// empty-pair weight can merge) is <= CLASS_N[c]; pool capacity CLASS_N(CLASS_N-1) directed weights
__constant__ int c_class_n[N_CLASSES] = {16, 24, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 181};
constexpr int CLASS_N_HOST[N_CLASSES] = {16, 24, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 181};
constexpr int CLASS_T_HOST[N_CLASSES] = {64, 64, 64, 64, 64, 128, 128, 128, 128, 128, 256, 256, 256};
constexpr int LG_CS = 512;     // staging capacity (entries) of one sorted chunk
constexpr int LG_CH = 128;     // target chunk size (all merges of a planar neighbourhood happen in the top ~100 weights)
constexpr int LG_BINS = 256;   // weight histogram bins
constexpr int MAX_NEIGH = 181;
constexpr int REC_PAD = 17;  // smem row stride of a 16-float record (bank-conflict free)

struct Box { double mn[3], mx[3]; };

// PCL dynamic bounding box, per insertion epoch (octree_pointcloud.hpp adoptBoundingBoxToPoint):
// points with index in [viol[e], viol[e+1]) were keyed against origin mn[e]; growth events that
// happened later shift those keys by shift[e] (the old root becomes a child of the new root).
struct EpochTable {
  int n;
  long long viol[MAX_EPOCHS];
  double mn[MAX_EPOCHS][3];
  uint32_t shift[MAX_EPOCHS][3];
};

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// ---- stage 0: smallest index in [start,end) of a finite point outside the current box
//      (have_box == 0: of any finite point).  12 B/point read. ----
__global__ void __launch_bounds__(256) k_find_outside(const float* __restrict__ xyz, int stride, int64_t start, int64_t end,
                                                    Box box, int have_box, unsigned long long* __restrict__ found,
                                                    unsigned long long* __restrict__ nonfinite) {
  unsigned long long best = ~0ull;
  unsigned int nf = 0;
  for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = xyz + i * stride;
    float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) { nf++; continue; }
    bool out = !have_box || (double)x < box.mn[0] || (double)y < box.mn[1] || (double)z < box.mn[2] ||
               (double)x >= box.mx[0] || (double)y >= box.mx[1] || (double)z >= box.mx[2];
    if (out) { best = (unsigned long long)i; break; }  // indices ascend per thread
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (best != ~0ull) atomicMin(found, best);
    if (nonfinite && nf) atomicAdd(nonfinite, (unsigned long long)nf);
  }
}

// the coordinates of the point k_find_outside found, next to its index (one host read per growth epoch instead of two)
__global__ void k_fetch_found(const float* __restrict__ xyz, int stride, const unsigned long long* __restrict__ found, float* __restrict__ out3) {
  const unsigned long long i = *found;
  if (i == ~0ull) return;
  const float* p = xyz + i * stride;
  out3[0] = p[0]; out3[1] = p[1]; out3[2] = p[2];
}

// ---- stage 0 on the device: the growth loop of PCL's dynamic bounding box without a host round trip per epoch.
//      The state (box, depth, epoch table, growth events, scan cursor) lives in device memory; one ROUND is
//      k_origin_scan (first finite point outside the box in the window [cursor, cursor + window)) followed by
//      k_origin_adopt (one thread: adoptBoundingBoxToPoint / getKeyBitSize of octree_pointcloud.hpp on that point, or the
//      window moves on).  The host enqueues a batch of rounds back to back and reads the state once; rounds after the
//      end are empty launches.  Every operation of the adopt step is an IEEE double operation in the order of the host
//      version (OctState::adopt), so both give the same box bit for bit (-fmad=false). ----
constexpr int MAX_GROW_EVENTS = 40;
struct OriginState {
  double mn[3], mx[3], res;
  unsigned depth;
  int defined, done, error;          // error: 1 = depth > 21, 2 = too many epochs
  long long cursor, window, n;
  unsigned long long found;          // smallest violating index of the current window (~0 = none)
  int n_epochs, n_events;
  long long viol[MAX_EPOCHS];
  double ep_mn[MAX_EPOCHS][3];
  int events_before[MAX_EPOCHS];
  unsigned ev_lowered[MAX_GROW_EVENTS], ev_depth_old[MAX_GROW_EVENTS];
};
constexpr long long ORIGIN_WINDOW0 = 1ll << 18;   // violations come early or never: a short window first, then x64 per miss
__global__ void k_origin_init(OriginState* __restrict__ st, double res, long long n) {
  st->res = res; st->depth = 0; st->defined = 0; st->done = n <= 0 ? 1 : 0; st->error = 0;
  st->cursor = 0; st->window = ORIGIN_WINDOW0; st->n = n; st->found = ~0ull; st->n_epochs = 0; st->n_events = 0;
  for (int a = 0; a < 3; a++) { st->mn[a] = 0; st->mx[a] = 0; }
}
__global__ void __launch_bounds__(256) k_origin_scan(const float* __restrict__ xyz, int stride, OriginState* __restrict__ st) {
  if (st->done) return;
  const long long start = st->cursor;
  const long long end = min(st->n, start + st->window);
  const int have_box = st->defined;
  const double m0 = st->mn[0], m1 = st->mn[1], m2 = st->mn[2], M0 = st->mx[0], M1 = st->mx[1], M2 = st->mx[2];
  unsigned long long best = ~0ull;
  for (long long i = start + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
    const float* p = xyz + i * stride;
    const float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) continue;
    const bool out = !have_box || (double)x < m0 || (double)y < m1 || (double)z < m2 || (double)x >= M0 || (double)y >= M1 || (double)z >= M2;
    if (out) { best = (unsigned long long)i; break; }   // indices ascend per thread
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&st->found, best);
}
__global__ void k_origin_adopt(const float* __restrict__ xyz, int stride, OriginState* __restrict__ st) {
  if (st->done) return;
  const unsigned long long f = st->found;
  if (f == ~0ull) {       // nothing outside the box in this window
    st->cursor += st->window;
    st->window *= 64;
    if (st->cursor >= st->n) st->done = 1;
    return;
  }
  const float* q = xyz + (long long)f * stride;
  const float p[3] = {q[0], q[1], q[2]};
  const float eps = 1.1920928955078125e-07f;   // std::numeric_limits<float>::epsilon()
  double mn[3] = {st->mn[0], st->mn[1], st->mn[2]}, mx[3] = {st->mx[0], st->mx[1], st->mx[2]};
  const double res = st->res;
  unsigned depth = st->depth;
  bool defined = st->defined != 0;
  int nev = st->n_events;
  while (true) {
    bool up[3], any = false;
    for (int a = 0; a < 3; a++) {
      const bool lo = p[a] < mn[a];
      up[a] = p[a] >= mx[a];
      any = any || lo || up[a];
    }
    if (!any && defined) break;
    if (defined) {
      double side = (double)(1 << depth) * res;
      unsigned lowered = 0;
      for (int a = 0; a < 3; a++) if (!up[a]) { mn[a] -= side; lowered |= 1u << a; }
      if (nev < MAX_GROW_EVENTS) { st->ev_lowered[nev] = lowered; st->ev_depth_old[nev] = depth; }
      nev++;
      depth++;
      side = (double)(1 << depth) * res - eps;
      for (int a = 0; a < 3; a++) mx[a] = mn[a] + side;
      if (depth > 30) break;
    } else {
      for (int a = 0; a < 3; a++) { mn[a] = p[a] - res / 2; mx[a] = p[a] + res / 2; }
      unsigned mk = 2;
      for (int a = 0; a < 3; a++) mk = max(mk, (unsigned)((mx[a] - mn[a]) / res));
      depth = (unsigned)ceil(log((double)mk) / log(2.0) - eps);
      const double side = (double)(1 << depth) * res - eps;
      for (int a = 0; a < 3; a++) {
        const double over = (side - (mx[a] - mn[a])) / 2.0;
        mn[a] -= over; mx[a] += over;
      }
      defined = true;
    }
  }
  for (int a = 0; a < 3; a++) { st->mn[a] = mn[a]; st->mx[a] = mx[a]; }
  st->depth = depth; st->defined = 1; st->n_events = nev;
  const int e = st->n_epochs;
  if (depth > 21 || nev > MAX_GROW_EVENTS) { st->error = 1; st->done = 1; return; }
  if (e >= MAX_EPOCHS) { st->error = 2; st->done = 1; return; }
  st->viol[e] = (long long)f;
  for (int a = 0; a < 3; a++) st->ep_mn[e][a] = mn[a];
  st->events_before[e] = nev;
  st->n_epochs = e + 1;
  st->cursor = (long long)f + 1;
  st->window = ORIGIN_WINDOW0;
  st->found = ~0ull;
  if (st->cursor >= st->n) st->done = 1;
}

// ---- stage 1a: octree key per point -> sortable 64-bit code, value = point index.
//      key.a = (unsigned)((p.a - min_a)/res) in double (genOctreeKeyforPoint).  12 B read, 12 B written. ----
//      K = uint32_t when the key (3 * depth bits + the sentinel bit) fits 32 bits: the sort then moves 8 instead of 12 bytes
//      per point and pass.
template <class K>
__global__ void __launch_bounds__(256) k_quantise(const float* __restrict__ xyz, int stride, int64_t n, EpochTable ep, double res,
                                                int depth, int descending, K* __restrict__ keys,
                                                uint32_t* __restrict__ vals, uint32_t* __restrict__ key3_out, int gidx_w = 0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + i * stride;
  float x = p[0], y = p[1], z = p[2];
  vals[i] = (uint32_t)i;
  // slab tiles (vgs_tiles.cuh): the insertion epoch is decided by the point's GLOBAL index, carried in the record
  const long long gi = gidx_w ? (long long)__float_as_uint(p[3]) : (long long)i;
  const uint64_t mask = (depth * 3 >= 64) ? ~0ull : ((1ull << (3 * depth)) - 1ull);
  if (!finite3(x, y, z)) {
    keys[i] = (K)(1ull << (3 * depth));  // sorts after every real key
    if (key3_out) { key3_out[3 * i] = key3_out[3 * i + 1] = key3_out[3 * i + 2] = 0xffffffffu; }
    return;
  }
  int e = ep.n - 1;
  while (e > 0 && gi < ep.viol[e]) e--;
  uint32_t kx = (uint32_t)(((double)x - ep.mn[e][0]) / res) + ep.shift[e][0];
  uint32_t ky = (uint32_t)(((double)y - ep.mn[e][1]) / res) + ep.shift[e][1];
  uint32_t kz = (uint32_t)(((double)z - ep.mn[e][2]) / res) + ep.shift[e][2];
  uint64_t m = morton_encode(kx, ky, kz);
  keys[i] = (K)(descending ? (~m & mask) : m);
  if (key3_out) { key3_out[3 * i] = kx; key3_out[3 * i + 1] = ky; key3_out[3 * i + 2] = kz; }
}

// SVGS units: key = supervoxel label (SV.h:288-323); dropped labels get the key max_label, which sorts after every kept
// label (kept labels are < max_label), so the sort only has to look at the bits of max_label
__global__ void __launch_bounds__(256) k_label_keys(const int32_t* __restrict__ labels, const float* __restrict__ xyz, int stride, int64_t n,
                                                  int32_t max_label, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t l = labels[i];
  vals[i] = (uint32_t)i;
  bool ok = l > 0 && l < max_label;
  keys[i] = ok ? (uint64_t)(uint32_t)l : (uint64_t)(uint32_t)max_label;
}

// Built-in stand-in for the supervoxel generator (the reference calls PCL's VCCS, SV.h:265-284, which
// is third-party code): one supervoxel per occupied cell of a seed-resolution grid anchored at the
// octree origin.  Deterministic; NOT VCCS.
__global__ void __launch_bounds__(256) k_seed_cell_keys(const float* __restrict__ xyz, int stride, int64_t n, double ox, double oy, double oz,
                                                      double seed, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + i * stride;
  float x = p[0], y = p[1], z = p[2];
  vals[i] = (uint32_t)i;
  if (!finite3(x, y, z)) { keys[i] = 1ull << 63; return; }
  uint32_t cx = (uint32_t)(((double)x - ox) / seed), cy = (uint32_t)(((double)y - oy) / seed), cz = (uint32_t)(((double)z - oz) / seed);
  keys[i] = morton_encode(cx & 0x1fffffu, cy & 0x1fffffu, cz & 0x1fffffu);
}

// ---- stage 1b: segment heads of the sorted keys.  A head is a position whose key differs from its predecessor's; the unit
//      id of a position is the number of heads up to it, minus one.  The scan runs straight over the keys (reduce per
//      tile -> k_scan_tiles -> down sweep): no flag array is materialised; 8 B / point read, 4 B / point written. ----
template <class K>
__device__ __forceinline__ uint32_t head_flag(const K* __restrict__ keys, int64_t i) { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
template <class K>
__global__ void __launch_bounds__(SC_THREADS) k_heads_reduce(const K* __restrict__ keys, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t sm[33];
  const int64_t base = (int64_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    const int64_t i = base + (int64_t)k * SC_THREADS + threadIdx.x;
    if (i < n) s += head_flag(keys, i);
  }
  uint32_t tot;
  block_excl_scan(s, &tot, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// writes unit start offsets, the unit's sort key and the sorted-position -> unit id map
template <class K>
__global__ void __launch_bounds__(SC_THREADS) k_heads_down(const K* __restrict__ keys, int64_t n, const uint32_t* __restrict__ tile_sums,
                                                         uint32_t* __restrict__ ustart, uint64_t* __restrict__ ukey, uint32_t* __restrict__ pos_unit) {
  __shared__ uint32_t sm[33];
  const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_IPT;  // blocked arrangement
  K kk[SC_IPT + 1];
  kk[0] = base > 0 && base - 1 < n ? keys[base - 1] : (K)0;
  uint32_t f[SC_IPT];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    kk[k + 1] = (base + k < n) ? keys[base + k] : (K)0;
    f[k] = (base + k < n) ? ((base + k == 0 || kk[k + 1] != kk[k]) ? 1u : 0u) : 0u;
    s += f[k];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, &tot, sm) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    if (base + k < n) {
      const uint32_t u = ex + f[k] - 1u;
      pos_unit[base + k] = u;
      if (f[k]) { ustart[u] = (uint32_t)(base + k); ukey[u] = (uint64_t)kk[k + 1]; }
    }
    ex += f[k];
  }
}

// ---- stage 2: per-unit record (centroid, scatter, eigen33, normal, 8 eigen features).
//      One thread per unit, points visited in ascending index order so the fp32 sums are the
//      reference's sums bit for bit.  (Measured and rejected in round 2: gathering the points of a warp's 32 units into
//      shared memory with all lanes first and summing from there — 0.46 ms against 0.24 ms: the per-thread gathers already
//      overlap across the 2 waves of resident threads, and the staged walk pays bank conflicts.) ----
constexpr int FEAT_THREADS = 128;
__global__ void __launch_bounds__(FEAT_THREADS) k_features(const float* __restrict__ xyz, int stride, const uint32_t* __restrict__ perm,
                                                const uint32_t* __restrict__ ustart, int64_t nunits, int points_min, int svgs,
                                                float* __restrict__ rec, uint8_t* __restrict__ uflags,
                                                unsigned long long* __restrict__ n_used) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nunits) return;
  uint32_t s = ustart[u], e = ustart[u + 1];
  int cnt = (int)(e - s);
  bool used = svgs ? true : (cnt > points_min);
  float r[REC_FLOATS];
  unit_record(
      [&](int j, float& x, float& y, float& z) {
        const float* p = xyz + (int64_t)perm[s + j] * stride;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
      },
      cnt, used, svgs, r);
  float4* out = reinterpret_cast<float4*>(rec + u * REC_FLOATS);
  out[0] = make_float4(r[0], r[1], r[2], r[3]);
  out[1] = make_float4(r[4], r[5], r[6], r[7]);
  out[2] = make_float4(r[8], r[9], r[10], r[11]);
  out[3] = make_float4(r[12], r[13], r[14], r[15]);
  uflags[u] = (uint8_t)f2i(r[REC_FLAGS]);   // compact copy of the flags: the graph stages test F_USED of every neighbour
  if (used) atomicAdd(n_used, 1ull);
}

// plain (non-complemented) morton of each voxel, the hash-table key
__global__ void __launch_bounds__(256) k_plain_morton(const uint32_t* __restrict__ key3, int64_t n, uint64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = morton_encode(key3[3 * i], key3[3 * i + 1], key3[3 * i + 2]);
}

// ---- stage 3 (SVGS): radius search over supervoxel centroids (SV.h:1477-1521).  Uniform grid of
//      cell size >= r: a unit's neighbours lie in the 27 cells around its own. ----
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void __launch_bounds__(256) k_centroid_min(const float* __restrict__ rec, int64_t nu, uint32_t* __restrict__ gmin) {
  uint32_t m[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += (int64_t)gridDim.x * blockDim.x)
    for (int a = 0; a < 3; a++) m[a] = min(m[a], f2ord(rec[u * REC_FLOATS + a]));
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[a] = min(m[a], __shfl_xor_sync(0xffffffffu, m[a], o));
    if ((threadIdx.x & 31) == 0) atomicMin(&gmin[a], m[a]);
  }
}
__device__ __forceinline__ void cell_of(const float* c, const uint32_t* gmin, float cell, int& cx, int& cy, int& cz) {
  cx = (int)floorf((c[0] - ord2f(gmin[0])) / cell);
  cy = (int)floorf((c[1] - ord2f(gmin[1])) / cell);
  cz = (int)floorf((c[2] - ord2f(gmin[2])) / cell);
}
__global__ void __launch_bounds__(256) k_cell_keys(const float* __restrict__ rec, int64_t nu, const uint32_t* __restrict__ gmin, float cell,
                                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  int cx, cy, cz;
  cell_of(rec + u * REC_FLOATS, gmin, cell, cx, cy, cz);
  keys[u] = morton_encode((uint32_t)cx & 0x1fffffu, (uint32_t)cy & 0x1fffffu, (uint32_t)cz & 0x1fffffu);
  vals[u] = (uint32_t)u;
}
// one warp per unit; cstart/cunits = cell table (units sorted by cell key).  Lanes 0..26 look up the 27 cells around the
// unit's own at once (one round of hash probes instead of 27 dependent ones), a warp scan concatenates their member
// ranges, and the lanes walk that list 32 candidates at a time.  fill = 0: count only; fill = 1: lists ordered by
// (dist2, id) with a rank count (FLANN's order, independent of the order the candidates were met in).
__global__ void __launch_bounds__(128) k_adjacency_svgs(const float* __restrict__ rec, int64_t nu, const uint32_t* __restrict__ gmin,
                                                      float cell, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ cunits,
                                                      const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                      uint64_t mask, float r2, int fill, uint32_t* __restrict__ adj_cnt,
                                                      const uint32_t* __restrict__ adj_off, int32_t* __restrict__ adj_idx, int cap,
                                                      unsigned long long* __restrict__ overflow) {
  extern __shared__ unsigned char smraw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sd2 = reinterpret_cast<float*>(smraw) + (size_t)w * cap;
  int* sid = reinterpret_cast<int*>(smraw + (size_t)wpb * cap * sizeof(float)) + (size_t)w * cap;
  const int64_t v = (int64_t)blockIdx.x * wpb + w;
  if (v >= nu) return;
  const float qx = rec[v * REC_FLOATS], qy = rec[v * REC_FLOATS + 1], qz = rec[v * REC_FLOATS + 2];
  int cx, cy, cz;
  cell_of(rec + v * REC_FLOATS, gmin, cell, cx, cy, cz);
  // lane d < 27: member range of cell d
  uint32_t s0 = 0, n0 = 0;
  if (lane < 27) {
    const int x = cx + lane / 9 - 1, y = cy + (lane / 3) % 3 - 1, z = cz + lane % 3 - 1;
    if (x >= 0 && y >= 0 && z >= 0) {
      const int c = hash_lookup(tk, tv, mask, morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z));
      if (c >= 0) { s0 = cstart[c]; n0 = cstart[c + 1] - s0; }
    }
  }
  __syncwarp();                                              // the probe loops end at different iterations
  const uint32_t incl = warp_incl_scan(n0, lane);            // candidates of the cells up to and including this lane's
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  int count = 0;
  for (uint32_t b = 0; b < total; b += 32) {
    const uint32_t t = b + lane;
    const uint32_t tt = min(t, total - 1u);      // every lane takes part in the shuffles below
    // cell of candidate tt = first lane whose inclusive count exceeds tt (a 5-step search over the lanes' counts)
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const uint32_t probe = __shfl_sync(0xffffffffu, incl, min(lo + step - 1, 31));
      if (lo + step <= 31 && probe <= tt) lo += step;
    }
    const uint32_t incl_lo = __shfl_sync(0xffffffffu, incl, lo), n_lo = __shfl_sync(0xffffffffu, n0, lo), s_lo = __shfl_sync(0xffffffffu, s0, lo);
    int id = -1;
    float d2 = 0.f;
    if (t < total) {
      id = (int)cunits[s_lo + (tt - (incl_lo - n_lo))];
      const float* q = rec + (int64_t)id * REC_FLOATS;
      const float dx = qx - q[0], dy = qy - q[1], dz = qz - q[2];
      d2 = 0.f; d2 += dx * dx; d2 += dy * dy; d2 += dz * dz;
      if (!(d2 < r2)) id = -1;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, id >= 0);
    if (fill && id >= 0) {
      const int pos = count + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) { sd2[pos] = d2; sid[pos] = id; }
    }
    count += __popc(bal);
  }
  if (count >= cap) { if (lane == 0) atomicAdd(overflow, 1ull); count = cap - 1; }
  if (!fill) { if (lane == 0) adj_cnt[v] = (uint32_t)count; return; }
  __syncwarp();
  const uint32_t off = adj_off[v];
  for (int e = lane; e < count; e += 32) {
    const float d = sd2[e]; const int id = sid[e];
    int rank = 0;
    for (int j = 0; j < count; j++) {
      const float dj = sd2[j]; const int ij = sid[j];
      rank += (dj < d || (dj == d && ij < id)) ? 1 : 0;
    }
    adj_idx[off + rank] = id;
  }
}

// weight of any pair that involves an unused (all-empty) unit: one value per parameter set
__global__ void k_wempty(PairParams pp, float* __restrict__ out) {
  float z[REC_FLOATS];
  for (int q = 0; q < REC_FLOATS; q++) z[q] = 0.f;
  float a, b;
  pair_weights(z, z, pp, a, b);
  out[0] = a;
}

// ---- bin used units by the size of their local graph.  One warp per unit: counts the USED
//      neighbours; pairs with an unused unit are enumerated only if their constant weight could
//      merge (w_empty > cut bound), which never happens with the reference's parameter sets. ----
__global__ void __launch_bounds__(128) k_bin_classes(const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                   const uint8_t* __restrict__ uflags, int64_t nu, int64_t first, int64_t last, float cut,
                                                   int svgs, const float* __restrict__ wempty, uint32_t* __restrict__ class_count,
                                                   uint32_t* __restrict__ class_maxn, uint64_t* __restrict__ class_key,
                                                   uint32_t* __restrict__ class_val,
                                                   unsigned long long* __restrict__ stats /* [0]=sum n(n-1) [1]=max n [2]=overflow */) {
  const int lane = threadIdx.x & 31;
  const int64_t u = first + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (u >= last) return;
  if (!(uflags[u] & F_USED)) return;
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  int used = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    bool us = false;
    if (e < n) {
      const int64_t g = adj_idx[off + e];
      us = (__ldg(uflags + g) & F_USED) != 0;
    }
    used += __popc(__ballot_sync(0xffffffffu, us));
  }
  if (lane != 0) return;
  const float lb = (float)(1.0 - 2.0 * (double)cut + (double)cut / (double)n - 4e-7 * (double)(n + 8));
  const int need = (svgs || wempty[0] > lb) ? n : used;
  int c = 0;
  while (c < N_CLASSES && need > c_class_n[c]) c++;
  atomicMax(&stats[1], (unsigned long long)n);
  if (c >= N_CLASSES || n > 255) { atomicAdd(&stats[2], 1ull); return; }
  atomicAdd(&stats[0], (unsigned long long)n * (unsigned long long)(n - 1));
  atomicMax(&class_maxn[c], (uint32_t)n);
  atomicAdd(&class_count[c], 1u);
  class_key[u] = (uint64_t)c;      // one stable radix pass on this key orders every class list by unit id
  (void)class_val;
}
// (key, value) = (255 = not in any class, unit id) for every unit
__global__ void __launch_bounds__(256) k_class_init(uint64_t* __restrict__ class_key, uint32_t* __restrict__ class_val, int64_t nu) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < nu) { class_key[u] = 255ull; class_val[u] = (uint32_t)u; }
}

// ---- stage 4+5a, general kernel: local affinity graph + Felzenszwalb-style cut of ONE unit per CTA
//      (buildAdjacencyGraph VS.h:1796-1910 + cutGraphSegmentation VS.h:1913-2029).  Used for SVGS (units are not on a
//      lattice) and for the VGS units the row kernel (vgs_rows.cuh) hands back.
//      1. directed weights of all pairs of the neighbourhood, evaluated here from the records.  Weights
//         w <= 1-2k+k/n can never merge (cut bound) and are dropped.
//      2. the reference sorts all n^2 weights; here a 256-bin histogram of the weights delimits
//         chunks of ~512 entries that are sorted (bitonic, shared memory) and merged in descending
//         order (w desc, flat index asc) until the segment of local vertex 0 is final (S0 rule).
//      3. the merge scans 32 sorted entries per warp step: the first mergeable entry merges, the
//         later ones are re-evaluated against the new state. ----
struct GraphParams {
  PairParams pp;
  float cut;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_local_graph2(const uint32_t* __restrict__ list, uint32_t nlist,
                                                        const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                        const float* __restrict__ rec,
                                                        GraphParams gp, int ncap, int mcap, const float* __restrict__ wempty, int bucketed,
                                                        uint32_t* __restrict__ conn_cnt, int32_t* __restrict__ conn_idx) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int AUX = REC_PAD;
  float* A_w = reinterpret_cast<float*>(smraw);                                  // mcap: weight pool (append order)
  float* C_w = A_w + mcap;                                                       // LG_CS: sorted chunk
  float* s_int = C_w + LG_CS;                                                    // ncap
  int* s_gid = reinterpret_cast<int*>(s_int + ncap);                             // ncap
  int* s_aux = s_gid + ncap;                                                     // AUX*ncap: records
  unsigned* s_hist = reinterpret_cast<unsigned*>(s_aux + (size_t)AUX * ncap);    // LG_BINS+1
  unsigned short* A_f = reinterpret_cast<unsigned short*>(s_hist + 2 * LG_BINS + 2); // mcap: packed (col << 8) | row
  unsigned short* C_f = A_f + mcap;                                              // LG_CS
  unsigned short* s_seg = C_f + LG_CS;                                           // ncap
  unsigned short* s_size = s_seg + ncap;                                         // ncap
  unsigned short* s_ul = s_size + ncap;                                          // ncap: local ids of the used vertices
  unsigned short* B_i = s_ul + ncap;                                             // bucketed ? mcap : 0: pool indices ordered by bin
  unsigned* s_cur = s_hist + LG_BINS + 1;                                        // LG_BINS scatter cursors (layout: hist | cursors)
  __shared__ int s_m, s_cnt, s_c1, s_done, s_nseg, s_tot, s_nu, s_c0b, s_before;
  __shared__ float s_wempty, s_ratio;

  const int tid = threadIdx.x;
  if (blockIdx.x >= nlist) return;
  const uint32_t u = list[blockIdx.x];
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  const float k = gp.cut;
  if (tid == 0) { s_m = 0; s_nseg = n; s_done = (n <= 1) ? 1 : 0; s_ratio = 1.0f; }
  for (int i = tid; i <= LG_BINS; i += THREADS) s_hist[i] = 0;
  for (int i = tid; i < n; i += THREADS) {
    s_gid[i] = adj_idx[off + i];
    s_seg[i] = (unsigned short)i; s_size[i] = 1; s_int[i] = 1.0f;
  }
  __syncthreads();
  {
    float* s_rec = reinterpret_cast<float*>(s_aux);
    for (int t = tid; t < n * 4; t += THREADS) {
      int i = t >> 2, q = t & 3;
      float4 val = __ldg(reinterpret_cast<const float4*>(rec + (int64_t)s_gid[i] * REC_FLOATS) + q);
      float* d = s_rec + i * REC_PAD + q * 4;
      d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
    }
  }
  if (tid == 0) s_wempty = wempty[0];
  __syncthreads();
  if (tid < 32) {   // ordered list of the used local vertices
    int cnt = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
      const int v = b0 + tid;
      bool us = false;
      if (v < n) us = (s_aux[v * REC_PAD + REC_FLAGS] & F_USED) != 0;
      const uint32_t bal = __ballot_sync(0xffffffffu, us);
      if (us) s_ul[cnt + __popc(bal & ((1u << tid) - 1u))] = (unsigned short)v;
      cnt += __popc(bal);
    }
    if (tid == 0) s_nu = cnt;
  }
  __syncthreads();
  // --- 1. directed weights of all unordered pairs -> pool + histogram ---
  const float lb = (float)(1.0 - 2.0 * (double)k + (double)k / (double)n - 4e-7 * (double)(n + 8));
  const float scale = (float)LG_BINS / fmaxf(1.0f - lb, 1e-3f);
  // Pairs with an unused (all-empty) unit all have the same weight w_empty; when that weight cannot
  // merge (w_empty <= lb, the normal case) only pairs of USED vertices are enumerated.
  const bool all_pairs = s_wempty > lb;
  const int nv = all_pairs ? n : s_nu;
  const int npairs = nv * (nv - 1) / 2;
  if (!all_pairs) {
    // vertices that are not enumerated have no entry at all: they stay singletons for ever and must not
    // keep the termination tests alive (segment count, smallest live threshold)
    for (int i = tid; i < n; i += THREADS)
      if (!(s_aux[i * REC_PAD + REC_FLAGS] & F_USED)) s_size[i] = 0;
    if (tid == 0) { s_nseg = nv; if (nv <= 1) s_done = 1; }
  }
  for (int p = tid; p < npairs; p += THREADS) {
    int r = p / (nv - 1), c = p - r * (nv - 1);
    int a, b;
    if (c < nv - 1 - r) { a = r; b = r + 1 + c; }
    else { a = nv - 1 - r; b = a + 1 + (c - (nv - 1 - r)); }
    if (!all_pairs) { a = s_ul[a]; b = s_ul[b]; }
    float w_ab, w_ba;
    {
      const float* s_rec = reinterpret_cast<const float*>(s_aux);
      pair_weights(s_rec + a * REC_PAD, s_rec + b * REC_PAD, gp.pp, w_ab, w_ba);
    }
    // matrix entry (row i, col j) = weight(v1=idx[i], v2=idx[j]); the reference's flat index is
    // col*n + row (VS.h:1922) and it reads v1 = col, v2 = row back (VS.h:1959-1960).  Stored packed
    // as (col << 8) | row, which orders exactly like the flat index (row < n <= 256).
    if (w_ab > lb) {
      int s = atomicAdd(&s_m, 1); A_w[s] = w_ab; A_f[s] = (unsigned short)((b << 8) | a);
      atomicAdd(&s_hist[min(LG_BINS - 1, (int)((1.0f - w_ab) * scale))], 1u);
    }
    if (w_ba > lb) {
      int s = atomicAdd(&s_m, 1); A_w[s] = w_ba; A_f[s] = (unsigned short)((a << 8) | b);
      atomicAdd(&s_hist[min(LG_BINS - 1, (int)((1.0f - w_ba) * scale))], 1u);
    }
  }
  __syncthreads();
  const int m = s_m;
  // --- 2+3. chunks of descending weight: gather -> sort -> merge, until nothing more can merge.
  //     Entries whose two vertices already share a segment are no-ops for ever (segments only
  //     grow) and are skipped at gather time; s_ratio tracks how many entries survive that, so the
  //     bin range of the next chunk is sized for ~LG_CH surviving entries. ---
  // inclusive prefix sums of the histogram (warp 0), so a chunk boundary is a short search
  if (tid < 32) {
    constexpr int PER = LG_BINS / 32;
    unsigned loc[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) { sum += s_hist[tid * PER + q]; loc[q] = sum; }
    unsigned inc = warp_incl_scan(sum, tid);
    const unsigned base = inc - sum;
#pragma unroll
    for (int q = 0; q < PER; q++) s_hist[tid * PER + q] = base + loc[q];
  }
  __syncthreads();
  if (bucketed) {   // order the pool indices by bin once, so a chunk is a contiguous index range
    for (int b = tid; b < LG_BINS; b += THREADS) s_cur[b] = b > 0 ? s_hist[b - 1] : 0u;
    __syncthreads();
    for (int i = tid; i < m; i += THREADS) {
      const int bin = min(LG_BINS - 1, (int)((1.0f - A_w[i]) * scale));
      B_i[atomicAdd(&s_cur[bin], 1u)] = (unsigned short)i;
    }
    __syncthreads();
  }
  int c0 = 0;
  while (c0 < LG_BINS && !s_done && m > 0) {
    if (tid == 0) {
      const int before = c0 > 0 ? (int)s_hist[c0 - 1] : 0;    // entries in bins < c0
      // target: (count in [c0,c1)) * ratio <= LG_CH, at least one non-empty bin
      const float ratio = s_ratio;
      const int budget = before + max(1, (int)((float)LG_CH / ratio));
      int lo = c0, hi = LG_BINS;       // largest c1 with prefix[c1-1] <= budget
      while (lo < hi) { int mid = (lo + hi + 1) >> 1; if ((int)s_hist[mid - 1] <= budget) lo = mid; else hi = mid - 1; }
      int c1 = lo;
      if (c1 <= c0 || (int)s_hist[c1 - 1] == before) {   // next non-empty bin alone exceeds the budget (or none left)
        int l2 = c0, h2 = LG_BINS;      // smallest c1 with prefix[c1-1] > before
        while (l2 < h2) { int mid = (l2 + h2) >> 1; if (mid >= 1 && (int)s_hist[mid - 1] > before) h2 = mid; else l2 = mid + 1; }
        c1 = l2;
        if (c1 > LG_BINS) c1 = LG_BINS;
        if (c1 >= 1 && c1 <= LG_BINS && (int)s_hist[c1 - 1] == before) c1 = LG_BINS;   // nothing left
      }
      s_c1 = c1; s_cnt = 0; s_tot = (c1 >= 1 ? (int)s_hist[c1 - 1] : 0) - before; s_before = before;
      s_c0b = (s_tot == 0) ? LG_BINS : c0;
    }
    __syncthreads();
    c0 = s_c0b;
    const int c1 = s_c1;
    if (c0 >= LG_BINS) break;
    const int L = s_tot;
    const bool single_big = (L > LG_CS) && s_ratio >= 1.0f;   // one bin alone overflows the staging buffer
    int kept = 0;
    if (!single_big) {
      // gather the chunk's still-useful entries
      if (bucketed) {
        const int xs = s_before, xe = xs + L;
        for (int x = xs + tid; x < xe; x += THREADS) {
          const int i = B_i[x];
          const unsigned short f = A_f[i];
          if (s_seg[f >> 8] != s_seg[f & 255]) {
            int s = atomicAdd(&s_cnt, 1);
            if (s < LG_CS) { C_w[s] = A_w[i]; C_f[s] = f; }
          }
        }
      } else {
        for (int i = tid; i < m; i += THREADS) {
          const float w = A_w[i];
          const int bin = min(LG_BINS - 1, (int)((1.0f - w) * scale));
          if (bin >= c0 && bin < c1) {
            const unsigned short f = A_f[i];
            if (s_seg[f >> 8] != s_seg[f & 255]) {
              int s = atomicAdd(&s_cnt, 1);
              if (s < LG_CS) { C_w[s] = w; C_f[s] = f; }
            }
          }
        }
      }
      __syncthreads();
      kept = s_cnt;
      if (kept > LG_CS) {       // estimate was too optimistic: retry this range conservatively
        __syncthreads();
        if (tid == 0) s_ratio = 1.0f;
        __syncthreads();
        continue;
      }
      if (kept == 0) {          // every entry of this range is already inside one segment
        __syncthreads();
        c0 = c1;
        continue;
      }
    }
    const int nwin = single_big ? (L + LG_CS - 1) / LG_CS : 1;
    for (int win = 0; win < nwin; win++) {
      int Lw;
      if (!single_big) {
        Lw = kept;
        int P = 32;
        while (P < Lw) P <<= 1;
        for (int i = Lw + tid; i < P; i += THREADS) { C_w[i] = -1.0f; C_f[i] = 0xffff; }
        __syncthreads();
        // bitonic sort: (w desc, packed index asc)
        for (int kk = 2; kk <= P; kk <<= 1) {
          for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += THREADS) {
              const int x = i ^ j;
              if (x > i) {
                const float wi = C_w[i], wx = C_w[x];
                const unsigned short fi = C_f[i], fx = C_f[x];
                const bool x_before_i = (wx > wi) || (wx == wi && fx < fi);
                const bool up = (i & kk) == 0;
                if (x_before_i == up) { C_w[i] = wx; C_w[x] = wi; C_f[i] = fx; C_f[x] = fi; }
              }
            }
            __syncthreads();
          }
        }
      } else {
        // one bin holds more entries than the staging buffer: place by exact rank, window by window
        Lw = min(LG_CS, L - win * LG_CS);
        for (int i = tid; i < m; i += THREADS) {
          const float w = A_w[i];
          const int bx = min(LG_BINS - 1, (int)((1.0f - w) * scale));
          if (bx < c0 || bx >= c1) continue;
          const unsigned short f = A_f[i];
          int rank = 0;
          for (int y = 0; y < m; y++) {
            const float wy = A_w[y];
            const int by = min(LG_BINS - 1, (int)((1.0f - wy) * scale));
            if (by >= c0 && by < c1 && ((wy > w) || (wy == w && A_f[y] < f))) rank++;
          }
          if (rank >= win * LG_CS && rank < (win + 1) * LG_CS) { C_w[rank - win * LG_CS] = w; C_f[rank - win * LG_CS] = f; }
        }
        __syncthreads();
      }
      // merge (warp 0)
      if (tid < 32) {
        const int lane = tid;
        int nseg = s_nseg;
        // S0 rule: only the segment of local vertex 0 is emitted, and its next merge needs an entry with
        // w > max(thr1, thr2) >= thr(S0) = Int(S0) - k/|S0| (unchanged until S0 merges again).  Entries come in
        // descending order, so once the next weight is <= thr(S0) the emitted segment is final — exact, and far
        // earlier than waiting for the smallest live threshold of ALL segments (an outlier singleton keeps that at 1-k).
        const int s0v = s_seg[0];
        const float thr0 = s_int[s0v] - k / (float)(int)s_size[s0v];
        const bool below = (Lw > 0) && !(C_w[0] > thr0);
        bool stop = below || nseg <= 1;
        for (int base = 0; base < Lw && !stop; base += 32) {
          const int e = base + lane;
          const bool valid = e < Lw;
          const float w = valid ? C_w[e] : 0.f;
          const int f = valid ? (int)C_f[e] : 0;
          const int v1 = f >> 8, v2 = f & 255;
          uint32_t todo = __ballot_sync(0xffffffffu, valid);
          while (todo) {
            bool pred = false, a_wins = true;
            int sa = 0, sb = 0;
            if ((todo >> lane) & 1u) {
              sa = s_seg[v1]; sb = s_seg[v2];
              if (sa != sb) {
                const float m1 = s_int[sa] - k / (float)(int)s_size[sa];
                const float m2 = s_int[sb] - k / (float)(int)s_size[sb];
                a_wins = (m1 >= m2);
                pred = w > (a_wins ? m1 : m2);
              }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, pred);
            if (!bal) break;
            const int Lm = __ffs(bal) - 1;
            const int keep = __shfl_sync(0xffffffffu, a_wins ? sa : sb, Lm);
            const int drop = __shfl_sync(0xffffffffu, a_wins ? sb : sa, Lm);
            const float wl = __shfl_sync(0xffffffffu, w, Lm);
            __syncwarp();     // every lane has read its segments / thresholds (the ballot above orders them; this states it)
            for (int v = lane; v < n; v += 32) if (s_seg[v] == drop) s_seg[v] = (unsigned short)keep;
            if (lane == 0) { s_int[keep] = wl; s_size[keep] = (unsigned short)(s_size[keep] + s_size[drop]); s_size[drop] = 0; }
            nseg--;
            __syncwarp();
            todo &= ~((2u << Lm) - 1u);
            if (nseg <= 1) { stop = true; break; }
          }
        }
        __syncwarp();
        if (lane == 0) {
          s_nseg = nseg;
          if (nseg <= 1 || below) s_done = 1;
          if (!single_big) {
            float r = 1.5f * (float)(kept + 8) / (float)(L + 8);
            s_ratio = fminf(1.0f, fmaxf(r, 1.0f / 64.0f));
          }
        }
      }
      __syncthreads();
      if (s_done) break;
    }
    c0 = c1;
  }
  // --- emit the segment that contains local vertex 0 (the unit itself) ---
  if (tid < 32) {
    const int lane = tid;
    const int s0 = s_seg[0];
    int cnt = 0;
    for (int b = 0; b < n; b += 32) {
      const int v = b + lane;
      const bool in = v < n && s_seg[v] == s0;
      const uint32_t bal = __ballot_sync(0xffffffffu, in);
      if (in) conn_idx[off + cnt + __popc(bal & ((1u << lane) - 1u))] = s_gid[v];
      cnt += __popc(bal);
    }
    if (lane == 0) conn_cnt[u] = (uint32_t)cnt;
  }
}

// ---- stage 5b: crossValidation (VS.h:2111-2179): keep j in L[i] iff i in L[j].  One warp per unit. ----
__global__ void __launch_bounds__(128) k_mutual(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ cnt0,
                                              const int32_t* __restrict__ idx0, int64_t nu, uint32_t* __restrict__ cnt1,
                                              int32_t* __restrict__ idx1) {
  const int lane = threadIdx.x & 31;
  int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  const uint32_t off = adj_off[u];
  const int c = (int)cnt0[u];
  int kept = 0;
  for (int b = 0; b < c; b += 32) {
    int e = b + lane;
    bool keep = false;
    int j = -1;
    if (e < c) {
      j = idx0[off + e];
      if (c <= 1) keep = true;  // lists of size <= 1 are left untouched (VS.h:2120)
      else {
        const uint32_t oj = adj_off[j];
        const int cj = (int)cnt0[j];
        for (int t = 0; t < cj; t++) if (idx0[oj + t] == (int)u) { keep = true; break; }
      }
    }
    uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (keep) idx1[off + kept + __popc(bal & ((1u << lane) - 1u))] = j;
    kept += __popc(bal);
  }
  if (lane == 0) cnt1[u] = (uint32_t)kept;
}

// ---- stage 5c: closestCheck (VS.h:2181-2303) as a fixed-point iteration.  The reference visits
//      units in id order and links in place, so a single unit i may attach to a smaller-id single
//      that already attached.  attach[i] depends only on attach[c], c < i, hence iterating to a
//      fixed point reproduces the sequential result.  One warp per single unit (k_closest_round_warp). ----

// units whose list is {self} after the mutual filter and that have enough neighbours (VS.h:2199-2201); woff (optional) =
// start of the single's n + 1 kept candidate weights in a compact buffer, wtotal = its running size
__global__ void __launch_bounds__(256) k_collect_singles(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ cnt1, int64_t nu,
                                                       int adjacency_min, uint32_t* __restrict__ list, uint32_t* __restrict__ counts /* [0]=eligible [1]=all singles */,
                                                       unsigned long long* __restrict__ woff = nullptr, unsigned long long* __restrict__ wtotal = nullptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu || cnt1[i] != 1u) return;
  atomicAdd(&counts[1], 1u);
  const int n = (int)(adj_off[i + 1] - adj_off[i]);
  if (!(n + 1 > adjacency_min)) return;
  const uint32_t slot = atomicAdd(&counts[0], 1u);
  list[slot] = (uint32_t)i;
  if (woff) woff[slot] = atomicAdd(wtotal, (unsigned long long)(n + 1));
}
// one warp per single unit: lanes evaluate the candidates; the winner is the LAST candidate with the
// largest weight (`>=` in VS.h:2281), candidates in the order slot 0 (= the COUNT), then the neighbours.
// A candidate with a list of more than one unit is always eligible; a candidate that is itself a single is eligible only
// once it has attached, and only for singles with a larger id — the one rule that makes the reference's loop order
// dependent.  Weights never change between rounds, only that eligibility: the first round (first != 0) evaluates the
// weight of every candidate that is or can become eligible and keeps it (n + 1 floats per single from woff[slot] in
// wcache; -1 = never eligible) and files the singles that have such "dynamic" candidates in dep_list; the later rounds
// visit only dep_list (length read on the device: no host round trip in between) and pick from the kept weights.
__global__ void __launch_bounds__(128) k_closest_round_warp(const uint32_t* __restrict__ singles, uint32_t nsingles, const uint32_t* __restrict__ adj_off,
                                                          const int32_t* __restrict__ adj_idx, const uint32_t* __restrict__ cnt1,
                                                          const float* __restrict__ rec, int64_t nu, PairParams pp, int32_t* attach,
                                                          uint32_t* __restrict__ changed, int first, float* __restrict__ wcache,
                                                          const unsigned long long* __restrict__ woff, uint32_t* __restrict__ dep_list,
                                                          uint32_t* __restrict__ dep_count) {
  const int lane = threadIdx.x & 31;
  uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (first) { if (li >= nsingles) return; }
  else { if (li >= *dep_count) return; li = dep_list[li]; }
  const int64_t i = singles[li];
  const uint32_t off = adj_off[i];
  const int n = (int)(adj_off[i + 1] - off);
  const unsigned long long wbase = woff[li];
  float ri[REC_FLOATS], rc[REC_FLOATS];
  if (first)
    for (int q = 0; q < REC_FLOATS; q++) ri[q] = rec[i * REC_FLOATS + q];
  float best = 0.f;
  int bj = -1, bi = -1;
  bool dynamic = false;
  for (int j = lane; j <= n; j += 32) {
    const int64_t c = (j == 0) ? (int64_t)n : (int64_t)adj_idx[off + j - 1];
    float* slot = wcache + wbase + j;
    float w = -1.0f;
    bool is_dynamic = false;
    if (c >= 0 && c < nu) {
      const uint32_t cc = cnt1[c];
      is_dynamic = cc == 1u && c < i;
      if (first) {
        if (cc > 1u || is_dynamic) {
          for (int t = 0; t < REC_FLOATS; t++) rc[t] = rec[c * REC_FLOATS + t];
          float w_ab, w_ba;
          pair_weights(ri, rc, pp, w_ab, w_ba);
          w = w_ab >= 0.f ? w_ab : -1.0f;            // a NaN weight never wins (`>=` is false)
        }
      } else {
        w = *slot;
      }
    }
    if (first) *slot = w;
    if (is_dynamic) dynamic = true;
    if (w < 0.f || (is_dynamic && ((volatile int32_t*)attach)[c] < 0)) continue;
    if (w >= best) { best = w; bj = j; bi = (int)c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ow = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oj >= 0 && (bj < 0 || ow > best || (ow == best && oj > bj))) { best = ow; bj = oj; bi = oi; }
  }
  const bool any_dynamic = __any_sync(0xffffffffu, dynamic);
  if (lane == 0) {
    if (bi != attach[i]) { attach[i] = bi; *changed = 1u; }
    if (first && any_dynamic) dep_list[atomicAdd(dep_count, 1u)] = li;
  }
}

// ---- stage 5d: connected components by lock-free union-find, root = smallest unit id ----
// find with path halving: a non-root's pointer only ever moves to one of its ancestors (roots are hooked under
// smaller roots only), so the plain store races benignly with the CAS in uf_union, which touches roots only
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = ((volatile int*)parent)[x];
  while (p != x) {
    const int gp = ((volatile int*)parent)[p];
    if (gp != p) ((volatile int*)parent)[x] = gp;
    x = p; p = gp;
  }
  return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}
__global__ void __launch_bounds__(256) k_iota(int* __restrict__ p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}
// initial forest: every unit points at its smallest linked unit with a smaller id (connect list / closest-check
// partner), which already performs one union per unit without any atomics.  One warp per unit.
__global__ void __launch_bounds__(128) k_cc_init(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ cnt1,
                                               const int32_t* __restrict__ idx1, const int32_t* __restrict__ attach, int64_t nu,
                                               int* __restrict__ parent, const uint8_t* __restrict__ own = nullptr) {
  const int lane = threadIdx.x & 31;
  int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  // slab tiles: only the links of OWNED voxels are trusted (halo voxels have truncated neighbourhoods); the partner
  // of an owned single may be named by global id (>= 0x40000000): that link is merged across slabs by key
  if (own && !(own[u] & 1)) { if (lane == 0) parent[u] = (int)u; return; }
  const uint32_t off = adj_off[u];
  const int c = (int)cnt1[u];
  int p = (int)u;
  for (int e = lane; e < c; e += 32) p = min(p, idx1[off + e]);
  p = __reduce_min_sync(0xffffffffu, p);
  if (lane == 0) {
    const int a = attach[u];
    if (a >= 0 && (!own || a < 0x40000000)) p = min(p, a);
    parent[u] = p;
  }
}
// pointer jumping over the initial forest (parents only move to ancestors, so in-place updates are safe)
__global__ void __launch_bounds__(256) k_cc_jump(int* parent, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = ((volatile int*)parent)[i];
  const int gp = ((volatile int*)parent)[p];
  if (gp != p) ((volatile int*)parent)[i] = gp;
}
__global__ void __launch_bounds__(128) k_cc_hook(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ cnt1,
                                               const int32_t* __restrict__ idx1, const int32_t* __restrict__ attach, int64_t nu,
                                               int* parent, const uint8_t* __restrict__ own = nullptr) {
  const int lane = threadIdx.x & 31;
  int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  if (own && !(own[u] & 1)) return;
  const uint32_t off = adj_off[u];
  const int c = (int)cnt1[u];
  // after the pointer jumping nearly every unit points at its root: two units with the same parent are already in
  // one tree, which one (coalesced + one random) load pair decides without walking the trees
  const int pu = ((volatile int*)parent)[u];
  for (int e = lane; e < c; e += 32) {
    int j = idx1[off + e];
    if (j > (int)u && ((volatile int*)parent)[j] != pu) uf_union(parent, (int)u, j);
  }
  if (lane == 0) { int a = attach[u]; if (a >= 0 && (!own || a < 0x40000000) && ((volatile int*)parent)[a] != pu) uf_union(parent, (int)u, a); }
}
__global__ void __launch_bounds__(256) k_cc_flatten(int* parent, int64_t n, int* __restrict__ root) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) root[i] = uf_find(parent, (int)i);
}

// ---- stage 6: cluster sizes / smallest point index, export filter, per-point canonical labels ----
__global__ void __launch_bounds__(256) k_cluster_stats(const int* __restrict__ root, const uint32_t* __restrict__ ustart,
                                                     const uint32_t* __restrict__ perm, int64_t nu, uint32_t* __restrict__ csize,
                                                     uint32_t* __restrict__ cminpt) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = u < nu;
  const int lane = threadIdx.x & 31;
  const int r = ok ? root[u] : -1 - lane;              // out-of-range lanes form singleton groups
  const uint32_t mp = ok ? perm[ustart[u]] : 0xffffffffu;   // the unit's points ascend by index
  // neighbouring units mostly share their cluster: one atomic pair per distinct root and warp
  const uint32_t peers = __match_any_sync(0xffffffffu, r);
  const uint32_t mn = __reduce_min_sync(peers, mp);
  if (ok && lane == __ffs(peers) - 1) {
    atomicAdd(&csize[r], (uint32_t)__popc(peers));
    atomicMin(&cminpt[r], mn);
  }
}
__global__ void __launch_bounds__(256) k_cluster_count(const int* __restrict__ root, const uint32_t* __restrict__ csize, int64_t nu,
                                                     int min_size_excl, unsigned long long* __restrict__ out2) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool is_root = u < nu && root[u] == (int)u;
  bool exp_ = is_root && (int)csize[u] > min_size_excl;
  uint32_t b1 = __ballot_sync(0xffffffffu, is_root), b2 = __ballot_sync(0xffffffffu, exp_);
  if ((threadIdx.x & 31) == 0) {
    if (b1) atomicAdd(&out2[0], (unsigned long long)__popc(b1));
    if (b2) atomicAdd(&out2[1], (unsigned long long)__popc(b2));
  }
}
__global__ void __launch_bounds__(256) k_point_labels(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ pos_unit,
                                                    const int* __restrict__ root, const uint32_t* __restrict__ csize,
                                                    const uint32_t* __restrict__ cminpt, int64_t n, int64_t n_valid,
                                                    int min_size_excl, int32_t* __restrict__ label, int32_t* __restrict__ point_unit) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t i = perm[p];
  int32_t lab = -1, pu = -1;
  if (p < n_valid) {
    uint32_t u = pos_unit[p];
    pu = (int32_t)u;
    if (root) {
      int r = root[u];
      if ((int)csize[r] > min_size_excl) lab = (int32_t)cminpt[r];
    }
  }
  if (label) label[i] = lab;
  if (point_unit) point_unit[i] = pu;
}

// ---- cluster export (drawColorMapofPointsinClusters VS.h:947-1014 / getClusterIdx VS.h:117) on the device:
//      exported clusters in ascending seed (= root) id, voxels ascending inside a cluster, points ascending inside a voxel ----
// flag of the exported roots (scanned into the cluster rank)
__global__ void __launch_bounds__(256) k_export_flags(const int* __restrict__ root, const uint32_t* __restrict__ csize, int64_t nu, int min_size_excl,
                                                    uint32_t* __restrict__ flag) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < nu) flag[u] = (root[u] == (int)u && (int)csize[u] > min_size_excl) ? 1u : 0u;
}
// sort key of a voxel = rank of its cluster (not exported: nclusters, sorts last); value = voxel id
__global__ void __launch_bounds__(256) k_export_keys(const int* __restrict__ root, const uint32_t* __restrict__ csize, const uint32_t* __restrict__ rank,
                                                   int64_t nu, int min_size_excl, uint32_t nclusters, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const int r = root[u];
  keys[u] = (int)csize[r] > min_size_excl ? rank[r] : nclusters;
  vals[u] = (uint32_t)u;
}
// number of points of the voxel at every sorted position (0 for the voxels that are not exported)
__global__ void __launch_bounds__(256) k_export_sizes(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svox, const uint32_t* __restrict__ ustart,
                                                    int64_t nu, uint32_t nclusters, uint32_t* __restrict__ sz) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nu) return;
  const uint32_t u = svox[p];
  sz[p] = skeys[p] < nclusters ? ustart[u + 1] - ustart[u] : 0u;
}
// one warp per sorted position: the voxel's point indices to their place; the first voxel of a cluster files its offset
__global__ void __launch_bounds__(128) k_export_points(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svox, const uint32_t* __restrict__ dst,
                                                     const uint32_t* __restrict__ ustart, const uint32_t* __restrict__ perm, int64_t nu,
                                                     uint32_t nclusters, long long* __restrict__ offsets, int32_t* __restrict__ point_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= nu) return;
  const uint32_t c = skeys[p];
  if (c >= nclusters) return;
  const uint32_t u = svox[p], d = dst[p], s = ustart[u], n = ustart[u + 1] - s;
  if (lane == 0 && (p == 0 || skeys[p - 1] != c)) offsets[c] = (long long)d;
  for (uint32_t i = lane; i < n; i += 32) point_idx[d + i] = (int32_t)perm[s + i];
}

}  // namespace vgs

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

// ---- stage 1a: octree key per point -> sortable 64-bit code, value = point index.
//      key.a = (unsigned)((p.a - min_a)/res) in double (genOctreeKeyforPoint).  12 B read, 12 B written. ----
__global__ void __launch_bounds__(256) k_quantise(const float* __restrict__ xyz, int stride, int64_t n, EpochTable ep, double res,
                                                int depth, int descending, uint64_t* __restrict__ keys,
                                                uint32_t* __restrict__ vals, uint32_t* __restrict__ key3_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + i * stride;
  float x = p[0], y = p[1], z = p[2];
  vals[i] = (uint32_t)i;
  const uint64_t mask = (depth * 3 >= 64) ? ~0ull : ((1ull << (3 * depth)) - 1ull);
  if (!finite3(x, y, z)) {
    keys[i] = 1ull << (3 * depth);  // sorts after every real key
    if (key3_out) { key3_out[3 * i] = key3_out[3 * i + 1] = key3_out[3 * i + 2] = 0xffffffffu; }
    return;
  }
  int e = ep.n - 1;
  while (e > 0 && i < ep.viol[e]) e--;
  uint32_t kx = (uint32_t)(((double)x - ep.mn[e][0]) / res) + ep.shift[e][0];
  uint32_t ky = (uint32_t)(((double)y - ep.mn[e][1]) / res) + ep.shift[e][1];
  uint32_t kz = (uint32_t)(((double)z - ep.mn[e][2]) / res) + ep.shift[e][2];
  uint64_t m = morton_encode(kx, ky, kz);
  keys[i] = descending ? (~m & mask) : m;
  if (key3_out) { key3_out[3 * i] = kx; key3_out[3 * i + 1] = ky; key3_out[3 * i + 2] = kz; }
}

// SVGS units: key = supervoxel label (SV.h:288-323), dropped labels sort last
__global__ void __launch_bounds__(256) k_label_keys(const int32_t* __restrict__ labels, const float* __restrict__ xyz, int stride, int64_t n,
                                                  int32_t max_label, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t l = labels[i];
  vals[i] = (uint32_t)i;
  bool ok = l > 0 && l < max_label;
  keys[i] = ok ? (uint64_t)(uint32_t)l : (1ull << 32);
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

// ---- stage 1b: segment heads of the sorted keys ----
__global__ void __launch_bounds__(256) k_head_flags(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
// scan[i] = exclusive scan of flags.  Writes unit start offsets, the unit's sort key and the
// sorted-position -> unit id map.
__global__ void __launch_bounds__(256) k_head_write(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flags,
                                                  const uint32_t* __restrict__ scan, int64_t n, uint32_t* __restrict__ ustart,
                                                  uint64_t* __restrict__ ukey, uint32_t* __restrict__ pos_unit) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t f = flags[i];
  uint32_t u = scan[i] + f - 1u;
  pos_unit[i] = u;
  if (f) { ustart[u] = (uint32_t)i; ukey[u] = keys[i]; }
}

// ---- stage 2: per-unit record (centroid, scatter, eigen33, normal, 8 eigen features).
//      One thread per unit, points visited in ascending index order so the fp32 sums are the
//      reference's sums bit for bit. ----
__global__ void __launch_bounds__(128) k_features(const float* __restrict__ xyz, int stride, const uint32_t* __restrict__ perm,
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

// voxel key + centre of every leaf (setVoxelCenters VS.h:146-189, getVoxelCenterFromOctreeKey
// VS.h:2102-2109): centre = (float)(((double)key + 0.5f) * res_f + min_f) with the FLOAT members
// the reference narrows in setVoxelSize / setBoundingBox (VS.h:127, 136-142, 1121-1123).
__global__ void __launch_bounds__(256) k_voxel_geometry(const uint64_t* __restrict__ ukey, int64_t nu, int depth, int descending,
                                                      float res_f, float mnx, float mny, float mnz,
                                                      uint32_t* __restrict__ key3, float* __restrict__ center) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  uint64_t m = ukey[u];
  const uint64_t mask = (1ull << (3 * depth)) - 1ull;
  if (descending) m = ~m & mask;
  uint32_t kx, ky, kz;
  morton_decode(m, kx, ky, kz);
  key3[3 * u] = kx; key3[3 * u + 1] = ky; key3[3 * u + 2] = kz;
  center[3 * u] = (float)(((double)kx + 0.5f) * res_f + mnx);
  center[3 * u + 1] = (float)(((double)ky + 0.5f) * res_f + mny);
  center[3 * u + 2] = (float)(((double)kz + 0.5f) * res_f + mnz);
}

// plain (non-complemented) morton of each voxel, the hash-table key
__global__ void __launch_bounds__(256) k_plain_morton(const uint32_t* __restrict__ key3, int64_t n, uint64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = morton_encode(key3[3 * i], key3[3 * i + 1], key3[3 * i + 2]);
}

// ---- stage 3 (VGS): FLANN radius search over voxel centres == lattice stencil probed through the
//      hash table, then the float test dist2 < (float)(r*r) and ordering by (dist2, id)
//      (VS.h:223-265, FLANN L2_Simple + RadiusResultSet).  One warp per voxel.
//      fill == 0: write the neighbour count; fill == 1: write the ordered list at adj_off[v];
//      fill == 2 (one probing pass instead of two): write the count AND the ordered list into a fixed-stride
//      staging row (adj_idx + v * cap); k_adjacency_compact moves the rows to their CSR offsets after the scan. ----
__global__ void __launch_bounds__(128) k_adjacency(const uint32_t* __restrict__ key3, const float* __restrict__ center, int64_t nv,
                                                 int depth, const int4* __restrict__ stencil, int nst,
                                                 const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                 uint64_t mask, float r2, int fill, uint32_t* __restrict__ adj_cnt,
                                                 const uint32_t* __restrict__ adj_off, int32_t* __restrict__ adj_idx, int cap) {
  extern __shared__ unsigned char smraw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sd2 = reinterpret_cast<float*>(smraw) + (size_t)w * cap;
  int* sid = reinterpret_cast<int*>(smraw + (size_t)wpb * cap * sizeof(float)) + (size_t)w * cap;
  int64_t v = (int64_t)blockIdx.x * wpb + w;
  if (v >= nv) return;
  const uint32_t kx = key3[3 * v], ky = key3[3 * v + 1], kz = key3[3 * v + 2];
  const float qx = center[3 * v], qy = center[3 * v + 1], qz = center[3 * v + 2];
  const int64_t lim = 1ll << depth;
  int count = 0;
  for (int b = 0; b < nst; b += 32) {
    int s = b + lane;
    int id = -1;
    float d2 = 0.f;
    if (s < nst) {
      int4 o = stencil[s];
      int64_t x = (int64_t)kx + o.x, y = (int64_t)ky + o.y, z = (int64_t)kz + o.z;
      if (x >= 0 && y >= 0 && z >= 0 && x < lim && y < lim && z < lim) {
        id = hash_lookup(tk, tv, mask, morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z));
        if (id >= 0) {
          float dx = qx - center[3 * (int64_t)id], dy = qy - center[3 * (int64_t)id + 1], dz = qz - center[3 * (int64_t)id + 2];
          d2 = 0.f; d2 += dx * dx; d2 += dy * dy; d2 += dz * dz;
          if (!(d2 < r2)) id = -1;
        }
      }
    }
    uint32_t bal = __ballot_sync(0xffffffffu, id >= 0);
    if (fill && id >= 0) {
      int pos = count + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) { sd2[pos] = d2; sid[pos] = id; }
    }
    count += __popc(bal);
  }
  if (fill != 1 && lane == 0) adj_cnt[v] = (uint32_t)count;
  if (!fill) return;
  __syncwarp();
  if (count > cap) count = cap;
  if (fill == 2) adj_idx += v * cap;
  const uint32_t off = fill == 2 ? 0u : adj_off[v];
  for (int e = lane; e < count; e += 32) {
    float d = sd2[e]; int id = sid[e];
    int rank = 0;
    for (int j = 0; j < count; j++) {
      float dj = sd2[j]; int ij = sid[j];
      rank += (dj < d || (dj == d && ij < id)) ? 1 : 0;
    }
    adj_idx[off + rank] = id;
  }
}

__global__ void __launch_bounds__(256) k_adjacency_compact(const int32_t* __restrict__ stage, int cap, const uint32_t* __restrict__ adj_off,
                                                         int64_t nv, int32_t* __restrict__ adj_idx) {
  const int lane = threadIdx.x & 31;
  int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= nv) return;
  const uint32_t off = adj_off[v];
  const int c = (int)(adj_off[v + 1] - off);
  for (int e = lane; e < c; e += 32) adj_idx[off + e] = stage[v * cap + e];
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
// one warp per unit; cstart/cunits = cell table (units sorted by cell key); fill as in k_adjacency
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
  int count = 0;
  for (int d = 0; d < 27; d++) {
    const int x = cx + d / 9 - 1, y = cy + (d / 3) % 3 - 1, z = cz + d % 3 - 1;
    if (x < 0 || y < 0 || z < 0) continue;
    const int c = hash_lookup(tk, tv, mask, morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z));
    if (c < 0) continue;
    const uint32_t s = cstart[c], e = cstart[c + 1];
    for (uint32_t b = s; b < e; b += 32) {
      const uint32_t i = b + lane;
      int id = -1;
      float d2 = 0.f;
      if (i < e) {
        id = (int)cunits[i];
        const float* t = rec + (int64_t)id * REC_FLOATS;
        const float dx = qx - t[0], dy = qy - t[1], dz = qz - t[2];
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
                                                   unsigned long long* __restrict__ stats /* [0]=sum n(n-1) [1]=max n [2]=overflow */,
                                                   uint8_t* __restrict__ need_rows) {
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
      if (need_rows && us) need_rows[g] = 1;
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

// ---- stage 4 (VGS, cached): every unordered pair of USED voxels closer than two stencil radii is
//      evaluated ONCE and stored by lattice offset: table[a*half + code(key_b - key_a)] =
//      (w(a->b), w(b->a)), a = the voxel whose offset to b is lexicographically positive.
//      A pair's weight does not depend on the local graph it appears in (buildAdjacencyGraph
//      VS.h:1796-1910 recomputes it for every centre voxel).  One warp per voxel: hash probes of
//      the offset list, hits queued in shared memory, 32 pair evaluations per warp step. ----
__global__ void __launch_bounds__(128) k_pair_cache(const uint32_t* __restrict__ key3, const float* __restrict__ rec, int64_t nv,
                                                  int depth, const int4* __restrict__ st2, int nst2,
                                                  const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                  uint64_t mask, PairParams pp, float2* __restrict__ table, int half,
                                                  const uint8_t* __restrict__ need_rows, const uint8_t* __restrict__ uflags) {
  __shared__ int pend_b[4][64];
  __shared__ int pend_i[4][64];
  __shared__ float s_ra[4][REC_FLOATS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t v = (int64_t)blockIdx.x * 4 + w;
  if (v >= nv) return;
  if (!(uflags[v] & F_USED)) return;
  if (need_rows && !need_rows[v]) return;   // multi-GPU: only rows read by this rank's local graphs
  if (lane < REC_FLOATS) s_ra[w][lane] = rec[v * REC_FLOATS + lane];
  __syncwarp();
  const uint32_t kx = key3[3 * v], ky = key3[3 * v + 1], kz = key3[3 * v + 2];
  const int64_t lim = 1ll << depth;
  int npend = 0;
  auto process = [&](int cnt) {
    if (lane < cnt) {
      const int b = pend_b[w][lane], idx = pend_i[w][lane];
      float rb[REC_FLOATS];
      const float4* src = reinterpret_cast<const float4*>(rec + (int64_t)b * REC_FLOATS);
#pragma unroll
      for (int q = 0; q < 4; q++) { float4 t = __ldg(src + q); rb[4 * q] = t.x; rb[4 * q + 1] = t.y; rb[4 * q + 2] = t.z; rb[4 * q + 3] = t.w; }
      float w_ab, w_ba;
      pair_weights(s_ra[w], rb, pp, w_ab, w_ba);
      table[(size_t)v * half + idx] = make_float2(w_ab, w_ba);
    }
  };
  for (int base = 0; base < nst2; base += 32) {
    const int s = base + lane;
    int id = -1, code = 0;
    if (s < nst2) {
      int4 o = st2[s];
      code = o.w;
      int64_t x = (int64_t)kx + o.x, y = (int64_t)ky + o.y, z = (int64_t)kz + o.z;
      if (x >= 0 && y >= 0 && z >= 0 && x < lim && y < lim && z < lim) {
        id = hash_lookup(tk, tv, mask, morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z));
        if (id >= 0 && !(__ldg(uflags + id) & F_USED)) id = -1;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, id >= 0);
    if (id >= 0) { int pos = npend + __popc(bal & ((1u << lane) - 1u)); pend_b[w][pos] = id; pend_i[w][pos] = code; }
    npend += __popc(bal);
    __syncwarp();
    if (npend >= 32) {
      process(32);
      __syncwarp();
      const int rem = npend - 32;
      int tb = 0, ti = 0;
      if (lane < rem) { tb = pend_b[w][32 + lane]; ti = pend_i[w][32 + lane]; }
      __syncwarp();
      if (lane < rem) { pend_b[w][lane] = tb; pend_i[w][lane] = ti; }
      __syncwarp();
      npend = rem;
    }
  }
  process(npend);
}

// ---- occupancy bitmap of the used voxels over the octree key cube: bit ((x << depth | y) << depth | z).
//      A z-run of the lattice is one or two words, so a stencil column costs one load instead of one hash
//      probe per offset (most offsets miss: 22 % of the pair-cache stencil is occupied on the 10 M scene). ----
__global__ void __launch_bounds__(256) k_bitmap_set(const uint32_t* __restrict__ key3, const uint8_t* __restrict__ uflags, int64_t nv,
                                                  int depth, uint32_t* __restrict__ bm) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv || (uflags && !(uflags[v] & F_USED))) return;   // uflags == null: every voxel
  const uint64_t b = ((((uint64_t)key3[3 * v] << depth) | key3[3 * v + 1]) << depth) | key3[3 * v + 2];
  atomicOr(&bm[b >> 5], 1u << (b & 31));
}

// k_adjacency (fill == 2) with the lattice search done on the all-voxel bitmap: one load per stencil column,
// hash lookups only for occupied cells.  Same float test, same (dist2, id) order.
__global__ void __launch_bounds__(128) k_adjacency_bm(const uint32_t* __restrict__ key3, const float* __restrict__ center, int64_t nv,
                                                    int depth, const int4* __restrict__ cols, int ncol, int rho, const uint32_t* __restrict__ bm,
                                                    const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                    uint64_t mask, float r2, uint32_t* __restrict__ adj_cnt, int32_t* __restrict__ stage, int cap) {
  extern __shared__ unsigned char smraw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sd2 = reinterpret_cast<float*>(smraw) + (size_t)w * cap;
  int* sid = reinterpret_cast<int*>(smraw + (size_t)wpb * cap * sizeof(float)) + (size_t)w * cap;
  unsigned short* q = reinterpret_cast<unsigned short*>(smraw + (size_t)wpb * cap * 8) + (size_t)w * cap;
  int64_t v = (int64_t)blockIdx.x * wpb + w;
  if (v >= nv) return;
  const int kx = (int)key3[3 * v], ky = (int)key3[3 * v + 1], kz = (int)key3[3 * v + 2];
  const float qx = center[3 * v], qy = center[3 * v + 1], qz = center[3 * v + 2];
  const int lim = 1 << depth;
  const int S = 2 * rho + 1;
  int nq = 0;
  for (int base = 0; base < ncol; base += 32) {
    const int ci = base + lane;
    uint32_t hits = 0;
    int cbase = 0;
    if (ci < ncol) {
      const int4 o = cols[ci];
      const int x = kx + o.x, y = ky + o.y;
      if (x >= 0 && y >= 0 && x < lim && y < lim) {
        const int z0 = max(kz - rho, 0), z1 = min(kz + rho, lim - 1);
        const uint64_t b0 = ((((uint64_t)x << depth) | (uint64_t)y) << depth) | (uint64_t)z0;
        const uint64_t two = (uint64_t)__ldg(bm + (b0 >> 5)) | ((uint64_t)__ldg(bm + (b0 >> 5) + 1) << 32);
        uint32_t run = (uint32_t)(two >> (b0 & 31)) & (uint32_t)((1ull << (z1 - z0 + 1)) - 1ull);
        run <<= (z0 - (kz - rho));
        hits = run & (uint32_t)o.z;
        cbase = ((o.x + rho) * S + (o.y + rho)) * S;
      }
    }
    const int cnt = __popc(hits);
    const int incl = (int)warp_incl_scan((unsigned)cnt, lane);
    int pos = nq + incl - cnt;
    while (hits) {
      const int j = __ffs(hits) - 1;
      hits &= hits - 1;
      q[pos++] = (unsigned short)(cbase + j);
    }
    nq += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  int count = 0;
  for (int b = 0; b < nq; b += 32) {
    const int e = b + lane;
    int id = -1;
    float d2 = 0.f;
    if (e < nq) {
      const int c = q[e];
      const int dz = c % S - rho, dy = (c / S) % S - rho, dx = c / (S * S) - rho;
      id = hash_lookup(tk, tv, mask, morton_encode((uint32_t)(kx + dx), (uint32_t)(ky + dy), (uint32_t)(kz + dz)));
      if (id >= 0) {
        float ex = qx - center[3 * (int64_t)id], ey = qy - center[3 * (int64_t)id + 1], ez = qz - center[3 * (int64_t)id + 2];
        d2 = 0.f; d2 += ex * ex; d2 += ey * ey; d2 += ez * ez;
        if (!(d2 < r2)) id = -1;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, id >= 0);
    if (id >= 0) {
      const int pos = count + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) { sd2[pos] = d2; sid[pos] = id; }
    }
    count += __popc(bal);
  }
  if (lane == 0) adj_cnt[v] = (uint32_t)count;
  __syncwarp();
  if (count > cap) count = cap;
  stage += v * cap;
  for (int e = lane; e < count; e += 32) {
    const float d = sd2[e]; const int id = sid[e];
    int rank = 0;
    for (int j = 0; j < count; j++) {
      const float dj = sd2[j]; const int ij = sid[j];
      rank += (dj < d || (dj == d && ij < id)) ? 1 : 0;
    }
    stage[rank] = id;
  }
}

// k_pair_cache with the partner search done on the bitmap: the stencil is stored as columns (dx, dy, 13-bit mask
// of dz), a lane reads the z-run of its column, the set bits are queued as table codes (the code encodes the
// offset), and 32 queued pairs are evaluated per warp step (the hash table only resolves the ids of real partners).
constexpr int PC_QCAP = 32 + 32 * 13;
__global__ void __launch_bounds__(128) k_pair_cache_bm(const uint32_t* __restrict__ key3, const float* __restrict__ rec, int64_t nv,
                                                     int depth, const int4* __restrict__ cols, int ncol, int r2, const uint32_t* __restrict__ bm,
                                                     const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                     uint64_t mask, PairParams pp, float2* __restrict__ table, int half,
                                                     const uint8_t* __restrict__ need_rows, const uint8_t* __restrict__ uflags) {
  __shared__ unsigned short pend[4][PC_QCAP];
  __shared__ float s_ra[4][REC_FLOATS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t v = (int64_t)blockIdx.x * 4 + w;
  if (v >= nv) return;
  if (!(uflags[v] & F_USED)) return;
  if (need_rows && !need_rows[v]) return;   // multi-GPU: only rows read by this rank's local graphs
  if (lane < REC_FLOATS) s_ra[w][lane] = rec[v * REC_FLOATS + lane];
  __syncwarp();
  const int kx = (int)key3[3 * v], ky = (int)key3[3 * v + 1], kz = (int)key3[3 * v + 2];
  const int lim = 1 << depth;
  const int S = 2 * r2 + 1;
  int npend = 0;
  auto process = [&](int first, int cnt) {
    if (lane < cnt) {
      const int code = pend[w][first + lane];
      const int c = code + half + 1;                 // ((dx+r2)*S + (dy+r2))*S + (dz+r2)
      const int dz = c % S - r2, dy = (c / S) % S - r2, dx = c / (S * S) - r2;
      const int b = hash_lookup(tk, tv, mask, morton_encode((uint32_t)(kx + dx), (uint32_t)(ky + dy), (uint32_t)(kz + dz)));
      float rb[REC_FLOATS];
      const float4* src = reinterpret_cast<const float4*>(rec + (int64_t)b * REC_FLOATS);
#pragma unroll
      for (int q = 0; q < 4; q++) { float4 t = __ldg(src + q); rb[4 * q] = t.x; rb[4 * q + 1] = t.y; rb[4 * q + 2] = t.z; rb[4 * q + 3] = t.w; }
      float w_ab, w_ba;
      pair_weights(s_ra[w], rb, pp, w_ab, w_ba);
      table[(size_t)v * half + code] = make_float2(w_ab, w_ba);
    }
  };
  for (int base = 0; base < ncol; base += 32) {
    const int ci = base + lane;
    uint32_t hits = 0;
    int cbase = 0;
    if (ci < ncol) {
      const int4 o = cols[ci];
      const int x = kx + o.x, y = ky + o.y;
      if (x >= 0 && y >= 0 && x < lim && y < lim) {
        const int z0 = max(kz - r2, 0), z1 = min(kz + r2, lim - 1);
        const uint64_t b0 = ((((uint64_t)x << depth) | (uint64_t)y) << depth) | (uint64_t)z0;
        const uint64_t two = (uint64_t)__ldg(bm + (b0 >> 5)) | ((uint64_t)__ldg(bm + (b0 >> 5) + 1) << 32);
        uint32_t run = (uint32_t)(two >> (b0 & 31)) & ((1u << (z1 - z0 + 1)) - 1u);
        run <<= (z0 - (kz - r2));                    // bit j <-> dz = j - r2
        hits = run & (uint32_t)o.z;
        cbase = ((o.x + r2) * S + (o.y + r2)) * S - half - 1;
      }
    }
    const int cnt = __popc(hits);
    const int incl = (int)warp_incl_scan((unsigned)cnt, lane);
    int pos = npend + incl - cnt;
    while (hits) {
      const int j = __ffs(hits) - 1;
      hits &= hits - 1;
      pend[w][pos++] = (unsigned short)(cbase + j);
    }
    npend += __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    while (npend >= 32) { process(npend - 32, 32); npend -= 32; }
    __syncwarp();
  }
  process(0, npend);
}

// ---- stage 4+5a: local affinity graph + Felzenszwalb-style cut of ONE unit per CTA
//      (buildAdjacencyGraph VS.h:1796-1910 + cutGraphSegmentation VS.h:1913-2029).
//      1. directed weights of all pairs of the neighbourhood: CACHED -> one 8-byte load per unordered
//         pair from the offset-indexed table; else evaluated here from the records.  Weights
//         w <= 1-2k+k/n can never merge (cut bound) and are dropped.
//      2. the reference sorts all n^2 weights; here a 256-bin histogram of the weights delimits
//         chunks of ~512 entries that are sorted (bitonic, shared memory) and merged in descending
//         order (w desc, flat index asc) until exactly nothing more can merge: one segment left, or
//         the next weight does not exceed the smallest live threshold Int(C) - k/|C|.
//      3. the merge scans 32 sorted entries per warp step: the first mergeable entry merges, the
//         later ones are re-evaluated against the new state. ----
struct GraphParams {
  PairParams pp;
  float cut;
};

template <int THREADS, bool CACHED>
__global__ void __launch_bounds__(THREADS) k_local_graph2(const uint32_t* __restrict__ list, uint32_t nlist,
                                                        const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                        const float* __restrict__ rec, const uint32_t* __restrict__ key3,
                                                        GraphParams gp, int ncap, int mcap, const float2* __restrict__ table,
                                                        int half, int r2, const float* __restrict__ wempty, int bucketed,
                                                        uint32_t* __restrict__ conn_cnt, int32_t* __restrict__ conn_idx) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int AUX = CACHED ? 4 : REC_PAD;
  float* A_w = reinterpret_cast<float*>(smraw);                                  // mcap: weight pool (append order)
  float* C_w = A_w + mcap;                                                       // LG_CS: sorted chunk
  float* s_int = C_w + LG_CS;                                                    // ncap
  int* s_gid = reinterpret_cast<int*>(s_int + ncap);                             // ncap
  int* s_aux = s_gid + ncap;                                                     // AUX*ncap: (kx,ky,kz,flags) | records
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
  if (CACHED) {
    for (int i = tid; i < n; i += THREADS) {
      const int64_t g = s_gid[i];
      s_aux[4 * i] = (int)key3[3 * g]; s_aux[4 * i + 1] = (int)key3[3 * g + 1]; s_aux[4 * i + 2] = (int)key3[3 * g + 2];
      s_aux[4 * i + 3] = f2i(__ldg(rec + g * REC_FLOATS + REC_FLAGS));
    }
  } else {
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
      if (v < n) us = ((CACHED ? s_aux[4 * v + 3] : s_aux[v * REC_PAD + REC_FLAGS]) & F_USED) != 0;
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
      if (!((CACHED ? s_aux[4 * i + 3] : s_aux[i * REC_PAD + REC_FLAGS]) & F_USED)) s_size[i] = 0;
    if (tid == 0) { s_nseg = nv; if (nv <= 1) s_done = 1; }
  }
  for (int p = tid; p < npairs; p += THREADS) {
    int r = p / (nv - 1), c = p - r * (nv - 1);
    int a, b;
    if (c < nv - 1 - r) { a = r; b = r + 1 + c; }
    else { a = nv - 1 - r; b = a + 1 + (c - (nv - 1 - r)); }
    if (!all_pairs) { a = s_ul[a]; b = s_ul[b]; }
    float w_ab, w_ba;
    if (CACHED) {
      const int fa = s_aux[4 * a + 3], fb = s_aux[4 * b + 3];
      if (!(fa & F_USED) || !(fb & F_USED)) { w_ab = w_ba = s_wempty; }
      else {
        int dx = s_aux[4 * b] - s_aux[4 * a], dy = s_aux[4 * b + 1] - s_aux[4 * a + 1], dz = s_aux[4 * b + 2] - s_aux[4 * a + 2];
        const bool pos = dx > 0 || (dx == 0 && (dy > 0 || (dy == 0 && dz > 0)));
        if (!pos) { dx = -dx; dy = -dy; dz = -dz; }
        const int S = 2 * r2 + 1;
        const int code = ((dx + r2) * S + (dy + r2)) * S + (dz + r2) - half - 1;
        const float2 e = __ldg(table + (size_t)(pos ? s_gid[a] : s_gid[b]) * half + code);
        w_ab = pos ? e.x : e.y;
        w_ba = pos ? e.y : e.x;
      }
    } else {
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
            for (int v = lane; v < n; v += 32) if (s_seg[v] == drop) s_seg[v] = (unsigned short)keep;
            if (lane == 0) { s_int[keep] = wl; s_size[keep] = (unsigned short)(s_size[keep] + s_size[drop]); s_size[drop] = 0; }
            nseg--;
            __syncwarp();
            todo &= ~((2u << Lm) - 1u);
            if (nseg <= 1) { stop = true; break; }
          }
        }
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

// ---- stage 4+5a, warp-per-unit variant (VGS with the pair cache): same algorithm and results as
//      k_local_graph2, restructured so that nothing waits at a block barrier (the ncu profile of the CTA
//      version showed 37 % of all stall samples at the barrier behind warp 0's serial merge) and so that a
//      unit needs ~8 KB of shared memory instead of 30-40 KB: the weights are NOT kept — pass 0 reads every
//      pair once from the offset-indexed table and stores only its histogram bin (1 byte per directed
//      entry); each chunk then re-reads just the entries of its bin range.  Units whose weight
//      distribution defeats this (one bin larger than the staging buffer, or empty-pair weights that can
//      merge) are appended to `fallback` and handled by the CTA kernel. ----
constexpr int LW_CS = 256;      // staging capacity (entries) per warp
constexpr int LW_CH = 28;       // target number of USEFUL entries per chunk (one per lane: register sort)
constexpr int LW_WARPS = 4;     // units per CTA
constexpr int LW_ILP = 2;       // table gathers in flight per lane in pass A (4 measured the same)
constexpr int LW_BINS = 255;    // real bins 0..254; bin value 255 marks a dropped entry
__host__ __device__ inline size_t lw_slice_bytes(int ncap, int mcap) {
  (void)mcap;   // the bin-ordered entry list lives in a global scratch slice (written once, read once: L2)
  size_t b = (size_t)LW_CS * 4 + (size_t)ncap * 12 + 512 * 4 + (size_t)LW_CS * 2 + (size_t)ncap * 2 + 16;
  return (b + 15) & ~(size_t)15;
}

// min CTAs per SM = 8 caps the kernel at 64 registers; 48 registers / 10 CTAs measured 8 % slower, 6-7 CTAs the same
__global__ void __launch_bounds__(LW_WARPS * 32, 8) k_local_graph_warp(const uint32_t* __restrict__ list, uint32_t nlist,
                                                                  const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                                  const uint8_t* __restrict__ uflags, const uint32_t* __restrict__ key3,
                                                                  float k, int ncap, int mcap, const float2* __restrict__ table, int half,
                                                                  int r2, const float* __restrict__ wempty, uint32_t* __restrict__ conn_cnt,
                                                                  int32_t* __restrict__ conn_idx, uint32_t* __restrict__ fallback,
                                                                  uint32_t* __restrict__ fallback_count,
                                                                  unsigned short* __restrict__ scratch, int chunk_target,
                                                                  unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const uint32_t li = blockIdx.x * LW_WARPS + wq;
  if (li >= nlist) return;
  unsigned char* base = smraw + (size_t)wq * lw_slice_bytes(ncap, mcap);
  float* C_w = reinterpret_cast<float*>(base);                       // LW_CS
  float* s_thr = C_w + LW_CS;                                        // ncap: merge threshold Int(C) - k/|C| of segment C
  int* s_gid = reinterpret_cast<int*>(s_thr + ncap);                 // ncap
  int* s_key = s_gid + ncap;                                         // ncap: (dx+64) | (dy+64)<<8 | (dz+64)<<16 relative to the centre
  unsigned* s_hist = reinterpret_cast<unsigned*>(s_key + ncap);      // 256: histogram, then inclusive prefix sums
  unsigned* s_cur = s_hist + 256;                                    // 256: scatter cursors
  unsigned short* C_f = reinterpret_cast<unsigned short*>(s_cur + 256);   // LW_CS
  // global scratch slice of this unit (written once, read once: L2): mcap entry codes ordered by bin,
  // then mcap bytes holding the histogram bin of every directed entry in enumeration order
  unsigned short* ids = scratch + (size_t)li * (size_t)(mcap + mcap / 2);
  unsigned char* gbins = reinterpret_cast<unsigned char*>(ids + mcap);
  unsigned char* s_seg = reinterpret_cast<unsigned char*>(C_f + LW_CS);   // ncap
  unsigned char* s_size = s_seg + ncap;                              // ncap
  const uint32_t lt = (1u << lane) - 1u;

  const uint32_t u = list[li];
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  const int cx = (int)key3[3 * (int64_t)u], cy = (int)key3[3 * (int64_t)u + 1], cz = (int)key3[3 * (int64_t)u + 2];
  for (int i = lane; i < 256; i += 32) s_hist[i] = 0;
  const int S = 2 * r2 + 1;
  int nv = 0;
  // vertex table of the USED neighbours only, in adjacency order (local vertex 0 = the unit itself): unused voxels
  // cannot merge (their weight is w_empty <= cut bound, else the unit goes to the general kernel), and the
  // tie-break order (col * n + row over all neighbours, VS.h:1922) is preserved by the monotone renumbering
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int i = b0 + lane;
    bool us = false;
    int64_t g = 0;
    if (i < n) {
      g = adj_idx[off + i];
      us = (__ldg(uflags + g) & F_USED) != 0;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, us);
    if (us) {
      const int j = nv + __popc(bal & lt);
      s_gid[j] = (int)g;
      // mixed-radix lattice code of the offset to the centre: code(b) - code(a) orders pairs lexicographically
      s_key[j] = (((int)key3[3 * g] - cx) * S + ((int)key3[3 * g + 1] - cy)) * S + ((int)key3[3 * g + 2] - cz);
      s_seg[j] = (unsigned char)j; s_size[j] = 1; s_thr[j] = 1.0f - k / 1.0f;
    }
    nv += __popc(bal);
  }
  __syncwarp();
  const float lb = (float)(1.0 - 2.0 * (double)k + (double)k / (double)n - 4e-7 * (double)(n + 8));
  const float scale = (float)LW_BINS / fmaxf(1.0f - lb, 1e-3f);
  bool to_fallback = wempty[0] > lb;     // empty pairs could merge: the general kernel enumerates them
  // table slot of the unordered pair {a,b}: row = the vertex whose offset to the other is lexicographically
  // positive, column = |code(b) - code(a)| - 1  (== ((dx+r2)*S + (dy+r2))*S + (dz+r2) - half - 1 of k_pair_cache)
  auto fetch = [&](int a, int b, float& w_ab, float& w_ba) {
    const int diff = s_key[b] - s_key[a];
    const bool pos = diff > 0;
    const float2 e = __ldg(table + (size_t)(pos ? s_gid[a] : s_gid[b]) * half + (pos ? diff : -diff) - 1);
    w_ab = pos ? e.x : e.y;
    w_ba = pos ? e.y : e.x;
  };
  int nseg = nv;
  bool stop = false;
  // one batch of <= 32 entries in descending order, one per lane: the first mergeable entry merges,
  // the later ones are re-evaluated against the new state (cutGraphSegmentation VS.h:1955-2001)
  auto merge_batch = [&](float w, int f, bool valid) {
    const int v1 = f >> 8, v2 = f & 255;
    uint32_t todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
      bool pred = false, a_wins = true;
      int sa = 0, sb = 0;
      if ((todo >> lane) & 1u) {
        sa = s_seg[v1]; sb = s_seg[v2];
        if (sa != sb) {
          const float m1 = s_thr[sa], m2 = s_thr[sb];     // Int(C) - k / |C|, kept up to date at every merge
          a_wins = (m1 >= m2);
          pred = w > (a_wins ? m1 : m2);
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, pred);
      if (!bal) break;
      const int Lm = __ffs(bal) - 1;
      const int keepl = __shfl_sync(0xffffffffu, a_wins ? sa : sb, Lm);
      const int drop = __shfl_sync(0xffffffffu, a_wins ? sb : sa, Lm);
      const float wl = __shfl_sync(0xffffffffu, w, Lm);
      for (int v = lane; v < nv; v += 32) if (s_seg[v] == drop) s_seg[v] = (unsigned char)keepl;
      if (lane == 0) {
        const int nsz = (int)s_size[keepl] + (int)s_size[drop];
        s_thr[keepl] = wl - k / (float)nsz; s_size[keepl] = (unsigned char)nsz; s_size[drop] = 0;
      }
      nseg--;
      __syncwarp();
      todo &= ~((2u << Lm) - 1u);
      if (nseg <= 1) { stop = true; break; }
    }
  };
  if (!to_fallback && nv > 1) {
    // --- pass A: histogram of the kept weights (the weights themselves are not stored) ---
    // pairs (ia < ib) of used-list positions in row-major order; each lane steps 32 pairs at a time
    // (row wrap by subtraction, no division).  The bins are parked in the scratch slice so that pass B
    // does not have to gather from the table again.
    {
      int ia = 0, rem = lane, p = lane;    // rem = offset inside row ia, row length nv-1-ia
      while (ia < nv - 1 && rem >= nv - 1 - ia) { rem -= nv - 1 - ia; ia++; }
      // LW_ILP pairs in flight per lane: the table gathers are dependent-latency bound (ncu: long_scoreboard)
      while (ia < nv - 1) {
        int va[LW_ILP], vc[LW_ILP];
        bool ok[LW_ILP];
#pragma unroll
        for (int q = 0; q < LW_ILP; q++) {
          ok[q] = ia < nv - 1;
          va[q] = 0; vc[q] = 0;
          if (ok[q]) {
            va[q] = ia; vc[q] = ia + 1 + rem;
            rem += 32;
            while (ia < nv - 1 && rem >= nv - 1 - ia) { rem -= nv - 1 - ia; ia++; }
          }
        }
        float wf[LW_ILP], wr[LW_ILP];
#pragma unroll
        for (int q = 0; q < LW_ILP; q++) {
          wf[q] = 0.f; wr[q] = 0.f;
          if (ok[q]) fetch(va[q], vc[q], wf[q], wr[q]);
        }
#pragma unroll
        for (int q = 0; q < LW_ILP; q++) {
          if (ok[q]) {
            int b0 = 255, b1 = 255;
            if (wf[q] > lb) { b0 = min(LW_BINS - 1, (int)((1.0f - wf[q]) * scale)); atomicAdd(&s_hist[b0], 1u); }
            if (wr[q] > lb) { b1 = min(LW_BINS - 1, (int)((1.0f - wr[q]) * scale)); atomicAdd(&s_hist[b1], 1u); }
            reinterpret_cast<unsigned short*>(gbins)[p + 32 * q] = (unsigned short)(b0 | (b1 << 8));
          }
        }
        p += 32 * LW_ILP;
      }
    }
    __syncwarp();
    {   // inclusive prefix sums over the 256 counters; cursors = exclusive starts
      unsigned loc[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { sum += s_hist[lane * 8 + q]; loc[q] = sum; }
      const unsigned inc = warp_incl_scan(sum, lane);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const unsigned cnt = q ? loc[q] - loc[q - 1] : loc[0];
        s_hist[lane * 8 + q] = inc - sum + loc[q];
        s_cur[lane * 8 + q] = inc - sum + loc[q] - cnt;
      }
    }
    __syncwarp();
    // --- pass B: entries ordered by bin; an entry is stored as (ia << 8) | (ib << 1) | dir with ia < ib
    //     positions in the used list, dir 0 = a->b, 1 = b->a ---
    {
      int ia = 0, rem = lane, p = lane;
      while (ia < nv - 1 && rem >= nv - 1 - ia) { rem -= nv - 1 - ia; ia++; }
      const int np = nv * (nv - 1) / 2;
      int bb_n = p < np ? (int)reinterpret_cast<const unsigned short*>(gbins)[p] : 0;
      while (p < np) {   // the parked bins are read one step ahead of their use
        const int ib = ia + 1 + rem;
        const int bb = bb_n;
        const int code = (ia << 8) | (ib << 1);
        rem += 32; p += 32;
        bb_n = p < np ? (int)reinterpret_cast<const unsigned short*>(gbins)[p] : 0;
        while (ia < nv - 1 && rem >= nv - 1 - ia) { rem -= nv - 1 - ia; ia++; }
        const int b0 = bb & 255, b1 = bb >> 8;
        if (b0 != 255) ids[atomicAdd(&s_cur[b0], 1u)] = (unsigned short)code;
        if (b1 != 255) ids[atomicAdd(&s_cur[b1], 1u)] = (unsigned short)(code | 1);
      }
    }
    __syncwarp();
    const int m = (int)s_hist[255];
    float ratio = 1.0f;
    int c0 = 0;
    bool done = false;
    while (!done && c0 < LW_BINS && m > 0) {
      // --- chunk boundary: bins [c0, c1) with about LW_CH / ratio entries ---
      // (prefix sums are monotone: a boundary is a count of bins, found by all lanes together)
      const int bef = c0 > 0 ? (int)s_hist[c0 - 1] : 0;
      const int budget = bef + min(LW_CS, max(1, (int)((float)chunk_target / ratio)));   // never more than the staging buffer
      int nA = 0, nB = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int bq = lane * 8 + q;
        const int pv = (int)s_hist[bq];
        const bool in = bq >= c0 && bq < LW_BINS;
        nA += (in && pv <= budget) ? 1 : 0;     // bins whose cumulative count still fits the budget
        nB += (in && pv <= bef) ? 1 : 0;        // leading empty bins
      }
      nA = __reduce_add_sync(0xffffffffu, nA);
      nB = __reduce_add_sync(0xffffffffu, nB);
      int c1 = c0 + nA;
      if (nA <= nB) c1 = min(LW_BINS, c0 + nB + 1);   // the next non-empty bin alone exceeds the budget: take it alone
      const int tot = (int)s_hist[c1 - 1] - bef;
      if (tot == 0) break;                 // nothing left
      if (tot > LW_CS) { to_fallback = true; break; }   // one bin alone overflows the staging buffer
      const int cntE = tot;               // the chunk is the contiguous id range [bef, bef + tot)
      // --- fetch the still-useful ones (endpoints in different segments) ---
      int kept = 0;
      float rw = -1.0f;
      int rf = 0xffff;
      // phase 1: codes of the still-useful entries, compacted; the entry list is read one step ahead
      {
        int e_nxt = lane < cntE ? (int)ids[bef + lane] : 0;
        for (int x0 = 0; x0 < cntE; x0 += 32) {
          const int e = e_nxt;
          const int xn = x0 + 32 + lane;
          e_nxt = xn < cntE ? (int)ids[bef + xn] : 0;
          const bool keep = (x0 + lane < cntE) && s_seg[e >> 8] != s_seg[(e >> 1) & 127];
          const uint32_t bal = __ballot_sync(0xffffffffu, keep);
          if (keep) C_f[kept + __popc(bal & lt)] = (unsigned short)e;
          kept += __popc(bal);
        }
      }
      __syncwarp();
      // phase 2: one dense gather of their weights (instead of one dependent gather per 32 visited entries).
      // entry (row i, col j) = weight(idx[i] -> idx[j]); packed (col << 8) | row orders like col*n+row (VS.h:1922)
      for (int i = lane; i < kept; i += 32) {
        const int e = C_f[i];
        const int a = e >> 8, b = (e >> 1) & 127;
        float w_ab, w_ba;
        fetch(a, b, w_ab, w_ba);
        float w; int f;
        if (e & 1) { w = w_ba; f = (a << 8) | b; } else { w = w_ab; f = (b << 8) | a; }
        if (kept <= 32) { rw = w; rf = f; } else { C_w[i] = w; C_f[i] = (unsigned short)f; }
      }
      __syncwarp();
      if (kept > 0) {
        const bool below_pending = true;
        (void)below_pending;
        bool below;
        if (kept <= 32) {
          // one entry per lane: bitonic sort across the warp with shuffles, (w desc, packed index asc)
#pragma unroll
          for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
              const float wo = __shfl_xor_sync(0xffffffffu, rw, j);
              const int fo = __shfl_xor_sync(0xffffffffu, rf, j);
              const bool other_first = (wo > rw) || (wo == rw && fo < rf);
              const bool up = (lane & kk) == 0, lower = (lane & j) == 0;
              if ((up == lower) ? other_first : !other_first) { rw = wo; rf = fo; }
            }
          }
          const float w0 = __shfl_sync(0xffffffffu, rw, 0);
          below = !(w0 > s_thr[s_seg[0]]);   // S0 rule (see k_local_graph2): the emitted segment cannot merge any more
          stop = below || nseg <= 1;
          if (!stop) merge_batch(rw, rf, lane < kept);
        } else {
          int P = 64;
          while (P < kept) P <<= 1;
          for (int i = kept + lane; i < P; i += 32) { C_w[i] = -1.0f; C_f[i] = 0xffff; }
          __syncwarp();
          for (int kk = 2; kk <= P; kk <<= 1) {       // bitonic sort in shared memory
            for (int j = kk >> 1; j > 0; j >>= 1) {
              for (int i = lane; i < P; i += 32) {
                const int x = i ^ j;
                if (x > i) {
                  const float wi = C_w[i], wx = C_w[x];
                  const unsigned short fi = C_f[i], fx = C_f[x];
                  const bool x_before_i = (wx > wi) || (wx == wi && fx < fi);
                  const bool up = (i & kk) == 0;
                  if (x_before_i == up) { C_w[i] = wx; C_w[x] = wi; C_f[i] = fx; C_f[x] = fi; }
                }
              }
              __syncwarp();
            }
          }
          below = !(C_w[0] > s_thr[s_seg[0]]);
          stop = below || nseg <= 1;
          for (int bs = 0; bs < kept && !stop; bs += 32) {
            const int e = bs + lane;
            const bool valid = e < kept;
            merge_batch(valid ? C_w[e] : 0.f, valid ? (int)C_f[e] : 0, valid);
          }
        }
        if (nseg <= 1 || below) done = true;
      }
      ratio = fminf(1.0f, fmaxf(1.25f * (float)(kept + 2) / (float)(tot + 2), 1.0f / 256.0f));
      c0 = c1;
      if (dbg && lane == 0) { atomicAdd(&dbg[0], 1ull); atomicAdd(&dbg[1], (unsigned long long)tot); atomicAdd(&dbg[2], (unsigned long long)kept); }
    }
    if (dbg && lane == 0) { atomicAdd(&dbg[3], 1ull); atomicAdd(&dbg[4], (unsigned long long)m); atomicAdd(&dbg[5], (unsigned long long)nseg); atomicAdd(&dbg[6], (unsigned long long)nv); }
  }
  if (to_fallback) {
    if (lane == 0) fallback[atomicAdd(fallback_count, 1u)] = u;
    return;
  }
  // --- emit the segment that contains local vertex 0 (the unit itself) ---
  const int s0 = s_seg[0];
  int cnt = 0;
  for (int b = 0; b < nv; b += 32) {
    const int v = b + lane;
    const bool in = v < nv && s_seg[v] == s0;
    const uint32_t bal = __ballot_sync(0xffffffffu, in);
    if (in) conn_idx[off + cnt + __popc(bal & lt)] = s_gid[v];
    cnt += __popc(bal);
  }
  if (lane == 0) conn_cnt[u] = (uint32_t)cnt;
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

// units whose list is {self} after the mutual filter and that have enough neighbours (VS.h:2199-2201)
__global__ void __launch_bounds__(256) k_collect_singles(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ cnt1, int64_t nu,
                                                       int adjacency_min, uint32_t* __restrict__ list, uint32_t* __restrict__ counts /* [0]=eligible [1]=all singles */) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu || cnt1[i] != 1u) return;
  atomicAdd(&counts[1], 1u);
  const int n = (int)(adj_off[i + 1] - adj_off[i]);
  if (!(n + 1 > adjacency_min)) return;
  list[atomicAdd(&counts[0], 1u)] = (uint32_t)i;
}
// one warp per single unit: lanes evaluate the candidates; the winner is the LAST candidate with the
// largest weight (`>=` in VS.h:2281), candidates in the order slot 0 (= the COUNT), then the neighbours
__global__ void __launch_bounds__(128) k_closest_round_warp(const uint32_t* __restrict__ list, uint32_t nlist, const uint32_t* __restrict__ adj_off,
                                                          const int32_t* __restrict__ adj_idx, const uint32_t* __restrict__ cnt1,
                                                          const float* __restrict__ rec, int64_t nu, PairParams pp, int32_t* attach,
                                                          uint32_t* __restrict__ changed) {
  const int lane = threadIdx.x & 31;
  const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (li >= nlist) return;
  const int64_t i = list[li];
  const uint32_t off = adj_off[i];
  const int n = (int)(adj_off[i + 1] - off);
  float ri[REC_FLOATS], rc[REC_FLOATS];
  for (int q = 0; q < REC_FLOATS; q++) ri[q] = rec[i * REC_FLOATS + q];
  float best = 0.f;
  int bj = -1, bi = -1;
  for (int j = lane; j <= n; j += 32) {
    const int64_t c = (j == 0) ? (int64_t)n : (int64_t)adj_idx[off + j - 1];
    if (c < 0 || c >= nu) continue;
    const uint32_t cc = cnt1[c];
    if (!(cc > 1u || (cc == 1u && c < i && ((volatile int32_t*)attach)[c] >= 0))) continue;
    for (int t = 0; t < REC_FLOATS; t++) rc[t] = rec[c * REC_FLOATS + t];
    float w_ab, w_ba;
    pair_weights(ri, rc, pp, w_ab, w_ba);
    if (w_ab >= best) { best = w_ab; bj = j; bi = (int)c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ow = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oj >= 0 && (bj < 0 || ow > best || (ow == best && oj > bj))) { best = ow; bj = oj; bi = oi; }
  }
  if (lane == 0 && bi != attach[i]) { attach[i] = bi; *changed = 1u; }
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
                                               int* __restrict__ parent) {
  const int lane = threadIdx.x & 31;
  int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  const uint32_t off = adj_off[u];
  const int c = (int)cnt1[u];
  int p = (int)u;
  for (int e = lane; e < c; e += 32) p = min(p, idx1[off + e]);
  p = __reduce_min_sync(0xffffffffu, p);
  if (lane == 0) {
    const int a = attach[u];
    if (a >= 0) p = min(p, a);
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
                                               int* parent) {
  const int lane = threadIdx.x & 31;
  int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  const uint32_t off = adj_off[u];
  const int c = (int)cnt1[u];
  for (int e = lane; e < c; e += 32) {
    int j = idx1[off + e];
    if (j > (int)u) uf_union(parent, (int)u, j);
  }
  if (lane == 0) { int a = attach[u]; if (a >= 0) uf_union(parent, (int)u, a); }
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

}  // namespace vgs

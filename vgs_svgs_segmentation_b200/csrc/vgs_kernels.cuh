// vgs_kernels.cuh — the stage kernels of the VGS/SVGS hot path for sm_100a.
// One kernel per row of SURVEY.md §2's kernel table; algorithmic bytes per unit in DESIGN.md.
#pragma once
#include "vgs_prims.cuh"

namespace vgs {

constexpr int MAX_EPOCHS = 40;
constexpr int N_CLASSES = 7;
// class c holds neighbourhoods with n(n-1) <= CLASS_M[c] directed off-diagonal weights
__constant__ int c_class_n[N_CLASSES] = {16, 32, 45, 64, 91, 128, 181};
constexpr int CLASS_N_HOST[N_CLASSES] = {16, 32, 45, 64, 91, 128, 181};
constexpr int CLASS_M_HOST[N_CLASSES] = {256, 1024, 2048, 4096, 8192, 16384, 32768};
constexpr int CLASS_T_HOST[N_CLASSES] = {64, 128, 128, 256, 256, 512, 512};
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
                                                float* __restrict__ rec, unsigned long long* __restrict__ n_used) {
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
//      fill == 0: write the neighbour count; fill == 1: write the ordered list at adj_off[v]. ----
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
  if (!fill) { if (lane == 0) adj_cnt[v] = (uint32_t)count; return; }
  __syncwarp();
  if (count > cap) count = cap;
  const uint32_t off = adj_off[v];
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

// ---- bin used units by neighbourhood size class ----
__global__ void __launch_bounds__(256) k_bin_classes(const uint32_t* __restrict__ adj_off, const float* __restrict__ rec, int64_t nu,
                                                   uint32_t* __restrict__ class_count, uint32_t* __restrict__ class_list,
                                                   unsigned long long* __restrict__ stats /* [0]=sum n^2 [1]=max n [2]=overflow */) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  int fl = f2i(rec[u * REC_FLOATS + REC_FLAGS]);
  if (!(fl & F_USED)) return;
  int n = (int)(adj_off[u + 1] - adj_off[u]);
  int c = 0;
  while (c < N_CLASSES && n > c_class_n[c]) c++;
  atomicMax(&stats[1], (unsigned long long)n);
  if (c >= N_CLASSES) { atomicAdd(&stats[2], 1ull); return; }
  atomicAdd(&stats[0], (unsigned long long)n * (unsigned long long)(n - 1));
  uint32_t pos = atomicAdd(&class_count[c], 1u);
  class_list[(int64_t)c * nu + pos] = (uint32_t)u;
}

// ---- stage 4+5a: local affinity graph + Felzenszwalb-style cut of ONE unit per CTA
//      (buildAdjacencyGraph VS.h:1796-1910 + cutGraphSegmentation VS.h:1913-2029).
//      smem: neighbour records, directed weights (float) + flat index (u16), segment state.
//      Weights w <= 1-2k+k/n can never merge (DESIGN.md §cut bound) and are dropped before the
//      sort; order = (w desc, flat index asc); merge scans 32 sorted entries per warp step. ----
struct GraphParams {
  PairParams pp;
  float cut;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_local_graph(const uint32_t* __restrict__ list, uint32_t nlist,
                                                       const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                       const float* __restrict__ rec, GraphParams gp, int ncap, int mcap,
                                                       uint32_t* __restrict__ conn_cnt, int32_t* __restrict__ conn_idx) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float* s_rec = reinterpret_cast<float*>(smraw);                       // ncap * REC_PAD
  float* s_w = s_rec + (size_t)ncap * REC_PAD;                          // mcap
  float* s_int = s_w + mcap;                                            // ncap
  int* s_gid = reinterpret_cast<int*>(s_int + ncap);                    // ncap
  unsigned short* s_f = reinterpret_cast<unsigned short*>(s_gid + ncap);  // mcap
  unsigned short* s_seg = s_f + mcap;                                   // ncap
  unsigned short* s_size = s_seg + ncap;                                // ncap
  __shared__ int s_m;
  __shared__ int s_nseg;

  const int tid = threadIdx.x;
  if (blockIdx.x >= nlist) return;
  const uint32_t u = list[blockIdx.x];
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  if (tid == 0) { s_m = 0; s_nseg = n; }
  for (int i = tid; i < n; i += THREADS) {
    s_gid[i] = adj_idx[off + i];
    s_seg[i] = (unsigned short)i; s_size[i] = 1; s_int[i] = 1.0f;
  }
  __syncthreads();
  for (int t = tid; t < n * 4; t += THREADS) {  // 4 x float4 per record
    int i = t >> 2, q = t & 3;
    float4 val = __ldg(reinterpret_cast<const float4*>(rec + (int64_t)s_gid[i] * REC_FLOATS) + q);
    float* d = s_rec + i * REC_PAD + q * 4;
    d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
  }
  __syncthreads();
  // --- directed weights of all unordered pairs ---
  const float k = gp.cut;
  const float lb = (float)(1.0 - 2.0 * (double)k + (double)k / (double)n - 4e-7 * (double)(n + 8));
  const int npairs = n * (n - 1) / 2;
  for (int p = tid; p < npairs; p += THREADS) {
    int r = p / (n - 1), c = p - r * (n - 1);
    int a, b;
    if (c < n - 1 - r) { a = r; b = r + 1 + c; }
    else { a = n - 1 - r; b = a + 1 + (c - (n - 1 - r)); }
    float w_ab, w_ba;
    pair_weights(s_rec + a * REC_PAD, s_rec + b * REC_PAD, gp.pp, w_ab, w_ba);
    // matrix entry (row i, col j) = weight(v1=idx[i], v2=idx[j]); flat index = col*n + row
    if (w_ab > lb) { int s = atomicAdd(&s_m, 1); s_w[s] = w_ab; s_f[s] = (unsigned short)(b * n + a); }
    if (w_ba > lb) { int s = atomicAdd(&s_m, 1); s_w[s] = w_ba; s_f[s] = (unsigned short)(a * n + b); }
  }
  __syncthreads();
  const int m = s_m;
  int mpad = 32;
  while (mpad < m) mpad <<= 1;
  for (int i = m + tid; i < mpad; i += THREADS) { s_w[i] = -1.0f; s_f[i] = 0xffff; }
  __syncthreads();
  // --- bitonic sort: (w desc, f asc) ---
  for (int kk = 2; kk <= mpad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < mpad; i += THREADS) {
        int x = i ^ j;
        if (x > i) {
          float wi = s_w[i], wx = s_w[x];
          unsigned short fi = s_f[i], fx = s_f[x];
          bool x_before_i = (wx > wi) || (wx == wi && fx < fi);
          bool up = (i & kk) == 0;
          if (x_before_i == up) { s_w[i] = wx; s_w[x] = wi; s_f[i] = fx; s_f[x] = fi; }
        }
      }
      __syncthreads();
    }
  }
  // --- merge (warp 0): 32 sorted entries per step, first mergeable entry merges, rest re-evaluated ---
  if (tid < 32) {
    const int lane = tid;
    int nseg = n;
    for (int base = 0; base < m && nseg > 1; base += 32) {
      int e = base + lane;
      bool valid = e < m;
      float w = valid ? s_w[e] : 0.f;
      int f = valid ? (int)s_f[e] : 0;
      int v1 = f / n, v2 = f - v1 * n;
      uint32_t todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        bool pred = false;
        int sa = 0, sb = 0;
        float thr = 0.f;
        bool a_wins = true;
        if ((todo >> lane) & 1u) {
          sa = s_seg[v1]; sb = s_seg[v2];
          if (sa != sb) {
            float m1 = s_int[sa] - k / (float)(int)s_size[sa];
            float m2 = s_int[sb] - k / (float)(int)s_size[sb];
            a_wins = (m1 >= m2);
            thr = a_wins ? m1 : m2;
            pred = w > thr;
          }
        }
        uint32_t bal = __ballot_sync(0xffffffffu, pred);
        if (!bal) break;
        int L = __ffs(bal) - 1;
        int keep = __shfl_sync(0xffffffffu, a_wins ? sa : sb, L);
        int drop = __shfl_sync(0xffffffffu, a_wins ? sb : sa, L);
        float wl = __shfl_sync(0xffffffffu, w, L);
        for (int v = lane; v < n; v += 32) if (s_seg[v] == drop) s_seg[v] = (unsigned short)keep;
        if (lane == 0) { s_int[keep] = wl; s_size[keep] = (unsigned short)(s_size[keep] + s_size[drop]); s_size[drop] = 0; }
        nseg--;
        __syncwarp();
        todo &= ~((2u << L) - 1u);  // entries up to and including L are settled
        if (nseg <= 1) break;
      }
    }
    // --- emit the segment that contains local vertex 0 (the unit itself) ---
    const int s0 = s_seg[0];
    int cnt = 0;
    for (int b = 0; b < n; b += 32) {
      int v = b + lane;
      bool in = v < n && s_seg[v] == s0;
      uint32_t bal = __ballot_sync(0xffffffffu, in);
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
//      fixed point reproduces the sequential result.  One thread per single unit. ----
__global__ void __launch_bounds__(128) k_closest_round(const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                     const uint32_t* __restrict__ cnt1, const float* __restrict__ rec, int64_t nu,
                                                     int adjacency_min, PairParams pp, int32_t* attach, uint32_t* __restrict__ changed,
                                                     unsigned long long* __restrict__ n_singles) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu) return;
  if (cnt1[i] != 1u) return;
  if (n_singles) atomicAdd(n_singles, 1ull);
  const uint32_t off = adj_off[i];
  const int n = (int)(adj_off[i + 1] - off);
  if (!(n + 1 > adjacency_min)) return;   // voxels_adjacency_idx_[i].size() = count + 1 (VS.h:2201)
  float ri[REC_FLOATS], rc[REC_FLOATS];
  for (int q = 0; q < REC_FLOATS; q++) ri[q] = rec[i * REC_FLOATS + q];
  float best = 0.f;
  int bi = -1;
  for (int j = 0; j <= n; j++) {
    int64_t c = (j == 0) ? (int64_t)n : (int64_t)adj_idx[off + j - 1];  // slot 0 is the COUNT (VS.h:2243)
    if (c < 0 || c >= nu) continue;
    uint32_t cc = cnt1[c];
    bool q = cc > 1u || (cc == 1u && c < i && ((volatile int32_t*)attach)[c] >= 0);
    if (!q) continue;
    for (int t = 0; t < REC_FLOATS; t++) rc[t] = rec[c * REC_FLOATS + t];
    float w_ab, w_ba;
    pair_weights(ri, rc, pp, w_ab, w_ba);
    if (w_ab >= best) { best = w_ab; bi = (int)c; }
  }
  if (bi != attach[i]) { attach[i] = bi; *changed = 1u; }
}

// ---- stage 5d: connected components by lock-free union-find, root = smallest unit id ----
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = ((volatile int*)parent)[x];
  while (p != x) { x = p; p = ((volatile int*)parent)[x]; }
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
  if (u >= nu) return;
  int r = root[u];
  atomicAdd(&csize[r], 1u);
  atomicMin(&cminpt[r], perm[ustart[u]]);  // the unit's points ascend by index
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

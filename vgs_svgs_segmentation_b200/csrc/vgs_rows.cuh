// vgs_rows.cuh — VGS lattice kernels built around an occupancy grid of the occupied key range ("BitGrid") and
// per-voxel rows of pair weights ordered by weight.
//
//   stage 3  k_adj_count / k_adj_fill : radius adjacency (findAllVoxelAdjacency VS.h:223-265) written straight to the CSR,
//                                       lists ordered by (dist2, id) with a rank sort inside each lattice distance class
//   stage 4  k_rows_fill / k_rows_sort: every unordered pair of used voxels that can meet in a local graph is evaluated ONCE
//                                       (buildAdjacencyGraph VS.h:1796-1910 re-evaluates it in every graph) and both directed
//                                       weights are filed in the row of their source voxel; rows are ordered by weight cell
//   stage 5a k_local_graph_rows       : cutGraphSegmentation (VS.h:1913-2029) of one voxel per warp, consuming only the heavy
//                                       prefix of the rows of its neighbours (S0 rule: the cut of the voxel itself is final
//                                       once the next weight is <= Int(S0) - k/|S0|)
//   stage 5b k_mutual_mask            : crossValidation (VS.h:2111-2179) on lattice-offset bit masks
#pragma once
#include <type_traits>
#include "vgs_kernels.cuh"

namespace vgs {

// ---- occupancy grid over the occupied key range + a margin, so stencil probes need no bounds checks ----
struct BitGrid {
  int x0, y0, z0;          // key of cell (0,0,0) = smallest occupied key - margin
  uint32_t ny, nz;         // extents incl. margins; bit = ((x-x0)*ny + (y-y0))*nz + (z-z0)
};
__device__ __forceinline__ uint64_t bg_bit(const BitGrid& g, int x, int y, int z) {
  return ((uint64_t)(uint32_t)(x - g.x0) * g.ny + (uint32_t)(y - g.y0)) * g.nz + (uint32_t)(z - g.z0);
}
// bits [b0, b0 + len) of the grid, len <= 32 (bit j = cell z0 + j of the run)
__device__ __forceinline__ uint32_t bg_run(const uint32_t* __restrict__ bm, uint64_t b0, int len) {
  const uint64_t two = (uint64_t)__ldg(bm + (b0 >> 5)) | ((uint64_t)__ldg(bm + (b0 >> 5) + 1) << 32);
  return (uint32_t)(two >> (b0 & 31)) & (len >= 32 ? 0xffffffffu : ((1u << len) - 1u));
}
__global__ void __launch_bounds__(256) k_bitgrid_set(const uint32_t* __restrict__ key3, const uint8_t* __restrict__ uflags, int64_t nv,
                                                   BitGrid g, uint32_t* __restrict__ bm_all, uint32_t* __restrict__ bm_used) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const uint64_t b = bg_bit(g, (int)key3[3 * v], (int)key3[3 * v + 1], (int)key3[3 * v + 2]);
  atomicOr(&bm_all[b >> 5], 1u << (b & 31));
  if (uflags[v] & F_USED) atomicOr(&bm_used[b >> 5], 1u << (b & 31));
}

// voxel key + smallest / largest occupied key per axis (the extent of the occupancy grids)
__global__ void __launch_bounds__(256) k_voxel_keys(const uint64_t* __restrict__ ukey, int64_t nu, int depth, int descending,
                                                  uint32_t* __restrict__ key3, uint32_t* __restrict__ kminmax /* [6] */) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t kx = 0xffffffffu, ky = 0xffffffffu, kz = 0xffffffffu, Kx = 0, Ky = 0, Kz = 0;
  if (u < nu && !(ukey[u] >> (3 * depth))) {     // the segment of the non-finite points (sentinel key) has no voxel key
    uint64_t m = ukey[u];
    const uint64_t mask = (1ull << (3 * depth)) - 1ull;
    if (descending) m = ~m & mask;
    morton_decode(m, kx, ky, kz);
    key3[3 * u] = kx; key3[3 * u + 1] = ky; key3[3 * u + 2] = kz;
    Kx = kx; Ky = ky; Kz = kz;
  }
  kx = __reduce_min_sync(0xffffffffu, kx); ky = __reduce_min_sync(0xffffffffu, ky); kz = __reduce_min_sync(0xffffffffu, kz);
  Kx = __reduce_max_sync(0xffffffffu, Kx); Ky = __reduce_max_sync(0xffffffffu, Ky); Kz = __reduce_max_sync(0xffffffffu, Kz);
  // one set of global atomics per CTA (six hot addresses: per-warp atomics serialise in L2)
  __shared__ uint32_t s_mm[6];
  if (threadIdx.x < 6) s_mm[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_mm[0], kx); atomicMin(&s_mm[1], ky); atomicMin(&s_mm[2], kz);
    atomicMax(&s_mm[3], Kx); atomicMax(&s_mm[4], Ky); atomicMax(&s_mm[5], Kz);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(&kminmax[threadIdx.x], s_mm[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(&kminmax[threadIdx.x], s_mm[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_voxel_centers(const uint32_t* __restrict__ key3, int64_t nu, float res_f, float mnx, float mny,
                                                     float mnz, float* __restrict__ center) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  center[3 * u] = (float)(((double)key3[3 * u] + 0.5f) * res_f + mnx);
  center[3 * u + 1] = (float)(((double)key3[3 * u + 1] + 0.5f) * res_f + mny);
  center[3 * u + 2] = (float)(((double)key3[3 * u + 2] + 0.5f) * res_f + mnz);
}

// geometry of the lattice searches, filled by the host
struct LatticeGeom {
  float res_f, mnx, mny, mnz;   // float-narrowed members the centres are computed from (VS.h:2102-2109)
  float r2;                     // (float)(graph_size^2): FLANN's strict test dist2 < r2
  int rho;                      // reach of the radius stencil (cells)
  int r2c;                      // reach of the pair stencil = 2 * rho
};
// centre coordinate of key k exactly as k_voxel_centers stores it
__device__ __forceinline__ float centre_of(uint32_t k, float res_f, float mn) { return (float)(((double)k + 0.5f) * res_f + mn); }
// FLANN L2_Simple: sequential float accumulation of squared differences
__device__ __forceinline__ float flann_d2(float qx, float qy, float qz, float cx, float cy, float cz) {
  const float dx = qx - cx, dy = qy - cy, dz = qz - cz;
  float d2 = 0.f; d2 += dx * dx; d2 += dy * dy; d2 += dz * dz;
  return d2;
}

// 5-bit packed lattice offsets relative to a centre voxel: (o + rho) per axis, x | y << 5 | z << 10
__host__ __device__ __forceinline__ uint32_t pack5(int x, int y, int z) { return (uint32_t)x | ((uint32_t)y << 5) | ((uint32_t)z << 10); }
// 6-bit fields (bit 5 of each field is a guard bit for the field-parallel range test of the consumer)
__host__ __device__ __forceinline__ uint32_t pack6(int x, int y, int z) { return (uint32_t)x | ((uint32_t)y << 6) | ((uint32_t)z << 12); }

// ---- stage 3a: neighbour count of every voxel (all-voxel grid, radius stencil, float test on the centres computed
//      from the keys) and, for used voxels, the length of its weight row (used-voxel grid, pair stencil: a voxel files
//      the pairs with its partners in the lexicographically positive half; the other half is filed by the partners).
//      One warp per voxel, lanes over stencil columns (dx, dy, mask of dz). ----
struct CountStats { unsigned long long sum_nn, max_n, n_long, n_used; };
__global__ void __launch_bounds__(256) k_adj_count(const uint32_t* __restrict__ key3, const uint8_t* __restrict__ uflags, int64_t nv,
                                                 LatticeGeom lg, BitGrid g, const uint32_t* __restrict__ bm_all,
                                                 const uint32_t* __restrict__ bm_used, const int4* __restrict__ adj_cols, int n_adj_cols,
                                                 const int4* __restrict__ pc_cols, int n_pc_cols, uint32_t* __restrict__ adj_cnt,
                                                 uint32_t* __restrict__ row_len, int long_len,
                                                 uint32_t* __restrict__ long_rows, CountStats* __restrict__ stats) {
  __shared__ unsigned long long s_sum[8];
  __shared__ unsigned s_max[8], s_used[8];
  __shared__ float s_cax[8][96];      // per warp: centres of the stencil cells per axis (rho <= 15)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t v = (int64_t)blockIdx.x * 8 + w;
  unsigned long long my_sum = 0; unsigned my_max = 0, my_used = 0;
  if (v < nv) {
    const int kx = (int)key3[3 * v], ky = (int)key3[3 * v + 1], kz = (int)key3[3 * v + 2];
    // centre coordinates of the 2 rho + 1 cells per axis, once per voxel (the float test needs FLANN's float centres)
    const int S = 2 * lg.rho + 1;
    float* cax = s_cax[w];
    for (int t = lane; t < 3 * S; t += 32) {
      const int a = t / S, d = t - a * S;
      cax[a * 32 + d] = a == 0 ? centre_of(kx - lg.rho + d, lg.res_f, lg.mnx) : (a == 1 ? centre_of(ky - lg.rho + d, lg.res_f, lg.mny) : centre_of(kz - lg.rho + d, lg.res_f, lg.mnz));
    }
    __syncwarp();
    const float qx = cax[lg.rho], qy = cax[32 + lg.rho], qz = cax[64 + lg.rho];
    int cnt = 0;
    for (int ci = lane; ci < n_adj_cols; ci += 32) {
      const int4 o = adj_cols[ci];
      const int x = kx + o.x, y = ky + o.y;
      uint32_t hits = bg_run(bm_all, bg_bit(g, x, y, kz - lg.rho), S) & (uint32_t)o.z;
      const float cx = cax[o.x + lg.rho], cy = cax[32 + o.y + lg.rho];
      while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        cnt += flann_d2(qx, qy, qz, cx, cy, cax[64 + j]) < lg.r2 ? 1 : 0;
      }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    int npos = 0;
    const bool used = (uflags[v] & F_USED) != 0;
    if (used) {
      const int span = 2 * lg.r2c + 1;
      for (int ci = lane; ci < n_pc_cols; ci += 32) {
        const int4 o = pc_cols[ci];
        npos += __popc(bg_run(bm_used, bg_bit(g, kx + o.x, ky + o.y, kz - lg.r2c), span) & (uint32_t)o.z);
      }
      npos = __reduce_add_sync(0xffffffffu, npos);
    }
    if (lane == 0) {
      adj_cnt[v] = (uint32_t)cnt;
      const uint32_t len = (uint32_t)npos;
      row_len[v] = len;
      if ((int)len > long_len) long_rows[atomicAdd(&stats->n_long, 1ull)] = (uint32_t)v;
      my_sum = used ? (unsigned long long)cnt * (unsigned long long)(cnt - 1) : 0ull;
      my_max = (unsigned)cnt; my_used = used ? 1u : 0u;
    }
  }
  if (lane == 0) { s_sum[w] = my_sum; s_max[w] = my_max; s_used[w] = my_used; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0; unsigned m = 0, us = 0;
    for (int i = 0; i < 8; i++) { s += s_sum[i]; m = max(m, s_max[i]); us += s_used[i]; }
    if (s) atomicAdd(&stats->sum_nn, s);
    if (us) atomicAdd(&stats->n_used, (unsigned long long)us);
    atomicMax(&stats->max_n, (unsigned long long)m);
  }
}

// dense voxel-id grid over the cells of the BitGrid (4 B per cell, -1 = empty): a neighbour's id is ONE load (a z-run of
// the stencil shares a sector) instead of a Morton encode + hash probes.  Used when the grid fits the budget.
__global__ void __launch_bounds__(256) k_idgrid_set(const uint32_t* __restrict__ key3, int64_t nv, BitGrid g, int32_t* __restrict__ idg) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  idg[bg_bit(g, (int)key3[3 * v], (int)key3[3 * v + 1], (int)key3[3 * v + 2])] = (int32_t)v;
}
__device__ __forceinline__ int voxel_at(const int32_t* __restrict__ idg, const BitGrid& g, const unsigned long long* __restrict__ tk,
                                        const uint32_t* __restrict__ tv, uint64_t hmask, int x, int y, int z) {
  if (idg) return __ldg(idg + bg_bit(g, x, y, z));
  return hash_lookup(tk, tv, hmask, morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z));
}

// ---- stage 3b: ordered neighbour lists straight into the CSR.  Candidates are keyed by their stencil slot (slots are
//      sorted by integer distance class on the host); inside a class the order is FLANN's (float dist2, id), found by a
//      rank count over the PRESENT members of the class only — the classes are >= res^2 apart, far above the float noise
//      of dist2 (the host passes ONE class = a full rank sort when the cloud is so far from the origin that this is not
//      certain).  adj_code = 5-bit packed offset of the neighbour (pack5), used by the graph and mutual-filter kernels. ----
struct AdjTables {
  const uint32_t* code_lut;       // ((dx+rho)*S + (dy+rho))*S + (dz+rho) -> slot | (dx+rho) << 16 | (dy+rho) << 21 | (dz+rho) << 26 (slot 0xffff: none)
  const uint16_t* slot_code5;     // slot -> pack5(dx+rho, dy+rho, dz+rho)
  const uint16_t* cls_first;      // slot -> first slot of its distance class
  const uint16_t* cls_last;       // slot -> last slot of its distance class
};
constexpr int ADJ_WARPS = 4;
__host__ __device__ inline size_t adj_fill_smem(int nst, int S) {
  const int nst8 = (nst + 7) & ~7, S4 = (S + 3) & ~3;
  // per warp: slot keys + compacted keys (8 B), hit queue / present-slot list (2 B), presence words + counts, axis centres
  return (size_t)ADJ_WARPS * ((size_t)nst8 * 18 + (size_t)(nst8 / 32 + 2) * 8 + (size_t)S4 * 12);
}
__global__ void __launch_bounds__(ADJ_WARPS * 32) k_adj_fill(const uint32_t* __restrict__ key3, int64_t nv, LatticeGeom lg, BitGrid g,
                                                           const uint32_t* __restrict__ bm_all, const int4* __restrict__ adj_cols,
                                                           int n_adj_cols, AdjTables tb, int nst, const int32_t* __restrict__ idg,
                                                           const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                           uint64_t hmask, const uint32_t* __restrict__ adj_off, int32_t* __restrict__ adj_idx,
                                                           uint16_t* __restrict__ adj_code, unsigned* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int rho = lg.rho, S = 2 * rho + 1, S4 = (S + 3) & ~3;
  const int nst8 = (nst + 7) & ~7, nblk = nst8 / 32 + 2;
  unsigned char* base = smraw;
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(base) + (size_t)w * nst8;   // per slot: (d2 bits << 32) | id, ~0 = absent
  base += (size_t)ADJ_WARPS * nst8 * 8;
  unsigned long long* pkey = reinterpret_cast<unsigned long long*>(base) + (size_t)w * nst8;   // keys of the present slots, slot order
  base += (size_t)ADJ_WARPS * nst8 * 8;
  unsigned short* q = reinterpret_cast<unsigned short*>(base) + (size_t)w * nst8;             // hit queue, then present-slot list
  base += (size_t)ADJ_WARPS * nst8 * 2;
  unsigned* sP = reinterpret_cast<unsigned*>(base) + (size_t)w * nblk * 2;                     // presence words
  unsigned* sC = sP + nblk;                                                                   // present slots before the word
  base += (size_t)ADJ_WARPS * nblk * 8;
  float* cax = reinterpret_cast<float*>(base) + (size_t)w * 3 * S4;                            // centre coordinates of the S cells per axis
  const int64_t v = (int64_t)blockIdx.x * ADJ_WARPS + w;
  if (v >= nv) return;
  const int kx = (int)key3[3 * v], ky = (int)key3[3 * v + 1], kz = (int)key3[3 * v + 2];
  for (int t = lane; t < 3 * S; t += 32) {
    const int a = t / S, d = t - a * S;
    cax[a * S4 + d] = a == 0 ? centre_of(kx - rho + d, lg.res_f, lg.mnx) : (a == 1 ? centre_of(ky - rho + d, lg.res_f, lg.mny) : centre_of(kz - rho + d, lg.res_f, lg.mnz));
  }
  for (int s = lane; s < nst; s += 32) skey[s] = ~0ull;
  int nq = 0;
  for (int b0 = 0; b0 < n_adj_cols; b0 += 32) {
    const int ci = b0 + lane;
    uint32_t hits = 0;
    int cbase = 0;
    if (ci < n_adj_cols) {
      const int4 o = adj_cols[ci];
      hits = bg_run(bm_all, bg_bit(g, kx + o.x, ky + o.y, kz - rho), S) & (uint32_t)o.z;
      cbase = ((o.x + rho) * S + (o.y + rho)) * S;
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
  const float qx = cax[rho], qy = cax[S4 + rho], qz = cax[2 * S4 + rho];
  for (int e0 = 0; e0 < nq; e0 += 32) {
    const int e = e0 + lane;
    int id = -1, slot = 0;
    float d2 = 0.f;
    if (e < nq) {
      const uint32_t lut = __ldg(tb.code_lut + q[e]);
      const int ox = (lut >> 16) & 31, oy = (lut >> 21) & 31, oz = (lut >> 26) & 31;
      slot = (int)(lut & 0xffffu);
      d2 = flann_d2(qx, qy, qz, cax[ox], cax[S4 + oy], cax[2 * S4 + oz]);
      if (d2 < lg.r2) {
        id = voxel_at(idg, g, tk, tv, hmask, kx + ox - rho, ky + oy - rho, kz + oz - rho);
        if (id < 0) atomicOr(err, 1u);               // grid and voxel table disagree: cannot happen
      }
    }
    __syncwarp();     // the probe loops end at different iterations
    if (id >= 0) skey[slot] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)id;
  }
  __syncwarp();
  // presence words + running counts; present slots and their keys compacted in slot order
  int np = 0;
  for (int s0 = 0; s0 < nst; s0 += 32) {
    const int s = s0 + lane;
    const unsigned long long me = s < nst ? skey[s] : ~0ull;
    const uint32_t bal = __ballot_sync(0xffffffffu, me != ~0ull);
    if (lane == 0) { sP[s0 >> 5] = bal; sC[s0 >> 5] = (unsigned)np; }
    if (me != ~0ull) {
      const int p = np + __popc(bal & ((1u << lane) - 1u));
      pkey[p] = me; q[p] = (unsigned short)s;
    }
    np += __popc(bal);
  }
  __syncwarp();
  const uint32_t off = adj_off[v];
  const uint32_t total = adj_off[v + 1] - off;
  if ((uint32_t)np != total) { if (lane == 0) atomicOr(err, 2u); return; }   // count and fill disagree: cannot happen
  for (int e = lane; e < np; e += 32) {
    const unsigned long long me = pkey[e];
    const int s = q[e];
    const int f = tb.cls_first[s], l = tb.cls_last[s] + 1;
    // present members of the class = compacted entries [pb, pe)
    const int pb = (int)sC[f >> 5] + __popc(sP[f >> 5] & ((1u << (f & 31)) - 1u));
    const int pe = l >= nst ? np : (int)sC[l >> 5] + __popc(sP[l >> 5] & ((1u << (l & 31)) - 1u));
    int rank = 0;
    for (int t = pb; t < pe; t++) rank += pkey[t] < me ? 1 : 0;
    adj_idx[off + pb + rank] = (int32_t)(unsigned)(me & 0xffffffffull);
    adj_code[off + pb + rank] = tb.slot_code5[s];
  }
}

// compact list of the used voxels in id order (flag -> exclusive scan -> write): the row and local-graph kernels launch one
// warp per USED voxel, so no CTA slot is spent on the voxels with too few points
__global__ void __launch_bounds__(256) k_used_flags(const uint8_t* __restrict__ uflags, int64_t nv, uint32_t* __restrict__ flag) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) flag[v] = (uflags[v] & F_USED) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_used_write(const uint8_t* __restrict__ uflags, const uint32_t* __restrict__ pos, int64_t nv,
                                                  uint32_t* __restrict__ list) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv && (uflags[v] & F_USED)) list[pos[v]] = (uint32_t)v;
}

// sort key of a used voxel for the local-graph launch order: big neighbourhoods first (longest job first keeps the tail of
// the launch short); equal sizes stay in id order (stable pass), so neighbours still share their rows in L2
__global__ void __launch_bounds__(256) k_order_keys(const uint32_t* __restrict__ used_list, uint32_t n_used, const uint32_t* __restrict__ adj_off,
                                                  int shift, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_used) return;
  const uint32_t u = used_list[i];
  const uint32_t n = min(adj_off[u + 1] - adj_off[u], 255u);
  keys[i] = (255u - n) >> shift;
  vals[i] = u;
}

// ---- stage 4: weight rows.  One 16-byte entry per UNORDERED pair {a, b} of used voxels that can meet in a local graph,
//      filed in the row of the voxel a whose offset to b is lexicographically positive:
//        .x = w(a -> b)   .y = w(b -> a)   .z = cell(max weight) << 18 | pack6(d + r2c), d = key_b - key_a   .w = cell(min weight)
//      cell(w) = floor((1 - w) * 4096) orders a row coarsely (k_rows_sort); the consumer sorts exactly inside the cells
//      it takes.  A row is written by one warp only: no atomics, no scattered stores. ----
constexpr int ROW_CELLS = 4096;
#ifndef VGS_ROWS_CAP
#define VGS_ROWS_CAP 192     // measured (10 M-point site): 256 -> 3.05 ms, 192 -> 2.85 ms, 128 -> 2.72 ms + 0.2 ms of long rows
#endif
#ifndef VGS_RF_MINB
#define VGS_RF_MINB 8
#endif
constexpr int ROWS_SHORT_CAP = VGS_ROWS_CAP;   // rows up to this length are assembled and ordered in shared memory by k_rows_fill; longer ones by k_rows_sort
__host__ __device__ __forceinline__ uint32_t row_cell(float w) {
  // monotone non-increasing in w; NaN and w <= 0 fall into the last cell
  if (!(w > 0.f)) return ROW_CELLS - 1;
  const int q = (int)((1.0f - w) * (float)ROW_CELLS);
  return q < 0 ? 0u : (q > ROW_CELLS - 1 ? (uint32_t)(ROW_CELLS - 1) : (uint32_t)q);
}
// every weight of a cell >= c is <= this bound (1 - w is exact for w >= 0.5, else off by <= 2^-25; the slack covers it)
__device__ __forceinline__ float row_cell_upper(int c) { return (1.0f - (float)c * (1.0f / (float)ROW_CELLS)) + 1.2e-7f; }

// One warp per used voxel.  Rows of up to ROWS_SHORT_CAP entries are assembled in shared memory, ordered by weight cell
// there (LSD radix sort of (cell << 8 | position) words, 2 passes x 6 bits) and written once, in order; longer rows go out
// unsorted and are ordered by k_rows_sort over the long-row list.
constexpr int RF_WARPS = 4;
__host__ __device__ inline size_t rows_fill_smem_warp(int S) {
  const size_t qcap = 32 + 32 * (size_t)S;          // a lane queues at most one z-run (<= S hits) per column step
  return (size_t)ROWS_SHORT_CAP * 16 + 2 * (size_t)ROWS_SHORT_CAP * 4 + 64 * 4 + ((qcap * 2 + 15) & ~(size_t)15) + REC_FLOATS * 4;
}
__global__ void __launch_bounds__(RF_WARPS * 32, VGS_RF_MINB) k_rows_fill(const uint32_t* __restrict__ used_list, uint32_t n_used, const uint32_t* __restrict__ key3,
                                                 const float* __restrict__ rec, LatticeGeom lg,
                                                 BitGrid g, const uint32_t* __restrict__ bm_used, const int4* __restrict__ pc_cols, int n_pc_cols,
                                                 const int32_t* __restrict__ idg, const unsigned long long* __restrict__ tk,
                                                 const uint32_t* __restrict__ tv, uint64_t hmask,
                                                 PairParams pp, const uint32_t* __restrict__ row_off,
                                                 uint4* __restrict__ rows, unsigned* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t li = (int64_t)blockIdx.x * RF_WARPS + w;
  if (li >= (int64_t)n_used) return;
  const int64_t v = used_list[li];
  const int r2 = lg.r2c, S = 2 * r2 + 1;
  unsigned char* mine = smraw + (size_t)w * rows_fill_smem_warp(S);
  uint4* ent = reinterpret_cast<uint4*>(mine);                                  // ROWS_SHORT_CAP entries
  uint32_t* kA = reinterpret_cast<uint32_t*>(ent + ROWS_SHORT_CAP);             // sort words (cell << 8 | position)
  uint32_t* kB = kA + ROWS_SHORT_CAP;
  unsigned* s_cnt = kB + ROWS_SHORT_CAP;                                        // 64 digit counters
  unsigned short* pend = reinterpret_cast<unsigned short*>(s_cnt + 64);         // queue of stencil codes
  float* s_ra = reinterpret_cast<float*>(mine + rows_fill_smem_warp(S) - REC_FLOATS * 4);
  if (lane < REC_FLOATS) s_ra[lane] = rec[v * REC_FLOATS + lane];
  __syncwarp();
  const int kx = (int)key3[3 * v], ky = (int)key3[3 * v + 1], kz = (int)key3[3 * v + 2];
  const uint32_t my_row = row_off[v];
  const int len = (int)(row_off[v + 1] - my_row);
  const bool in_smem = len <= ROWS_SHORT_CAP;
  int npend = 0, done = 0;
  auto process = [&](int first, int cnt) {
    const bool act = lane < cnt;
    int b = -1, dx = 0, dy = 0, dz = 0;
    if (act) {
      const int c = pend[first + lane];              // ((dx+r2)*S + (dy+r2))*S + (dz+r2)
      dz = c % S - r2; dy = (c / S) % S - r2; dx = c / (S * S) - r2;
      b = voxel_at(idg, g, tk, tv, hmask, kx + dx, ky + dy, kz + dz);
      if (b < 0) atomicOr(err, 4u);
    }
    __syncwarp();     // the probe loops end at different iterations: reconverge before the long pair evaluation
    if (act && b >= 0) {
      float rb[REC_FLOATS];
      const float4* src = reinterpret_cast<const float4*>(rec + (int64_t)b * REC_FLOATS);
#pragma unroll
      for (int qd = 0; qd < 4; qd++) { float4 t = __ldg(src + qd); rb[4 * qd] = t.x; rb[4 * qd + 1] = t.y; rb[4 * qd + 2] = t.z; rb[4 * qd + 3] = t.w; }
      float w_ab, w_ba;
      pair_weights(s_ra, rb, pp, w_ab, w_ba);
      const uint32_t ca = row_cell(w_ab), cb = row_cell(w_ba);
      const uint4 e = make_uint4(__float_as_uint(w_ab), __float_as_uint(w_ba), (min(ca, cb) << 18) | pack6(dx + r2, dy + r2, dz + r2), max(ca, cb));
      const int at = done + lane;
      if (in_smem) {
        if (at < ROWS_SHORT_CAP) { ent[at] = e; kA[at] = (min(ca, cb) << 8) | (uint32_t)at; }
      } else rows[my_row + at] = e;
    }
    done += cnt;
  };
  // one call site of the pair evaluation (it is ~1 100 instructions: a second copy does not fit the instruction cache)
  for (int base = 0;; base += 32) {
    const bool flush = base >= n_pc_cols;
    if (!flush) {
      const int ci = base + lane;
      uint32_t hits = 0;
      int cbase = 0;
      if (ci < n_pc_cols) {
        const int4 o = pc_cols[ci];
        hits = bg_run(bm_used, bg_bit(g, kx + o.x, ky + o.y, kz - r2), S) & (uint32_t)o.z;
        cbase = ((o.x + r2) * S + (o.y + r2)) * S;
      }
      const int cnt = __popc(hits);
      const int incl = (int)warp_incl_scan((unsigned)cnt, lane);
      int pos = npend + incl - cnt;
      while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        pend[pos++] = (unsigned short)(cbase + j);
      }
      npend += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    while (npend >= 32 || (flush && npend > 0)) {
      const int take = min(32, npend);
      process(npend - take, take);
      npend -= take;
    }
    __syncwarp();
    if (flush) break;
  }
  if (!in_smem) return;
  if (done != len) { if (lane == 0) atomicOr(err, 8u); return; }     // count and fill disagree: cannot happen
  __syncwarp();
  // ---- order the row by weight cell (stable LSD radix, 2 x 6 bits) and write it once ----
  const uint32_t lt = (1u << lane) - 1u;
  if (len > 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const int shift = 8 + 6 * pass;
      s_cnt[lane] = 0; s_cnt[lane + 32] = 0;
      __syncwarp();
      for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        const unsigned d = i < len ? ((kA[i] >> shift) & 63u) : 0xffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (i < len && (peers & lt) == 0) s_cnt[d] += __popc(peers);
        __syncwarp();
      }
      {   // exclusive prefix sums over the 64 counters: lane owns bins 2*lane, 2*lane + 1
        const unsigned c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
        const unsigned inc = warp_incl_scan(c0 + c1, lane);
        __syncwarp();
        s_cnt[2 * lane] = inc - c0 - c1;
        s_cnt[2 * lane + 1] = inc - c1;
      }
      __syncwarp();
      for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        uint32_t e = 0;
        unsigned d = 0xffffu;
        if (i < len) { e = kA[i]; d = (e >> shift) & 63u; }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        unsigned pos = 0;
        if (i < len) pos = s_cnt[d] + __popc(peers & lt);
        __syncwarp();
        if (i < len && (peers & lt) == 0) s_cnt[d] += __popc(peers);
        if (i < len) kB[pos] = e;
        __syncwarp();
      }
      uint32_t* t = kA; kA = kB; kB = t;
    }
  }
  for (int i = lane; i < len; i += 32) rows[my_row + i] = ent[kA[i] & 255u];
  static_assert(ROWS_SHORT_CAP <= 256, "the sort words carry an 8-bit position");
}

// ---- rows longer than the shared-memory assembly of k_rows_fill: one row per CTA (one warp), ordered by weight cell with
//      the same 2 x 6-bit LSD radix on whole 16-byte entries.  A row is one contiguous, 16-byte aligned span of up to
//      ~16 KB: it arrives in shared memory with ONE bulk async copy (cp.async.bulk, completion on an mbarrier) and leaves
//      with one bulk store — the copy engine moves it while the warp only waits, instead of len/32 rounds of per-lane
//      16-byte loads and stores. ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) k_rows_sort(const uint32_t* __restrict__ row_off, const uint32_t* __restrict__ list, uint32_t nlist, int cap,
                                                uint4* __restrict__ rows) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int lane = threadIdx.x;
  uint4* A = reinterpret_cast<uint4*>(smraw);
  uint4* B = A + cap;
  __shared__ unsigned s_cnt[64];
  __shared__ __align__(8) unsigned long long s_bar;
  if (blockIdx.x >= nlist) return;
  const int64_t v = list[blockIdx.x];
  const uint32_t off = row_off[v];
  const int len = (int)(row_off[v + 1] - off);
  if (len <= 1 || len > cap) return;
  const uint32_t bytes = (uint32_t)len * 16u;
  // --- global -> shared: one bulk copy, the warp waits on the mbarrier's transaction count ---
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the init is visible to the async proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(A)),
                 "l"(rows + off), "r"(bytes), "r"(smem_addr(&s_bar))
                 : "memory");
  }
  __syncwarp();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(&s_bar)) : "memory");
  }
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    const int shift = 18 + 6 * pass;
    s_cnt[lane] = 0; s_cnt[lane + 32] = 0;
    __syncwarp();
    for (int i0 = 0; i0 < len; i0 += 32) {
      const int i = i0 + lane;
      const unsigned d = i < len ? ((A[i].z >> shift) & 63u) : 0xffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (i < len && (peers & lt) == 0) s_cnt[d] += __popc(peers);
      __syncwarp();
    }
    {   // exclusive prefix sums over the 64 counters: lane owns bins 2*lane, 2*lane + 1
      const unsigned c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
      const unsigned inc = warp_incl_scan(c0 + c1, lane);
      __syncwarp();
      s_cnt[2 * lane] = inc - c0 - c1;
      s_cnt[2 * lane + 1] = inc - c1;
    }
    __syncwarp();
    for (int i0 = 0; i0 < len; i0 += 32) {
      const int i = i0 + lane;
      uint4 e = make_uint4(0, 0, 0, 0);
      unsigned d = 0xffffu;
      if (i < len) { e = A[i]; d = (e.z >> shift) & 63u; }
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      unsigned pos = 0;
      if (i < len) pos = s_cnt[d] + __popc(peers & lt);
      __syncwarp();
      if (i < len && (peers & lt) == 0) s_cnt[d] += __popc(peers);
      if (i < len) B[pos] = e;
      __syncwarp();
    }
    uint4* t = A; A = B; B = t;
  }
  // --- shared -> global: the warp's writes to A are made visible to the async proxy, then one bulk store ---
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rows + off), "r"(smem_addr(A)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the shared buffer is read before the CTA retires
  }
}

// ---- stage 5a: cutGraphSegmentation (VS.h:1913-2029) of one voxel per warp (= per CTA) from the weight rows.
//      The reference sorts all n^2 weights of the neighbourhood; the merge loop only ever acts on an entry whose two
//      vertices lie in different segments, and the emitted segment (that of local vertex 0) is final once the next
//      weight is <= thr(S0) = Int(S0) - k/|S0| (S0 rule, exact).  Here the warp walks the rows of its used neighbours
//      in rounds of descending weight cells [c0, c1): an entry of row j is the pair {j, partner at lattice offset d};
//      it belongs to this local graph iff the partner lies inside the neighbourhood (offset table lookup), and its two
//      directed weights are staged iff the two vertices are in different segments (the lighter direction waits in a small
//      deferred list when its cell is not part of the round).  Staged entries are sorted exactly (w desc, flat index asc)
//      and merged like the reference does; the round size adapts to the number of staged entries.
//      Output: lattice-offset bit mask of the connect list (+ its size); units this kernel cannot handle (staging
//      overflow inside one cell, deferred list full) go to the fallback list of the general kernel. ----
// bitonic sort of 32 * NR (weight, flat index) entries held NR per lane (element e = lane + 32 r) into the order
// (w desc, f asc): shuffles for partner distances < 32, register swaps inside the lane above
template <int NR>
__device__ __forceinline__ void warp_bitonic(float (&w)[NR], int (&f)[NR], int lane) {
  auto before = [](float wa, int fa, float wb, int fb) { return (wa > wb) || (wa == wb && fa < fb); };
#pragma unroll 1
  for (int kk = 2; kk <= 32 * NR; kk <<= 1) {
#pragma unroll 1
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      if (jj >= 32) {
        // partner = another register of the same lane; the register distance is a compile-time constant in every copy
#pragma unroll
        for (int dr = 1; dr < NR; dr <<= 1) {
          if (jj == 32 * dr) {
#pragma unroll
            for (int r = 0; r < NR; r++) {
              if ((r & dr) == 0) {
                const bool up = ((32 * r) & kk) == 0;       // lane < 32 <= jj < kk: only the register index decides
                const bool ok = up ? !before(w[r | dr], f[r | dr], w[r], f[r]) : !before(w[r], f[r], w[r | dr], f[r | dr]);
                if (!ok) { const float tw = w[r]; w[r] = w[r | dr]; w[r | dr] = tw; const int tf = f[r]; f[r] = f[r | dr]; f[r | dr] = tf; }
              }
            }
          }
        }
      } else {
        const bool lower = (lane & jj) == 0;
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const float wo = __shfl_xor_sync(0xffffffffu, w[r], jj);
          const int fo = __shfl_xor_sync(0xffffffffu, f[r], jj);
          const bool up = ((lane + 32 * r) & kk) == 0;
          const bool other_first = before(wo, fo, w[r], f[r]);
          if ((up == lower) ? other_first : !other_first) { w[r] = wo; f[r] = fo; }
        }
      }
    }
  }
}

constexpr int LR_NCAP = 184;     // used neighbours of one voxel (MAX_NEIGH = 181)
constexpr int LR_CS = 256;       // staging capacity (entries)
constexpr int LR_TARGET = 40;    // staged entries aimed at per round (<= 64: sorted in registers)
constexpr int LR_DEF = 16;       // deferred entries (lighter direction of a pair whose cells straddle a round boundary)
__host__ __device__ inline size_t lr_smem_bytes(int lbits, int mwords, int ncap) {
  // C_w, s_thr, s_kn, s_off6, s_cur, s_end, s_sav, s_mask, s_dw (4 B) | C_f, s_nc, s_df (2 B) | s_seg, s_size, s_loc (1 B);
  // ncap = vertex capacity (multiple of 4, >= the largest neighbourhood of the scene, <= LR_NCAP)
  size_t b = (size_t)LR_CS * 4 + (size_t)ncap * 4 * 6 + (size_t)mwords * 4 + (size_t)LR_DEF * 4 + (size_t)LR_CS * 2 + (size_t)ncap * 2 +
             (size_t)LR_DEF * 4 + (size_t)ncap * 2 + ((size_t)1 << (3 * lbits));
  return (b + 15) & ~(size_t)15;
}
#ifndef VGS_LR_WARPS
#define VGS_LR_WARPS 1        // measured: 1 voxel per CTA 6.07 ms, 2 per CTA 6.7-7.2 ms (10 M-point site)
#endif
#ifndef VGS_LR_MINB
#define VGS_LR_MINB 32       // <= 64 registers: 32 resident CTAs per SM (measured 5.16 -> 5.03 ms against 28)
#endif
constexpr int LR_WARPS = VGS_LR_WARPS;      // voxels per CTA (independent warps, no block barrier)
__global__ void __launch_bounds__(32 * LR_WARPS, VGS_LR_MINB) k_local_graph_rows(const uint32_t* __restrict__ used_list, uint32_t n_used, const uint32_t* __restrict__ adj_off,
                                                            const int32_t* __restrict__ adj_idx, const uint16_t* __restrict__ adj_code,
                                                            const uint8_t* __restrict__ uflags, float k, int rho, int lbits, int mwords, int ncap,
                                                            const uint32_t* __restrict__ row_off, const uint4* __restrict__ rows,
                                                            const float* __restrict__ wempty, uint32_t* __restrict__ conn_cnt,
                                                            uint32_t* __restrict__ conn_mask, uint32_t* __restrict__ fallback,
                                                            uint32_t* __restrict__ fallback_count, int force_fb_mod, int target,
                                                            unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smraw_all[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t li = (int64_t)blockIdx.x * LR_WARPS + wid;
  if (li >= (int64_t)n_used) return;
  const int64_t u = used_list[li];          // unused voxels keep the empty connect lists the host memset gave them (VS.h:403-409)
  unsigned char* smraw = smraw_all + (size_t)wid * lr_smem_bytes(lbits, mwords, ncap);
  float* C_w = reinterpret_cast<float*>(smraw);                         // LR_CS
  float* s_thr = C_w + LR_CS;                                           // ncap: Int(C) - k/|C| of segment C
  float* s_kn = s_thr + ncap;                                           // ncap: k / n, n = 1 .. ncap (VS.h:1963: float / int)
  uint32_t* s_off6 = reinterpret_cast<uint32_t*>(s_kn + ncap);          // ncap: pack6(o + rho) of the vertex
  uint32_t* s_cur = s_off6 + ncap;                                      // ncap: next unread entry of the vertex's row
  uint32_t* s_end = s_cur + ncap;                                       // ncap
  uint32_t* s_sav = s_end + ncap;                                       // ncap: cursors at the start of the round
  uint32_t* s_mask = s_sav + ncap;                                      // mwords: output mask
  float* s_dw = reinterpret_cast<float*>(s_mask + mwords);              // LR_DEF: deferred weights
  unsigned short* C_f = reinterpret_cast<unsigned short*>(s_dw + LR_DEF);     // LR_CS
  unsigned short* s_nc = C_f + LR_CS;                                   // ncap: cell of the next unread entry (ROW_CELLS = exhausted)
  unsigned short* s_df = s_nc + ncap;                                   // LR_DEF x 2: deferred (flat index, cell)
  unsigned char* s_seg = reinterpret_cast<unsigned char*>(s_df + 2 * LR_DEF); // ncap
  unsigned char* s_size = s_seg + ncap;                                 // ncap
  unsigned char* s_loc = s_size + ncap;                                 // 1 << 3*lbits: lattice offset -> local vertex, 0xff = none
  __shared__ int s_cnt_w[LR_WARPS], s_ndef_w[LR_WARPS];
  int& s_cnt = s_cnt_w[wid];
  int& s_ndef = s_ndef_w[wid];
  const uint32_t lt = (1u << lane) - 1u;
  const int nloc = 1 << (3 * lbits);
  const uint32_t lmask = (1u << lbits) - 1u;
  const int S1 = 2 * rho + 1;

  for (int i = lane; i < mwords; i += 32) s_mask[i] = 0;
  for (int i = lane; i < nloc / 4; i += 32) reinterpret_cast<uint32_t*>(s_loc)[i] = 0xffffffffu;
  if (lane == 0) s_ndef = 0;
  __syncwarp();
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  int nv = 0;
  // vertex table of the USED neighbours in adjacency order (local vertex 0 = the voxel itself); unused voxels cannot
  // merge (their weight is w_empty <= cut bound, else the unit goes to the general kernel), and the tie-break order
  // (col * n + row over all neighbours, VS.h:1922) is preserved by the monotone renumbering
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int i = b0 + lane;
    bool us = false;
    int64_t gq = 0;
    uint32_t c5 = 0;
    if (i < n) {
      gq = adj_idx[off + i];
      c5 = adj_code[off + i];
      us = (__ldg(uflags + gq) & F_USED) != 0;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, us);
    if (us) {
      const int j = nv + __popc(bal & lt);
      if (j < ncap) {
        const uint32_t ox = c5 & 31u, oy = (c5 >> 5) & 31u, oz = (c5 >> 10) & 31u;
        s_off6[j] = pack6((int)ox, (int)oy, (int)oz);
        s_loc[ox | (oy << lbits) | (oz << (2 * lbits))] = (unsigned char)j;
        const uint32_t r0 = __ldg(row_off + gq), r1 = __ldg(row_off + gq + 1);
        s_cur[j] = r0; s_end[j] = r1;
        s_nc[j] = r0 < r1 ? (unsigned short)(__ldg(&rows[r0].z) >> 18) : (unsigned short)ROW_CELLS;
        s_seg[j] = (unsigned char)j; s_size[j] = 1; s_thr[j] = 1.0f - k / 1.0f;
        s_kn[j] = k / (float)(j + 1);
      }
    }
    nv += __popc(bal);
  }
  __syncwarp();
  const float lb = (float)(1.0 - 2.0 * (double)k + (double)k / (double)n - 4e-7 * (double)(n + 8));
  bool to_fallback = (wempty[0] > lb) || nv > ncap;     // empty pairs could merge / too many vertices: general kernel
  if (force_fb_mod > 0 && (u % force_fb_mod) == 0) to_fallback = true;   // test knob VGS_B200_FORCE_FALLBACK
  int nseg = nv;
  bool stop = false;
  // one batch of <= 32 entries in descending order, one per lane: the first mergeable entry merges,
  // the later ones are re-evaluated against the new state (cutGraphSegmentation VS.h:1955-2001)
  auto merge_batch = [&](float w, int f, bool valid) {
    const int v1 = f >> 8, v2 = f & 255;
    // the lane's two segments live in registers and follow the merges of this batch
    int sa = 0, sb = 0;
    float m1 = 0.f, m2 = 0.f;
    if (valid) { sa = s_seg[v1]; sb = s_seg[v2]; m1 = s_thr[sa]; m2 = s_thr[sb]; }
    uint32_t todo = __ballot_sync(0xffffffffu, valid && sa != sb);
    while (todo) {
      const bool mine = (todo >> lane) & 1u;
      const bool a_wins = (m1 >= m2);
      const bool pred = mine && w > (a_wins ? m1 : m2);
      const uint32_t bal = __ballot_sync(0xffffffffu, pred);
      if (!bal) break;
      const int Lm = __ffs(bal) - 1;
      const int packed = __shfl_sync(0xffffffffu, a_wins ? (sa | (sb << 8)) : (sb | (sa << 8)), Lm);
      const int keepl = packed & 255, drop = packed >> 8;
      const float wl = __shfl_sync(0xffffffffu, w, Lm);
      {   // relabel, four vertices per word (bytes past nv are never read)
        const uint32_t drop4 = (uint32_t)drop * 0x01010101u, keep4 = (uint32_t)keepl * 0x01010101u;
        uint32_t* seg32 = reinterpret_cast<uint32_t*>(s_seg);
        for (int v4 = lane; v4 * 4 < nv; v4 += 32) {
          const uint32_t x = seg32[v4];
          const uint32_t m = __vcmpeq4(x, drop4);
          if (m) seg32[v4] = (x & ~m) | (keep4 & m);
        }
      }
      const int nsz = (int)s_size[keepl] + (int)s_size[drop];
      const float nthr = wl - s_kn[nsz - 1];          // Int(C) - k/|C|; a table: an IEEE division here doubles the kernel time
      __syncwarp();
      if (lane == 0) { s_thr[keepl] = nthr; s_size[keepl] = (unsigned char)nsz; s_size[drop] = 0; }
      if (sa == drop) sa = keepl;
      if (sb == drop) sb = keepl;
      if (sa == keepl) m1 = nthr;
      if (sb == keepl) m2 = nthr;
      nseg--;
      __syncwarp();
      todo &= ~((2u << Lm) - 1u);
      todo &= __ballot_sync(0xffffffffu, sa != sb);
      if (nseg <= 1) { stop = true; break; }
    }
  };
  if (!to_fallback && nv > 1) {
    // field-parallel range test of pack6(o_j + rho) + pack6(d + 2 rho) = pack6(p + 3 rho): p inside the neighbourhood
    // cube iff every field lies in [2 rho, 4 rho]
    const uint32_t G = 0x20820u, ONES = 0x1041u;
    const uint32_t add_hi = (uint32_t)(31 - 4 * rho) * ONES, sub_lo = (uint32_t)(2 * rho) * ONES;
    int c0 = ROW_CELLS;
    for (int j = lane; j < nv; j += 32) c0 = min(c0, (int)s_nc[j]);
    c0 = __reduce_min_sync(0xffffffffu, c0);
    int span = ROW_CELLS / 256;
    while (!stop && c0 < ROW_CELLS) {
      // S0 rule on the cell bound: every unread or deferred entry has a weight <= row_cell_upper(c0)
      if (!(row_cell_upper(c0) > s_thr[s_seg[0]])) break;
      const int c1 = min(ROW_CELLS, c0 + span);
      const int ndef0 = s_ndef;
      __syncwarp();
      if (lane == 0) s_cnt = 0;
      __syncwarp();
      int cmin = ROW_CELLS;
      for (int j = lane; j < nv; j += 32) {
        int nc = (int)s_nc[j];
        uint32_t c = s_cur[j];
        s_sav[j] = c;
        if (nc >= c1) { cmin = min(cmin, nc); continue; }
        const uint32_t e = s_end[j];
        const uint32_t oj = s_off6[j];
        const int sj = s_seg[j];
        uint4 en = __ldg(&rows[c]);
        while (true) {
          if ((int)(en.z >> 18) >= c1) { nc = (int)(en.z >> 18); break; }
          c++;
          uint4 nx = en;
          if (c < e) nx = __ldg(&rows[c]);        // the next entry is in flight while this one is filed
          const uint32_t sum = oj + (en.z & 0x3ffffu);
          const bool inside = (((sum + add_hi) & G) == 0u) && ((((sum | G) - sub_lo) & G) == G);
          if (inside) {
            const uint32_t p = sum - sub_lo;
            const int i = s_loc[(p & lmask) | (((p >> 6) & lmask) << lbits) | (((p >> 12) & lmask) << (2 * lbits))];
            if (i != 0xff && s_seg[i] != sj) {
              // entry (row a, col b) = weight(idx[a] -> idx[b]); packed (col << 8) | row orders like col*n+row (VS.h:1922)
              const float w_ji = __uint_as_float(en.x), w_ij = __uint_as_float(en.y);
              const unsigned short f_ji = (unsigned short)((i << 8) | j), f_ij = (unsigned short)((j << 8) | i);
              if (en.x == en.y) {
                // both directions carry the same weight (bit for bit): the second one in the order (larger flat index)
                // can never act — if the first merges, the second joins one segment; if the first is refused because a
                // threshold is >= w, no entry of the same weight in between can merge (hence lower) that segment
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < LR_CS) { C_w[slot] = w_ji; C_f[slot] = min(f_ji, f_ij); }
              } else if ((int)en.w < c1) {       // both directions belong to this round
                const int slot = atomicAdd(&s_cnt, 2);
                if (slot + 1 < LR_CS) { C_w[slot] = w_ji; C_f[slot] = f_ji; C_w[slot + 1] = w_ij; C_f[slot + 1] = f_ij; }
              } else {                    // the lighter direction waits for its cell
                const bool ji_heavy = row_cell(w_ji) <= row_cell(w_ij);
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < LR_CS) { C_w[slot] = ji_heavy ? w_ji : w_ij; C_f[slot] = ji_heavy ? f_ji : f_ij; }
                const int ds = atomicAdd(&s_ndef, 1);
                if (ds < LR_DEF) { s_dw[ds] = ji_heavy ? w_ij : w_ji; s_df[2 * ds] = ji_heavy ? f_ij : f_ji; s_df[2 * ds + 1] = (unsigned short)en.w; }
              }
            }
          }
          if (c >= e) { nc = ROW_CELLS; break; }
          en = nx;
        }
        s_cur[j] = c;
        s_nc[j] = (unsigned short)nc;
        cmin = min(cmin, nc);
      }
      __syncwarp();
      if (s_ndef > LR_DEF) { to_fallback = true; break; }
      // deferred entries whose cell has come up join the round (those added in this round have cells >= c1)
      int took_cell = -1;     // cell of the deferred entry this lane consumed in this round (for a roll-back)
      {
        const int nd = min(s_ndef, LR_DEF);
        if (lane < nd) {
          const int dc = (int)s_df[2 * lane + 1];
          if (lane < ndef0 && dc < c1) {
            const int f = (int)s_df[2 * lane];
            if (s_seg[f >> 8] != s_seg[f & 255]) {
              const int slot = atomicAdd(&s_cnt, 1);
              if (slot < LR_CS) { C_w[slot] = s_dw[lane]; C_f[slot] = (unsigned short)f; }
            }
            took_cell = dc;
            s_df[2 * lane + 1] = (unsigned short)0xffff;     // consumed (kept in place: the list is tiny)
          } else if (dc != 0xffff) cmin = min(cmin, dc);     // still waiting: the next round must not start behind it
        }
      }
      __syncwarp();
      const int kept = s_cnt;
      if (kept > LR_CS) {
        if (span == 1) { to_fallback = true; break; }       // one cell alone overflows the staging buffer
        // roll the cursors and the deferred list back and retry with a narrower round
        for (int j = lane; j < nv; j += 32) {
          const uint32_t sv = s_sav[j];
          if (s_cur[j] != sv) { s_cur[j] = sv; s_nc[j] = (unsigned short)(__ldg(&rows[sv].z) >> 18); }
        }
        if (took_cell >= 0) s_df[2 * lane + 1] = (unsigned short)took_cell;
        if (lane == 0) s_ndef = ndef0;
        __syncwarp();
        span = max(1, span / 4);
        continue;
      }
      if (dbg && lane == 0) { atomicAdd(&dbg[0], 1ull); atomicAdd(&dbg[2], (unsigned long long)kept); }
      if (kept > 0) {
        // entries sorted across the warp in registers, NR per lane, (w desc, packed index asc), written back in order
        // and merged in batches of 32 (one copy of the merge loop: the kernel has to stay inside the instruction cache)
        auto sort_regs = [&](auto tag) {
          constexpr int NR = decltype(tag)::value;
          float w[NR]; int f[NR];
#pragma unroll
          for (int r = 0; r < NR; r++) {
            const int e = lane + 32 * r;
            w[r] = e < kept ? C_w[e] : -1.0f;
            f[r] = e < kept ? (int)C_f[e] : 0xffff;
          }
          warp_bitonic<NR>(w, f, lane);
#pragma unroll
          for (int r = 0; r < NR; r++) { C_w[lane + 32 * r] = w[r]; C_f[lane + 32 * r] = (unsigned short)f[r]; }
        };
        if (kept <= 32) sort_regs(std::integral_constant<int, 1>{});
        else if (kept <= 64) sort_regs(std::integral_constant<int, 2>{});
        else if (kept <= 128) sort_regs(std::integral_constant<int, 4>{});
        else sort_regs(std::integral_constant<int, 8>{});
        __syncwarp();
#pragma unroll 1
        for (int bs = 0; bs < kept && !stop; bs += 32) {
          const int e = bs + lane;
          merge_batch(C_w[e], (int)C_f[e], e < kept);
        }
      }
      // next round: first cell that still holds an unread / deferred entry; size adapted to the yield of this one
      c0 = __reduce_min_sync(0xffffffffu, cmin);
      span = min(ROW_CELLS / 4, max(1, (span * (target + 4)) / (kept + 4)));
    }
    if (dbg && lane == 0) { atomicAdd(&dbg[3], 1ull); atomicAdd(&dbg[5], (unsigned long long)nseg); atomicAdd(&dbg[6], (unsigned long long)nv); }
  }
  if (to_fallback) {
    if (lane == 0) fallback[atomicAdd(fallback_count, 1u)] = (uint32_t)u;
    return;
  }
  // --- emit the segment that contains local vertex 0 (the voxel itself) as a lattice-offset mask ---
  __syncwarp();
  const int s0 = s_seg[0];
  int cnt = 0;
  for (int b = 0; b < nv; b += 32) {
    const int v = b + lane;
    const bool in = v < nv && s_seg[v] == s0;
    if (in) {
      const uint32_t o6 = s_off6[v];
      const int lin = ((int)(o6 & 63u) * S1 + (int)((o6 >> 6) & 63u)) * S1 + (int)((o6 >> 12) & 63u);
      atomicOr(&s_mask[lin >> 5], 1u << (lin & 31));
    }
    cnt += __popc(__ballot_sync(0xffffffffu, in));
  }
  __syncwarp();
  for (int i = lane; i < mwords; i += 32) conn_mask[(size_t)u * mwords + i] = s_mask[i];
  if (lane == 0) conn_cnt[u] = (uint32_t)cnt;
}

// general-kernel results (connect LIST at the adjacency offsets) -> lattice mask, for the units of the fallback list
__global__ void __launch_bounds__(128) k_conn_list_to_mask(const uint32_t* __restrict__ list, uint32_t nlist, const uint32_t* __restrict__ adj_off,
                                                         const int32_t* __restrict__ adj_idx, const uint16_t* __restrict__ adj_code,
                                                         const uint32_t* __restrict__ conn_cnt, const int32_t* __restrict__ conn_idx, int rho,
                                                         int mwords, uint32_t* __restrict__ conn_mask) {
  const int lane = threadIdx.x & 31;
  const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (li >= nlist) return;
  const uint32_t u = list[li];
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off), c = (int)conn_cnt[u];
  const int S1 = 2 * rho + 1;
  for (int i = lane; i < mwords; i += 32) conn_mask[(size_t)u * mwords + i] = 0;
  __syncwarp();
  // both lists are in adjacency order: walk them together
  for (int e = lane; e < n; e += 32) {
    const int g = adj_idx[off + e];
    bool in = false;
    for (int t = 0; t < c; t++) if (conn_idx[off + t] == g) { in = true; break; }
    if (in) {
      const uint32_t c5 = adj_code[off + e];
      const int lin = ((int)(c5 & 31u) * S1 + (int)((c5 >> 5) & 31u)) * S1 + (int)((c5 >> 10) & 31u);
      atomicOr(&conn_mask[(size_t)u * mwords + (lin >> 5)], 1u << (lin & 31));
    }
  }
}

// lattice mask -> connect list at the adjacency offsets, adjacency order (debug blobs / SVGS-style consumers)
__global__ void __launch_bounds__(128) k_conn_mask_to_list(const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                         const uint16_t* __restrict__ adj_code, int64_t nu, int rho, int mwords,
                                                         const uint32_t* __restrict__ conn_mask, int32_t* __restrict__ conn_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  const int S1 = 2 * rho + 1;
  int kept = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    bool in = false;
    int g = -1;
    if (e < n) {
      g = adj_idx[off + e];
      const uint32_t c5 = adj_code[off + e];
      const int lin = ((int)(c5 & 31u) * S1 + (int)((c5 >> 5) & 31u)) * S1 + (int)((c5 >> 10) & 31u);
      in = (conn_mask[(size_t)u * mwords + (lin >> 5)] >> (lin & 31)) & 1u;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, in);
    if (in) conn_idx[off + kept + __popc(bal & ((1u << lane) - 1u))] = g;
    kept += __popc(bal);
  }
}

// ---- stage 5b: crossValidation (VS.h:2111-2179) on the masks: j stays in L[i] iff i is in L[j], i.e. iff the bit of the
//      mirrored offset is set in the mask of j.  Lists of size <= 1 are left as they are.  One warp per voxel. ----
__global__ void __launch_bounds__(128) k_mutual_mask(const uint32_t* __restrict__ adj_off, const int32_t* __restrict__ adj_idx,
                                                   const uint16_t* __restrict__ adj_code, const uint32_t* __restrict__ cnt0,
                                                   const uint32_t* __restrict__ mask0, int64_t nu, int rho, int mwords,
                                                   uint32_t* __restrict__ cnt1, int32_t* __restrict__ idx1) {
  const int lane = threadIdx.x & 31;
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nu) return;
  const int c = (int)cnt0[u];
  if (c == 0) { if (lane == 0) cnt1[u] = 0; return; }
  const uint32_t off = adj_off[u];
  const int n = (int)(adj_off[u + 1] - off);
  const int S1 = 2 * rho + 1, top = S1 * S1 * S1 - 1;
  int kept = 0;
  for (int b = 0; b < n; b += 32) {
    const int e = b + lane;
    bool keep = false;
    int j = -1;
    if (e < n) {
      j = adj_idx[off + e];
      const uint32_t c5 = adj_code[off + e];
      const int lin = ((int)(c5 & 31u) * S1 + (int)((c5 >> 5) & 31u)) * S1 + (int)((c5 >> 10) & 31u);
      if ((mask0[(size_t)u * mwords + (lin >> 5)] >> (lin & 31)) & 1u) {
        if (c <= 1) keep = true;
        else {
          const int ml = top - lin;                    // offset of u seen from j
          keep = (__ldg(&mask0[(size_t)j * mwords + (ml >> 5)]) >> (ml & 31)) & 1u;
        }
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (keep) idx1[off + kept + __popc(bal & ((1u << lane) - 1u))] = j;
    kept += __popc(bal);
  }
  if (lane == 0) cnt1[u] = (uint32_t)kept;
}

}  // namespace vgs

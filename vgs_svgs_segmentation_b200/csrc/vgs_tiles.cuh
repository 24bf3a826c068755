// vgs_tiles.cuh — kernels of the multi-GPU split: ONE scene cut into spatial slabs of the voxel
// lattice along one axis, every rank segments its slab plus a halo of voxel layers, components are merged across
// slabs.  The reference is single-process (VS.h:372-421 runs on one cloud); what must survive the split is its
// result: voxel ids are PCL's global leaf order and closestCheck (VS.h:2181-2303) is sequential in that order and
// reads slot 0 of an adjacency list — the neighbour COUNT — as a voxel id (VS.h:2243), so the split keeps
//   * the global origin / growth epochs (k_quantise with global point indices),
//   * the relative voxel order (local ids are the global order restricted to the tile),
//   * a table of the globally first 256 voxels (the only ids a neighbour count can name),
//   * closest-check rounds with an exchange of the boundary singles' state between rounds,
//   * a cross-slab union-find over voxel keys, cluster sizes and smallest point indices summed over ranks.
#pragma once
#include "vgs_rows.cuh"

namespace vgs {

constexpr int TILE_MAX_RANKS = 16;
constexpr int LOW_IDS = 256;                       // a neighbour count is <= 255
constexpr int32_t ATT_REMOTE = 0x7fffffff;         // halo single: attached according to its owner
constexpr int32_t ATT_FAR = 0x40000000;            // | global voxel id (< 256): partner named by a neighbour count

struct SlabCuts {
  int axis, nranks, halo;
  int cut[TILE_MAX_RANKS + 1];                     // slab r owns key[axis] in [cut[r], cut[r+1])
};

// octree key of one point (genOctreeKeyforPoint with PCL's growth epochs; same arithmetic as k_quantise)
__device__ __forceinline__ void point_key3(float x, float y, float z, long long gidx, const EpochTable& ep, double res, uint32_t& kx,
                                           uint32_t& ky, uint32_t& kz) {
  int e = ep.n - 1;
  while (e > 0 && gidx < ep.viol[e]) e--;
  kx = (uint32_t)(((double)x - ep.mn[e][0]) / res) + ep.shift[e][0];
  ky = (uint32_t)(((double)y - ep.mn[e][1]) / res) + ep.shift[e][1];
  kz = (uint32_t)(((double)z - ep.mn[e][2]) / res) + ep.shift[e][2];
}

// per-axis histogram of the key coordinates of every `step`-th point of this rank's slice (the cuts only balance the
// load, a deterministic subsample is as good as the full slice): hist[a * nbins + (key_a >> shift)].  Shared-memory
// counters per CTA (a flat ground puts most points into one z bin), flushed with one atomic per touched bin.
__global__ void __launch_bounds__(256) k_slab_hist(const float* __restrict__ xyz, int stride, int64_t n, long long gfirst, EpochTable ep, double res,
                                                 int nbins, int shift, int step, unsigned long long* __restrict__ hist) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < 3 * nbins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int64_t ns = (n + step - 1) / step;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = s * step;
    const float* p = xyz + i * stride;
    const float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) continue;
    uint32_t kx, ky, kz;
    point_key3(x, y, z, gfirst + i, ep, res, kx, ky, kz);
    kx >>= shift; ky >>= shift; kz >>= shift;
    if ((int)kx < nbins) atomicAdd(&sh[kx], 1u);
    if ((int)ky < nbins) atomicAdd(&sh[nbins + ky], 1u);
    if ((int)kz < nbins) atomicAdd(&sh[2 * nbins + kz], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nbins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// destination mask of a point: its owner slab and every slab whose halo reaches it
__device__ __forceinline__ uint32_t slab_dests(int ka, const SlabCuts& sc) {
  uint32_t m = 0;
  for (int r = 0; r < sc.nranks; r++)
    if (ka >= sc.cut[r] - sc.halo && ka < sc.cut[r + 1] + sc.halo) m |= 1u << r;
  return m;
}
constexpr int ROUTE_PER_WARP = 256;     // consecutive points per warp: the stable partition keeps the index order
// pass 1: per warp and destination, the number of points sent there: cnt[d * nwarps + warp]
__global__ void __launch_bounds__(256) k_route_count(const float* __restrict__ xyz, int stride, int64_t n, long long gfirst, EpochTable ep,
                                                   double res, SlabCuts sc, int64_t nwarps, uint32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= nwarps) return;
  uint32_t c[TILE_MAX_RANKS];
#pragma unroll
  for (int d = 0; d < TILE_MAX_RANKS; d++) c[d] = 0;
  for (int r = 0; r < ROUTE_PER_WARP / 32; r++) {
    const int64_t i = wg * ROUTE_PER_WARP + r * 32 + lane;
    uint32_t m = 0;
    if (i < n) {
      const float* p = xyz + i * stride;
      const float x = p[0], y = p[1], z = p[2];
      if (finite3(x, y, z)) {
        uint32_t k3[3];
        point_key3(x, y, z, gfirst + i, ep, res, k3[0], k3[1], k3[2]);
        m = slab_dests((int)k3[sc.axis], sc);
      }
    }
#pragma unroll
    for (int d = 0; d < TILE_MAX_RANKS; d++)
      if (d < sc.nranks) c[d] += __popc(__ballot_sync(0xffffffffu, (m >> d) & 1u));
  }
  if (lane == 0)
    for (int d = 0; d < sc.nranks; d++) cnt[(int64_t)d * nwarps + wg] = c[d];
}
// pass 2: records (x, y, z, global index bits) grouped by destination, index order kept inside a group
__global__ void __launch_bounds__(256) k_route_scatter(const float* __restrict__ xyz, int stride, int64_t n, long long gfirst, EpochTable ep,
                                                     double res, SlabCuts sc, int64_t nwarps, const uint32_t* __restrict__ offs,
                                                     float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= nwarps) return;
  uint32_t base[TILE_MAX_RANKS];
#pragma unroll
  for (int d = 0; d < TILE_MAX_RANKS; d++) base[d] = d < sc.nranks ? offs[(int64_t)d * nwarps + wg] : 0u;
  const uint32_t lt = (1u << lane) - 1u;
  for (int r = 0; r < ROUTE_PER_WARP / 32; r++) {
    const int64_t i = wg * ROUTE_PER_WARP + r * 32 + lane;
    uint32_t m = 0;
    float4 recd = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      const float* p = xyz + i * stride;
      recd.x = p[0]; recd.y = p[1]; recd.z = p[2];
      recd.w = __uint_as_float((uint32_t)(gfirst + i));
      if (finite3(recd.x, recd.y, recd.z)) {
        uint32_t k3[3];
        point_key3(recd.x, recd.y, recd.z, gfirst + i, ep, res, k3[0], k3[1], k3[2]);
        m = slab_dests((int)k3[sc.axis], sc);
      }
    }
#pragma unroll
    for (int d = 0; d < TILE_MAX_RANKS; d++) {
      if (d < sc.nranks) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (m >> d) & 1u);
        if ((m >> d) & 1u) out[base[d] + __popc(bal & lt)] = recd;
        base[d] += __popc(bal);
      }
    }
  }
}

// ownership of the local voxels: bit 0 = owned (key[axis] inside the slab), bit 1 = shared with another rank
// (inside somebody's halo, or a halo voxel here)
__global__ void __launch_bounds__(256) k_tile_owner(const uint32_t* __restrict__ key3, int64_t nu, SlabCuts sc, int rank, uint8_t* __restrict__ own) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const int ka = (int)key3[3 * u + sc.axis];
  const bool owned = ka >= sc.cut[rank] && ka < sc.cut[rank + 1];
  const uint32_t m = slab_dests(ka, sc);
  own[u] = (owned ? 1 : 0) | ((!owned || (m & ~(1u << rank))) ? 2 : 0);
}

// ---- table of the globally first LOW_IDS voxels: every rank contributes its first owned voxels (local id order =
//      global order), the host merges them by sort key ----
struct LowEntry {
  unsigned long long sort_key;   // the octree sort key that defines the voxel id order (ukey)
  unsigned long long plain;      // plain morton key (hash key of the voxel)
  float rec[REC_FLOATS];
  uint32_t cnt1;                 // connect-list size after the mutual filter
  int32_t local;                 // local id on the contributing rank
};
__global__ void __launch_bounds__(256) k_own_flags(const uint8_t* __restrict__ own, int64_t nu, uint32_t* __restrict__ flags) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < nu) flags[u] = own[u] & 1u;
}
// rank[u] = number of owned voxels before u (exclusive scan of the flags): the first LOW_IDS owned voxels are exported
__global__ void __launch_bounds__(256) k_low_export(const uint8_t* __restrict__ own, const uint32_t* __restrict__ rank,
                                                  const unsigned long long* __restrict__ ukey, const unsigned long long* __restrict__ plain,
                                                  const float* __restrict__ rec, const uint32_t* __restrict__ cnt1, int64_t nu,
                                                  LowEntry* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu || !(own[u] & 1)) return;
  const uint32_t pos = rank[u];
  if (pos >= (uint32_t)LOW_IDS) return;
  LowEntry e;
  e.sort_key = ukey[u]; e.plain = plain[u]; e.cnt1 = cnt1[u]; e.local = (int32_t)u;
  for (int q = 0; q < REC_FLOATS; q++) e.rec[q] = rec[u * REC_FLOATS + q];
  out[pos] = e;
}
// gidlo[u] = global voxel id if u is one of the globally first LOW_IDS voxels, else 0xffff
__global__ void k_low_import(const LowEntry* __restrict__ low, int n_low, const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                             uint64_t hmask, uint16_t* __restrict__ gidlo) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_low) return;
  const int u = hash_lookup(tk, tv, hmask, low[p].plain);
  if (u >= 0) gidlo[u] = (uint16_t)p;
}

// ---- closestCheck round of the OWNED singles of a tile (k_closest_round_warp with the global low-id table for the
//      candidate that the reference reads from the COUNT slot) ----
__global__ void __launch_bounds__(128) k_closest_round_tile(const uint32_t* __restrict__ list, uint32_t nlist, const uint32_t* __restrict__ adj_off,
                                                          const int32_t* __restrict__ adj_idx, const uint32_t* __restrict__ cnt1,
                                                          const float* __restrict__ rec, int64_t nu, PairParams pp, const uint8_t* __restrict__ own,
                                                          const LowEntry* __restrict__ low, const int32_t* __restrict__ low_att, int n_low,
                                                          const uint16_t* __restrict__ gidlo, int32_t* attach, uint32_t* __restrict__ changed) {
  const int lane = threadIdx.x & 31;
  const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (li >= nlist) return;
  const int64_t i = list[li];
  if (!(own[i] & 1)) return;            // halo singles get their state from their owner
  const uint32_t off = adj_off[i];
  const int n = (int)(adj_off[i + 1] - off);
  float ri[REC_FLOATS], rc[REC_FLOATS];
  for (int q = 0; q < REC_FLOATS; q++) ri[q] = rec[i * REC_FLOATS + q];
  const int gi = (int)gidlo[i];          // global id if < LOW_IDS, else 0xffff (larger than any count)
  float best = 0.f;
  int bj = -1, bi = -1;
  for (int j = lane; j <= n; j += 32) {
    if (j == 0) {
      // slot 0 of the adjacency list is the neighbour COUNT, read as a (global) voxel id (VS.h:2243)
      const int c = n;
      if (c >= n_low) continue;
      const uint32_t cc = low[c].cnt1;
      if (!(cc > 1u || (cc == 1u && c < gi && low_att[c]))) continue;
      for (int t = 0; t < REC_FLOATS; t++) rc[t] = low[c].rec[t];
      float w_ab, w_ba;
      pair_weights(ri, rc, pp, w_ab, w_ba);
      if (w_ab >= best) { best = w_ab; bj = 0; bi = ATT_FAR | c; }
      continue;
    }
    const int64_t c = (int64_t)adj_idx[off + j - 1];
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
// state of the owned singles another rank can see (shared voxels): plain key | attached << 63
__global__ void __launch_bounds__(256) k_boundary_export(const uint32_t* __restrict__ list, uint32_t nlist, const uint8_t* __restrict__ own,
                                                       const unsigned long long* __restrict__ plain, const int32_t* __restrict__ attach,
                                                       unsigned long long* __restrict__ out, uint32_t* __restrict__ n_out) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= nlist) return;
  const uint32_t i = list[li];
  if ((own[i] & 3) != 3) return;        // owned and visible elsewhere
  out[atomicAdd(n_out, 1u)] = plain[i] | (attach[i] >= 0 ? (1ull << 63) : 0ull);
}
__global__ void __launch_bounds__(256) k_boundary_import(const unsigned long long* __restrict__ in, int64_t n_in, const unsigned long long* __restrict__ tk,
                                                       const uint32_t* __restrict__ tv, uint64_t hmask, const uint8_t* __restrict__ own,
                                                       int32_t* __restrict__ attach, uint32_t* __restrict__ changed) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_in) return;
  const unsigned long long v = in[t];
  if (v == ~0ull) return;               // padding
  const int u = hash_lookup(tk, tv, hmask, v & ~(1ull << 63));
  if (u < 0 || (own[u] & 1)) return;    // not here, or my own voxel
  const int32_t want = (v >> 63) ? ATT_REMOTE : -1;
  if (attach[u] != want) { attach[u] = want; *changed = 1u; }
}
// attached flags of the low-id voxels this rank owns
__global__ void k_low_attached(const LowEntry* __restrict__ low, int n_low, const uint16_t* __restrict__ gidlo, const uint8_t* __restrict__ own,
                               const int32_t* __restrict__ attach, int64_t nu, int32_t* __restrict__ out /* LOW_IDS, zeroed */) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const int g = (int)gidlo[u];
  if (g < n_low && (own[u] & 1) && attach[u] >= 0) out[g] = 1;
}

// ---- cross-slab merge ----
// (key(v), key(root(v))) of every voxel another rank can name and that is not its own root — the shared voxels, and the
// owned ones among the globally first LOW_IDS voxels (named by neighbour counts) — plus (key(i), key of the far
// partner) of the owned singles attached through the count slot
__global__ void __launch_bounds__(256) k_pairs_export(const uint8_t* __restrict__ own, const uint16_t* __restrict__ gidlo, const int* __restrict__ root,
                                                    const unsigned long long* __restrict__ plain, const int32_t* __restrict__ attach,
                                                    const LowEntry* __restrict__ low, int64_t nu, unsigned long long* __restrict__ pairs,
                                                    unsigned long long* __restrict__ n_pairs) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const int r = root[u];
  const bool named = (own[u] & 2) || ((own[u] & 1) && gidlo[u] != 0xffff);
  if (named && r != (int)u) {
    const unsigned long long s = atomicAdd(n_pairs, 1ull);
    pairs[2 * s] = plain[u]; pairs[2 * s + 1] = plain[r];
  }
  const int32_t a = attach[u];
  if ((own[u] & 1) && a >= ATT_FAR && a != ATT_REMOTE) {
    const unsigned long long s = atomicAdd(n_pairs, 1ull);
    pairs[2 * s] = plain[u]; pairs[2 * s + 1] = low[a & 0xffff].plain;
  }
}
// union-find over the keys of all ranks' pairs: slot of a key in the open-addressing table = node id
__device__ __forceinline__ int key_slot(const unsigned long long* __restrict__ tk, uint64_t mask, uint64_t key) {
  uint64_t s = hash64(key) & mask;
  while (true) {
    const unsigned long long k = tk[s];
    if (k == key) return (int)s;
    if (k == HASH_EMPTY) return -1;
    s = (s + 1) & mask;
  }
}
__global__ void __launch_bounds__(256) k_merge_insert(const unsigned long long* __restrict__ pairs, int64_t n2, unsigned long long* __restrict__ tk,
                                                    uint64_t mask) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n2) return;
  const uint64_t key = pairs[t];
  uint64_t s = hash64(key) & mask;
  while (true) {
    const unsigned long long old = atomicCAS(&tk[s], (unsigned long long)HASH_EMPTY, (unsigned long long)key);
    if (old == HASH_EMPTY || old == key) return;
    s = (s + 1) & mask;
  }
}
__global__ void __launch_bounds__(256) k_merge_union(const unsigned long long* __restrict__ pairs, int64_t np, const unsigned long long* __restrict__ tk,
                                                   uint64_t mask, int* parent) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= np) return;
  const int a = key_slot(tk, mask, pairs[2 * t]), b = key_slot(tk, mask, pairs[2 * t + 1]);
  if (a >= 0 && b >= 0) uf_union(parent, a, b);
}
// representative of a merged component = its smallest key (independent of the slot layout, hence of the rank)
__global__ void __launch_bounds__(256) k_merge_minkey(const unsigned long long* __restrict__ tk, int64_t cap, int* parent,
                                                    unsigned long long* __restrict__ minkey) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap || tk[s] == HASH_EMPTY) return;
  atomicMin(&minkey[uf_find(parent, (int)s)], tk[s]);
}
// per local voxel: representative key of its global component (own root key when the component is not shared)
__global__ void __launch_bounds__(256) k_merge_lookup(const int* __restrict__ root, const unsigned long long* __restrict__ plain, int64_t nu,
                                                    const unsigned long long* __restrict__ tk, uint64_t mask, int* parent,
                                                    const unsigned long long* __restrict__ minkey, unsigned long long* __restrict__ rep,
                                                    uint8_t* __restrict__ is_cross) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const int r = root[u];
  const unsigned long long kr = plain[r];
  const int s = mask ? key_slot(tk, mask, kr) : -1;
  if (s >= 0) { rep[u] = minkey[uf_find(parent, s)]; is_cross[u] = 1; }
  else { rep[u] = kr; is_cross[u] = 0; }
}
// cluster size / smallest point index over the OWNED voxels of a tile (k_cluster_stats with ownership and global indices)
__global__ void __launch_bounds__(256) k_tile_stats(const int* __restrict__ root, const uint8_t* __restrict__ own, const uint32_t* __restrict__ ustart,
                                                  const uint32_t* __restrict__ perm, const float* __restrict__ xyz4, int64_t nu,
                                                  uint32_t* __restrict__ csize, uint32_t* __restrict__ cminpt) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu || !(own[u] & 1)) return;
  const int r = root[u];
  atomicAdd(&csize[r], 1u);
  atomicMin(&cminpt[r], __float_as_uint(xyz4[(int64_t)perm[ustart[u]] * 4 + 3]));   // the voxel's points ascend by global index
}
struct StatRec { unsigned long long rep; uint32_t size, minpt; };
// partial (size, minpt) of the cross-rank components, one record per local root that has owned members
__global__ void __launch_bounds__(256) k_stats_export(const int* __restrict__ root, const uint8_t* __restrict__ is_cross,
                                                    const unsigned long long* __restrict__ rep, const uint32_t* __restrict__ csize,
                                                    const uint32_t* __restrict__ cminpt, int64_t nu, StatRec* __restrict__ out,
                                                    unsigned long long* __restrict__ n_out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu || root[u] != (int)u || !is_cross[u] || csize[u] == 0) return;
  const unsigned long long s = atomicAdd(n_out, 1ull);
  out[s].rep = rep[u]; out[s].size = csize[u]; out[s].minpt = cminpt[u];
}
// totals per representative: tot[slot of rep key] accumulated from every rank's records
__global__ void __launch_bounds__(256) k_stats_accumulate(const StatRec* __restrict__ in, int64_t n, const unsigned long long* __restrict__ tk,
                                                        uint64_t mask, uint32_t* __restrict__ tsize, uint32_t* __restrict__ tminpt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = key_slot(tk, mask, in[t].rep);
  if (s < 0) return;
  atomicAdd(&tsize[s], in[t].size);
  atomicMin(&tminpt[s], in[t].minpt);
}
// (global index, canonical label) of the points of the owned voxels, compacted in sorted-position order
__global__ void __launch_bounds__(256) k_tile_labels(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ pos_unit, const int* __restrict__ root,
                                                   const uint8_t* __restrict__ own, const uint8_t* __restrict__ is_cross,
                                                   const unsigned long long* __restrict__ rep, const uint32_t* __restrict__ csize,
                                                   const uint32_t* __restrict__ cminpt, const unsigned long long* __restrict__ tk, uint64_t mask,
                                                   const uint32_t* __restrict__ tsize, const uint32_t* __restrict__ tminpt,
                                                   const float* __restrict__ xyz4, int64_t n_valid, int min_size_excl, uint2* __restrict__ out,
                                                   unsigned long long* __restrict__ n_out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  uint2 o = make_uint2(0, 0);
  if (p < n_valid) {
    const uint32_t u = pos_unit[p];
    if (own[u] & 1) {
      ok = true;
      const int r = root[u];
      uint32_t size, minpt;
      if (is_cross[u]) {
        const int s = key_slot(tk, mask, rep[u]);
        size = tsize[s]; minpt = tminpt[s];
      } else { size = csize[r]; minpt = cminpt[r]; }
      o.x = __float_as_uint(xyz4[(int64_t)perm[p] * 4 + 3]);
      o.y = (int)size > min_size_excl ? minpt : 0xffffffffu;
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  unsigned long long base = 0;
  const int lane = threadIdx.x & 31;
  if (lane == 0 && bal) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (ok) out[base + __popc(bal & ((1u << lane) - 1u))] = o;
}
// out[gidx - gfirst] = label for the records that came home
__global__ void __launch_bounds__(256) k_labels_scatter(const uint2* __restrict__ in, int64_t n, long long gfirst, int64_t n_out, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long i = (long long)in[t].x - gfirst;
  if (i >= 0 && i < n_out) out[i] = (int32_t)in[t].y;
}
// destination rank of a label record = the rank whose index slice holds the point: cnt[d * nwarps + warp]
struct SliceStarts { int nranks; long long first[TILE_MAX_RANKS + 1]; };
__device__ __forceinline__ int slice_of(long long g, const SliceStarts& ss) {
  int d = 0;
  while (d + 1 < ss.nranks && g >= ss.first[d + 1]) d++;
  return d;
}
__global__ void __launch_bounds__(256) k_home_count(const uint2* __restrict__ recs, int64_t n, SliceStarts ss, int64_t nwarps, uint32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= nwarps) return;
  uint32_t c[TILE_MAX_RANKS];
#pragma unroll
  for (int d = 0; d < TILE_MAX_RANKS; d++) c[d] = 0;
  for (int r = 0; r < ROUTE_PER_WARP / 32; r++) {
    const int64_t i = wg * ROUTE_PER_WARP + r * 32 + lane;
    const int dst = i < n ? slice_of((long long)recs[i].x, ss) : -1;
#pragma unroll
    for (int d = 0; d < TILE_MAX_RANKS; d++)
      if (d < ss.nranks) c[d] += __popc(__ballot_sync(0xffffffffu, dst == d));
  }
  if (lane == 0)
    for (int d = 0; d < ss.nranks; d++) cnt[(int64_t)d * nwarps + wg] = c[d];
}
__global__ void __launch_bounds__(256) k_home_scatter(const uint2* __restrict__ recs, int64_t n, SliceStarts ss, int64_t nwarps,
                                                    const uint32_t* __restrict__ offs, uint2* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= nwarps) return;
  uint32_t base[TILE_MAX_RANKS];
#pragma unroll
  for (int d = 0; d < TILE_MAX_RANKS; d++) base[d] = d < ss.nranks ? offs[(int64_t)d * nwarps + wg] : 0u;
  const uint32_t lt = (1u << lane) - 1u;
  for (int r = 0; r < ROUTE_PER_WARP / 32; r++) {
    const int64_t i = wg * ROUTE_PER_WARP + r * 32 + lane;
    uint2 v = make_uint2(0, 0);
    int dst = -1;
    if (i < n) { v = recs[i]; dst = slice_of((long long)v.x, ss); }
#pragma unroll
    for (int d = 0; d < TILE_MAX_RANKS; d++) {
      if (d < ss.nranks) {
        const uint32_t bal = __ballot_sync(0xffffffffu, dst == d);
        if (dst == d) out[base[d] + __popc(bal & lt)] = v;
        base[d] += __popc(bal);
      }
    }
  }
}

}  // namespace vgs

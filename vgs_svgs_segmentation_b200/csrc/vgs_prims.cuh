// vgs_prims.cuh — device-wide building blocks written for this library (no CUB/Thrust):
// exclusive scan (u32), stable LSD radix sort of (u32 / u64 key, u32 value), 64-bit open-addressing hash.
// All HBM-bound; grids are sized from the element count, tiles are 2-4 K elements per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vgs_math.cuh"

namespace vgs {

constexpr int SC_THREADS = 256, SC_IPT = 8, SC_TILE = SC_THREADS * SC_IPT;
constexpr int RS_THREADS = 256, RS_IPT = 16, RS_TILE = RS_THREADS * RS_IPT, RS_WARPS = RS_THREADS / 32;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024); returns exclusive prefix,
// total in *total (valid for all threads)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* smem33) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) smem33[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < nw ? smem33[lane] : 0;
    uint32_t si = warp_incl_scan(s, lane);
    smem33[lane] = si - s;
    if (lane == 31) smem33[32] = si;
  }
  __syncthreads();
  uint32_t r = inc - v + smem33[w];
  *total = smem33[32];
  __syncthreads();
  return r;
}

// ---- scan: reduce per tile -> scan tile sums (one CTA) -> downsweep ----
__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t sm[33];
  int64_t base = (int64_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    int64_t i = base + (int64_t)k * SC_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t tot;
  block_excl_scan(s, &tot, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// in-place exclusive scan of tile_sums[0..nt), total (u64) to *total64 and tile_sums[nt]
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* __restrict__ tile_sums, int64_t nt, unsigned long long* __restrict__ total64) {
  __shared__ uint32_t sm[33];
  unsigned long long run = 0;
  for (int64_t b = 0; b < nt; b += 1024) {
    int64_t i = b + threadIdx.x;
    uint32_t v = i < nt ? tile_sums[i] : 0;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, &tot, sm);
    if (i < nt) tile_sums[i] = (uint32_t)(run + ex);
    run += tot;
  }
  if (threadIdx.x == 0) { tile_sums[nt] = (uint32_t)run; *total64 = run; }
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_down(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                                                        const uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t sm[33];
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_IPT;  // blocked arrangement
  uint32_t v[SC_IPT];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, &tot, sm) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

// ---- radix sort pass: per-tile digit histogram -> (scan, digit-major) -> stable scatter ----
template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K* __restrict__ keys, int64_t n, int shift,
                                                       uint32_t* __restrict__ hist, int64_t nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; k++) {
    int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// (launch bound 4 CTAs per SM: without it the unrolled ranking takes 127 registers, 2 CTAs per SM, and the pass is
//  latency-bound at 21 % active warps — ncu, profiles/r02)
template <class K>
__global__ void __launch_bounds__(RS_THREADS, 4) k_rs_scatter(const K* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                          K* __restrict__ kout, uint32_t* __restrict__ vout, int64_t n,
                                                          int shift, const uint32_t* __restrict__ offs, int64_t nblk) {
  __shared__ uint32_t wc[RS_WARPS][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
  __syncthreads();
  // warp w owns the contiguous chunk [w*512, (w+1)*512) of the tile, 16 rounds of 32 (order kept)
  const int64_t cbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (RS_TILE / RS_WARPS);
  K k[RS_IPT];
  uint32_t v[RS_IPT];
  const uint32_t lt = (1u << lane) - 1u;
  // all loads first (32 independent requests in flight per thread), then the ranking: the warp syncs of the ranking loop
  // are compiler barriers, loads left inside it would be issued one round trip at a time
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    int64_t i = cbase + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? kin[i] : 0;
    v[r] = ok ? vin[i] : 0;
  }
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    int64_t i = cbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t d = ok ? ((uint32_t)(k[r] >> shift) & 255u) : 0xffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & lt) == 0) wc[w][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    int d = threadIdx.x;  // RS_THREADS == 256 digits
    uint32_t run = offs[(int64_t)d * nblk + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) { uint32_t t = wc[ww][d]; wc[ww][d] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    int64_t i = cbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t d = ok ? ((uint32_t)(k[r] >> shift) & 255u) : 0xffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t pos = 0;
    if (ok) pos = wc[w][d] + __popc(peers & lt);
    __syncwarp();
    if (ok && (peers & lt) == 0) wc[w][d] += __popc(peers);
    __syncwarp();
    if (ok) { kout[pos] = k[r]; vout[pos] = v[r]; }
  }
}

// ---- 64-bit key -> u32 value hash table (open addressing, linear probing, L2-resident) ----
constexpr uint64_t HASH_EMPTY = 0xffffffffffffffffull;

__global__ void k_hash_insert(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ tk,
                              uint32_t* __restrict__ tv, uint64_t mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = keys[i];
  uint64_t s = hash64(key) & mask;
  while (true) {
    unsigned long long old = atomicCAS(&tk[s], (unsigned long long)HASH_EMPTY, (unsigned long long)key);
    if (old == HASH_EMPTY || old == key) { tv[s] = (uint32_t)i; return; }
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ int hash_lookup(const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                           uint64_t mask, uint64_t key) {
  uint64_t s = hash64(key) & mask;
  while (true) {
    unsigned long long k = __ldg(&tk[s]);
    if (k == key) return (int)__ldg(&tv[s]);
    if (k == HASH_EMPTY) return -1;
    s = (s + 1) & mask;
  }
}

}  // namespace vgs

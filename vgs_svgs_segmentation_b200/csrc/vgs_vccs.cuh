// vgs_vccs.cuh — supervoxel generator for SVGS: what the reference obtains from
// pcl::SupervoxelClustering in createSupervoxels (supervoxel_segmentation.h:245-284; VCCS, Papon et al. 2013),
// as data-parallel kernels.  The algorithm and every float operation are those of oracle/vccs_oracle.cpp with
// schedule 1 (synchronous expansion rounds, order-independent fixed-point centroid sums); the tests require the
// two to agree bit for bit, and compare segmentation quality with the oracle's sequential PCL schedule.
// Third-party algorithm: parity with PCL itself is unpinned.
#pragma once
#include "vgs_prims.cuh"
#include "vgs_rows.cuh"

namespace vgs {

constexpr float VCCS_FMAX = 3.402823466e+38f;
// key of the generator's own voxel hash table: the three lattice coordinates side by side (21 bits each) — the lookups of
// the neighbour table, the seed filter and the reseeding probe hundreds of cells per warp, and a Morton interleave per
// probe was most of their instructions
__device__ __forceinline__ uint64_t vccs_key(uint32_t x, uint32_t y, uint32_t z) { return ((uint64_t)x << 42) | ((uint64_t)y << 21) | (uint64_t)z; }

// Occupancy bits of the voxel lattice over the occupied key range + a margin (BitGrid of vgs_rows.cuh; bits == nullptr when
// the range is too large for one).  On a surface 70-85 % of the cells a cube search visits are empty: one bit test (lanes
// that walk along z share a word) answers those without the hash probe.  A cell outside the grid holds no voxel.
struct VccsBits { BitGrid g; uint32_t nx; const uint32_t* bits; };
__device__ __forceinline__ bool vccs_maybe_occupied(const VccsBits& b, long long x, long long y, long long z) {
  if (!b.bits) return true;
  const long long rx = x - b.g.x0, ry = y - b.g.y0, rz = z - b.g.z0;
  if (rx < 0 || ry < 0 || rz < 0 || rx >= (long long)b.nx || ry >= (long long)b.g.ny || rz >= (long long)b.g.nz) return false;
  const uint64_t bit = ((uint64_t)rx * b.g.ny + (uint64_t)ry) * b.g.nz + (uint64_t)rz;
  return (__ldg(b.bits + (bit >> 5)) >> (bit & 31)) & 1u;
}
__global__ void __launch_bounds__(256) k_vccs_bits_set(const uint32_t* __restrict__ key3, int64_t V, BitGrid g, uint32_t* __restrict__ bits) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const uint64_t b = bg_bit(g, (int)key3[3 * v], (int)key3[3 * v + 1], (int)key3[3 * v + 2]);
  atomicOr(&bits[b >> 5], 1u << (b & 31));
}

// ---- computeVoxelData: voxel = float mean of its points (ascending point index), lattice key, point -> voxel ----
__global__ void __launch_bounds__(256) k_vccs_voxels(const float* __restrict__ xyz, int stride, const uint32_t* __restrict__ perm,
                                                   const uint32_t* __restrict__ vstart, const uint64_t* __restrict__ vkey, int64_t V,
                                                   int depth, int descending, float* __restrict__ vxyz, uint32_t* __restrict__ key3,
                                                   uint64_t* __restrict__ plain, int32_t* __restrict__ pt_voxel) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  uint64_t m = vkey[v];
  const uint64_t mask = (1ull << (3 * depth)) - 1ull;
  if (descending) m = ~m & mask;
  uint32_t kx, ky, kz;
  morton_decode(m, kx, ky, kz);
  key3[3 * v] = kx; key3[3 * v + 1] = ky; key3[3 * v + 2] = kz;
  plain[v] = vccs_key(kx, ky, kz);
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const uint32_t b = vstart[v], e = vstart[v + 1];
  for (uint32_t j = b; j < e; j++) {
    const uint32_t p = perm[j];
    const float* q = xyz + (int64_t)p * stride;
    sx += q[0]; sy += q[1]; sz += q[2];
    pt_voxel[p] = (int32_t)v;
  }
  const float c = (float)(e - b);
  vxyz[3 * v] = sx / c; vxyz[3 * v + 1] = sy / c; vxyz[3 * v + 2] = sz / c;
}

// ---- OctreePointCloudAdjacency::computeNeighbors: the 27 lattice cells around a voxel, self included.
//      Table layout nb[v * 27 + j] (measured against the slot-major layout nb[j * V + v]: expand 95 vs 112 us, plane-fit
//      moments 131 vs 184 us — a thread walks its own 108 bytes through L1, 27 far-apart streams per warp cost more);
//      consumed once by k_vccs_nb_count / k_vccs_nb_compact ----
__global__ void __launch_bounds__(256) k_vccs_neighbours(const uint32_t* __restrict__ key3, int64_t V, int depth,
                                                       const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv,
                                                       uint64_t mask, VccsBits ob, int32_t* __restrict__ nb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * 27) return;
  const int64_t v = i / 27;
  const int j = (int)(i - v * 27);
  const int64_t lim = 1ll << depth;
  const int64_t x = (int64_t)key3[3 * v] + (j / 9 - 1), y = (int64_t)key3[3 * v + 1] + ((j / 3) % 3 - 1), z = (int64_t)key3[3 * v + 2] + (j % 3 - 1);
  int id = -1;
  if (x >= 0 && y >= 0 && z >= 0 && x < lim && y < lim && z < lim && vccs_maybe_occupied(ob, x, y, z))
    id = hash_lookup(tk, tv, mask, vccs_key((uint32_t)x, (uint32_t)y, (uint32_t)z));
  nb[i] = id;
}

// The rounds below only ever want the EXISTING neighbours (about 10 of the 27 slots on a surface): the table is compacted once
// into a CSR (slot order kept), which cuts the bytes every expansion round and plane fit streams by ~2.7x.
__global__ void __launch_bounds__(256) k_vccs_nb_count(const int32_t* __restrict__ nb, int64_t V, uint32_t* __restrict__ cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > V) return;
  int c = 0;
  if (v < V)
    for (int j = 0; j < 27; j++) c += nb[v * 27 + j] >= 0 ? 1 : 0;
  cnt[v] = (uint32_t)c;          // cnt[V] = 0: the scan then yields the end of the last list
}
__global__ void __launch_bounds__(256) k_vccs_nb_compact(const int32_t* __restrict__ nb, int64_t V, const uint32_t* __restrict__ off,
                                                       int32_t* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  uint32_t p = off[v];
  for (int j = 0; j < 27; j++) {
    const int u = nb[v * 27 + j];
    if (u >= 0) out[p++] = u;
  }
}

__device__ __forceinline__ float vccs_dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }

// ---- normals: plane fit (pcl::computePointNormal) over the multiset { [v] } + for u in N(v): u, N(u);
//      filter -2: initial fit (self counted once more, everybody counts); owner != null: only voxels of v's supervoxel
//      (SupervoxelHelper::refineNormals).
//      The sums are ORDER INDEPENDENT: voxel centroids in 2^-20 fixed point, moments in 64-bit integers (exact).  That makes
//      the two-ring walk factorise: k_vccs_moments gives every voxel u the moments of its own 27-neighbourhood relative to
//      itself (27 gathers), k_vccs_normals sums the records of the 27 neighbours translated by q_u - q_v (27 gathers of
//      80 bytes) — instead of the 27 x 27 = 729 gathers per voxel of the literal walk (0.73 ms -> 0.1 ms per fit at 1.5 M
//      voxels).  oracle/vccs_oracle.cpp (fit_moments / fit_normal_fixed) is the same arithmetic. ----
struct __align__(16) VMom { long long n, s[3], m[6]; };
__device__ __forceinline__ void vccs_q3(const float* __restrict__ vxyz, int64_t w, long long q[3]) {
  q[0] = __double2ll_rn((double)vxyz[3 * w] * 1048576.0);
  q[1] = __double2ll_rn((double)vxyz[3 * w + 1] * 1048576.0);
  q[2] = __double2ll_rn((double)vxyz[3 * w + 2] * 1048576.0);
}
__global__ void __launch_bounds__(128) k_vccs_moments(int64_t V, const float* __restrict__ vxyz, const uint32_t* __restrict__ nb_off,
                                                    const int32_t* __restrict__ nb, const int32_t* __restrict__ owner, VMom* __restrict__ mom) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= V) return;
  int f = -2;
  if (owner) { f = owner[u]; if (f < 0) return; }        // the record of an unowned voxel is never read
  long long qu[3];
  vccs_q3(vxyz, u, qu);
  VMom A;
  A.n = 0; A.s[0] = A.s[1] = A.s[2] = 0;
  for (int i = 0; i < 6; i++) A.m[i] = 0;
  for (uint32_t k = nb_off[u], ke = nb_off[u + 1]; k < ke; k++) {
    const int w = nb[k];
    if (f >= 0 && owner[w] != f) continue;
    long long qw[3];
    vccs_q3(vxyz, w, qw);
    const long long x = qw[0] - qu[0], y = qw[1] - qu[1], z = qw[2] - qu[2];
    A.n++; A.s[0] += x; A.s[1] += y; A.s[2] += z;
    A.m[0] += x * x; A.m[1] += x * y; A.m[2] += x * z; A.m[3] += y * y; A.m[4] += y * z; A.m[5] += z * z;
  }
  longlong2* out = reinterpret_cast<longlong2*>(mom + u);
  out[0] = make_longlong2(A.n, A.s[0]); out[1] = make_longlong2(A.s[1], A.s[2]); out[2] = make_longlong2(A.m[0], A.m[1]);
  out[3] = make_longlong2(A.m[2], A.m[3]); out[4] = make_longlong2(A.m[4], A.m[5]);
}
__global__ void __launch_bounds__(128) k_vccs_normals(int64_t V, const float* __restrict__ vxyz, const uint32_t* __restrict__ nb_off,
                                                    const int32_t* __restrict__ nb, const int32_t* __restrict__ owner, const VMom* __restrict__ mom,
                                                    float* __restrict__ nrm) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int filter = -2;
  if (owner) { filter = owner[v]; if (filter < 0) return; }
  const float Kx = vxyz[3 * v], Ky = vxyz[3 * v + 1], Kz = vxyz[3 * v + 2];
  long long qv[3];
  vccs_q3(vxyz, v, qv);
  long long n = filter == -2 ? 1 : 0, S0 = 0, S1 = 0, S2 = 0, M0 = 0, M1 = 0, M2 = 0, M3 = 0, M4 = 0, M5 = 0;
  for (uint32_t j = nb_off[v], je = nb_off[v + 1]; j < je; j++) {
    const int u = nb[j];
    if (filter >= 0 && owner[u] != filter) continue;
    long long qu[3];
    vccs_q3(vxyz, u, qu);
    const long long tx = qu[0] - qv[0], ty = qu[1] - qv[1], tz = qu[2] - qv[2];
    const longlong2* in = reinterpret_cast<const longlong2*>(mom + u);
    const longlong2 r0 = __ldg(in), r1 = __ldg(in + 1), r2 = __ldg(in + 2), r3 = __ldg(in + 3), r4 = __ldg(in + 4);
    const long long c = r0.x + 1;                 // u itself + its neighbourhood
    const long long sx = r0.y, sy = r1.x, sz = r1.y;
    n += c;
    S0 += sx + c * tx; S1 += sy + c * ty; S2 += sz + c * tz;
    M0 += r2.x + 2 * sx * tx + c * tx * tx;
    M1 += r2.y + sx * ty + tx * sy + c * tx * ty;
    M2 += r3.x + sx * tz + tx * sz + c * tx * tz;
    M3 += r3.y + 2 * sy * ty + c * ty * ty;
    M4 += r4.x + sy * tz + ty * sz + c * ty * tz;
    M5 += r4.y + 2 * sz * tz + c * tz * tz;
  }
  float nx, ny, nz;
  if (n < 3) {
    nx = ny = nz = __int_as_float(0x7fc00000);
  } else {
    const double cd = (double)n, sc = 9.094947017729282379150390625e-13;   // 2^-40: fixed point^2 -> m^2
    const double mx = (double)S0 / cd, my = (double)S1 / cd, mz = (double)S2 / cd;
    const Sym3 cov{(float)(((double)M0 / cd - mx * mx) * sc), (float)(((double)M1 / cd - mx * my) * sc), (float)(((double)M2 / cd - mx * mz) * sc),
                   (float)(((double)M3 / cd - my * my) * sc), (float)(((double)M4 / cd - my * mz) * sc), (float)(((double)M5 / cd - mz * mz) * sc)};
    // pcl::eigen33(mat, eigenvalue, eigenvector): eigenvector of the smallest eigenvalue
    float scale = fmaxf(fmaxf(fmaxf(fabsf(cov.a00), fabsf(cov.a01)), fmaxf(fabsf(cov.a02), fabsf(cov.a11))),
                        fmaxf(fabsf(cov.a12), fabsf(cov.a22)));
    if (scale <= FLT_MIN) scale = 1.0f;
    const Sym3 a{cov.a00 / scale, cov.a01 / scale, cov.a02 / scale, cov.a11 / scale, cov.a12 / scale, cov.a22 / scale};
    float ev[3];
    roots3(a, ev);
    F3 v1, v2, v3, nn; float l1, l2, l3;
    row_crosses(a, ev[0], v1, v2, v3, l1, l2, l3);
    pick_cross(v1, v2, v3, l1, l2, l3, nn);
    // flipNormalTowardsViewpoint(point, 0, 0, 0, normal); normal[3] = 0; normalize
    if (vccs_dot3(0.0f - Kx, 0.0f - Ky, 0.0f - Kz, nn.x, nn.y, nn.z) < 0) { nn.x *= -1; nn.y *= -1; nn.z *= -1; }
    const float z = (nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z;
    if (z > 0.f) { const float sq = sqrtf(z); nn.x /= sq; nn.y /= sq; nn.z /= sq; }
    nx = nn.x; ny = nn.y; nz = nn.z;
  }
  nrm[3 * v] = nx; nrm[3 * v + 1] = ny; nrm[3 * v + 2] = nz;
}

// ---- selectInitialSupervoxelSeeds ----
// seed-grid cell of every voxel centroid (grid anchored at the octree origin); key = x-major morton of the cell
__global__ void __launch_bounds__(256) k_vccs_cell_keys(const float* __restrict__ vxyz, int64_t V, double ox, double oy, double oz, double seed,
                                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ cell3) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const long long cx = (long long)floor(((double)vxyz[3 * v] - ox) / seed), cy = (long long)floor(((double)vxyz[3 * v + 1] - oy) / seed),
                  cz = (long long)floor(((double)vxyz[3 * v + 2] - oz) / seed);
  cell3[3 * v] = (int32_t)cx; cell3[3 * v + 1] = (int32_t)cy; cell3[3 * v + 2] = (int32_t)cz;
  keys[v] = morton_encode((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
  vals[v] = (uint32_t)v;
}

// voxel nearest to the centre of every occupied cell: each voxel offers itself to the 27 cells around its own
__global__ void __launch_bounds__(256) k_vccs_seed_nearest(const float* __restrict__ vxyz, const int32_t* __restrict__ cell3, int64_t V,
                                                         double ox, double oy, double oz, double seed,
                                                         const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv, uint64_t mask,
                                                         unsigned long long* __restrict__ best) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float px = vxyz[3 * v], py = vxyz[3 * v + 1], pz = vxyz[3 * v + 2];
  for (int j = 0; j < 27; j++) {
    const long long cx = (long long)cell3[3 * v] + (j / 9 - 1), cy = (long long)cell3[3 * v + 1] + ((j / 3) % 3 - 1), cz = (long long)cell3[3 * v + 2] + (j % 3 - 1);
    if (cx < 0 || cy < 0 || cz < 0) continue;
    const int c = hash_lookup(tk, tv, mask, morton_encode((uint32_t)cx, (uint32_t)cy, (uint32_t)cz));
    if (c < 0) continue;
    const float mx = (float)(((double)cx + 0.5) * seed + ox), my = (float)(((double)cy + 0.5) * seed + oy), mz = (float)(((double)cz + 0.5) * seed + oz);
    const float dx = px - mx, dy = py - my, dz = pz - mz;
    const float d2 = dx * dx + (dy * dy + dz * dz);
    atomicMin(&best[c], ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(uint32_t)v);
  }
}

// several cells may share their nearest voxel: the smallest cell index keeps it
__global__ void __launch_bounds__(256) k_vccs_seed_claim(const unsigned long long* __restrict__ best, int64_t NC, int32_t* __restrict__ claim) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  const unsigned long long b = best[c];
  if (b == ~0ull) return;
  atomicMin(&claim[(uint32_t)b], (int32_t)c);
}

// kept when more than min_points voxel centroids lie within search_radius of the seed voxel (FLANN: dist2 < r2).
// One warp per cell.
__global__ void __launch_bounds__(128) k_vccs_seed_filter(const unsigned long long* __restrict__ best, const int32_t* __restrict__ claim, int64_t NC,
                                                        const uint32_t* __restrict__ key3, const float* __restrict__ vxyz, int depth,
                                                        const unsigned long long* __restrict__ tk, const uint32_t* __restrict__ tv, uint64_t mask,
                                                        VccsBits ob, float r2, float min_points, int reach, uint32_t* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= NC) return;
  const unsigned long long b = best[c];
  bool keep = false;
  if (b != ~0ull && claim[(uint32_t)b] == (int32_t)c) {
    const int64_t s = (int64_t)(uint32_t)b;
    const float sx = vxyz[3 * s], sy = vxyz[3 * s + 1], sz = vxyz[3 * s + 2];
    const int side = 2 * reach + 1;
    const int64_t lim = 1ll << depth;
    int num = 0;
    for (int t = lane; t < side * side * side; t += 32) {
      const int64_t x = (int64_t)key3[3 * s] + (t / (side * side) - reach), y = (int64_t)key3[3 * s + 1] + ((t / side) % side - reach),
                    z = (int64_t)key3[3 * s + 2] + (t % side - reach);
      if (x < 0 || y < 0 || z < 0 || x >= lim || y >= lim || z >= lim || !vccs_maybe_occupied(ob, x, y, z)) continue;
      const int w = hash_lookup(tk, tv, mask, vccs_key((uint32_t)x, (uint32_t)y, (uint32_t)z));
      if (w < 0) continue;
      const float dx = vxyz[3 * (int64_t)w] - sx, dy = vxyz[3 * (int64_t)w + 1] - sy, dz = vxyz[3 * (int64_t)w + 2] - sz;
      if (dx * dx + (dy * dy + dz * dz) < r2) num++;
    }
    num = __reduce_add_sync(0xffffffffu, num);
    keep = (float)num > min_points;
  }
  if (lane == 0) flag[c] = keep ? 1u : 0u;
}

// createSupervoxelHelpers: supervoxel h (label h + 1) starts as its seed voxel
__global__ void __launch_bounds__(256) k_vccs_helpers(const unsigned long long* __restrict__ best, const uint32_t* __restrict__ flag,
                                                    const uint32_t* __restrict__ rank, int64_t NC, const float* __restrict__ vxyz,
                                                    const float* __restrict__ nrm, float* __restrict__ hc, float* __restrict__ hn,
                                                    uint8_t* __restrict__ alive, int32_t* __restrict__ owner) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC || !flag[c]) return;
  const int64_t h = rank[c], s = (int64_t)(uint32_t)best[c];
  for (int a = 0; a < 3; a++) { hc[3 * h + a] = vxyz[3 * s + a]; hn[3 * h + a] = nrm[3 * s + a]; }
  alive[h] = 1;
  owner[s] = (int32_t)h;
}

// ---- expandSupervoxels, one synchronous round: a voxel goes to the supervoxel (among the owners of its 26
//      neighbours) whose centroid is nearest in voxelDataDistance, if that beats its best distance so far;
//      ties to the smaller label (the sequential reference lets the first, i.e. smaller, label win) ----
__global__ void __launch_bounds__(256) k_vccs_expand(int64_t V, const uint32_t* __restrict__ nb_off, const int32_t* __restrict__ nb, const int32_t* __restrict__ owner_old,
                                                   int32_t* __restrict__ owner_new, float* __restrict__ dist, const float* __restrict__ vxyz,
                                                   const float* __restrict__ nrm, const float* __restrict__ hc, const float* __restrict__ hn,
                                                   const uint8_t* __restrict__ alive, float seed_res, float wc, float ws, float wn,
                                                   unsigned long long* __restrict__ acc, unsigned long long* __restrict__ cnt) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int own = owner_old[v];
  float bd = dist[v];
  int bh = own;
  const float px = vxyz[3 * v], py = vxyz[3 * v + 1], pz = vxyz[3 * v + 2];
  const float qx = nrm[3 * v], qy = nrm[3 * v + 1], qz = nrm[3 * v + 2];
  // (measured and rejected: fetching the 27 neighbour ids and then their 27 owners into registers first — 104 vs 95 us)
  int last = -1;
  for (uint32_t j = nb_off[v], je = nb_off[v + 1]; j < je; j++) {
    const int u = nb[j];
    const int h = owner_old[u];
    if (h < 0 || h == own || h == last || !alive[h]) continue;   // h == last: same candidate again, same distance
    last = h;
    const float dx = hc[3 * (int64_t)h] - px, dy = hc[3 * (int64_t)h + 1] - py, dz = hc[3 * (int64_t)h + 2] - pz;
    const float spatial = sqrtf(dx * dx + (dy * dy + dz * dz)) / seed_res;
    const float color = 0.0f;
    const float cosn = 1.0f - fabsf(vccs_dot3(hn[3 * (int64_t)h], hn[3 * (int64_t)h + 1], hn[3 * (int64_t)h + 2], qx, qy, qz));
    const float d = cosn * wn + color * wc + spatial * ws;
    if (d < bd || (d == bd && bh != own && h < bh)) { bd = d; bh = h; }
  }
  owner_new[v] = bh;
  dist[v] = bd;
  // SupervoxelHelper::updateCentroid: the fixed-point sums are exact integers, so a voxel that changes hands is simply taken
  // out of its old supervoxel's sums and put into the new one's (k_vccs_accumulate builds them once per expansion phase)
  if (bh != own) {
    const unsigned long long x = (unsigned long long)__double2ll_rn((double)px * 1048576.0), y = (unsigned long long)__double2ll_rn((double)py * 1048576.0),
                             z = (unsigned long long)__double2ll_rn((double)pz * 1048576.0);
    const bool hasn = qx == qx;           // NaN normals (isolated voxels) contribute nothing
    unsigned long long nx = 0, ny = 0, nz = 0;
    if (hasn) {
      nx = (unsigned long long)__double2ll_rn((double)qx * 1073741824.0); ny = (unsigned long long)__double2ll_rn((double)qy * 1073741824.0);
      nz = (unsigned long long)__double2ll_rn((double)qz * 1073741824.0);
    }
    if (own >= 0) {
      unsigned long long* a = acc + (int64_t)own * 6;
      atomicAdd(a + 0, 0ull - x); atomicAdd(a + 1, 0ull - y); atomicAdd(a + 2, 0ull - z);
      if (hasn) { atomicAdd(a + 3, 0ull - nx); atomicAdd(a + 4, 0ull - ny); atomicAdd(a + 5, 0ull - nz); }
      atomicAdd(cnt + own, ~0ull);        // - 1
    }
    unsigned long long* a = acc + (int64_t)bh * 6;       // bh >= 0: a voxel never goes back to "no owner"
    atomicAdd(a + 0, x); atomicAdd(a + 1, y); atomicAdd(a + 2, z);
    if (hasn) { atomicAdd(a + 3, nx); atomicAdd(a + 4, ny); atomicAdd(a + 5, nz); }
    atomicAdd(cnt + bh, 1ull);
  }
}

// ---- SupervoxelHelper::updateCentroid with order-independent sums: 2^-20 fixed point for xyz, 2^-30 for normals.
//      Runs ONCE per expansion phase (over the seeds); the rounds keep the sums current with the voxels that change hands.
//      Consecutive voxels (Morton order) mostly share their supervoxel: a thread walks VCCS_ACC_RUN consecutive voxels and
//      issues its seven atomics only when the owner changes.  (Measured and rejected: warp-level aggregation with
//      match_any / REDUX over run masks — 138 us per round against 56 us for one set of atomics per voxel.) ----
constexpr int VCCS_ACC_RUN = 8;
__global__ void __launch_bounds__(256) k_vccs_accumulate(int64_t V, const int32_t* __restrict__ owner, const float* __restrict__ vxyz,
                                                       const float* __restrict__ nrm, unsigned long long* __restrict__ acc,
                                                       unsigned long long* __restrict__ cnt) {
  const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VCCS_ACC_RUN;
  if (v0 >= V) return;
  int cur = -1;
  long long sx = 0, sy = 0, sz = 0, nx = 0, ny = 0, nz = 0;
  unsigned long long c = 0;
  auto flush = [&]() {
    if (cur < 0 || c == 0) return;
    unsigned long long* a = acc + (int64_t)cur * 6;
    atomicAdd(a + 0, (unsigned long long)sx); atomicAdd(a + 1, (unsigned long long)sy); atomicAdd(a + 2, (unsigned long long)sz);
    if (nx) atomicAdd(a + 3, (unsigned long long)nx);
    if (ny) atomicAdd(a + 4, (unsigned long long)ny);
    if (nz) atomicAdd(a + 5, (unsigned long long)nz);
    atomicAdd(cnt + cur, c);
  };
#pragma unroll
  for (int k = 0; k < VCCS_ACC_RUN; k++) {
    const int64_t v = v0 + k;
    if (v >= V) break;
    const int h = owner[v];
    if (h != cur) { flush(); cur = h; sx = sy = sz = nx = ny = nz = 0; c = 0; }
    if (h < 0) continue;
    sx += __double2ll_rn((double)vxyz[3 * v] * 1048576.0);
    sy += __double2ll_rn((double)vxyz[3 * v + 1] * 1048576.0);
    sz += __double2ll_rn((double)vxyz[3 * v + 2] * 1048576.0);
    const float fx = nrm[3 * v];
    if (fx == fx) {   // NaN normals (isolated voxels) contribute nothing
      nx += __double2ll_rn((double)fx * 1073741824.0);
      ny += __double2ll_rn((double)nrm[3 * v + 1] * 1073741824.0);
      nz += __double2ll_rn((double)nrm[3 * v + 2] * 1073741824.0);
    }
    c++;
  }
  flush();
}

__global__ void __launch_bounds__(256) k_vccs_centroids(int64_t H, const unsigned long long* __restrict__ acc, const unsigned long long* __restrict__ cnt,
                                                      float* __restrict__ hc, float* __restrict__ hn, uint8_t* __restrict__ alive) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H || !alive[h]) return;
  if (cnt[h] == 0) { alive[h] = 0; return; }
  const long long* a = reinterpret_cast<const long long*>(acc) + h * 6;
  const double c = (double)cnt[h];
  hc[3 * h] = (float)((double)a[0] / 1048576.0 / c);
  hc[3 * h + 1] = (float)((double)a[1] / 1048576.0 / c);
  hc[3 * h + 2] = (float)((double)a[2] / 1048576.0 / c);
  float nx = (float)((double)a[3] / 1073741824.0), ny = (float)((double)a[4] / 1073741824.0), nz = (float)((double)a[5] / 1073741824.0);
  const float z = (nx * nx + ny * ny) + nz * nz;
  if (z > 0.f) { const float s = sqrtf(z); nx /= s; ny /= s; nz /= s; }
  hn[3 * h] = nx; hn[3 * h + 1] = ny; hn[3 * h + 2] = nz;
}

// ---- reseedSupervoxels: the voxel nearest to the supervoxel's centroid (kd-tree nearestKSearch in PCL), found
//      exactly by growing lattice cubes: a hit closer than the cube's inscribed radius cannot be beaten from
//      outside.  One warp per supervoxel. ----
__global__ void __launch_bounds__(128) k_vccs_reseed(int64_t H, const float* __restrict__ hc, const uint8_t* __restrict__ alive, double ox, double oy,
                                                   double oz, double res, int depth, const unsigned long long* __restrict__ tk,
                                                   const uint32_t* __restrict__ tv, uint64_t mask, VccsBits ob, const float* __restrict__ vxyz,
                                                   int32_t* __restrict__ seedv) {
  const int lane = threadIdx.x & 31;
  int64_t h = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (h >= H) return;
  int found = -1;
  if (alive[h]) {
    const float cx = hc[3 * h], cy = hc[3 * h + 1], cz = hc[3 * h + 2];
    const long long kx = (long long)floor(((double)cx - ox) / res), ky = (long long)floor(((double)cy - oy) / res), kz = (long long)floor(((double)cz - oz) / res);
    const long long lim = 1ll << depth;
    for (int R = 2; R <= 32 && found < 0; R *= 2) {
      const int side = 2 * R + 1;
      unsigned long long bestk = ~0ull;
      for (int t = lane; t < side * side * side; t += 32) {
        const long long x = kx + (t / (side * side) - R), y = ky + ((t / side) % side - R), z = kz + (t % side - R);
        if (x < 0 || y < 0 || z < 0 || x >= lim || y >= lim || z >= lim || !vccs_maybe_occupied(ob, x, y, z)) continue;
        const int v = hash_lookup(tk, tv, mask, vccs_key((uint32_t)x, (uint32_t)y, (uint32_t)z));
        if (v < 0) continue;
        const float dx = vxyz[3 * (int64_t)v] - cx, dy = vxyz[3 * (int64_t)v + 1] - cy, dz = vxyz[3 * (int64_t)v + 2] - cz;
        const float d2 = dx * dx + (dy * dy + dz * dz);
        const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(uint32_t)v;
        bestk = k < bestk ? k : bestk;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestk, o);
        bestk = other < bestk ? other : bestk;
      }
      if (bestk != ~0ull && sqrt((double)__uint_as_float((uint32_t)(bestk >> 32))) < (double)R * res * (1.0 - 1e-5)) found = (int)(uint32_t)bestk;
    }
  }
  if (lane == 0) seedv[h] = found;
}

__global__ void __launch_bounds__(256) k_vccs_reset(int64_t V, int32_t* __restrict__ owner, float* __restrict__ dist, int32_t* __restrict__ claim) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  owner[v] = -1; dist[v] = VCCS_FMAX; claim[v] = 0x7fffffff;
}
__global__ void __launch_bounds__(256) k_vccs_reseed_claim(int64_t H, const int32_t* __restrict__ seedv, const uint8_t* __restrict__ alive, int32_t* __restrict__ claim) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H || !alive[h] || seedv[h] < 0) return;
  atomicMin(&claim[seedv[h]], (int32_t)h);
}
// a supervoxel without a voxel, or whose voxel went to a smaller label, ends here
__global__ void __launch_bounds__(256) k_vccs_reseed_apply(int64_t H, const int32_t* __restrict__ seedv, uint8_t* __restrict__ alive,
                                                         const int32_t* __restrict__ claim, int32_t* __restrict__ owner) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H || !alive[h]) return;
  const int s = seedv[h];
  if (s < 0 || claim[s] != (int32_t)h) { alive[h] = 0; return; }
  owner[s] = (int32_t)h;
}

// ---- getLabeledCloud / getMaxLabel ----
__global__ void __launch_bounds__(256) k_vccs_point_labels(int64_t n, const int32_t* __restrict__ pt_voxel, const int32_t* __restrict__ owner,
                                                         int32_t* __restrict__ labels) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = pt_voxel[i];
  const int o = v >= 0 ? owner[v] : -1;
  labels[i] = o >= 0 ? o + 1 : 0;
}
__global__ void __launch_bounds__(256) k_vccs_max_label(int64_t H, const uint8_t* __restrict__ alive, int32_t* __restrict__ out) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H || !alive[h]) return;
  atomicMax(out, (int32_t)h + 1);
}

}  // namespace vgs

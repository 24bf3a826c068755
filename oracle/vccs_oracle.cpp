/*
 * vccs_oracle.cpp — CPU restatement of the supervoxel generator the reference calls in
 * createSupervoxels (supervoxel_segmentation.h:245-284): pcl::SupervoxelClustering ("VCCS", Papon et al.
 * CVPR 2013) with extract() + refineSupervoxels(5), colour/spatial/normal importances from the task file.
 *
 * TEST INFRASTRUCTURE ONLY (see vgs_oracle.h).  PARITY UNPINNED: VCCS is third-party code (PCL 1.8.1,
 * segmentation/impl/supervoxel_clustering.hpp) that is absent from /root/reference; this file restates its
 * published algorithm from the paper and the recalled PCL implementation:
 *   computeVoxelData        voxel = mean of its points; normal = plane fit over the voxel, its 26-neighbours and
 *                           their neighbours (a multiset: a voxel reached several times counts several times),
 *                           flipped towards the origin
 *   selectInitialSupervoxelSeeds   occupied cells of a seed_resolution grid -> voxel nearest to the cell centre,
 *                           kept when more than 0.05*pi*r^2/res^2 voxels lie within r = seed_resolution/2
 *   expandSupervoxels       depth = int(1.8*seed/res); depth-1 rounds: every supervoxel looks at the neighbours
 *                           of its voxels and takes those whose distance  wn*(1-|n.n'|) + wc*dc + ws*|dx|/seed
 *                           to ITS centroid is below the voxel's best distance so far; then centroids are updated
 *   refineSupervoxels(k)    k x { normals re-fitted inside each supervoxel; reseed at the voxel nearest to the
 *                           centroid; expand again }
 *   getLabeledCloud / getMaxLabel   per-point label of the owning supervoxel (0 = none), largest label
 * Deliberate deviations (none can be pinned against PCL here): the seed grid is anchored at the voxel octree's
 * origin (PCL grows a second dynamic bounding box over the voxel centroids); the covariance sums are shifted by
 * the voxel's own centroid (PCL 1.8.1 sums raw coordinates in float); colours are absent (the reference copies an
 * XYZ cloud into XYZRGBA, so every colour distance is 0).
 *
 * Two schedules:  schedule 0 = PCL's sequential order (supervoxels expand one after the other and see each
 * other's claims inside a round; float sums in voxel order);  schedule 1 = synchronous rounds (every voxel picks
 * the best claim of the round, ties to the smaller label; centroid sums in 2^-20 / 2^-30 fixed point and the plane-fit
 * moments in exact 2^-20 fixed-point integer arithmetic, hence order independent) — the schedule the CUDA generator
 * implements, and with which it must agree bit for bit.
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <set>
#include <unordered_map>
#include <vector>

#include "vgs_oracle.h"

namespace {

struct F3 { float x, y, z; };
inline float sqn3(F3 a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
inline F3 cross3(F3 a, F3 b) { return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float dot3(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// pcl::computeRoots2 / computeRoots, Scalar = float, libm calls correctly rounded (oracle math = 1)
inline void roots2(float b, float c, float* r) {
  r[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
inline void roots3(const float m[6], float* r) {   // m = a00 a01 a02 a11 a12 a22
  float c0 = m[0] * m[3] * m[5] + 2.0f * m[1] * m[2] * m[4] - m[0] * m[4] * m[4] - m[3] * m[2] * m[2] - m[5] * m[1] * m[1];
  float c1 = m[0] * m[3] - m[1] * m[1] + m[0] * m[5] - m[2] * m[2] + m[3] * m[5] - m[4] * m[4];
  float c2 = m[0] + m[3] + m[5];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) { roots2(c2, c1, r); return; }
  const float inv3 = (float)(1.0 / 3.0);
  const float sqrt3 = std::sqrt(3.0f);
  float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.f) a_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.f) q = 0.f;
  float rho = std::sqrt(-a_3);
  float theta = (float)std::atan2((double)std::sqrt(-q), (double)half_b) * inv3;
  float ct = (float)std::cos((double)theta), st = (float)std::sin((double)theta);
  r[0] = c2_3 + 2.0f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  float t;
  if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  if (r[1] >= r[2]) {
    t = r[1]; r[1] = r[2]; r[2] = t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  }
  if (r[0] <= 0.f) roots2(c2, c1, r);
}
// pcl::eigen33(mat, eigenvalue, eigenvector): eigenvector of the smallest eigenvalue
inline F3 smallest_eigenvector(const float cov[6]) {
  float scale = 0.f;
  for (int i = 0; i < 6; i++) scale = std::fmax(scale, std::fabs(cov[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float a[6];
  for (int i = 0; i < 6; i++) a[i] = cov[i] / scale;
  float ev[3];
  roots3(a, ev);
  F3 r0{a[0] - ev[0], a[1], a[2]}, r1{a[1], a[3] - ev[0], a[4]}, r2{a[2], a[4], a[5] - ev[0]};
  F3 v1 = cross3(r0, r1), v2 = cross3(r0, r2), v3 = cross3(r1, r2);
  float l1 = sqn3(v1), l2 = sqn3(v2), l3 = sqn3(v3);
  F3 v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  float s = std::sqrt(l);
  return F3{v.x / s, v.y / s, v.z / s};
}

struct Vccs {
  int64_t V = 0;
  const uint32_t* key = nullptr;
  std::vector<F3> xyz, nrm;
  std::vector<int32_t> nb;         // V x 27, (dx+1)*9 + (dy+1)*3 + (dz+1), self included (PCL computeNeighbors), -1 none
  std::vector<int32_t> owner;      // helper index, -1 none
  std::vector<float> dist;
  float res = 0, seed = 0, wc = 0, ws = 0, wn = 0;
  double origin[3] = {0, 0, 0};
  std::unordered_map<uint64_t, int32_t> cell;   // packed lattice key -> voxel
  // helpers
  std::vector<F3> hc, hn;          // centroid xyz / normal
  std::vector<uint8_t> alive;
  std::vector<std::set<int32_t>> leaves;   // schedule 0 only

  static uint64_t pack(int64_t x, int64_t y, int64_t z) { return ((uint64_t)x << 42) | ((uint64_t)y << 21) | (uint64_t)z; }
  int32_t find(int64_t x, int64_t y, int64_t z) const {
    if (x < 0 || y < 0 || z < 0 || x >= (1 << 21) || y >= (1 << 21) || z >= (1 << 21)) return -1;
    auto it = cell.find(pack(x, y, z));
    return it == cell.end() ? -1 : it->second;
  }

  // plane fit over the multiset { [v] + } for u in N(v) (owner filter): u, N(u) (owner filter)
  F3 fit_normal(int32_t v, int32_t filter /* -2: initial (self pushed first, no filter) */) const {
    const F3 K = xyz[v];
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    auto push = [&](int32_t w) {
      float x = xyz[w].x - K.x, y = xyz[w].y - K.y, z = xyz[w].z - K.z;
      acc[0] += x * x; acc[1] += x * y; acc[2] += x * z; acc[3] += y * y; acc[4] += y * z; acc[5] += z * z;
      acc[6] += x; acc[7] += y; acc[8] += z;
      cnt++;
    };
    if (filter == -2) push(v);
    for (int j = 0; j < 27; j++) {
      int32_t u = nb[(size_t)v * 27 + j];
      if (u < 0 || (filter >= 0 && owner[u] != filter)) continue;
      push(u);
      for (int k = 0; k < 27; k++) {
        int32_t w = nb[(size_t)u * 27 + k];
        if (w < 0 || (filter >= 0 && owner[w] != filter)) continue;
        push(w);
      }
    }
    const float nanv = std::numeric_limits<float>::quiet_NaN();
    if (cnt < 3) return F3{nanv, nanv, nanv};
    for (int i = 0; i < 9; i++) acc[i] /= (float)cnt;
    float cov[6] = {acc[0] - acc[6] * acc[6], acc[1] - acc[6] * acc[7], acc[2] - acc[6] * acc[8],
                    acc[3] - acc[7] * acc[7], acc[4] - acc[7] * acc[8], acc[5] - acc[8] * acc[8]};
    F3 n = smallest_eigenvector(cov);
    // flipNormalTowardsViewpoint(point, 0, 0, 0, normal); normal[3] = 0; normalize
    F3 vp{0.0f - K.x, 0.0f - K.y, 0.0f - K.z};
    if (dot3(vp, n) < 0) { n.x *= -1; n.y *= -1; n.z *= -1; }
    float z = (n.x * n.x + n.y * n.y) + n.z * n.z;
    if (z > 0.f) { float s = std::sqrt(z); n.x /= s; n.y /= s; n.z /= s; }
    return n;
  }

  // ---- schedule 1: the same plane fit with ORDER-INDEPENDENT sums (what the CUDA generator computes).  Voxel centroids
  //      in 2^-20 fixed point (64-bit integers); every voxel u first gets the moments of its own 27-neighbourhood relative
  //      to itself (fit_moments); the two-ring multiset of v is then the sum of its neighbours' records translated by
  //      q_u - q_v.  Integer arithmetic is exact, so 27 + 27 gathers give what the 27 x 27 walk of fit_normal sums. ----
  struct Mom { int64_t n, s[3], m[6]; };
  std::vector<Mom> mom;
  static void q3(const F3& p, int64_t q[3]) {
    q[0] = (int64_t)std::llrint((double)p.x * 1048576.0); q[1] = (int64_t)std::llrint((double)p.y * 1048576.0);
    q[2] = (int64_t)std::llrint((double)p.z * 1048576.0);
  }
  void fit_moments(bool filtered) {
    mom.assign((size_t)V, Mom{0, {0, 0, 0}, {0, 0, 0, 0, 0, 0}});
    for (int64_t u = 0; u < V; u++) {
      const int32_t f = filtered ? owner[u] : -2;
      if (filtered && f < 0) continue;
      int64_t qu[3];
      q3(xyz[u], qu);
      Mom A{0, {0, 0, 0}, {0, 0, 0, 0, 0, 0}};
      for (int k = 0; k < 27; k++) {
        const int32_t w = nb[(size_t)u * 27 + k];
        if (w < 0 || (f >= 0 && owner[w] != f)) continue;
        int64_t qw[3];
        q3(xyz[w], qw);
        const int64_t x = qw[0] - qu[0], y = qw[1] - qu[1], z = qw[2] - qu[2];
        A.n++; A.s[0] += x; A.s[1] += y; A.s[2] += z;
        A.m[0] += x * x; A.m[1] += x * y; A.m[2] += x * z; A.m[3] += y * y; A.m[4] += y * z; A.m[5] += z * z;
      }
      mom[u] = A;
    }
  }
  F3 fit_normal_fixed(int32_t v, int32_t filter /* -2: initial (self counted once more, no filter) */) const {
    const F3 K = xyz[v];
    int64_t qv[3];
    q3(K, qv);
    int64_t n = filter == -2 ? 1 : 0, S[3] = {0, 0, 0}, M[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 27; j++) {
      const int32_t u = nb[(size_t)v * 27 + j];
      if (u < 0 || (filter >= 0 && owner[u] != filter)) continue;
      int64_t qu[3];
      q3(xyz[u], qu);
      const int64_t tx = qu[0] - qv[0], ty = qu[1] - qv[1], tz = qu[2] - qv[2];
      const Mom& A = mom[u];
      const int64_t c = A.n + 1;                    // u itself + its neighbourhood
      n += c;
      S[0] += A.s[0] + c * tx; S[1] += A.s[1] + c * ty; S[2] += A.s[2] + c * tz;
      M[0] += A.m[0] + 2 * A.s[0] * tx + c * tx * tx;
      M[1] += A.m[1] + A.s[0] * ty + tx * A.s[1] + c * tx * ty;
      M[2] += A.m[2] + A.s[0] * tz + tx * A.s[2] + c * tx * tz;
      M[3] += A.m[3] + 2 * A.s[1] * ty + c * ty * ty;
      M[4] += A.m[4] + A.s[1] * tz + ty * A.s[2] + c * ty * tz;
      M[5] += A.m[5] + 2 * A.s[2] * tz + c * tz * tz;
    }
    const float nanv = std::numeric_limits<float>::quiet_NaN();
    if (n < 3) return F3{nanv, nanv, nanv};
    const double cd = (double)n, sc = 9.094947017729282379150390625e-13;   // 2^-40: fixed point^2 -> m^2
    const double mx = (double)S[0] / cd, my = (double)S[1] / cd, mz = (double)S[2] / cd;
    float cov[6] = {(float)(((double)M[0] / cd - mx * mx) * sc), (float)(((double)M[1] / cd - mx * my) * sc),
                    (float)(((double)M[2] / cd - mx * mz) * sc), (float)(((double)M[3] / cd - my * my) * sc),
                    (float)(((double)M[4] / cd - my * mz) * sc), (float)(((double)M[5] / cd - mz * mz) * sc)};
    F3 nn = smallest_eigenvector(cov);
    F3 vp{0.0f - K.x, 0.0f - K.y, 0.0f - K.z};
    if (dot3(vp, nn) < 0) { nn.x *= -1; nn.y *= -1; nn.z *= -1; }
    float z = (nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z;
    if (z > 0.f) { float sq = std::sqrt(z); nn.x /= sq; nn.y /= sq; nn.z /= sq; }
    return nn;
  }

  // voxelDataDistance(centroid of helper h, voxel v)
  float distance(int32_t h, int32_t v) const {
    F3 d{hc[h].x - xyz[v].x, hc[h].y - xyz[v].y, hc[h].z - xyz[v].z};
    float spatial = std::sqrt(sqn3(d)) / seed;
    float color = 0.0f;
    float cosn = 1.0f - std::fabs(dot3(hn[h], nrm[v]));
    return cosn * wn + color * wc + spatial * ws;
  }

  // exact nearest voxel to a point by growing lattice cubes; -1 when nothing within 32 voxel layers
  int32_t nearest_voxel(F3 c) const {
    const double r = (double)res;
    int64_t k[3] = {(int64_t)std::floor(((double)c.x - origin[0]) / r), (int64_t)std::floor(((double)c.y - origin[1]) / r),
                    (int64_t)std::floor(((double)c.z - origin[2]) / r)};
    for (int R = 2; R <= 32; R *= 2) {
      float best = std::numeric_limits<float>::max();
      int32_t bv = -1;
      for (int64_t x = k[0] - R; x <= k[0] + R; x++)
        for (int64_t y = k[1] - R; y <= k[1] + R; y++)
          for (int64_t z = k[2] - R; z <= k[2] + R; z++) {
            int32_t v = find(x, y, z);
            if (v < 0) continue;
            F3 d{xyz[v].x - c.x, xyz[v].y - c.y, xyz[v].z - c.z};
            float d2 = sqn3(d);
            if (d2 < best || (d2 == best && v < bv)) { best = d2; bv = v; }
          }
      if (bv >= 0 && std::sqrt((double)best) < (double)R * r * (1.0 - 1e-5)) return bv;
    }
    return -1;
  }

  void update_centroid_float(int32_t h) {   // SupervoxelHelper::updateCentroid, schedule 0
    F3 sx{0, 0, 0}, sn{0, 0, 0};
    for (int32_t v : leaves[h]) {
      sn.x += nrm[v].x; sn.y += nrm[v].y; sn.z += nrm[v].z;
      sx.x += xyz[v].x; sx.y += xyz[v].y; sx.z += xyz[v].z;
    }
    float z = (sn.x * sn.x + sn.y * sn.y) + sn.z * sn.z;
    if (z > 0.f) { float s = std::sqrt(z); sn.x /= s; sn.y /= s; sn.z /= s; }
    float c = (float)leaves[h].size();
    hn[h] = sn;
    hc[h] = F3{sx.x / c, sx.y / c, sx.z / c};
  }

  void update_centroids_fixed() {           // schedule 1: order-independent sums
    const int64_t H = (int64_t)hc.size();
    std::vector<long long> acc((size_t)H * 6, 0);
    std::vector<int64_t> cnt((size_t)H, 0);
    for (int64_t v = 0; v < V; v++) {
      int32_t h = owner[v];
      if (h < 0) continue;
      long long* a = &acc[(size_t)h * 6];
      a[0] += std::llrint((double)xyz[v].x * 1048576.0); a[1] += std::llrint((double)xyz[v].y * 1048576.0);
      a[2] += std::llrint((double)xyz[v].z * 1048576.0);
      // NaN normals (isolated voxels) contribute nothing
      if (nrm[v].x == nrm[v].x) {
        a[3] += std::llrint((double)nrm[v].x * 1073741824.0); a[4] += std::llrint((double)nrm[v].y * 1073741824.0);
        a[5] += std::llrint((double)nrm[v].z * 1073741824.0);
      }
      cnt[h]++;
    }
    for (int64_t h = 0; h < H; h++) {
      if (!alive[h]) continue;
      if (cnt[h] == 0) { alive[h] = 0; continue; }
      const long long* a = &acc[(size_t)h * 6];
      const double c = (double)cnt[h];
      hc[h] = F3{(float)((double)a[0] / 1048576.0 / c), (float)((double)a[1] / 1048576.0 / c), (float)((double)a[2] / 1048576.0 / c)};
      F3 sn{(float)((double)a[3] / 1073741824.0), (float)((double)a[4] / 1073741824.0), (float)((double)a[5] / 1073741824.0)};
      float z = (sn.x * sn.x + sn.y * sn.y) + sn.z * sn.z;
      if (z > 0.f) { float s = std::sqrt(z); sn.x /= s; sn.y /= s; sn.z /= s; }
      hn[h] = sn;
    }
  }

  void expand_sequential(int depth) {       // expandSupervoxels + SupervoxelHelper::expand
    const int32_t H = (int32_t)hc.size();
    for (int it = 1; it < depth; it++) {
      for (int32_t h = 0; h < H; h++) {
        if (!alive[h]) continue;
        std::vector<int32_t> fresh;
        for (int32_t u : leaves[h])
          for (int j = 0; j < 27; j++) {
            int32_t v = nb[(size_t)u * 27 + j];
            if (v < 0 || owner[v] == h) continue;
            float d = distance(h, v);
            if (d < dist[v]) {
              dist[v] = d;
              if (owner[v] >= 0) leaves[owner[v]].erase(v);
              owner[v] = h;
              fresh.push_back(v);
            }
          }
        for (int32_t v : fresh) leaves[h].insert(v);
      }
      for (int32_t h = 0; h < H; h++) {
        if (!alive[h]) continue;
        if (leaves[h].empty()) alive[h] = 0; else update_centroid_float(h);
      }
    }
  }

  void expand_synchronous(int depth) {
    std::vector<int32_t> next((size_t)V);
    for (int it = 1; it < depth; it++) {
      for (int64_t v = 0; v < V; v++) {
        const int32_t own = owner[v];
        float bd = dist[v];
        int32_t bh = own;
        for (int j = 0; j < 27; j++) {
          int32_t u = nb[(size_t)v * 27 + j];
          if (u < 0) continue;
          int32_t h = owner[u];
          if (h < 0 || h == own || !alive[h]) continue;
          float d = distance(h, (int32_t)v);
          if (d < bd || (d == bd && bh != own && h < bh)) { bd = d; bh = h; }
        }
        next[v] = bh;
        dist[v] = bd;
      }
      owner.swap(next);
      update_centroids_fixed();
    }
  }
};

}  // namespace

extern "C" int vgso_vccs(const float* xyz, int64_t n, int stride, int64_t V, const uint32_t* vox_key, const int64_t* vox_off,
                         const int32_t* vox_pts, const double* origin3, const vgso_vccs_params* p, int32_t* point_label,
                         int32_t* max_label, float* vox_normal, int32_t* vox_label) {
  Vccs S;
  S.V = V; S.key = vox_key;
  S.res = p->voxel_res; S.seed = p->seed_res; S.wc = p->color_importance; S.ws = p->spatial_importance; S.wn = p->normal_importance;
  for (int a = 0; a < 3; a++) S.origin[a] = origin3[a];
  // --- computeVoxelData: voxel centroid = float sum of its points (ascending index) / count ---
  S.xyz.resize((size_t)V);
  for (int64_t v = 0; v < V; v++) {
    float sx = 0, sy = 0, sz = 0;
    for (int64_t j = vox_off[v]; j < vox_off[v + 1]; j++) {
      const float* q = xyz + (int64_t)vox_pts[j] * stride;
      sx += q[0]; sy += q[1]; sz += q[2];
    }
    float c = (float)(vox_off[v + 1] - vox_off[v]);
    S.xyz[v] = F3{sx / c, sy / c, sz / c};
    S.cell[Vccs::pack(vox_key[3 * v], vox_key[3 * v + 1], vox_key[3 * v + 2])] = (int32_t)v;
  }
  S.nb.assign((size_t)V * 27, -1);
  for (int64_t v = 0; v < V; v++)
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++)
          S.nb[(size_t)v * 27 + (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)] =
              S.find((int64_t)vox_key[3 * v] + dx, (int64_t)vox_key[3 * v + 1] + dy, (int64_t)vox_key[3 * v + 2] + dz);
  S.owner.assign((size_t)V, -1);
  S.dist.assign((size_t)V, std::numeric_limits<float>::max());
  S.nrm.resize((size_t)V);
  if (p->schedule == 1) {
    S.fit_moments(false);
    for (int64_t v = 0; v < V; v++) S.nrm[v] = S.fit_normal_fixed((int32_t)v, -2);
  } else {
    for (int64_t v = 0; v < V; v++) S.nrm[v] = S.fit_normal((int32_t)v, -2);
  }
  if (vox_normal)
    for (int64_t v = 0; v < V; v++) { vox_normal[3 * v] = S.nrm[v].x; vox_normal[3 * v + 1] = S.nrm[v].y; vox_normal[3 * v + 2] = S.nrm[v].z; }

  // --- selectInitialSupervoxelSeeds ---
  const double sd = (double)p->seed_res;
  std::vector<std::pair<uint64_t, int32_t>> cells;   // (x-major morton of the seed cell, voxel)
  auto morton = [](uint64_t x, uint64_t y, uint64_t z) {
    uint64_t m = 0;
    for (int b = 20; b >= 0; b--) m = (m << 3) | (((x >> b) & 1) << 2) | (((y >> b) & 1) << 1) | ((z >> b) & 1);
    return m;
  };
  std::vector<int64_t> vcell((size_t)V * 3);
  for (int64_t v = 0; v < V; v++) {
    vcell[3 * v] = (int64_t)std::floor(((double)S.xyz[v].x - S.origin[0]) / sd);
    vcell[3 * v + 1] = (int64_t)std::floor(((double)S.xyz[v].y - S.origin[1]) / sd);
    vcell[3 * v + 2] = (int64_t)std::floor(((double)S.xyz[v].z - S.origin[2]) / sd);
    cells.emplace_back(morton((uint64_t)vcell[3 * v], (uint64_t)vcell[3 * v + 1], (uint64_t)vcell[3 * v + 2]), (int32_t)v);
  }
  std::sort(cells.begin(), cells.end());
  std::unordered_map<uint64_t, int32_t> cell_index;    // morton -> rank of the occupied cell
  std::vector<std::array<int64_t, 3>> cell_xyz;
  for (size_t i = 0; i < cells.size(); i++)
    if (i == 0 || cells[i].first != cells[i - 1].first) {
      cell_index[cells[i].first] = (int32_t)cell_xyz.size();
      int32_t v = cells[i].second;
      cell_xyz.push_back({vcell[3 * v], vcell[3 * v + 1], vcell[3 * v + 2]});
    }
  const int64_t NC = (int64_t)cell_xyz.size();
  std::vector<float> best_d((size_t)NC, std::numeric_limits<float>::max());
  std::vector<int32_t> best_v((size_t)NC, -1);
  for (int64_t v = 0; v < V; v++)
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
          int64_t cx = vcell[3 * v] + dx, cy = vcell[3 * v + 1] + dy, cz = vcell[3 * v + 2] + dz;
          if (cx < 0 || cy < 0 || cz < 0) continue;
          auto it = cell_index.find(morton((uint64_t)cx, (uint64_t)cy, (uint64_t)cz));
          if (it == cell_index.end()) continue;
          F3 ctr{(float)(((double)cx + 0.5) * sd + S.origin[0]), (float)(((double)cy + 0.5) * sd + S.origin[1]),
                 (float)(((double)cz + 0.5) * sd + S.origin[2])};
          F3 d{S.xyz[v].x - ctr.x, S.xyz[v].y - ctr.y, S.xyz[v].z - ctr.z};
          float d2 = sqn3(d);
          int32_t c = it->second;
          if (d2 < best_d[c] || (d2 == best_d[c] && (int32_t)v < best_v[c])) { best_d[c] = d2; best_v[c] = (int32_t)v; }
        }
  const float search_radius = 0.5f * p->seed_res;
  const float min_points = 0.05f * (search_radius) * (search_radius) * 3.1415926536f / (p->voxel_res * p->voxel_res);
  const float r2 = search_radius * search_radius;
  const int reach = (int)std::ceil((double)search_radius / (double)p->voxel_res) + 1;
  std::vector<int32_t> seeds;
  std::vector<uint8_t> taken((size_t)V, 0);
  for (int64_t c = 0; c < NC; c++) {
    int32_t s = best_v[c];
    if (s < 0 || taken[s]) continue;      // two cells sharing their nearest voxel: one supervoxel (smaller label)
    int num = 0;
    for (int dx = -reach; dx <= reach; dx++)
      for (int dy = -reach; dy <= reach; dy++)
        for (int dz = -reach; dz <= reach; dz++) {
          int32_t w = S.find((int64_t)vox_key[3 * s] + dx, (int64_t)vox_key[3 * s + 1] + dy, (int64_t)vox_key[3 * s + 2] + dz);
          if (w < 0) continue;
          F3 d{S.xyz[w].x - S.xyz[s].x, S.xyz[w].y - S.xyz[s].y, S.xyz[w].z - S.xyz[s].z};
          if (sqn3(d) < r2) num++;
        }
    if ((float)num > min_points) { seeds.push_back(s); taken[s] = 1; }
  }
  // --- createSupervoxelHelpers ---
  const int32_t H = (int32_t)seeds.size();
  S.hc.resize((size_t)H); S.hn.resize((size_t)H); S.alive.assign((size_t)H, 1);
  if (p->schedule == 0) S.leaves.resize((size_t)H);
  for (int32_t h = 0; h < H; h++) {
    S.owner[seeds[h]] = h;
    S.hc[h] = S.xyz[seeds[h]];
    S.hn[h] = S.nrm[seeds[h]];
    if (p->schedule == 0) { S.leaves[h].insert(seeds[h]); S.update_centroid_float(h); }   // addLeaf + updateCentroid
  }
  const int depth = (int)(1.8f * p->seed_res / p->voxel_res);
  if (p->schedule == 0) S.expand_sequential(depth); else S.expand_synchronous(depth);

  // --- refineSupervoxels ---
  for (int it = 0; it < p->refine_iterations; it++) {
    if (p->schedule == 1) S.fit_moments(true);
    for (int64_t v = 0; v < V; v++)
      if (S.owner[v] >= 0) S.nrm[v] = p->schedule == 1 ? S.fit_normal_fixed((int32_t)v, S.owner[v]) : S.fit_normal((int32_t)v, S.owner[v]);
    std::vector<int32_t> seedv((size_t)H, -1);
    for (int32_t h = 0; h < H; h++)
      if (S.alive[h]) seedv[h] = S.nearest_voxel(S.hc[h]);
    std::fill(S.owner.begin(), S.owner.end(), -1);
    std::fill(S.dist.begin(), S.dist.end(), std::numeric_limits<float>::max());
    if (p->schedule == 0) {
      for (int32_t h = 0; h < H; h++) {
        S.leaves[h].clear();
        if (!S.alive[h]) continue;
        if (seedv[h] >= 0) { S.leaves[h].insert(seedv[h]); S.owner[seedv[h]] = h; }
      }
      S.expand_sequential(depth);
    } else {
      for (int32_t h = 0; h < H; h++) {
        if (!S.alive[h]) continue;
        if (seedv[h] < 0 || S.owner[seedv[h]] >= 0) { S.alive[h] = 0; continue; }   // no voxel, or taken by a smaller label
        S.owner[seedv[h]] = h;
      }
      S.expand_synchronous(depth);
    }
  }

  // --- getLabeledCloud / getMaxLabel ---
  int32_t ml = 0;
  for (int32_t h = 0; h < H; h++) {
    bool live = S.alive[h] != 0;
    if (p->schedule == 0) live = live && !S.leaves[h].empty();
    if (live) ml = std::max(ml, h + 1);
  }
  *max_label = ml;
  for (int64_t i = 0; i < n; i++) point_label[i] = 0;
  for (int64_t v = 0; v < V; v++) {
    int32_t l = S.owner[v] >= 0 ? S.owner[v] + 1 : 0;
    if (vox_label) vox_label[v] = l;
    for (int64_t j = vox_off[v]; j < vox_off[v + 1]; j++) point_label[vox_pts[j]] = l;
  }
  return (int)H;
}

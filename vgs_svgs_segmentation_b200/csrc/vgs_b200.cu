// vgs_b200.cu — host pipeline + C ABI (include/vgs_b200.h) of the VGS/SVGS hot path on sm_100a.
// No CPU fallback: every entry point needs a CUDA device.
#include "../../include/vgs_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include "vgs_kernels.cuh"
#include "vgs_rows.cuh"
#include "vgs_vccs.cuh"
#include "vgs_hostio.cuh"

using namespace vgs;

namespace {

thread_local std::string g_create_error;

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace

struct vgs_context {
  int mode = 0, device = 0, leaf_order = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int64_t launches = 0;

  // input
  const float* d_xyz = nullptr;
  int stride = 3;
  int64_t n = 0;
  DBuf xyz_own;
  const int32_t* d_labels = nullptr;
  int32_t max_label = 0;
  DBuf labels_own;

  // octree
  float voxel_size = 0;
  Box box{};
  int depth = 0;
  int64_t n_finite = 0;
  EpochTable ep{};
  bool voxelized = false;
  int64_t n_voxels = 0;   // octree leaves (VGS units; SVGS: getVoxelNum only)

  // units
  int64_t nu = 0;         // units: voxels (VGS) / supervoxels (SVGS)
  int64_t n_valid = 0;    // points that belong to a unit (sorted positions [0, n_valid))
  bool have_graph = false;       // connect lists of stage 4+5a complete (own range computed or imported)
  bool units_external = false;   // SVGS units made by vgs_make_supervoxels_grid (not from labels)
  bool have_units = false, have_features = false, have_adj = false, have_segments = false, have_geometry = false;
  float bb_f[6] = {0, 0, 0, 0, 0, 0};   // float-narrowed bounding box members (VS.h:1123)
  int64_t n_used = 0, n_adj = 0, n_pairs = 0, max_n = 0, n_singles = 0, n_attached = 0, closest_rounds = 0;
  int last_voxels_min = std::numeric_limits<int>::min();
  bool have_cluster_stats = false;
  bool have_csr = false;            // cluster export (vgs_get_clusters_csr) cached on the device
  int csr_min_excl = 0;
  int64_t csr_total = 0;
  DBuf csr_off, csr_idx;
  int64_t n_clusters_all = 0, n_clusters_exp = 0;
  int points_min = 0;
  float graph_size = 0;

  // device buffers
  DBuf keysA, keysB, valsA, valsB, hist, tiles, flags, scan, small;
  DBuf ustart, ukey, pos_unit, rec, key3, center, plainm, tk, tv, stencil;
  DBuf adj_cnt, adj_off, adj_idx, adj_code, class_count, class_list;
  DBuf conn0_cnt, conn0_idx, conn1_cnt, conn1_idx, attach, parent, root, csize, cminpt, labels_out, tmp;
  int sort_key_bytes = 8;       // width of the voxel sort keys of the last vgs_voxelize
  uint32_t* d_perm = nullptr;   // sorted point indices
  uint32_t* vox_perm = nullptr;  // point indices in voxel order as vgs_voxelize left them (with ustart / ukey: the voxel table)
  bool vox_table_valid = false;  // ... until another sort reuses the shared buffers (SVGS unit builders)
  uint64_t hmask = 0;
  // lattice searches (VGS): host tables rebuilt only when (voxel_size, graph_size, float-noise bound) change
  std::vector<int4> stencil_host;   // radius stencil, sorted by integer distance class
  std::vector<int4> adj_cols_host, pc_cols_host;   // columns (dx, dy, mask of dz) of the radius / pair stencil
  int st_rho = 0, st_r2 = 0, lbits = 3, mwords = 0;
  float st_vs = -1.f, st_gs = -1.f;
  double st_noise = -1.0;
  bool rows_ok = false;             // pair-weight rows usable (stencil reach <= 5 cells)
  int64_t max_row_len = 0, n_rows = 0, n_long = 0, grid_bytes = 0;
  uint32_t kminmax[6] = {0, 0, 0, 0, 0, 0};   // smallest / largest occupied voxel key per axis
  BitGrid grid{};
  LatticeGeom lgeo{};
  DBuf d_adj_cols, d_pc_cols, tb_slot, tb_code5, tb_first, tb_last;
  DBuf bm_all, bm_used, idgrid, row_len, row_off, rows, long_rows, cstats, conn_mask;
  bool use_idgrid = false;
  uint64_t idgrid_budget = 8ull << 30;   // bytes (VGS_B200_IDGRID_MB; 0 = always the hash table)
  DBuf fallback, uflags, singles, singles_dep, singles_best, closest_w, used_list, origin_state;
  bool conn0_is_mask = false;       // connect lists of stage 5a held as lattice-offset masks (VGS row kernel)
  int64_t n_fallback = 0;
  DBuf ckeysA, ckeysB, cvalsA, cvalsB, cstart, ckey, cpos, gridmin;   // SVGS centroid grid
  int use_pair_cache = 1;           // 0 = evaluate weights inside every local graph (general kernel; VGS_B200_NO_PAIR_CACHE)
  int cc_jumps = 6;                 // pointer-jumping rounds over the initial component forest
  int lg_order = 0;                 // local-graph launch order: 0 = voxel id, b = by neighbourhood size in 2^b classes, big first (VGS_B200_LG_ORDER)
  int lr_target = LR_TARGET;        // staged entries the local-graph rounds aim at (VGS_B200_LR_TARGET)
  int force_fallback = 0;           // test knob VGS_B200_FORCE_FALLBACK=m: the row kernel hands every m-th voxel to the general kernel
  // supervoxel generator (vgs_make_supervoxels_vccs): its own voxel table and working set
  struct {
    DBuf keysA, keysB, valsA, valsB, start, key, pos, xyz, key3, plain, ptvox, nb, nb_cnt, nb_off, nb_csr, nrm, mom, bits, owner, owner2, dist, claim;
    DBuf ckA, ckB, cvA, cvB, cstart, ckey, cpos, cell3, best, flag, rank, seedv, hc, hn, alive, acc, tk, tv, tk2, tv2;
  } vc;
  int64_t vccs_seeds = 0;

  std::vector<uint32_t> host_u32;   // host staging of small index arrays
  vgs_timings tm{};
  cudaEvent_t ev[32] = {};
  // the size-class launches of the local graph stage are independent: they run on the handle's stream plus these,
  // forked / joined with events, so that the small classes of big units fill the tails of the large ones
  static constexpr int N_AUX = 7;
  cudaStream_t aux[N_AUX] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[N_AUX] = {};
  int class_streams = 6;            // VGS_B200_CLASS_STREAMS (1 = everything on the handle's stream; measured flat from 4 to 8)
  float* tm_slot[16] = {};
  unsigned tm_pending = 0;
  // per-kernel-group timers (vgs_kernel_timings): one event pair per group and run
  static constexpr int NK = 20;
  cudaEvent_t kev[2 * NK] = {};
  float k_ms[NK] = {};
  int k_launches[NK] = {};
  unsigned k_pending = 0;

  vgs_status fail(vgs_status s, const std::string& m) { err = m; return s; }
  vgs_status fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at vgs_b200.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    err = buf;
    return VGS_ERR_CUDA;
  }
};

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) return h->fail_cuda(e_, #x, __LINE__);             \
  } while (0)
// In-step host synchronisation.  Default: cudaStreamSynchronize.  VGS_B200_SPIN_SYNC=1 polls the stream instead of
// sleeping (lower wake-up latency on an idle host, worse on an oversubscribed one: measured both ways on shared boxes).
static int g_spin_sync = -1;
static inline cudaError_t stream_wait(cudaStream_t st) {
  if (g_spin_sync < 0) { const char* e = getenv("VGS_B200_SPIN_SYNC"); g_spin_sync = (e && e[0] == '1') ? 1 : 0; }
  if (!g_spin_sync) return cudaStreamSynchronize(st);
  for (;;) {
    const cudaError_t r = cudaStreamQuery(st);
    if (r != cudaErrorNotReady) return r;
  }
}

#define LAUNCH(kernel, grid, block, smem, ...)                                \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), h->stream>>>(__VA_ARGS__);              \
    h->launches++;                                                            \
    CK(cudaGetLastError());                                                   \
  } while (0)

#define LAUNCH_ON(st, kernel, grid, block, smem, ...)                        \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                   \
    h->launches++;                                                            \
    CK(cudaGetLastError());                                                   \
  } while (0)

namespace {

// exclusive scan of n u32 (in -> out, may alias); total (u64) read back to the host if total != null
vgs_status scan_u32(vgs_handle h, const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* total_host,
                    unsigned long long* d_total = nullptr) {
  int64_t nt = cdiv(n, SC_TILE);
  if (nt < 1) nt = 1;
  CK(h->tiles.reserve((size_t)(nt + 1) * 4 + 16));
  CK(h->small.reserve(4096));
  if (!d_total) d_total = h->small.as<unsigned long long>() + 8;
  LAUNCH(k_scan_reduce, (unsigned)nt, SC_THREADS, 0, in, n, h->tiles.as<uint32_t>());
  LAUNCH(k_scan_tiles, 1, 1024, 0, h->tiles.as<uint32_t>(), nt, d_total);
  LAUNCH(k_scan_down, (unsigned)nt, SC_THREADS, 0, in, out, n, h->tiles.as<uint32_t>());
  if (total_host) {
    CK(cudaMemcpyAsync(total_host, d_total, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
  }
  return VGS_OK;
}

// stable LSD radix sort of (key,val) over the low nbits bits; result pointers returned
template <class K>
vgs_status radix_sort(vgs_handle h, int64_t n, int nbits, K** keys_out, uint32_t** vals_out,
                      K* ka = nullptr, K* kb = nullptr, uint32_t* va = nullptr, uint32_t* vb = nullptr) {
  if (!ka) { ka = h->keysA.as<K>(); kb = h->keysB.as<K>(); va = h->valsA.as<uint32_t>(); vb = h->valsB.as<uint32_t>(); }
  int64_t nblk = cdiv(n, RS_TILE);
  if (nblk < 1) nblk = 1;
  CK(h->hist.reserve((size_t)nblk * 256 * 4));
  for (int shift = 0; shift < nbits; shift += 8) {
    LAUNCH(k_rs_hist<K>, (unsigned)nblk, RS_THREADS, 0, ka, n, shift, h->hist.as<uint32_t>(), nblk);
    vgs_status s = scan_u32(h, h->hist.as<uint32_t>(), h->hist.as<uint32_t>(), nblk * 256, nullptr);
    if (s) return s;
    LAUNCH(k_rs_scatter<K>, (unsigned)nblk, RS_THREADS, 0, ka, va, kb, vb, n, shift, h->hist.as<uint32_t>(), nblk);
    std::swap(ka, kb); std::swap(va, vb);
  }
  *keys_out = ka; *vals_out = va;
  return VGS_OK;
}

// sorted keys -> unit table.  n_valid = number of sorted positions with key < sentinel.
template <class K>
vgs_status build_units(vgs_handle h, const K* keys, int64_t n_valid, int64_t* n_units, DBuf* ustart = nullptr,
                       DBuf* ukey = nullptr, DBuf* pos_unit = nullptr) {
  if (!ustart) { ustart = &h->ustart; ukey = &h->ukey; pos_unit = &h->pos_unit; }
  if (n_valid <= 0) { *n_units = 0; return VGS_OK; }
  CK(pos_unit->reserve((size_t)n_valid * 4));
  const int64_t nt = std::max<int64_t>(1, cdiv(n_valid, SC_TILE));
  CK(h->tiles.reserve((size_t)(nt + 1) * 4 + 16));
  CK(h->small.reserve(4096));
  unsigned long long* d_total = h->small.as<unsigned long long>() + 8;
  LAUNCH(k_heads_reduce<K>, (unsigned)nt, SC_THREADS, 0, keys, n_valid, h->tiles.as<uint32_t>());
  LAUNCH(k_scan_tiles, 1, 1024, 0, h->tiles.as<uint32_t>(), nt, d_total);
  unsigned long long total = 0;
  CK(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  *n_units = (int64_t)total;
  CK(ustart->reserve((size_t)(total + 1) * 4));
  CK(ukey->reserve((size_t)(total + 1) * 8));
  LAUNCH(k_heads_down<K>, (unsigned)nt, SC_THREADS, 0, keys, n_valid, h->tiles.as<uint32_t>(), ustart->as<uint32_t>(), ukey->as<uint64_t>(),
         pos_unit->as<uint32_t>());
  uint32_t endv = (uint32_t)n_valid;
  CK(cudaMemcpyAsync(ustart->as<uint32_t>() + total, &endv, 4, cudaMemcpyHostToDevice, h->stream));
  return VGS_OK;
}

// ---- PCL dynamic bounding box on the host side of stage 0 (octree_pointcloud.hpp
//      adoptBoundingBoxToPoint / getKeyBitSize); the device only finds the next violating point ----
struct OctState {
  double mn[3], mx[3], res;
  unsigned depth = 0;
  bool defined = false;
  struct Ev { unsigned lowered, depth_old; };
  std::vector<Ev> events;
  void adopt(const float p[3]) {
    const float eps = std::numeric_limits<float>::epsilon();
    while (true) {
      bool up[3], any = false;
      for (int a = 0; a < 3; a++) {
        bool lo = p[a] < mn[a];
        up[a] = p[a] >= mx[a];
        any = any || lo || up[a];
      }
      if (!any && defined) break;
      if (defined) {
        double side = (double)(1 << depth) * res;
        unsigned lowered = 0;
        for (int a = 0; a < 3; a++) if (!up[a]) { mn[a] -= side; lowered |= 1u << a; }
        events.push_back(Ev{lowered, depth});
        depth++;
        side = (double)(1 << depth) * res - eps;
        for (int a = 0; a < 3; a++) mx[a] = mn[a] + side;
        if (depth > 30) break;
      } else {
        for (int a = 0; a < 3; a++) { mn[a] = p[a] - res / 2; mx[a] = p[a] + res / 2; }
        unsigned mk = 2;
        for (int a = 0; a < 3; a++) mk = std::max(mk, (unsigned)((mx[a] - mn[a]) / res));
        depth = (unsigned)std::ceil(std::log((double)mk) / std::log(2.0) - eps);
        double side = (double)(1 << depth) * res - eps;
        for (int a = 0; a < 3; a++) {
          double over = (side - (mx[a] - mn[a])) / 2.0;
          mn[a] -= over; mx[a] += over;
        }
        defined = true;
      }
    }
  }
};

vgs_status find_next(vgs_handle h, int64_t cursor, const OctState& st, int64_t* found, float* xyz3) {
  unsigned long long* d_found = h->small.as<unsigned long long>();      // [0] index, [1..2] the point's coordinates (3 floats)
  Box b;
  for (int a = 0; a < 3; a++) { b.mn[a] = st.mn[a]; b.mx[a] = st.mx[a]; }
  int64_t window = 1 << 18;
  int64_t s = cursor;
  *found = h->n;
  while (s < h->n) {
    int64_t e = std::min(h->n, s + window);
    unsigned long long init = ~0ull;
    CK(cudaMemcpyAsync(d_found, &init, 8, cudaMemcpyHostToDevice, h->stream));
    int64_t blocks = std::min<int64_t>(cdiv(e - s, 256), 148 * 8);
    LAUNCH(k_find_outside, (unsigned)blocks, 256, 0, h->d_xyz, h->stride, s, e, b, st.defined ? 1 : 0, d_found, nullptr);
    LAUNCH(k_fetch_found, 1, 1, 0, h->d_xyz, h->stride, d_found, reinterpret_cast<float*>(d_found + 1));
    struct { unsigned long long idx; float p[4]; } r;
    CK(cudaMemcpyAsync(&r, d_found, 24, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    if (r.idx != ~0ull) { *found = (int64_t)r.idx; xyz3[0] = r.p[0]; xyz3[1] = r.p[1]; xyz3[2] = r.p[2]; return VGS_OK; }
    s = e;
    window *= 64;     // violations come early or never: a short window first, then (nearly) the rest in one launch
  }
  return VGS_OK;
}

// Growth epochs of PCL's dynamic bounding box: every point that falls outside the current box (in insertion order)
// grows it and starts an epoch; later growth events shift the keys of earlier epochs (the old root becomes a child).
// Shared by vgs_voxelize (one device) and the slab group (global insertion order over all ranks).
struct OriginBuilder {
  OctState st;
  EpochTable ep;
  std::vector<size_t> events_before;   // #growth events before each epoch started
  void begin(double res) { st = OctState(); st.res = res; ep.n = 0; events_before.clear(); }
  // the point with (global) index idx violates the current box
  const char* add(long long idx, const float p[3]) {
    st.adopt(p);
    if (st.depth > 21) return "octree depth > 21 bits per axis (extent / voxel_size too large)";
    if (ep.n >= MAX_EPOCHS) return "too many bounding-box growth epochs";
    ep.viol[ep.n] = idx;
    for (int a = 0; a < 3; a++) ep.mn[ep.n][a] = st.mn[a];
    events_before.push_back(st.events.size());
    ep.n++;
    return nullptr;
  }
  void finish() {   // growth events after an epoch move its keys by 1 << depth_old on lowered axes
    for (int e = 0; e < ep.n; e++) {
      for (int a = 0; a < 3; a++) ep.shift[e][a] = 0;
      for (size_t k = events_before[e]; k < st.events.size(); k++)
        for (int a = 0; a < 3; a++)
          if ((st.events[k].lowered >> a) & 1u) ep.shift[e][a] += 1u << st.events[k].depth_old;
    }
  }
  void install(vgs_handle h) const {
    h->ep = ep;
    for (int a = 0; a < 3; a++) { h->box.mn[a] = st.mn[a]; h->box.mx[a] = st.mx[a]; }
    for (int a = 0; a < 3; a++) { h->bb_f[a] = (float)st.mn[a]; h->bb_f[3 + a] = (float)st.mx[a]; }
    h->have_geometry = false;
    h->depth = (int)st.depth;
  }
};

// Stage timers: CUDA events recorded on the handle's stream WITHOUT a host synchronisation (a sync per
// stage costs a host round trip and exposes the step to host scheduling noise); the elapsed times are
// read when somebody asks for them (vgs_stage_timings, end of vgs_run), after one synchronisation.
struct StageTimer {
  vgs_handle h; int idx; cudaEvent_t a, b;
  StageTimer(vgs_handle h_, float* slot_, int idx_) : h(h_), idx(idx_) {
    a = h->ev[idx * 2]; b = h->ev[idx * 2 + 1];
    h->tm_slot[idx] = slot_;
    cudaEventRecord(a, h->stream);
  }
  void stop() {
    cudaEventRecord(b, h->stream);
    h->tm_pending |= 1u << idx;
  }
};

// kernel-group timer: brackets the launches of one group (names in vgs_kernel_timings)
enum KId { K_ORIGIN = 0, K_QUANTISE, K_SORT, K_HEADS, K_FEATURES, K_HASH, K_GRID, K_ADJ_COUNT, K_ADJ_FILL, K_ROWS_FILL, K_ROWS_SORT,
           K_GRAPH_ROWS, K_GRAPH_GENERAL, K_MUTUAL, K_CLOSEST, K_COMPONENTS, K_LABELS, K_VCCS, K_SVGS_UNITS, K_SVGS_ADJ };
struct KTimer {
  vgs_handle h; int id; int64_t l0;
  KTimer(vgs_handle h_, int id_) : h(h_), id(id_), l0(h_->launches) { cudaEventRecord(h->kev[2 * id], h->stream); }
  void stop() {
    cudaEventRecord(h->kev[2 * id + 1], h->stream);
    h->k_pending |= 1u << id;
    h->k_launches[id] = (int)(h->launches - l0);
  }
};

void resolve_timers(vgs_handle h) {
  if (!h->tm_pending && !h->k_pending) return;
  cudaStreamSynchronize(h->stream);
  for (int i = 0; i < vgs_context::NK; i++)
    if ((h->k_pending >> i) & 1u) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, h->kev[2 * i], h->kev[2 * i + 1]) == cudaSuccess) h->k_ms[i] = ms;
    }
  h->k_pending = 0;
  for (int i = 0; i < 15; i++)
    if ((h->tm_pending >> i) & 1u) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, h->ev[i * 2], h->ev[i * 2 + 1]) == cudaSuccess && h->tm_slot[i]) *h->tm_slot[i] = ms;
    }
  h->tm_pending = 0;
}

std::vector<int4> make_stencil(float voxel_size_f, float graph_size_f, double noise) {
  // integer lattice offsets whose ideal centre distance could pass the float test dist2 < (float)(r*r);
  // noise = bound on the float error of dist2 for this cloud (grows with the distance from the origin)
  double res = (double)voxel_size_f, r = (double)graph_size_f;
  int rho = (int)std::ceil(std::sqrt(r * r + noise) / res) + 1;
  double lim = r * r * (1.0 + 1e-4) + 1e-9 + noise;
  std::vector<int4> st;
  for (int dx = -rho; dx <= rho; dx++)
    for (int dy = -rho; dy <= rho; dy++)
      for (int dz = -rho; dz <= rho; dz++) {
        double d2 = res * res * (double)(dx * dx + dy * dy + dz * dz);
        if (d2 < lim) st.push_back(make_int4(dx, dy, dz, 0));
      }
  return st;
}

}  // namespace

extern "C" {

vgs_status vgs_device_count(int* n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (n) *n = (e == cudaSuccess) ? c : 0;
  return (e == cudaSuccess && c > 0) ? VGS_OK : VGS_ERR_NO_DEVICE;
}

vgs_status vgs_create(vgs_handle* out, const vgs_config* cfg) {
  if (!out || !cfg) { g_create_error = "vgs_create: null argument"; return VGS_ERR_INVALID; }
  *out = nullptr;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess || c <= 0) {
    g_create_error = std::string("vgs_create: no CUDA device (") + cudaGetErrorString(e) + "); libvgs_b200 has no CPU path";
    return VGS_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= c) { g_create_error = "vgs_create: bad device ordinal"; return VGS_ERR_INVALID; }
  if (cfg->mode != VGS_MODE_VGS && cfg->mode != VGS_MODE_SVGS) { g_create_error = "vgs_create: bad mode"; return VGS_ERR_INVALID; }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return VGS_ERR_CUDA; }
  vgs_handle h = new vgs_context();
  h->mode = cfg->mode; h->device = cfg->device; h->leaf_order = cfg->leaf_order;
  if (cfg->stream) h->stream = (cudaStream_t)cfg->stream;
  else {
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete h; return VGS_ERR_CUDA; }
    h->own_stream = true;
  }
  for (auto& ev : h->ev) cudaEventCreate(&ev);
  for (auto& ev : h->kev) cudaEventCreate(&ev);
  for (int i = 0; i < vgs_context::N_AUX; i++) {
    cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (const char* e_cs = getenv("VGS_B200_CLASS_STREAMS")) { int v = atoi(e_cs); if (v >= 1 && v <= 1 + vgs_context::N_AUX) h->class_streams = v; }
  if (const char* e_nc = getenv("VGS_B200_NO_PAIR_CACHE")) h->use_pair_cache = (e_nc[0] == '1') ? 0 : 1;
  if (const char* e_ff = getenv("VGS_B200_FORCE_FALLBACK")) { int v = atoi(e_ff); if (v >= 1) h->force_fallback = v; }
  if (const char* e_ig = getenv("VGS_B200_IDGRID_MB")) { long long v = atoll(e_ig); if (v >= 0) h->idgrid_budget = (uint64_t)v << 20; }
  if (const char* e_lo = getenv("VGS_B200_LG_ORDER")) { int v = atoi(e_lo); if (v >= 0 && v <= 8) h->lg_order = v; }
  if (const char* e_lt = getenv("VGS_B200_LR_TARGET")) { int v = atoi(e_lt); if (v >= 8 && v <= 200) h->lr_target = v; }
  if (const char* e_cj = getenv("VGS_B200_CC_JUMPS")) { int v = atoi(e_cj); if (v >= 0 && v <= 32) h->cc_jumps = v; }
  // opt in to large dynamic shared memory (227 KB per CTA on sm_100, static part included)
  {
    auto optin = [&](const void* fn, size_t want_total) -> cudaError_t {
      cudaFuncAttributes fa;
      cudaError_t r = cudaFuncGetAttributes(&fa, fn);
      if (r != cudaSuccess) return r;
      return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(want_total - fa.sharedSizeBytes));
    };
    const size_t kMax = 227 * 1024;
    cudaError_t r = optin((const void*)k_local_graph2<64>, kMax);
    if (r == cudaSuccess) r = optin((const void*)k_local_graph2<128>, kMax);
    if (r == cudaSuccess) r = optin((const void*)k_local_graph2<256>, kMax);
    if (r == cudaSuccess) r = optin((const void*)k_local_graph_rows, kMax);
    if (r == cudaSuccess) r = optin((const void*)k_adj_fill, 200 * 1024);
    if (r == cudaSuccess) r = optin((const void*)k_rows_sort, 200 * 1024);
    if (r == cudaSuccess) r = optin((const void*)k_rows_fill, 200 * 1024);
    if (r != cudaSuccess) {
      g_create_error = std::string("kernel attribute setup failed (is this an sm_100 device?): ") + cudaGetErrorString(r);
      cudaGetLastError();
      delete h;
      return VGS_ERR_CUDA;
    }
  }
  *out = h;
  return VGS_OK;
}

void vgs_destroy(vgs_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DBuf* all[] = {&h->xyz_own, &h->labels_own, &h->keysA, &h->keysB, &h->valsA, &h->valsB, &h->hist, &h->tiles, &h->flags, &h->scan,
                 &h->small, &h->ustart, &h->ukey, &h->pos_unit, &h->rec, &h->key3, &h->center, &h->plainm, &h->tk, &h->tv,
                 &h->stencil, &h->adj_cnt, &h->adj_off, &h->adj_idx, &h->adj_code, &h->class_count, &h->class_list, &h->conn0_cnt,
                 &h->conn0_idx, &h->conn1_cnt, &h->conn1_idx, &h->attach, &h->parent, &h->root, &h->csize, &h->cminpt,
                 &h->labels_out, &h->tmp, &h->fallback, &h->uflags, &h->singles, &h->ckeysA, &h->ckeysB, &h->cvalsA, &h->cvalsB,
                 &h->cstart, &h->ckey, &h->cpos, &h->gridmin, &h->d_adj_cols, &h->d_pc_cols, &h->tb_slot, &h->tb_code5, &h->tb_first,
                 &h->tb_last, &h->bm_all, &h->bm_used, &h->idgrid, &h->row_len, &h->row_off, &h->rows, &h->long_rows,
                 &h->cstats, &h->conn_mask, &h->csr_off, &h->csr_idx, &h->used_list, &h->origin_state, &h->singles_dep, &h->singles_best, &h->closest_w};
  for (DBuf* b : all) b->release();
  auto& c = h->vc;
  DBuf* vcb[] = {&c.keysA, &c.keysB, &c.valsA, &c.valsB, &c.start, &c.key, &c.pos, &c.xyz, &c.key3, &c.plain, &c.ptvox, &c.nb, &c.nb_cnt, &c.nb_off, &c.nb_csr, &c.nrm, &c.mom, &c.bits,
                 &c.owner, &c.owner2, &c.dist, &c.claim, &c.ckA, &c.ckB, &c.cvA, &c.cvB, &c.cstart, &c.ckey, &c.cpos, &c.cell3, &c.best,
                 &c.flag, &c.rank, &c.seedv, &c.hc, &c.hn, &c.alive, &c.acc, &c.tk, &c.tv, &c.tk2, &c.tv2};
  for (DBuf* b : vcb) b->release();
  for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->kev) if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < vgs_context::N_AUX; i++) { if (h->aux[i]) cudaStreamDestroy(h->aux[i]); if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
}

// ---- pooled lifecycle: a parked handle keeps its device buffers, streams and events (cudaMalloc / cudaFree of the ~1.6 GB
//      working set of a 10 M-point cloud cost more than the segmentation itself) ----
namespace {
std::mutex g_pool_mu;
std::vector<vgs_handle> g_pool;
constexpr size_t POOL_MAX = 4;
void reset_for_reuse(vgs_handle h) {
  h->err.clear();
  h->launches = 0;
  h->d_xyz = nullptr; h->n = 0; h->stride = 3;
  h->d_labels = nullptr; h->max_label = 0;
  h->voxel_size = 0; h->depth = 0; h->n_finite = 0; h->n_voxels = 0; h->nu = 0; h->n_valid = 0;
  h->voxelized = h->have_graph = h->units_external = h->have_units = h->have_features = h->have_adj = h->have_segments = false;
  h->have_geometry = h->have_cluster_stats = h->have_csr = h->conn0_is_mask = false;
  h->last_voxels_min = std::numeric_limits<int>::min();
  h->n_used = h->n_adj = h->n_pairs = h->max_n = h->n_singles = h->n_attached = h->closest_rounds = 0;
  h->n_clusters_all = h->n_clusters_exp = 0; h->n_fallback = 0; h->csr_total = 0;
  h->d_perm = nullptr; h->vox_perm = nullptr; h->vox_table_valid = false;
  h->tm = vgs_timings{};
  h->tm_pending = 0; h->k_pending = 0;
  for (int i = 0; i < vgs_context::NK; i++) { h->k_ms[i] = 0.f; h->k_launches[i] = 0; }
}
}  // namespace

vgs_status vgs_acquire(vgs_handle* out, const vgs_config* cfg) {
  if (!out || !cfg) { g_create_error = "vgs_acquire: null argument"; return VGS_ERR_INVALID; }
  if (!cfg->stream) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); i++) {
      vgs_handle h = g_pool[i];
      if (h->device == cfg->device && h->mode == cfg->mode && h->leaf_order == cfg->leaf_order) {
        g_pool.erase(g_pool.begin() + (long)i);
        *out = h;
        return VGS_OK;
      }
    }
  }
  return vgs_create(out, cfg);
}

void vgs_release(vgs_handle h) {
  if (!h) return;
  static const bool no_pool = [] { const char* e = getenv("VGS_B200_NO_POOL"); return e && e[0] == '1'; }();   // A/B knob
  if (h->own_stream && !no_pool) {
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) == cudaSuccess) {
      reset_for_reuse(h);
      std::lock_guard<std::mutex> lock(g_pool_mu);
      if (g_pool.size() < POOL_MAX) { g_pool.push_back(h); return; }
    } else {
      cudaGetLastError();
    }
  }
  vgs_destroy(h);
}

void vgs_pool_trim(void) {
  std::vector<vgs_handle> all;
  {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    all.swap(g_pool);
  }
  for (vgs_handle h : all) vgs_destroy(h);
}

const char* vgs_last_error(vgs_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

vgs_status vgs_set_points(vgs_handle h, const float* xyz, int64_t n, int stride_bytes, int on_device) {
  if (!h) return VGS_ERR_INVALID;
  if (!xyz || n <= 0 || (stride_bytes != 12 && stride_bytes != 16)) return h->fail(VGS_ERR_INVALID, "vgs_set_points: bad argument");
  if (n >= (1ll << 31)) return h->fail(VGS_ERR_LIMIT, "vgs_set_points: n must be < 2^31 per handle (point indices are int, VS.h:163)");
  CK(cudaSetDevice(h->device));
  h->n = n; h->stride = stride_bytes / 4;
  h->voxelized = h->have_units = h->have_features = h->have_adj = h->have_segments = false;
  h->d_labels = nullptr;
  h->tm_pending = 0;
  h->k_pending = 0;
  for (int i = 0; i < vgs_context::NK; i++) { h->k_ms[i] = 0.f; h->k_launches[i] = 0; }
  h->tm = vgs_timings{};
  if (on_device) { h->d_xyz = xyz; }
  else {
    StageTimer t(h, &h->tm.h2d_ms, 0);
    CK(h->xyz_own.reserve((size_t)n * stride_bytes));
    CK(vgs_hostio::to_device(h->xyz_own.p, xyz, (size_t)n * stride_bytes, h->device, h->stream));
    h->d_xyz = h->xyz_own.as<float>();
    t.stop();
  }
  return VGS_OK;
}

vgs_status vgs_set_supervoxel_labels(vgs_handle h, const int32_t* labels, int32_t max_label, int on_device) {
  if (!h) return VGS_ERR_INVALID;
  if (h->mode != VGS_MODE_SVGS) return h->fail(VGS_ERR_STATE, "vgs_set_supervoxel_labels: handle is not in SVGS mode");
  if (!labels || h->n <= 0) return h->fail(VGS_ERR_STATE, "vgs_set_supervoxel_labels: call vgs_set_points first");
  CK(cudaSetDevice(h->device));
  if (on_device) h->d_labels = labels;
  else {
    CK(h->labels_own.reserve((size_t)h->n * 4));
    CK(vgs_hostio::to_device(h->labels_own.p, labels, (size_t)h->n * 4, h->device, h->stream));
    h->d_labels = h->labels_own.as<int32_t>();
  }
  h->max_label = max_label;
  h->units_external = false;
  h->have_units = h->have_features = h->have_adj = h->have_segments = false;
  return VGS_OK;
}

// ---- stage 1: keys, sort, voxel table (the epoch table / box / depth of stage 0 are installed in the handle).
//      gidx_w: the points are 16-byte records whose 4th word is the point's GLOBAL index (slab tiles). ----
static vgs_status voxelize_sorted(vgs_handle h, int gidx_w) {
  const int64_t n = h->n;
  {
    StageTimer t(h, &h->tm.voxelize_ms, 2);
    CK(h->keysA.reserve((size_t)n * 8)); CK(h->keysB.reserve((size_t)n * 8));
    CK(h->valsA.reserve((size_t)n * 4)); CK(h->valsB.reserve((size_t)n * 4));
    const int desc = h->leaf_order == VGS_LEAF_DESCENDING ? 1 : 0;
    const int nbits = 3 * h->depth + 1;
    int64_t nunits = 0;
    uint32_t* vs = nullptr;
    vgs_status s;
    std::optional<KTimer> kh;
    // keys of <= 32 bits (octrees up to depth 10: 307 m at 0.15 m) are sorted as 32-bit words
    auto run = [&](auto tag) -> vgs_status {
      using K = decltype(tag);
      KTimer kq(h, K_QUANTISE);
      LAUNCH(k_quantise<K>, (unsigned)cdiv(n, 256), 256, 0, h->d_xyz, h->stride, n, h->ep, (double)h->voxel_size, h->depth, desc,
             h->keysA.as<K>(), h->valsA.as<uint32_t>(), (uint32_t*)nullptr, gidx_w);
      kq.stop();
      K* ks;
      KTimer ksrt(h, K_SORT);
      vgs_status s_ = radix_sort<K>(h, n, nbits, &ks, &vs);
      if (s_) return s_;
      ksrt.stop();
      h->sort_key_bytes = (int)sizeof(K);
      kh.emplace(h, K_HEADS);
      // number of finite points = first sorted position whose key has the sentinel bit: count via head scan
      // (sentinel keys form at most one extra segment at the end)
      return build_units<K>(h, ks, n, &nunits);
    };
    s = nbits <= 32 ? run(uint32_t{}) : run(uint64_t{});
    if (s) return s;
    // voxel keys + occupied key range (the extent of the occupancy grids of the lattice searches)
    uint32_t* d_kmm = h->small.as<uint32_t>() + 208;
    {
      const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
      CK(cudaMemcpyAsync(d_kmm, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
      CK(h->key3.reserve((size_t)nunits * 12 + 16));
      LAUNCH(k_voxel_keys, (unsigned)cdiv(nunits, 256), 256, 0, h->ukey.as<uint64_t>(), nunits, h->depth,
             h->leaf_order == VGS_LEAF_DESCENDING ? 1 : 0, h->key3.as<uint32_t>(), d_kmm);
    }
    // peel the sentinel segment if present
    uint64_t lastkey = 0; uint32_t laststart = 0;
    CK(cudaMemcpyAsync(&lastkey, h->ukey.as<uint64_t>() + (nunits - 1), 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&laststart, h->ustart.as<uint32_t>() + (nunits - 1), 4, cudaMemcpyDeviceToHost, h->stream));
    kh->stop();
    CK(cudaMemcpyAsync(h->kminmax, d_kmm, sizeof(h->kminmax), cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    int64_t n_fin = n;
    if (lastkey >> (3 * h->depth)) { n_fin = laststart; nunits--; }
    h->n_finite = n_fin;
    h->n_voxels = nunits;
    h->vox_perm = vs; h->vox_table_valid = true;
    if (h->mode == VGS_MODE_VGS) {
      h->d_perm = vs;
      h->nu = nunits; h->n_valid = n_fin;
      h->have_units = true;
    }
    t.stop();
  }
  h->voxelized = true;
  return VGS_OK;
}

vgs_status vgs_voxelize(vgs_handle h, float voxel_size) {
  if (!h) return VGS_ERR_INVALID;
  if (!h->d_xyz) return h->fail(VGS_ERR_STATE, "vgs_voxelize: call vgs_set_points first");
  if (!(voxel_size > 0)) return h->fail(VGS_ERR_INVALID, "vgs_voxelize: voxel_size must be > 0");
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n;
  h->voxel_size = voxel_size;
  h->voxelized = false;
  h->vox_table_valid = false;
  h->have_units = h->have_features = h->have_adj = h->have_segments = false;   // the sort buffers are shared
  h->units_external = false;
  CK(h->small.reserve(4096));

  // ---- stage 0: PCL dynamic bounding box (origin) ----
  OriginBuilder ob;
  ob.begin((double)voxel_size);
  static const bool host_origin = [] { const char* e = getenv("VGS_B200_HOST_ORIGIN"); return e && e[0] == '1'; }();   // A/B knob: one host round trip per epoch
  {
    StageTimer t(h, &h->tm.origin_ms, 1);
    KTimer kt(h, K_ORIGIN);
    if (host_origin) {
      int64_t cursor = 0;
      while (true) {
        int64_t idx;
        float p[3];
        vgs_status s = find_next(h, cursor, ob.st, &idx, p);
        if (s) return s;
        if (idx >= n) break;
        if (const char* why = ob.add(idx, p)) return h->fail(VGS_ERR_LIMIT, std::string("vgs_voxelize: ") + why);
        cursor = idx + 1;
      }
    } else {
      // the growth loop runs on the device (k_origin_scan / k_origin_adopt rounds, state in device memory): the host
      // enqueues a batch of rounds and reads the state once; a typical cloud is done after 6-9 rounds
      CK(h->origin_state.reserve(sizeof(OriginState) + 64));
      OriginState* d_st = h->origin_state.as<OriginState>();
      LAUNCH(k_origin_init, 1, 1, 0, d_st, (double)voxel_size, (long long)n);
      OriginState hs;
      for (int batch = 0;; batch++) {
        const int rounds = batch == 0 ? 10 : 16;
        for (int r = 0; r < rounds; r++) {
          LAUNCH(k_origin_scan, 148 * 8, 256, 0, h->d_xyz, h->stride, d_st);
          LAUNCH(k_origin_adopt, 1, 1, 0, h->d_xyz, h->stride, d_st);
        }
        CK(cudaMemcpyAsync(&hs, d_st, sizeof(OriginState), cudaMemcpyDeviceToHost, h->stream));
        CK(stream_wait(h->stream));
        if (hs.done) break;
        if (batch > 64) return h->fail(VGS_ERR_LIMIT, "vgs_voxelize: the bounding-box growth loop did not finish");
      }
      if (hs.error == 1) return h->fail(VGS_ERR_LIMIT, "vgs_voxelize: octree depth > 21 bits per axis (extent / voxel_size too large)");
      if (hs.error == 2) return h->fail(VGS_ERR_LIMIT, "vgs_voxelize: too many bounding-box growth epochs");
      // the device state in the host builder's terms (shifts of the epoch keys, installation into the handle)
      ob.st.defined = hs.defined != 0;
      ob.st.depth = hs.depth;
      for (int a = 0; a < 3; a++) { ob.st.mn[a] = hs.mn[a]; ob.st.mx[a] = hs.mx[a]; }
      for (int k = 0; k < hs.n_events; k++) ob.st.events.push_back(OctState::Ev{hs.ev_lowered[k], hs.ev_depth_old[k]});
      ob.ep.n = hs.n_epochs;
      for (int e = 0; e < hs.n_epochs; e++) {
        ob.ep.viol[e] = hs.viol[e];
        for (int a = 0; a < 3; a++) ob.ep.mn[e][a] = hs.ep_mn[e][a];
        ob.events_before.push_back((size_t)hs.events_before[e]);
      }
    }
    if (!ob.st.defined) return h->fail(VGS_ERR_INVALID, "vgs_voxelize: no finite point in the cloud");
    ob.finish();
    kt.stop();
    t.stop();
  }
  ob.install(h);
  return voxelize_sorted(h, 0);
}

vgs_status vgs_get_bounding_box(vgs_handle h, double out6[6]) {
  if (!h || !out6) return VGS_ERR_INVALID;
  if (!h->voxelized) return h->fail(VGS_ERR_STATE, "vgs_get_bounding_box: call vgs_voxelize first");
  for (int a = 0; a < 3; a++) { out6[a] = h->box.mn[a]; out6[3 + a] = h->box.mx[a]; }
  return VGS_OK;
}

vgs_status vgs_voxel_count(vgs_handle h, int64_t* nv) {
  if (!h || !nv) return VGS_ERR_INVALID;
  if (!h->voxelized) return h->fail(VGS_ERR_STATE, "vgs_voxel_count: call vgs_voxelize first");
  *nv = h->n_voxels;
  return VGS_OK;
}

static vgs_status build_svgs_units(vgs_handle h) {
  if (!h->d_labels) return h->fail(VGS_ERR_STATE, "SVGS: call vgs_set_supervoxel_labels first");
  h->vox_table_valid = false;     // the shared sort buffers and the unit table are rebuilt over the labels
  const int64_t n = h->n;
  int32_t ml = h->max_label;
  if (ml <= 0) ml = std::numeric_limits<int32_t>::max();
  CK(h->keysA.reserve((size_t)n * 8)); CK(h->keysB.reserve((size_t)n * 8));
  CK(h->valsA.reserve((size_t)n * 4)); CK(h->valsB.reserve((size_t)n * 4));
  LAUNCH(k_label_keys, (unsigned)cdiv(n, 256), 256, 0, h->d_labels, h->d_xyz, h->stride, n, ml, h->keysA.as<uint64_t>(), h->valsA.as<uint32_t>());
  uint64_t* ks; uint32_t* vs;
  int label_bits = 1;
  while (label_bits < 32 && ((uint32_t)ml >> label_bits)) label_bits++;      // keys are <= ml: only its bits are sorted
  vgs_status s = radix_sort<uint64_t>(h, n, label_bits, &ks, &vs);
  if (s) return s;
  int64_t nunits = 0;
  s = build_units<uint64_t>(h, ks, n, &nunits);
  if (s) return s;
  uint64_t lastkey = 0; uint32_t laststart = 0;
  CK(cudaMemcpyAsync(&lastkey, h->ukey.as<uint64_t>() + (nunits - 1), 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&laststart, h->ustart.as<uint32_t>() + (nunits - 1), 4, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  int64_t nval = n;
  if (lastkey == (uint64_t)(uint32_t)ml) { nval = laststart; nunits--; }   // the segment of the dropped labels
  h->d_perm = vs;
  h->nu = nunits; h->n_valid = nval;
  h->have_units = true;
  return VGS_OK;
}

vgs_status vgs_make_supervoxels_grid(vgs_handle h, float seed_size) {
  if (!h) return VGS_ERR_INVALID;
  if (h->mode != VGS_MODE_SVGS) return h->fail(VGS_ERR_STATE, "vgs_make_supervoxels_grid: handle is not in SVGS mode");
  if (!h->voxelized) return h->fail(VGS_ERR_STATE, "vgs_make_supervoxels_grid: call vgs_voxelize first (the grid is anchored at the octree origin)");
  if (!(seed_size > 0)) return h->fail(VGS_ERR_INVALID, "vgs_make_supervoxels_grid: seed_size must be > 0");
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n;
  h->vox_table_valid = false;
  CK(h->keysA.reserve((size_t)n * 8)); CK(h->keysB.reserve((size_t)n * 8));
  CK(h->valsA.reserve((size_t)n * 4)); CK(h->valsB.reserve((size_t)n * 4));
  LAUNCH(k_seed_cell_keys, (unsigned)cdiv(n, 256), 256, 0, h->d_xyz, h->stride, n, h->box.mn[0], h->box.mn[1], h->box.mn[2],
         (double)seed_size, h->keysA.as<uint64_t>(), h->valsA.as<uint32_t>());
  uint64_t* ks; uint32_t* vs;
  vgs_status s = radix_sort<uint64_t>(h, n, 64, &ks, &vs);
  if (s) return s;
  int64_t nunits = 0;
  s = build_units<uint64_t>(h, ks, n, &nunits);
  if (s) return s;
  uint64_t lastkey = 0; uint32_t laststart = 0;
  CK(cudaMemcpyAsync(&lastkey, h->ukey.as<uint64_t>() + (nunits - 1), 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&laststart, h->ustart.as<uint32_t>() + (nunits - 1), 4, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  int64_t nval = n;
  if (lastkey >> 63) { nval = laststart; nunits--; }
  h->d_perm = vs;
  h->nu = nunits; h->n_valid = nval;
  h->have_units = true;
  h->units_external = true;
  h->have_features = h->have_adj = h->have_segments = false;
  return VGS_OK;
}

// createSupervoxels (SV.h:245-284): pcl::SupervoxelClustering(voxel_resolution, seed_resolution) + extract +
// refineSupervoxels(k), restated as data-parallel kernels (vgs_vccs.cuh; CPU restatement oracle/vccs_oracle.cpp,
// schedule 1).  Installs the per-point labels and getMaxLabel() exactly as vgs_set_supervoxel_labels would.
vgs_status vgs_make_supervoxels_vccs(vgs_handle h, float seed_resolution, float color_importance, float spatial_importance,
                                     float normal_importance, int refine_iterations) {
  if (!h) return VGS_ERR_INVALID;
  if (h->mode != VGS_MODE_SVGS) return h->fail(VGS_ERR_STATE, "vgs_make_supervoxels_vccs: handle is not in SVGS mode");
  if (!h->voxelized) return h->fail(VGS_ERR_STATE, "vgs_make_supervoxels_vccs: call vgs_voxelize first (voxel_resolution = its voxel size)");
  if (!(seed_resolution > 0) || refine_iterations < 0) return h->fail(VGS_ERR_INVALID, "vgs_make_supervoxels_vccs: bad seed_resolution / refine_iterations");
  CK(cudaSetDevice(h->device));
  auto& c = h->vc;
  const int64_t n = h->n;
  const int desc = h->leaf_order == VGS_LEAF_DESCENDING ? 1 : 0;
  KTimer kvccs(h, K_VCCS);
  // --- voxel table: the one vgs_voxelize left behind (sorted point indices, voxel starts and keys) while it is intact;
  //     the shared sort buffers are reused by the unit builders, so a later call sorts the points again ---
  int64_t V = 0;
  const uint32_t* vs = nullptr;
  const uint32_t* vstart = nullptr;
  const uint64_t* vkeys = nullptr;
  vgs_status s;
  if (h->vox_table_valid) {
    V = h->n_voxels; vs = h->vox_perm; vstart = h->ustart.as<uint32_t>(); vkeys = h->ukey.as<uint64_t>();
  } else {
    CK(c.keysA.reserve((size_t)n * 8)); CK(c.keysB.reserve((size_t)n * 8));
    CK(c.valsA.reserve((size_t)n * 4)); CK(c.valsB.reserve((size_t)n * 4));
    LAUNCH(k_quantise<uint64_t>, (unsigned)cdiv(n, 256), 256, 0, h->d_xyz, h->stride, n, h->ep, (double)h->voxel_size, h->depth, desc,
           c.keysA.as<uint64_t>(), c.valsA.as<uint32_t>(), (uint32_t*)nullptr);
    uint64_t* ks; uint32_t* vs_;
    s = radix_sort<uint64_t>(h, n, 3 * h->depth + 1, &ks, &vs_, c.keysA.as<uint64_t>(), c.keysB.as<uint64_t>(), c.valsA.as<uint32_t>(),
                             c.valsB.as<uint32_t>());
    if (s) return s;
    s = build_units<uint64_t>(h, ks, n, &V, &c.start, &c.key, &c.pos);
    if (s) return s;
    {
      uint64_t lastkey = 0;
      CK(cudaMemcpyAsync(&lastkey, c.key.as<uint64_t>() + (V - 1), 8, cudaMemcpyDeviceToHost, h->stream));
      CK(stream_wait(h->stream));
      if (lastkey >> (3 * h->depth)) V--;   // the segment of the non-finite points
    }
    if (V != h->n_voxels) return h->fail(VGS_ERR_STATE, "vgs_make_supervoxels_vccs: voxel table differs from vgs_voxelize's");
    vs = vs_; vstart = c.start.as<uint32_t>(); vkeys = c.key.as<uint64_t>();
  }
  CK(c.xyz.reserve((size_t)V * 12 + 16)); CK(c.key3.reserve((size_t)V * 12 + 16)); CK(c.plain.reserve((size_t)V * 8 + 16));
  CK(c.ptvox.reserve((size_t)n * 4)); CK(c.nb.reserve((size_t)V * 27 * 4)); CK(c.nrm.reserve((size_t)V * 12 + 16));
  CK(c.owner.reserve((size_t)V * 4)); CK(c.owner2.reserve((size_t)V * 4)); CK(c.dist.reserve((size_t)V * 4)); CK(c.claim.reserve((size_t)V * 4));
  CK(cudaMemsetAsync(c.ptvox.p, 0xff, (size_t)n * 4, h->stream));
  LAUNCH(k_vccs_voxels, (unsigned)cdiv(V, 256), 256, 0, h->d_xyz, h->stride, vs, vstart, vkeys, V, h->depth, desc,
         c.xyz.as<float>(), c.key3.as<uint32_t>(), c.plain.as<uint64_t>(), c.ptvox.as<int32_t>());
  uint64_t capacity = 64;
  while (capacity < (uint64_t)V * 2) capacity <<= 1;
  const uint64_t vmask = capacity - 1;
  CK(c.tk.reserve(capacity * 8)); CK(c.tv.reserve(capacity * 4));
  CK(cudaMemsetAsync(c.tk.p, 0xff, capacity * 8, h->stream));
  LAUNCH(k_hash_insert, (unsigned)cdiv(V, 256), 256, 0, c.plain.as<uint64_t>(), V, c.tk.as<unsigned long long>(), c.tv.as<uint32_t>(), vmask);
  // occupancy bits of the voxel lattice over the occupied key range (+ the largest search reach): the cube searches below
  // test a bit before they probe the hash table
  VccsBits ob{};
  {
    const int margin = 34;        // reseeding looks up to 32 cells away
    ob.g.x0 = (int)h->kminmax[0] - margin; ob.g.y0 = (int)h->kminmax[1] - margin; ob.g.z0 = (int)h->kminmax[2] - margin;
    const uint64_t nx = (uint64_t)(h->kminmax[3] - h->kminmax[0] + 1) + 2 * margin;
    ob.g.ny = (uint32_t)(h->kminmax[4] - h->kminmax[1] + 1) + 2 * margin;
    ob.g.nz = (uint32_t)(h->kminmax[5] - h->kminmax[2] + 1) + 2 * margin;
    ob.nx = (uint32_t)nx;
    const uint64_t nbits = nx * ob.g.ny * ob.g.nz;
    if (nbits <= ((uint64_t)1 << 33)) {
      const size_t bytes = (size_t)(nbits / 8) + 64;
      CK(c.bits.reserve(bytes));
      CK(cudaMemsetAsync(c.bits.p, 0, bytes, h->stream));
      LAUNCH(k_vccs_bits_set, (unsigned)cdiv(V, 256), 256, 0, c.key3.as<uint32_t>(), V, ob.g, c.bits.as<uint32_t>());
      ob.bits = c.bits.as<uint32_t>();
    }
  }
  LAUNCH(k_vccs_neighbours, (unsigned)cdiv(V * 27, 256), 256, 0, c.key3.as<uint32_t>(), V, h->depth, c.tk.as<unsigned long long>(),
         c.tv.as<uint32_t>(), vmask, ob, c.nb.as<int32_t>());
  // the existing neighbours as a CSR (slot order kept): what every round below streams
  CK(c.nb_cnt.reserve((size_t)(V + 1) * 4 + 16)); CK(c.nb_off.reserve((size_t)(V + 1) * 4 + 16)); CK(c.nb_csr.reserve((size_t)V * 27 * 4 + 16));
  LAUNCH(k_vccs_nb_count, (unsigned)cdiv(V + 1, 256), 256, 0, c.nb.as<int32_t>(), V, c.nb_cnt.as<uint32_t>());
  s = scan_u32(h, c.nb_cnt.as<uint32_t>(), c.nb_off.as<uint32_t>(), V + 1, nullptr);
  if (s) return s;
  LAUNCH(k_vccs_nb_compact, (unsigned)cdiv(V, 256), 256, 0, c.nb.as<int32_t>(), V, c.nb_off.as<uint32_t>(), c.nb_csr.as<int32_t>());
  const uint32_t* nb_off = c.nb_off.as<uint32_t>();
  const int32_t* nb_csr = c.nb_csr.as<int32_t>();
  CK(c.mom.reserve((size_t)V * sizeof(VMom) + 16));
  LAUNCH(k_vccs_moments, (unsigned)cdiv(V, 128), 128, 0, V, c.xyz.as<float>(), nb_off, nb_csr, (const int32_t*)nullptr, c.mom.as<VMom>());
  LAUNCH(k_vccs_normals, (unsigned)cdiv(V, 128), 128, 0, V, c.xyz.as<float>(), nb_off, nb_csr, (const int32_t*)nullptr, c.mom.as<VMom>(),
         c.nrm.as<float>());

  // --- seeds ---
  const double seed_d = (double)seed_resolution;
  CK(c.ckA.reserve((size_t)V * 8)); CK(c.ckB.reserve((size_t)V * 8)); CK(c.cvA.reserve((size_t)V * 4)); CK(c.cvB.reserve((size_t)V * 4));
  CK(c.cell3.reserve((size_t)V * 12 + 16));
  LAUNCH(k_vccs_cell_keys, (unsigned)cdiv(V, 256), 256, 0, c.xyz.as<float>(), V, h->box.mn[0], h->box.mn[1], h->box.mn[2], seed_d,
         c.ckA.as<uint64_t>(), c.cvA.as<uint32_t>(), c.cell3.as<int32_t>());
  uint64_t* cks; uint32_t* cvs;
  // seed cells are counted from the octree origin: a coordinate is below side / seed + 1, the morton key needs 3 x its bits
  int cell_bits = 1;
  {
    const double side = std::ldexp((double)h->voxel_size, h->depth);
    const double cells = std::floor(side / seed_d) + 2.0;
    while (cell_bits < 21 && std::ldexp(1.0, cell_bits) < cells) cell_bits++;
  }
  s = radix_sort<uint64_t>(h, V, 3 * cell_bits, &cks, &cvs, c.ckA.as<uint64_t>(), c.ckB.as<uint64_t>(), c.cvA.as<uint32_t>(), c.cvB.as<uint32_t>());
  if (s) return s;
  int64_t NC = 0;
  s = build_units<uint64_t>(h, cks, V, &NC, &c.cstart, &c.ckey, &c.cpos);
  if (s) return s;
  uint64_t ccap = 64;
  while (ccap < (uint64_t)NC * 2) ccap <<= 1;
  const uint64_t cmask = ccap - 1;
  CK(c.tk2.reserve(ccap * 8)); CK(c.tv2.reserve(ccap * 4));
  CK(cudaMemsetAsync(c.tk2.p, 0xff, ccap * 8, h->stream));
  LAUNCH(k_hash_insert, (unsigned)cdiv(NC, 256), 256, 0, c.ckey.as<uint64_t>(), NC, c.tk2.as<unsigned long long>(), c.tv2.as<uint32_t>(), cmask);
  CK(c.best.reserve((size_t)NC * 8)); CK(c.flag.reserve((size_t)(NC + 1) * 4)); CK(c.rank.reserve((size_t)(NC + 1) * 4));
  CK(cudaMemsetAsync(c.best.p, 0xff, (size_t)NC * 8, h->stream));
  LAUNCH(k_vccs_seed_nearest, (unsigned)cdiv(V, 256), 256, 0, c.xyz.as<float>(), c.cell3.as<int32_t>(), V, h->box.mn[0], h->box.mn[1], h->box.mn[2],
         seed_d, c.tk2.as<unsigned long long>(), c.tv2.as<uint32_t>(), cmask, c.best.as<unsigned long long>());
  LAUNCH(k_vccs_reset, (unsigned)cdiv(V, 256), 256, 0, V, c.owner.as<int32_t>(), c.dist.as<float>(), c.claim.as<int32_t>());
  LAUNCH(k_vccs_seed_claim, (unsigned)cdiv(NC, 256), 256, 0, c.best.as<unsigned long long>(), NC, c.claim.as<int32_t>());
  const float voxel_res = h->voxel_size;
  const float search_radius = 0.5f * seed_resolution;
  const float min_points = 0.05f * (search_radius) * (search_radius) * 3.1415926536f / (voxel_res * voxel_res);
  // two centroids closer than r lie in cells at most floor(r / res) + 1 apart per axis
  const int reach = (int)std::floor((double)search_radius / (double)voxel_res * (1.0 + 1e-5)) + 1;
  if (reach > 40) return h->fail(VGS_ERR_LIMIT, "vgs_make_supervoxels_vccs: seed_resolution / voxel_resolution too large");
  LAUNCH(k_vccs_seed_filter, (unsigned)cdiv(NC * 32, 128), 128, 0, c.best.as<unsigned long long>(), c.claim.as<int32_t>(), NC, c.key3.as<uint32_t>(),
         c.xyz.as<float>(), h->depth, c.tk.as<unsigned long long>(), c.tv.as<uint32_t>(), vmask, ob, search_radius * search_radius, min_points, reach,
         c.flag.as<uint32_t>());
  unsigned long long Htot = 0;
  s = scan_u32(h, c.flag.as<uint32_t>(), c.rank.as<uint32_t>(), NC, &Htot);
  if (s) return s;
  const int64_t H = (int64_t)Htot;
  h->vccs_seeds = H;
  CK(h->labels_own.reserve((size_t)n * 4));
  if (H == 0) {
    CK(cudaMemsetAsync(h->labels_own.p, 0, (size_t)n * 4, h->stream));
    h->d_labels = h->labels_own.as<int32_t>(); h->max_label = 0;
    h->units_external = false;
    h->have_units = h->have_features = h->have_adj = h->have_segments = false;
    return h->fail(VGS_ERR_INVALID, "vgs_make_supervoxels_vccs: no seed survived (cloud too sparse for this seed_resolution)");
  }
  CK(c.hc.reserve((size_t)H * 12 + 16)); CK(c.hn.reserve((size_t)H * 12 + 16)); CK(c.alive.reserve((size_t)H + 16));
  CK(c.acc.reserve((size_t)H * 56 + 16)); CK(c.seedv.reserve((size_t)H * 4));
  CK(cudaMemsetAsync(c.alive.p, 0, (size_t)H, h->stream));
  LAUNCH(k_vccs_helpers, (unsigned)cdiv(NC, 256), 256, 0, c.best.as<unsigned long long>(), c.flag.as<uint32_t>(), c.rank.as<uint32_t>(), NC,
         c.xyz.as<float>(), c.nrm.as<float>(), c.hc.as<float>(), c.hn.as<float>(), c.alive.as<uint8_t>(), c.owner.as<int32_t>());

  // --- expandSupervoxels(depth): depth - 1 synchronous rounds, centroids after each ---
  const int depth = (int)(1.8f * seed_resolution / voxel_res);
  int32_t* own_a = c.owner.as<int32_t>();
  int32_t* own_b = c.owner2.as<int32_t>();
  auto expand = [&]() -> vgs_status {
    // centroid sums of the phase's starting owners (the seeds), then kept current by the rounds
    unsigned long long* d_acc = c.acc.as<unsigned long long>();
    unsigned long long* d_cnt = d_acc + (size_t)H * 6;                  // acc (6 x u64 per helper) and cnt live in one buffer
    CK(cudaMemsetAsync(c.acc.p, 0, (size_t)H * 56, h->stream));
    LAUNCH(k_vccs_accumulate, (unsigned)cdiv(cdiv(V, VCCS_ACC_RUN), 256), 256, 0, V, own_a, c.xyz.as<float>(), c.nrm.as<float>(), d_acc, d_cnt);
    for (int it = 1; it < depth; it++) {
      LAUNCH(k_vccs_expand, (unsigned)cdiv(V, 256), 256, 0, V, nb_off, nb_csr, own_a, own_b, c.dist.as<float>(), c.xyz.as<float>(),
             c.nrm.as<float>(), c.hc.as<float>(), c.hn.as<float>(), c.alive.as<uint8_t>(), seed_resolution, color_importance,
             spatial_importance, normal_importance, d_acc, d_cnt);
      std::swap(own_a, own_b);
      LAUNCH(k_vccs_centroids, (unsigned)cdiv(H, 256), 256, 0, H, d_acc, d_cnt, c.hc.as<float>(),
             c.hn.as<float>(), c.alive.as<uint8_t>());
    }
    return VGS_OK;
  };
  s = expand();
  if (s) return s;
  // --- refineSupervoxels(k): normals inside each supervoxel, reseed at the voxel nearest to the centroid, expand ---
  for (int it = 0; it < refine_iterations; it++) {
    LAUNCH(k_vccs_moments, (unsigned)cdiv(V, 128), 128, 0, V, c.xyz.as<float>(), nb_off, nb_csr, (const int32_t*)own_a, c.mom.as<VMom>());
    LAUNCH(k_vccs_normals, (unsigned)cdiv(V, 128), 128, 0, V, c.xyz.as<float>(), nb_off, nb_csr, (const int32_t*)own_a, c.mom.as<VMom>(),
           c.nrm.as<float>());
    LAUNCH(k_vccs_reseed, (unsigned)cdiv(H * 32, 128), 128, 0, H, c.hc.as<float>(), c.alive.as<uint8_t>(), h->box.mn[0], h->box.mn[1], h->box.mn[2],
           (double)voxel_res, h->depth, c.tk.as<unsigned long long>(), c.tv.as<uint32_t>(), vmask, ob, c.xyz.as<float>(), c.seedv.as<int32_t>());
    LAUNCH(k_vccs_reset, (unsigned)cdiv(V, 256), 256, 0, V, own_a, c.dist.as<float>(), c.claim.as<int32_t>());
    LAUNCH(k_vccs_reseed_claim, (unsigned)cdiv(H, 256), 256, 0, H, c.seedv.as<int32_t>(), c.alive.as<uint8_t>(), c.claim.as<int32_t>());
    LAUNCH(k_vccs_reseed_apply, (unsigned)cdiv(H, 256), 256, 0, H, c.seedv.as<int32_t>(), c.alive.as<uint8_t>(), c.claim.as<int32_t>(), own_a);
    s = expand();
    if (s) return s;
  }
  // --- getLabeledCloud / getMaxLabel, installed like vgs_set_supervoxel_labels ---
  int32_t* d_ml = reinterpret_cast<int32_t*>(h->small.as<unsigned long long>() + 60);
  CK(cudaMemsetAsync(d_ml, 0, 4, h->stream));
  LAUNCH(k_vccs_max_label, (unsigned)cdiv(H, 256), 256, 0, H, c.alive.as<uint8_t>(), d_ml);
  LAUNCH(k_vccs_point_labels, (unsigned)cdiv(n, 256), 256, 0, n, c.ptvox.as<int32_t>(), (const int32_t*)own_a, h->labels_own.as<int32_t>());
  kvccs.stop();
  int32_t ml = 0;
  CK(cudaMemcpyAsync(&ml, d_ml, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  h->d_labels = h->labels_own.as<int32_t>();
  h->max_label = ml;
  h->units_external = false;
  h->have_units = h->have_features = h->have_adj = h->have_segments = false;
  return VGS_OK;
}

// the labels the SVGS units are built from (set by the caller or made by vgs_make_supervoxels_vccs) and max_label
vgs_status vgs_get_supervoxel_labels(vgs_handle h, int32_t* labels, int32_t* max_label, int on_device) {
  if (!h) return VGS_ERR_INVALID;
  if (h->mode != VGS_MODE_SVGS || !h->d_labels) return h->fail(VGS_ERR_STATE, "vgs_get_supervoxel_labels: no supervoxel labels yet");
  CK(cudaSetDevice(h->device));
  if (labels) {
    CK(cudaMemcpyAsync(labels, h->d_labels, (size_t)h->n * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
  }
  if (max_label) *max_label = h->max_label;
  return VGS_OK;
}

vgs_status vgs_unit_count(vgs_handle h, int64_t* nunits) {
  if (!h || !nunits) return VGS_ERR_INVALID;
  if (!h->have_units) return h->fail(VGS_ERR_STATE, "vgs_unit_count: units not built yet");
  *nunits = h->nu;
  return VGS_OK;
}

vgs_status vgs_compute_features(vgs_handle h, int points_min) {
  if (!h) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  StageTimer t(h, &h->tm.features_ms, 3);
  if (h->mode == VGS_MODE_SVGS && !(h->units_external && h->have_units)) {
    KTimer ku(h, K_SVGS_UNITS);
    vgs_status s = build_svgs_units(h);
    if (s) return s;
    ku.stop();
  }
  if (!h->have_units) return h->fail(VGS_ERR_STATE, "vgs_compute_features: call vgs_voxelize first (test:54-62)");
  h->points_min = points_min;
  h->have_features = h->have_adj = h->have_segments = false;
  const int64_t nu = h->nu;
  if (nu <= 0) return h->fail(VGS_ERR_INVALID, "vgs_compute_features: no units");
  CK(h->rec.reserve((size_t)nu * REC_FLOATS * 4));
  unsigned long long* d_used = h->small.as<unsigned long long>() + 16;
  CK(cudaMemsetAsync(d_used, 0, 8, h->stream));
  CK(h->uflags.reserve((size_t)nu + 16));
  KTimer kf(h, K_FEATURES);
  LAUNCH(k_features, (unsigned)cdiv(nu, FEAT_THREADS), FEAT_THREADS, 0, h->d_xyz, h->stride, h->d_perm, h->ustart.as<uint32_t>(), nu, points_min,
         h->mode == VGS_MODE_SVGS ? 1 : 0, h->rec.as<float>(), h->uflags.as<uint8_t>(), d_used);
  kf.stop();
  if (h->mode == VGS_MODE_SVGS) {   // VGS: the count arrives with the adjacency totals (one host round trip less)
    unsigned long long used = 0;
    CK(cudaMemcpyAsync(&used, d_used, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    h->n_used = (int64_t)used;
  }
  h->have_features = true;
  t.stop();
  return VGS_OK;
}

static vgs_status ensure_geometry(vgs_handle h) {
  if (h->mode != VGS_MODE_VGS || !h->voxelized) return h->fail(VGS_ERR_STATE, "voxel geometry needs vgs_voxelize in VGS mode");
  if (h->have_geometry) return VGS_OK;
  const int64_t nu = h->nu;
  CK(h->center.reserve((size_t)nu * 12 + 16));
  float res_f = (float)(double)h->voxel_size;   // setVoxelSize narrows the resolution to float (VS.h:127)
  LAUNCH(k_voxel_centers, (unsigned)cdiv(nu, 256), 256, 0, h->key3.as<uint32_t>(), nu, res_f, h->bb_f[0], h->bb_f[1], h->bb_f[2],
         h->center.as<float>());
  h->have_geometry = true;
  return VGS_OK;
}

vgs_status vgs_set_bounding_box(vgs_handle h, const double in6[6]) {
  if (!h || !in6) return VGS_ERR_INVALID;
  if (!h->voxelized) return h->fail(VGS_ERR_STATE, "vgs_set_bounding_box: call vgs_voxelize first (test:56-57)");
  for (int a = 0; a < 6; a++) h->bb_f[a] = (float)in6[a];
  h->have_geometry = false;
  h->have_adj = h->have_segments = false;
  return VGS_OK;
}

vgs_status vgs_get_voxel_centers(vgs_handle h, float* xyz) {
  if (!h || !xyz) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  { vgs_status s = ensure_geometry(h); if (s) return s; }
  CK(cudaMemcpyAsync(xyz, h->center.p, (size_t)h->nu * 12, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  return VGS_OK;
}

vgs_status vgs_get_unit_adjacency(vgs_handle h, int64_t unit, int32_t* ids, int cap, int* n) {
  if (!h || !n) return VGS_ERR_INVALID;
  if (!h->have_adj) return h->fail(VGS_ERR_STATE, "vgs_get_unit_adjacency: call vgs_find_adjacency first");
  if (unit < 0 || unit >= h->nu) return h->fail(VGS_ERR_INVALID, "vgs_get_unit_adjacency: unit id out of range");
  CK(cudaSetDevice(h->device));
  uint32_t off[2] = {0, 0};
  CK(cudaMemcpyAsync(off, h->adj_off.as<uint32_t>() + unit, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  const int len = (int)(off[1] - off[0]);
  *n = len;
  if (ids && cap > 0 && len > 0) {
    CK(cudaMemcpyAsync(ids, h->adj_idx.as<int32_t>() + off[0], (size_t)std::min(len, cap) * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
  }
  return VGS_OK;
}

// Host tables of the lattice searches, rebuilt (and uploaded) only when (voxel_size, graph_size, float-noise bound) change:
// radius stencil sorted by integer distance class, its columns, slot tables, and the pair stencil (differences of two
// stencil offsets, lexicographically positive half) as columns.
static vgs_status build_lattice_tables(vgs_handle h, float graph_size, double noise) {
  if (h->st_vs == h->voxel_size && h->st_gs == graph_size && h->st_noise == noise) return VGS_OK;
  std::vector<int4> st = make_stencil(h->voxel_size, graph_size, noise);
  for (int4& o : st) o.w = o.x * o.x + o.y * o.y + o.z * o.z;
  std::sort(st.begin(), st.end(), [](const int4& a, const int4& b) {
    if (a.w != b.w) return a.w < b.w;
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z < b.z;
  });
  int rho = 0;
  for (const int4& o : st) rho = std::max(rho, std::max(std::abs(o.x), std::max(std::abs(o.y), std::abs(o.z))));
  if (rho > 15) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: graph_size / voxel_size too large for the stencil search");
  const int nst = (int)st.size();
  if (nst >= 65535 || adj_fill_smem(nst, 2 * rho + 1) > 200 * 1024) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: graph_size / voxel_size too large for the stencil search");
  auto columns = [](const std::vector<int4>& offs, int reach) {
    std::vector<int4> cols;
    for (const int4& o : offs) {
      bool found = false;
      for (int4& c : cols) if (c.x == o.x && c.y == o.y) { c.z |= 1 << (o.z + reach); found = true; break; }
      if (!found) cols.push_back(make_int4(o.x, o.y, 1 << (o.z + reach), 0));
    }
    return cols;
  };
  h->stencil_host = st;
  h->st_rho = rho;
  h->adj_cols_host = columns(st, rho);
  // slot tables
  const int S = 2 * rho + 1;
  std::vector<uint32_t> code_lut((size_t)S * S * S);
  std::vector<uint16_t> code5(nst), cf(nst), cl(nst);
  for (int x = 0; x < S; x++)
    for (int y = 0; y < S; y++)
      for (int z = 0; z < S; z++) code_lut[(size_t)(x * S + y) * S + z] = 0xffffu | ((uint32_t)x << 16) | ((uint32_t)y << 21) | ((uint32_t)z << 26);
  // distance classes are only trusted when they are far apart compared with the float noise of dist2
  const double res = (double)h->voxel_size;
  const bool classes_ok = noise * 4.0 < res * res;
  for (int s = 0; s < nst; s++) {
    uint32_t& e = code_lut[(size_t)((st[s].x + rho) * S + (st[s].y + rho)) * S + (st[s].z + rho)];
    e = (e & 0xffff0000u) | (uint32_t)s;
    code5[s] = (uint16_t)pack5(st[s].x + rho, st[s].y + rho, st[s].z + rho);
  }
  for (int s = 0; s < nst;) {
    int e = s;
    while (e + 1 < nst && (classes_ok ? st[e + 1].w == st[s].w : true)) e++;
    for (int t = s; t <= e; t++) { cf[t] = (uint16_t)s; cl[t] = (uint16_t)e; }
    s = e + 1;
  }
  // pair stencil
  const int r2 = 2 * rho, S2 = 2 * r2 + 1;
  h->st_r2 = r2;
  h->rows_ok = rho <= 5 && S * S * S <= 65535;
  h->pc_cols_host.clear();
  if (h->rows_ok) {
    std::vector<char> seen((size_t)S2 * S2 * S2, 0);
    std::vector<int4> st2;
    for (const int4& p1 : st)
      for (const int4& p2 : st) {
        int dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
        if (!(dx > 0 || (dx == 0 && (dy > 0 || (dy == 0 && dz > 0))))) continue;
        int code = ((dx + r2) * S2 + (dy + r2)) * S2 + (dz + r2);
        if (seen[code]) continue;
        seen[code] = 1;
        st2.push_back(make_int4(dx, dy, dz, 0));
      }
    h->pc_cols_host = columns(st2, r2);
    h->max_row_len = (int64_t)st2.size();
  }
  h->lbits = rho <= 3 ? 3 : 4;
  h->mwords = (S * S * S + 31) / 32;
  // upload (synchronous with respect to the host vectors: the copies are made from pageable memory)
  CK(h->d_adj_cols.reserve(h->adj_cols_host.size() * sizeof(int4) + 16));
  CK(cudaMemcpyAsync(h->d_adj_cols.p, h->adj_cols_host.data(), h->adj_cols_host.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  CK(h->d_pc_cols.reserve(h->pc_cols_host.size() * sizeof(int4) + 16));
  if (!h->pc_cols_host.empty())
    CK(cudaMemcpyAsync(h->d_pc_cols.p, h->pc_cols_host.data(), h->pc_cols_host.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  CK(h->tb_slot.reserve(code_lut.size() * 4 + 16)); CK(h->tb_code5.reserve((size_t)nst * 2 + 16));
  CK(h->tb_first.reserve((size_t)nst * 2 + 16)); CK(h->tb_last.reserve((size_t)nst * 2 + 16));
  CK(cudaMemcpyAsync(h->tb_slot.p, code_lut.data(), code_lut.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->tb_code5.p, code5.data(), (size_t)nst * 2, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->tb_first.p, cf.data(), (size_t)nst * 2, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->tb_last.p, cl.data(), (size_t)nst * 2, cudaMemcpyHostToDevice, h->stream));
  CK(stream_wait(h->stream));     // the host vectors go out of scope
  h->st_vs = h->voxel_size; h->st_gs = graph_size; h->st_noise = noise;
  return VGS_OK;
}

vgs_status vgs_find_adjacency(vgs_handle h, float graph_size) {
  if (!h) return VGS_ERR_INVALID;
  if (!h->have_features) return h->fail(VGS_ERR_STATE, "vgs_find_adjacency: call vgs_compute_features first");
  if (!(graph_size > 0)) return h->fail(VGS_ERR_INVALID, "vgs_find_adjacency: graph_size must be > 0");
  CK(cudaSetDevice(h->device));
  StageTimer t(h, &h->tm.adjacency_ms, 4);
  if (h->mode == VGS_MODE_SVGS) {
    // findAllSupervoxelNeighbors SV.h:1477-1521: FLANN radius search over supervoxel CENTROIDS
    // (arbitrary floats).  A uniform grid of cell size 1.01 r only prunes candidates; the float test
    // and the (dist2, id) order are FLANN's.
    h->graph_size = graph_size;
    h->have_adj = h->have_segments = false;
    h->stencil_host.clear();
    const int64_t nu = h->nu;
    const float cell = graph_size * 1.01f;
    KTimer ksa(h, K_SVGS_ADJ);
    CK(h->gridmin.reserve(64));
    const uint32_t init[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
    CK(cudaMemcpyAsync(h->gridmin.p, init, 12, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_centroid_min, (unsigned)std::min<int64_t>(cdiv(nu, 256), 1024), 256, 0, h->rec.as<float>(), nu, h->gridmin.as<uint32_t>());
    CK(h->ckeysA.reserve((size_t)nu * 8)); CK(h->ckeysB.reserve((size_t)nu * 8));
    CK(h->cvalsA.reserve((size_t)nu * 4)); CK(h->cvalsB.reserve((size_t)nu * 4));
    LAUNCH(k_cell_keys, (unsigned)cdiv(nu, 256), 256, 0, h->rec.as<float>(), nu, h->gridmin.as<uint32_t>(), cell,
           h->ckeysA.as<uint64_t>(), h->cvalsA.as<uint32_t>());
    uint64_t* ks; uint32_t* vs;
    // cells are counted from the smallest centroid, which lies inside the octree box: coordinates are below side / cell + 2
    int cell_bits = 1;
    {
      const double side = std::ldexp((double)h->voxel_size, h->depth);
      const double cells = std::floor(side / (double)cell) + 3.0;
      while (cell_bits < 21 && std::ldexp(1.0, cell_bits) < cells) cell_bits++;
    }
    vgs_status s = radix_sort<uint64_t>(h, nu, 3 * cell_bits, &ks, &vs, h->ckeysA.as<uint64_t>(), h->ckeysB.as<uint64_t>(), h->cvalsA.as<uint32_t>(),
                              h->cvalsB.as<uint32_t>());
    if (s) return s;
    int64_t ncells = 0;
    s = build_units<uint64_t>(h, ks, nu, &ncells, &h->cstart, &h->ckey, &h->cpos);
    if (s) return s;
    uint64_t capacity = 64;
    while (capacity < (uint64_t)ncells * 2) capacity <<= 1;
    h->hmask = capacity - 1;
    CK(h->tk.reserve(capacity * 8)); CK(h->tv.reserve(capacity * 4));
    CK(cudaMemsetAsync(h->tk.p, 0xff, capacity * 8, h->stream));
    LAUNCH(k_hash_insert, (unsigned)cdiv(ncells, 256), 256, 0, h->ckey.as<uint64_t>(), ncells, h->tk.as<unsigned long long>(),
           h->tv.as<uint32_t>(), h->hmask);
    double r = (double)graph_size;
    float r2 = (float)(r * r);
    const int cap = 256, wpb = 4;
    size_t smem = (size_t)wpb * cap * 8;
    CK(h->adj_cnt.reserve((size_t)(nu + 1) * 4)); CK(h->adj_off.reserve((size_t)(nu + 1) * 4));
    unsigned long long* d_over = h->small.as<unsigned long long>() + 56;
    CK(cudaMemsetAsync(d_over, 0, 8, h->stream));
    LAUNCH(k_adjacency_svgs, (unsigned)cdiv(nu, wpb), wpb * 32, smem, h->rec.as<float>(), nu, h->gridmin.as<uint32_t>(), cell,
           h->cstart.as<uint32_t>(), vs, h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(), h->hmask, r2, 0,
           h->adj_cnt.as<uint32_t>(), (const uint32_t*)nullptr, (int32_t*)nullptr, cap, d_over);
    unsigned long long total = 0, over = 0;
    s = scan_u32(h, h->adj_cnt.as<uint32_t>(), h->adj_off.as<uint32_t>(), nu, &total);
    if (s) return s;
    CK(cudaMemcpyAsync(&over, d_over, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    if (over) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: a supervoxel has more than 255 neighbours within graph_size");
    uint32_t tot32 = (uint32_t)total;
    CK(cudaMemcpyAsync(h->adj_off.as<uint32_t>() + nu, &tot32, 4, cudaMemcpyHostToDevice, h->stream));
    h->n_adj = (int64_t)total;
    CK(h->adj_idx.reserve((size_t)total * 4 + 16));
    LAUNCH(k_adjacency_svgs, (unsigned)cdiv(nu, wpb), wpb * 32, smem, h->rec.as<float>(), nu, h->gridmin.as<uint32_t>(), cell,
           h->cstart.as<uint32_t>(), vs, h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(), h->hmask, r2, 1,
           h->adj_cnt.as<uint32_t>(), h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), cap, d_over);
    ksa.stop();
    CK(stream_wait(h->stream));
    h->have_adj = true;
    t.stop();
    return VGS_OK;
  }
  // ---- VGS: findAllVoxelAdjacency (VS.h:223-265) on the voxel lattice ----
  h->graph_size = graph_size;
  h->have_adj = h->have_segments = false;
  const int64_t nu = h->nu;
  const float res_f = (float)(double)h->voxel_size;
  // float noise of a squared centre distance: centres carry <= 1/2 ulp of the largest coordinate each
  double maxc = 0;
  for (int a = 0; a < 3; a++) {
    maxc = std::max(maxc, std::fabs(((double)h->kminmax[a] + 0.5) * res_f + h->bb_f[a]));
    maxc = std::max(maxc, std::fabs(((double)h->kminmax[3 + a] + 0.5) * res_f + h->bb_f[a]));
  }
  int ex = 0;
  std::frexp(maxc, &ex);
  const double ulp = std::ldexp(1.0, ex - 24);
  const double noise = maxc > 0 ? 12.0 * (double)graph_size * ulp : 0.0;
  { vgs_status s_ = build_lattice_tables(h, graph_size, noise); if (s_) return s_; }
  const int nst = (int)h->stencil_host.size(), rho = h->st_rho, r2c = h->st_r2;
  // hash table: plain morton -> voxel id
  uint64_t capacity = 64;
  while (capacity < (uint64_t)nu * 2) capacity <<= 1;
  h->hmask = capacity - 1;
  CK(h->plainm.reserve((size_t)nu * 8));
  CK(h->tk.reserve(capacity * 8)); CK(h->tv.reserve(capacity * 4));
  KTimer khash(h, K_HASH);
  CK(cudaMemsetAsync(h->tk.p, 0xff, capacity * 8, h->stream));
  LAUNCH(k_plain_morton, (unsigned)cdiv(nu, 256), 256, 0, h->key3.as<uint32_t>(), nu, h->plainm.as<uint64_t>());
  LAUNCH(k_hash_insert, (unsigned)cdiv(nu, 256), 256, 0, h->plainm.as<uint64_t>(), nu, h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(), h->hmask);
  khash.stop();
  // occupancy grids over the occupied key range (+ margin: no bounds checks in the searches)
  const int margin = std::max(rho, r2c) + 1;
  BitGrid& g = h->grid;
  g.x0 = (int)h->kminmax[0] - margin; g.y0 = (int)h->kminmax[1] - margin; g.z0 = (int)h->kminmax[2] - margin;
  const uint64_t nx = (uint64_t)(h->kminmax[3] - h->kminmax[0] + 1) + 2 * margin;
  g.ny = (uint32_t)(h->kminmax[4] - h->kminmax[1] + 1) + 2 * margin;
  g.nz = (uint32_t)(h->kminmax[5] - h->kminmax[2] + 1) + 2 * margin;
  const uint64_t nbits = nx * g.ny * g.nz;
  if (nbits > ((uint64_t)1 << 35)) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: occupied key range too large for the occupancy grid (> 4 GB): tile the scene");
  const size_t bm_bytes = (size_t)(nbits / 8) + 64;
  CK(h->bm_all.reserve(bm_bytes)); CK(h->bm_used.reserve(bm_bytes));
  KTimer kgrid(h, K_GRID);
  CK(cudaMemsetAsync(h->bm_all.p, 0, bm_bytes, h->stream));
  CK(cudaMemsetAsync(h->bm_used.p, 0, bm_bytes, h->stream));
  LAUNCH(k_bitgrid_set, (unsigned)cdiv(nu, 256), 256, 0, h->key3.as<uint32_t>(), h->uflags.as<uint8_t>(), nu, g, h->bm_all.as<uint32_t>(),
         h->bm_used.as<uint32_t>());
  // dense id grid (4 B per cell) when it fits the budget: neighbour ids by one load instead of hash probes
  h->use_idgrid = false;
  if (nbits * 4 <= h->idgrid_budget) {
    CK(h->idgrid.reserve((size_t)nbits * 4 + 64));
    CK(cudaMemsetAsync(h->idgrid.p, 0xff, (size_t)nbits * 4, h->stream));
    LAUNCH(k_idgrid_set, (unsigned)cdiv(nu, 256), 256, 0, h->key3.as<uint32_t>(), nu, g, h->idgrid.as<int32_t>());
    h->use_idgrid = true;
  }
  kgrid.stop();
  h->grid_bytes = (int64_t)bm_bytes + (h->use_idgrid ? (int64_t)nbits * 2 : 0);
  LatticeGeom& lg = h->lgeo;
  lg.res_f = res_f; lg.mnx = h->bb_f[0]; lg.mny = h->bb_f[1]; lg.mnz = h->bb_f[2];
  { double r = (double)graph_size; lg.r2 = (float)(r * r); }
  lg.rho = rho; lg.r2c = r2c;
  // counts: neighbours of every voxel, weight-row length of every used voxel
  const bool want_rows = h->rows_ok && h->use_pair_cache;
  CK(h->adj_cnt.reserve((size_t)(nu + 1) * 4)); CK(h->adj_off.reserve((size_t)(nu + 1) * 4));
  CK(h->row_len.reserve((size_t)(nu + 1) * 4)); CK(h->row_off.reserve((size_t)(nu + 1) * 4));
  CK(h->long_rows.reserve((size_t)nu * 4 + 16)); CK(h->cstats.reserve(256));
  CK(cudaMemsetAsync(h->cstats.p, 0, 256, h->stream));
  KTimer kcnt(h, K_ADJ_COUNT);
  LAUNCH(k_adj_count, (unsigned)cdiv(nu, 8), 256, 0, h->key3.as<uint32_t>(), h->uflags.as<uint8_t>(), nu, lg, g, h->bm_all.as<uint32_t>(),
         h->bm_used.as<uint32_t>(), h->d_adj_cols.as<int4>(), (int)h->adj_cols_host.size(), h->d_pc_cols.as<int4>(),
         want_rows ? (int)h->pc_cols_host.size() : 0, h->adj_cnt.as<uint32_t>(), h->row_len.as<uint32_t>(),
         ROWS_SHORT_CAP, h->long_rows.as<uint32_t>(), h->cstats.as<CountStats>());
  unsigned long long* d_tot_adj = h->small.as<unsigned long long>() + 8;
  unsigned long long* d_tot_rows = h->small.as<unsigned long long>() + 9;
  vgs_status s = scan_u32(h, h->adj_cnt.as<uint32_t>(), h->adj_off.as<uint32_t>(), nu, nullptr, d_tot_adj);
  if (s) return s;
  s = scan_u32(h, h->row_len.as<uint32_t>(), h->row_off.as<uint32_t>(), nu, nullptr, d_tot_rows);
  if (s) return s;
  CK(cudaMemcpyAsync(h->adj_off.as<uint32_t>() + nu, d_tot_adj, 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->row_off.as<uint32_t>() + nu, d_tot_rows, 4, cudaMemcpyDeviceToDevice, h->stream));
  kcnt.stop();
  unsigned long long totals[2] = {0, 0};
  CountStats cs{};
  CK(cudaMemcpyAsync(totals, d_tot_adj, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&cs, h->cstats.p, sizeof(cs), cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  if (totals[0] >= (1ull << 32)) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: more than 2^32 adjacency entries on one device");
  if (totals[1] >= (1ull << 32)) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: more than 2^32 pair-weight entries on one device");
  h->n_adj = (int64_t)totals[0];
  h->n_rows = (int64_t)totals[1];
  h->n_long = (int64_t)cs.n_long;
  h->n_pairs = (int64_t)cs.sum_nn; h->max_n = (int64_t)cs.max_n; h->n_used = (int64_t)cs.n_used;
  if (cs.max_n > 255) return h->fail(VGS_ERR_LIMIT, "vgs_find_adjacency: a voxel has more than 255 neighbours within graph_size");
  CK(h->adj_idx.reserve((size_t)h->n_adj * 4 + 16)); CK(h->adj_code.reserve((size_t)h->n_adj * 2 + 16));
  unsigned* d_err = h->small.as<unsigned>() + 200;
  CK(cudaMemsetAsync(d_err, 0, 4, h->stream));
  AdjTables tb{h->tb_slot.as<uint32_t>(), h->tb_code5.as<uint16_t>(), h->tb_first.as<uint16_t>(), h->tb_last.as<uint16_t>()};
  const int32_t* d_idg = h->use_idgrid ? h->idgrid.as<int32_t>() : nullptr;
  KTimer kfill(h, K_ADJ_FILL);
  LAUNCH(k_adj_fill, (unsigned)cdiv(nu, ADJ_WARPS), ADJ_WARPS * 32, adj_fill_smem(nst, 2 * rho + 1), h->key3.as<uint32_t>(), nu, lg, g, h->bm_all.as<uint32_t>(),
         h->d_adj_cols.as<int4>(), (int)h->adj_cols_host.size(), tb, nst, d_idg, h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(), h->hmask,
         h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), h->adj_code.as<uint16_t>(), d_err);
  kfill.stop();
  h->have_adj = true;
  t.stop();
  return VGS_OK;
}

static vgs_status segment_graph(vgs_handle h, const vgs_sigmas* sg, float cut_thred);
static vgs_status segment_finish(vgs_handle h, const vgs_sigmas* sg, float cut_thred, int adjacency_min);

vgs_status vgs_segment(vgs_handle h, const vgs_sigmas* sg, float cut_thred, int adjacency_min) {
  if (!h || !sg) return VGS_ERR_INVALID;
  if (!h->have_adj) return h->fail(VGS_ERR_STATE, "vgs_segment: call vgs_find_adjacency first");
  vgs_status s = segment_graph(h, sg, cut_thred);
  if (s) return s;
  return segment_finish(h, sg, cut_thred, adjacency_min);
}

// general kernel (one CTA per unit, weights evaluated from the records) over a list of units
static vgs_status launch_general(vgs_handle h, cudaStream_t st, const uint32_t* list, uint32_t count, int n_class, int T, uint32_t max_n,
                                 const GraphParams& gp, const float* d_wempty) {
  const int ncap = (int)((max_n + 3u) & ~3u), mcap = n_class * (n_class - 1);
  size_t smem = (size_t)mcap * 6 + (size_t)LG_CS * 6 + (size_t)ncap * (14 + 4 * REC_PAD) + (2 * LG_BINS + 2) * 4 + 64;
  const int bucketed = (smem + (size_t)mcap * 2 <= 220 * 1024) ? 1 : 0;   // bin-ordered pool index if it fits
  if (bucketed) smem += (size_t)mcap * 2;
#define LG(TT)                                                                                                      \
  do {                                                                                                              \
    auto kfn = k_local_graph2<TT>;                                                                                  \
    LAUNCH_ON(st, kfn, count, TT, smem, list, count, h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(),           \
              h->rec.as<float>(), gp, ncap, mcap, d_wempty, bucketed, h->conn0_cnt.as<uint32_t>(),                  \
              h->conn0_idx.as<int32_t>());                                                                          \
  } while (0)
  if (T == 64) LG(64); else if (T == 128) LG(128); else LG(256);
#undef LG
  return VGS_OK;
}

static vgs_status segment_graph(vgs_handle h, const vgs_sigmas* sg, float cut_thred) {
  CK(cudaSetDevice(h->device));
  const int64_t nu = h->nu;
  h->have_segments = false;
  h->have_cluster_stats = false; h->have_csr = false;
  h->conn0_is_mask = false;
  GraphParams gp;
  gp.pp = PairParams{sg->sig_p, sg->sig_n, sg->sig_o, sg->sig_e, sg->sig_c, sg->sig_w, h->mode == VGS_MODE_SVGS ? 1 : 0};
  gp.cut = cut_thred;
  const size_t E = (size_t)h->n_adj;
  CK(h->conn0_cnt.reserve((size_t)nu * 4));
  CK(h->conn1_cnt.reserve((size_t)nu * 4)); CK(h->conn1_idx.reserve(E * 4 + 16));
  CK(h->attach.reserve((size_t)nu * 4)); CK(h->parent.reserve((size_t)nu * 4)); CK(h->root.reserve((size_t)nu * 4));
  StageTimer t(h, &h->tm.graph_ms, 5);
  float* d_wempty = h->small.as<float>() + 160;
  LAUNCH(k_wempty, 1, 1, 0, gp.pp, d_wempty);
  unsigned long long* d_dbg = nullptr;
  if (getenv("VGS_B200_DEBUG_COUNTERS")) {
    d_dbg = h->small.as<unsigned long long>() + 64;
    CK(cudaMemsetAsync(d_dbg, 0, 64, h->stream));
  }
  const bool rows_path = h->mode == VGS_MODE_VGS && h->rows_ok && h->use_pair_cache;
  if (rows_path) {
    // ---- stage 4: weight rows (every unordered pair of used voxels once) ----
    const LatticeGeom& lg = h->lgeo;
    unsigned* d_err = h->small.as<unsigned>() + 200;
    uint32_t* d_fb_count = h->small.as<uint32_t>() + 201;
    CK(cudaMemsetAsync(d_err, 0, 8, h->stream));
    // compact list of the used voxels (id order): one warp per USED voxel in the two kernels below
    const int64_t n_used = h->n_used;
    {
      StageTimer tpc(h, &h->tm.pair_cache_ms, 11);
      CK(h->rows.reserve((size_t)h->n_rows * 16 + 64));
      KTimer krf(h, K_ROWS_FILL);
      CK(h->used_list.reserve((size_t)nu * 4 + 16));
      CK(h->flags.reserve((size_t)nu * 4 + 16)); CK(h->scan.reserve((size_t)nu * 4 + 16));
      LAUNCH(k_used_flags, (unsigned)cdiv(nu, 256), 256, 0, h->uflags.as<uint8_t>(), nu, h->flags.as<uint32_t>());
      { vgs_status s_ = scan_u32(h, h->flags.as<uint32_t>(), h->scan.as<uint32_t>(), nu, nullptr); if (s_) return s_; }
      LAUNCH(k_used_write, (unsigned)cdiv(nu, 256), 256, 0, h->uflags.as<uint8_t>(), h->scan.as<uint32_t>(), nu, h->used_list.as<uint32_t>());
      if (n_used > 0)
      LAUNCH(k_rows_fill, (unsigned)cdiv(n_used, RF_WARPS), RF_WARPS * 32, RF_WARPS * rows_fill_smem_warp(2 * lg.r2c + 1), h->used_list.as<uint32_t>(),
             (uint32_t)n_used, h->key3.as<uint32_t>(), h->rec.as<float>(), lg, h->grid, h->bm_used.as<uint32_t>(),
             h->d_pc_cols.as<int4>(), (int)h->pc_cols_host.size(), h->use_idgrid ? h->idgrid.as<int32_t>() : nullptr, h->tk.as<unsigned long long>(),
             h->tv.as<uint32_t>(), h->hmask, gp.pp, h->row_off.as<uint32_t>(), h->rows.as<uint4>(), d_err);
      krf.stop();
      if (h->n_long > 0) {   // rows longer than the shared-memory assembly of k_rows_fill
        KTimer krs(h, K_ROWS_SORT);
        const int cap_long = (int)((h->max_row_len + 63) & ~(int64_t)63);
        LAUNCH(k_rows_sort, (unsigned)h->n_long, 32, (size_t)2 * cap_long * 16, h->row_off.as<uint32_t>(), h->long_rows.as<uint32_t>(),
               (uint32_t)h->n_long, cap_long, h->rows.as<uint4>());
        krs.stop();
      }
      tpc.stop();
    }
    // ---- stage 5a: one warp per voxel ----
    const int mw = h->mwords;
    CK(h->conn_mask.reserve((size_t)nu * mw * 4 + 16));
    CK(h->fallback.reserve((size_t)nu * 4 + 16));
    KTimer kgr(h, K_GRAPH_ROWS);
    const int ncap = (int)std::min<int64_t>(LR_NCAP, (h->max_n + 3) & ~(int64_t)3);   // vertices of the largest neighbourhood
    // launch order: big neighbourhoods first (VGS_B200_LG_ORDER=0: id order)
    const uint32_t* d_order = h->used_list.as<uint32_t>();
    if (h->lg_order > 0 && n_used > 0) {
      CK(h->ckeysA.reserve((size_t)nu * 8 + 16)); CK(h->ckeysB.reserve((size_t)nu * 8 + 16));
      CK(h->cvalsA.reserve((size_t)nu * 4 + 16)); CK(h->cvalsB.reserve((size_t)nu * 4 + 16));
      const int shift = h->lg_order >= 8 ? 0 : 8 - h->lg_order;      // lg_order = number of key bits (classes of 2^(8-bits) sizes)
      LAUNCH(k_order_keys, (unsigned)cdiv(n_used, 256), 256, 0, h->used_list.as<uint32_t>(), (uint32_t)n_used, h->adj_off.as<uint32_t>(), shift,
             h->ckeysA.as<uint32_t>(), h->cvalsA.as<uint32_t>());
      uint32_t* ok; uint32_t* ov;
      vgs_status s_ = radix_sort<uint32_t>(h, n_used, 8 - shift, &ok, &ov, h->ckeysA.as<uint32_t>(), h->ckeysB.as<uint32_t>(), h->cvalsA.as<uint32_t>(),
                                           h->cvalsB.as<uint32_t>());
      if (s_) return s_;
      d_order = ov;
    }
    CK(cudaMemsetAsync(h->conn0_cnt.p, 0, (size_t)nu * 4, h->stream));            // unused voxels: empty connect lists
    CK(cudaMemsetAsync(h->conn_mask.p, 0, (size_t)nu * mw * 4, h->stream));
    if (n_used > 0)
      LAUNCH(k_local_graph_rows, (unsigned)cdiv(n_used, LR_WARPS), 32 * LR_WARPS, LR_WARPS * lr_smem_bytes(h->lbits, mw, ncap), d_order,
             (uint32_t)n_used, h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), h->adj_code.as<uint16_t>(), h->uflags.as<uint8_t>(), cut_thred, lg.rho,
             h->lbits, mw, ncap, h->row_off.as<uint32_t>(), h->rows.as<uint4>(), d_wempty, h->conn0_cnt.as<uint32_t>(), h->conn_mask.as<uint32_t>(),
             h->fallback.as<uint32_t>(), d_fb_count, h->force_fallback, h->lr_target, d_dbg);
    kgr.stop();
    uint32_t fe[2] = {0, 0};   // [0] = error bits of the fill kernels, [1] = units handed back
    CK(cudaMemcpyAsync(fe, d_err, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    if (fe[0]) return h->fail(VGS_ERR_CUDA, "vgs_segment: internal inconsistency between the occupancy grid and the voxel table (error bits " + std::to_string(fe[0]) + ")");
    h->n_fallback = fe[1];
    if (d_dbg) {
      unsigned long long dd[8];
      CK(cudaMemcpyAsync(dd, d_dbg, 64, cudaMemcpyDeviceToHost, h->stream));
      CK(stream_wait(h->stream));
      fprintf(stderr, "[vgs debug] units %llu rounds %llu (%.2f/unit) staged %llu (%.1f/unit) final_nseg/unit %.2f nv/unit %.1f fallback %u\n", dd[3], dd[0],
              (double)dd[0] / (double)std::max(1ull, dd[3]), dd[2], (double)dd[2] / (double)std::max(1ull, dd[3]),
              (double)dd[5] / (double)std::max(1ull, dd[3]), (double)dd[6] / (double)std::max(1ull, dd[3]), fe[1]);
    }
    if (fe[1]) {   // units the row kernel handed back: general kernel sized for the largest neighbourhood, then list -> mask
      CK(h->conn0_idx.reserve(E * 4 + 16));
      KTimer kgg(h, K_GRAPH_GENERAL);
      vgs_status s_ = launch_general(h, h->stream, h->fallback.as<uint32_t>(), fe[1], CLASS_N_HOST[N_CLASSES - 1], 256, (uint32_t)h->max_n, gp, d_wempty);
      if (s_) return s_;
      LAUNCH(k_conn_list_to_mask, (unsigned)cdiv((int64_t)fe[1] * 32, 128), 128, 0, h->fallback.as<uint32_t>(), fe[1], h->adj_off.as<uint32_t>(),
             h->adj_idx.as<int32_t>(), h->adj_code.as<uint16_t>(), h->conn0_cnt.as<uint32_t>(), h->conn0_idx.as<int32_t>(), lg.rho, mw,
             h->conn_mask.as<uint32_t>());
      kgg.stop();
    }
    h->conn0_is_mask = true;
    t.stop();
    h->have_graph = true;
    return VGS_OK;
  }
  // ---- general path (SVGS; VGS when the weight rows are disabled or the stencil is too wide): units binned by the
  //      size of their local graph, one launch per class over up to 6 streams ----
  CK(h->conn0_idx.reserve(E * 4 + 16));
  CK(h->class_count.reserve(256));
  KTimer kgg(h, K_GRAPH_GENERAL);
  CK(h->ckeysA.reserve((size_t)nu * 8 + 16)); CK(h->ckeysB.reserve((size_t)nu * 8 + 16));
  CK(h->cvalsA.reserve((size_t)nu * 4 + 16)); CK(h->cvalsB.reserve((size_t)nu * 4 + 16));
  LAUNCH(k_class_init, (unsigned)cdiv(nu, 256), 256, 0, h->ckeysA.as<uint64_t>(), h->cvalsA.as<uint32_t>(), nu);
  unsigned long long* d_stats = h->small.as<unsigned long long>() + 24;
  uint32_t* d_maxn = h->class_count.as<uint32_t>() + 16;
  CK(cudaMemsetAsync(h->class_count.p, 0, 256, h->stream));
  CK(cudaMemsetAsync(d_stats, 0, 32, h->stream));
  CK(cudaMemsetAsync(h->conn0_cnt.p, 0, (size_t)nu * 4, h->stream));
  LAUNCH(k_bin_classes, (unsigned)cdiv(nu * 32 + 1, 128), 128, 0, h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(),
         h->uflags.as<uint8_t>(), nu, (int64_t)0, nu, cut_thred, h->mode == VGS_MODE_SVGS ? 1 : 0, d_wempty, h->class_count.as<uint32_t>(),
         d_maxn, h->ckeysA.as<uint64_t>(), h->cvalsA.as<uint32_t>(), d_stats);
  // class lists ordered by unit id: one stable 8-bit radix pass on (class, unit id)
  uint64_t* cls_keys; uint32_t* cls_sorted;
  {
    vgs_status s_ = radix_sort<uint64_t>(h, nu, 8, &cls_keys, &cls_sorted, h->ckeysA.as<uint64_t>(), h->ckeysB.as<uint64_t>(),
                               h->cvalsA.as<uint32_t>(), h->cvalsB.as<uint32_t>());
    if (s_) return s_;
  }
  uint32_t cc[N_CLASSES], cmaxn[N_CLASSES];
  unsigned long long stats[3];
  CK(cudaMemcpyAsync(cc, h->class_count.p, sizeof(cc), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cmaxn, d_maxn, sizeof(cmaxn), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  h->n_pairs = (int64_t)stats[0]; h->max_n = (int64_t)stats[1];
  if (stats[2]) return h->fail(VGS_ERR_LIMIT, "vgs_segment: a local graph has more than 181 enumerated (or 255 total) units (graph_size / unit spacing too large)");
  size_t class_off[N_CLASSES + 1];
  class_off[0] = 0;
  for (int c = 0; c < N_CLASSES; c++) class_off[c + 1] = class_off[c] + cc[c];
  // classes of big units first (few units, long per-unit time), spread over the class streams
  const int nstreams = h->class_streams;
  if (nstreams > 1) {
    CK(cudaEventRecord(h->ev_fork, h->stream));
    for (int i = 0; i < nstreams - 1; i++) CK(cudaStreamWaitEvent(h->aux[i], h->ev_fork, 0));
  }
  int launch_no = 0;
  for (int c = N_CLASSES - 1; c >= 0; c--) {
    if (!cc[c]) continue;
    const int si = launch_no++ % nstreams;
    cudaStream_t st = si == 0 ? h->stream : h->aux[si - 1];
    vgs_status s_ = launch_general(h, st, cls_sorted + class_off[c], cc[c], CLASS_N_HOST[c], CLASS_T_HOST[c], cmaxn[c], gp, d_wempty);
    if (s_) return s_;
  }
  if (nstreams > 1)
    for (int i = 0; i < nstreams - 1; i++) {
      CK(cudaEventRecord(h->ev_join[i], h->aux[i]));
      CK(cudaStreamWaitEvent(h->stream, h->ev_join[i], 0));
    }
  kgg.stop();
  t.stop();
  h->have_graph = true;
  return VGS_OK;
}

// ---- stage 5b: mutual filter (crossValidation VS.h:2111-2179) ----
static vgs_status stage_mutual(vgs_handle h) {
  const int64_t nu = h->nu;
  const size_t E = (size_t)h->n_adj;
  CK(h->conn1_cnt.reserve((size_t)nu * 4)); CK(h->conn1_idx.reserve(E * 4 + 16));
  StageTimer t(h, &h->tm.mutual_ms, 6);
  KTimer km(h, K_MUTUAL);
  if (h->conn0_is_mask)
    LAUNCH(k_mutual_mask, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), h->adj_code.as<uint16_t>(),
           h->conn0_cnt.as<uint32_t>(), h->conn_mask.as<uint32_t>(), nu, h->lgeo.rho, h->mwords, h->conn1_cnt.as<uint32_t>(),
           h->conn1_idx.as<int32_t>());
  else
    LAUNCH(k_mutual, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->conn0_cnt.as<uint32_t>(),
           h->conn0_idx.as<int32_t>(), nu, h->conn1_cnt.as<uint32_t>(), h->conn1_idx.as<int32_t>());
  km.stop();
  t.stop();
  return VGS_OK;
}

static vgs_status segment_finish(vgs_handle h, const vgs_sigmas* sg, float cut_thred, int adjacency_min) {
  CK(cudaSetDevice(h->device));
  const int64_t nu = h->nu;
  h->have_cluster_stats = false; h->have_csr = false;
  GraphParams gp;
  gp.pp = PairParams{sg->sig_p, sg->sig_n, sg->sig_o, sg->sig_e, sg->sig_c, sg->sig_w, h->mode == VGS_MODE_SVGS ? 1 : 0};
  gp.cut = cut_thred;
  CK(h->attach.reserve((size_t)nu * 4)); CK(h->parent.reserve((size_t)nu * 4)); CK(h->root.reserve((size_t)nu * 4));
  { vgs_status s_ = stage_mutual(h); if (s_) return s_; }
  // ---- stage 5c: closest check ----
  {
    StageTimer t(h, &h->tm.closest_ms, 7);
    KTimer kc(h, K_CLOSEST);
    CK(cudaMemsetAsync(h->attach.p, 0xff, (size_t)nu * 4, h->stream));
    uint32_t* d_changed = h->small.as<uint32_t>() + 128;
    uint32_t* d_scnt = h->small.as<uint32_t>() + 132;
    CK(h->singles.reserve((size_t)nu * 4 + 16));
    CK(h->singles_best.reserve((size_t)nu * 8 + 16));          // start of every single's kept candidate weights
    CK(cudaMemsetAsync(d_scnt, 0, 16, h->stream));             // [0] eligible, [1] all singles, [2..3] kept weights (u64)
    LAUNCH(k_collect_singles, (unsigned)cdiv(nu, 256), 256, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(), nu, adjacency_min,
           h->singles.as<uint32_t>(), d_scnt, h->singles_best.as<unsigned long long>(), reinterpret_cast<unsigned long long*>(d_scnt + 2));
    uint32_t scnt[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(scnt, d_scnt, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(stream_wait(h->stream));
    const unsigned long long n_kept = (unsigned long long)scnt[2] | ((unsigned long long)scnt[3] << 32);
    // The first round visits every eligible single, keeps the candidate weights and files the singles whose result can
    // still move (a candidate that is a smaller-id single) in a sublist; the following rounds visit only that sublist (its
    // length is read on the device) and pick from the kept weights.
    // Rounds are launched three at a time with one flag each and read back together (a round costs microseconds, a host
    // round trip more): the fixed point is reached when a round changes nothing.
    int rounds = 0;
    bool more = scnt[0] > 0;
    uint32_t* d_dep_cnt = h->small.as<uint32_t>() + 136;
    if (more) {
      CK(h->singles_dep.reserve((size_t)scnt[0] * 4 + 16));
      CK(h->closest_w.reserve((size_t)n_kept * 4 + 16));              // candidate weights of the eligible singles
      CK(cudaMemsetAsync(d_dep_cnt, 0, 4, h->stream));
    }
    while (more) {
      constexpr int BATCH = 3;
      CK(cudaMemsetAsync(d_changed, 0, 4 * BATCH, h->stream));
      for (int b = 0; b < BATCH; b++)
        LAUNCH(k_closest_round_warp, (unsigned)cdiv((int64_t)scnt[0] * 32, 128), 128, 0, h->singles.as<uint32_t>(), scnt[0],
               h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), h->conn1_cnt.as<uint32_t>(), h->rec.as<float>(), nu, gp.pp,
               h->attach.as<int32_t>(), d_changed + b, (rounds == 0 && b == 0) ? 1 : 0, h->closest_w.as<float>(), h->singles_best.as<unsigned long long>(),
               h->singles_dep.as<uint32_t>(), d_dep_cnt);
      uint32_t changed[BATCH] = {0, 0, 0};
      CK(cudaMemcpyAsync(changed, d_changed, 4 * BATCH, cudaMemcpyDeviceToHost, h->stream));
      CK(stream_wait(h->stream));
      for (int b = 0; b < BATCH && more; b++) { rounds++; if (!changed[b]) more = false; }
      if (rounds > 100000) return h->fail(VGS_ERR_LIMIT, "vgs_segment: closest-check did not converge");
    }
    unsigned long long singles = scnt[1];
    h->n_singles = (int64_t)singles;
    h->closest_rounds = rounds;
    kc.stop();
    t.stop();
  }
  // ---- stage 5d: components ----
  {
    StageTimer t(h, &h->tm.components_ms, 8);
    KTimer kcc(h, K_COMPONENTS);
    LAUNCH(k_cc_init, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(), h->conn1_idx.as<int32_t>(),
           h->attach.as<int32_t>(), nu, h->parent.as<int>());
    for (int r = 0; r < h->cc_jumps; r++) LAUNCH(k_cc_jump, (unsigned)cdiv(nu, 256), 256, 0, h->parent.as<int>(), nu);
    LAUNCH(k_cc_hook, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(),
           h->conn1_idx.as<int32_t>(), h->attach.as<int32_t>(), nu, h->parent.as<int>());
    LAUNCH(k_cc_flatten, (unsigned)cdiv(nu, 256), 256, 0, h->parent.as<int>(), nu, h->root.as<int>());
    kcc.stop();
    t.stop();
  }
  h->have_segments = true;
  return VGS_OK;
}

static vgs_status ensure_cluster_stats(vgs_handle h, int voxels_min) {
  if (!h->have_segments) return h->fail(VGS_ERR_STATE, "results requested before vgs_segment");
  const int min_excl = h->mode == VGS_MODE_SVGS ? -1 : voxels_min;
  if (h->have_cluster_stats && h->last_voxels_min == min_excl) return VGS_OK;
  const int64_t nu = h->nu;
  CK(h->csize.reserve((size_t)nu * 4)); CK(h->cminpt.reserve((size_t)nu * 4));
  CK(cudaMemsetAsync(h->csize.p, 0, (size_t)nu * 4, h->stream));
  CK(cudaMemsetAsync(h->cminpt.p, 0xff, (size_t)nu * 4, h->stream));
  LAUNCH(k_cluster_stats, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), h->ustart.as<uint32_t>(), h->d_perm, nu,
         h->csize.as<uint32_t>(), h->cminpt.as<uint32_t>());
  unsigned long long* d_cnt = h->small.as<unsigned long long>() + 48;
  CK(cudaMemsetAsync(d_cnt, 0, 16, h->stream));
  LAUNCH(k_cluster_count, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), h->csize.as<uint32_t>(), nu, min_excl, d_cnt);
  unsigned long long c2[2];
  CK(cudaMemcpyAsync(c2, d_cnt, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(stream_wait(h->stream));
  h->n_clusters_all = (int64_t)c2[0]; h->n_clusters_exp = (int64_t)c2[1];
  h->have_cluster_stats = true; h->last_voxels_min = min_excl;
  return VGS_OK;
}

vgs_status vgs_cluster_count(vgs_handle h, int voxels_min, int64_t* n_all, int64_t* n_exported) {
  if (!h) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  vgs_status s = ensure_cluster_stats(h, voxels_min);
  if (s) return s;
  if (n_all) *n_all = h->n_clusters_all;
  if (n_exported) *n_exported = h->n_clusters_exp;
  return VGS_OK;
}

vgs_status vgs_get_point_labels(vgs_handle h, int voxels_min, int32_t* labels, int on_device) {
  if (!h || !labels) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  StageTimer t(h, &h->tm.labels_ms, 9);
  KTimer kl(h, K_LABELS);
  vgs_status s = ensure_cluster_stats(h, voxels_min);
  if (s) return s;
  const int min_excl = h->mode == VGS_MODE_SVGS ? -1 : voxels_min;
  int32_t* d_out = labels;
  if (!on_device) { CK(h->labels_out.reserve((size_t)h->n * 4)); d_out = h->labels_out.as<int32_t>(); }
  LAUNCH(k_point_labels, (unsigned)cdiv(h->n, 256), 256, 0, h->d_perm, h->pos_unit.as<uint32_t>(), h->root.as<int>(),
         h->csize.as<uint32_t>(), h->cminpt.as<uint32_t>(), h->n, h->n_valid, min_excl, d_out, (int32_t*)nullptr);
  kl.stop();
  t.stop();
  if (!on_device) {
    StageTimer t2(h, &h->tm.d2h_ms, 10);
    CK(vgs_hostio::to_host(labels, d_out, (size_t)h->n * 4, h->device, h->stream));   // synchronous: the caller's buffer is valid on return
    t2.stop();
  }
  return VGS_OK;
}

vgs_status vgs_get_clusters_csr(vgs_handle h, int voxels_min, int64_t* n_clusters, int64_t* n_points_total, int64_t* offsets,
                                int32_t* point_idx) {
  if (!h) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  vgs_status s = ensure_cluster_stats(h, voxels_min);
  if (s) return s;
  // The grouping is done on the device (cluster rank by a scan over the exported roots, voxels ordered by cluster with one
  // stable radix sort, point indices copied voxel by voxel) and cached: the size query and the fetch of the drop-in
  // classes cost one pass.  Cluster order = ascending smallest voxel id = the seeds of clusteringVoxels (VS.h:2064).
  const int min_excl = h->mode == VGS_MODE_SVGS ? -1 : voxels_min;
  const int64_t nu = h->nu;
  if (!(h->have_csr && h->csr_min_excl == min_excl)) {
    const uint32_t nc = (uint32_t)h->n_clusters_exp;
    CK(h->flags.reserve((size_t)nu * 4 + 16)); CK(h->scan.reserve((size_t)nu * 4 + 16));
    LAUNCH(k_export_flags, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), h->csize.as<uint32_t>(), nu, min_excl, h->flags.as<uint32_t>());
    if ((s = scan_u32(h, h->flags.as<uint32_t>(), h->scan.as<uint32_t>(), nu, nullptr))) return s;
    CK(h->ckeysA.reserve((size_t)nu * 8 + 16)); CK(h->ckeysB.reserve((size_t)nu * 8 + 16));
    CK(h->cvalsA.reserve((size_t)nu * 4 + 16)); CK(h->cvalsB.reserve((size_t)nu * 4 + 16));
    LAUNCH(k_export_keys, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), h->csize.as<uint32_t>(), h->scan.as<uint32_t>(), nu, min_excl, nc,
           h->ckeysA.as<uint32_t>(), h->cvalsA.as<uint32_t>());
    int bits = 1;
    while ((1ull << bits) <= (unsigned long long)nc) bits++;
    uint32_t* sk; uint32_t* sv;
    if ((s = radix_sort<uint32_t>(h, nu, bits, &sk, &sv, h->ckeysA.as<uint32_t>(), h->ckeysB.as<uint32_t>(), h->cvalsA.as<uint32_t>(),
                                  h->cvalsB.as<uint32_t>()))) return s;
    LAUNCH(k_export_sizes, (unsigned)cdiv(nu, 256), 256, 0, sk, sv, h->ustart.as<uint32_t>(), nu, nc, h->flags.as<uint32_t>());
    unsigned long long total = 0;
    if ((s = scan_u32(h, h->flags.as<uint32_t>(), h->scan.as<uint32_t>(), nu, &total))) return s;
    CK(h->csr_off.reserve(((size_t)nc + 1) * 8 + 16)); CK(h->csr_idx.reserve((size_t)total * 4 + 16));
    LAUNCH(k_export_points, (unsigned)cdiv(nu * 32, 128), 128, 0, sk, sv, h->scan.as<uint32_t>(), h->ustart.as<uint32_t>(), h->d_perm, nu, nc,
           h->csr_off.as<long long>(), h->csr_idx.as<int32_t>());
    const long long tot = (long long)total;
    CK(cudaMemcpyAsync(h->csr_off.as<long long>() + nc, &tot, 8, cudaMemcpyHostToDevice, h->stream));
    CK(stream_wait(h->stream));
    h->csr_total = (int64_t)total;
    h->have_csr = true; h->csr_min_excl = min_excl;
  }
  if (n_clusters) *n_clusters = h->n_clusters_exp;
  if (n_points_total) *n_points_total = h->csr_total;
  if (offsets && point_idx) {
    static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
    CK(cudaMemcpyAsync(offsets, h->csr_off.p, ((size_t)h->n_clusters_exp + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    if (h->csr_total > 0) CK(vgs_hostio::to_host(point_idx, h->csr_idx.p, (size_t)h->csr_total * 4, h->device, h->stream));
    CK(stream_wait(h->stream));
  }
  return VGS_OK;
}

vgs_status vgs_run(vgs_handle h, const vgs_params* p, int32_t* labels, int on_device) {
  if (!h || !p) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaEvent_t a = h->ev[30], b = h->ev[31];
  CK(cudaEventRecord(a, h->stream));
  vgs_status s;
  if ((s = vgs_voxelize(h, p->voxel_size))) return s;
  if ((s = vgs_compute_features(h, p->points_min))) return s;
  if ((s = vgs_find_adjacency(h, p->graph_size))) return s;
  if ((s = vgs_segment(h, &p->sig, p->cut_thred, p->adjacency_min))) return s;
  if (labels && (s = vgs_get_point_labels(h, p->voxels_min, labels, on_device))) return s;
  CK(cudaEventRecord(b, h->stream));
  CK(cudaEventSynchronize(b));
  CK(cudaEventElapsedTime(&h->tm.total_ms, a, b));
  resolve_timers(h);
  return VGS_OK;
}

vgs_status vgs_get_counts(vgs_handle h, vgs_counts* out) {
  if (!h || !out) return VGS_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  out->n_points = h->n; out->n_finite = h->n_finite; out->n_voxels = h->n_voxels; out->n_units = h->nu;
  out->n_used = h->n_used; out->n_adjacency = h->n_adj; out->n_pairs = h->n_pairs; out->n_singles = h->n_singles;
  out->n_clusters_all = h->n_clusters_all; out->n_clusters_exported = h->n_clusters_exp; out->octree_depth = h->depth;
  out->closest_rounds = h->closest_rounds; out->max_neighbours = h->max_n; out->reserved[0] = h->n_fallback;
  return VGS_OK;
}

vgs_status vgs_stage_timings(vgs_handle h, vgs_timings* out) {
  if (!h || !out) return VGS_ERR_INVALID;
  cudaSetDevice(h->device);
  resolve_timers(h);
  *out = h->tm;
  out->kernel_launches = h->launches;
  return VGS_OK;
}

vgs_status vgs_kernel_timings(vgs_handle h, vgs_kernel_timing* out, int* n) {
  if (!h || !n) return VGS_ERR_INVALID;
  cudaSetDevice(h->device);
  resolve_timers(h);
  static const char* names[vgs_context::NK] = {
      "origin: k_origin_scan + k_origin_adopt rounds", "keys: k_quantise", "sort: k_rs_hist + scan + k_rs_scatter per digit", "heads: k_heads_reduce + k_scan_tiles + k_heads_down + k_voxel_keys",
      "features: k_features", "hash: k_plain_morton + k_hash_insert", "grids: memset + k_bitgrid_set", "adjacency count: k_adj_count + 2 scans",
      "adjacency fill: k_adj_fill", "weight rows: k_rows_fill (evaluate, order by weight cell, write once)", "weight rows: k_rows_sort (rows longer than 256 entries)", "local graphs: k_local_graph_rows",
      "local graphs: k_local_graph2 (general / fallback)", "mutual filter: k_mutual(_mask)", "closest check: k_collect_singles + k_closest_round_warp rounds",
      "components: k_cc_init + k_cc_jump + k_cc_hook + k_cc_flatten", "labels: k_cluster_stats + k_cluster_count + k_point_labels",
      "supervoxels: VCCS generator (voxel table, normals, seeds, expansion + refinement rounds)", "supervoxel units: k_label_keys + sort + heads",
      "supervoxel adjacency: centroid grid + k_adjacency_svgs"};
  const int64_t N = h->n, V = h->nu, E = h->n_adj, R = h->n_rows, MW = h->mwords;
  const int64_t KB = h->sort_key_bytes;     // bytes per sort key
  const int passes = (3 * h->depth + 1 + 7) / 8;
  // algorithmic bytes = compulsory HBM traffic with inputs / outputs materialised once
  const int64_t bytes[vgs_context::NK] = {
      12 * N, 12 * N + (KB + 4) * N, (int64_t)passes * 2 * (KB + 4) * N, KB * N + 4 * N + 28 * V, 16 * N + 64 * V, 32 * V, 13 * V + 2 * h->grid_bytes,
      20 * V, 16 * V + 6 * E, 64 * V + 16 * R, 32 * h->n_long * h->max_row_len, 6 * E + 16 * R + 4 * MW * V, 6 * E + 64 * V, 10 * E + 4 * MW * V, 0, 8 * E + 8 * V, 8 * N + 4 * V, 16 * N, 40 * N, 64 * V + 4 * E};
  int c = 0;
  for (int i = 0; i < vgs_context::NK && out && c < *n; i++) {
    if (h->k_launches[i] == 0 && h->k_ms[i] == 0.f) continue;
    out[c].name = names[i]; out[c].ms = h->k_ms[i]; out[c].alg_bytes = bytes[i]; out[c].launches = h->k_launches[i]; out[c].reserved = 0;
    c++;
  }
  *n = c;
  return VGS_OK;
}

vgs_status vgs_debug_get(vgs_handle h, vgs_blob_kind kind, void* dst, size_t* bytes) {
  if (!h || !bytes) return VGS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const int64_t n = h->n, nu = h->nu;
  auto copy_out = [&](const void* src, size_t nbytes) -> vgs_status {
    if (!dst) { *bytes = nbytes; return VGS_OK; }
    if (*bytes < nbytes) return h->fail(VGS_ERR_INVALID, "vgs_debug_get: buffer too small");
    CK(vgs_hostio::to_host(dst, src, nbytes, h->device, h->stream));
    *bytes = nbytes;
    return VGS_OK;
  };
  switch (kind) {
    case VGS_BLOB_POINT_KEY: {
      if (!h->voxelized) return h->fail(VGS_ERR_STATE, "debug: not voxelized");
      if (!dst) { *bytes = (size_t)n * 12; return VGS_OK; }
      CK(h->tmp.reserve((size_t)n * 12 + (size_t)n * 12));
      uint32_t* k3 = h->tmp.as<uint32_t>();
      // keys/vals scratch of a throw-away quantise pass (the sorted arrays stay untouched)
      DBuf sk, sv;
      CK(sk.reserve((size_t)n * 8)); CK(sv.reserve((size_t)n * 4));
      LAUNCH(k_quantise<uint64_t>, (unsigned)cdiv(n, 256), 256, 0, h->d_xyz, h->stride, n, h->ep, (double)h->voxel_size, h->depth,
             h->leaf_order == VGS_LEAF_DESCENDING ? 1 : 0, sk.as<uint64_t>(), sv.as<uint32_t>(), k3);
      vgs_status s = copy_out(k3, (size_t)n * 12);
      sk.release(); sv.release();
      return s;
    }
    case VGS_BLOB_POINT_UNIT: {
      if (!h->have_units) return h->fail(VGS_ERR_STATE, "debug: units not built");
      if (!dst) { *bytes = (size_t)n * 4; return VGS_OK; }
      CK(h->tmp.reserve((size_t)n * 4));
      LAUNCH(k_point_labels, (unsigned)cdiv(n, 256), 256, 0, h->d_perm, h->pos_unit.as<uint32_t>(), (const int*)nullptr,
             (const uint32_t*)nullptr, (const uint32_t*)nullptr, n, h->n_valid, 0, (int32_t*)nullptr, h->tmp.as<int32_t>());
      return copy_out(h->tmp.p, (size_t)n * 4);
    }
    case VGS_BLOB_UNIT_KEY: { vgs_status s = ensure_geometry(h); if (s) return s; return copy_out(h->key3.p, (size_t)nu * 12); }
    case VGS_BLOB_UNIT_CENTER: { vgs_status s = ensure_geometry(h); if (s) return s; return copy_out(h->center.p, (size_t)nu * 12); }
    case VGS_BLOB_UNIT_OFFSETS: {
      if (!h->have_units) break;
      if (!dst) { *bytes = (size_t)(nu + 1) * 8; return VGS_OK; }
      if (*bytes < (size_t)(nu + 1) * 8) return h->fail(VGS_ERR_INVALID, "vgs_debug_get: buffer too small");
      std::vector<uint32_t> t32((size_t)nu + 1);
      CK(stream_wait(h->stream));
      CK(cudaMemcpy(t32.data(), h->ustart.p, ((size_t)nu + 1) * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i <= nu; i++) ((int64_t*)dst)[i] = t32[i];
      *bytes = (size_t)(nu + 1) * 8;
      return VGS_OK;
    }
    case VGS_BLOB_UNIT_POINTS: if (!h->have_units) break; return copy_out(h->d_perm, (size_t)h->n_valid * 4);
    case VGS_BLOB_RECORDS: if (!h->have_features) break; return copy_out(h->rec.p, (size_t)nu * REC_FLOATS * 4);
    case VGS_BLOB_ADJ_OFFSETS: {
      if (!h->have_adj) break;
      if (!dst) { *bytes = (size_t)(nu + 1) * 8; return VGS_OK; }
      if (*bytes < (size_t)(nu + 1) * 8) return h->fail(VGS_ERR_INVALID, "vgs_debug_get: buffer too small");
      std::vector<uint32_t> t32((size_t)nu + 1);
      CK(stream_wait(h->stream));
      CK(cudaMemcpy(t32.data(), h->adj_off.p, ((size_t)nu + 1) * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i <= nu; i++) ((int64_t*)dst)[i] = t32[i];
      *bytes = (size_t)(nu + 1) * 8;
      return VGS_OK;
    }
    case VGS_BLOB_ADJ_IDX: if (!h->have_adj) break; return copy_out(h->adj_idx.p, (size_t)h->n_adj * 4);
    case VGS_BLOB_CONN0_COUNT: if (!h->have_segments) break; return copy_out(h->conn0_cnt.p, (size_t)nu * 4);
    case VGS_BLOB_CONN0_IDX: {
      if (!h->have_segments) break;
      if (!dst) { *bytes = (size_t)h->n_adj * 4; return VGS_OK; }
      if (h->conn0_is_mask) {   // the row kernel keeps the lists as lattice masks: expand them in adjacency order
        CK(h->conn0_idx.reserve((size_t)h->n_adj * 4 + 16));
        LAUNCH(k_conn_mask_to_list, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(),
               h->adj_code.as<uint16_t>(), nu, h->lgeo.rho, h->mwords, h->conn_mask.as<uint32_t>(), h->conn0_idx.as<int32_t>());
      }
      return copy_out(h->conn0_idx.p, (size_t)h->n_adj * 4);
    }
    case VGS_BLOB_CONN1_COUNT: if (!h->have_segments) break; return copy_out(h->conn1_cnt.p, (size_t)nu * 4);
    case VGS_BLOB_CONN1_IDX: if (!h->have_segments) break; return copy_out(h->conn1_idx.p, (size_t)h->n_adj * 4);
    case VGS_BLOB_ATTACH: if (!h->have_segments) break; return copy_out(h->attach.p, (size_t)nu * 4);
    case VGS_BLOB_UNIT_ROOT: if (!h->have_segments) break; return copy_out(h->root.p, (size_t)nu * 4);
    default: return h->fail(VGS_ERR_INVALID, "vgs_debug_get: unknown blob kind");
  }
  return h->fail(VGS_ERR_STATE, "vgs_debug_get: blob not available yet");
}

}  // extern "C"

#include "vgs_group.inl"

// vgs_group.inl — ONE scene on several GPUs: spatial slabs of the voxel lattice + halo layers,
// segmented per rank with the single-device pipeline, components merged across slabs.  Included at the end of
// vgs_b200.cu (same translation unit: it drives the stage functions of the handle).
//
// The reference is single-process (VS.h:372-421 runs on one cloud).  What has to survive the split is its RESULT:
//   * voxel keys come from PCL's dynamic bounding box, which depends on the global insertion order -> the growth
//     epochs are found by rounds of "first point outside the box" over all ranks (min over ranks);
//   * voxel ids are the global leaf order; only their relative order matters, except that closestCheck reads slot 0 of
//     an adjacency list — the neighbour COUNT — as a voxel id (VS.h:2243): a table of the globally first 256 voxels;
//   * closestCheck (VS.h:2181-2303) is sequential in voxel order: fixed-point rounds with an exchange of the state of
//     the singles next to a cut;
//   * clusters that cross a cut are merged by a union-find over voxel KEYS, replicated on every rank.
// Halo = 3 * rho + 1 voxel layers (rho = reach of the radius stencil): an owned single needs the mutual-filtered list
// size of its neighbours (rho), those need the cuts of their neighbours (2 rho), those the records of theirs (3 rho).
//
// Two transports behind the same schedule: NCCL (one rank per process, libnccl.so.2 bound at run time with dlopen so
// that single-GPU users need no NCCL) and a loopback group (all ranks in this process on one device, exchanges are
// device copies) that makes the split testable on a single GPU.
#include <dlfcn.h>
#include <nccl.h>

#include "vgs_tiles.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api(std::string& err) {
  static NcclApi api;
  static bool tried = false;
  if (api.lib) return &api;
  if (tried) { err = "libnccl.so.2 could not be loaded"; return nullptr; }
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);   // the copy torch already mapped, if any
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return nullptr; }
#define NSYM(field, name)                                                        \
  *(void**)(&api.field) = dlsym(lib, name);                                      \
  if (!api.field) { err = std::string("libnccl: missing symbol ") + name; return nullptr; }
  NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommDestroy, "ncclCommDestroy")
  NSYM(AllGather, "ncclAllGather") NSYM(AllReduce, "ncclAllReduce") NSYM(Broadcast, "ncclBroadcast") NSYM(Send, "ncclSend")
  NSYM(Recv, "ncclRecv") NSYM(GroupStart, "ncclGroupStart") NSYM(GroupEnd, "ncclGroupEnd") NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
  api.lib = lib;
  return &api;
}

// element-wise combination of small arrays across the ranks of a loopback group
__global__ void k_comb_sum_u64(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
__global__ void k_comb_max_i32(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = max(dst[i], src[i]);
}
// out = 1 if any of the n flags is set (the round's "anything changed" word, all-reduced with the attached flags)
__global__ void k_fold_flags(const uint32_t* __restrict__ flags, int n, int32_t* __restrict__ out) {
  const bool any = threadIdx.x < (unsigned)n && flags[threadIdx.x] != 0;
  if (__ballot_sync(0xffffffffu, any) && threadIdx.x == 0) *out = 1;
}
// exported clusters of one rank: local components no other rank knows about, and (on one rank only) merged components
__global__ void __launch_bounds__(256) k_group_cluster_count(const int* __restrict__ root, const uint8_t* __restrict__ is_cross,
                                                           const uint32_t* __restrict__ csize, int64_t nu, int min_size_excl,
                                                           unsigned long long* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = u < nu && root[u] == (int)u && !is_cross[u] && (int)csize[u] > min_size_excl;
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(out, (unsigned long long)__popc(bal));
}
__global__ void __launch_bounds__(256) k_group_rep_count(const unsigned long long* __restrict__ tk, int* parent,
                                                       const unsigned long long* __restrict__ minkey, const uint32_t* __restrict__ tsize,
                                                       int64_t cap, int min_size_excl, unsigned long long* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // one slot per merged component: the slot that holds its representative (smallest) key; sizes are filed there
  bool ok = false;
  if (s < cap && tk[s] != HASH_EMPTY) ok = minkey[uf_find(parent, (int)s)] == tk[s] && (int)tsize[s] > min_size_excl;
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(out, (unsigned long long)__popc(bal));
}

// state of one rank of the group
struct TileCtx {
  const float* d_xyz = nullptr;     // the rank's slice of the cloud (device)
  int stride = 3;
  int64_t n = 0;
  long long gfirst = 0;
  int32_t* labels_user = nullptr;
  int64_t cursor = 0, found = -1;
  bool origin_done = false;
  int64_t n_tile = 0, n_lab = 0, n_owned_vox = 0;
  uint32_t n_bnd = 0, n_singles = 0;
  DBuf xyz_in, hist, rcnt, roff, sendrec, tilerec, tsmall, own, gidlo, low_loc, low_all, low_glob, low_att, low_att_out, bnd_out, bnd_all,
      pairs, pairs_all, mtk, mparent, mminkey, rep, is_cross, tsize, tminpt, stat_out, stat_all, lab_recs, lab_send, lab_recv, labels_dev,
      xs, xr;
  std::vector<size_t> sbytes, soff, rbytes, roffs;
};

}  // namespace

struct vgs_group_s {
  int nranks = 1, nlocal = 1, rank0 = 0, device = 0;
  bool use_nccl = false;
  ncclComm_t comm = nullptr;
  NcclApi* api = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<vgs_handle> hs;
  std::vector<TileCtx> tc;
  std::string err;
  void* pin = nullptr;            // pinned host staging of the small exchanges
  size_t pin_bytes = 0;
  vgs_group_counts counts{};
  vgs_group_timings tm{};
  cudaEvent_t ev[16] = {};
  vgs_status fail(vgs_status s, const std::string& m) { err = m; return s; }
  vgs_status fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at vgs_group.inl:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    err = buf;
    return VGS_ERR_CUDA;
  }
  vgs_status fail_nccl(ncclResult_t r, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "NCCL error %d (%s) at vgs_group.inl:%d: %s", (int)r, api ? api->GetErrorString(r) : "?", line, what);
    err = buf;
    return VGS_ERR_CUDA;
  }
};

#define GCK(x)                                                                \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) return g->fail_cuda(e_, #x, __LINE__);             \
  } while (0)
#define NCK(x)                                                                \
  do {                                                                        \
    ncclResult_t r_ = (x);                                                    \
    if (r_ != ncclSuccess) return g->fail_nccl(r_, #x, __LINE__);             \
  } while (0)

namespace {

template <class F>
vgs_status each_rank(vgs_group g, F f) {
  for (int lr = 0; lr < g->nlocal; lr++) {
    vgs_status s = f(lr, g->hs[lr], g->tc[lr]);
    if (s) {
      if (g->err.empty() || !g->hs[lr]->err.empty()) g->err = "rank " + std::to_string(g->rank0 + lr) + ": " + g->hs[lr]->err;
      return s;
    }
  }
  return VGS_OK;
}

// ---- collectives.  "host" variants move a few bytes per rank (through a pinned staging buffer and the device when the
//      transport is NCCL); "dev" variants move device buffers. ----
// all-gather of `bytes` per rank: send[lr] -> out (rank-major, the same on every rank)
vgs_status xg_host(vgs_group g, const void* const* send, void* out, size_t bytes) {
  if (!g->use_nccl) {
    for (int r = 0; r < g->nranks; r++) memcpy((char*)out + (size_t)r * bytes, send[r], bytes);
    return VGS_OK;
  }
  TileCtx& t = g->tc[0];
  const size_t pad = (bytes + 15) & ~(size_t)15;
  if (pad * (g->nranks + 1) > g->pin_bytes) return g->fail(VGS_ERR_LIMIT, "group: small exchange larger than the staging buffer");
  GCK(t.xs.reserve(pad)); GCK(t.xr.reserve(pad * g->nranks));
  memcpy(g->pin, send[0], bytes);
  GCK(cudaMemcpyAsync(t.xs.p, g->pin, pad, cudaMemcpyHostToDevice, g->stream));
  NCK(g->api->AllGather(t.xs.p, t.xr.p, pad, ncclChar, g->comm, g->stream));
  GCK(cudaMemcpyAsync((char*)g->pin + pad, t.xr.p, pad * g->nranks, cudaMemcpyDeviceToHost, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  for (int r = 0; r < g->nranks; r++) memcpy((char*)out + (size_t)r * bytes, (char*)g->pin + pad + (size_t)r * pad, bytes);
  return VGS_OK;
}
// all-gather of device buffers with one size per rank: recv[lr] holds the pieces rank-major and packed
vgs_status xg_dev(vgs_group g, void* const* send, void* const* recv, const size_t* bytes) {
  std::vector<size_t> off(g->nranks + 1, 0);
  for (int r = 0; r < g->nranks; r++) off[r + 1] = off[r] + bytes[r];
  if (off[g->nranks] == 0) return VGS_OK;
  if (!g->use_nccl) {
    for (int d = 0; d < g->nranks; d++)
      for (int r = 0; r < g->nranks; r++)
        if (bytes[r]) GCK(cudaMemcpyAsync((char*)recv[d] + off[r], send[r], bytes[r], cudaMemcpyDeviceToDevice, g->stream));
    return VGS_OK;
  }
  NCK(g->api->GroupStart());
  for (int r = 0; r < g->nranks; r++)
    if (bytes[r]) NCK(g->api->Broadcast(send[0], (char*)recv[0] + off[r], bytes[r], ncclChar, r, g->comm, g->stream));
  NCK(g->api->GroupEnd());
  return VGS_OK;
}
// all-to-all of device buffers: cnt[src * nranks + dst] bytes go from src to dst; send pieces are packed in dst order,
// received pieces are packed in src order
vgs_status xa_dev(vgs_group g, void* const* send, void* const* recv, const std::vector<size_t>& cnt) {
  const int R = g->nranks;
  if (!g->use_nccl) {
    for (int d = 0; d < R; d++) {
      size_t roff = 0;
      for (int s = 0; s < R; s++) {
        size_t soff = 0;
        for (int q = 0; q < d; q++) soff += cnt[(size_t)s * R + q];
        const size_t b = cnt[(size_t)s * R + d];
        if (b) GCK(cudaMemcpyAsync((char*)recv[d] + roff, (char*)send[s] + soff, b, cudaMemcpyDeviceToDevice, g->stream));
        roff += b;
      }
    }
    return VGS_OK;
  }
  const int me = g->rank0;
  size_t soff = 0, roff = 0;
  NCK(g->api->GroupStart());
  for (int r = 0; r < R; r++) {
    const size_t sb = cnt[(size_t)me * R + r], rb = cnt[(size_t)r * R + me];
    if (r == me) {
      if (sb) GCK(cudaMemcpyAsync((char*)recv[0] + roff, (char*)send[0] + soff, sb, cudaMemcpyDeviceToDevice, g->stream));
    } else {
      if (sb) NCK(g->api->Send((char*)send[0] + soff, sb, ncclChar, r, g->comm, g->stream));
      if (rb) NCK(g->api->Recv((char*)recv[0] + roff, rb, ncclChar, r, g->comm, g->stream));
    }
    soff += sb; roff += rb;
  }
  NCK(g->api->GroupEnd());
  return VGS_OK;
}
// in-place all-reduce of small device arrays (kind 0: sum of u64, 1: max of i32)
vgs_status xr_dev(vgs_group g, void* const* buf, int64_t count, int kind) {
  if (!g->use_nccl) {
    for (int r = 1; r < g->nranks; r++) {
      if (kind == 0) k_comb_sum_u64<<<(unsigned)cdiv(count, 256), 256, 0, g->stream>>>((unsigned long long*)buf[0], (const unsigned long long*)buf[r], count);
      else k_comb_max_i32<<<(unsigned)cdiv(count, 256), 256, 0, g->stream>>>((int32_t*)buf[0], (const int32_t*)buf[r], count);
      g->hs[0]->launches++;
    }
    GCK(cudaGetLastError());
    for (int r = 1; r < g->nranks; r++)
      GCK(cudaMemcpyAsync(buf[r], buf[0], (size_t)count * (kind == 0 ? 8 : 4), cudaMemcpyDeviceToDevice, g->stream));
    return VGS_OK;
  }
  if (kind == 0) NCK(g->api->AllReduce(buf[0], buf[0], (size_t)count, ncclUint64, ncclSum, g->comm, g->stream));
  else NCK(g->api->AllReduce(buf[0], buf[0], (size_t)count, ncclInt32, ncclMax, g->comm, g->stream));
  return VGS_OK;
}

// slab cuts from the per-axis key histograms: the axis with the widest occupied range, cut where the running point
// count passes r/R of the total.  cut[0] / cut[R] are open ends.  Pure host code (tests/test_slabs_host.py).
void choose_cuts(const unsigned long long* hist, int nbins, int shift, int nranks, int* axis_out, int* cuts /* nranks + 1 */) {
  int best_axis = 0, best_span = -1;
  for (int a = 0; a < 3; a++) {
    int lo = -1, hi = -1;
    for (int b = 0; b < nbins; b++) if (hist[(size_t)a * nbins + b]) { if (lo < 0) lo = b; hi = b; }
    const int span = lo < 0 ? 0 : hi - lo + 1;
    if (span > best_span) { best_span = span; best_axis = a; }
  }
  const unsigned long long* hv = hist + (size_t)best_axis * nbins;
  unsigned long long total = 0;
  for (int b = 0; b < nbins; b++) total += hv[b];
  cuts[0] = -(1 << 29);
  cuts[nranks] = 1 << 29;
  unsigned long long run = 0;
  int b = 0;
  for (int r = 1; r < nranks; r++) {
    const unsigned long long want = (total * (unsigned long long)r + nranks - 1) / nranks;
    while (b < nbins && run + hv[b] <= want && run < want) { run += hv[b]; b++; }
    // the cut sits on a bin boundary; never before the previous cut
    int c = b << shift;
    if (c < cuts[r - 1] && r > 1) c = cuts[r - 1];
    cuts[r] = c;
  }
  *axis_out = best_axis;
}

vgs_status group_alloc(vgs_group g) {
  g->tc.resize(g->nlocal);
  g->pin_bytes = 1 << 20;
  GCK(cudaMallocHost(&g->pin, g->pin_bytes));
  for (auto& e : g->ev) GCK(cudaEventCreate(&e));
  for (int lr = 0; lr < g->nlocal; lr++) GCK(g->tc[lr].tsmall.reserve(4096));
  return VGS_OK;
}

}  // namespace

extern "C" {

vgs_status vgs_slab_choose_cuts(const uint64_t* hist, int nbins, int shift, int nranks, int* axis, int* cuts) {
  if (!hist || nbins <= 0 || shift < 0 || nranks < 1 || nranks > TILE_MAX_RANKS || !axis || !cuts) return VGS_ERR_INVALID;
  choose_cuts((const unsigned long long*)hist, nbins, shift, nranks, axis, cuts);
  return VGS_OK;
}

vgs_status vgs_group_unique_id(void* id128) {
  if (!id128) return VGS_ERR_INVALID;
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) { g_create_error = "vgs_group_unique_id: " + err; return VGS_ERR_CUDA; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) { g_create_error = "vgs_group_unique_id: ncclGetUniqueId failed"; return VGS_ERR_CUDA; }
  memcpy(id128, &id, 128);
  return VGS_OK;
}

static vgs_status group_make_handles(vgs_group g, const vgs_config* cfg) {
  vgs_config c = *cfg;
  c.mode = VGS_MODE_VGS;
  g->device = c.device;
  if (cudaSetDevice(c.device) != cudaSuccess) return g->fail(VGS_ERR_CUDA, "vgs_group_create: cudaSetDevice failed");
  if (c.stream) g->stream = (cudaStream_t)c.stream;
  else {
    GCK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    g->own_stream = true;
  }
  c.stream = (void*)g->stream;
  for (int lr = 0; lr < g->nlocal; lr++) {
    vgs_handle h = nullptr;
    vgs_status s = vgs_create(&h, &c);
    if (s) return g->fail(s, g_create_error);
    g->hs.push_back(h);
  }
  return group_alloc(g);
}

vgs_status vgs_group_create_local(vgs_group* out, const vgs_config* cfg, int nranks) {
  if (!out || !cfg || nranks < 1 || nranks > TILE_MAX_RANKS) { g_create_error = "vgs_group_create_local: bad argument"; return VGS_ERR_INVALID; }
  *out = nullptr;
  if (cfg->mode != VGS_MODE_VGS) { g_create_error = "vgs_group_create: the slab split is built for VGS (voxel lattice)"; return VGS_ERR_INVALID; }
  vgs_group g = new vgs_group_s();
  g->nranks = g->nlocal = nranks; g->rank0 = 0; g->use_nccl = false;
  vgs_status s = group_make_handles(g, cfg);
  if (s) { g_create_error = g->err; vgs_group_destroy(g); return s; }
  *out = g;
  return VGS_OK;
}

vgs_status vgs_group_create_nccl(vgs_group* out, const vgs_config* cfg, int nranks, int rank, const void* id128) {
  if (!out || !cfg || !id128 || nranks < 1 || nranks > TILE_MAX_RANKS || rank < 0 || rank >= nranks) {
    g_create_error = "vgs_group_create_nccl: bad argument";
    return VGS_ERR_INVALID;
  }
  *out = nullptr;
  if (cfg->mode != VGS_MODE_VGS) { g_create_error = "vgs_group_create: the slab split is built for VGS (voxel lattice)"; return VGS_ERR_INVALID; }
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) { g_create_error = "vgs_group_create_nccl: " + err; return VGS_ERR_CUDA; }
  vgs_group g = new vgs_group_s();
  g->nranks = nranks; g->nlocal = 1; g->rank0 = rank; g->use_nccl = true; g->api = api;
  vgs_status s = group_make_handles(g, cfg);
  if (s) { g_create_error = g->err; vgs_group_destroy(g); return s; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = api->CommInitRank(&g->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    g_create_error = std::string("vgs_group_create_nccl: ncclCommInitRank: ") + api->GetErrorString(r);
    vgs_group_destroy(g);
    return VGS_ERR_CUDA;
  }
  *out = g;
  return VGS_OK;
}

void vgs_group_destroy(vgs_group g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->comm && g->api) g->api->CommDestroy(g->comm);
  for (auto& t : g->tc) {
    DBuf* all[] = {&t.xyz_in, &t.hist, &t.rcnt, &t.roff, &t.sendrec, &t.tilerec, &t.tsmall, &t.own, &t.gidlo, &t.low_loc, &t.low_all, &t.low_glob,
                   &t.low_att, &t.low_att_out, &t.bnd_out, &t.bnd_all, &t.pairs, &t.pairs_all, &t.mtk, &t.mparent, &t.mminkey, &t.rep, &t.is_cross,
                   &t.tsize, &t.tminpt, &t.stat_out, &t.stat_all, &t.lab_recs, &t.lab_send, &t.lab_recv, &t.labels_dev, &t.xs, &t.xr};
    for (DBuf* b : all) b->release();
  }
  for (vgs_handle h : g->hs) vgs_destroy(h);
  for (auto& e : g->ev) if (e) cudaEventDestroy(e);
  if (g->pin) cudaFreeHost(g->pin);
  if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

const char* vgs_group_last_error(vgs_group g) { return g ? g->err.c_str() : g_create_error.c_str(); }

vgs_handle vgs_group_handle(vgs_group g, int local_rank) {
  if (!g || local_rank < 0 || local_rank >= g->nlocal) return nullptr;
  return g->hs[local_rank];
}

vgs_status vgs_group_get_counts(vgs_group g, vgs_group_counts* out) {
  if (!g || !out) return VGS_ERR_INVALID;
  *out = g->counts;
  return VGS_OK;
}
vgs_status vgs_group_get_timings(vgs_group g, vgs_group_timings* out) {
  if (!g || !out) return VGS_ERR_INVALID;
  *out = g->tm;
  return VGS_OK;
}

vgs_status vgs_group_run(vgs_group g, const vgs_params* p, const float* const* xyz, const int64_t* n_per_rank, int stride_bytes, int on_device,
                         int32_t* const* labels) {
  if (!g || !p || !xyz || !n_per_rank || !labels) return VGS_ERR_INVALID;
  if (stride_bytes != 12 && stride_bytes != 16) return g->fail(VGS_ERR_INVALID, "vgs_group_run: stride_bytes must be 12 or 16");
  if (!(p->voxel_size > 0) || !(p->graph_size > 0)) return g->fail(VGS_ERR_INVALID, "vgs_group_run: voxel_size / graph_size must be > 0");
  GCK(cudaSetDevice(g->device));
  g->err.clear();
  const int R = g->nranks, L = g->nlocal;
  cudaStream_t st = g->stream;
  int evi = 0;
  auto mark = [&]() { if (evi < 16) cudaEventRecord(g->ev[evi++], st); };
  mark();   // 0
  vgs_status s;

  // ---- slices: rank r holds the points [first[r], first[r+1]) of the cloud ----
  std::vector<long long> nloc(L), nall(R), first(R + 1, 0);
  {
    std::vector<const void*> sp(L);
    for (int lr = 0; lr < L; lr++) { nloc[lr] = (long long)n_per_rank[lr]; sp[lr] = &nloc[lr]; if (nloc[lr] < 0) return g->fail(VGS_ERR_INVALID, "vgs_group_run: negative point count"); }
    if ((s = xg_host(g, sp.data(), nall.data(), 8))) return s;
    for (int r = 0; r < R; r++) first[r + 1] = first[r] + nall[r];
    if (first[R] <= 0) return g->fail(VGS_ERR_INVALID, "vgs_group_run: empty cloud");
    if (first[R] >= (1ll << 31)) return g->fail(VGS_ERR_LIMIT, "vgs_group_run: the cloud must have < 2^31 points (labels are int32 point indices)");
  }
  s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
    t.n = nloc[lr]; t.stride = stride_bytes / 4; t.gfirst = first[g->rank0 + lr]; t.labels_user = labels[lr];
    t.cursor = 0; t.found = -1; t.origin_done = t.n == 0;
    if (t.n > 0 && !xyz[lr]) return h->fail(VGS_ERR_INVALID, "vgs_group_run: null slice");
    if (on_device || t.n == 0) t.d_xyz = xyz[lr];
    else {
      CK(t.xyz_in.reserve((size_t)t.n * stride_bytes));
      CK(cudaMemcpyAsync(t.xyz_in.p, xyz[lr], (size_t)t.n * stride_bytes, cudaMemcpyHostToDevice, st));
      t.d_xyz = t.xyz_in.as<float>();
    }
    h->d_xyz = t.d_xyz; h->stride = t.stride; h->n = t.n;
    h->voxel_size = p->voxel_size;
    h->voxelized = h->have_units = h->have_features = h->have_adj = h->have_segments = false;
    h->units_external = false;
    h->tm_pending = 0; h->k_pending = 0;
    for (int i = 0; i < vgs_context::NK; i++) { h->k_ms[i] = 0.f; h->k_launches[i] = 0; }
    h->tm = vgs_timings{};
    CK(h->small.reserve(4096));
    return VGS_OK;
  });
  if (s) return s;

  // ---- stage 0 over all ranks: growth epochs of PCL's bounding box in global insertion order ----
  OriginBuilder ob;
  ob.begin((double)p->voxel_size);
  int origin_rounds = 0;
  {
    struct Cand { long long gidx; float p[3]; int pad; };
    std::vector<Cand> mine(L), all(R);
    while (true) {
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        mine[lr].gidx = std::numeric_limits<long long>::max();
        if (t.origin_done) return VGS_OK;
        int64_t idx;
        vgs_status s_ = find_next(h, t.cursor, ob.st, &idx, mine[lr].p);
        if (s_) return s_;
        if (idx >= t.n) { t.origin_done = true; return VGS_OK; }   // the box only grows: this slice is inside for good
        mine[lr].gidx = t.gfirst + idx;
        t.found = idx;
        return VGS_OK;
      });
      if (s) return s;
      std::vector<const void*> sp(L);
      for (int lr = 0; lr < L; lr++) sp[lr] = &mine[lr];
      if ((s = xg_host(g, sp.data(), all.data(), sizeof(Cand)))) return s;
      int win = -1;
      for (int r = 0; r < R; r++)
        if (all[r].gidx != std::numeric_limits<long long>::max() && (win < 0 || all[r].gidx < all[win].gidx)) win = r;
      if (win < 0) break;
      if (const char* why = ob.add(all[win].gidx, all[win].p)) return g->fail(VGS_ERR_LIMIT, std::string("vgs_group_run: ") + why);
      origin_rounds++;
      for (int lr = 0; lr < L; lr++) {
        TileCtx& t = g->tc[lr];
        if (t.origin_done) continue;
        t.cursor = (g->rank0 + lr == win) ? t.found + 1 : t.found;   // a loser's point is re-tested against the grown box
      }
    }
    if (!ob.st.defined) return g->fail(VGS_ERR_INVALID, "vgs_group_run: no finite point in the cloud");
    ob.finish();
  }
  const int depth = (int)ob.st.depth;
  mark();   // 1: origin

  // ---- reach of the radius stencil for the whole scene (the same on every rank) -> halo ----
  int rho_g = 0;
  {
    double maxc = 0;
    for (int a = 0; a < 3; a++) { maxc = std::max(maxc, std::fabs(ob.st.mn[a])); maxc = std::max(maxc, std::fabs(ob.st.mx[a])); }
    int ex = 0;
    std::frexp(maxc, &ex);
    const double noise = maxc > 0 ? 12.0 * (double)p->graph_size * std::ldexp(1.0, ex - 24) : 0.0;
    for (const int4& o : make_stencil(p->voxel_size, p->graph_size, noise))
      rho_g = std::max(rho_g, std::max(std::abs(o.x), std::max(std::abs(o.y), std::abs(o.z))));
  }
  SlabCuts sc{};
  sc.nranks = R; sc.halo = 3 * rho_g + 1;

  // ---- slab cuts from the key histograms of a subsample ----
  {
    const int shift = std::max(0, depth - 11), nbins = 1 << (depth - shift);
    std::vector<void*> hp(L);
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      CK(t.hist.reserve((size_t)3 * nbins * 8));
      CK(cudaMemsetAsync(t.hist.p, 0, (size_t)3 * nbins * 8, st));
      if (t.n > 0) {
        const int step = 8;
        const int64_t ns = cdiv(t.n, step);
        LAUNCH(k_slab_hist, (unsigned)std::min<int64_t>(cdiv(ns, 256), 148 * 4), 256, (size_t)3 * nbins * 4, t.d_xyz, t.stride, t.n, t.gfirst, ob.ep,
               ob.st.res, nbins, shift, step, t.hist.as<unsigned long long>());
      }
      hp[lr] = t.hist.p;
      return VGS_OK;
    });
    if (s) return s;
    if ((s = xr_dev(g, hp.data(), (int64_t)3 * nbins, 0))) return s;
    std::vector<unsigned long long> hh((size_t)3 * nbins);
    GCK(cudaMemcpyAsync(hh.data(), g->tc[0].hist.p, hh.size() * 8, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    choose_cuts(hh.data(), nbins, shift, R, &sc.axis, sc.cut);
  }
  mark();   // 2: cuts

  // ---- route every point to its owner slab and to the slabs whose halo reaches it ----
  std::vector<size_t> cntmat((size_t)R * R, 0);
  {
    std::vector<unsigned long long> mine((size_t)L * R, 0), allc((size_t)R * R, 0);
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      if (t.n == 0) return VGS_OK;
      const int64_t nw = cdiv(t.n, ROUTE_PER_WARP);
      CK(t.rcnt.reserve((size_t)nw * R * 4 + 16)); CK(t.roff.reserve((size_t)(nw * R + 1) * 4 + 16));
      LAUNCH(k_route_count, (unsigned)cdiv(nw * 32, 256), 256, 0, t.d_xyz, t.stride, t.n, t.gfirst, ob.ep, ob.st.res, sc, nw, t.rcnt.as<uint32_t>());
      unsigned long long total = 0;
      vgs_status s_ = scan_u32(h, t.rcnt.as<uint32_t>(), t.roff.as<uint32_t>(), nw * R, &total);
      if (s_) return s_;
      if (total >= (1ull << 32)) return h->fail(VGS_ERR_LIMIT, "vgs_group_run: more than 2^32 routed records from one slice");
      std::vector<uint32_t> starts(R + 1, 0);
      CK(cudaMemcpy2DAsync(starts.data(), 4, t.roff.p, (size_t)nw * 4, 4, R, cudaMemcpyDeviceToHost, st));
      CK(stream_wait(st));
      starts[R] = (uint32_t)total;
      for (int d = 0; d < R; d++) mine[(size_t)lr * R + d] = (unsigned long long)(starts[d + 1] - starts[d]) * 16ull;
      CK(t.sendrec.reserve((size_t)total * 16 + 16));
      LAUNCH(k_route_scatter, (unsigned)cdiv(nw * 32, 256), 256, 0, t.d_xyz, t.stride, t.n, t.gfirst, ob.ep, ob.st.res, sc, nw, t.roff.as<uint32_t>(),
             t.sendrec.as<float4>());
      return VGS_OK;
    });
    if (s) return s;
    std::vector<const void*> sp(L);
    for (int lr = 0; lr < L; lr++) sp[lr] = &mine[(size_t)lr * R];
    if ((s = xg_host(g, sp.data(), allc.data(), (size_t)R * 8))) return s;
    for (size_t i = 0; i < cntmat.size(); i++) cntmat[i] = (size_t)allc[i];
    std::vector<void*> sendp(L), recvp(L);
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      size_t rb = 0;
      for (int r = 0; r < R; r++) rb += cntmat[(size_t)r * R + g->rank0 + lr];
      t.n_tile = (int64_t)(rb / 16);
      if (t.n_tile >= (1ll << 31)) return h->fail(VGS_ERR_LIMIT, "vgs_group_run: a slab holds >= 2^31 points: use more ranks");
      CK(t.tilerec.reserve(rb + 16));
      sendp[lr] = t.sendrec.p; recvp[lr] = t.tilerec.p;
      return VGS_OK;
    });
    if (s) return s;
    if ((s = xa_dev(g, sendp.data(), recvp.data(), cntmat))) return s;
  }
  mark();   // 3: route

  // ---- the single-device pipeline on every tile, up to the mutual filter ----
  int64_t sum_tile_pts = 0, sum_tile_vox = 0, sum_adj = 0;
  s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
    h->n = t.n_tile; h->d_xyz = t.tilerec.as<float>(); h->stride = 4;
    h->nu = 0; h->n_valid = 0; h->n_adj = 0; h->n_voxels = 0; h->n_used = 0;
    ob.install(h);
    sum_tile_pts += t.n_tile;
    if (t.n_tile == 0) return VGS_OK;
    vgs_status s_;
    if ((s_ = voxelize_sorted(h, 1))) return s_;
    if ((s_ = vgs_compute_features(h, p->points_min))) return s_;
    if ((s_ = vgs_find_adjacency(h, p->graph_size))) return s_;
    if (h->st_rho > rho_g) return h->fail(VGS_ERR_LIMIT, "vgs_group_run: stencil reach of a tile exceeds the reach the halo was sized for");
    if ((s_ = segment_graph(h, &p->sig, p->cut_thred))) return s_;
    CK(h->attach.reserve((size_t)h->nu * 4)); CK(h->parent.reserve((size_t)h->nu * 4)); CK(h->root.reserve((size_t)h->nu * 4));
    if ((s_ = stage_mutual(h))) return s_;
    CK(t.own.reserve((size_t)h->nu + 16));
    LAUNCH(k_tile_owner, (unsigned)cdiv(h->nu, 256), 256, 0, h->key3.as<uint32_t>(), h->nu, sc, g->rank0 + lr, t.own.as<uint8_t>());
    sum_tile_vox += h->nu; sum_adj += h->n_adj;
    return VGS_OK;
  });
  if (s) return s;
  mark();   // 4: tiles

  // ---- table of the globally first LOW_IDS voxels (ids a neighbour count can name, VS.h:2243) ----
  int n_low = 0;
  {
    std::vector<void*> sp(L), rp(L);
    std::vector<size_t> bytes(R, (size_t)LOW_IDS * sizeof(LowEntry));
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      CK(t.low_loc.reserve((size_t)LOW_IDS * sizeof(LowEntry))); CK(t.low_all.reserve((size_t)R * LOW_IDS * sizeof(LowEntry)));
      CK(t.low_glob.reserve((size_t)LOW_IDS * sizeof(LowEntry)));
      CK(t.low_att.reserve((LOW_IDS + 1) * 4)); CK(t.low_att_out.reserve((LOW_IDS + 1) * 4));
      CK(cudaMemsetAsync(t.low_loc.p, 0xff, (size_t)LOW_IDS * sizeof(LowEntry), st));     // absent entries sort last
      CK(cudaMemsetAsync(t.low_att.p, 0, LOW_IDS * 4, st));
      if (h->nu > 0) {
        CK(h->flags.reserve((size_t)h->nu * 4)); CK(h->scan.reserve((size_t)h->nu * 4));
        LAUNCH(k_own_flags, (unsigned)cdiv(h->nu, 256), 256, 0, t.own.as<uint8_t>(), h->nu, h->flags.as<uint32_t>());
        vgs_status s_ = scan_u32(h, h->flags.as<uint32_t>(), h->scan.as<uint32_t>(), h->nu, nullptr);
        if (s_) return s_;
        LAUNCH(k_low_export, (unsigned)cdiv(h->nu, 256), 256, 0, t.own.as<uint8_t>(), h->scan.as<uint32_t>(), h->ukey.as<unsigned long long>(),
               h->plainm.as<unsigned long long>(), h->rec.as<float>(), h->conn1_cnt.as<uint32_t>(), h->nu, t.low_loc.as<LowEntry>());
      }
      sp[lr] = t.low_loc.p; rp[lr] = t.low_all.p;
      return VGS_OK;
    });
    if (s) return s;
    if ((s = xg_dev(g, sp.data(), rp.data(), bytes.data()))) return s;
    std::vector<LowEntry> all((size_t)R * LOW_IDS);
    GCK(cudaMemcpyAsync(all.data(), g->tc[0].low_all.p, all.size() * sizeof(LowEntry), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    std::vector<LowEntry> valid;
    for (const LowEntry& e : all) if (e.local >= 0) valid.push_back(e);
    std::sort(valid.begin(), valid.end(), [](const LowEntry& a, const LowEntry& b) { return a.sort_key < b.sort_key; });
    if ((int)valid.size() > LOW_IDS) valid.resize(LOW_IDS);
    n_low = (int)valid.size();
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      if (n_low > 0) CK(cudaMemcpyAsync(t.low_glob.p, valid.data(), (size_t)n_low * sizeof(LowEntry), cudaMemcpyHostToDevice, st));
      CK(t.gidlo.reserve((size_t)h->nu * 2 + 16));
      CK(cudaMemsetAsync(t.gidlo.p, 0xff, (size_t)h->nu * 2 + 16, st));
      if (n_low > 0 && h->nu > 0)
        LAUNCH(k_low_import, (unsigned)cdiv(n_low, 128), 128, 0, t.low_glob.as<LowEntry>(), n_low, h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(),
               h->hmask, t.gidlo.as<uint16_t>());
      return VGS_OK;
    });
    if (s) return s;
    GCK(cudaStreamSynchronize(st));   // `valid` goes out of scope
  }
  mark();   // 5: low ids

  // ---- closest check: fixed point over all ranks ----
  PairParams pp{p->sig.sig_p, p->sig.sig_n, p->sig.sig_o, p->sig.sig_e, p->sig.sig_c, p->sig.sig_w, 0};
  int global_rounds = 0;
  int64_t total_singles = 0;
  {
    // tsmall (u32 words): [0] low count, [4] changed, [8] eligible singles, [9] all singles, [12] boundary records
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      t.n_singles = 0; t.n_bnd = 0;
      if (h->nu == 0) return VGS_OK;
      CK(cudaMemsetAsync(h->attach.p, 0xff, (size_t)h->nu * 4, st));
      CK(h->singles.reserve((size_t)h->nu * 4 + 16));
      uint32_t* d_scnt = t.tsmall.as<uint32_t>() + 8;
      CK(cudaMemsetAsync(d_scnt, 0, 8, st));
      LAUNCH(k_collect_singles, (unsigned)cdiv(h->nu, 256), 256, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(), h->nu, p->adjacency_min,
             h->singles.as<uint32_t>(), d_scnt);
      uint32_t scnt[2] = {0, 0};
      CK(cudaMemcpyAsync(scnt, d_scnt, 8, cudaMemcpyDeviceToHost, st));
      CK(stream_wait(st));
      t.n_singles = scnt[0];
      CK(t.bnd_out.reserve((size_t)t.n_singles * 8 + 16));
      return VGS_OK;
    });
    if (s) return s;
    std::vector<size_t> bnd_bytes(R, 0);
    bool have_bnd_counts = false;
    size_t bnd_total = 0;
    // One global round = three local rounds of the owned singles (launched back to back, one flag each) + the exchange of
    // the state of the singles next to a cut + the attached flags of the low-id voxels.  The "anything changed" flags of
    // the round ride on the same all-reduce as the attached flags (word LOW_IDS), so a round costs ONE host read.
    while (true) {
      std::vector<void*> sp(L), rp(L), ap(L);
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        uint32_t* d_changed = t.tsmall.as<uint32_t>() + 4;      // [0..2] local rounds, [3] import
        uint32_t* d_nb = t.tsmall.as<uint32_t>() + 12;
        CK(cudaMemsetAsync(d_changed, 0, 16, st));
        CK(cudaMemsetAsync(d_nb, 0, 4, st));
        CK(cudaMemsetAsync(t.low_att_out.p, 0, (LOW_IDS + 1) * 4, st));
        if (t.n_singles > 0) {
          for (int it = 0; it < 3; it++)
            LAUNCH(k_closest_round_tile, (unsigned)cdiv((int64_t)t.n_singles * 32, 128), 128, 0, h->singles.as<uint32_t>(), t.n_singles,
                   h->adj_off.as<uint32_t>(), h->adj_idx.as<int32_t>(), h->conn1_cnt.as<uint32_t>(), h->rec.as<float>(), h->nu, pp, t.own.as<uint8_t>(),
                   t.low_glob.as<LowEntry>(), t.low_att.as<int32_t>(), n_low, t.gidlo.as<uint16_t>(), h->attach.as<int32_t>(), d_changed + it);
          if (R > 1)
            LAUNCH(k_boundary_export, (unsigned)cdiv(t.n_singles, 256), 256, 0, h->singles.as<uint32_t>(), t.n_singles, t.own.as<uint8_t>(),
                   h->plainm.as<unsigned long long>(), h->attach.as<int32_t>(), t.bnd_out.as<unsigned long long>(), d_nb);
        }
        if (R > 1 && h->nu > 0 && n_low > 0)
          LAUNCH(k_low_attached, (unsigned)cdiv(h->nu, 256), 256, 0, t.low_glob.as<LowEntry>(), n_low, t.gidlo.as<uint16_t>(), t.own.as<uint8_t>(),
                 h->attach.as<int32_t>(), h->nu, t.low_att_out.as<int32_t>());
        if (R > 1 && !have_bnd_counts) {
          CK(cudaMemcpyAsync(&t.n_bnd, d_nb, 4, cudaMemcpyDeviceToHost, st));
          CK(stream_wait(st));
        }
        sp[lr] = t.bnd_out.p; ap[lr] = t.low_att_out.p;
        return VGS_OK;
      });
      if (s) return s;
      global_rounds++;
      if (R > 1) {
        if (!have_bnd_counts) {   // the set of exported singles is the same in every round: sizes are exchanged once
          std::vector<unsigned long long> mine(L), allb(R);
          std::vector<const void*> cp(L);
          for (int lr = 0; lr < L; lr++) { mine[lr] = (unsigned long long)g->tc[lr].n_bnd * 8ull; cp[lr] = &mine[lr]; }
          if ((s = xg_host(g, cp.data(), allb.data(), 8))) return s;
          for (int r = 0; r < R; r++) { bnd_bytes[r] = (size_t)allb[r]; bnd_total += bnd_bytes[r]; }
          for (int lr = 0; lr < L; lr++) GCK(g->tc[lr].bnd_all.reserve(bnd_total + 16));
          have_bnd_counts = true;
        }
        for (int lr = 0; lr < L; lr++) rp[lr] = g->tc[lr].bnd_all.p;
        if ((s = xg_dev(g, sp.data(), rp.data(), bnd_bytes.data()))) return s;
      }
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        uint32_t* d_changed = t.tsmall.as<uint32_t>() + 4;
        if (R > 1 && bnd_total > 0 && h->nu > 0)
          LAUNCH(k_boundary_import, (unsigned)cdiv((int64_t)(bnd_total / 8), 256), 256, 0, t.bnd_all.as<unsigned long long>(), (int64_t)(bnd_total / 8),
                 h->tk.as<unsigned long long>(), h->tv.as<uint32_t>(), h->hmask, t.own.as<uint8_t>(), h->attach.as<int32_t>(), d_changed + 3);
        LAUNCH(k_fold_flags, 1, 32, 0, d_changed, 4, t.low_att_out.as<int32_t>() + LOW_IDS);
        return VGS_OK;
      });
      if (s) return s;
      if (R > 1 && (s = xr_dev(g, ap.data(), LOW_IDS + 1, 1))) return s;
      int32_t any = 0;
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        CK(cudaMemcpyAsync(t.low_att.p, t.low_att_out.p, LOW_IDS * 4, cudaMemcpyDeviceToDevice, st));
        if (lr == 0) {
          CK(cudaMemcpyAsync(&any, t.low_att_out.as<int32_t>() + LOW_IDS, 4, cudaMemcpyDeviceToHost, st));
          CK(stream_wait(st));
        }
        return VGS_OK;
      });
      if (s) return s;
      if (!any) break;
      if (global_rounds > 100000) return g->fail(VGS_ERR_LIMIT, "vgs_group_run: closest check did not converge");
    }
    for (int lr = 0; lr < L; lr++) total_singles += g->tc[lr].n_singles;
  }
  mark();   // 6: closest

  // ---- components inside every tile (links of owned voxels only) ----
  s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
    const int64_t nu = h->nu;
    if (nu == 0) return VGS_OK;
    StageTimer tcc(h, &h->tm.components_ms, 8);
    KTimer kcc(h, K_COMPONENTS);
    LAUNCH(k_cc_init, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(), h->conn1_idx.as<int32_t>(),
           h->attach.as<int32_t>(), nu, h->parent.as<int>(), t.own.as<uint8_t>());
    for (int r = 0; r < h->cc_jumps; r++) LAUNCH(k_cc_jump, (unsigned)cdiv(nu, 256), 256, 0, h->parent.as<int>(), nu);
    LAUNCH(k_cc_hook, (unsigned)cdiv(nu * 32, 128), 128, 0, h->adj_off.as<uint32_t>(), h->conn1_cnt.as<uint32_t>(), h->conn1_idx.as<int32_t>(),
           h->attach.as<int32_t>(), nu, h->parent.as<int>(), t.own.as<uint8_t>());
    LAUNCH(k_cc_flatten, (unsigned)cdiv(nu, 256), 256, 0, h->parent.as<int>(), nu, h->root.as<int>());
    kcc.stop();
    tcc.stop();
    h->have_segments = true;
    return VGS_OK;
  });
  if (s) return s;
  mark();   // 7: components

  // ---- cross-slab merge: union-find over the keys of all ranks' (voxel, root) pairs, replicated on every rank ----
  const int min_excl = p->voxels_min;
  size_t n_pairs_total = 0, merge_cap = 0;
  {
    std::vector<unsigned long long> mine(L, 0), allp(R, 0);
    std::vector<void*> sp(L), rp(L);
    // tsmall (u64 words from byte 256): [32] pairs, [33] stat records, [34] label records, [35] clusters
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      if (h->nu == 0) return VGS_OK;
      unsigned long long* d_np = t.tsmall.as<unsigned long long>() + 32;
      CK(cudaMemsetAsync(d_np, 0, 32, st));
      CK(t.pairs.reserve(((size_t)h->nu + (size_t)t.n_singles) * 16 + 16));
      LAUNCH(k_pairs_export, (unsigned)cdiv(h->nu, 256), 256, 0, t.own.as<uint8_t>(), t.gidlo.as<uint16_t>(), h->root.as<int>(), h->plainm.as<unsigned long long>(),
             h->attach.as<int32_t>(), t.low_glob.as<LowEntry>(), h->nu, t.pairs.as<unsigned long long>(), d_np);
      CK(cudaMemcpyAsync(&mine[lr], d_np, 8, cudaMemcpyDeviceToHost, st));
      CK(stream_wait(st));
      return VGS_OK;
    });
    if (s) return s;
    std::vector<const void*> cp(L);
    for (int lr = 0; lr < L; lr++) cp[lr] = &mine[lr];
    if ((s = xg_host(g, cp.data(), allp.data(), 8))) return s;
    std::vector<size_t> bytes(R);
    for (int r = 0; r < R; r++) { bytes[r] = (size_t)allp[r] * 16; n_pairs_total += (size_t)allp[r]; }
    if (n_pairs_total > 0) {
      merge_cap = 64;
      while (merge_cap < n_pairs_total * 4) merge_cap <<= 1;
      if (merge_cap >= (1ull << 31)) return g->fail(VGS_ERR_LIMIT, "vgs_group_run: too many cross-slab pairs");
    }
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      CK(t.pairs_all.reserve(n_pairs_total * 16 + 16));
      sp[lr] = t.pairs.p; rp[lr] = t.pairs_all.p;
      return VGS_OK;
    });
    if (s) return s;
    if ((s = xg_dev(g, sp.data(), rp.data(), bytes.data()))) return s;
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      const int64_t nu = h->nu;
      const uint64_t mask = merge_cap ? merge_cap - 1 : 0;
      if (merge_cap) {
        CK(t.mtk.reserve(merge_cap * 8)); CK(t.mparent.reserve(merge_cap * 4)); CK(t.mminkey.reserve(merge_cap * 8));
        CK(t.tsize.reserve(merge_cap * 4)); CK(t.tminpt.reserve(merge_cap * 4));
        CK(cudaMemsetAsync(t.mtk.p, 0xff, merge_cap * 8, st));
        CK(cudaMemsetAsync(t.mminkey.p, 0xff, merge_cap * 8, st));
        CK(cudaMemsetAsync(t.tsize.p, 0, merge_cap * 4, st));
        CK(cudaMemsetAsync(t.tminpt.p, 0xff, merge_cap * 4, st));
        LAUNCH(k_merge_insert, (unsigned)cdiv((int64_t)n_pairs_total * 2, 256), 256, 0, t.pairs_all.as<unsigned long long>(), (int64_t)n_pairs_total * 2,
               t.mtk.as<unsigned long long>(), mask);
        LAUNCH(k_iota, (unsigned)cdiv((int64_t)merge_cap, 256), 256, 0, t.mparent.as<int>(), (int64_t)merge_cap);
        LAUNCH(k_merge_union, (unsigned)cdiv((int64_t)n_pairs_total, 256), 256, 0, t.pairs_all.as<unsigned long long>(), (int64_t)n_pairs_total,
               t.mtk.as<unsigned long long>(), mask, t.mparent.as<int>());
        LAUNCH(k_merge_minkey, (unsigned)cdiv((int64_t)merge_cap, 256), 256, 0, t.mtk.as<unsigned long long>(), (int64_t)merge_cap, t.mparent.as<int>(),
               t.mminkey.as<unsigned long long>());
      }
      if (nu == 0) return VGS_OK;
      CK(t.rep.reserve((size_t)nu * 8)); CK(t.is_cross.reserve((size_t)nu + 16));
      LAUNCH(k_merge_lookup, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), h->plainm.as<unsigned long long>(), nu, t.mtk.as<unsigned long long>(),
             mask, t.mparent.as<int>(), t.mminkey.as<unsigned long long>(), t.rep.as<unsigned long long>(), t.is_cross.as<uint8_t>());
      // sizes / smallest point index over the owned voxels
      CK(h->csize.reserve((size_t)nu * 4)); CK(h->cminpt.reserve((size_t)nu * 4));
      CK(cudaMemsetAsync(h->csize.p, 0, (size_t)nu * 4, st));
      CK(cudaMemsetAsync(h->cminpt.p, 0xff, (size_t)nu * 4, st));
      LAUNCH(k_tile_stats, (unsigned)cdiv(nu, 256), 256, 0, h->root.as<int>(), t.own.as<uint8_t>(), h->ustart.as<uint32_t>(), h->d_perm,
             h->d_xyz, nu, h->csize.as<uint32_t>(), h->cminpt.as<uint32_t>());
      return VGS_OK;
    });
    if (s) return s;
    if (merge_cap) {
      std::vector<unsigned long long> ms(L, 0), alls(R, 0);
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        if (h->nu == 0) return VGS_OK;
        unsigned long long* d_ns = t.tsmall.as<unsigned long long>() + 33;
        CK(t.stat_out.reserve((size_t)h->nu * sizeof(StatRec) + 16));
        LAUNCH(k_stats_export, (unsigned)cdiv(h->nu, 256), 256, 0, h->root.as<int>(), t.is_cross.as<uint8_t>(), t.rep.as<unsigned long long>(),
               h->csize.as<uint32_t>(), h->cminpt.as<uint32_t>(), h->nu, t.stat_out.as<StatRec>(), d_ns);
        CK(cudaMemcpyAsync(&ms[lr], d_ns, 8, cudaMemcpyDeviceToHost, st));
        CK(stream_wait(st));
        return VGS_OK;
      });
      if (s) return s;
      for (int lr = 0; lr < L; lr++) cp[lr] = &ms[lr];
      if ((s = xg_host(g, cp.data(), alls.data(), 8))) return s;
      size_t nstat = 0;
      for (int r = 0; r < R; r++) { bytes[r] = (size_t)alls[r] * sizeof(StatRec); nstat += (size_t)alls[r]; }
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        CK(t.stat_all.reserve(nstat * sizeof(StatRec) + 16));
        sp[lr] = t.stat_out.p; rp[lr] = t.stat_all.p;
        return VGS_OK;
      });
      if (s) return s;
      if ((s = xg_dev(g, sp.data(), rp.data(), bytes.data()))) return s;
      s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
        if (nstat > 0)
          LAUNCH(k_stats_accumulate, (unsigned)cdiv((int64_t)nstat, 256), 256, 0, t.stat_all.as<StatRec>(), (int64_t)nstat, t.mtk.as<unsigned long long>(),
                 merge_cap - 1, t.tsize.as<uint32_t>(), t.tminpt.as<uint32_t>());
        return VGS_OK;
      });
      if (s) return s;
    }
  }
  mark();   // 8: merge

  // ---- labels of the owned points, sent home to the rank that holds the point's slice ----
  SliceStarts ss{};
  ss.nranks = R;
  for (int r = 0; r <= R; r++) ss.first[r] = first[r];
  int64_t clusters_local = 0;
  {
    std::vector<unsigned long long> mine((size_t)L * R, 0), allc((size_t)R * R, 0);
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      t.n_lab = 0;
      if (h->nu == 0) return VGS_OK;
      const uint64_t mask = merge_cap ? merge_cap - 1 : 0;
      unsigned long long* d_nl = t.tsmall.as<unsigned long long>() + 34;
      CK(t.lab_recs.reserve((size_t)h->n_valid * 8 + 16));
      StageTimer tl(h, &h->tm.labels_ms, 9);
      KTimer kl(h, K_LABELS);
      LAUNCH(k_tile_labels, (unsigned)cdiv(h->n_valid, 256), 256, 0, h->d_perm, h->pos_unit.as<uint32_t>(), h->root.as<int>(), t.own.as<uint8_t>(),
             t.is_cross.as<uint8_t>(), t.rep.as<unsigned long long>(), h->csize.as<uint32_t>(), h->cminpt.as<uint32_t>(), t.mtk.as<unsigned long long>(),
             mask, t.tsize.as<uint32_t>(), t.tminpt.as<uint32_t>(), h->d_xyz, h->n_valid, min_excl, t.lab_recs.as<uint2>(), d_nl);
      // exported clusters: local ones here, merged ones on the first rank only
      LAUNCH(k_group_cluster_count, (unsigned)cdiv(h->nu, 256), 256, 0, h->root.as<int>(), t.is_cross.as<uint8_t>(), h->csize.as<uint32_t>(), h->nu,
             min_excl, d_nl + 1);
      if (merge_cap && g->rank0 + lr == 0)
        LAUNCH(k_group_rep_count, (unsigned)cdiv((int64_t)merge_cap, 256), 256, 0, t.mtk.as<unsigned long long>(), t.mparent.as<int>(),
               t.mminkey.as<unsigned long long>(), t.tsize.as<uint32_t>(), (int64_t)merge_cap, min_excl, d_nl + 1);
      kl.stop();
      tl.stop();
      unsigned long long two[2] = {0, 0};
      CK(cudaMemcpyAsync(two, d_nl, 16, cudaMemcpyDeviceToHost, st));
      CK(stream_wait(st));
      t.n_lab = (int64_t)two[0];
      clusters_local += (int64_t)two[1];
      if (t.n_lab == 0) return VGS_OK;
      const int64_t nw = cdiv(t.n_lab, ROUTE_PER_WARP);
      CK(t.rcnt.reserve((size_t)nw * R * 4 + 16)); CK(t.roff.reserve((size_t)(nw * R + 1) * 4 + 16));
      LAUNCH(k_home_count, (unsigned)cdiv(nw * 32, 256), 256, 0, t.lab_recs.as<uint2>(), t.n_lab, ss, nw, t.rcnt.as<uint32_t>());
      unsigned long long total = 0;
      vgs_status s_ = scan_u32(h, t.rcnt.as<uint32_t>(), t.roff.as<uint32_t>(), nw * R, &total);
      if (s_) return s_;
      std::vector<uint32_t> starts(R + 1, 0);
      CK(cudaMemcpy2DAsync(starts.data(), 4, t.roff.p, (size_t)nw * 4, 4, R, cudaMemcpyDeviceToHost, st));
      CK(stream_wait(st));
      starts[R] = (uint32_t)total;
      for (int d = 0; d < R; d++) mine[(size_t)lr * R + d] = (unsigned long long)(starts[d + 1] - starts[d]) * 8ull;
      CK(t.lab_send.reserve((size_t)total * 8 + 16));
      LAUNCH(k_home_scatter, (unsigned)cdiv(nw * 32, 256), 256, 0, t.lab_recs.as<uint2>(), t.n_lab, ss, nw, t.roff.as<uint32_t>(), t.lab_send.as<uint2>());
      return VGS_OK;
    });
    if (s) return s;
    std::vector<const void*> cp(L);
    for (int lr = 0; lr < L; lr++) cp[lr] = &mine[(size_t)lr * R];
    if ((s = xg_host(g, cp.data(), allc.data(), (size_t)R * 8))) return s;
    std::vector<size_t> cm((size_t)R * R);
    for (size_t i = 0; i < cm.size(); i++) cm[i] = (size_t)allc[i];
    std::vector<void*> sendp(L), recvp(L);
    std::vector<size_t> rbytes(L, 0);
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      for (int r = 0; r < R; r++) rbytes[lr] += cm[(size_t)r * R + g->rank0 + lr];
      CK(t.lab_recv.reserve(rbytes[lr] + 16));
      CK(t.lab_send.reserve(16));
      sendp[lr] = t.lab_send.p; recvp[lr] = t.lab_recv.p;
      return VGS_OK;
    });
    if (s) return s;
    if ((s = xa_dev(g, sendp.data(), recvp.data(), cm))) return s;
    s = each_rank(g, [&](int lr, vgs_handle h, TileCtx& t) -> vgs_status {
      if (t.n == 0) return VGS_OK;
      int32_t* d_out = t.labels_user;
      if (!on_device) { CK(t.labels_dev.reserve((size_t)t.n * 4)); d_out = t.labels_dev.as<int32_t>(); }
      if (!d_out) return h->fail(VGS_ERR_INVALID, "vgs_group_run: null label buffer");
      CK(cudaMemsetAsync(d_out, 0xff, (size_t)t.n * 4, st));     // non-finite / unlabelled points: -1
      const int64_t nr = (int64_t)(rbytes[lr] / 8);
      if (nr > 0)
        LAUNCH(k_labels_scatter, (unsigned)cdiv(nr, 256), 256, 0, t.lab_recv.as<uint2>(), nr, t.gfirst, t.n, d_out);
      if (!on_device) CK(cudaMemcpyAsync(t.labels_user, d_out, (size_t)t.n * 4, cudaMemcpyDeviceToHost, st));
      return VGS_OK;
    });
    if (s) return s;
  }
  mark();   // 9: labels
  GCK(cudaStreamSynchronize(st));

  // ---- counts / timings ----
  for (int lr = 0; lr < L; lr++) resolve_timers(g->hs[lr]);
  {
    long long mine[6] = {sum_tile_pts, sum_tile_vox, clusters_local, total_singles, sum_adj, 0};
    std::vector<long long> all((size_t)R * 6);
    // loopback: the sums already run over all ranks; NCCL: one contribution per process
    if (g->use_nccl) {
      const void* sp[1] = {mine};
      if ((s = xg_host(g, sp, all.data(), sizeof(mine)))) return s;
      for (int k = 0; k < 6; k++) { mine[k] = 0; for (int r = 0; r < R; r++) mine[k] += all[(size_t)r * 6 + k]; }
    }
    vgs_group_counts& c = g->counts;
    memset(&c, 0, sizeof(c));
    c.n_ranks = R; c.n_points = first[R]; c.n_tile_points = mine[0]; c.n_tile_voxels = mine[1]; c.n_clusters_exported = mine[2];
    c.n_singles = mine[3]; c.n_adjacency = mine[4]; c.octree_depth = depth; c.halo = sc.halo; c.axis = sc.axis; c.origin_rounds = origin_rounds;
    c.closest_rounds = global_rounds; c.n_cross_pairs = (int64_t)n_pairs_total;
    for (int r = 0; r <= R && r <= 16; r++) c.cuts[r] = sc.cut[r];
  }
  {
    float ms[16] = {};
    for (int i = 1; i < evi; i++) cudaEventElapsedTime(&ms[i], g->ev[i - 1], g->ev[i]);
    vgs_group_timings& t = g->tm;
    memset(&t, 0, sizeof(t));
    t.origin_ms = ms[1]; t.cuts_ms = ms[2]; t.route_ms = ms[3]; t.tiles_ms = ms[4]; t.low_ms = ms[5]; t.closest_ms = ms[6];
    t.components_ms = ms[7]; t.merge_ms = ms[8]; t.labels_ms = ms[9];
    cudaEventElapsedTime(&t.total_ms, g->ev[0], g->ev[evi - 1]);
    int64_t l = 0;
    for (vgs_handle h : g->hs) l += h->launches;
    t.kernel_launches = l;
  }
  return VGS_OK;
}

}  // extern "C"

/*
 * vgs_b200.h — C ABI of libvgs_b200.so: the VGS / SVGS segmentation hot path of
 * Yusheng-Xu/VGS-SVGS-Segmentation as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary: the C++ classes in include/vgs_dropin/ (same class and member names
 * as the reference's voxel_segmentation.h / supervoxel_segmentation.h) forward to these entry
 * points; INTEGRATION.md shows the binding.  Plain C, opaque handle, int status, caller-owned
 * buffers, no torch / CUDA types in the signatures.  One handle = one CUDA device + one stream (work forked onto
 * the handle's internal side streams — the independent size-class launches of vgs_segment — is joined back into that
 * stream before the call returns); a handle is not thread-safe, different handles are independent.
 *
 * There is NO CPU fallback: every entry point fails with VGS_ERR_NO_DEVICE / VGS_ERR_CUDA when
 * no sm_100 device is usable.
 *
 * Reference citations: VS.h = voxel_segmentation.h, SV.h = supervoxel_segmentation.h,
 * test = /root/reference/test.
 */
#ifndef VGS_B200_H_
#define VGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgs_context* vgs_handle;
typedef int vgs_status;

enum {
  VGS_OK = 0,
  VGS_ERR_INVALID = 1,   /* bad argument */
  VGS_ERR_CUDA = 2,      /* CUDA runtime error, see vgs_last_error */
  VGS_ERR_STATE = 3,     /* stage called out of order (the reference's call order is mandatory, test:51-76) */
  VGS_ERR_LIMIT = 4,     /* input exceeds a documented limit (octree depth > 21, neighbourhood > 181, ...) */
  VGS_ERR_NO_DEVICE = 5  /* no CUDA device: there is no CPU path */
};

enum { VGS_MODE_VGS = 0, VGS_MODE_SVGS = 1 };
enum { VGS_LEAF_DESCENDING = 0 /* PCL 1.8.1 */, VGS_LEAF_ASCENDING = 1 /* PCL >= 1.9 */ };

typedef struct vgs_config {
  int32_t mode;        /* VGS_MODE_VGS | VGS_MODE_SVGS */
  int32_t device;      /* CUDA device ordinal */
  void* stream;        /* cudaStream_t to launch on; NULL = the handle creates its own non-blocking stream
                          (pass cudaStreamLegacy to run on the legacy default stream) */
  int32_t leaf_order;  /* VGS_LEAF_* : octree leaf-iterator direction that defines voxel ids */
  int32_t reserved[5];
} vgs_config;

typedef struct vgs_sigmas {  /* test:28-33 / test:113-118 */
  float sig_p, sig_n, sig_o, sig_e, sig_c, sig_w;
} vgs_sigmas;

typedef struct vgs_timings {  /* CUDA-event milliseconds of the last run of each stage */
  float h2d_ms, origin_ms, voxelize_ms, features_ms, adjacency_ms, graph_ms, mutual_ms, closest_ms,
      components_ms, labels_ms, d2h_ms, total_ms;
  int64_t kernel_launches;   /* kernels of this library launched since vgs_create / last reset */
  float pair_cache_ms;       /* part of graph_ms spent building the pair-weight cache (VGS) */
  float reserved[3];
} vgs_timings;

typedef struct vgs_counts {
  int64_t n_points, n_finite, n_voxels, n_units, n_used, n_adjacency, n_pairs, n_singles, n_attached,
      n_clusters_all, n_clusters_exported, octree_depth, closest_rounds, max_neighbours, reserved[2];
} vgs_counts;

/* blobs for vgs_debug_get (device arrays copied to the caller's host buffer) */
typedef enum vgs_blob_kind {
  VGS_BLOB_POINT_KEY = 1,     /* u32 x3 xN  final octree key per point (0xFFFFFFFF non-finite)   */
  VGS_BLOB_POINT_UNIT = 2,    /* i32 xN     voxel / supervoxel id per point, -1 none             */
  VGS_BLOB_UNIT_KEY = 3,      /* u32 x3 xV  voxel keys in voxel-id order (VGS)                   */
  VGS_BLOB_UNIT_CENTER = 4,   /* f32 x3 xV  voxel centres (VGS)                                  */
  VGS_BLOB_UNIT_OFFSETS = 5,  /* i64 xV+1                                                        */
  VGS_BLOB_UNIT_POINTS = 6,   /* i32        point indices per unit, ascending                    */
  VGS_BLOB_RECORDS = 7,       /* f32 x16 xV centroid3 normal3 eigen8 count flags                 */
  VGS_BLOB_ADJ_OFFSETS = 11,  /* i64 xV+1                                                        */
  VGS_BLOB_ADJ_IDX = 12,      /* i32 xE     (dist2, id) ascending, self first                    */
  VGS_BLOB_CONN0_COUNT = 13,  /* i32 xV     connect-list sizes after the local graph cut         */
  VGS_BLOB_CONN0_IDX = 14,    /* i32 xE     lists stored at the ADJ offsets, ascending ids       */
  VGS_BLOB_CONN1_COUNT = 15,  /* after the mutual-link filter                                    */
  VGS_BLOB_CONN1_IDX = 16,
  VGS_BLOB_ATTACH = 17,       /* i32 xV     closestCheck partner, -1 none                        */
  VGS_BLOB_UNIT_ROOT = 19     /* i32 xV     component root = smallest unit id of the cluster     */
} vgs_blob_kind;

/* -- lifecycle -- */
vgs_status vgs_create(vgs_handle* out, const vgs_config* cfg);
void vgs_destroy(vgs_handle h);
const char* vgs_last_error(vgs_handle h);       /* h may be NULL: error of the last failed vgs_create */
vgs_status vgs_device_count(int* n);            /* VGS_ERR_NO_DEVICE when none */
/* Pooled lifecycle for callers that make one object per cloud (the drop-in classes: the reference constructs a
 * VoxelBasedSegmentation per call of segmentationVGS, test:51).  vgs_release parks the handle — its device buffers,
 * streams and events — in a process-wide pool instead of freeing them; vgs_acquire hands out a parked handle of the same
 * (device, mode, leaf_order) with all per-cloud state reset, or creates one.  Handles made on a caller's stream
 * (cfg->stream != NULL) are never pooled.  vgs_pool_trim destroys every parked handle (call it before cudaDeviceReset). */
vgs_status vgs_acquire(vgs_handle* out, const vgs_config* cfg);
void vgs_release(vgs_handle h);
void vgs_pool_trim(void);

/* -- input: replaces setInputCloud + getCloudPointNum (VS.h:94-102, test:52-53).  xyz = n points,
 *    stride_bytes 12 (packed) or 16 (pcl::PointXYZ).  on_device != 0: xyz is a device pointer that
 *    must stay valid until the run ends (no copy is made); else a host pointer, copied H2D on the handle's stream:
 *    asynchronously when the buffer is pinned, so it must stay valid and unchanged until the next call that
 *    synchronises (vgs_voxelize and every later stage do; vgs_run does before it returns).  Large PAGEABLE host
 *    buffers (here and in the result getters) move through the library's pinned staging ring, filled / drained by a
 *    few host threads — several times the bandwidth of a plain pageable cudaMemcpy; such a call returns only when the
 *    caller's buffer has been read / written completely. -- */
vgs_status vgs_set_points(vgs_handle h, const float* xyz, int64_t n, int stride_bytes, int on_device);

/* -- stage 0+1: replaces OctreePointCloud(res) + addPointsFromInputCloud + setVoxelCenters
 *    (VS.h:84, test:51-60, VS.h:146-189): PCL's dynamic bounding box is emulated exactly. -- */
vgs_status vgs_voxelize(vgs_handle h, float voxel_size);
vgs_status vgs_get_bounding_box(vgs_handle h, double out6[6]);          /* PCL getBoundingBox, test:56 */
/* setBoundingBox VS.h:133-144: the values are narrowed to float and used as the origin of the voxel
 * centres (VS.h:2102-2109); the octree keys themselves are not affected (as in the reference). */
vgs_status vgs_set_bounding_box(vgs_handle h, const double in6[6]);
vgs_status vgs_voxel_count(vgs_handle h, int64_t* n_voxels);            /* getVoxelNum VS.h:104 */
vgs_status vgs_get_voxel_centers(vgs_handle h, float* xyz /* V x 3 */); /* getVoxelCenters VS.h:191 */
/* getOneVoxelAdjacency VS.h:268-287: ids of the units within graph_size of `unit`, nearest first, itself first; *n = list
 * length (ids may be NULL to query it), at most cap ids are written.  Needs vgs_find_adjacency. */
vgs_status vgs_get_unit_adjacency(vgs_handle h, int64_t unit, int32_t* ids, int cap, int* n);

/* -- SVGS units: per-point supervoxel labels as pcl::SupervoxelClustering::getLabeledCloud gives
 *    them (SV.h:283-323; 0 = unlabelled, labels >= max_label are dropped, max_label <= 0 keeps all).
 *    on_device as above. -- */
vgs_status vgs_set_supervoxel_labels(vgs_handle h, const int32_t* label_per_point, int32_t max_label, int on_device);
/* Stand-in for createSupervoxels when no VCCS labels are supplied: one supervoxel per occupied cell of a
 * seed_size grid anchored at the octree origin (deterministic; NOT PCL's VCCS, whose parity is unpinned). */
vgs_status vgs_make_supervoxels_grid(vgs_handle h, float seed_size);
/* createSupervoxels SV.h:245-284: the supervoxel generator itself.  The reference calls PCL's VCCS
 * (pcl::SupervoxelClustering(voxel_resolution, seed_resolution), setColor/Spatial/NormalImportance SV.h:269-271,
 * extract SV.h:277, refineSupervoxels(5) SV.h:278, getLabeledCloud / getMaxLabel SV.h:283-284).  This restates that
 * algorithm with synchronous expansion rounds (deterministic, order independent); voxel_resolution is the size given
 * to vgs_voxelize.  Third-party algorithm: parity with PCL is unpinned, the CPU restatement is oracle/vccs_oracle.cpp.
 * Installs the labels exactly as vgs_set_supervoxel_labels(labels, getMaxLabel()) would. */
vgs_status vgs_make_supervoxels_vccs(vgs_handle h, float seed_resolution, float color_importance, float spatial_importance,
                                     float normal_importance, int refine_iterations);
/* labels currently installed (label_per_point may be NULL) and their max_label */
vgs_status vgs_get_supervoxel_labels(vgs_handle h, int32_t* label_per_point, int32_t* max_label, int on_device);
vgs_status vgs_unit_count(vgs_handle h, int64_t* n_units);              /* getSuperVoxelNum SV.h:118 */

/* -- stage 2: calcualteVoxelCloudAttributes VS.h:290-369 / calcualteSupervoxelCloudAttributes SV.h:1238 -- */
vgs_status vgs_compute_features(vgs_handle h, int points_min);
/* -- stage 3: findAllVoxelAdjacency VS.h:223-265 / findAllSupervoxelNeighbors SV.h:1477-1521 -- */
vgs_status vgs_find_adjacency(vgs_handle h, float graph_size);
/* -- stage 4+5: segmentVoxelCloudWithGraphModel VS.h:372-421 / SV.h:383-420 (local graphs, cut,
 *    crossValidation, closestCheck, clustering) -- */
vgs_status vgs_segment(vgs_handle h, const vgs_sigmas* s, float cut_thred, int adjacency_min);

/* -- results: drawColorMapofPointsinClusters + getClusterIdx / getClusterNum (VS.h:947-1014, 111-121;
 *    SV.h:2109-2126).  voxels_min filters VGS clusters (size > voxels_min); ignored for SVGS. -- */
vgs_status vgs_cluster_count(vgs_handle h, int voxels_min, int64_t* n_all, int64_t* n_exported);
/* canonical labels: label = smallest point index of the point's exported cluster, -1 = not exported.
 * on_device != 0: label_per_point is a device buffer of n int32. */
vgs_status vgs_get_point_labels(vgs_handle h, int voxels_min, int32_t* label_per_point, int on_device);
/* exported clusters as CSR in the reference's cluster order (ascending smallest voxel id); inside a
 * cluster voxels ascend by id and points by index (the reference's DFS member order is not kept).
 * Pass NULL buffers to query sizes: *n_clusters, *n_points_total. */
vgs_status vgs_get_clusters_csr(vgs_handle h, int voxels_min, int64_t* n_clusters, int64_t* n_points_total,
                                int64_t* offsets, int32_t* point_idx);

/* -- whole pipeline in one call (what bench.py times): voxelize .. labels -- */
typedef struct vgs_params {
  float voxel_size, graph_size;
  vgs_sigmas sig;
  float cut_thred;
  int32_t points_min, adjacency_min, voxels_min;
} vgs_params;
vgs_status vgs_run(vgs_handle h, const vgs_params* p, int32_t* label_per_point, int on_device);

/* -- introspection -- */
vgs_status vgs_get_counts(vgs_handle h, vgs_counts* out);
vgs_status vgs_stage_timings(vgs_handle h, vgs_timings* out);
vgs_status vgs_debug_get(vgs_handle h, vgs_blob_kind kind, void* dst, size_t* bytes); /* dst NULL: size only */
/* per-kernel-group CUDA-event times of the last run with the algorithmic HBM bytes each group is accounted with
 * (DESIGN.md section 4); *n = capacity in, number of groups that ran out.  name points to static storage. */
typedef struct vgs_kernel_timing {
  const char* name;
  float ms;
  int32_t launches;
  int64_t alg_bytes;
  int64_t reserved;
} vgs_kernel_timing;
vgs_status vgs_kernel_timings(vgs_handle h, vgs_kernel_timing* out, int* n);

/* ---- ONE scene on several GPUs: spatial slabs of the voxel lattice + halo layers, cross-slab component merge
 *      (DESIGN.md section 6).  The reference is single-process: segmentVoxelCloudWithGraphModel
 *      (VS.h:372-421) on the whole cloud; a group reproduces ITS labels bit for bit — global PCL origin (growth epochs in
 *      global insertion order), global voxel order for closestCheck (VS.h:2181-2303, incl. the COUNT slot read as a voxel
 *      id, VS.h:2243), components merged over voxel keys.  VGS only.
 *      Rank r holds the points [first_r, first_r + n_r) of the cloud (slices in rank order = index order) and gets
 *      back the canonical labels of exactly those points (label = smallest GLOBAL point index of the cluster, -1 none).
 *      Two transports: NCCL (one rank per process; libnccl.so.2 is bound at run time with dlopen) and a loopback group
 *      (all ranks in this process on one device; exchanges are device copies) for single-GPU testing / tiling. ---- */
typedef struct vgs_group_s* vgs_group;
typedef struct vgs_group_counts {
  int64_t n_ranks, n_points, n_tile_points /* owned + halo, all ranks */, n_tile_voxels, n_adjacency, n_singles,
      n_clusters_exported, n_cross_pairs, octree_depth, halo, axis, origin_rounds, closest_rounds;
  int64_t cuts[17];            /* slab r owns key[axis] in [cuts[r], cuts[r+1]) */
  int64_t reserved[2];
} vgs_group_counts;
typedef struct vgs_group_timings {   /* CUDA-event milliseconds on the group's stream (this process) */
  float origin_ms, cuts_ms, route_ms, tiles_ms, low_ms, closest_ms, components_ms, merge_ms, labels_ms, total_ms;
  int64_t kernel_launches;
  float reserved[4];
} vgs_group_timings;
/* rank 0 makes the 128-byte id (ncclGetUniqueId); the caller ships it to the other ranks (torch.distributed, MPI, a file) */
vgs_status vgs_group_unique_id(void* id128);
vgs_status vgs_group_create_nccl(vgs_group* out, const vgs_config* cfg, int nranks, int rank, const void* id128);
vgs_status vgs_group_create_local(vgs_group* out, const vgs_config* cfg, int nranks);
void vgs_group_destroy(vgs_group g);
const char* vgs_group_last_error(vgs_group g);   /* g may be NULL: error of the last failed create */
/* xyz / n / labels: one entry per LOCAL rank (1 for an NCCL group, nranks for a loopback group).  A collective call. */
vgs_status vgs_group_run(vgs_group g, const vgs_params* p, const float* const* xyz, const int64_t* n, int stride_bytes,
                         int on_device, int32_t* const* labels);
vgs_status vgs_group_get_counts(vgs_group g, vgs_group_counts* out);
vgs_status vgs_group_get_timings(vgs_group g, vgs_group_timings* out);
vgs_handle vgs_group_handle(vgs_group g, int local_rank);   /* the rank's handle: per-kernel timings / counts of its tile */
/* host-only: the slab cuts the group derives from per-axis key histograms hist[axis * nbins + (key >> shift)] */
vgs_status vgs_slab_choose_cuts(const uint64_t* hist, int nbins, int shift, int nranks, int* axis, int* cuts /* nranks + 1 */);

#ifdef __cplusplus
}
#endif
#endif /* VGS_B200_H_ */

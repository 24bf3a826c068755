/*
 * vgs_oracle.h — C API of the CPU ORACLE for the VGS / SVGS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the timed CPU baseline).  libvgs_b200.so never links it.
 *
 * PARITY UNPINNED: the reference (Yusheng-Xu/VGS-SVGS-Segmentation) ships no tests, no golden
 * vectors and does not compile (voxel_segmentation.h:2279 calls an undefined function); its
 * arithmetic partly lives in PCL 1.8.1 / FLANN / Eigen which are absent here.  This oracle is a
 * restatement of voxel_segmentation.h / supervoxel_segmentation.h plus the recalled third-party
 * behaviour, pinned only by the analytic known-answer tests of
 * SURVEY.md Appendix C (tests/test_oracle_kat.py).
 */
#ifndef VGS_ORACLE_H_
#define VGS_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgso_params {
  int32_t mode;          /* 0 = VGS (voxel_segmentation.h), 1 = SVGS (supervoxel_segmentation.h) */
  float voxel_size;      /* test:26 / test:109 */
  float graph_size;      /* test:27 / test:111 */
  float sig_p, sig_n, sig_o, sig_e, sig_c, sig_w; /* test:28-33 */
  float cut_thred;       /* test:34 */
  int32_t points_min;    /* test:35 */
  int32_t adjacency_min; /* test:36 */
  int32_t voxels_min;    /* test:37 */
  int32_t leaf_order;    /* 0 = descending x-major Morton (PCL 1.8.1 iterator), 1 = ascending (PCL>=1.9) */
  int32_t math;          /* 0 = glibc float libm as a GCC build would call it,
                            1 = correctly rounded float libm (evaluate in double, round once) */
  float near_tol;        /* relative tolerance for the near-threshold decision report (0 => 1e-5) */
} vgso_params;

typedef struct vgso_handle_s* vgso_handle;

/* blob kinds for vgso_get(); element type in brackets */
enum {
  VGSO_BBOX = 0,           /* [f64 x6] PCL bounding box min xyz, max xyz                        */
  VGSO_POINT_KEY = 1,      /* [u32 x3 xN] final octree key per point (0xFFFFFFFF for non-finite)*/
  VGSO_POINT_UNIT = 2,     /* [i32 xN] voxel id (VGS) / supervoxel id (SVGS) per point, -1 none */
  VGSO_UNIT_KEY = 3,       /* [u32 x3 xV] VGS voxel keys in voxel-id order                      */
  VGSO_UNIT_CENTER = 4,    /* [f32 x3 xV] VGS voxel centres (VS.h:2102-2109)                    */
  VGSO_UNIT_OFFSETS = 5,   /* [i64 xV+1] CSR offsets into VGSO_UNIT_POINTS                      */
  VGSO_UNIT_POINTS = 6,    /* [i32] point indices per unit, ascending                           */
  VGSO_CENTROID = 7,       /* [f32 x3 xV]                                                       */
  VGSO_NORMAL = 8,         /* [f32 x3 xV]                                                       */
  VGSO_EIGEN = 9,          /* [f32 x8 xV] (zeros for unused units)                              */
  VGSO_USED = 10,          /* [u8 xV]                                                           */
  VGSO_ADJ_OFFSETS = 11,   /* [i64 xV+1]                                                        */
  VGSO_ADJ_IDX = 12,       /* [i32 xE] radius neighbours, (dist2, id) ascending, self first     */
  VGSO_CONN0_OFFSETS = 13, /* [i64 xV+1] connect lists after cutGraphSegmentation               */
  VGSO_CONN0_IDX = 14,     /* [i32] ascending ids inside each list (reference order not kept)   */
  VGSO_CONN1_OFFSETS = 15, /* after crossValidation                                             */
  VGSO_CONN1_IDX = 16,
  VGSO_CONN2_OFFSETS = 17, /* after closestCheck                                                */
  VGSO_CONN2_IDX = 18,
  VGSO_UNIT_CLUSTER = 19,  /* [i32 xV] cluster index in reference order (seed = smallest id)    */
  VGSO_POINT_LABEL = 20,   /* [i32 xN] canonical label = min point index of the exported cluster, -1 absent */
  VGSO_CLUSTER_OFFSETS = 21,/* [i64] exported clusters (getClusterIdx), reference DFS order     */
  VGSO_CLUSTER_POINTS = 22,/* [i32]                                                             */
  VGSO_NEAR_EDGES = 23,    /* [i32 x3] (centre unit, unit a, unit b) of near-threshold decisions*/
  VGSO_STATS = 24,         /* [i64 x16] see vgs_oracle.cpp                                      */
  VGSO_ATTACH = 25         /* [i32 xV] closestCheck result: partner id, -1 none                 */
};

vgso_handle vgso_create(const vgso_params* p);
void vgso_destroy(vgso_handle h);

/* xyz: N points, stride_floats floats apart (3 or 4).  labels: SVGS per-point supervoxel labels
 * as pcl::SupervoxelClustering::getLabeledCloud() would give (0 = unlabelled) or NULL for VGS.
 * max_label: getMaxLabel() (SV.h:284); <=0 => max(labels)+1 so that no label is dropped.
 * Returns 0 on success. */
int vgso_run(vgso_handle h, const float* xyz, int64_t n, int stride_floats,
             const int32_t* labels, int32_t max_label);

/* Returns a pointer to oracle-owned storage valid until the next vgso_run/destroy. */
const void* vgso_get(vgso_handle h, int kind, int64_t* count);

/* stand-alone pieces for known-answer tests */
/* out[0..4] = S,A,T,E,C ; out[5] = weight ; eig8 arrays; flags bit0 pos valid, bit1 normal valid, bit2 eigen valid */
void vgso_pair(const vgso_params* p, const float* c1, const float* n1, const float* e1, int flags1,
               const float* c2, const float* n2, const float* e2, int flags2, float* out6);
/* pcl::eigen33 restatement: mat row-major 3x3 -> evals[3] ascending, evecs row-major (columns = vectors) */
void vgso_eigen33(const float* mat9, float* evals3, float* evecs9, int math);
/* unit features from a point list: out = centroid3, normal3, eig8 */
void vgso_features(const vgso_params* p, const float* xyz, int64_t n, float* out14);
/* cutGraphSegmentation on a dense n x n matrix (row-major W[row*n+col]); returns member count,
 * members (local ids ascending) into out */
int vgso_cut(float cut_thred, const float* w, int n, int32_t* out);
/* threads for the per-unit local-graph loop of vgso_run (default 1 = as the reference runs it).  Results do not
 * depend on the value.  Returns the number in effect (1 when the oracle was built without OpenMP). */
int vgso_set_threads(int n);

/* ---- supervoxel generator (vccs_oracle.cpp): restatement of pcl::SupervoxelClustering as the reference
 *      drives it in createSupervoxels (supervoxel_segmentation.h:245-284).  Third-party algorithm, PARITY
 *      UNPINNED (see the header of vccs_oracle.cpp). ---- */
typedef struct vgso_vccs_params {
  float voxel_res, seed_res;                                   /* SV.h:266 */
  float color_importance, spatial_importance, normal_importance; /* SV.h:269-271 */
  int32_t refine_iterations;                                   /* SV.h:278: refineSupervoxels(5, ...) */
  int32_t schedule; /* 0 = PCL's sequential expansion, 1 = synchronous rounds + fixed-point sums (the CUDA schedule) */
} vgso_vccs_params;
/* Inputs: the cloud and its voxel table as vgso_run (mode 0) produced it at voxel_res (keys, CSR of point lists in
 * voxel-id order, origin = bounding-box minimum).  Outputs: label per point (0 = none), getMaxLabel(), optionally the
 * initial voxel normals (V x 3) and the final label per voxel.  Returns the number of seeds. */
int vgso_vccs(const float* xyz, int64_t n, int stride, int64_t n_voxels, const uint32_t* vox_key, const int64_t* vox_off,
              const int32_t* vox_pts, const double* origin3, const vgso_vccs_params* p, int32_t* point_label,
              int32_t* max_label, float* vox_normal, int32_t* vox_label);

#ifdef __cplusplus
}
#endif
#endif

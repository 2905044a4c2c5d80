/*
 * esikf_b200.h — C ABI of the B200-native ESIKF measurement update for FAST-LIVO2.
 *
 * Drop-in boundary for the two calls LIVMapper makes per tick (reference @ 0d2c034):
 *   VoxelMapManager::StateEstimation(StatesGroup&)        include/voxel_map.h:229, called at src/LIVMapper.cpp:370
 *   VIOManager::computeJacobianAndUpdateEKF(cv::Mat img)  include/vio.h:153,      called at src/vio.cpp:1810
 * plus the per-frame producers of the VIO inputs named by the north star:
 *   VIOManager::getImagePatch                              include/vio.h:151  (src/vio.cpp:203-225)
 *   VIOManager::warpAffine / getWarpMatrixAffineHomography include/vio.h:161-162 (src/vio.cpp:252-318)
 *
 * Plain pointers and sizes only; every function returns 0 on success or a negative
 * esikf_status; nothing throws across this boundary. All `double*` state arguments use the
 * packed POD below (fp64, row-major). The library is CUDA-only: there is no CPU fallback, and
 * esikf_create fails with ESIKF_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef ESIKF_B200_H_
#define ESIKF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct esikf_ctx esikf_ctx;

enum esikf_status {
  ESIKF_OK = 0,
  ESIKF_ERR_NO_DEVICE = -1,   /* no CUDA device / wrong architecture */
  ESIKF_ERR_CUDA = -2,        /* a CUDA runtime call failed; see esikf_last_error */
  ESIKF_ERR_ARG = -3,         /* bad argument (null pointer, negative size, key out of range ...) */
  ESIKF_ERR_STATE = -4,       /* call order (e.g. lio_run before map_upload / set_scan) */
  ESIKF_ERR_COMM = -5         /* NCCL / peer-memory set-up failed, or a bounded in-kernel wait (grid barrier, peer mailbox)
                                 expired: the update that reported it is invalid (returned by the fetch calls) */
};

/* StatesGroup (include/common_lib.h:126-223) as a packed POD, 386 doubles:
 *   [0:9) rot_end row-major | [9:12) pos_end | [12] inv_expo_time | [13:16) vel_end |
 *   [16:19) bias_g | [19:22) bias_a | [22:25) gravity | [25:386) cov 19x19 row-major
 * Error-state order (common_lib.h:167-206): dtheta(0:3) dp(3:6) dinv_expo(6) dv(7:10) dbg(10:13) dba(13:16) dg(16:19). */
#define ESIKF_STATE_DOUBLES 386

/* One plane of the adaptive voxel map: VoxelPlane (include/voxel_map.h:69-94) reduced to what
 * build_single_residual / the Jacobian loop read (src/voxel_map.cpp:721-754, 425-449).
 * plane_var holds the upper triangle (row-major, i<=j) of the 6x6 plane_var_. 256 bytes. */
typedef struct esikf_plane {
  double center[3];
  double normal[3];
  double plane_var[21];
  float d;       /* VoxelPlane::d_ */
  float radius;  /* VoxelPlane::radius_ */
  int32_t layer; /* depth of the node inside its root voxel (PointToPlane::layer_) */
  int32_t path;  /* 3 bits per layer: leaf index at layer 1 | leaf index at layer 2 << 3 | ... */
  int32_t pad[6];
} esikf_plane;

/* VoxelMapConfig (include/voxel_map.h:35-52), hot-path subset; loadVoxelConfig src/voxel_map.cpp:36-53 */
typedef struct esikf_lio_cfg {
  double voxel_size;      /* lio/voxel_size      max_voxel_size_ */
  double sigma_num;       /* lio/sigma_num */
  double dept_err;        /* lio/dept_err */
  double beam_err;        /* lio/beam_err */
  int32_t max_layer;      /* lio/max_layer */
  int32_t max_iterations; /* lio/max_iterations (<= 8) */
} esikf_lio_cfg;

/* extrinsics: LIVMapper.cpp:125-126 (extR_/extT_), vio.cpp:29-39,57-65 (Rcl/Pcl) */
typedef struct esikf_extrinsics {
  double extR[9]; /* lidar -> imu rotation, row-major */
  double extT[3];
  double Rcl[9]; /* lidar -> camera */
  double Pcl[3];
} esikf_extrinsics;

/* vk::AbstractCamera as used by the path (cam->world2cam at vio.cpp:1574, fx()/fy() at :45-46).
 * model 0 = Pinhole + radtan d[0..4] (config/camera_pinhole.yaml), 1 = EquidistantCamera k1..k4 in d[0..3]
 * (config/camera_fisheye_HILTI22.yaml). Intrinsics are already multiplied by the camera `scale`. */
typedef struct esikf_camera {
  int32_t model, width, height, pad_;
  double fx, fy, cx, cy;
  double d[5];
} esikf_camera;

/* vio/ parameters (LIVMapper.cpp:50-117): patch_size is fixed at 8 (every shipped config). */
typedef struct esikf_vio_cfg {
  double img_point_cov;
  int32_t patch_pyrimid_level; /* <= 8 */
  int32_t max_iterations;      /* <= 8 */
  int32_t exposure_estimate_en;
  int32_t inverse_composition_en; /* vio/inverse_composition_en (LIVMapper.cpp:60): updateStateInverse instead of updateState;
                                     needs esikf_vio_set_inverse_refs. 0 in every shipped config. */
} esikf_vio_cfg;

/* Per-call diagnostics (what the reference prints at src/voxel_map.cpp:404-405). */
typedef struct esikf_lio_stats {
  int32_t iters;                /* iterations executed */
  int32_t effct_feat_num[8];    /* matched points per iteration (global over all ranks) */
  int32_t converged[8];
  int32_t pad_;
  double total_residual[8];     /* sum |dis_to_plane| per iteration */
  double HTH[8][36];            /* 6x6 H^T R^-1 H per iteration */
  double HTz[8][6];
  double solution[8][19];
} esikf_lio_stats;

typedef struct esikf_vio_stats {
  int32_t total_iters;
  int32_t iters_per_level[8];
  int32_t accepted_per_level[8];
  int32_t pad_;
  float error_trace[8][8];      /* [level][iteration] mean squared photometric error */
  double HTH[8][8][49];         /* [level][iteration] 7x7 H^T H (accepted iterations) */
  double HTz[8][8][7];
  double solution[8][8][19];
} esikf_vio_stats;

/* ---------------------------------------------------------------- context */
int esikf_create(esikf_ctx **out, int device);
void esikf_destroy(esikf_ctx *ctx);
const char *esikf_last_error(const esikf_ctx *ctx);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on. */
void *esikf_stream(esikf_ctx *ctx);
int esikf_synchronize(esikf_ctx *ctx);
/* Page-locked host memory for callers that stage the per-tick buffers themselves (the C++ shim does): copies from / into
 * it are DMA transfers instead of driver-staged pageable copies. NULL on failure; free(NULL) is a no-op. */
void *esikf_host_alloc(size_t bytes);
void esikf_host_free(void *p);
/* Number of kernels launched by this context since creation (bench.py "gpu_launches"). */
int64_t esikf_launch_count(const esikf_ctx *ctx);
/* solve_mode: 0 = Woodbury 6x6/7x7 form of (H^T H + P^-1)^-1 (default), 1 = literal two 19x19
 * partial-pivot inversions as at src/voxel_map.cpp:468 / src/vio.cpp:1661. */
int esikf_set_solve_mode(esikf_ctx *ctx, int mode);
/* loop_mode: how the iteration loop of an update is driven. Results of all modes are bit-identical.
 *   2 (default; 1 is accepted as an alias) = one persistent cooperative kernel per update: every CTA keeps the state on
 *       chip, sums the per-CTA partial vectors after ONE grid barrier per iteration and runs the gain solve itself. With
 *       peer GPUs attached (esikf_peer_attach) the same kernel also carries the NVLink exchange of the information vector.
 *   0 = one residual + one solve launch per iteration (also used with an NCCL communicator or kernel timing on). */
int esikf_set_loop_mode(esikf_ctx *ctx, int mode);
/* Measurement variants, OR-ed flags; 0 = none (default). Same results bit for bit.
 *   ESIKF_TUNE_STAGE_LDG : LIO plane records are brought into shared memory by coalesced half-warp __ldg copies instead of
 *                          one cp.async.bulk (TMA engine) per lane — the round-1 staging, kept to measure against.
 *   ESIKF_TUNE_VIO_TMA   : VIO tap footprints (the 11 x 11 strided taps of a patch, vio.cpp:1595-1631) are brought in by ONE
 *                          tiled TMA load per patch (cp.async.bulk.tensor.2d of 11 image rows, elementStrides {1, s}) instead
 *                          of four byte loads per lane. Needs an image row pitch that is a multiple of 16 bytes; otherwise,
 *                          for tap strides 16 / 32 and for footprints leaving the image the per-lane loads are used. */
#define ESIKF_TUNE_STAGE_LDG 1u
#define ESIKF_TUNE_VIO_TMA 2u
int esikf_set_tuning(esikf_ctx *ctx, uint32_t flags);
int esikf_set_extrinsics(esikf_ctx *ctx, const esikf_extrinsics *ext);
/* Only the lidar -> imu part (extR_, extT_ of VoxelMapManager, LIVMapper.cpp:125-126); the camera part set before is kept.
 * Cheap when nothing changed (no copy, no synchronisation): the LIO shim calls it every tick. */
int esikf_set_lidar_extrinsics(esikf_ctx *ctx, const double extR[9], const double extT[3]);

/* ---------------------------------------------------------------- voxel map mirror
 * Device mirror of `std::unordered_map<VOXEL_LOCATION, VoxelOctoTree*> voxel_map_`
 * (include/voxel_map.h:194). `keys` are n_roots x 3 int64 VOXEL_LOCATIONs; root r owns the
 * ordered candidate planes [first[r], first[r]+count[r]) — the DFS order in which
 * build_single_residual (src/voxel_map.cpp:771-784) visits the octree. */
int esikf_map_upload(esikf_ctx *ctx, const int64_t *keys, const int32_t *first, const int32_t *count, int32_t n_roots,
                     const esikf_plane *planes, int32_t n_planes, double voxel_size);
/* Overwrite planes whose VoxelPlane::is_update_ flag was set by UpdateVoxelMap (voxel_map.h:86). */
int esikf_map_patch(esikf_ctx *ctx, const int32_t *plane_ids, const esikf_plane *planes, int32_t n);

/* ---------------------------------------------------------------- device-resident voxel map
 * VoxelMapManager::BuildVoxelMap / UpdateVoxelMap (src/voxel_map.cpp:532-591, 609-641) with the octrees, the nodes' point
 * lists (temp_points_) and the plane refits (init_plane, :55-135) kept on the GPU: after a LIO update the map absorbs the
 * scan without any plane, point list or key crossing PCIe, and the next esikf_lio_run reads the refitted records in place.
 * Replaces the per-tick host refit + flatten + esikf_map_patch of a host-owned map (esikf_map_upload stays available; the
 * two forms are exclusive: whichever was set up last owns the context's map). */
typedef struct esikf_map_cfg {
  double voxel_size;          /* lio/voxel_size       max_voxel_size_ */
  double min_eigen_value;     /* lio/min_eigen_value  planner_threshold_ */
  double dept_err, beam_err;  /* lio/dept_err, lio/beam_err (BuildVoxelMap's own calcBodyCov, :546) */
  int32_t max_layer;          /* lio/max_layer (<= 7) */
  int32_t max_points_num;     /* lio/max_points_num */
  int32_t layer_init_num[8];  /* lio/layer_init_num, entries 0..max_layer are used */
  int32_t pad;
  /* capacities, fixed at init (nothing is allocated per tick); 0 = default derived from root_capacity (default 2^20):
   * nodes 4 x, plane records 4 x, stored points 64 x the roots. Exceeding one is reported by the update call
   * (ESIKF_ERR_STATE, flags in esikf_map_stats.errors) and invalidates the map. */
  int64_t root_capacity, node_capacity, record_capacity, point_capacity;
} esikf_map_cfg;
typedef struct esikf_map_stats {
  int32_t roots, nodes, records; /* root voxels, octree nodes, record slots handed out (incl. dead blocks of relocated lists) */
  int32_t touched_roots;         /* roots the last build / update replayed */
  int32_t errors;                /* 1 nodes | 2 point pool | 4 records | 8 hash | 16 key outside +-2^20 | 32 octree depth */
  int32_t pad;
  int64_t pool_points;           /* point slots handed out */
} esikf_map_stats;
/* Empty device map with the given configuration (drops any map the context held). */
int esikf_map_device_init(esikf_ctx *ctx, const esikf_map_cfg *cfg);
/* BuildVoxelMap on the resident scan (esikf_lio_set_scan) at pose `state` (packed, host): LIVMapper.cpp:356-366. Needs an
 * empty map. */
int esikf_map_device_build(esikf_ctx *ctx, const double *state);
/* LIVMapper.cpp:413-424: world points / covariances of the resident scan with the posterior of the update that just ran
 * (state == NULL: the one resident on the device; else a packed host state), then UpdateVoxelMap(pv_list_). Call after
 * esikf_lio_run / esikf_lio_update of this scan and before the VIO update overwrites the resident state. */
int esikf_map_device_update(esikf_ctx *ctx, const double *state);
/* UpdateVoxelMap(input_points) with the caller's own lists: point_w [n][3], var [n][9] row-major (host). */
int esikf_map_device_update_points(esikf_ctx *ctx, const double *point_w, const double *var, int32_t n);
/* mapSliding / clearMemOutOfMap (src/voxel_map.cpp:924-971): root voxels whose key lies outside [key_min, key_max]
 * (component-wise, inclusive) are dropped; the survivors are copied into a second arena of the same capacities, which also
 * reclaims dead record blocks and point-list slack (NULL bounds: compaction only). The per-point plane ids of the last
 * update are void afterwards (esikf_lio_fetch normals / ids refer to the old records): call it between ticks. */
int esikf_map_device_slide(esikf_ctx *ctx, const int64_t key_min[3], const int64_t key_max[3]);
int esikf_map_device_stats(esikf_ctx *ctx, esikf_map_stats *out);
/* The map in esikf_map_upload's flat form (roots in no particular order). keys == NULL: sizes only. */
int esikf_map_device_download(esikf_ctx *ctx, int64_t *keys, int32_t *first, int32_t *count, int32_t roots_cap, esikf_plane *planes, int32_t planes_cap,
                              int32_t *n_roots, int32_t *n_planes);

/* ---------------------------------------------------------------- LIO update (StateEstimation)
 * Staged form (device-resident between calls):
 *   set_scan : feats_down_body_ (xyz float32, n points) -> device + per-frame calcBodyCov/crossmat
 *              (src/voxel_map.cpp:349-360). shard_begin/shard_count select this rank's slice.
 *   run      : the iteration loop (src/voxel_map.cpp:372-500) asynchronously on esikf_stream.
 *   fetch    : state_ / stats / per-point results back to host (synchronises the stream). */
int esikf_lio_set_scan(esikf_ctx *ctx, const float *pts_xyz, int32_t n);
int esikf_lio_run(esikf_ctx *ctx, const double *state_in, const double *state_prop, const esikf_lio_cfg *cfg);
int esikf_lio_fetch(esikf_ctx *ctx, double *state_out, esikf_lio_stats *stats /* nullable */,
                    int32_t *match_plane /* n, plane id matched in the last iteration or -1: ptpl_list_ */,
                    int32_t *normal_plane /* n, plane id behind pv.normal (sticky across iterations) or -1 */,
                    float *dis_to_plane /* n, PointToPlane::dis_to_plane_ of the last iteration */);
/* One-shot form with host buffers = set_scan + run + fetch (what the C++ shim calls). */
int esikf_lio_update(esikf_ctx *ctx, const float *pts_xyz, int32_t n, const double *state_in, const double *state_prop,
                     const esikf_lio_cfg *cfg, double *state_out, esikf_lio_stats *stats, int32_t *match_plane,
                     int32_t *normal_plane, float *dis_to_plane);
/* body_cov_list_ / cross_mat_list_ (include/voxel_map.h:215-216) of the current scan: n x 9 doubles each. */
/* pv_list_[i].normal of the last update (src/voxel_map.cpp:744; zero for an unmatched point): [n][3] doubles. Valid after the
 * map absorbed the scan too (esikf_map_device_update snapshots the normals before plane records may move). */
int esikf_lio_fetch_normals(esikf_ctx *ctx, double *normals);
int esikf_lio_fetch_point_cov(esikf_ctx *ctx, double *body_cov9 /* nullable */, double *cross_mat9 /* nullable */);

/* ---------------------------------------------------------------- VIO update (computeJacobianAndUpdateEKF)
 * set_camera    : vk camera + vio/ parameters.
 * set_image     : the current 8-bit grey image, width x height of the camera.
 * set_patches   : visual_submap (include/vio.h:26-57): voxel_points[i]->pos_ (n x 3 f64), warp_patch
 *                 (n x levels*64 f32, [level][row][col]), search_levels (n), inv_expo_list (n).
 * run / fetch   : the coarse-to-fine loop (src/vio.cpp:784-802, 1520-1688). */
int esikf_vio_set_camera(esikf_ctx *ctx, const esikf_camera *cam, const esikf_vio_cfg *cfg);
int esikf_vio_set_image(esikf_ctx *ctx, const uint8_t *img, int32_t width, int32_t height);
int esikf_vio_set_patches(esikf_ctx *ctx, const double *pos, const float *warp_patch, const int32_t *search_levels,
                          const double *inv_expo_list, int32_t n);
int esikf_vio_run(esikf_ctx *ctx, const double *state_in, const double *state_prop);
int esikf_vio_fetch(esikf_ctx *ctx, double *state_out, esikf_vio_stats *stats /* nullable */, float *errors /* n, nullable */);
int esikf_vio_update(esikf_ctx *ctx, const uint8_t *img, int32_t width, int32_t height, const double *pos,
                     const float *warp_patch, const int32_t *search_levels, const double *inv_expo_list, int32_t n,
                     const double *state_in, const double *state_prop, double *state_out, esikf_vio_stats *stats,
                     float *errors);

/* ---------------------------------------------------------------- per-frame patch producers
 * get_image_patch: batched getImagePatch (src/vio.cpp:203-225) on the image set by set_image:
 *                  pc n x 2 f64 -> patch n x 64 f32 at `level`.
 * warp_patches   : batched getWarpMatrixAffineHomography + getBestSearchLevel + warpAffine over all
 *                  pyramid levels (src/vio.cpp:701-714, 739-742). Reference images are registered
 *                  once (Feature::img_, include/feature.h) and addressed by index. */
int esikf_vio_get_image_patch(esikf_ctx *ctx, const double *pc, int32_t n, int32_t level, float *patch_out);
int esikf_vio_set_ref_images(esikf_ctx *ctx, const uint8_t *const *imgs, int32_t n_imgs, int32_t width, int32_t height);
int esikf_vio_warp_patches(esikf_ctx *ctx, int32_t n, const int32_t *ref_img_index, const double *px_ref /* n x 2 */,
                           const double *pos_w /* n x 3 */, const double *normal_w /* n x 3 */,
                           const double *T_ref_w /* n x 12: R(9) t(3) of Feature::T_f_w_ */,
                           const double *T_cur_w /* 12: new_frame_->T_f_w_ */, double *A_cur_ref_out /* n x 4, nullable */,
                           int32_t *search_level_out /* n */, float *warp_patch_out /* n x levels*64 */,
                           int32_t keep_on_device /* 1: also install as the patches of set_patches */);
/* set_inverse_refs: what the inverse-compositional variant (src/vio.cpp:1327-1518) reads of every visual point's reference
 *                  feature (include/feature.h): image (index into set_ref_images), px_, f_ (unit bearing), the rotation
 *                  of T_f_w_ (row-major) and pos() = T_f_w_.inverse().translation(). n must equal the n of set_patches. */
int esikf_vio_set_inverse_refs(esikf_ctx *ctx, int32_t n, const int32_t *ref_img_index, const double *ref_px /* n x 2 */,
                               const double *ref_f /* n x 3 */, const double *ref_R /* n x 9 */, const double *ref_pos /* n x 3 */);
/* warp_affine    : batched warpAffine alone (src/vio.cpp:292-318, include/vio.h:161-162) with caller-provided affine
 *                  matrices A_cur_ref (n x 4, row-major [a00 a01 a10 a11]) and search levels; writes all pyramid levels,
 *                  patch i level l at warp_patch_out[(i*levels + l)*64 ...] like visual_submap->warp_patch[i]. */
int esikf_vio_warp_affine(esikf_ctx *ctx, int32_t n, const int32_t *ref_img_index, const double *px_ref /* n x 2 */,
                          const double *A_cur_ref /* n x 4 */, const int32_t *search_level /* n */,
                          float *warp_patch_out /* n x levels*64 */);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU)
 * The residual point / patch set is sharded by contiguous blocks; each iteration all-reduces the
 * packed information buffer [HTH upper | HTz | M | sum|d| ...] and every rank solves redundantly.
 * esikf_comm_unique_id: rank 0 creates the NCCL id, the host side broadcasts it (torch.distributed).
 * esikf_comm_init     : ncclCommInitRank on this context. shard = [begin, begin+count) of the scan /
 *                       patch arrays given to set_scan / set_patches (all ranks pass the full arrays).
 * With more than one rank the per-point / per-patch outputs of a rank (match_plane, normal_plane, dis_to_plane,
 * errors) are valid on that rank's esikf_shard_range slice only; the state and the stats are identical on every rank. */
int esikf_comm_unique_id(char out[128]);
int esikf_comm_init(esikf_ctx *ctx, int32_t rank, int32_t nranks, const char unique_id[128]);
int esikf_comm_rank(const esikf_ctx *ctx, int32_t *rank, int32_t *nranks);
/* NVLink peer-memory path (no NCCL in the loop): every rank exports a CUDA-IPC handle of its 72-double mailbox, the host
 * side all-gathers the 64-byte handles, esikf_peer_attach maps the peers' mailboxes. The persistent update kernels then
 * all-reduce the information buffer themselves (stores into every peer's mailbox + flag spin) and stay one launch per
 * update. <= 8 ranks of one NVSwitch box; every rank must issue the same sequence of updates. */
int esikf_peer_export(esikf_ctx *ctx, char out[64]);
int esikf_peer_attach(esikf_ctx *ctx, int32_t rank, int32_t nranks, const char *handles /* nranks x 64 bytes */);
/* The contiguous slice [begin, begin+count) of n units owned by `rank` (host-only helper, no device needed). */
int esikf_shard_range(int32_t n, int32_t rank, int32_t nranks, int32_t *begin, int32_t *count);

/* ---------------------------------------------------------------- measurement hooks (bench.py)
 * Average device time (ms, CUDA events on esikf_stream) of `reps` back-to-back launches of one
 * kernel on the resident inputs: which = 0 LIO residual kernel, 1 solve kernel, 2 VIO patch kernel
 * (level given by `arg`), 3 per-frame LIO precompute. flush_l2 != 0 writes a >L2 buffer between launches
 * (excluded from the timing). */
int esikf_profile_kernel(esikf_ctx *ctx, int32_t which, int32_t arg, int32_t reps, int32_t flush_l2, float *avg_ms);

/* Per-launch device times of the update loops. enable != 0 brackets every residual / patch / solve launch of the
 * following lio_run / vio_run calls with CUDA events on esikf_stream (adds sub-microsecond gaps: use a separate,
 * untimed pass). get: ms of the launches of the LAST run, in launch order; LIO: residual[8], solve[8];
 * VIO: patch[64], solve[64] indexed (levels-1-level)*max_iterations + iteration. Unlaunched slots are 0. */
int esikf_set_kernel_timing(esikf_ctx *ctx, int32_t enable);
int esikf_get_kernel_timing(esikf_ctx *ctx, float *lio_residual_ms /* 8 */, float *lio_solve_ms /* 8 */,
                            float *vio_patch_ms /* 64 */, float *vio_solve_ms /* 64 */);

/* Phase timestamps of the persistent update kernels (ns, %globaltimer, written by CTA 0): 8 per iteration slot
 * [start, consts loaded, slice done, all CTAs arrived, partials summed, solved, state published, spare];
 * slots 0..7 = LIO iterations, 8..71 = VIO (levels-1-level)*max_iterations + iteration. 0 = slot not executed. */
int esikf_set_phase_stamps(esikf_ctx *ctx, int32_t enable);
int esikf_get_phase_stamps(esikf_ctx *ctx, uint64_t *out /* 800: 576 phase stamps + 64 fine-grained debug stamps + 160 per-CTA slice-end stamps of LIO iteration 3 */);

#ifdef __cplusplus
}
#endif
#endif /* ESIKF_B200_H_ */

// fl2_shim.hpp — C++ host side of the drop-in boundary: mirrors of the reference's VoxelMapManager / VIOManager call
// surface (same member names, same call order, same "void return + public member fields" convention) on top of the C ABI
// of include/esikf_b200.h. The reference's own headers need Eigen / PCL / ROS / OpenCV / vikit (absent from this image),
// so the boundary types are restated here as layout-compatible PODs; INTEGRATION.md shows the two-line adaptation from the
// real Eigen types (Eigen fixed-size matrices are plain arrays of doubles).
//
// Reference interface being replaced (hku-mars/FAST-LIVO2 @ 0d2c034):
//   VoxelMapManager::StateEstimation(StatesGroup &)                    include/voxel_map.h:229   src/voxel_map.cpp:338-511
//   VoxelMapManager::{state_, feats_down_body_, feats_down_size_, pv_list_, ptpl_list_, effct_feat_num_, position_last_,
//                     cross_mat_list_, body_cov_list_, extR_, extT_, config_setting_, voxel_map_}   include/voxel_map.h:187-218
//   VIOManager::computeJacobianAndUpdateEKF(cv::Mat)                   include/vio.h:153         src/vio.cpp:784-802
//   VIOManager::{state, state_propagat, visual_submap, total_points, G, H_T_H, ...}                  include/vio.h
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/esikf_b200.h"

namespace fl2b200 {

struct V3D { double v[3] = {0, 0, 0}; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; } };
struct M3D { double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; double &operator()(int r, int c) { return m[3 * r + c]; } const double &operator()(int r, int c) const { return m[3 * r + c]; } };  // row-major

// include/common_lib.h:126-223
struct StatesGroup {
  M3D rot_end;
  V3D pos_end, vel_end;
  double inv_expo_time = 1.0;
  V3D bias_g, bias_a, gravity;
  double cov[19 * 19];  // row-major
  StatesGroup();
  void pack(double *out386) const;
  void unpack(const double *in386);
};

// include/common_lib.h:102-123 (fields the hot path reads / writes)
struct pointWithVar {
  V3D point_b, point_i, point_w;
  M3D var_nostate, body_var, var, point_crossmat;
  V3D normal;
  pointWithVar();
};

// include/voxel_map.h:54-67
struct PointToPlane {
  V3D point_b_, point_w_, normal_, center_;
  double plane_var_[36];
  M3D body_cov_;
  int layer_ = 0;
  double d_ = 0, eigen_value_ = 0;
  bool is_valid_ = false;
  float dis_to_plane_ = 0;
};

// include/voxel_map.h:69-94
struct VoxelPlane {
  V3D center_, normal_, y_normal_, x_normal_;
  M3D covariance_;
  double plane_var_[36] = {0};  // row-major 6x6
  float radius_ = 0, min_eigen_value_ = 1, mid_eigen_value_ = 1, max_eigen_value_ = 1, d_ = 0;
  int points_size_ = 0;
  bool is_plane_ = false, is_init_ = false;
  int id_ = 0;
  bool is_update_ = false;
};

// include/voxel_map.h:96-118
struct VOXEL_LOCATION {
  int64_t x, y, z;
  VOXEL_LOCATION(int64_t vx = 0, int64_t vy = 0, int64_t vz = 0) : x(vx), y(vy), z(vz) {}
  bool operator==(const VOXEL_LOCATION &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelLocationHash {
  int64_t operator()(const VOXEL_LOCATION &s) const {
    const int64_t P = 116101, N = 10000000000LL;
    return ((((s.z) * P) % N + (s.y)) * P) % N + (s.x);
  }
};

// include/voxel_map.h:129-183 (structure only; the map bookkeeping stays with the reference's own code)
struct VoxelOctoTree {
  VoxelPlane *plane_ptr_ = nullptr;
  int layer_ = 0;
  VoxelOctoTree *leaves_[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double voxel_center_[3] = {0, 0, 0};
  float quater_length_ = 0;
  bool init_octo_ = false;
  ~VoxelOctoTree();
};
typedef std::unordered_map<VOXEL_LOCATION, VoxelOctoTree *, VoxelLocationHash> VoxelMap;

// include/voxel_map.h:35-52
struct VoxelMapConfig {
  double max_voxel_size_ = 0.5;
  int max_layer_ = 2;
  int max_iterations_ = 5;
  double beam_err_ = 0.05, dept_err_ = 0.02, sigma_num_ = 3.0;
  // map-building part (read by the device-resident map only; a host-owned map is built by the reference's own code)
  double planner_threshold_ = 0.0025;  // lio/min_eigen_value
  int max_points_num_ = 50;            // lio/max_points_num
  std::vector<int> layer_init_num_{5, 5, 5, 5, 5};
  long long device_root_capacity_ = 0;  // 0: esikf_map_device_init's default (2^20 root voxels)
  bool map_sliding_en = false;          // local_map/map_sliding_en
  int half_map_size = 100;              // local_map/half_map_size
  double sliding_thresh = 8.0;          // local_map/sliding_thresh
};

struct PointXYZ { float x, y, z; };  // pcl::PointXYZINormal's xyz as consumed at src/voxel_map.cpp:351,520-521

// The flattened map handed to esikf_map_upload: root keys + DFS-ordered candidate planes.
struct FlatVoxelMap {
  std::vector<int64_t> keys;
  std::vector<int32_t> first, count;
  std::vector<esikf_plane> planes;
  std::vector<const VoxelPlane *> plane_src;  // flat plane id -> the VoxelPlane it mirrors (for map_patch / pv.normal)
};
// DFS of every root in leaf order 0..7, plane nodes terminate their branch: exactly what build_single_residual visits
// (src/voxel_map.cpp:721, 771-784). Throws nothing; returns false with `err` set on inconsistent roots.
bool FlattenVoxelMap(const VoxelMap &map, const VoxelMapConfig &cfg, FlatVoxelMap &out, std::string *err);
// What changed between two flattenings of the same map: false if the structure differs (roots added / removed, a candidate
// list grew or shrank or moved — needs a full upload), true with the ids of the plane records whose content changed
// (refitted planes: what esikf_map_patch takes). is_update_ is not used: the reference sets it on every fit and never clears it.
bool DiffFlatVoxelMaps(const FlatVoxelMap &synced, const FlatVoxelMap &now, std::vector<int32_t> &changed_ids);

// Grow-only page-locked host buffer (esikf_host_alloc): the staging areas the shim hands to the C ABI, so that every
// per-tick copy is a DMA from / into pinned memory instead of a driver-staged pageable copy.
template <typename T> class PinnedBuf {
 public:
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { esikf_host_free(p_); }
  T *get(size_t n) {
    if (n > cap_) {
      esikf_host_free(p_);
      cap_ = n + n / 4 + 16;
      p_ = static_cast<T *>(esikf_host_alloc(cap_ * sizeof(T)));
      if (!p_) cap_ = 0;
    }
    return p_;
  }
  T *data() { return p_; }

 private:
  T *p_ = nullptr;
  size_t cap_ = 0;
};

class VoxelMapManager {
 public:
  VoxelMapConfig config_setting_;
  VoxelMap &voxel_map_;
  std::vector<PointXYZ> feats_down_body_;
  int feats_down_size_ = 0;
  int effct_feat_num_ = 0;
  M3D extR_;
  V3D extT_;
  StatesGroup state_;
  V3D position_last_;
  std::vector<M3D> cross_mat_list_, body_cov_list_;
  std::vector<pointWithVar> pv_list_;
  std::vector<PointToPlane> ptpl_list_;
  bool fill_point_lists_ = true;   // pv_list_ / ptpl_list_ / cross_mat_list_ / body_cov_list_ filled by StateEstimation (off: state_ only)
  bool lazy_point_lists_ = false;  // with fill_point_lists_: fill them on MaterializePointLists() instead of inside StateEstimation —
                                   // the 14 MB device->host copy of the per-point covariances and the host loops over the scan
                                   // only happen for callers that read the lists (LIVMapper's host-side UpdateVoxelMap does)
  int last_status_ = 0;            // esikf_status of the last call (the reference's calls return void)
  int last_iters_ = 0;             // iterations executed by the last StateEstimation
  std::string last_error_;

  VoxelMapManager(VoxelMapConfig &config_setting, VoxelMap &voxel_map, int device = 0);
  ~VoxelMapManager();
  // call after BuildVoxelMap / UpdateVoxelMap / mapSliding changed voxel_map_: re-flatten, then patch the refitted plane
  // records in place (esikf_map_patch) when the candidate lists kept their shape, full upload (esikf_map_upload) otherwise
  void SyncDeviceMap();
  void MarkMapDirty() { map_synced_ = false; }
  int last_sync_patched_ = -1;     // planes patched by the last SyncDeviceMap, -1 = it was a full upload
  void StateEstimation(StatesGroup &state_propagat);  // include/voxel_map.h:229
  // ---- device-resident map: the octrees, point lists and refits live on the GPU (esikf_map_device_*); voxel_map_ is not used
  // and SyncDeviceMap is never needed. EnableDeviceMap once after construction, then the reference's own call sequence:
  //   BuildVoxelMap()              first LiDAR frame, from feats_down_body_ and state_      (include/voxel_map.h:231, LIVMapper.cpp:356-366)
  //   UpdateVoxelMap()             LIVMapper.cpp:413-424 in ONE call: world points + covariances with the posterior and the
  //                                update, all on the device (nothing but a status crosses PCIe)
  //   UpdateVoxelMap(input_points) the reference's signature (include/voxel_map.h:232) with host lists
  void EnableDeviceMap();
  bool device_map_ = false;
  void BuildVoxelMap();
  void UpdateVoxelMap();
  void UpdateVoxelMap(const std::vector<pointWithVar> &input_points);
  void mapSliding();  // include/voxel_map.h:247 / src/voxel_map.cpp:924-948 (device-resident map); uses position_last_, sliding_thresh, half_map_size
  V3D last_slide_position;
  void MaterializePointLists();                       // fills the four lists from the last StateEstimation (idempotent per call of it)
  esikf_ctx *context() { return ctx_; }

 private:
  esikf_ctx *ctx_ = nullptr;
  FlatVoxelMap flat_;
  bool map_synced_ = false, device_has_map_ = false;
  PinnedBuf<float> st_pts_, st_dis_;
  PinnedBuf<int32_t> st_match_, st_normal_;
  PinnedBuf<double> st_state_, st_cov_, st_normals_;
  bool lists_pending_ = false;
};

// include/vio.h:26-57 restated over flat storage
struct SubSparseMap {
  std::vector<float> errors;
  std::vector<std::vector<float>> warp_patch;  // [i][level*64 + row*8 + col]
  std::vector<int> search_levels;
  std::vector<V3D> voxel_points_pos;           // voxel_points[i]->pos_
  std::vector<double> inv_expo_list;
};

struct GrayImage { const uint8_t *data = nullptr; int cols = 0, rows = 0; };  // cv::Mat CV_8UC1 continuous

class VIOManager {
 public:
  StatesGroup *state = nullptr, *state_propagat = nullptr;  // raw pointers into LIVMapper (LIVMapper.cpp:135-136)
  SubSparseMap *visual_submap = nullptr;
  int total_points = 0;
  int patch_size = 8, patch_pyrimid_level = 4, max_iterations = 5;
  double img_point_cov = 100;
  bool exposure_estimate_en = true;
  esikf_camera cam{};
  M3D Rcl;
  V3D Pcl;
  M3D extR;  // setImuToLidarExtrinsic / setLidarToCameraExtrinsic (src/vio.cpp:29-39)
  V3D extT;
  int last_status_ = 0;
  int last_total_iters_ = 0;  // iterations executed by the last computeJacobianAndUpdateEKF
  // The per-patch mirrors below are called once per point and level by the reference's retrieval code: the image they are
  // handed is uploaded only when it is not the one of the previous call (same pixel pointer and size). A caller that rewrites
  // the pixels in place calls InvalidatePatchImages().
  void InvalidatePatchImages() { patch_img_ = patch_ref_img_ = nullptr; }
  std::string last_error_;

  explicit VIOManager(esikf_ctx *shared_ctx);  // shares the device context (and stream) of the VoxelMapManager
  void initializeVIO();                         // src/vio.cpp:41-160 (the parts the update needs)
  void computeJacobianAndUpdateEKF(const GrayImage &img);  // include/vio.h:153
  // Per-patch helpers with the reference's signatures (include/vio.h:151, 161-162; V2D / Matrix2d as plain arrays). One
  // launch and one small read-back per call: the frame path should use the batched esikf_vio_get_image_patch /
  // esikf_vio_warp_patches.
  void getImagePatch(const GrayImage &img, const double pc[2], float *patch_tmp, int level);
  void warpAffine(const double A_cur_ref[4] /* row-major 2x2 */, const GrayImage &img_ref, const double px_ref[2], int level_ref, int search_level,
                  int pyramid_level, int halfpatch_size, float *patch);

 private:
  const uint8_t *patch_img_ = nullptr, *patch_ref_img_ = nullptr;  // what the per-patch helpers uploaded last
  int patch_img_w_ = 0, patch_img_h_ = 0, patch_ref_w_ = 0, patch_ref_h_ = 0;
  esikf_ctx *ctx_ = nullptr;
  PinnedBuf<double> st_pos_, st_ie_, st_state_;
  PinnedBuf<float> st_wp_, st_err_;
  PinnedBuf<int32_t> st_sl_;
  PinnedBuf<uint8_t> st_img_;
};

}  // namespace fl2b200

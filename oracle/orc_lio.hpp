// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header). PINNED to the reference source: oracle/_ref/libfl2_ref_lio.so is
// the reference's own src/voxel_map.cpp compiled against stand-in headers (oracle/ref_voxel_map.cpp) and
// tests/test_oracle_ref_pin.py holds StateEstimation of this restatement to it bit for bit (associations) / to 1e-12 (states).
// The map-construction part (BuildVoxelMap / UpdateVoxelMap / init_plane) is compiled into that library too but is exercised by
// the pin only through the maps it is handed; it is cross-checked against the numpy builder of the generator instead.
//
// CPU restatement of the reference's LIO ESIKF measurement update and of the voxel-map
// construction that manufactures its input. Every function cites the reference lines it
// follows (paths relative to hku-mars/FAST-LIVO2 @ 0d2c034).
#pragma once
#include <mutex>
#include <unordered_map>
#include "orc_math.hpp"

namespace orc {

typedef Mat<6, 6> M6;

// include/common_lib.h:102-123 (same fields, same ~384-byte footprint)
struct pointWithVar {
  V3 point_b, point_i, point_w;
  M3 var_nostate, body_var, var, point_crossmat;
  V3 normal;
  pointWithVar() {
    var_nostate = var = body_var = point_crossmat = M3::Zero();
    point_b = point_i = point_w = normal = V3::Zero();
  }
};

// include/voxel_map.h:54-67
struct PointToPlane {
  V3 point_b_, point_w_, normal_, center_;
  M6 plane_var_;
  M3 body_cov_;
  int layer_;
  double d_;
  double eigen_value_;
  bool is_valid_;
  float dis_to_plane_;
  int plane_id_;  // oracle bookkeeping only: index of the plane in the flat map (-1 if none)
};

// include/voxel_map.h:69-94
struct VoxelPlane {
  V3 center_, normal_, y_normal_, x_normal_;
  M3 covariance_;
  M6 plane_var_;
  float radius_ = 0;
  float min_eigen_value_ = 1, mid_eigen_value_ = 1, max_eigen_value_ = 1;
  float d_ = 0;
  int points_size_ = 0;
  bool is_plane_ = false;
  bool is_init_ = false;
  int id_ = 0;
  bool is_update_ = false;
  int flat_id_ = -1;  // oracle bookkeeping: index in the flat map
  VoxelPlane() {
    plane_var_ = M6::Zero();
    covariance_ = M3::Zero();
    center_ = normal_ = y_normal_ = x_normal_ = V3::Zero();
  }
};

// include/voxel_map.h:96-118
struct VOXEL_LOCATION {
  int64_t x, y, z;
  VOXEL_LOCATION(int64_t vx = 0, int64_t vy = 0, int64_t vz = 0) : x(vx), y(vy), z(vz) {}
  bool operator==(const VOXEL_LOCATION &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelHash {
  int64_t operator()(const VOXEL_LOCATION &s) const {
    const int64_t P = 116101, N = 10000000000LL;
    return ((((s.z) * P) % N + (s.y)) * P) % N + (s.x);
  }
};

struct VoxelMapConfig {  // include/voxel_map.h:35-52 (hot-path subset)
  double max_voxel_size_ = 0.5;
  int max_layer_ = 2;
  int max_iterations_ = 5;
  std::vector<int> layer_init_num_{5, 5, 5, 5, 5};
  int max_points_num_ = 50;
  double planner_threshold_ = 0.0025;
  double beam_err_ = 0.05;
  double dept_err_ = 0.02;
  double sigma_num_ = 3.0;
};

// include/voxel_map.h:129-183
struct VoxelOctoTree {
  std::vector<pointWithVar> temp_points_;
  VoxelPlane *plane_ptr_;
  int layer_;
  int octo_state_;
  VoxelOctoTree *leaves_[8];
  double voxel_center_[3];
  std::vector<int> layer_init_num_;
  float quater_length_;
  float planer_threshold_;
  int points_size_threshold_;
  int update_size_threshold_;
  int max_points_num_;
  int max_layer_;
  int new_points_;
  bool init_octo_;
  bool update_enable_;
  VoxelOctoTree(int max_layer, int layer, int points_size_threshold, int max_points_num, float planer_threshold)
      : layer_(layer), planer_threshold_(planer_threshold), points_size_threshold_(points_size_threshold),
        max_points_num_(max_points_num), max_layer_(max_layer) {
    octo_state_ = 0;
    new_points_ = 0;
    update_size_threshold_ = 5;
    init_octo_ = false;
    update_enable_ = true;
    for (int i = 0; i < 8; i++) leaves_[i] = nullptr;
    plane_ptr_ = new VoxelPlane;
  }
  ~VoxelOctoTree() {
    for (int i = 0; i < 8; i++) delete leaves_[i];
    delete plane_ptr_;
  }
  void init_plane(const std::vector<pointWithVar> &points, VoxelPlane *plane);
  void init_octo_tree();
  void cut_octo_tree();
  void UpdateOctoTree(const pointWithVar &pv);
};

typedef std::unordered_map<VOXEL_LOCATION, VoxelOctoTree *, VoxelHash> VoxelMap;

void calcBodyCov(V3 &pb, const float range_inc, const float degree_inc, M3 &cov);

// Flat plane record shared with the product boundary (include/esikf_b200.h, esikf_plane_t).
struct FlatPlane {
  double center[3];
  double normal[3];
  double plane_var[21];  // upper triangle, row-major
  float d;
  float radius;
  int32_t layer;
  int32_t path;  // 3 bits per layer: leaf index at layer 1 | (leaf index at layer 2) << 3 | ...
  int32_t pad[6];
};
static_assert(sizeof(FlatPlane) == 256, "flat plane record is 256 bytes");

struct LioStats {
  int iters;               // iterations executed
  int effct_feat_num[8];   // per iteration
  double total_residual[8];
  double HTH[8][36];       // per iteration 6x6 H^T R^-1 H
  double HTz[8][6];
  double solution[8][19];
  int converged[8];
};

class VoxelMapManager {  // include/voxel_map.h:187-256 (hot-path subset)
 public:
  VoxelMapConfig config_setting_;
  VoxelMap voxel_map_;
  std::vector<float> feats_down_body_;   // xyz f32, N*3 (PCL cloud in the reference)
  std::vector<float> feats_down_world_;  // xyz f32
  M3 extR_;
  V3 extT_;
  StatesGroup state_;
  V3 position_last_;
  int feats_down_size_ = 0;
  int effct_feat_num_ = 0;
  std::vector<M3> cross_mat_list_;
  std::vector<M3> body_cov_list_;
  std::vector<pointWithVar> pv_list_;
  std::vector<PointToPlane> ptpl_list_;
  std::vector<int> ptpl_index_;       // oracle bookkeeping: point index of each ptpl_list_ entry
  std::vector<int> normal_plane_id_;  // oracle bookkeeping: flat plane id behind pv.normal (sticky), -1 if never
  int omp_threads_ = 1;               // MP_PROC_NUM (CMakeLists.txt:46-58 caps it at 4)
  bool faithful_cost_ = false;        // keep the reference's allocation / mutex cost structure (CPU baseline)
  LioStats stats_;

  ~VoxelMapManager();
  void StateEstimation(StatesGroup &state_propagat);
  void TransformLidar(const M3 &rot, const V3 &t, const std::vector<float> &input_cloud, std::vector<float> &trans_cloud);
  void BuildVoxelMap();
  void UpdateVoxelMap(const std::vector<pointWithVar> &input_points);
  void BuildResidualListOMP(std::vector<pointWithVar> &pv_list, std::vector<PointToPlane> &ptpl_list);
  void build_single_residual(pointWithVar &pv, const VoxelOctoTree *current_octo, const int current_layer, bool &is_sucess,
                             double &prob, PointToPlane &single_ptpl, int &plane_id);
  // oracle-side flattening / un-flattening of the map (independent of the product flattener)
  void Flatten(std::vector<int64_t> &keys, std::vector<int32_t> &first, std::vector<int32_t> &count,
               std::vector<FlatPlane> &planes);
  void FromFlat(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const FlatPlane *planes,
                int n_planes);
};

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header). PINNED to the reference source: oracle/_ref/libfl2_ref_vio.so is
// the reference's own src/vio.cpp compiled against stand-in headers (oracle/ref_vio.cpp) and tests/test_oracle_ref_pin_vio.py
// holds this restatement to it (update, inverse-compositional variant, patch producers). PARITY UNPINNED remains true for the
// one dependency restated on BOTH sides of that comparison:
//
// CPU restatement of the reference's VIO ESIKF measurement update (src/vio.cpp) and of the
// third-party camera arithmetic it calls (vikit, xuankuzcr/rpg_vikit, NO version pin in the
// reference: README.md:80-84). The vikit parts are restated from its published algorithm
// (pinhole + radtan, equidistant fisheye, interpolateMat_8u).
#pragma once
#include "orc_math.hpp"

namespace orc {

// vikit AbstractCamera / PinholeCamera / EquidistantCamera (intrinsics already multiplied by
// `scale`, as vikit's constructors do; config/camera_pinhole.yaml:1-11).
struct Camera {
  int model = 0;  // 0 = Pinhole (radtan d0..d4), 1 = EquidistantCamera (k1..k4)
  int width = 0, height = 0;
  double fx = 0, fy = 0, cx = 0, cy = 0;
  double d[5] = {0, 0, 0, 0, 0};
  V2 world2cam(const V3 &xyz_c) const;
  V3 cam2world(const V2 &px) const;
};

struct SE3 {  // Sophus::SE3 (commit a621ff) restated as (R, t): x' = R x + t
  M3 R;
  V3 t;
  SE3() : R(M3::Identity()), t(V3::Zero()) {}
  SE3(const M3 &R_, const V3 &t_) : R(R_), t(t_) {}
  SE3 inverse() const { return SE3(T(R), -(T(R) * t)); }
  SE3 operator*(const SE3 &o) const { return SE3(R * o.R, R * o.t + t); }
  V3 operator*(const V3 &x) const { return R * x + t; }
};

struct Image {  // cv::Mat CV_8UC1, continuous
  const uint8_t *data = nullptr;
  int cols = 0, rows = 0;
};

float interpolateMat_8u(const Image &mat, float u, float v);  // vk::interpolateMat_8u

struct VioStats {
  int iters_per_level[8];     // iterations executed at each level (index = level)
  int accepted_per_level[8];  // accepted updates at each level
  float error_trace[8][8];    // [level][iteration] mean squared photometric error
  int total_iters;
  double HTH[8][8][49];        // [level][iteration] 7x7 H^T H (accepted iterations only)
  double HTz[8][8][7];
  double solution[8][8][19];
};

class VIOManager {  // include/vio.h (hot-path subset)
 public:
  // configuration (LIVMapper.cpp:50-117 -> vio.cpp:41-160)
  int patch_size = 8, patch_size_total = 64, patch_size_half = 4, patch_pyrimid_level = 4;
  int max_iterations = 5;
  double img_point_cov = 100;
  bool exposure_estimate_en = true;
  int width = 0, height = 0;
  double fx = 0, fy = 0;
  Camera cam;
  M3 Rli, Rci, Rcl, Jdphi_dR, Jdp_dR, Jdp_dt, Rcw;
  V3 Pli, Pci, Pcl, Pcw;
  int omp_threads_ = 1;
  // state pointers into LIVMapper (LIVMapper.cpp:135-136)
  StatesGroup *state = nullptr;
  StatesGroup *state_propagat = nullptr;
  // visual_submap (SubSparseMap, include/vio.h:26-57) flattened
  int total_points = 0;
  std::vector<double> pos;             // voxel_points[i]->pos_, Np*3
  std::vector<float> warp_patch;       // Np * patch_pyrimid_level*64
  std::vector<int> search_levels;      // Np
  std::vector<double> inv_expo_list;   // Np
  std::vector<float> errors;           // Np (out)
  M19 G, H_T_H;
  VioStats stats_;
  // inverse-compositional variant (vio/inverse_composition_en, LIVMapper.cpp:60,140): what it reads of every point's
  // reference feature (include/feature.h: img_, px_, f_, T_f_w_), flattened
  bool inverse_composition_en = false, has_ref_patch_cache = false;
  std::vector<Image> ref_imgs;         // Feature::img_ (one per reference frame)
  std::vector<int> ref_img_index;      // Np
  std::vector<double> ref_px;          // Np*2  ref_patch->px_
  std::vector<double> ref_f;           // Np*3  ref_patch->f_ (unit bearing)
  std::vector<double> ref_R;           // Np*9  ref_patch->T_f_w_.rotation_matrix()
  std::vector<double> ref_pos;         // Np*3  ref_patch->pos() = T_f_w_.inverse().translation()
  std::vector<double> H_sub_inv;       // (Np*64) x 6

  void setImuToLidarExtrinsic(const V3 &transl, const M3 &rot);
  void setLidarToCameraExtrinsic(const M3 &R, const V3 &P);
  void initializeVIO();
  void computeProjectionJacobian(const V3 &p, Mat<2, 3> &J);
  void getImagePatch(const Image &img, const V2 &pc, float *patch_tmp, int level);
  void getWarpMatrixAffineHomography(const Camera &cam, const V2 &px_ref, const V3 &xyz_ref, const V3 &normal_ref, const SE3 &T_cur_ref,
                                     const int level_ref, M2 &A_cur_ref);
  void warpAffine(const M2 &A_cur_ref, const Image &img_ref, const V2 &px_ref, const int level_ref, const int search_level,
                  const int pyramid_level, const int halfpatch_size, float *patch);
  int getBestSearchLevel(const M2 &A_cur_ref, const int max_level);
  void computeJacobianAndUpdateEKF(const Image &img);
  void updateState(const Image &img, int level);
  void precomputeReferencePatches(int level);
  void updateStateInverse(const Image &img, int level);
};

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header). PARITY UNPINNED.
// Restatement of src/voxel_map.cpp (LIO ESIKF update + voxel map construction).
#include "orc_lio.hpp"
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// src/voxel_map.cpp:15-34. DEG2RAD is PCL's macro ((x)*0.017453293), pcl/pcl_macros.h.
void calcBodyCov(V3 &pb, const float range_inc, const float degree_inc, M3 &cov) {
  if (pb[2] == 0) pb[2] = 0.0001;
  float range = std::sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
  float range_var = range_inc * range_inc;
  double dv = std::pow(std::sin((degree_inc)*0.017453293), 2);
  Mat<2, 2> direction_var;
  direction_var(0, 0) = dv, direction_var(0, 1) = 0, direction_var(1, 0) = 0, direction_var(1, 1) = dv;
  V3 direction = pb;
  direction = direction / norm(direction);
  M3 direction_hat = skew(direction);
  V3 base_vector1 = v3(1, 1, -(direction[0] + direction[1]) / direction[2]);
  base_vector1 = base_vector1 / norm(base_vector1);
  V3 base_vector2 = cross(base_vector1, direction);
  base_vector2 = base_vector2 / norm(base_vector2);
  Mat<3, 2> N;
  N(0, 0) = base_vector1[0], N(0, 1) = base_vector2[0];
  N(1, 0) = base_vector1[1], N(1, 1) = base_vector2[1];
  N(2, 0) = base_vector1[2], N(2, 1) = base_vector2[2];
  Mat<3, 2> A = ((double)range * direction_hat) * N;
  cov = (direction * (double)range_var) * T(direction) + (A * direction_var) * T(A);
}

// src/voxel_map.cpp:55-135
void VoxelOctoTree::init_plane(const std::vector<pointWithVar> &points, VoxelPlane *plane) {
  plane->plane_var_ = M6::Zero();
  plane->covariance_ = M3::Zero();
  plane->center_ = V3::Zero();
  plane->normal_ = V3::Zero();
  plane->points_size_ = points.size();
  plane->radius_ = 0;
  for (const auto &pv : points) {
    plane->covariance_ = plane->covariance_ + pv.point_w * T(pv.point_w);
    plane->center_ = plane->center_ + pv.point_w;
  }
  plane->center_ = plane->center_ / (double)plane->points_size_;
  plane->covariance_ = plane->covariance_ / (double)plane->points_size_ - plane->center_ * T(plane->center_);
  double evalsReal[3];
  M3 evecs;
  eig_sym3(plane->covariance_, evalsReal, evecs);  // Eigen::EigenSolver at :70
  int evalsMin = 0, evalsMax = 0;
  for (int i = 1; i < 3; i++) {
    if (evalsReal[i] < evalsReal[evalsMin]) evalsMin = i;
    if (evalsReal[i] > evalsReal[evalsMax]) evalsMax = i;
  }
  int evalsMid = 3 - evalsMin - evalsMax;
  M3 J_Q = M3::Identity() * (1.0 / plane->points_size_);
  if (evalsReal[evalsMin] < planer_threshold_) {
    V3 evecMin = v3(evecs(0, evalsMin), evecs(1, evalsMin), evecs(2, evalsMin));
    for (size_t i = 0; i < points.size(); i++) {
      Mat<6, 3> J;
      M3 F;
      for (int m = 0; m < 3; m++) {
        if (m != evalsMin) {
          V3 em = v3(evecs(0, m), evecs(1, m), evecs(2, m));
          Mat<1, 3> F_m = (T(points[i].point_w - plane->center_) / ((plane->points_size_) * (evalsReal[evalsMin] - evalsReal[m]))) *
                          (em * T(evecMin) + evecMin * T(em));
          for (int c = 0; c < 3; c++) F(m, c) = F_m(0, c);
        } else {
          for (int c = 0; c < 3; c++) F(m, c) = 0;
        }
      }
      set_block(J, 0, 0, evecs * F);
      set_block(J, 3, 0, J_Q);
      plane->plane_var_ = plane->plane_var_ + (J * points[i].var) * T(J);
    }
    plane->normal_ = evecMin;
    plane->y_normal_ = v3(evecs(0, evalsMid), evecs(1, evalsMid), evecs(2, evalsMid));
    plane->x_normal_ = v3(evecs(0, evalsMax), evecs(1, evalsMax), evecs(2, evalsMax));
    plane->min_eigen_value_ = evalsReal[evalsMin];
    plane->mid_eigen_value_ = evalsReal[evalsMid];
    plane->max_eigen_value_ = evalsReal[evalsMax];
    plane->radius_ = std::sqrt(evalsReal[evalsMax]);
    plane->d_ = -(plane->normal_[0] * plane->center_[0] + plane->normal_[1] * plane->center_[1] + plane->normal_[2] * plane->center_[2]);
    plane->is_plane_ = true;
    plane->is_update_ = true;
    if (!plane->is_init_) plane->is_init_ = true;
  } else {
    plane->is_update_ = true;
    plane->is_plane_ = false;
  }
}

// src/voxel_map.cpp:137-161
void VoxelOctoTree::init_octo_tree() {
  if ((int)temp_points_.size() > points_size_threshold_) {
    init_plane(temp_points_, plane_ptr_);
    if (plane_ptr_->is_plane_ == true) {
      octo_state_ = 0;
      if ((int)temp_points_.size() > max_points_num_) {
        update_enable_ = false;
        std::vector<pointWithVar>().swap(temp_points_);
        new_points_ = 0;
      }
    } else {
      octo_state_ = 1;
      cut_octo_tree();
    }
    init_octo_ = true;
    new_points_ = 0;
  }
}

// src/voxel_map.cpp:163-217
void VoxelOctoTree::cut_octo_tree() {
  if (layer_ >= max_layer_) {
    octo_state_ = 0;
    return;
  }
  for (size_t i = 0; i < temp_points_.size(); i++) {
    int xyz[3] = {0, 0, 0};
    if (temp_points_[i].point_w[0] > voxel_center_[0]) xyz[0] = 1;
    if (temp_points_[i].point_w[1] > voxel_center_[1]) xyz[1] = 1;
    if (temp_points_[i].point_w[2] > voxel_center_[2]) xyz[2] = 1;
    int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
    if (leaves_[leafnum] == nullptr) {
      leaves_[leafnum] = new VoxelOctoTree(max_layer_, layer_ + 1, layer_init_num_[layer_ + 1], max_points_num_, planer_threshold_);
      leaves_[leafnum]->layer_init_num_ = layer_init_num_;
      leaves_[leafnum]->voxel_center_[0] = voxel_center_[0] + (2 * xyz[0] - 1) * quater_length_;
      leaves_[leafnum]->voxel_center_[1] = voxel_center_[1] + (2 * xyz[1] - 1) * quater_length_;
      leaves_[leafnum]->voxel_center_[2] = voxel_center_[2] + (2 * xyz[2] - 1) * quater_length_;
      leaves_[leafnum]->quater_length_ = quater_length_ / 2;
    }
    leaves_[leafnum]->temp_points_.push_back(temp_points_[i]);
    leaves_[leafnum]->new_points_++;
  }
  for (unsigned i = 0; i < 8; i++) {
    if (leaves_[i] != nullptr) {
      if ((int)leaves_[i]->temp_points_.size() > leaves_[i]->points_size_threshold_) {
        init_plane(leaves_[i]->temp_points_, leaves_[i]->plane_ptr_);
        if (leaves_[i]->plane_ptr_->is_plane_) {
          leaves_[i]->octo_state_ = 0;
          if ((int)leaves_[i]->temp_points_.size() > leaves_[i]->max_points_num_) {
            leaves_[i]->update_enable_ = false;
            std::vector<pointWithVar>().swap(leaves_[i]->temp_points_);
            new_points_ = 0;
          }
        } else {
          leaves_[i]->octo_state_ = 1;
          leaves_[i]->cut_octo_tree();
        }
        leaves_[i]->init_octo_ = true;
        leaves_[i]->new_points_ = 0;
      }
    }
  }
}

// src/voxel_map.cpp:219-290
void VoxelOctoTree::UpdateOctoTree(const pointWithVar &pv) {
  if (!init_octo_) {
    new_points_++;
    temp_points_.push_back(pv);
    if ((int)temp_points_.size() > points_size_threshold_) init_octo_tree();
  } else {
    if (plane_ptr_->is_plane_) {
      if (update_enable_) {
        new_points_++;
        temp_points_.push_back(pv);
        if (new_points_ > update_size_threshold_) {
          init_plane(temp_points_, plane_ptr_);
          new_points_ = 0;
        }
        if ((int)temp_points_.size() >= max_points_num_) {
          update_enable_ = false;
          std::vector<pointWithVar>().swap(temp_points_);
          new_points_ = 0;
        }
      }
    } else {
      if (layer_ < max_layer_) {
        int xyz[3] = {0, 0, 0};
        if (pv.point_w[0] > voxel_center_[0]) xyz[0] = 1;
        if (pv.point_w[1] > voxel_center_[1]) xyz[1] = 1;
        if (pv.point_w[2] > voxel_center_[2]) xyz[2] = 1;
        int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
        if (leaves_[leafnum] != nullptr) {
          leaves_[leafnum]->UpdateOctoTree(pv);
        } else {
          leaves_[leafnum] = new VoxelOctoTree(max_layer_, layer_ + 1, layer_init_num_[layer_ + 1], max_points_num_, planer_threshold_);
          leaves_[leafnum]->layer_init_num_ = layer_init_num_;
          leaves_[leafnum]->voxel_center_[0] = voxel_center_[0] + (2 * xyz[0] - 1) * quater_length_;
          leaves_[leafnum]->voxel_center_[1] = voxel_center_[1] + (2 * xyz[1] - 1) * quater_length_;
          leaves_[leafnum]->voxel_center_[2] = voxel_center_[2] + (2 * xyz[2] - 1) * quater_length_;
          leaves_[leafnum]->quater_length_ = quater_length_ / 2;
          leaves_[leafnum]->UpdateOctoTree(pv);
        }
      } else {
        if (update_enable_) {
          new_points_++;
          temp_points_.push_back(pv);
          if (new_points_ > update_size_threshold_) {
            init_plane(temp_points_, plane_ptr_);
            new_points_ = 0;
          }
          if ((int)temp_points_.size() > max_points_num_) {
            update_enable_ = false;
            std::vector<pointWithVar>().swap(temp_points_);
            new_points_ = 0;
          }
        }
      }
    }
  }
}

VoxelMapManager::~VoxelMapManager() {
  for (auto &kv : voxel_map_) delete kv.second;
}

// src/voxel_map.cpp:513-530. The PCL cloud stores float xyz, so p_w is narrowed to float here.
void VoxelMapManager::TransformLidar(const M3 &rot, const V3 &t, const std::vector<float> &input_cloud, std::vector<float> &trans_cloud) {
  std::vector<float>().swap(trans_cloud);
  trans_cloud.reserve(input_cloud.size());
  for (size_t i = 0; i < input_cloud.size() / 3; i++) {
    V3 p = v3(input_cloud[3 * i], input_cloud[3 * i + 1], input_cloud[3 * i + 2]);
    p = (rot * (extR_ * p + extT_) + t);
    trans_cloud.push_back((float)p[0]);
    trans_cloud.push_back((float)p[1]);
    trans_cloud.push_back((float)p[2]);
  }
}

// src/voxel_map.cpp:532-591
void VoxelMapManager::BuildVoxelMap() {
  float voxel_size = config_setting_.max_voxel_size_;
  float planer_threshold = config_setting_.planner_threshold_;
  int max_layer = config_setting_.max_layer_;
  int max_points_num = config_setting_.max_points_num_;
  std::vector<int> layer_init_num = config_setting_.layer_init_num_;
  std::vector<pointWithVar> input_points;
  for (size_t i = 0; i < feats_down_world_.size() / 3; i++) {
    pointWithVar pv;
    pv.point_w = v3(feats_down_world_[3 * i], feats_down_world_[3 * i + 1], feats_down_world_[3 * i + 2]);
    V3 point_this = v3(feats_down_body_[3 * i], feats_down_body_[3 * i + 1], feats_down_body_[3 * i + 2]);
    M3 var;
    calcBodyCov(point_this, config_setting_.dept_err_, config_setting_.beam_err_, var);
    M3 point_crossmat = skew(point_this);
    M3 RE = state_.rot_end * extR_;
    var = (RE * var) * T(RE) + ((-point_crossmat) * block<3, 3>(state_.cov, 0, 0)) * T(-point_crossmat) + block<3, 3>(state_.cov, 3, 3);
    pv.var = var;
    input_points.push_back(pv);
  }
  unsigned plsize = input_points.size();
  for (unsigned i = 0; i < plsize; i++) {
    const pointWithVar p_v = input_points[i];
    float loc_xyz[3];
    for (int j = 0; j < 3; j++) {
      loc_xyz[j] = p_v.point_w[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
    }
    VOXEL_LOCATION position((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto iter = voxel_map_.find(position);
    if (iter != voxel_map_.end()) {
      voxel_map_[position]->temp_points_.push_back(p_v);
      voxel_map_[position]->new_points_++;
    } else {
      VoxelOctoTree *octo_tree = new VoxelOctoTree(max_layer, 0, layer_init_num[0], max_points_num, planer_threshold);
      voxel_map_[position] = octo_tree;
      voxel_map_[position]->quater_length_ = voxel_size / 4;
      voxel_map_[position]->voxel_center_[0] = (0.5 + position.x) * voxel_size;
      voxel_map_[position]->voxel_center_[1] = (0.5 + position.y) * voxel_size;
      voxel_map_[position]->voxel_center_[2] = (0.5 + position.z) * voxel_size;
      voxel_map_[position]->temp_points_.push_back(p_v);
      voxel_map_[position]->new_points_++;
      voxel_map_[position]->layer_init_num_ = layer_init_num;
    }
  }
  for (auto iter = voxel_map_.begin(); iter != voxel_map_.end(); ++iter) iter->second->init_octo_tree();
}

// src/voxel_map.cpp:609-641
void VoxelMapManager::UpdateVoxelMap(const std::vector<pointWithVar> &input_points) {
  float voxel_size = config_setting_.max_voxel_size_;
  float planer_threshold = config_setting_.planner_threshold_;
  int max_layer = config_setting_.max_layer_;
  int max_points_num = config_setting_.max_points_num_;
  std::vector<int> layer_init_num = config_setting_.layer_init_num_;
  unsigned plsize = input_points.size();
  for (unsigned i = 0; i < plsize; i++) {
    const pointWithVar p_v = input_points[i];
    float loc_xyz[3];
    for (int j = 0; j < 3; j++) {
      loc_xyz[j] = p_v.point_w[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
    }
    VOXEL_LOCATION position((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto iter = voxel_map_.find(position);
    if (iter != voxel_map_.end()) {
      voxel_map_[position]->UpdateOctoTree(p_v);
    } else {
      VoxelOctoTree *octo_tree = new VoxelOctoTree(max_layer, 0, layer_init_num[0], max_points_num, planer_threshold);
      voxel_map_[position] = octo_tree;
      voxel_map_[position]->layer_init_num_ = layer_init_num;
      voxel_map_[position]->quater_length_ = voxel_size / 4;
      voxel_map_[position]->voxel_center_[0] = (0.5 + position.x) * voxel_size;
      voxel_map_[position]->voxel_center_[1] = (0.5 + position.y) * voxel_size;
      voxel_map_[position]->voxel_center_[2] = (0.5 + position.z) * voxel_size;
      voxel_map_[position]->UpdateOctoTree(p_v);
    }
  }
}

// src/voxel_map.cpp:713-786
void VoxelMapManager::build_single_residual(pointWithVar &pv, const VoxelOctoTree *current_octo, const int current_layer, bool &is_sucess,
                                            double &prob, PointToPlane &single_ptpl, int &plane_id) {
  int max_layer = config_setting_.max_layer_;
  double sigma_num = config_setting_.sigma_num_;
  double radius_k = 3;
  V3 p_w = pv.point_w;
  if (current_octo->plane_ptr_->is_plane_) {
    VoxelPlane &plane = *current_octo->plane_ptr_;
    float dis_to_plane = std::fabs(plane.normal_[0] * p_w[0] + plane.normal_[1] * p_w[1] + plane.normal_[2] * p_w[2] + plane.d_);
    float dis_to_center = (plane.center_[0] - p_w[0]) * (plane.center_[0] - p_w[0]) + (plane.center_[1] - p_w[1]) * (plane.center_[1] - p_w[1]) +
                          (plane.center_[2] - p_w[2]) * (plane.center_[2] - p_w[2]);
    float range_dis = std::sqrt(dis_to_center - dis_to_plane * dis_to_plane);
    if (range_dis <= radius_k * plane.radius_) {
      Mat<1, 6> J_nq;
      for (int k = 0; k < 3; k++) {
        J_nq(0, k) = p_w[k] - plane.center_[k];
        J_nq(0, 3 + k) = -plane.normal_[k];
      }
      double sigma_l = ((J_nq * plane.plane_var_) * T(J_nq))[0];
      sigma_l += ((T(plane.normal_) * pv.var) * plane.normal_)[0];
      if (dis_to_plane < sigma_num * std::sqrt(sigma_l)) {
        is_sucess = true;
        double this_prob = 1.0 / (std::sqrt(sigma_l)) * std::exp(-0.5 * dis_to_plane * dis_to_plane / sigma_l);
        if (this_prob > prob) {
          prob = this_prob;
          pv.normal = plane.normal_;
          single_ptpl.body_cov_ = pv.body_var;
          single_ptpl.point_b_ = pv.point_b;
          single_ptpl.point_w_ = pv.point_w;
          single_ptpl.plane_var_ = plane.plane_var_;
          single_ptpl.normal_ = plane.normal_;
          single_ptpl.center_ = plane.center_;
          single_ptpl.d_ = plane.d_;
          single_ptpl.layer_ = current_layer;
          single_ptpl.dis_to_plane_ = plane.normal_[0] * p_w[0] + plane.normal_[1] * p_w[1] + plane.normal_[2] * p_w[2] + plane.d_;
          plane_id = plane.flat_id_;
        }
        return;
      } else {
        return;
      }
    } else {
      return;
    }
  } else {
    if (current_layer < max_layer) {
      for (size_t leafnum = 0; leafnum < 8; leafnum++) {
        if (current_octo->leaves_[leafnum] != nullptr) {
          VoxelOctoTree *leaf_octo = current_octo->leaves_[leafnum];
          build_single_residual(pv, leaf_octo, current_layer + 1, is_sucess, prob, single_ptpl, plane_id);
        }
      }
      return;
    } else {
      return;
    }
  }
}

// src/voxel_map.cpp:643-711
void VoxelMapManager::BuildResidualListOMP(std::vector<pointWithVar> &pv_list, std::vector<PointToPlane> &ptpl_list) {
  double voxel_size = config_setting_.max_voxel_size_;
  std::mutex mylock;
  ptpl_list.clear();
  ptpl_index_.clear();
  std::vector<PointToPlane> all_ptpl_list(pv_list.size());
  std::vector<bool> useful_ptpl(pv_list.size());
  std::vector<size_t> index(pv_list.size());
  for (size_t i = 0; i < index.size(); ++i) {
    index[i] = i;
    useful_ptpl[i] = false;
  }
#ifdef _OPENMP
  omp_set_num_threads(omp_threads_);
#pragma omp parallel for
#endif
  for (int i = 0; i < (int)index.size(); i++) {
    pointWithVar &pv = pv_list[i];
    float loc_xyz[3];
    for (int j = 0; j < 3; j++) {
      loc_xyz[j] = pv.point_w[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
    }
    VOXEL_LOCATION position((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto iter = voxel_map_.find(position);
    if (iter != voxel_map_.end()) {
      VoxelOctoTree *current_octo = iter->second;
      PointToPlane single_ptpl;
      single_ptpl.plane_id_ = -1;
      int plane_id = -1;
      bool is_sucess = false;
      double prob = 0;
      build_single_residual(pv, current_octo, 0, is_sucess, prob, single_ptpl, plane_id);
      if (!is_sucess) {
        VOXEL_LOCATION near_position = position;
        if (loc_xyz[0] > (current_octo->voxel_center_[0] + current_octo->quater_length_)) near_position.x = near_position.x + 1;
        else if (loc_xyz[0] < (current_octo->voxel_center_[0] - current_octo->quater_length_)) near_position.x = near_position.x - 1;
        if (loc_xyz[1] > (current_octo->voxel_center_[1] + current_octo->quater_length_)) near_position.y = near_position.y + 1;
        else if (loc_xyz[1] < (current_octo->voxel_center_[1] - current_octo->quater_length_)) near_position.y = near_position.y - 1;
        if (loc_xyz[2] > (current_octo->voxel_center_[2] + current_octo->quater_length_)) near_position.z = near_position.z + 1;
        else if (loc_xyz[2] < (current_octo->voxel_center_[2] - current_octo->quater_length_)) near_position.z = near_position.z - 1;
        auto iter_near = voxel_map_.find(near_position);
        if (iter_near != voxel_map_.end()) build_single_residual(pv, iter_near->second, 0, is_sucess, prob, single_ptpl, plane_id);
      }
      // Documented deviation: the reference would push an uninitialised PointToPlane when a
      // plane passes the gate but this_prob > prob never holds (NaN/0 probability). We
      // require a chosen plane (plane_id >= 0).
      if (is_sucess && plane_id >= 0) {
        single_ptpl.plane_id_ = plane_id;
        normal_plane_id_[i] = plane_id;
        mylock.lock();
        useful_ptpl[i] = true;
        all_ptpl_list[i] = single_ptpl;
        mylock.unlock();
      } else {
        mylock.lock();
        useful_ptpl[i] = false;
        mylock.unlock();
      }
    }
  }
  for (size_t i = 0; i < useful_ptpl.size(); i++) {
    if (useful_ptpl[i]) {
      ptpl_list.push_back(all_ptpl_list[i]);
      ptpl_index_.push_back((int)i);
    }
  }
}

// src/voxel_map.cpp:338-511
void VoxelMapManager::StateEstimation(StatesGroup &state_propagat) {
  cross_mat_list_.clear();
  cross_mat_list_.reserve(feats_down_size_);
  body_cov_list_.clear();
  body_cov_list_.reserve(feats_down_size_);
  memset(&stats_, 0, sizeof(stats_));

  for (int i = 0; i < feats_down_size_; i++) {
    V3 point_this = v3(feats_down_body_[3 * i], feats_down_body_[3 * i + 1], feats_down_body_[3 * i + 2]);
    if (point_this[2] == 0) point_this[2] = 0.001;
    M3 var;
    calcBodyCov(point_this, config_setting_.dept_err_, config_setting_.beam_err_, var);
    body_cov_list_.push_back(var);
    point_this = extR_ * point_this + extT_;
    cross_mat_list_.push_back(skew(point_this));
  }

  std::vector<pointWithVar>().swap(pv_list_);
  pv_list_.resize(feats_down_size_);
  normal_plane_id_.assign(feats_down_size_, -1);

  int rematch_num = 0;
  M19 G = M19::Zero(), H_T_H = M19::Zero(), I_STATE = M19::Identity();

  bool flg_EKF_converged, EKF_stop_flg = 0;
  for (int iterCount = 0; iterCount < config_setting_.max_iterations_; iterCount++) {
    double total_residual = 0.0;
    std::vector<float> world_lidar;
    TransformLidar(state_.rot_end, state_.pos_end, feats_down_body_, world_lidar);
    M3 rot_var = block<3, 3>(state_.cov, 0, 0);
    M3 t_var = block<3, 3>(state_.cov, 3, 3);
    for (int i = 0; i < feats_down_size_; i++) {
      pointWithVar &pv = pv_list_[i];
      pv.point_b = v3(feats_down_body_[3 * i], feats_down_body_[3 * i + 1], feats_down_body_[3 * i + 2]);
      pv.point_w = v3(world_lidar[3 * i], world_lidar[3 * i + 1], world_lidar[3 * i + 2]);
      M3 cov = body_cov_list_[i];
      M3 point_crossmat = cross_mat_list_[i];
      cov = (state_.rot_end * cov) * T(state_.rot_end) + ((-point_crossmat) * rot_var) * (-T(point_crossmat)) + t_var;
      pv.var = cov;
      pv.body_var = body_cov_list_[i];
    }
    ptpl_list_.clear();

    BuildResidualListOMP(pv_list_, ptpl_list_);

    for (size_t i = 0; i < ptpl_list_.size(); i++) total_residual += std::fabs(ptpl_list_[i].dis_to_plane_);
    effct_feat_num_ = ptpl_list_.size();
    // (per-iteration cout at :404-405 suppressed)

    // Hsub / Hsub_T_R_inv / R_inv / meas_vec : dynamic Eigen matrices at :409-412
    std::vector<double> Hsub((size_t)effct_feat_num_ * 6), Hsub_T_R_inv((size_t)effct_feat_num_ * 6), R_inv(effct_feat_num_),
        meas_vec(effct_feat_num_, 0.0);
    for (int i = 0; i < effct_feat_num_; i++) {
      auto &ptpl = ptpl_list_[i];
      V3 point_this = ptpl.point_b_;
      point_this = extR_ * point_this + extT_;
      M3 point_crossmat = skew(point_this);
      V3 point_world = state_propagat.rot_end * point_this + state_propagat.pos_end;
      Mat<1, 6> J_nq;
      for (int k = 0; k < 3; k++) {
        J_nq(0, k) = point_world[k] - ptpl.center_[k];
        J_nq(0, 3 + k) = -ptpl.normal_[k];
      }
      M3 RE = state_propagat.rot_end * extR_;
      M3 var = (RE * ptpl.body_cov_) * T(RE);
      double sigma_l = ((J_nq * ptpl.plane_var_) * T(J_nq))[0];
      R_inv[i] = 1.0 / (0.001 + sigma_l + ((T(ptpl.normal_) * var) * ptpl.normal_)[0]);
      V3 A = (point_crossmat * T(state_.rot_end)) * ptpl.normal_;
      for (int k = 0; k < 3; k++) {
        Hsub[i * 6 + k] = A[k];
        Hsub[i * 6 + 3 + k] = ptpl.normal_[k];
        Hsub_T_R_inv[i * 6 + k] = A[k] * R_inv[i];
        Hsub_T_R_inv[i * 6 + 3 + k] = ptpl.normal_[k] * R_inv[i];
      }
      meas_vec[i] = -ptpl.dis_to_plane_;
    }
    EKF_stop_flg = false;
    flg_EKF_converged = false;
    // HTz = Hsub_T_R_inv * meas_vec ; H_T_H(6x6) = Hsub_T_R_inv * Hsub   (:464-466)
    Mat<6, 1> HTz = Mat<6, 1>::Zero();
    M6 HTH6 = M6::Zero();
    for (int i = 0; i < effct_feat_num_; i++) {
      for (int r = 0; r < 6; r++) {
        HTz[r] += Hsub_T_R_inv[i * 6 + r] * meas_vec[i];
        for (int c = 0; c < 6; c++) HTH6(r, c) += Hsub_T_R_inv[i * 6 + r] * Hsub[i * 6 + c];
      }
    }
    set_block(H_T_H, 0, 0, HTH6);
    M19 K_1 = inverse_pplu(H_T_H + inverse_pplu(state_.cov));  // :468
    Mat<19, 6> K16 = block<19, 6>(K_1, 0, 0);
    Mat<19, 6> G6 = K16 * HTH6;  // :469
    set_block(G, 0, 0, G6);
    V19 vec = state_propagat.boxminus(state_);  // :470
    V19 solution = K16 * HTz + vec - G6 * block<6, 1>(vec, 0, 0);  // :471-472
    state_.boxplus(solution);  // :474
    V3 rot_add = block<3, 1>(solution, 0, 0);
    V3 t_add = block<3, 1>(solution, 3, 0);
    if ((norm(rot_add) * 57.3 < 0.01) && (norm(t_add) * 100 < 0.015)) flg_EKF_converged = true;  // :477

    {
      int it = iterCount < 8 ? iterCount : 7;
      stats_.iters = iterCount + 1;
      stats_.effct_feat_num[it] = effct_feat_num_;
      stats_.total_residual[it] = total_residual;
      memcpy(stats_.HTH[it], HTH6.a, sizeof(double) * 36);
      memcpy(stats_.HTz[it], HTz.a, sizeof(double) * 6);
      memcpy(stats_.solution[it], solution.a, sizeof(double) * 19);
      stats_.converged[it] = flg_EKF_converged;
    }

    // :482
    if (flg_EKF_converged || ((rematch_num == 0) && (iterCount == (config_setting_.max_iterations_ - 2)))) rematch_num++;
    // :485-498
    if (!EKF_stop_flg && (rematch_num >= 2 || (iterCount == config_setting_.max_iterations_ - 1))) {
      state_.cov = (I_STATE - G) * state_.cov;
      position_last_ = state_.pos_end;
      EKF_stop_flg = true;
    }
    if (EKF_stop_flg) break;
  }
}

// Oracle-side flattening: DFS of each root in leaf order 0..7 (the visiting order of
// build_single_residual, :771-784); a plane node terminates its branch (:721).
static void flatten_node(const VoxelOctoTree *node, int layer, int max_layer, int path, std::vector<FlatPlane> &planes) {
  if (node->plane_ptr_->is_plane_) {
    const VoxelPlane &p = *node->plane_ptr_;
    FlatPlane f;
    memset(&f, 0, sizeof(f));
    for (int k = 0; k < 3; k++) f.center[k] = p.center_[k], f.normal[k] = p.normal_[k];
    int t = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) f.plane_var[t++] = p.plane_var_(i, j);
    f.d = p.d_;
    f.radius = p.radius_;
    f.layer = layer;
    f.path = path;
    node->plane_ptr_->flat_id_ = (int)planes.size();
    planes.push_back(f);
    return;
  }
  if (layer < max_layer)
    for (int l = 0; l < 8; l++)
      if (node->leaves_[l] != nullptr) flatten_node(node->leaves_[l], layer + 1, max_layer, path | (l << (3 * layer)), planes);
}

void VoxelMapManager::Flatten(std::vector<int64_t> &keys, std::vector<int32_t> &first, std::vector<int32_t> &count,
                              std::vector<FlatPlane> &planes) {
  keys.clear(), first.clear(), count.clear(), planes.clear();
  for (auto &kv : voxel_map_) {
    keys.push_back(kv.first.x), keys.push_back(kv.first.y), keys.push_back(kv.first.z);
    first.push_back((int)planes.size());
    flatten_node(kv.second, 0, config_setting_.max_layer_, 0, planes);
    count.push_back((int)planes.size() - first.back());
  }
}

// Rebuild a pointer octree (what the reference walks) from the flat arrays. Only plane
// nodes and the interior nodes leading to them are recreated; nodes that hold no plane
// contribute nothing to build_single_residual.
void VoxelMapManager::FromFlat(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const FlatPlane *planes,
                               int n_planes) {
  (void)n_planes;
  for (auto &kv : voxel_map_) delete kv.second;
  voxel_map_.clear();
  float voxel_size = config_setting_.max_voxel_size_;
  int max_layer = config_setting_.max_layer_;
  for (int r = 0; r < n_roots; r++) {
    VOXEL_LOCATION position(keys[3 * r], keys[3 * r + 1], keys[3 * r + 2]);
    VoxelOctoTree *root = new VoxelOctoTree(max_layer, 0, 5, config_setting_.max_points_num_, (float)config_setting_.planner_threshold_);
    root->quater_length_ = voxel_size / 4;                       // :578
    root->voxel_center_[0] = (0.5 + position.x) * voxel_size;    // :579-581
    root->voxel_center_[1] = (0.5 + position.y) * voxel_size;
    root->voxel_center_[2] = (0.5 + position.z) * voxel_size;
    root->init_octo_ = true;
    voxel_map_[position] = root;
    for (int c = 0; c < count[r]; c++) {
      const FlatPlane &f = planes[first[r] + c];
      VoxelOctoTree *node = root;
      for (int l = 0; l < f.layer; l++) {
        int leaf = (f.path >> (3 * l)) & 7;
        if (node->leaves_[leaf] == nullptr) {
          node->leaves_[leaf] = new VoxelOctoTree(max_layer, l + 1, 5, config_setting_.max_points_num_, (float)config_setting_.planner_threshold_);
          int xyz[3] = {(leaf >> 2) & 1, (leaf >> 1) & 1, leaf & 1};
          for (int k = 0; k < 3; k++) node->leaves_[leaf]->voxel_center_[k] = node->voxel_center_[k] + (2 * xyz[k] - 1) * node->quater_length_;
          node->leaves_[leaf]->quater_length_ = node->quater_length_ / 2;
          node->leaves_[leaf]->init_octo_ = true;
        }
        node = node->leaves_[leaf];
      }
      VoxelPlane &p = *node->plane_ptr_;
      for (int k = 0; k < 3; k++) p.center_[k] = f.center[k], p.normal_[k] = f.normal[k];
      int t = 0;
      for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) {
          p.plane_var_(i, j) = f.plane_var[t];
          p.plane_var_(j, i) = f.plane_var[t];
          t++;
        }
      p.d_ = f.d;
      p.radius_ = f.radius;
      p.is_plane_ = true;
      p.is_init_ = true;
      p.flat_id_ = first[r] + c;
    }
  }
}

}  // namespace orc

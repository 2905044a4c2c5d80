#pragma once
// Stand-in for Sophus (pinned by the reference at commit a621ff, README.md:67-71): the SE3 interface src/vio.cpp and the
// Frame / Feature headers use. That Sophus stores the rotation as a unit quaternion; the stand-in keeps the matrix, so
// rotation_matrix() returns exactly what the constructor received (differences to the real class: <= 1e-16 relative, and only
// in the warp producers and the published frame pose — not in the update's residuals or Jacobians).
#include <Eigen/Dense>
namespace Sophus {
class SO3 {};
class SE3 {
  Eigen::Matrix3d R_;
  Eigen::Vector3d t_;
 public:
  SE3() : R_(Eigen::Matrix3d::Identity()), t_(Eigen::Vector3d::Zero()) {}
  SE3(const Eigen::Matrix3d &R, const Eigen::Vector3d &t) : R_(R), t_(t) {}
  Eigen::Matrix3d rotation_matrix() const { return R_; }
  const Eigen::Vector3d &translation() const { return t_; }
  Eigen::Vector3d &translation() { return t_; }
  SE3 inverse() const {
    const Eigen::Matrix3d Rt = R_.transpose();
    return SE3(Rt, (Rt * t_) * -1.0);
  }
  SE3 operator*(const SE3 &o) const { return SE3(R_ * o.R_, t_ + R_ * o.t_); }
  Eigen::Vector3d operator*(const Eigen::Vector3d &p) const { return R_ * p + t_; }
};
}  // namespace Sophus

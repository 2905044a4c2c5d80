#pragma once
// Stand-in for vk::EquidistantCamera (fisheye, k1..k4; config/camera_fisheye_HILTI22.yaml): restated from the published model
// (theta_d = theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8), fixed-point inversion), same as oracle/orc_vio.cpp.
// Not included by the reference's sources (they only see vk::AbstractCamera*); used by oracle/ref_vio.cpp to pin the update with
// this camera behind the abstract interface.
#include <cmath>
#include "abstract_camera.h"
namespace vk {
class EquidistantCamera : public AbstractCamera {
  double fx_, fy_, cx_, cy_, scale_, k_[4];
 public:
  EquidistantCamera(double width, double height, double scale, double fx, double fy, double cx, double cy, double k1, double k2, double k3, double k4)
      : AbstractCamera((int)(width * scale), (int)(height * scale), scale), fx_(fx * scale), fy_(fy * scale), cx_(cx * scale), cy_(cy * scale), scale_(scale) {
    k_[0] = k1, k_[1] = k2, k_[2] = k3, k_[3] = k4;
  }
  Eigen::Vector3d cam2world(const double &u, const double &v) const override {
    const double x0 = (u - cx_) / fx_, y0 = (v - cy_) / fy_;
    double x = x0, y = y0;
    const double theta_d = std::sqrt(x0 * x0 + y0 * y0);
    if (theta_d > 1e-8) {
      double theta = theta_d;
      for (int it = 0; it < 10; it++) {
        const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        theta = theta_d / (1 + k_[0] * t2 + k_[1] * t4 + k_[2] * t6 + k_[3] * t8);
      }
      const double scaling = std::tan(theta) / theta_d;
      x = x0 * scaling, y = y0 * scaling;
    }
    Eigen::Vector3d xyz;
    xyz[0] = x, xyz[1] = y, xyz[2] = 1.0;
    return xyz.normalized();
  }
  Eigen::Vector3d cam2world(const Eigen::Vector2d &px) const override { return cam2world(px[0], px[1]); }
  Eigen::Vector2d world2cam(const Eigen::Vector3d &xyz_c) const override {
    Eigen::Vector2d uv;
    uv[0] = xyz_c[0] / xyz_c[2], uv[1] = xyz_c[1] / xyz_c[2];
    return world2cam(uv);
  }
  Eigen::Vector2d world2cam(const Eigen::Vector2d &uv) const override {
    const double x = uv[0], y = uv[1];
    const double r = std::sqrt(x * x + y * y);
    const double theta = std::atan(r);
    const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    const double theta_d = theta * (1 + k_[0] * t2 + k_[1] * t4 + k_[2] * t6 + k_[3] * t8);
    const double scaling = (r > 1e-8) ? theta_d / r : 1.0;
    Eigen::Vector2d px;
    px[0] = fx_ * x * scaling + cx_, px[1] = fy_ * y * scaling + cy_;
    return px;
  }
  double errorMultiplier2() const override { return std::fabs(fx_); }
  double errorMultiplier() const override { return std::fabs(4.0 * fx_ * fy_); }
  double fx() const override { return fx_; }
  double fy() const override { return fy_; }
  double cx() const override { return cx_; }
  double cy() const override { return cy_; }
  double scale() const override { return scale_; }
};
}  // namespace vk

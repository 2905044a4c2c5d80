#pragma once
// Stand-in for vk::PinholeCamera: pinhole + radial-tangential distortion (d0 d1 d2 d3 d4 = k1 k2 p1 p2 k3), intrinsics and size
// pre-multiplied by `scale` (config/camera_pinhole.yaml). Restated from the published model, same as oracle/orc_vio.cpp.
#include <cmath>
#include <opencv2/opencv.hpp>
#include "abstract_camera.h"
namespace vk {
class PinholeCamera : public AbstractCamera {
  double fx_, fy_, cx_, cy_, scale_;
  bool distortion_;
  double d_[5];
 public:
  PinholeCamera(double width, double height, double scale, double fx, double fy, double cx, double cy, double d0 = 0.0, double d1 = 0.0, double d2 = 0.0, double d3 = 0.0,
                double d4 = 0.0)
      : AbstractCamera((int)(width * scale), (int)(height * scale), scale), fx_(fx * scale), fy_(fy * scale), cx_(cx * scale), cy_(cy * scale), scale_(scale),
        distortion_(std::fabs(d0) > 0.0000001) {
    d_[0] = d0, d_[1] = d1, d_[2] = d2, d_[3] = d3, d_[4] = d4;
  }
  Eigen::Vector3d cam2world(const double &u, const double &v) const override {
    double x0 = (u - cx_) / fx_, y0 = (v - cy_) / fy_, x = x0, y = y0;
    if (distortion_) {  // cv::undistortPoints' fixed-point iteration (5 rounds)
      for (int it = 0; it < 5; it++) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((d_[4] * r2 + d_[1]) * r2 + d_[0]) * r2);
        const double dx = 2 * d_[2] * x * y + d_[3] * (r2 + 2 * x * x);
        const double dy = d_[2] * (r2 + 2 * y * y) + 2 * d_[3] * x * y;
        x = (x0 - dx) * icdist, y = (y0 - dy) * icdist;
      }
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
    Eigen::Vector2d px;
    if (!distortion_) {
      px[0] = fx_ * uv[0] + cx_, px[1] = fy_ * uv[1] + cy_;
    } else {
      const double x = uv[0], y = uv[1];
      const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
      const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
      const double cdist = 1 + d_[0] * r2 + d_[1] * r4 + d_[4] * r6;
      const double xd = x * cdist + d_[2] * a1 + d_[3] * a2, yd = y * cdist + d_[2] * a3 + d_[3] * a1;
      px[0] = xd * fx_ + cx_, px[1] = yd * fy_ + cy_;
    }
    return px;
  }
  double errorMultiplier2() const override { return std::fabs(fx_); }
  double errorMultiplier() const override { return std::fabs(4.0 * fx_ * fy_); }
  double fx() const override { return fx_; }
  double fy() const override { return fy_; }
  double cx() const override { return cx_; }
  double cy() const override { return cy_; }
  double scale() const override { return scale_; }
  void undistortImage(const cv::Mat &raw, cv::Mat &rectified) { rectified = raw.clone(); }  // visualisation only (vio.cpp:1771)
};
}  // namespace vk

#pragma once
// Stand-in for vikit_common's vk::AbstractCamera (xuankuzcr/rpg_vikit, un-vendored dependency of the reference, no version pin):
// the interface src/vio.cpp calls. The arithmetic of the concrete models (pinhole_camera.h) is a restatement of the published
// algorithm — the one part of the VIO pin that is NOT the reference's own code.
#include <Eigen/Dense>
namespace vk {
class AbstractCamera {
 protected:
  int width_ = 0, height_ = 0;
 public:
  AbstractCamera() {}
  AbstractCamera(int width, int height, double) : width_(width), height_(height) {}
  virtual ~AbstractCamera() {}
  virtual Eigen::Vector3d cam2world(const double &x, const double &y) const = 0;
  virtual Eigen::Vector3d cam2world(const Eigen::Vector2d &px) const = 0;
  virtual Eigen::Vector2d world2cam(const Eigen::Vector3d &xyz_c) const = 0;
  virtual Eigen::Vector2d world2cam(const Eigen::Vector2d &uv) const = 0;
  virtual double errorMultiplier2() const = 0;
  virtual double errorMultiplier() const = 0;
  virtual double fx() const = 0;
  virtual double fy() const = 0;
  virtual double cx() const = 0;
  virtual double cy() const = 0;
  virtual double scale() const = 0;
  inline int width() const { return width_; }
  inline int height() const { return height_; }
  inline bool isInFrame(const Eigen::Vector2i &obs, int boundary = 0) const {
    return obs[0] >= boundary && obs[0] < width() - boundary && obs[1] >= boundary && obs[1] < height() - boundary;
  }
  inline bool isInFrame(const Eigen::Vector2i &obs, int boundary, int level) const {
    return obs[0] >= boundary && obs[0] < width() / (1 << level) - boundary && obs[1] >= boundary && obs[1] < height() / (1 << level) - boundary;
  }
};
}  // namespace vk

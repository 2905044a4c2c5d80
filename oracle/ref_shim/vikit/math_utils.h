#pragma once
#include <Eigen/Dense>
namespace vk {
inline Eigen::Vector2d project2d(const Eigen::Vector3d &v) { Eigen::Vector2d r; r[0] = v[0] / v[2], r[1] = v[1] / v[2]; return r; }
}

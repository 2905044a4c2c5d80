#pragma once
#include <memory>
#include <vector>
#include "point_types.h"
namespace pcl {
template <class P> class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  std::vector<P> points;
  size_t size() const { return points.size(); }
  void reserve(size_t n) { points.reserve(n); }
  void clear() { points.clear(); }
  void resize(size_t n) { points.resize(n); }
  void push_back(const P &p) { points.push_back(p); }
  void swap(PointCloud &o) { points.swap(o.points); }
  P &operator[](size_t i) { return points[i]; }
  const P &operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl

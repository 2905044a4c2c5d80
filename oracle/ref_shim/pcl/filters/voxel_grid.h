#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class P> class VoxelGrid {  // named by include/vio.h only (member declaration); never exercised on the pinned path
 public:
  void setLeafSize(float, float, float) {}
  void setInputCloud(const typename PointCloud<P>::Ptr &) {}
  void filter(PointCloud<P> &) {}
};
}  // namespace pcl

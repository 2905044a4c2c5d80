#pragma once
#include <string>
#include <vector>
#include "../geometry_msgs/Quaternion.h"
#include "../ros/ros.h"
namespace std_msgs { struct Header { std::string frame_id; ros::Time stamp; }; struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, ADD = 0, DELETE = 2 };
  std_msgs::Header header;
  std::string ns;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
};
}  // namespace visualization_msgs

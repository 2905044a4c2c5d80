#pragma once
#include <memory>
namespace sensor_msgs { struct Imu { typedef std::shared_ptr<const Imu> ConstPtr; typedef std::shared_ptr<Imu> Ptr; }; }

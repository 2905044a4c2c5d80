#pragma once
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace sensor_msgs { struct Imu { typedef std::shared_ptr<const Imu> ConstPtr; typedef std::shared_ptr<Imu> Ptr; }; }

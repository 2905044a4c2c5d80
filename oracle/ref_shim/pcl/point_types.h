// Stand-in for the PCL point types the reference's LIO path touches (oracle/_ref build only).
#pragma once
#include <cstdint>
#ifndef DEG2RAD
#define DEG2RAD(x) ((x)*0.017453293)  // pcl/pcl_macros.h
#endif
#ifndef RAD2DEG
#define RAD2DEG(x) ((x)*57.29578)
#endif
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointXYZINormal { float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointXYZRGB { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0; };
struct PointXYZRGBA { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0, a = 0; };
}  // namespace pcl

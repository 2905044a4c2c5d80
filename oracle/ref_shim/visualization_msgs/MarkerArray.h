#pragma once
#include "Marker.h"
namespace visualization_msgs { struct MarkerArray { std::vector<Marker> markers; }; }

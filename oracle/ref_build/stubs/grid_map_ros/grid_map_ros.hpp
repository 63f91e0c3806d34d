// Stand-in for <grid_map_ros/grid_map_ros.hpp> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

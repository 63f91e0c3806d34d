// Stand-in for <tf2_geometry_msgs/tf2_geometry_msgs.hpp> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

// Stand-in for <geometry_msgs/msg/point.hpp> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

// Stand-in for <tf2/LinearMath/Quaternion.h> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

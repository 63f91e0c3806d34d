// Stand-in for <pcl/kdtree/kdtree_flann.h> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

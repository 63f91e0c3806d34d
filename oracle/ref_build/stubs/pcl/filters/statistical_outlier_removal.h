// Stand-in for <pcl/filters/statistical_outlier_removal.h> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

// Stand-in for <pcl/filters/extract_indices.h> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

// Stand-in for <opencv2/opencv.hpp> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

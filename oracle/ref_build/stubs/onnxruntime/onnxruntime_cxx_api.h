// Stand-in for <onnxruntime/onnxruntime_cxx_api.h> (absent in the build container): everything lives in gv_standins.hpp.
#pragma once
#include "gv_standins.hpp"

// Forwarding header: the legacy (v1) include root used by the reference's examples and pybind
// (e.g. examples/00_TensorRT/infer.cc:33-37 include "tensorrt/laboratory/utils.h").
#pragma once
#include "trtlab/tensorrt/tensorrt.h"

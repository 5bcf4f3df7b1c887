// Forwarding header: keeps the reference's include path (trtlab/tensorrt/include/trtlab/tensorrt/infer_runner.h).
#pragma once
#include "trtlab/tensorrt/tensorrt.h"

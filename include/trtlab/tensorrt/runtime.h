// Forwarding header: keeps the reference's include path (trtlab/tensorrt/include/trtlab/tensorrt/runtime.h).
#pragma once
#include "trtlab/tensorrt/tensorrt.h"

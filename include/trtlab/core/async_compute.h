// Forwarding header for the trtlab/core subset the hot path consumes (see hotpath_core.h).
#pragma once
#include "trtlab/core/hotpath_core.h"

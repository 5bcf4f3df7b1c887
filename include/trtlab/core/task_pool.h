// Forwarding header: DeferredShortTaskPool is defined with the batcher/dispatcher.
#pragma once
#include "trtlab/core/batcher.h"

// Forwarding header: the Dispatcher lives next to the batcher it drives.
#pragma once
#include "trtlab/core/batcher.h"

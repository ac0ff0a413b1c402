/* tiny-cuda-nn/common_device.h -- applications include this header for the device-side helpers of the reference
 * (common_device.h:40-1100). The hot path of this library keeps its device code inside libtcnn_b200; what application code of the
 * reference's samples uses from here is the host / launch layer, which lives in common.h. */
#pragma once
#include "common.h"

/* tiny-cuda-nn/network.h -- see config.h (the shim keeps these classes in one header). */
#pragma once
#include "config.h"

// TEST INFRASTRUCTURE: stands in for <cuda_fp16.h> when the device sources are compiled for the host (tests/emu)
#pragma once
#include "cuda_emu.h"

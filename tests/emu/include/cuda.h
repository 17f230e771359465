// TEST INFRASTRUCTURE: stands in for <cuda.h> when the device sources are compiled for the host (tests/emu)
#pragma once
#include "cuda_emu.h"

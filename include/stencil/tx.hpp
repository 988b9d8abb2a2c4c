#pragma once

#include "stencil/tx_common.hpp"

#if STENCIL_USE_CUDA == 1 && defined(__NVCC__)
#include "stencil/tx_cuda.cuh"
#endif

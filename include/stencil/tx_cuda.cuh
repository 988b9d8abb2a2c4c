#pragma once
// The CUDA transport layer.  The reference implements six sender/receiver families here (peer-access
// kernels, pack + cudaMemcpyPeerAsync + unpack, CUDA-IPC colocated variants, host-staged and
// CUDA-aware MPI).  On one NVSwitch node all of them collapse into one mechanism -- each source GPU
// runs a single fused kernel that stores every outgoing halo directly into the neighbours' ghost
// cells (stencil/stencil.hpp, DistributedDomain::exchange) -- so this header only pulls in the pieces
// of the old layer that user code can still name.

#include <mpi.h>

#include <nvToolsExt.h>

#include "stencil/copy.cuh"
#include "stencil/cuda_runtime.hpp"
#include "stencil/local_domain.cuh"
#include "stencil/logging.hpp"
#include "stencil/packer.cuh"
#include "stencil/rcstream.hpp"
#include "stencil/rt.hpp"
#include "stencil/timer.hpp"
#include "stencil/translator.cuh"
#include "stencil/tx_common.hpp"

#pragma once
// LOG_SPEW / DEBUG / INFO / WARN / ERROR / FATAL to stderr, filtered at compile time by
// STENCIL_OUTPUT_LEVEL (5 = everything ... 0 = fatal only).  LOG_FATAL exits with status 1.

#include <cstdlib>
#include <iostream>

#include "stencil/mpi.hpp"

#ifndef STENCIL_OUTPUT_LEVEL
#define STENCIL_OUTPUT_LEVEL 3
#endif

#define STENCIL_LOG_LINE(tag, x)                                                                                       \
  std::cerr << tag "[" << __FILE__ << ":" << __LINE__ << "]{" << mpi::world_rank() << "} " << x << "\n";

#if STENCIL_OUTPUT_LEVEL >= 5
#define LOG_SPEW(x) STENCIL_LOG_LINE("SPEW", x)
#else
#define LOG_SPEW(x)
#endif

#if STENCIL_OUTPUT_LEVEL >= 4
#define LOG_DEBUG(x) STENCIL_LOG_LINE("DEBUG", x)
#else
#define LOG_DEBUG(x)
#endif

#if STENCIL_OUTPUT_LEVEL >= 3
#define LOG_INFO(x) STENCIL_LOG_LINE("INFO", x)
#else
#define LOG_INFO(x)
#endif

#if STENCIL_OUTPUT_LEVEL >= 2
#define LOG_WARN(x) STENCIL_LOG_LINE("WARN", x)
#else
#define LOG_WARN(x)
#endif

#if STENCIL_OUTPUT_LEVEL >= 1
#define LOG_ERROR(x) STENCIL_LOG_LINE("ERROR", x)
#else
#define LOG_ERROR(x)
#endif

#if STENCIL_OUTPUT_LEVEL >= 0
#define LOG_FATAL(x)                                                                                                   \
  STENCIL_LOG_LINE("FATAL", x)                                                                                         \
  exit(1);
#else
#define LOG_FATAL(x) exit(1);
#endif

#pragma once
// NVML glue.  NVML(call) is the library's fatal-error wrapper for nvmlReturn_t (print where, terminate -- the same
// convention as CUDA_RUNTIME), nvml::lazy_init() initialises the library exactly once per process, whichever thread
// asks first.

#include <nvml.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace nvml {

inline void require(nvmlReturn_t status, const char *file, int line) {
  if (status == NVML_SUCCESS) return;
  std::fprintf(stderr, "nvml Error: %s in %s : %d\n", nvmlErrorString(status), file, line);
  std::exit(-1);
}

} // namespace nvml

// the reference's spelling of the same check
inline void checkNvml(nvmlReturn_t status, const char *file, const int line) { nvml::require(status, file, line); }

#define NVML(call) nvml::require((call), __FILE__, __LINE__);

namespace nvml {

inline void lazy_init() {
  static std::once_flag once;
  std::call_once(once, [] { NVML(nvmlInit()); });
}

} // namespace nvml

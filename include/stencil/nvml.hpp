#pragma once

#include <cstdio>
#include <cstdlib>

#include <nvml.h>

inline void checkNvml(nvmlReturn_t result, const char *file, const int line) {
  if (NVML_SUCCESS == result) return;
  std::fprintf(stderr, "nvml Error: %s in %s : %d\n", nvmlErrorString(result), file, line);
  std::exit(-1);
}

#define NVML(stmt) checkNvml(stmt, __FILE__, __LINE__);

namespace nvml {
inline void lazy_init() {
  static bool done = false;
  if (!done) {
    NVML(nvmlInit());
    done = true;
  }
}
} // namespace nvml

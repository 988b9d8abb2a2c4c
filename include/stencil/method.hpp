#pragma once
// Method: bit set selecting which transports the planner may use.  In this implementation every
// message between GPUs of one NVSwitch node is carried by the fused direct-write kernel; the flags
// keep their reference meaning for planning (which messages count as "kernel" / "peer" / "colo" /
// "remote") and for exchange_bytes_for_method().

#include <string>

enum class Method : int {
  None = 0,
  CudaMpi = 1,
  ColoPackMemcpyUnpack = 2,
  ColoQuantityKernel = 4,
  ColoRegionKernel = 8,
  ColoMemcpy3d = 16,
  ColoDomainKernel = 32,
  CudaMemcpyPeer = 64,
  CudaKernel = 128,
  Default = CudaMpi + ColoPackMemcpyUnpack + CudaMemcpyPeer + CudaKernel
};

inline Method operator|(Method a, Method b) { return static_cast<Method>(static_cast<int>(a) | static_cast<int>(b)); }
inline Method &operator|=(Method &a, Method b) { return a = a | b; }
inline Method operator&(Method a, Method b) { return static_cast<Method>(static_cast<int>(a) & static_cast<int>(b)); }
// "any bit in common"
inline bool operator&&(Method a, Method b) { return (a & b) != Method::None; }
inline bool any(Method a) noexcept { return a != Method::None; }

inline std::string to_string(const Method &m) {
  struct Name {
    Method bit;
    const char *text;
  };
#if STENCIL_USE_CUDA_AWARE_MPI == 1
  const char *mpiName = "cuda-aware";
#else
  const char *mpiName = "staged";
#endif
  const Name names[] = {{Method::CudaMpi, mpiName},
                        {Method::ColoPackMemcpyUnpack, "colo-pmu"},
                        {Method::ColoQuantityKernel, "colo-q"},
                        {Method::ColoRegionKernel, "colo-r"},
                        {Method::ColoMemcpy3d, "colo-m3"},
                        {Method::CudaMemcpyPeer, "peer"},
                        {Method::CudaKernel, "kernel"}};
  std::string out;
  for (const Name &n : names) {
    if (m && n.bit) {
      if (!out.empty()) out += "|";
      out += n.text;
    }
  }
  return out;
}

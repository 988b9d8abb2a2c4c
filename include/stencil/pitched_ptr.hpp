#pragma once
// PitchedPtr<T>: a typed cudaPitchedPtr.  Field order {pitch, ptr, xsize, ysize} is ABI: user kernels
// receive it by value inside Accessor (reference include/stencil/pitched_ptr.hpp:15-19).

#include <cassert>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define STENCIL_HD __host__ __device__ __forceinline__
#else
#define STENCIL_HD inline
#endif

template <typename T> struct PitchedPtr {
  size_t pitch; // bytes between rows
  T *ptr;
  size_t xsize; // logical row width in bytes
  size_t ysize; // rows per plane

  PitchedPtr() : pitch(0), ptr(nullptr), xsize(0), ysize(0) {}
  PitchedPtr(size_t pitch_, T *ptr_, size_t xsize_, size_t ysize_) : pitch(pitch_), ptr(ptr_), xsize(xsize_), ysize(ysize_) {
    assert(xsize % sizeof(T) == 0);
  }
  explicit PitchedPtr(const cudaPitchedPtr &p) : pitch(p.pitch), ptr(reinterpret_cast<T *>(p.ptr)), xsize(p.xsize), ysize(p.ysize) {}

  explicit operator cudaPitchedPtr() {
    cudaPitchedPtr p = {};
    p.ptr = ptr;
    p.pitch = pitch;
    p.xsize = xsize;
    p.ysize = ysize;
    return p;
  }

  template <typename U> bool operator!=(const PitchedPtr<U> &o) const noexcept {
    return (const void *)ptr != (const void *)o.ptr || xsize != o.xsize || ysize != o.ysize || pitch != o.pitch;
  }

  STENCIL_HD T &at(size_t x, size_t y, size_t z) noexcept {
    return *reinterpret_cast<T *>(reinterpret_cast<char *>(ptr) + (z * ysize + y) * pitch + x * sizeof(T));
  }
  STENCIL_HD const T &at(size_t x, size_t y, size_t z) const noexcept {
    return *reinterpret_cast<const T *>(reinterpret_cast<const char *>(ptr) + (z * ysize + y) * pitch + x * sizeof(T));
  }
};

#undef STENCIL_HD

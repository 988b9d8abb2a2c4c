#pragma once
// Accessor<T>: index a quantity by GLOBAL grid coordinate.  origin = coordinate of allocation
// element (0,0,0), i.e. subdomain origin minus the negative-side radius.  Layout
// {PitchedPtr<T>, Dim3} is ABI (reference include/stencil/accessor.hpp:14-17).

#include <cassert>

#include "stencil/dim3.hpp"
#include "stencil/pitched_ptr.hpp"

#ifdef __CUDACC__
#define STENCIL_HD __host__ __device__ __forceinline__
#else
#define STENCIL_HD inline
#endif

template <typename T> class Accessor {
  PitchedPtr<T> ptr_;
  Dim3 origin_;

public:
  Accessor(const PitchedPtr<T> &ptr, const Dim3 &origin) : ptr_(ptr), origin_(origin) {}

  // raw pointer + pitch in ELEMENTS (kept for the reference's tests)
  Accessor(T *raw, const Dim3 &origin, const Dim3 &elemPitch)
      : ptr_(elemPitch.x * sizeof(T), raw, elemPitch.x * sizeof(T), elemPitch.y), origin_(origin) {}

  STENCIL_HD T &operator[](const Dim3 &p) noexcept {
    assert(p.x >= origin_.x && p.y >= origin_.y && p.z >= origin_.z);
    return ptr_.at(size_t(p.x - origin_.x), size_t(p.y - origin_.y), size_t(p.z - origin_.z));
  }
  STENCIL_HD const T &operator[](const Dim3 &p) const noexcept {
    assert(p.x >= origin_.x && p.y >= origin_.y && p.z >= origin_.z);
    return ptr_.at(size_t(p.x - origin_.x), size_t(p.y - origin_.y), size_t(p.z - origin_.z));
  }

  const Dim3 &origin() const noexcept { return origin_; }
  const PitchedPtr<T> &ptr() const noexcept { return ptr_; }
};

#undef STENCIL_HD

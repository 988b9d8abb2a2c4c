#pragma once
// LocalDomain: one subdomain on one GPU.  Owns a "current" and a "next" allocation per quantity
// (x fastest, unpitched rows: pitch == raw_size().x * elem_size, because astaroth-style user kernels
// index i + j*mx + k*mx*my) plus device-side arrays of the cudaPitchedPtrs.
// API-compatible with the reference's include/stencil/local_domain.cuh.

#include <cstdint>
#include <string>
#include <vector>

#include "stencil/accessor.hpp"
#include "stencil/cuda_runtime.hpp"
#include "stencil/dim3.hpp"
#include "stencil/geometry.hpp"
#include "stencil/logging.hpp"
#include "stencil/pack_kernel.cuh"
#include "stencil/pitched_ptr.hpp"
#include "stencil/radius.hpp"
#include "stencil/rect3.hpp"

class DistributedDomain;

template <typename T> class DataHandle {
  friend class DistributedDomain;
  friend class LocalDomain;
  size_t id_;
  std::string name_;

public:
  DataHandle(size_t i, const std::string &name = "") : id_(i), name_(name) {}
  size_t id() const noexcept { return id_; }
};

enum class DataType { None, Float, Double };

class LocalDomain {
  friend class DistributedDomain;

  Dim3 sz_;     // compute extent (no ghost cells)
  Dim3 origin_; // global coordinate of the first compute cell
  Radius radius_;

  std::vector<cudaPitchedPtr> currDataPtrs_;
  std::vector<cudaPitchedPtr> nextDataPtrs_;
  std::vector<void *> allocBases_; // what cudaMalloc returned for every allocation above (freed by the destructor)
  std::vector<size_t> dataElemSize_;
  std::vector<std::string> dataName_;

  // device copies of the arrays above (consumed by multi_translate-style kernels)
  cudaPitchedPtr *devCurrDataPtrs_;
  cudaPitchedPtr *devNextDataPtrs_;
  size_t *devDataElemSize_;

  int dev_;

  Dim3 low_ghost() const noexcept { return Dim3(radius_.x(-1), radius_.y(-1), radius_.z(-1)); }

  // HBM layout rule: when every row of quantity i has the same 16-byte phase the allocation starts
  // this many bytes into its cudaMalloc block, so that the first COMPUTE cell of every row is 16-byte aligned.
  size_t lead_bytes(size_t i) const noexcept;

public:
  LocalDomain(Dim3 sz, Dim3 origin, int dev);
  ~LocalDomain();
  // Copies share nothing: a LocalDomain may only be copied BEFORE realize() (DistributedDomain fills its vector that way,
  // like the reference, src/stencil.cu:261-267); copying one that owns device memory would free it twice.
  LocalDomain(const LocalDomain &o);
  LocalDomain &operator=(const LocalDomain &) = delete;
  LocalDomain(LocalDomain &&o) noexcept;

  void set_device(CudaErrorsFatal fatal = CudaErrorsFatal::YES);

  int64_t num_data() const { return int64_t(currDataPtrs_.size()); }
  const Dim3 &origin() const noexcept { return origin_; }

  // add an untyped quantity of n-byte elements; returns its index
  int64_t add_data(size_t n, const std::string &name = "") {
    dataName_.push_back(name);
    dataElemSize_.push_back(n);
    currDataPtrs_.push_back(cudaPitchedPtr{});
    nextDataPtrs_.push_back(cudaPitchedPtr{});
    return int64_t(dataElemSize_.size()) - 1;
  }
  template <typename T> DataHandle<T> add_data(const std::string &name = "") { return DataHandle<T>(add_data(sizeof(T), name), name); }

  void set_radius(size_t r) { radius_ = Radius::constant(r); }
  void set_radius(const Radius &r) { radius_ = r; }
  const Radius &radius() const noexcept { return radius_; }

  template <typename T> PitchedPtr<T> get_curr(const DataHandle<T> handle) const {
    assert(handle.id_ < currDataPtrs_.size() && sizeof(T) == dataElemSize_[handle.id_]);
    return PitchedPtr<T>(currDataPtrs_[handle.id_]);
  }
  template <typename T> PitchedPtr<T> get_next(const DataHandle<T> handle) const {
    assert(handle.id_ < nextDataPtrs_.size() && sizeof(T) == dataElemSize_[handle.id_]);
    return PitchedPtr<T>(nextDataPtrs_[handle.id_]);
  }

  size_t elem_size(const size_t idx) const {
    assert(idx < dataElemSize_.size());
    return dataElemSize_[idx];
  }
  const std::vector<size_t> &elem_sizes() const { return dataElemSize_; }
  const size_t *dev_elem_sizes() const { return devDataElemSize_; }

  cudaPitchedPtr curr_data(size_t idx) const {
    assert(idx < currDataPtrs_.size());
    return currDataPtrs_[idx];
  }
  cudaPitchedPtr next_data(size_t idx) const {
    assert(idx < nextDataPtrs_.size());
    return nextDataPtrs_[idx];
  }
  const std::vector<cudaPitchedPtr> &curr_datas() const noexcept { return currDataPtrs_; }
  const std::vector<cudaPitchedPtr> &next_datas() const noexcept { return nextDataPtrs_; }
  cudaPitchedPtr *dev_curr_datas() const { return devCurrDataPtrs_; }
  cudaPitchedPtr *dev_next_datas() const { return devNextDataPtrs_; }

  // accessors index by GLOBAL coordinate; element (0,0,0) of the allocation is origin - low ghost
  template <typename T> Accessor<T> get_curr_accessor(const DataHandle<T> &dh) const noexcept {
    return Accessor<T>(get_curr(dh), origin_ - low_ghost());
  }
  template <typename T> Accessor<T> get_next_accessor(const DataHandle<T> &dh) const noexcept {
    return Accessor<T>(get_next(dh), origin_ - low_ghost());
  }

  Rect3 get_compute_region() const noexcept;
  // compute region grown by the face radii
  Rect3 get_full_region() const noexcept {
    return Rect3(origin_ - low_ghost(), origin_ + sz_ + Dim3(radius_.x(1), radius_.y(1), radius_.z(1)));
  }

  // allocation-relative position of the region on side `dir` (halo: ghost cells; !halo: outermost
  // compute cells); dir (0,0,0) = the compute region
  static Dim3 halo_pos(const Dim3 &dir, const Dim3 &sz, const Radius &radius, const bool halo) noexcept {
    return stencil::geom::halo_pos(dir, sz, radius, halo);
  }
  Dim3 halo_pos(const Dim3 &dir, const bool halo) const noexcept { return halo_pos(dir, sz_, radius_, halo); }

  // global coordinates of that region
  Rect3 halo_coords(const Dim3 &dir, const bool halo) const;

  static Dim3 halo_extent(const Dim3 &dir, const Dim3 &sz, const Radius &radius) {
    return stencil::geom::halo_extent(dir, sz, radius);
  }
  Dim3 halo_extent(const Dim3 &dir) const noexcept { return halo_extent(dir, sz_, radius_); }

  int64_t halo_bytes(const Dim3 &dir, const int64_t idx) const noexcept {
    return int64_t(dataElemSize_[idx] * halo_extent(dir).flatten());
  }

  Dim3 size() const noexcept { return sz_; }
  Dim3 raw_size() const noexcept { return stencil::geom::raw_size(sz_, radius_); }
  int gpu() const { return dev_; }

  void swap() noexcept;

  // bytes of the box [pos, pos+ext) of quantity qi, packed x-fastest
  std::vector<unsigned char> region_to_host(const Dim3 &pos, const Dim3 &ext, const size_t qi) const;
  std::vector<unsigned char> interior_to_host(const size_t qi) const {
    return region_to_host(halo_pos(Dim3(0, 0, 0), true), halo_extent(Dim3(0, 0, 0)), qi);
  }
  std::vector<unsigned char> quantity_to_host(const size_t qi) const { return region_to_host(Dim3(0, 0, 0), raw_size(), qi); }

  void realize();
};

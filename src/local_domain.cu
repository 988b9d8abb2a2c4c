#include "stencil/local_domain.cuh"

#include "stencil_b200.h"

#include <nvToolsExt.h>

#include <cstdlib>

size_t LocalDomain::lead_bytes(size_t i) const noexcept {
  static const bool enabled = [] {
    const char *e = std::getenv("SB_ALLOC_ALIGN");
    return !(e && e[0] == '0');
  }();
  const size_t es = dataElemSize_[i];
  const size_t rowBytes = size_t(radius_.x(-1) + sz_.x + radius_.x(1)) * es;
  if (!enabled || rowBytes % 16 != 0) return 0;
  return (16 - (size_t(radius_.x(-1)) * es) % 16) % 16;
}

LocalDomain::LocalDomain(Dim3 sz, Dim3 origin, int dev)
    : sz_(sz), origin_(origin), radius_(Radius::constant(0)), devCurrDataPtrs_(nullptr), devNextDataPtrs_(nullptr),
      devDataElemSize_(nullptr), dev_(dev) {}

LocalDomain::LocalDomain(const LocalDomain &o)
    : sz_(o.sz_), origin_(o.origin_), radius_(o.radius_), currDataPtrs_(o.currDataPtrs_), nextDataPtrs_(o.nextDataPtrs_),
      dataElemSize_(o.dataElemSize_), dataName_(o.dataName_), devCurrDataPtrs_(nullptr), devNextDataPtrs_(nullptr), devDataElemSize_(nullptr),
      dev_(o.dev_) {
  if (!o.allocBases_.empty() || o.devCurrDataPtrs_) LOG_FATAL("LocalDomain copied after realize(): the copy would free the same device memory again");
}

LocalDomain::LocalDomain(LocalDomain &&o) noexcept
    : sz_(o.sz_), origin_(o.origin_), radius_(o.radius_), currDataPtrs_(std::move(o.currDataPtrs_)), nextDataPtrs_(std::move(o.nextDataPtrs_)),
      allocBases_(std::move(o.allocBases_)), dataElemSize_(std::move(o.dataElemSize_)), dataName_(std::move(o.dataName_)),
      devCurrDataPtrs_(o.devCurrDataPtrs_), devNextDataPtrs_(o.devNextDataPtrs_), devDataElemSize_(o.devDataElemSize_), dev_(o.dev_) {
  o.allocBases_.clear();
  o.devCurrDataPtrs_ = o.devNextDataPtrs_ = nullptr;
  o.devDataElemSize_ = nullptr;
}

LocalDomain::~LocalDomain() {
  // frees exactly what realize() allocated (the block addresses recorded there: set_radius() after realize() must not
  // change what is passed to cudaFree)
  if (allocBases_.empty() && !devCurrDataPtrs_ && !devNextDataPtrs_ && !devDataElemSize_) return;
  CUDA_RUNTIME(cudaSetDevice(dev_));
  for (void *b : allocBases_) CUDA_RUNTIME(cudaFree(b));
  if (devCurrDataPtrs_) CUDA_RUNTIME(cudaFree(devCurrDataPtrs_));
  if (devNextDataPtrs_) CUDA_RUNTIME(cudaFree(devNextDataPtrs_));
  if (devDataElemSize_) CUDA_RUNTIME(cudaFree(devDataElemSize_));
}

void LocalDomain::set_device(CudaErrorsFatal fatal) {
  const cudaError_t err = cudaSetDevice(dev_);
  if (CudaErrorsFatal::YES == fatal) {
    CUDA_RUNTIME(err);
  }
}

Rect3 LocalDomain::get_compute_region() const noexcept { return Rect3(origin_, origin_ + sz_); }

Rect3 LocalDomain::halo_coords(const Dim3 &dir, const bool halo) const {
  // allocation-relative -> global: allocation element (0,0,0) sits at origin - low ghost
  const Dim3 lo = halo_pos(dir, halo) - low_ghost() + origin_;
  return Rect3(lo, lo + halo_extent(dir));
}

void LocalDomain::swap() noexcept {
  nvtxRangePush("swap");
  currDataPtrs_.swap(nextDataPtrs_);
  std::swap(devCurrDataPtrs_, devNextDataPtrs_);
  nvtxRangePop();
}

std::vector<unsigned char> LocalDomain::region_to_host(const Dim3 &pos, const Dim3 &ext, const size_t qi) const {
  const size_t es = elem_size(qi);
  std::vector<unsigned char> host(es * ext.flatten());
  if (host.empty()) return host;
  CUDA_RUNTIME(cudaSetDevice(dev_));
  void *dense = nullptr;
  CUDA_RUNTIME(cudaMalloc(&dense, host.size()));
  const cudaPitchedPtr c = curr_data(qi);
  const int64_t p[3] = {pos.x, pos.y, pos.z}, e[3] = {ext.x, ext.y, ext.z};
  if (SB_OK != sb_pack(dense, sb_pitched{c.ptr, int64_t(c.pitch), int64_t(c.ysize)}, p, e, int64_t(es), nullptr)) {
    LOG_FATAL("region_to_host: " << sb_last_error());
  }
  CUDA_RUNTIME(cudaMemcpy(host.data(), dense, host.size(), cudaMemcpyDeviceToHost));
  CUDA_RUNTIME(cudaFree(dense));
  return host;
}

void LocalDomain::realize() {
  CUDA_RUNTIME(cudaSetDevice(dev_));
  const Dim3 raw = raw_size();
  for (int64_t i = 0; i < num_data(); ++i) {
    const size_t rowBytes = size_t(raw.x) * dataElemSize_[i];
    const size_t bytes = rowBytes * size_t(raw.y) * size_t(raw.z);
    for (std::vector<cudaPitchedPtr> *store : {&currDataPtrs_, &nextDataPtrs_}) {
      cudaPitchedPtr p{};
      CUDA_RUNTIME(cudaMalloc(&p.ptr, bytes + 32));
      CUDA_RUNTIME(cudaMemset(p.ptr, 0, bytes + 32));
      allocBases_.push_back(p.ptr);
      p.ptr = static_cast<char *>(p.ptr) + lead_bytes(size_t(i));
      p.pitch = rowBytes; // unpitched on purpose, see the header
      p.xsize = rowBytes;
      p.ysize = size_t(raw.y);
      (*store)[i] = p;
    }
  }
  auto upload = [](const void *host, size_t bytes) {
    void *d = nullptr;
    CUDA_RUNTIME(cudaMalloc(&d, bytes ? bytes : 1));
    CUDA_RUNTIME(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    return d;
  };
  devCurrDataPtrs_ = static_cast<cudaPitchedPtr *>(upload(currDataPtrs_.data(), currDataPtrs_.size() * sizeof(cudaPitchedPtr)));
  devNextDataPtrs_ = static_cast<cudaPitchedPtr *>(upload(nextDataPtrs_.data(), nextDataPtrs_.size() * sizeof(cudaPitchedPtr)));
  devDataElemSize_ = static_cast<size_t *>(upload(dataElemSize_.data(), dataElemSize_.size() * sizeof(size_t)));
}

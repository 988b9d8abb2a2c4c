#pragma once
// Translator: a repeatable set of strided 3-D box copies across several allocations with possibly
// different element sizes.  All four flavours of the reference API (one kernel per transfer,
// cudaMemcpy3D, one kernel per region group, one kernel per domain) are served by the same
// single-launch copy plan here; the class names remain so existing call sites compile.

#include <cuda_runtime.h>

#include <vector>

#include "stencil/dim3.hpp"

struct sb_copy_plan;

class Translator {
public:
  // the same logical box in n allocations (host arrays of length n)
  struct RegionParams {
    cudaPitchedPtr *dstPtrs;
    Dim3 dstPos;
    const cudaPitchedPtr *srcPtrs;
    Dim3 srcPos;
    Dim3 extent;
    const size_t *elemSizes;
    int64_t n;

    RegionParams() = default;
  };

  Translator();
  virtual ~Translator();

  virtual void prepare(const std::vector<RegionParams> &params) = 0;

  // run all prepared copies, asynchronously in `stream`
  void async(cudaStream_t stream);

protected:
  struct Param {
    cudaPitchedPtr dstPtr;
    Dim3 dstPos;
    cudaPitchedPtr srcPtr;
    Dim3 srcPos;
    Dim3 extent;
    size_t elemSize;
    Param(const cudaPitchedPtr &d, const Dim3 &dp, const cudaPitchedPtr &s, const Dim3 &sp, const Dim3 &e, const size_t es)
        : dstPtr(d), dstPos(dp), srcPtr(s), srcPos(sp), extent(e), elemSize(es) {}
  };
  static std::vector<Param> convert(const std::vector<RegionParams> &params);

  // (re)build the plan on `device`
  void build(const std::vector<RegionParams> &params, int device);

private:
  sb_copy_plan *plan_;
};

class TranslatorKernel : public Translator {
  int device_;

public:
  TranslatorKernel(int device);
  void prepare(const std::vector<RegionParams> &params) override;
};

class TranslatorMemcpy3D : public Translator {
public:
  void prepare(const std::vector<RegionParams> &params) override; // runs on the current device
};

class TranslatorMultiKernel : public Translator {
  int device_;

public:
  TranslatorMultiKernel(int device);
  ~TranslatorMultiKernel();
  void prepare(const std::vector<RegionParams> &params) override;
};

class TranslatorDomainKernel : public Translator {
  int device_;

public:
  TranslatorDomainKernel(int device);
  ~TranslatorDomainKernel();
  TranslatorDomainKernel(const TranslatorDomainKernel &other) = delete;
  void prepare(const std::vector<RegionParams> &params) override;
};

// Translator family on the box-copy engine (see stencil/translator.cuh).
#include "stencil/translator.cuh"

#include "stencil_b200.h"

#include "stencil/cuda_runtime.hpp"
#include "stencil/logging.hpp"

Translator::Translator() : plan_(nullptr) {}

Translator::~Translator() {
  if (plan_) sb_copy_plan_destroy(plan_);
}

std::vector<Translator::Param> Translator::convert(const std::vector<RegionParams> &params) {
  std::vector<Param> out;
  for (const RegionParams &r : params)
    for (int64_t i = 0; i < r.n; ++i) out.emplace_back(r.dstPtrs[i], r.dstPos, r.srcPtrs[i], r.srcPos, r.extent, r.elemSizes[i]);
  return out;
}

void Translator::build(const std::vector<RegionParams> &params, int device) {
  if (plan_) {
    sb_copy_plan_destroy(plan_);
    plan_ = nullptr;
  }
  std::vector<sb_box_copy> copies;
  for (const Param &p : convert(params)) {
    sb_box_copy c{};
    c.dst = sb_pitched{p.dstPtr.ptr, int64_t(p.dstPtr.pitch), int64_t(p.dstPtr.ysize)};
    c.src = sb_pitched{p.srcPtr.ptr, int64_t(p.srcPtr.pitch), int64_t(p.srcPtr.ysize)};
    const Dim3 *pos[2] = {&p.dstPos, &p.srcPos};
    int64_t *out[2] = {c.dst_pos, c.src_pos};
    for (int k = 0; k < 2; ++k) {
      out[k][0] = pos[k]->x;
      out[k][1] = pos[k]->y;
      out[k][2] = pos[k]->z;
    }
    c.extent[0] = p.extent.x;
    c.extent[1] = p.extent.y;
    c.extent[2] = p.extent.z;
    c.elem_size = int64_t(p.elemSize);
    copies.push_back(c);
  }
  if (SB_OK != sb_copy_plan_create(&plan_, device, copies.data(), int64_t(copies.size()))) {
    LOG_FATAL("translator: " << sb_last_error());
  }
}

void Translator::async(cudaStream_t stream) {
  if (!plan_) {
    LOG_FATAL("Translator::async before prepare()");
  }
  if (SB_OK != sb_copy_plan_launch(plan_, stream)) {
    LOG_FATAL("translator: " << sb_last_error());
  }
}

TranslatorKernel::TranslatorKernel(int device) : device_(device) {}
void TranslatorKernel::prepare(const std::vector<RegionParams> &params) { build(params, device_); }

void TranslatorMemcpy3D::prepare(const std::vector<RegionParams> &params) {
  int dev = 0;
  CUDA_RUNTIME(cudaGetDevice(&dev));
  build(params, dev);
}

TranslatorMultiKernel::TranslatorMultiKernel(int device) : device_(device) {}
TranslatorMultiKernel::~TranslatorMultiKernel() {}
void TranslatorMultiKernel::prepare(const std::vector<RegionParams> &params) { build(params, device_); }

TranslatorDomainKernel::TranslatorDomainKernel(int device) : device_(device) {}
TranslatorDomainKernel::~TranslatorDomainKernel() {}
void TranslatorDomainKernel::prepare(const std::vector<RegionParams> &params) { build(params, device_); }

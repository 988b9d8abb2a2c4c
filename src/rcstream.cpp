#include "stencil/rcstream.hpp"

#include "stencil/logging.hpp"

RcStream::RcStream(int dev, Priority requestedPriority) : refs_(new size_t(1)), dev_(dev), stream_(0) {
  CUDA_RUNTIME(cudaSetDevice(dev_));
  int least = 0, greatest = 0; // numerically lower = more urgent
  CUDA_RUNTIME(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  if (least == greatest && Priority::HIGH == requestedPriority) {
    LOG_WARN("stream priority not supported");
  }
  const int prio = (Priority::HIGH == requestedPriority) ? greatest : 0;
  CUDA_RUNTIME(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, prio));
}

void RcStream::release() {
  if (0 == stream_) return; // moved-from
  if (0 == --*refs_) {
    CUDA_RUNTIME(cudaSetDevice(dev_));
    CUDA_RUNTIME(cudaStreamDestroy(stream_));
    delete refs_;
  }
  stream_ = 0;
}

#pragma once
// RcStream: a reference-counted non-blocking CUDA stream bound to a device, optionally at the
// highest priority.  Converts implicitly to cudaStream_t.

#include <cuda_runtime.h>

#include <cstddef>
#include <iostream>

#include "stencil/cuda_runtime.hpp"

class RcStream {
public:
  enum class Priority { DEFAULT, HIGH };

private:
  size_t *refs_;
  int dev_;
  cudaStream_t stream_;

  void release(); // drop one reference; destroy the stream with the last one

public:
  RcStream(int dev, Priority requestedPriority = Priority::DEFAULT);
  RcStream() : RcStream(0) {}
  ~RcStream() { release(); }

  RcStream(const RcStream &o) : refs_(o.refs_), dev_(o.dev_), stream_(o.stream_) { ++*refs_; }
  RcStream(RcStream &&o) : refs_(o.refs_), dev_(o.dev_), stream_(o.stream_) { o.stream_ = 0; }
  RcStream &operator=(const RcStream &o) {
    if (this != &o) {
      release();
      refs_ = o.refs_;
      dev_ = o.dev_;
      stream_ = o.stream_;
      ++*refs_;
    }
    return *this;
  }
  RcStream &operator=(RcStream &&o) {
    if (this != &o) {
      release();
      refs_ = o.refs_;
      dev_ = o.dev_;
      stream_ = o.stream_;
      o.stream_ = 0;
    }
    return *this;
  }

  operator cudaStream_t() const noexcept { return stream_; }
  int device() const noexcept { return dev_; }
  bool operator==(const RcStream &o) const noexcept { return stream_ == o.stream_; }
};

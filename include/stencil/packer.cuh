#pragma once
// DevicePacker / DeviceUnpacker: gather all quantities of a set of halo messages of one LocalDomain
// into (out of) ONE contiguous device buffer in a fixed wire order:
//   messages sorted by Message::by_size; per message, for each quantity:
//   offset = next_align_of(offset, elem_size); then extent.flatten() * elem_size bytes, x fastest.
// prepare() builds a persistent copy plan; pack()/unpack() are ONE kernel launch each (the reference
// replays a CUDA graph of one launch per message).  Plans exist for both swap parities, so packing
// after LocalDomain::swap() reads the new "current" buffers.

#include <thread>
#include <vector>

#include "stencil/align.cuh"
#include "stencil/local_domain.cuh"
#include "stencil/logging.hpp"
#include "stencil/tx_common.hpp"

struct sb_copy_plan;

class Packer {
public:
  virtual void prepare(LocalDomain *domain, const std::vector<Message> &messages) = 0;
  virtual void pack() = 0;
  virtual int64_t size() = 0; // bytes
  virtual void *data() = 0;
  virtual ~Packer() {}
};

class Unpacker {
public:
  virtual void prepare(LocalDomain *domain, const std::vector<Message> &messages) = 0;
  virtual void unpack() = 0;
  virtual int64_t size() = 0;
  virtual void *data() = 0;
  virtual ~Unpacker() {}
};

namespace stencil {
namespace detail {
// shared by packer and unpacker: the plan for each identity of "curr"
struct PackPlans {
  sb_copy_plan *plan[2] = {nullptr, nullptr};
  void *currAtPrepare = nullptr; // curr pointer of quantity 0 when plan[0] was built
  void destroy();
};
} // namespace detail
} // namespace stencil

class DevicePacker : public Packer {
  LocalDomain *domain_;
  std::vector<Message> dirs_;
  int64_t size_;
  char *devBuf_;
  cudaStream_t stream_; // not owned
  stencil::detail::PackPlans plans_;

public:
  DevicePacker(cudaStream_t stream);
  ~DevicePacker();

  void prepare(LocalDomain *domain, const std::vector<Message> &messages) override;
  void pack() override;
  int64_t size() override { return size_; }
  void *data() override { return devBuf_; }
};

class DeviceUnpacker : public Unpacker {
  LocalDomain *domain_;
  std::vector<Message> dirs_;
  int64_t size_;
  char *devBuf_;
  cudaStream_t stream_;
  stencil::detail::PackPlans plans_;

public:
  DeviceUnpacker(cudaStream_t stream);
  ~DeviceUnpacker();

  void prepare(LocalDomain *domain, const std::vector<Message> &messages) override;
  void unpack() override;
  int64_t size() override { return size_; }
  void *data() override { return devBuf_; }
};

#pragma once
// Message (one halo region sent in one direction), MPI tag construction, and the abstract
// sender / receiver state-machine interfaces of the transport layer.

#include "stencil/dim3.hpp"

#include <cassert>
#include <climits>
#include <cstdint>
#include <iostream>
#include <vector>

inline uint16_t ipc_tag_payload(uint8_t a, uint8_t b) noexcept { return uint16_t((uint16_t(a) << 8) | uint16_t(b)); }

class Message {
  Dim3 ext_; // only used to order messages (bigger first)

public:
  Dim3 dir_;
  int srcGPU_;
  int dstGPU_;

  Message(Dim3 dir, int srcGPU, int dstGPU) : Message(dir, srcGPU, dstGPU, Dim3(0, 0, 0)) {}
  Message(Dim3 dir, int srcGPU, int dstGPU, Dim3 ext) : ext_(ext), dir_(dir), srcGPU_(srcGPU), dstGPU_(dstGPU) {}

  // strict weak order: larger element count first, equal counts by direction.  Sender and
  // receiver sort with this so the packed wire order is identical on both sides.
  static bool by_size(const Message &a, const Message &b) noexcept {
    const size_t na = a.ext_.flatten(), nb = b.ext_.flatten();
    return na != nb ? na > nb : a < b;
  }

  bool operator<(const Message &o) const noexcept { return dir_ < o.dir_; }
  bool operator==(const Message &o) const noexcept { return dir_ == o.dir_ && srcGPU_ == o.srcGPU_ && dstGPU_ == o.dstGPU_; }
};

enum class MsgKind {
  ColocatedEvt = 0,
  ColocatedCurrMem = 1,
  ColocatedNextMem = 2,
  ColocatedBuf = 3,
  ColocatedDev = 4,
  ColocatedNotify = 5,
  ColocatedPtr = 6,
  Other = 7,
};

namespace stencil {
namespace detail {
// 2 bits per axis: 0 -> 00, +1 -> 01, -1 -> 10; x in the low bits
inline int dir_bits(const Dim3 &d) {
  assert(d.all_gt(-2) && d.all_lt(2));
  auto two = [](int64_t c) { return c == 0 ? 0 : (c == 1 ? 1 : 2); };
  return two(d.x) | (two(d.y) << 2) | (two(d.z) << 4);
}
} // namespace detail
} // namespace stencil

// 23-bit tag: [0,3) kind, [3,9) direction, [9,24) payload
template <MsgKind kind> inline int make_tag(const int payload, const Dim3 dir = Dim3(0, 0, 0)) {
  assert(payload >= 0 && payload < (1 << 15));
  const int tag = static_cast<int>(kind) | (stencil::detail::dir_bits(dir) << 3) | ((payload & 0x7FFF) << 9);
  assert(tag >= 0);
  return tag;
}

// 31-bit tag: [0,16) data index, [16,24) gpu, [24,31) direction
inline int make_tag(int gpu, int idx, Dim3 dir) {
  static_assert(sizeof(int) == 4, "int is the wrong size");
  assert(gpu >= 0 && gpu < (1 << 8));
  assert(idx >= 0 && idx < (1 << 16));
  const int tag = (idx & 0xFFFF) | ((gpu & 0xFF) << 16) | (stencil::detail::dir_bits(dir) << 24);
  assert(tag >= 0 && "tag must be non-negative");
  return tag;
}

// A sender with several phases:  send(); while (active()) if (next_ready()) next();  wait();
class StatefulSender {
public:
  virtual void start_prepare(const std::vector<Message> &outbox) = 0;
  virtual void finish_prepare() = 0;
  virtual void send() = 0;
  virtual bool active() = 0;
  virtual bool next_ready() = 0;
  virtual void next() = 0;
  virtual void wait() = 0;
  virtual ~StatefulSender() {}
};

class StatefulRecver {
public:
  virtual void start_prepare(const std::vector<Message> &inbox) = 0;
  virtual void finish_prepare() = 0;
  virtual void recv() = 0;
  virtual bool active() = 0;
  virtual bool next_ready() = 0;
  virtual void next() = 0;
  virtual void wait() = 0;
  virtual ~StatefulRecver() {}
};

// DevicePacker / DeviceUnpacker on the box-copy engine: one kernel launch per pack()/unpack().
#include "stencil/packer.cuh"

#include "stencil_b200.h"

#include <algorithm>

namespace {

struct WireEntry {
  Dim3 dir;
  int64_t q;
  int64_t offset; // byte offset in the packed buffer
};

// the wire layout shared by both sides: messages by_size, quantities aligned to their element size
int64_t wire_layout(const LocalDomain &dom, const std::vector<Message> &sorted, std::vector<WireEntry> *entries) {
  int64_t offset = 0;
  for (const Message &msg : sorted) {
    for (int64_t qi = 0; qi < dom.num_data(); ++qi) {
      offset = int64_t(next_align_of(size_t(offset), dom.elem_size(qi)));
      if (entries) entries->push_back(WireEntry{msg.dir_, qi, offset});
      // a send in +x fills the receiver's -x halo: the message has the extent of the halo on side -dir
      offset += dom.halo_bytes(msg.dir_ * -1, qi);
    }
  }
  return offset;
}

sb_pitched as_sb(const cudaPitchedPtr &p) { return sb_pitched{p.ptr, int64_t(p.pitch), int64_t(p.ysize)}; }

void set3(int64_t out[3], const Dim3 &d) {
  out[0] = d.x;
  out[1] = d.y;
  out[2] = d.z;
}

// build the plan for the domain's present "curr" (which = 0) or "next" (which = 1) allocations
sb_copy_plan *build_plan(const LocalDomain &dom, const std::vector<Message> &sorted, char *buf, bool packing, int which) {
  std::vector<WireEntry> wire;
  wire_layout(dom, sorted, &wire);
  std::vector<sb_box_copy> copies;
  for (const WireEntry &w : wire) {
    const Dim3 recvSide = w.dir * -1;
    const Dim3 ext = dom.halo_extent(recvSide);
    if (0 == ext.flatten()) {
      LOG_FATAL("asked to pack for direction " << w.dir << " but computed message size is 0, ext=" << ext);
    }
    const cudaPitchedPtr q = (0 == which) ? dom.curr_data(size_t(w.q)) : dom.next_data(size_t(w.q));
    const int64_t es = int64_t(dom.elem_size(size_t(w.q)));
    sb_box_copy c{};
    const sb_pitched dense{buf + w.offset, ext.x * es, ext.y};
    if (packing) {
      c.dst = dense;
      c.src = as_sb(q);
      set3(c.src_pos, dom.halo_pos(w.dir, false)); // outermost compute cells on side dir
    } else {
      c.dst = as_sb(q);
      c.src = dense;
      set3(c.dst_pos, dom.halo_pos(recvSide, true)); // ghost cells on the opposite side
    }
    set3(c.extent, ext);
    c.elem_size = es;
    copies.push_back(c);
  }
  sb_copy_plan *plan = nullptr;
  if (SB_OK != sb_copy_plan_create(&plan, dom.gpu(), copies.data(), int64_t(copies.size()))) {
    LOG_FATAL("packer: " << sb_last_error());
  }
  return plan;
}

} // namespace

void stencil::detail::PackPlans::destroy() {
  for (sb_copy_plan *&p : plan) {
    if (p) sb_copy_plan_destroy(p);
    p = nullptr;
  }
}

// ---------------------------------------------------------------------------------------- packer
DevicePacker::DevicePacker(cudaStream_t stream) : domain_(nullptr), size_(-1), devBuf_(nullptr), stream_(stream) {}

DevicePacker::~DevicePacker() {
  plans_.destroy();
  if (devBuf_ && domain_) {
    cudaSetDevice(domain_->gpu());
    cudaFree(devBuf_);
  }
}

void DevicePacker::prepare(LocalDomain *domain, const std::vector<Message> &messages) {
  domain_ = domain;
  dirs_ = messages;
  std::sort(dirs_.begin(), dirs_.end(), Message::by_size);
  size_ = wire_layout(*domain_, dirs_, nullptr);
  if (0 == size_) {
    LOG_FATAL("zero-size packer was prepared");
  }
  CUDA_RUNTIME(cudaSetDevice(domain_->gpu()));
  CUDA_RUNTIME(cudaMalloc(&devBuf_, size_t(size_)));
  plans_.currAtPrepare = domain_->num_data() ? domain_->curr_data(0).ptr : nullptr;
  plans_.plan[0] = build_plan(*domain_, dirs_, devBuf_, true, 0);
  plans_.plan[1] = build_plan(*domain_, dirs_, devBuf_, true, 1);
}

void DevicePacker::pack() {
  assert(size_ > 0);
  // after an odd number of swap()s the buffers that were "next" at prepare() are current
  const int which = (domain_->num_data() && domain_->curr_data(0).ptr != plans_.currAtPrepare) ? 1 : 0;
  if (SB_OK != sb_copy_plan_launch(plans_.plan[which], stream_)) {
    LOG_FATAL("pack: " << sb_last_error());
  }
}

// ---------------------------------------------------------------------------------------- unpacker
DeviceUnpacker::DeviceUnpacker(cudaStream_t stream) : domain_(nullptr), size_(-1), devBuf_(nullptr), stream_(stream) {}

DeviceUnpacker::~DeviceUnpacker() {
  plans_.destroy();
  if (devBuf_ && domain_) {
    cudaSetDevice(domain_->gpu());
    cudaFree(devBuf_);
  }
}

void DeviceUnpacker::prepare(LocalDomain *domain, const std::vector<Message> &messages) {
  domain_ = domain;
  dirs_ = messages;
  std::sort(dirs_.begin(), dirs_.end(), Message::by_size); // same order as the sender packed
  size_ = wire_layout(*domain_, dirs_, nullptr);
  if (0 == size_) {
    LOG_FATAL("0-size packer was prepared");
  }
  CUDA_RUNTIME(cudaSetDevice(domain_->gpu()));
  CUDA_RUNTIME(cudaMalloc(&devBuf_, size_t(size_)));
  plans_.currAtPrepare = domain_->num_data() ? domain_->curr_data(0).ptr : nullptr;
  plans_.plan[0] = build_plan(*domain_, dirs_, devBuf_, false, 0);
  plans_.plan[1] = build_plan(*domain_, dirs_, devBuf_, false, 1);
}

void DeviceUnpacker::unpack() {
  assert(size_ > 0);
  const int which = (domain_->num_data() && domain_->curr_data(0).ptr != plans_.currAtPrepare) ? 1 : 0;
  if (SB_OK != sb_copy_plan_launch(plans_.plan[which], stream_)) {
    LOG_FATAL("unpack: " << sb_last_error());
  }
}

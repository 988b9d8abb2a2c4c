// DistributedDomain for one NVSwitch node (see include/stencil/stencil.hpp).
//
// Where the reference plans, per message, one of six sender/receiver state machines and then polls
// them from the host (reference src/stencil.cu:327-464, 1002-1186), this implementation plans ONE
// fused copy kernel per local subdomain: every outgoing (direction x quantity) halo region is a
// segment of that kernel, read from the subdomain's outermost compute cells and stored straight
// into the ghost cells of the destination subdomain -- same GPU, or a peer GPU through the
// NVLink/NVSwitch peer mapping.  exchange() = N launches + N stream syncs, nothing else.
#include "stencil/stencil.hpp"

#include "stencil_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace {

sb_pitched as_sb(const cudaPitchedPtr &p) { return sb_pitched{p.ptr, int64_t(p.pitch), int64_t(p.ysize)}; }

void set3(int64_t out[3], const Dim3 &d) {
  out[0] = d.x;
  out[1] = d.y;
  out[2] = d.z;
}

} // namespace

DistributedDomain::DistributedDomain(size_t x, size_t y, size_t z)
    : size_(x, y, z), placement_(nullptr), radius_(Radius::constant(0)), flags_(Method::Default),
      strategy_(PlacementStrategy::NodeAware), boundary_(Topology::Boundary::PERIODIC), parity_(0), mailbox_(nullptr), epoch_(0), numBytesCudaMpi_(0), numBytesColoDirectAccess_(0),
      numBytesColoPackMemcpyUnpack_(0), numBytesCudaMemcpyPeer_(0), numBytesCudaKernel_(0) {
#ifdef STENCIL_SETUP_STATS
  timeMpiTopo_ = timeNodeGpus_ = timePeerEn_ = timePlacement_ = timePlan_ = timeRealize_ = timeCreate_ = 0;
#endif
#ifdef STENCIL_EXCHANGE_STATS
  timeExchange_ = timeSwap_ = 0;
#endif
  MPI_Comm_rank(MPI_COMM_WORLD, &rank_);
  MPI_Comm_size(MPI_COMM_WORLD, &worldSize_);
  if (const char *s = std::getenv("STENCIL_OUTPUT_PREFIX")) outputPrefix_ = s;

  mpiTopology_ = MpiTopology(MPI_COMM_WORLD);

  int deviceCount = 0;
  CUDA_RUNTIME(cudaGetDeviceCount(&deviceCount));
  if (deviceCount < 1) {
    LOG_FATAL("no CUDA device: this library has no CPU path");
  }

  // default device choice: colocated ranks deal the node's GPUs round-robin; more ranks than GPUs share
  const int coloSize = mpiTopology_.colocated_size(), coloRank = mpiTopology_.colocated_rank();
  if (coloSize <= deviceCount) {
    for (int id = coloRank; id < deviceCount; id += coloSize) gpus_.push_back(id);
  } else {
    gpus_.push_back(coloRank % deviceCount);
  }

  // all-pairs peer access between the devices of this node
  nvtxRangePush("peer_en");
  for (int a = 0; a < deviceCount; ++a)
    for (int b = 0; b < deviceCount; ++b) gpu_topo::enable_peer(a, b);
  nvtxRangePop();
  CUDA_RUNTIME(cudaGetLastError());
}

DistributedDomain::~DistributedDomain() {
  close_ranks();
  destroy_plans();
  streams_.clear();
  delete placement_;
  placement_ = nullptr;
}

void DistributedDomain::destroy_plans() {
  for (auto *sides : {&plans_, &unpackPlans_}) {
    for (int parity = 0; parity < 2; ++parity) {
      for (sb_copy_plan *p : (*sides)[parity])
        if (p) sb_copy_plan_destroy(p);
      (*sides)[parity].clear();
    }
  }
  for (size_t i = 0; i < stagingBufs_.size(); ++i) {
    cudaSetDevice(stagingDevs_[i]);
    cudaFree(stagingBufs_[i]);
  }
  stagingBufs_.clear();
  stagingDevs_.clear();
  for (size_t i = 0; i < phase1Done_.size(); ++i) {
    cudaSetDevice(domains_[i].gpu());
    cudaEventDestroy(phase1Done_[i]);
  }
  phase1Done_.clear();
  stageSenders_.clear();
}

void DistributedDomain::set_methods(Method flags) noexcept {
  if ((flags && Method::ColoQuantityKernel) && (flags && Method::ColoPackMemcpyUnpack)) {
    LOG_FATAL("can't use Direct Access and Pack-Memcpy-Unpack for colocated ranks");
  }
  flags_ = flags;
}

void DistributedDomain::set_output_prefix(const std::string &prefix) { outputPrefix_ = prefix; }

uint64_t DistributedDomain::exchange_bytes_for_method(const Method &method) const {
  uint64_t total = 0;
  if (method && Method::CudaMpi) total += numBytesCudaMpi_;
  if (method && Method::ColoQuantityKernel) total += numBytesColoDirectAccess_;
  if (method && Method::ColoPackMemcpyUnpack) total += numBytesColoPackMemcpyUnpack_;
  if (method && Method::CudaMemcpyPeer) total += numBytesCudaMemcpyPeer_;
  if (method && Method::CudaKernel) total += numBytesCudaKernel_;
  return total;
}

void DistributedDomain::do_placement() {
  nvtxRangePush("placement");
  assert(!placement_);
  switch (strategy_) {
  case PlacementStrategy::NodeAware:
    placement_ = new NodeAware(size_, mpiTopology_, radius_, gpus_);
    break;
  case PlacementStrategy::Trivial:
    placement_ = new Trivial(size_, mpiTopology_, gpus_);
    break;
  case PlacementStrategy::IntraNodeRandom:
    placement_ = new IntraNodeRandom(size_, mpiTopology_, radius_, gpus_);
    break;
  }
  nvtxRangePop();
  topology_ = Topology(placement_->dim(), boundary_);
}

void DistributedDomain::realize() {
  do_placement();

  // LocalDomain copies share raw device pointers: fill the vector while they are still null,
  // then allocate in place.
  domains_.reserve(gpus_.size());
  for (int64_t id = 0; id < int64_t(gpus_.size()); ++id) {
    const Dim3 idx = placement_->get_idx(rank_, int(id));
    LocalDomain sd(placement_->subdomain_size(idx), placement_->subdomain_origin(idx), placement_->get_cuda(idx));
    sd.set_radius(radius_);
    for (size_t q = 0; q < dataElemSize_.size(); ++q) sd.add_data(dataElemSize_[q], dataName_[q]);
    domains_.push_back(sd);
    domainIdx_.push_back(idx);
  }
  for (LocalDomain &d : domains_) d.realize();

  streams_.clear();
  for (const LocalDomain &d : domains_) streams_.push_back(RcStream(d.gpu(), RcStream::Priority::HIGH));

  nvtxRangePush("DistributedDomain::realize() plan");
  plan_exchange();
  nvtxRangePop();
  MPI_Barrier(MPI_COMM_WORLD);
}

// ---- other ranks of the node ------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxDomPerRank = 8, kMaxQuantities = 16, kMaxStagePerRank = kMaxDomPerRank * 26;
struct DomRecord {
  int64_t idx[3], raw[3];
  int nq;
  cudaIpcMemHandle_t curr[kMaxQuantities], next[kMaxQuantities]; // handle of the cudaMalloc block ...
  uint64_t currOff[kMaxQuantities], nextOff[kMaxQuantities];     // ... and the allocation's offset inside it (lead_bytes)
};
struct StageRecord {
  int64_t src[3], dst[3];
  cudaIpcMemHandle_t handle;
};
struct RankRecord {
  int ndom, nstage;
  cudaIpcMemHandle_t flags;
  DomRecord dom[kMaxDomPerRank];
  StageRecord stage[kMaxStagePerRank];
};
} // namespace

// Collective.  Publishes this rank's allocations, staging buffers and mailbox; maps what its neighbours published.
void DistributedDomain::share_with_ranks(const std::map<std::pair<Dim3, Dim3>, char *> &myStage) {
  if (domains_.size() > size_t(kMaxDomPerRank) || dataElemSize_.size() > size_t(kMaxQuantities) || myStage.size() > size_t(kMaxStagePerRank))
    LOG_FATAL("multi-rank exchange supports <= " << kMaxDomPerRank << " subdomains per rank and <= " << kMaxQuantities << " quantities");
  const int dev0 = domains_[0].gpu();
  for (const LocalDomain &d : domains_)
    if (d.gpu() != dev0) LOG_FATAL("with several ranks every rank drives ONE GPU (the IPC mappings are opened on one device)");
  CUDA_RUNTIME(cudaSetDevice(dev0));
  CUDA_RUNTIME(cudaMalloc(&mailbox_, 2 * size_t(worldSize_) * sizeof(uint32_t)));
  CUDA_RUNTIME(cudaMemset(mailbox_, 0, 2 * size_t(worldSize_) * sizeof(uint32_t)));
  CUDA_RUNTIME(cudaDeviceSynchronize());

  std::vector<RankRecord> all;
  all.resize(size_t(worldSize_));
  RankRecord &me = all[size_t(rank_)];
  std::memset(&me, 0, sizeof(me));
  me.ndom = int(domains_.size());
  CUDA_RUNTIME(cudaIpcGetMemHandle(&me.flags, mailbox_));
  for (size_t di = 0; di < domains_.size(); ++di) {
    const LocalDomain &d = domains_[di];
    DomRecord &r = me.dom[di];
    set3(r.idx, domainIdx_[di]);
    set3(r.raw, d.raw_size());
    r.nq = int(d.num_data());
    CUDA_RUNTIME(cudaSetDevice(d.gpu()));
    for (int64_t q = 0; q < d.num_data(); ++q) {
      const size_t lead = d.lead_bytes(size_t(q));
      CUDA_RUNTIME(cudaIpcGetMemHandle(&r.curr[q], static_cast<char *>(d.curr_data(size_t(q)).ptr) - lead));
      CUDA_RUNTIME(cudaIpcGetMemHandle(&r.next[q], static_cast<char *>(d.next_data(size_t(q)).ptr) - lead));
      r.currOff[q] = r.nextOff[q] = lead;
    }
  }
  for (const auto &kv : myStage) {
    StageRecord &sr = me.stage[me.nstage++];
    set3(sr.src, kv.first.first);
    set3(sr.dst, kv.first.second);
    CUDA_RUNTIME(cudaIpcGetMemHandle(&sr.handle, kv.second));
  }
  {
    const RankRecord mine = me; // Allgather's send buffer must not alias the receive buffer
    MPI_Allgather(&mine, int(sizeof(RankRecord)), MPI_BYTE, all.data(), int(sizeof(RankRecord)), MPI_BYTE, MPI_COMM_WORLD);
  }

  // which subdomains of other ranks do mine talk to?
  std::set<Dim3> wanted;
  std::set<int> nbrs;
  for (size_t di = 0; di < domains_.size(); ++di)
    for (int z = -1; z <= 1; ++z)
      for (int y = -1; y <= 1; ++y)
        for (int x = -1; x <= 1; ++x) {
          if (0 == x && 0 == y && 0 == z) continue;
          const Topology::OptionalNeighbor nb = topology_.get_neighbor(domainIdx_[di], Dim3(x, y, z));
          if (!nb.exists) continue;
          const int r = placement_->get_rank(nb.index);
          if (r == rank_) continue;
          wanted.insert(nb.index);
          nbrs.insert(r);
        }
  nbrRanks_.assign(nbrs.begin(), nbrs.end());
  peerFlags_.assign(size_t(worldSize_), nullptr);
  CUDA_RUNTIME(cudaSetDevice(dev0));
  auto open = [&](const cudaIpcMemHandle_t &h) {
    void *p = nullptr;
    CUDA_RUNTIME(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ipcOpened_.push_back(p);
    return static_cast<char *>(p);
  };
  for (int r : nbrRanks_) {
    const RankRecord &rec = all[size_t(r)];
    peerFlags_[size_t(r)] = reinterpret_cast<uint32_t *>(open(rec.flags));
    for (int i = 0; i < rec.ndom; ++i) {
      const DomRecord &dr = rec.dom[i];
      const Dim3 idx(dr.idx[0], dr.idx[1], dr.idx[2]);
      if (!wanted.count(idx)) continue;
      RemoteDomain rd;
      rd.raw = Dim3(dr.raw[0], dr.raw[1], dr.raw[2]);
      for (int q = 0; q < dr.nq; ++q) {
        rd.curr.push_back(open(dr.curr[q]) + dr.currOff[q]);
        rd.next.push_back(open(dr.next[q]) + dr.nextOff[q]);
      }
      remote_[idx] = rd;
    }
    for (int i = 0; i < rec.nstage; ++i) {
      const StageRecord &sr = rec.stage[i];
      const Dim3 src(sr.src[0], sr.src[1], sr.src[2]), dst(sr.dst[0], sr.dst[1], sr.dst[2]);
      if (placement_->get_rank(src) == rank_) remoteStage_[std::make_pair(src, dst)] = open(sr.handle);
    }
  }
}

void DistributedDomain::close_ranks() {
  if (!mailbox_) return;
  for (const LocalDomain &d : domains_) {
    cudaSetDevice(d.gpu());
    cudaDeviceSynchronize();
  }
  MPI_Barrier(MPI_COMM_WORLD); // nobody unmaps while a neighbour could still be writing
  cudaSetDevice(domains_[0].gpu());
  for (void *p : ipcOpened_) cudaIpcCloseMemHandle(p);
  ipcOpened_.clear();
  remote_.clear();
  remoteStage_.clear();
  MPI_Barrier(MPI_COMM_WORLD); // ... and nobody frees what a neighbour still has mapped
  cudaFree(mailbox_);
  mailbox_ = nullptr;
}

// exchange e, before the copies: tell every neighbour rank that my ghost cells may be overwritten, wait until they said so
void DistributedDomain::flags_begin() {
  if (nbrRanks_.empty()) return;
  const int dev = domains_[0].gpu();
  std::vector<uint32_t *> slots;
  for (int r : nbrRanks_) slots.push_back(peerFlags_[size_t(r)] + rank_);
  if (SB_OK != sb_signal(slots.data(), int(slots.size()), epoch_, dev, streams_[0])) LOG_FATAL("exchange: " << sb_last_error());
  for (int r : nbrRanks_)
    if (SB_OK != sb_wait(mailbox_ + r, 1, epoch_, dev, streams_[0])) LOG_FATAL("exchange: " << sb_last_error());
  if (streams_.size() > 1) { // the other local subdomains start after the handshake
    cudaEvent_t ev = phase1Done_[0];
    CUDA_RUNTIME(cudaSetDevice(dev));
    CUDA_RUNTIME(cudaEventRecord(ev, streams_[0]));
    for (size_t di = 1; di < streams_.size(); ++di) {
      CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
      CUDA_RUNTIME(cudaStreamWaitEvent(streams_[di], ev, 0));
    }
  }
}

// after the copies of every local subdomain: tell the neighbours they have landed, wait for theirs
void DistributedDomain::flags_finish() {
  if (nbrRanks_.empty()) return;
  const int dev = domains_[0].gpu();
  for (size_t di = 1; di < streams_.size(); ++di) {
    CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
    CUDA_RUNTIME(cudaEventRecord(phase1Done_[di], streams_[di]));
    CUDA_RUNTIME(cudaSetDevice(dev));
    CUDA_RUNTIME(cudaStreamWaitEvent(streams_[0], phase1Done_[di], 0));
  }
  std::vector<uint32_t *> slots;
  for (int r : nbrRanks_) slots.push_back(peerFlags_[size_t(r)] + worldSize_ + rank_);
  if (SB_OK != sb_signal(slots.data(), int(slots.size()), epoch_, dev, streams_[0])) LOG_FATAL("exchange: " << sb_last_error());
  for (int r : nbrRanks_)
    if (SB_OK != sb_wait(mailbox_ + worldSize_ + r, 1, epoch_, dev, streams_[0])) LOG_FATAL("exchange: " << sb_last_error());
  if (streams_.size() > 1) {
    CUDA_RUNTIME(cudaSetDevice(dev));
    CUDA_RUNTIME(cudaEventRecord(phase1Done_[0], streams_[0]));
    for (size_t di = 1; di < streams_.size(); ++di) {
      CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
      CUDA_RUNTIME(cudaStreamWaitEvent(streams_[di], phase1Done_[0], 0));
    }
  }
}

// For each local subdomain and each of the 26 directions with a non-zero radius on the receiving
// side: source box = outermost compute cells on side `dir`, destination box = ghost cells on side
// -dir of the neighbour, extent = the neighbour's halo extent on side -dir.  The neighbour may be a
// local subdomain (same or peer GPU) or a subdomain of another rank of this node (CUDA IPC mapping).
void DistributedDomain::plan_exchange() {
  destroy_plans();
  numBytesCudaMpi_ = numBytesColoDirectAccess_ = numBytesColoPackMemcpyUnpack_ = numBytesCudaMemcpyPeer_ = numBytesCudaKernel_ = 0;

  std::ofstream planFile(outputPrefix_ + "plan_" + std::to_string(rank_) + ".txt");
  planFile << "rank=" << rank_ << "\n\ndomains\n";
  for (size_t di = 0; di < domains_.size(); ++di)
    planFile << di << ":cuda" << domains_[di].gpu() << ":" << domainIdx_[di] << " sz=" << domains_[di].size() << "\n";
  planFile << "\n== fused direct-write messages ==\n";

  // ---- geometry of every message that starts or ends at a local subdomain ------------------------------
  struct Msg {
    Dim3 srcIdx, dstIdx, dir, srcPos, dstPos, ext;
  };
  auto gpu_key = [&](const Dim3 &idx) { return std::make_pair(placement_->get_rank(idx), placement_->get_cuda(idx)); };
  auto local_id = [&](const Dim3 &idx) { return placement_->get_rank(idx) == rank_ ? placement_->get_subdomain_id(idx) : -1; };
  // all messages src -> dst in one canonical order (direction z, y, x ascending): both ends derive the same list
  auto messages_from = [&](const Dim3 &srcIdx) {
    std::vector<Msg> out;
    const Dim3 srcSz = placement_->subdomain_size(srcIdx);
    for (int z = -1; z <= 1; ++z)
      for (int y = -1; y <= 1; ++y)
        for (int x = -1; x <= 1; ++x) {
          const Dim3 dir(x, y, z);
          if (Dim3(0, 0, 0) == dir) continue;
          // the neighbour on side dir needs our cells only if ITS stencil reaches back (-dir)
          if (0 == radius_.dir(dir * -1)) continue;
          const Topology::OptionalNeighbor nbr = topology_.get_neighbor(srcIdx, dir);
          if (!nbr.exists) continue;
          const Dim3 dstSz = placement_->subdomain_size(nbr.index);
          const Dim3 ext = LocalDomain::halo_extent(dir * -1, dstSz, radius_);
          if (0 == ext.flatten()) continue;
          out.push_back(Msg{srcIdx, nbr.index, dir, LocalDomain::halo_pos(dir, srcSz, radius_, false),
                            LocalDomain::halo_pos(dir * -1, dstSz, radius_, true), ext});
        }
    return out;
  };
  // thin rows (x-faces, x-edges, corners) bound for ANOTHER GPU are packed into one dense staging buffer in the receiver's
  // memory and scattered there: layout of the (src -> dst) buffer, identical on both sides
  constexpr int64_t kStageMaxRowBytes = 64;
  struct Staged {
    Msg m;
    int64_t q, offset;
  };
  auto staging_layout = [&](const Dim3 &srcIdx, const Dim3 &dstIdx, int64_t *total) {
    std::vector<Staged> out;
    int64_t off = 0;
    if (!(gpu_key(srcIdx) == gpu_key(dstIdx))) {
      for (const Msg &m : messages_from(srcIdx)) {
        if (!(m.dstIdx == dstIdx)) continue;
        for (size_t q = 0; q < dataElemSize_.size(); ++q) {
          const int64_t es = int64_t(dataElemSize_[q]);
          if (m.ext.x * es >= kStageMaxRowBytes) continue;
          off = (off + 15) & ~int64_t(15);
          out.push_back(Staged{m, int64_t(q), off});
          off += es * int64_t(m.ext.flatten());
        }
      }
    }
    *total = off;
    return out;
  };

  // ---- staging buffers I receive into (allocated here, published to the senders) ----------------------------
  std::map<std::pair<Dim3, Dim3>, char *> myStage;                  // (src idx, dst idx) -> buffer on dst's GPU
  std::map<std::pair<Dim3, Dim3>, std::vector<Staged>> myStageEntries;
  stageSenders_.assign(domains_.size(), {});
  for (size_t dj = 0; dj < domains_.size(); ++dj) {
    std::set<Dim3> sources;
    for (int z = -1; z <= 1; ++z)
      for (int y = -1; y <= 1; ++y)
        for (int x = -1; x <= 1; ++x) {
          if (0 == x && 0 == y && 0 == z) continue;
          const Topology::OptionalNeighbor nb = topology_.get_neighbor(domainIdx_[dj], Dim3(x, y, z));
          if (nb.exists) sources.insert(nb.index);
        }
    for (const Dim3 &srcIdx : sources) {
      int64_t total = 0;
      std::vector<Staged> entries = staging_layout(srcIdx, domainIdx_[dj], &total);
      if (0 == total) continue;
      void *buf = nullptr;
      CUDA_RUNTIME(cudaSetDevice(domains_[dj].gpu()));
      CUDA_RUNTIME(cudaMalloc(&buf, size_t(total)));
      stagingBufs_.push_back(buf);
      stagingDevs_.push_back(domains_[dj].gpu());
      const auto key = std::make_pair(srcIdx, domainIdx_[dj]);
      myStage[key] = static_cast<char *>(buf);
      myStageEntries[key] = entries;
      if (local_id(srcIdx) >= 0) stageSenders_[dj].push_back(size_t(local_id(srcIdx)));
    }
  }
  for (size_t di = 0; di < domains_.size(); ++di) {
    cudaEvent_t ev;
    CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
    CUDA_RUNTIME(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    phase1Done_.push_back(ev);
  }
  if (worldSize_ > 1) share_with_ranks(myStage);

  // pointer of quantity q of subdomain idx, in the naming of swap parity `parity` (all ranks swap in lock step)
  auto pitched_of = [&](const Dim3 &idx, size_t q, int parity) {
    const int id = local_id(idx);
    if (id >= 0) {
      const LocalDomain &d = domains_[size_t(id)];
      return as_sb(0 == parity ? d.curr_data(q) : d.next_data(q));
    }
    auto it = remote_.find(idx);
    if (it == remote_.end()) LOG_FATAL("subdomain " << idx << " of rank " << placement_->get_rank(idx) << " was not shared");
    const RemoteDomain &rd = it->second;
    return sb_pitched{0 == parity ? rd.curr[q] : rd.next[q], rd.raw.x * int64_t(dataElemSize_[q]), rd.raw.y};
  };

  // ---- the two plans per subdomain and swap parity -----------------------------------------------------
  for (int parity = 0; parity < 2; ++parity) {
    for (size_t di = 0; di < domains_.size(); ++di) {
      const LocalDomain &src = domains_[di];
      std::vector<sb_box_copy> copies;
      for (const Msg &m : messages_from(domainIdx_[di])) {
        const bool sameGpu = gpu_key(m.srcIdx) == gpu_key(m.dstIdx);
        const bool remote = placement_->get_rank(m.dstIdx) != rank_;
        if (!remote && !gpu_topo::peer(src.gpu(), placement_->get_cuda(m.dstIdx)))
          LOG_FATAL("GPU " << src.gpu() << " cannot map GPU " << placement_->get_cuda(m.dstIdx) << " (no P2P): unsupported on this path");
        // where this pair's thin rows go, if anywhere
        const auto key = std::make_pair(m.srcIdx, m.dstIdx);
        char *stageBase = nullptr;
        std::vector<Staged> entries;
        if (!sameGpu) {
          int64_t total = 0;
          entries = staging_layout(m.srcIdx, m.dstIdx, &total);
          if (total > 0) {
            stageBase = remote ? remoteStage_[key] : myStage[key];
            if (!stageBase) LOG_FATAL("no staging buffer for " << m.srcIdx << " -> " << m.dstIdx);
          }
        }
        uint64_t msgBytes = 0;
        for (size_t q = 0; q < dataElemSize_.size(); ++q) {
          sb_box_copy c{};
          c.elem_size = int64_t(dataElemSize_[q]);
          msgBytes += uint64_t(c.elem_size) * m.ext.flatten();
          c.src = pitched_of(m.srcIdx, q, parity);
          set3(c.src_pos, m.srcPos);
          set3(c.extent, m.ext);
          const Staged *st = nullptr;
          for (const Staged &cand : entries)
            if (cand.m.dir == m.dir && cand.q == int64_t(q)) st = &cand;
          if (st) { // dense staging buffer in the receiver's memory
            c.dst = sb_pitched{stageBase + st->offset, m.ext.x * c.elem_size, m.ext.y};
          } else {
            c.dst = pitched_of(m.dstIdx, q, parity);
            set3(c.dst_pos, m.dstPos);
          }
          copies.push_back(c);
        }
        if (0 == parity) { // attribute the bytes the way the reference's planner picks a transport (src/stencil.cu:383-445)
          const char *how = "";
          if (remote) {
            if (any_methods(Method::ColoPackMemcpyUnpack)) numBytesColoPackMemcpyUnpack_ += msgBytes, how = "colo-rank";
            else if (any_methods(Method::ColoQuantityKernel)) numBytesColoDirectAccess_ += msgBytes, how = "colo-rank";
            else if (any_methods(Method::CudaMpi)) numBytesCudaMpi_ += msgBytes, how = "colo-rank(mpi)";
            else LOG_FATAL("No method available to send required message " << m.dir << "\n");
          } else if (any_methods(Method::CudaKernel) && sameGpu) {
            numBytesCudaKernel_ += msgBytes, how = "same-gpu";
          } else if (any_methods(Method::CudaMemcpyPeer)) {
            numBytesCudaMemcpyPeer_ += msgBytes, how = "peer";
          } else if (any_methods(Method::CudaMpi)) {
            numBytesCudaMpi_ += msgBytes, how = "self-mpi";
          } else {
            LOG_FATAL("No method available to send required message " << m.dir << "\n");
          }
          planFile << m.srcIdx << "->" << m.dstIdx << " " << m.dir << " " << msgBytes << "B " << how << (stageBase ? " (thin rows staged)" : "") << "\n";
        }
      }
      sb_copy_plan *plan = nullptr;
      if (SB_OK != sb_copy_plan_create(&plan, src.gpu(), copies.data(), int64_t(copies.size()))) {
        LOG_FATAL("exchange plan: " << sb_last_error());
      }
      plans_[parity].push_back(plan);

      // phase 2 of this subdomain as a receiver
      std::vector<sb_box_copy> scatter;
      for (const auto &kv : myStageEntries) {
        if (!(kv.first.second == domainIdx_[di])) continue;
        for (const Staged &st : kv.second) {
          sb_box_copy c{};
          c.elem_size = int64_t(dataElemSize_[size_t(st.q)]);
          c.src = sb_pitched{myStage[kv.first] + st.offset, st.m.ext.x * c.elem_size, st.m.ext.y};
          c.dst = pitched_of(domainIdx_[di], size_t(st.q), parity);
          set3(c.dst_pos, st.m.dstPos);
          set3(c.extent, st.m.ext);
          scatter.push_back(c);
        }
      }
      sb_copy_plan *up = nullptr;
      if (!scatter.empty() && SB_OK != sb_copy_plan_create(&up, src.gpu(), scatter.data(), int64_t(scatter.size()))) {
        LOG_FATAL("exchange scatter plan: " << sb_last_error());
      }
      unpackPlans_[parity].push_back(up);
    }
  }
  planFile.close();
  parity_ = 0;

  // rank x rank payload bytes per exchange, for numpy.loadtxt (the reference dumps the same file: src/stencil.cu:476-503)
  {
    std::vector<double> mat(size_t(worldSize_) * size_t(worldSize_), 0.0);
    for (size_t di = 0; di < domains_.size(); ++di)
      for (const Msg &m : messages_from(domainIdx_[di])) {
        double bytes = 0;
        for (size_t q = 0; q < dataElemSize_.size(); ++q) bytes += double(dataElemSize_[q]) * double(m.ext.flatten());
        mat[size_t(rank_) * size_t(worldSize_) + size_t(placement_->get_rank(m.dstIdx))] += bytes;
      }
    MPI_Allreduce(MPI_IN_PLACE, mat.data(), int(mat.size()), MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    if (0 == rank_) {
      std::ofstream matFile(outputPrefix_ + "mat_npy_loadtxt.txt");
      for (int r = 0; r < worldSize_; ++r) {
        for (int c = 0; c < worldSize_; ++c) matFile << mat[size_t(r) * size_t(worldSize_) + size_t(c)] << " ";
        matFile << "\n";
      }
    }
  }

  // every rank learns the global volume
  for (uint64_t *b : {&numBytesCudaMpi_, &numBytesColoDirectAccess_, &numBytesColoPackMemcpyUnpack_, &numBytesCudaMemcpyPeer_, &numBytesCudaKernel_})
    MPI_Allreduce(MPI_IN_PLACE, b, 1, MPI_UINT64_T, MPI_SUM, MPI_COMM_WORLD);
}

void DistributedDomain::swap() {
#ifdef STENCIL_EXCHANGE_STATS
  MPI_Barrier(MPI_COMM_WORLD);
  const double start = MPI_Wtime();
#endif
  for (LocalDomain &d : domains_) d.swap();
  parity_ ^= 1; // the exchange plans exist for both identities of "current"
#ifdef STENCIL_EXCHANGE_STATS
  double elapsed = MPI_Wtime() - start, maxElapsed = -1;
  MPI_Reduce(&elapsed, &maxElapsed, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
  if (0 == rank_) timeSwap_ += maxElapsed;
#endif
}

std::vector<Rect3> DistributedDomain::get_interior() const {
  std::vector<Rect3> out;
  for (const LocalDomain &d : domains_) out.push_back(stencil::geom::interior(d.get_compute_region(), radius_));
  return out;
}

std::vector<std::vector<Rect3>> DistributedDomain::get_exterior() const {
  std::vector<std::vector<Rect3>> out;
  for (const LocalDomain &d : domains_) out.push_back(stencil::geom::exterior(d.get_compute_region(), radius_));
  return out;
}

const Rect3 DistributedDomain::get_compute_region() const noexcept { return Rect3(Dim3(0, 0, 0), size_); }

void DistributedDomain::exchange_async() {
  nvtxRangePush("DD::exchange_async");
  const std::vector<sb_copy_plan *> &plans = plans_[parity_];
  const std::vector<sb_copy_plan *> &scatter = unpackPlans_[parity_];
  const bool ranks = !nbrRanks_.empty();
  if (ranks) {
    ++epoch_;
    flags_begin();
  }
  bool anyStaged = false;
  for (size_t di = 0; di < plans.size(); ++di) {
    if (SB_OK != sb_copy_plan_launch(plans[di], streams_[di])) {
      LOG_FATAL("exchange: " << sb_last_error());
    }
    anyStaged = anyStaged || (scatter[di] != nullptr);
  }
  if (ranks) flags_finish(); // every remote sender's phase 1 has landed (and stream order covers the local ones below)
  if (anyStaged) {
    if (!ranks) {
      for (size_t di = 0; di < plans.size(); ++di) {
        CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
        CUDA_RUNTIME(cudaEventRecord(phase1Done_[di], streams_[di]));
      }
    }
    for (size_t di = 0; di < plans.size(); ++di) {
      if (!scatter[di]) continue;
      CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
      if (!ranks) // (with ranks, flags_finish already made every stream wait for every local phase 1)
        for (size_t sj : stageSenders_[di]) CUDA_RUNTIME(cudaStreamWaitEvent(streams_[di], phase1Done_[sj], 0));
      if (SB_OK != sb_copy_plan_launch(scatter[di], streams_[di])) {
        LOG_FATAL("exchange scatter: " << sb_last_error());
      }
    }
  }
  nvtxRangePop();
}

void DistributedDomain::exchange_wait() {
  for (size_t di = 0; di < streams_.size(); ++di) {
    CUDA_RUNTIME(cudaSetDevice(streams_[di].device()));
    CUDA_RUNTIME(cudaStreamSynchronize(streams_[di]));
  }
}

void DistributedDomain::exchange() {
  nvtxRangePush("DD::exchange()");
#ifdef STENCIL_EXCHANGE_STATS
  MPI_Barrier(MPI_COMM_WORLD);
  const double start = MPI_Wtime();
#endif
  exchange_async();
  exchange_wait();
#ifdef STENCIL_EXCHANGE_STATS
  double elapsed = MPI_Wtime() - start, maxElapsed = -1;
  MPI_Reduce(&elapsed, &maxElapsed, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
  if (0 == rank_) timeExchange_ += maxElapsed;
#endif
  nvtxRangePop();
}

void DistributedDomain::write_paraview(const std::string &prefix, bool zeroNaNs) {
  nvtxRangePush("write_paraview");
  for (size_t di = 0; di < domains_.size(); ++di) {
    const LocalDomain &dom = domains_[di];
    const std::string path = prefix + "_" + std::to_string(size_t(rank_) * domains_.size() + di) + ".txt";
    LOG_INFO("write paraview file " << path);
    std::vector<std::vector<unsigned char>> host;
    for (int64_t q = 0; q < dom.num_data(); ++q) host.push_back(dom.interior_to_host(size_t(q)));

    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) {
      LOG_ERROR("unable to open \"" << path << "\" for writing");
      nvtxRangePop();
      return;
    }
    std::fprintf(f, "Z,Y,X");
    for (int64_t q = 0; q < dom.num_data(); ++q) {
      const std::string &nm = dom.dataName_[size_t(q)];
      std::fprintf(f, ",%s", nm.empty() ? ("data" + std::to_string(q)).c_str() : nm.c_str());
    }
    std::fprintf(f, "\n");
    const Dim3 sz = dom.size(), org = dom.origin();
    size_t cell = 0;
    for (int64_t lz = 0; lz < sz.z; ++lz)
      for (int64_t ly = 0; ly < sz.y; ++ly)
        for (int64_t lx = 0; lx < sz.x; ++lx, ++cell) {
          std::fprintf(f, "%ld,%ld,%ld", long(org.z + lz), long(org.y + ly), long(org.x + lx));
          for (int64_t q = 0; q < dom.num_data(); ++q) {
            if (8 == dom.elem_size(size_t(q))) {
              double v = reinterpret_cast<const double *>(host[size_t(q)].data())[cell];
              if (zeroNaNs && std::isnan(v)) v = 0.0;
              std::fprintf(f, ",%.17f", v);
            } else if (4 == dom.elem_size(size_t(q))) {
              float v = reinterpret_cast<const float *>(host[size_t(q)].data())[cell];
              if (zeroNaNs && std::isnan(v)) v = 0.0f;
              std::fprintf(f, ",%.9f", v);
            }
          }
          std::fprintf(f, "\n");
        }
    std::fclose(f);
  }
  nvtxRangePop();
}

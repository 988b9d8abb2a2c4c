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
      strategy_(PlacementStrategy::NodeAware), parity_(0), numBytesCudaMpi_(0), numBytesColoDirectAccess_(0),
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
  topology_ = Topology(placement_->dim(), Topology::Boundary::PERIODIC);
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

// For each local subdomain and each of the 26 directions with a non-zero radius on the receiving
// side: source box = outermost compute cells on side `dir`, destination box = ghost cells on side
// -dir of the neighbour, extent = the neighbour's halo extent on side -dir.
void DistributedDomain::plan_exchange() {
  destroy_plans();
  numBytesCudaMpi_ = numBytesColoDirectAccess_ = numBytesColoPackMemcpyUnpack_ = numBytesCudaMemcpyPeer_ = numBytesCudaKernel_ = 0;

  std::ofstream planFile(outputPrefix_ + "plan_" + std::to_string(rank_) + ".txt");
  planFile << "rank=" << rank_ << "\n\ndomains\n";
  for (size_t di = 0; di < domains_.size(); ++di)
    planFile << di << ":cuda" << domains_[di].gpu() << ":" << domainIdx_[di] << " sz=" << domains_[di].size() << "\n";
  planFile << "\n== fused direct-write messages ==\n";

  // ---- the message list (geometry only) ----------------------------------------------------------
  struct Msg {
    size_t src, dst; // local domain ids
    Dim3 dir, srcPos, dstPos, ext;
  };
  std::vector<Msg> msgs;
  for (size_t di = 0; di < domains_.size(); ++di) {
    const LocalDomain &src = domains_[di];
    for (int z = -1; z <= 1; ++z) {
      for (int y = -1; y <= 1; ++y) {
        for (int x = -1; x <= 1; ++x) {
          const Dim3 dir(x, y, z);
          if (Dim3(0, 0, 0) == dir) continue;
          // the neighbour on side dir needs our cells only if ITS stencil reaches back (-dir)
          if (0 == radius_.dir(dir * -1)) continue;
          const Topology::OptionalNeighbor nbr = topology_.get_neighbor(domainIdx_[di], dir);
          if (!nbr.exists) continue;
          const int dstRank = placement_->get_rank(nbr.index);
          if (dstRank != rank_) {
            LOG_FATAL("subdomain " << nbr.index << " lives on rank " << dstRank
                                   << ": the C++ API drives one rank x N GPUs; use the one-process-per-GPU "
                                      "CUDA-IPC mode of stencil_b200 (python) for multi-process runs");
          }
          const size_t dj = size_t(placement_->get_subdomain_id(nbr.index));
          const LocalDomain &dst = domains_[dj];
          const Dim3 ext = LocalDomain::halo_extent(dir * -1, dst.size(), radius_);
          if (0 == ext.flatten()) continue;

          // attribute the bytes the way the reference's planner picks a transport
          uint64_t *bucket = nullptr;
          const char *how = "";
          if (any_methods(Method::CudaKernel) && src.gpu() == dst.gpu()) {
            bucket = &numBytesCudaKernel_;
            how = "same-gpu";
          } else if (any_methods(Method::CudaMemcpyPeer) && gpu_topo::peer(src.gpu(), dst.gpu())) {
            bucket = &numBytesCudaMemcpyPeer_;
            how = "peer";
          } else if (any_methods(Method::CudaMpi)) {
            bucket = &numBytesCudaMpi_;
            how = "self-mpi";
          } else {
            LOG_FATAL("No method available to send required message " << dir << "\n");
          }
          if (!gpu_topo::peer(src.gpu(), dst.gpu())) {
            LOG_FATAL("GPU " << src.gpu() << " cannot map GPU " << dst.gpu() << " (no P2P): unsupported on this path");
          }
          uint64_t msgBytes = 0;
          for (int64_t q = 0; q < src.num_data(); ++q) msgBytes += uint64_t(src.elem_size(size_t(q))) * ext.flatten();
          *bucket += msgBytes;
          planFile << di << "->" << dj << " " << dir << " " << msgBytes << "B " << how << "\n";
          msgs.push_back(Msg{di, dj, dir, src.halo_pos(dir, false), dst.halo_pos(dir * -1, true), ext});
        }
      }
    }
  }

  // ---- staging buffers: thin rows that cross GPUs ----------------------------------------------------
  constexpr int64_t kStageMaxRowBytes = 64;
  struct Staged {
    size_t msg;
    int64_t q;
    int64_t offset;
  };
  // per (src, dst) pair: buffer + entries
  std::map<std::pair<size_t, size_t>, std::vector<Staged>> stagedOf;
  std::map<std::pair<size_t, size_t>, int64_t> stagedBytes;
  for (size_t mi = 0; mi < msgs.size(); ++mi) {
    const Msg &m = msgs[mi];
    if (domains_[m.src].gpu() == domains_[m.dst].gpu()) continue;
    for (int64_t q = 0; q < domains_[m.src].num_data(); ++q) {
      const int64_t es = int64_t(domains_[m.src].elem_size(size_t(q)));
      if (m.ext.x * es >= kStageMaxRowBytes) continue;
      const auto key = std::make_pair(m.src, m.dst);
      int64_t &off = stagedBytes[key];
      off = (off + 15) & ~int64_t(15);
      stagedOf[key].push_back(Staged{mi, q, off});
      off += es * int64_t(m.ext.flatten());
    }
  }
  std::map<std::pair<size_t, size_t>, char *> stageBuf;
  stageSenders_.assign(domains_.size(), {});
  for (const auto &kv : stagedBytes) {
    const size_t dj = kv.first.second;
    void *buf = nullptr;
    CUDA_RUNTIME(cudaSetDevice(domains_[dj].gpu()));
    CUDA_RUNTIME(cudaMalloc(&buf, size_t(kv.second)));
    stagingBufs_.push_back(buf);
    stagingDevs_.push_back(domains_[dj].gpu());
    stageBuf[kv.first] = static_cast<char *>(buf);
    stageSenders_[dj].push_back(kv.first.first);
  }
  for (size_t di = 0; di < domains_.size(); ++di) {
    cudaEvent_t ev;
    CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
    CUDA_RUNTIME(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    phase1Done_.push_back(ev);
  }

  // ---- the two plans per subdomain and swap parity -----------------------------------------------------
  for (int parity = 0; parity < 2; ++parity) {
    for (size_t di = 0; di < domains_.size(); ++di) {
      const LocalDomain &src = domains_[di];
      std::vector<sb_box_copy> copies;
      for (size_t mi = 0; mi < msgs.size(); ++mi) {
        const Msg &m = msgs[mi];
        if (m.src != di) continue;
        const LocalDomain &dst = domains_[m.dst];
        const auto key = std::make_pair(m.src, m.dst);
        for (int64_t q = 0; q < src.num_data(); ++q) {
          sb_box_copy c{};
          c.elem_size = int64_t(src.elem_size(size_t(q)));
          c.src = as_sb(0 == parity ? src.curr_data(size_t(q)) : src.next_data(size_t(q)));
          set3(c.src_pos, m.srcPos);
          set3(c.extent, m.ext);
          const Staged *st = nullptr;
          auto it = stagedOf.find(key);
          if (it != stagedOf.end())
            for (const Staged &cand : it->second)
              if (cand.msg == mi && cand.q == q) st = &cand;
          if (st) { // dense staging buffer in the receiver's memory
            c.dst = sb_pitched{stageBuf[key] + st->offset, m.ext.x * c.elem_size, m.ext.y};
          } else {
            c.dst = as_sb(0 == parity ? dst.curr_data(size_t(q)) : dst.next_data(size_t(q)));
            set3(c.dst_pos, m.dstPos);
          }
          copies.push_back(c);
        }
      }
      sb_copy_plan *plan = nullptr;
      if (SB_OK != sb_copy_plan_create(&plan, src.gpu(), copies.data(), int64_t(copies.size()))) {
        LOG_FATAL("exchange plan: " << sb_last_error());
      }
      plans_[parity].push_back(plan);

      // phase 2 of this subdomain as a receiver
      std::vector<sb_box_copy> scatter;
      for (const auto &kv : stagedOf) {
        if (kv.first.second != di) continue;
        for (const Staged &st : kv.second) {
          const Msg &m = msgs[st.msg];
          sb_box_copy c{};
          c.elem_size = int64_t(src.elem_size(size_t(st.q)));
          c.src = sb_pitched{stageBuf[kv.first] + st.offset, m.ext.x * c.elem_size, m.ext.y};
          c.dst = as_sb(0 == parity ? src.curr_data(size_t(st.q)) : src.next_data(size_t(st.q)));
          set3(c.dst_pos, m.dstPos);
          set3(c.extent, m.ext);
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

  // every rank learns the global volume (one rank here, kept collective for API parity)
  MPI_Allreduce(MPI_IN_PLACE, &numBytesCudaMpi_, 1, MPI_UINT64_T, MPI_SUM, MPI_COMM_WORLD);
  MPI_Allreduce(MPI_IN_PLACE, &numBytesCudaMemcpyPeer_, 1, MPI_UINT64_T, MPI_SUM, MPI_COMM_WORLD);
  MPI_Allreduce(MPI_IN_PLACE, &numBytesCudaKernel_, 1, MPI_UINT64_T, MPI_SUM, MPI_COMM_WORLD);
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
  bool anyStaged = false;
  for (size_t di = 0; di < plans.size(); ++di) {
    if (SB_OK != sb_copy_plan_launch(plans[di], streams_[di])) {
      LOG_FATAL("exchange: " << sb_last_error());
    }
    anyStaged = anyStaged || (scatter[di] != nullptr);
  }
  if (anyStaged) {
    for (size_t di = 0; di < plans.size(); ++di) {
      CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
      CUDA_RUNTIME(cudaEventRecord(phase1Done_[di], streams_[di]));
    }
    for (size_t di = 0; di < plans.size(); ++di) {
      if (!scatter[di]) continue;
      CUDA_RUNTIME(cudaSetDevice(domains_[di].gpu()));
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

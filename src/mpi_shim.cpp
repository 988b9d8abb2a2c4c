// Single-process implementation of the MPI surface declared in include/mpi_shim/mpi.h.
// One rank, size 1: collectives are copies, point-to-point is a self-send matched by
// (communicator, tag) in posting order.  Buffers may be host or device memory: copies go through
// cudaMemcpy(cudaMemcpyDefault) (UVA) whenever either side is not plain host memory.
#include "mpi.h"

#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include <unistd.h>

namespace {

struct Pending {
  bool isSend;
  const void *sbuf;
  void *rbuf;
  size_t bytes;
  int tag;
  MPI_Comm comm;
  int reqId;
};

struct Req {
  bool live = false;
  bool done = false;
  int tag = 0;
  size_t bytes = 0;
};

std::mutex g_mu;
std::vector<Req> g_reqs(1); // id 0 is MPI_REQUEST_NULL
std::deque<Pending> g_sends, g_recvs;
bool g_init = false;
const int g_tagUb = (1 << 30) - 1;

inline size_t dt_size(MPI_Datatype dt) { return size_t(dt & 0xFF); }

bool is_plain_host(const void *p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError(); // no driver / unregistered memory: treat as host
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost;
}

void copy_any(void *dst, const void *src, size_t n) {
  if (0 == n || dst == src) return;
  if (is_plain_host(dst) && is_plain_host(src)) {
    std::memcpy(dst, src, n);
    return;
  }
  cudaError_t e = cudaMemcpy(dst, src, n, cudaMemcpyDefault);
  if (e != cudaSuccess) {
    std::fprintf(stderr, "mpi_shim: cudaMemcpy failed: %s\n", cudaGetErrorString(e));
    std::exit(-1);
  }
}

int new_req(int tag, size_t bytes) {
  for (size_t i = 1; i < g_reqs.size(); ++i) {
    if (!g_reqs[i].live) {
      g_reqs[i] = Req{true, false, tag, bytes};
      return int(i);
    }
  }
  g_reqs.push_back(Req{true, false, tag, bytes});
  return int(g_reqs.size() - 1);
}

bool tag_match(int want, int have) { return want == MPI_ANY_TAG || want == have; }

void fatal(const char *msg) {
  std::fprintf(stderr, "mpi_shim: %s\n", msg);
  std::exit(-1);
}

void check_peer(int peer) {
  if (peer != 0 && peer != MPI_ANY_SOURCE) fatal("single-process shim: peer rank must be 0");
}

template <typename T> void reduce_copy(const void *s, void *r, int count) {
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, sizeof(T) * size_t(count));
}

} // namespace

extern "C" {

int MPI_Init(int *, char ***) {
  g_init = true;
  return MPI_SUCCESS;
}
int MPI_Init_thread(int *, char ***, int required, int *provided) {
  g_init = true;
  if (provided) *provided = required;
  return MPI_SUCCESS;
}
int MPI_Initialized(int *flag) {
  *flag = g_init ? 1 : 0;
  return MPI_SUCCESS;
}
int MPI_Finalize(void) {
  g_init = false;
  return MPI_SUCCESS;
}
int MPI_Abort(MPI_Comm, int errorcode) { std::exit(errorcode); }

int MPI_Comm_rank(MPI_Comm, int *rank) {
  *rank = 0;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int *size) {
  *size = 1;
  return MPI_SUCCESS;
}
int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm *newcomm) {
  static int next = 16;
  *newcomm = next++;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *comm) {
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_get_attr(MPI_Comm, int keyval, void *attribute_val, int *flag) {
  if (keyval == MPI_TAG_UB) {
    *reinterpret_cast<const int **>(attribute_val) = &g_tagUb;
    *flag = 1;
  } else {
    *flag = 0;
  }
  return MPI_SUCCESS;
}
int MPI_Get_processor_name(char *name, int *resultlen) {
  if (0 != gethostname(name, MPI_MAX_PROCESSOR_NAME - 1)) std::strcpy(name, "localhost");
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *resultlen = int(std::strlen(name));
  return MPI_SUCCESS;
}

double MPI_Wtime(void) {
  using clk = std::chrono::steady_clock;
  static const clk::time_point t0 = clk::now();
  return std::chrono::duration<double>(clk::now() - t0).count();
}
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req) {
  check_peer(dest);
  std::lock_guard<std::mutex> lk(g_mu);
  const size_t bytes = size_t(count) * dt_size(dt);
  const int id = new_req(tag, bytes);
  *req = id;
  for (auto it = g_recvs.begin(); it != g_recvs.end(); ++it) {
    if (it->comm == comm && tag_match(it->tag, tag)) {
      if (bytes > it->bytes) fatal("message truncated");
      copy_any(it->rbuf, buf, bytes);
      g_reqs[it->reqId].done = true;
      g_reqs[it->reqId].tag = tag;
      g_reqs[it->reqId].bytes = bytes;
      g_reqs[id].done = true;
      g_recvs.erase(it);
      return MPI_SUCCESS;
    }
  }
  g_sends.push_back(Pending{true, buf, nullptr, bytes, tag, comm, id});
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req) {
  check_peer(source);
  std::lock_guard<std::mutex> lk(g_mu);
  const size_t bytes = size_t(count) * dt_size(dt);
  const int id = new_req(tag, bytes);
  *req = id;
  for (auto it = g_sends.begin(); it != g_sends.end(); ++it) {
    if (it->comm == comm && tag_match(tag, it->tag)) {
      if (it->bytes > bytes) fatal("message truncated");
      copy_any(buf, it->sbuf, it->bytes);
      g_reqs[it->reqId].done = true;
      g_reqs[id].done = true;
      g_reqs[id].tag = it->tag;
      g_reqs[id].bytes = it->bytes;
      g_sends.erase(it);
      return MPI_SUCCESS;
    }
  }
  g_recvs.push_back(Pending{false, nullptr, buf, bytes, tag, comm, id});
  return MPI_SUCCESS;
}

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (*req == MPI_REQUEST_NULL) {
    *flag = 1;
    return MPI_SUCCESS;
  }
  Req &r = g_reqs[*req];
  if (r.done) {
    if (status) {
      status->MPI_SOURCE = 0;
      status->MPI_TAG = r.tag;
      status->MPI_ERROR = MPI_SUCCESS;
      status->count_bytes_ = int(r.bytes);
    }
    r.live = false;
    *req = MPI_REQUEST_NULL;
    *flag = 1;
  } else {
    *flag = 0;
  }
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status) {
  int flag = 0;
  MPI_Test(req, &flag, status);
  if (!flag) fatal("MPI_Wait on an unmatched request would deadlock in a one-rank world");
  return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *statuses) {
  for (int i = 0; i < n; ++i) MPI_Wait(&reqs[i], statuses ? &statuses[i] : MPI_STATUS_IGNORE);
  return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
  MPI_Request r;
  MPI_Isend(buf, count, dt, dest, tag, comm, &r);
  int flag = 0;
  MPI_Test(&r, &flag, MPI_STATUS_IGNORE);
  // an unmatched blocking self-send completes when the matching receive is posted (eager semantics):
  // leave it queued; the request slot is reclaimed at match time.
  return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status) {
  MPI_Request r;
  MPI_Irecv(buf, count, dt, source, tag, comm, &r);
  return MPI_Wait(&r, status);
}

int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count) {
  *count = int(size_t(status->count_bytes_) / dt_size(dt));
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op, int, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf && recvbuf) copy_any(recvbuf, sendbuf, size_t(count) * dt_size(dt));
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) copy_any(recvbuf, sendbuf, size_t(count) * dt_size(dt));
  return MPI_SUCCESS;
}
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sdt, void *recvbuf, int, MPI_Datatype, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE) copy_any(recvbuf, sendbuf, size_t(sendcount) * dt_size(sdt));
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sdt, void *recvbuf, const int *, const int *displs,
                   MPI_Datatype rdt, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE)
    copy_any(static_cast<char *>(recvbuf) + size_t(displs ? displs[0] : 0) * dt_size(rdt), sendbuf,
             size_t(sendcount) * dt_size(sdt));
  return MPI_SUCCESS;
}
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sdt, void *recvbuf, int, MPI_Datatype, int, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE) copy_any(recvbuf, sendbuf, size_t(sendcount) * dt_size(sdt));
  return MPI_SUCCESS;
}
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Alltoallv(const void *sendbuf, const int *sendcounts, const int *sdispls, MPI_Datatype sdt, void *recvbuf,
                  const int *, const int *rdispls, MPI_Datatype rdt, MPI_Comm) {
  copy_any(static_cast<char *>(recvbuf) + size_t(rdispls[0]) * dt_size(rdt),
           static_cast<const char *>(sendbuf) + size_t(sdispls[0]) * dt_size(sdt), size_t(sendcounts[0]) * dt_size(sdt));
  return MPI_SUCCESS;
}

} // extern "C"

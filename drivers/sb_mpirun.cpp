// sb_mpirun: launcher for the node-local MPI shim (src/mpi_shim.cpp, lib/libmpi_shim.a).
//
//   sb_mpirun -n <ranks> [--] <program> [args...]
//
// Creates the job's POSIX shared-memory segment, starts <ranks> copies of the program with SB_MPI_RANK / SB_MPI_SIZE /
// SB_MPI_JOB set, waits for them, and tears the job down if any rank fails.  Stands in for `mpirun` on boxes without an
// MPI installation so that the reference's drivers can run one rank per GPU against this library
// (the rank -> GPU mapping is the library's: colocated rank i drives GPU i, src/stencil.cu).
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

extern "C" size_t sb_mpi_shim_segment_bytes(int size);
extern "C" void sb_mpi_shim_segment_init(void *base, int size);
extern "C" void sb_mpi_shim_segment_abort(void *base);

int main(int argc, char **argv) {
  int n = 0, first = 1;
  while (first < argc) {
    const std::string a = argv[first];
    if ((a == "-n" || a == "-np") && first + 1 < argc) {
      n = std::atoi(argv[first + 1]);
      first += 2;
    } else if (a == "--") {
      ++first;
      break;
    } else {
      break;
    }
  }
  if (n < 1 || n > 64 || first >= argc) {
    std::fprintf(stderr, "usage: %s -n <ranks (1..64)> [--] <program> [args...]\n", argv[0]);
    return 2;
  }
  const std::string job = "/sbmpi_" + std::to_string(long(getpid()));
  const size_t bytes = sb_mpi_shim_segment_bytes(n);
  const int fd = shm_open(job.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, off_t(bytes)) != 0) {
    std::perror("sb_mpirun: shared-memory segment");
    return 1;
  }
  void *base = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == MAP_FAILED) {
    std::perror("sb_mpirun: mmap");
    shm_unlink(job.c_str());
    return 1;
  }
  sb_mpi_shim_segment_init(base, n); // ftruncate zero-filled the rest

  std::vector<pid_t> kids;
  for (int r = 0; r < n; ++r) {
    const pid_t pid = fork();
    if (pid < 0) {
      std::perror("sb_mpirun: fork");
      break;
    }
    if (0 == pid) {
      setenv("SB_MPI_RANK", std::to_string(r).c_str(), 1);
      setenv("SB_MPI_SIZE", std::to_string(n).c_str(), 1);
      setenv("SB_MPI_JOB", job.c_str(), 1);
      execvp(argv[first], argv + first);
      std::fprintf(stderr, "sb_mpirun: cannot exec %s: %s\n", argv[first], std::strerror(errno));
      _exit(127);
    }
    kids.push_back(pid);
  }
  int rc = (int(kids.size()) == n) ? 0 : 1;
  size_t left = kids.size();
  if (rc) {
    sb_mpi_shim_segment_abort(base);
    for (pid_t p : kids) kill(p, SIGTERM);
  }
  while (left > 0) {
    int st = 0;
    const pid_t p = wait(&st);
    if (p < 0) {
      if (errno == EINTR) continue;
      break;
    }
    --left;
    const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
    if (code != 0 && 0 == rc) { // first failure: release everyone spinning in the shim, then make sure they go away
      rc = code;
      sb_mpi_shim_segment_abort(base);
      for (pid_t k : kids)
        if (k != p) kill(k, SIGTERM);
    }
  }
  munmap(base, bytes);
  shm_unlink(job.c_str());
  return rc;
}

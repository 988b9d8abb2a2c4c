#pragma once
// Timer: a stopwatch that accumulates over resume()/pause() pairs.  timers::cudaRuntime and timers::mpi are the two
// process-wide instances the reference drivers print at exit (bin/jacobi3d.cu:397-398).

#include <chrono>

class Timer {
public:
  Timer() = default;

  void resume() {
    if (running_) return;
    since_ = clock::now();
    running_ = true;
  }

  void pause() {
    if (!running_) return;
    accumulated_ += clock::now() - since_;
    running_ = false;
  }

  void clear();         // stop and forget everything
  double get_elapsed(); // seconds so far; leaves the timer paused

private:
  using clock = std::chrono::steady_clock;
  std::chrono::duration<double> accumulated_{0.0};
  clock::time_point since_{};
  bool running_{false};
};

namespace timers {
extern Timer cudaRuntime;
extern Timer mpi;
} // namespace timers

// Statement-style hooks kept for source compatibility (src/ of the reference brackets API calls with them); the
// accounting itself lives in rt.hpp (STENCIL_TIME_API_CALLS).
#define CR_TIC() ((void)0)
#define CR_TOC() ((void)0)
#define MPI_TIC() ((void)0)
#define MPI_TOC() ((void)0)

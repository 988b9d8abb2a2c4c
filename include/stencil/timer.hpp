#pragma once
// Accumulating wall-clock timer + the two global timers the drivers print at exit.

#include <chrono>

class Timer {
  using Clock = std::chrono::steady_clock;
  Clock::time_point start_{};
  std::chrono::duration<double> total_{0};
  bool running_ = false;

public:
  Timer() = default;

  void clear(); // stop and zero

  void pause() {
    if (running_) {
      total_ += Clock::now() - start_;
      running_ = false;
    }
  }
  void resume() {
    if (!running_) {
      running_ = true;
      start_ = Clock::now();
    }
  }
  double get_elapsed(); // seconds; pauses the timer
};

namespace timers {
extern Timer cudaRuntime;
extern Timer mpi;
} // namespace timers

// per-call timing of CUDA / MPI API calls is compiled out (as in the reference's Release build)
#define CR_TIC()
#define CR_TOC()
#define MPI_TIC()
#define MPI_TOC()

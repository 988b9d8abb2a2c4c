#include "stencil/timer.hpp"

namespace timers {
Timer cudaRuntime; // time inside CUDA runtime calls (when STENCIL_TIME_API_CALLS is defined)
Timer mpi;         // time inside MPI calls
} // namespace timers

void Timer::clear() {
  running_ = false;
  accumulated_ = std::chrono::duration<double>(0.0);
}

double Timer::get_elapsed() {
  pause();
  return accumulated_.count();
}

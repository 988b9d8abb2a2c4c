#include "stencil/timer.hpp"

double Timer::get_elapsed() {
  pause();
  return total_.count();
}

void Timer::clear() {
  pause();
  total_ = std::chrono::duration<double>(0);
}

namespace timers {
Timer cudaRuntime;
Timer mpi;
} // namespace timers

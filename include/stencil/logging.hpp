#pragma once
// LOG_SPEW / LOG_DEBUG / LOG_INFO / LOG_WARN / LOG_ERROR / LOG_FATAL(stream expression)
// One line per message on stderr: TAG[file:line]{rank} text.  STENCIL_OUTPUT_LEVEL selects at compile time how chatty
// the build is (5: everything ... 0: fatal only; default 3).  A message is composed in a local buffer and written with
// a single insertion so lines of different ranks do not interleave mid-line.  LOG_FATAL terminates with status 1.

#include <cstdlib>
#include <iostream>
#include <sstream>

#include "stencil/mpi.hpp"

#ifndef STENCIL_OUTPUT_LEVEL
#define STENCIL_OUTPUT_LEVEL 3
#endif

namespace stencil {
namespace detail {
inline void emit_log(const char *tag, const char *file, int line, const std::string &text) {
  std::ostringstream msg;
  msg << tag << '[' << file << ':' << line << "]{" << mpi::world_rank() << "} " << text << '\n';
  std::cerr << msg.str();
}
} // namespace detail
} // namespace stencil

// `expr` is a chain of << operands, e.g. LOG_INFO("n=" << n)
#define STENCIL_LOG_AT(level, tag, expr)                                                                               \
  do {                                                                                                                 \
    if ((level) <= STENCIL_OUTPUT_LEVEL) {                                                                             \
      std::ostringstream stencil_log_buf_;                                                                             \
      stencil_log_buf_ << expr;                                                                                        \
      stencil::detail::emit_log(tag, __FILE__, __LINE__, stencil_log_buf_.str());                                      \
    }                                                                                                                  \
  } while (0)

#define LOG_SPEW(expr) STENCIL_LOG_AT(5, "SPEW", expr)
#define LOG_DEBUG(expr) STENCIL_LOG_AT(4, "DEBUG", expr)
#define LOG_INFO(expr) STENCIL_LOG_AT(3, "INFO", expr)
#define LOG_WARN(expr) STENCIL_LOG_AT(2, "WARN", expr)
#define LOG_ERROR(expr) STENCIL_LOG_AT(1, "ERROR", expr)
#define LOG_FATAL(expr)                                                                                                \
  do {                                                                                                                 \
    STENCIL_LOG_AT(0, "FATAL", expr);                                                                                  \
    std::exit(1);                                                                                                      \
  } while (0)

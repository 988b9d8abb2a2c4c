#pragma once

#include <cstdint>
#include <cstdlib>

// smallest value >= x that is a multiple of a (a a power of two)
inline size_t __device__ __host__ next_align_of(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }

#pragma once
// Quadratic assignment: place subdomains (weights w = halo bytes between subdomain pairs) on GPUs
// (distances d = 1 / bandwidth between GPU pairs) minimising sum_ab w[a][b] * d[f[a]][f[b]].
// solve(): exhaustive over permutations with a 10 s budget (8! = 40320 on an 8-GPU node);
// solve_catch(): greedy best-pairwise-swap descent.  On an NVSwitch node every off-diagonal
// distance is equal, the cost is permutation invariant and both return the identity.

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <iostream>
#include <vector>

#include "stencil/logging.hpp"
#include "stencil/mat2d.hpp"

namespace qap {

namespace detail {

// 0 * inf counts as 0: no traffic over a missing link costs nothing
inline double cost_product(double we, double de) { return (0 == we || 0 == de) ? 0.0 : we * de; }

inline double cost(const Mat2D<double> &w, const Mat2D<double> &d, const std::vector<size_t> &f) {
  assert(w.shape().x == w.shape().y && d.shape() == w.shape() && w.shape().x == f.size());
  double total = 0;
  const size_t n = f.size();
  for (size_t a = 0; a < n; ++a)
    for (size_t b = 0; b < n; ++b) total += cost_product(w.at(a, b), d.at(f[a], f[b]));
  return total;
}

} // namespace detail

inline std::vector<size_t> solve(const Mat2D<double> &w, Mat2D<double> &d, double *costp = nullptr) {
  using Clock = std::chrono::steady_clock;
  const auto deadline = Clock::now() + std::chrono::seconds(10);
  assert(w.shape() == d.shape() && w.shape().x == w.shape().y);

  std::vector<size_t> f(w.shape().x);
  for (size_t i = 0; i < f.size(); ++i) f[i] = i;
  std::vector<size_t> best = f;
  double bestCost = detail::cost(w, d, f);
  while (std::next_permutation(f.begin(), f.end())) {
    if (Clock::now() > deadline) {
      LOG_WARN("qap::solve timed out");
      break;
    }
    const double c = detail::cost(w, d, f);
    if (c < bestCost) {
      bestCost = c;
      best = f;
    }
  }
  if (costp) *costp = bestCost;
  return best;
}

inline std::vector<size_t> solve_catch(const Mat2D<double> &w, Mat2D<double> &d, double *costp = nullptr) {
  assert(w.shape() == d.shape() && w.shape().x == w.shape().y);
  const size_t n = w.shape().x;
  std::vector<size_t> best(n);
  for (size_t i = 0; i < n; ++i) best[i] = i;
  double bestCost = detail::cost(w, d, best);

  // contribution of rows/columns i and j to the cost under assignment f
  auto touching = [&](const std::vector<size_t> &f, size_t i, size_t j) {
    double c = 0;
    for (size_t k = 0; k < n; ++k) {
      c += detail::cost_product(w.at(i, k), d.at(f[i], f[k]));
      c += detail::cost_product(w.at(j, k), d.at(f[j], f[k]));
      if (k != i && k != j) {
        c += detail::cost_product(w.at(k, i), d.at(f[k], f[i]));
        c += detail::cost_product(w.at(k, j), d.at(f[k], f[j]));
      }
    }
    return c;
  };

  for (bool improved = true; improved;) {
    improved = false;
    std::vector<size_t> roundBest = best;
    double roundCost = bestCost;
    for (size_t i = 0; i < n; ++i) {
      for (size_t j = i + 1; j < n; ++j) {
        std::vector<size_t> f = best;
        double c = bestCost - touching(f, i, j);
        std::swap(f[i], f[j]);
        c += touching(f, i, j);
        if (c < roundCost) {
          roundCost = c;
          roundBest = f;
          improved = true;
        }
      }
    }
    if (improved) {
      best = roundBest;
      bestCost = roundCost;
    }
  }
  if (costp) *costp = bestCost;
  return best;
}

} // namespace qap

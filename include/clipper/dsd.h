/**
 * @file dsd.h
 * @brief Exact densest edge-weighted subgraph, Goldberg's algorithm (mirror of reference dsd.h:20-55)
 *        Host-side; used by Rounding::DSD on the support of the solver's u.
 */
#pragma once
#include <vector>
#include "clipper/types.h"

namespace clipper {
namespace dsd {
  /// A: symmetric, upper triangle filled in; S restricts the search to a subgraph (empty: all)
  std::vector<int> solve(const SpAffinity& A, const std::vector<int>& S = {});
  std::vector<int> solve(const Eigen::MatrixXd& A, const std::vector<int>& S = {});
} // ns dsd
} // ns clipper

/**
 * @file maxclique.h
 * @brief Maximum-clique front-end (mirror of reference maxclique.h:14-28).  The reference
 *        delegates to the third-party PMC library (CMakeLists.txt:75-78); not part of the GPU hot
 *        path and not bundled: solve() behaves like a build without CLIPPER_HAS_PMC
 *        (maxclique.cpp:141-144).
 */
#pragma once
#include <cstddef>
#include <vector>
#include "clipper/types.h"

namespace clipper {
namespace maxclique {

enum class Method { EXACT, HEU, KCORE };

struct Params
{
  Method method = Method::EXACT;
  size_t threads = 24;
  int time_limit = 3600;
  bool verbose = false;
};

std::vector<int> solve(const Eigen::MatrixXd& A, const Params& params = {});

} // ns maxclique
} // ns clipper

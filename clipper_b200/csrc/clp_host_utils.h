// clp_host_utils.h -- host-side utilities of the C-ABI (no CUDA).
#pragma once
#include <stdint.h>
#include <vector>

namespace clp {

// Exact densest edge-weighted subgraph (Goldberg's parametric min-cut), restricted to the k
// nodes whose pairwise weights are W (dense column-major k x k, symmetric, diagonal ignored).
// n_total is the node count the reference's stopping rule uses (rows of the full matrix,
// ref dsd.cpp:198,285).  Returns local indices (ascending) of the selected nodes.
std::vector<int32_t> densest_subgraph_dense(const double* W, int32_t k, int64_t n_total);

}  // namespace clp

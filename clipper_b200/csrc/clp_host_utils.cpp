// clp_host_utils.cpp -- host-side utilities exported through the C-ABI.
//
//   clp_k2ij, clp_create_all_to_all   ref src/utils.cpp:87-97, include/clipper/utils.h:61-71
//   clp_find_k_largest / clp_find_above  ref src/utils.cpp:33-68  (the rounding step, K6)
//   clp_dsd_dense                     ref src/dsd.cpp:274-327 (Goldberg densest subgraph)
//
// These are O(m) / O(m log k) host steps on vectors the API returns to the host anyway; the
// O(m^2) work lives in clp_kernels.cuh.
#include "clp_host_utils.h"
#include "../../include/clipper_b200.h"

#include <algorithm>
#include <cmath>
#include <queue>
#include <utility>

namespace clp {

namespace {

// Dinic max-flow on a dense residual-capacity matrix (V = k + 2 nodes).
struct DenseFlow {
  int V;
  std::vector<double> R;  // residual capacities, row-major V x V
  std::vector<int> level, arc;
  explicit DenseFlow(int v) : V(v), R((size_t)v * v, 0.0), level(v), arc(v) {}
  double& cap(int a, int b) { return R[(size_t)a * V + b]; }

  bool bfs(int s, int t) {
    std::fill(level.begin(), level.end(), -1);
    std::vector<int> q;
    q.reserve(V);
    q.push_back(s);
    level[s] = 0;
    for (size_t head = 0; head < q.size(); ++head) {
      const int a = q[head];
      const double* row = &R[(size_t)a * V];
      for (int b = 0; b < V; ++b)
        if (level[b] < 0 && row[b] > 0.0) { level[b] = level[a] + 1; q.push_back(b); }
    }
    return level[t] >= 0;
  }

  // iterative blocking flow with current-arc pointers
  double augment(int s, int t) {
    double total = 0.0;
    std::vector<int> path;
    for (;;) {
      path.assign(1, s);
      bool found = false;
      while (!path.empty()) {
        const int a = path.back();
        if (a == t) { found = true; break; }
        bool advanced = false;
        for (int& b = arc[a]; b < V; ++b) {
          if (level[b] == level[a] + 1 && R[(size_t)a * V + b] > 0.0) { path.push_back(b); advanced = true; break; }
        }
        if (!advanced) {  // dead end: retreat and never try this node again in this phase
          level[a] = -2;
          path.pop_back();
        }
      }
      if (!found) break;
      double f = INFINITY;
      for (size_t q = 0; q + 1 < path.size(); ++q) f = std::min(f, R[(size_t)path[q] * V + path[q + 1]]);
      for (size_t q = 0; q + 1 < path.size(); ++q) {
        R[(size_t)path[q] * V + path[q + 1]] -= f;
        R[(size_t)path[q + 1] * V + path[q]] += f;
      }
      total += f;
    }
    return total;
  }

  double run(int s, int t) {
    double flow = 0.0;
    while (bfs(s, t)) {
      std::fill(arc.begin(), arc.end(), 0);
      flow += augment(s, t);
    }
    return flow;
  }

  // source side of the minimum cut = nodes reachable from s in the residual graph
  std::vector<char> source_side(int s) {
    std::vector<char> seen((size_t)V, 0);
    std::vector<int> q(1, s);
    seen[(size_t)s] = 1;
    for (size_t head = 0; head < q.size(); ++head) {
      const int a = q[head];
      for (int b = 0; b < V; ++b)
        if (!seen[(size_t)b] && R[(size_t)a * V + b] > 0.0) { seen[(size_t)b] = 1; q.push_back(b); }
    }
    return seen;
  }
};

}  // namespace

std::vector<int32_t> densest_subgraph_dense(const double* W, int32_t k, int64_t n_total) {
  std::vector<int32_t> best;
  if (k <= 0) return best;
  // ref dsd.cpp:285: the edge list holds every ordered pair of S, so m = |S|^2 - |S|
  const int64_t n_edges = (int64_t)k * k - k;
  const double half = (double)(n_edges / 2);  // integer division, as in dsd.cpp:25,33,196
  std::vector<double> degree((size_t)k, 0.0);
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b)
      if (a != b) degree[(size_t)a] += W[(size_t)b * k + a];
  double lo = 0.0, hi = half;
  const int s = k, t = k + 1;
  const double nn = (double)n_total * (double)(n_total - 1);
  while (nn * (hi - lo) >= 1.0) {  // ref dsd.cpp:216
    const double g = (hi + lo) / 2;
    DenseFlow net(k + 2);
    for (int a = 0; a < k; ++a) {
      net.cap(s, a) = half;
      net.cap(a, t) = half + 2 * g - degree[(size_t)a];
      for (int b = 0; b < k; ++b)
        if (a != b) net.cap(a, b) = W[(size_t)b * k + a];
    }
    net.run(s, t);
    const std::vector<char> side = net.source_side(s);
    int cnt = 0;
    for (int a = 0; a < k; ++a) cnt += side[(size_t)a];
    if (cnt == 0) {
      hi = g;
    } else {
      lo = g;
      best.clear();
      for (int a = 0; a < k; ++a)
        if (side[(size_t)a]) best.push_back(a);
    }
  }
  return best;
}

}  // namespace clp

extern "C" {

void clp_k2ij(uint64_t k, uint64_t n, uint64_t* i_out, uint64_t* j_out) {
  // closed-form inverse of the row-major enumeration of the strict upper triangle
  const uint64_t rem = n * (n - 1) / 2 - (k + 1);  // pairs that come after pair k
  const uint64_t tri = (uint64_t)std::floor((std::sqrt((double)(1 + 8 * rem)) - 1) / 2.);
  const uint64_t off = rem - tri * (tri + 1) / 2;
  if (i_out) *i_out = n - (tri + 1) - 1;
  if (j_out) *j_out = n - off - 1;
}

void clp_create_all_to_all(int64_t n1, int64_t n2, int32_t* A) {
  const int64_t m = n1 * n2;
  for (int64_t r = 0; r < m; ++r) {
    A[r] = (int32_t)(r / n2);
    A[m + r] = (int32_t)(r % n2);
  }
}

int32_t clp_find_above(const double* x, int64_t n, double thr, int32_t* out) {
  int32_t c = 0;
  for (int64_t i = 0; i < n; ++i)
    if (x[i] > thr) out[c++] = (int32_t)i;
  return c;
}

// Semantics of the reference's bounded min-heap (utils.cpp:33-55): keep k (value,index) pairs;
// a later element enters only if its value is strictly greater than the smallest kept value and
// then evicts the lexicographically smallest (value,index) pair; output descending.
int32_t clp_find_k_largest(const double* x, int64_t n, int32_t k, int32_t* out) {
  if (k < 1 || n < 1) return 0;
  if ((int64_t)k > n) k = (int32_t)n;  // the reference pops an empty heap here (UB); clamp instead
  typedef std::pair<double, int32_t> VI;
  std::vector<VI> heap;
  heap.reserve((size_t)k);
  auto worse_on_top = [](const VI& a, const VI& b) { return a > b; };  // min-heap on (value,index)
  for (int64_t i = 0; i < n; ++i) {
    if ((int64_t)heap.size() < k) {
      heap.emplace_back(x[i], (int32_t)i);
      std::push_heap(heap.begin(), heap.end(), worse_on_top);
    } else if (heap.front().first < x[i]) {
      std::pop_heap(heap.begin(), heap.end(), worse_on_top);
      heap.back() = VI(x[i], (int32_t)i);
      std::push_heap(heap.begin(), heap.end(), worse_on_top);
    }
  }
  std::sort(heap.begin(), heap.end(), [](const VI& a, const VI& b) { return a > b; });
  for (int32_t e = 0; e < k; ++e) out[e] = heap[(size_t)e].second;
  return k;
}

int32_t clp_dsd_dense(const double* A, int64_t n, const int32_t* S, int32_t nS, int32_t* out) {
  if (!A || n <= 0 || !out) return -1;
  std::vector<int32_t> sel;
  if (S && nS > 0) sel.assign(S, S + nS);
  else { sel.resize((size_t)n); for (int64_t i = 0; i < n; ++i) sel[(size_t)i] = (int32_t)i; }
  const int32_t k = (int32_t)sel.size();
  if (k > 8192) return -2;
  std::vector<double> W((size_t)k * k, 0.0);
  for (int b = 0; b < k; ++b)
    for (int a = 0; a < k; ++a) {
      if (a == b) continue;
      // "A is assumed symmetric and the upper triangle is filled in" (ref dsd.cpp:300-302)
      const int64_t i = sel[(size_t)a], j = sel[(size_t)b];
      W[(size_t)b * k + a] = (i < j) ? A[(size_t)j * n + i] : A[(size_t)i * n + j];
    }
  const std::vector<int32_t> loc = clp::densest_subgraph_dense(W.data(), k, n);
  std::vector<int32_t> nodes;
  for (int32_t a : loc) nodes.push_back(sel[(size_t)a]);
  std::sort(nodes.begin(), nodes.end());
  for (size_t q = 0; q < nodes.size(); ++q) out[q] = nodes[q];
  return (int32_t)nodes.size();
}

}  // extern "C"

// shell_tests.cpp -- the reference's C++ tests restated against the B200 C++ shell
// (include/clipper/*.h + libclipper.so).  Same call sequences and assertions as
//   reference test/affinity_test.cpp:14-108   (Affinity.EuclideanDistance)
//   reference test/clipper_test.cpp:15-207    (CLIPPER.EuclideanDistance, _UseGetSet, _UseSparseGetSet)
//   reference test/sdp_test.cpp:15-64         (CLIPPERSDR.FindGlobalMax)
//   reference test/dsd_test.cpp:14-80         (DSD.Solve, DSD.SolveRestrictedGraph)
// written with a 20-line assertion harness because gtest and Eigen/Geometry are not installed.
// `shell_tests --cpu-only` runs the tests that need no GPU (DSD, utils).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <clipper/clipper.h>
#include <clipper/utils.h>

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::printf("  FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define ASSERT_EQ(a, b) do { ++g_checks; if (!((a) == (b))) { ++g_fail; std::printf("  FAIL %s:%d  %s == %s\n", __FILE__, __LINE__, #a, #b); return; } } while (0)

// model: 4 points; data = T_MD^-1 * model with T_MD = (Rz(pi/8), t = (5,3,0)), first 3 points
static void toy(Eigen::MatrixXd& model, Eigen::MatrixXd& data) {
  model = Eigen::MatrixXd(3, 4);
  const double pts[4][3] = {{0, 0, 0}, {2, 0, 0}, {0, 3, 0}, {2, 2, 0}};
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 3; ++i) model(i, j) = pts[j][i];
  const double th = M_PI / 8, c = std::cos(th), s = std::sin(th), t[3] = {5, 3, 0};
  data = Eigen::MatrixXd(3, 3);
  for (int j = 0; j < 3; ++j) {
    const double x = model(0, j) - t[0], y = model(1, j) - t[1], z = model(2, j) - t[2];
    data(0, j) = c * x + s * y;   // R^T (p - t)
    data(1, j) = -s * x + c * y;
    data(2, j) = z;
  }
}

static clipper::invariants::EuclideanDistancePtr make_invariant() {
  clipper::invariants::EuclideanDistance::Params iparams;
  return std::make_shared<clipper::invariants::EuclideanDistance>(iparams);
}

static Eigen::MatrixXd m20() {
  Eigen::MatrixXd M = Eigen::MatrixXd::Identity(20, 20);
  const struct { int i, j; double v; } e[] = {
    {0,18,0.2964},{1,13,0.0138},{2,11,0.0016},{2,18,0.0747},{3,5,0.0555},{3,6,0.2547},{3,13,0.0102},{3,15,0.7715},
    {4,5,0.0063},{4,7,0.3846},{4,9,0.0003},{4,10,0.0014},{4,15,0.0063},{5,12,0.9927},{5,15,0.9722},{6,8,0.0023},
    {6,11,0.8775},{7,8,0.0001},{8,9,0.7914},{8,13,0.0617},{8,16,0.9938},{8,19,0.0007},{9,12,0.0001},{9,13,0.0091},
    {9,15,0.2503},{9,16,0.0222},{9,17,0.0549},{10,19,0.0008},{11,18,0.7007},{12,14,0.9978},{13,17,0.0003},
    {14,15,0.0012},{14,19,0.0074},{15,16,0.0026},{15,17,0.0217},{17,18,0.0007}};
  for (const auto& x : e) { M(x.i, x.j) = x.v; M(x.j, x.i) = x.v; }
  return M;
}

// reference test/affinity_test.cpp:14-108
static void Affinity_EuclideanDistance() {
  clipper::Params params;
  clipper::CLIPPER clipper(make_invariant(), params);
  Eigen::MatrixXd model, data;
  toy(model, data);
  clipper.scorePairwiseConsistency(model, data);
  clipper::Association A = clipper.getInitialAssociations();
  const int n = (int)(model.cols() * data.cols());
  EXPECT_EQ(A.rows(), n);
  EXPECT_EQ(A.cols(), 2);
  for (size_t i = 0; i < (size_t)model.cols(); i++)
    for (size_t j = 0; j < (size_t)data.cols(); j++) {
      const size_t k = i * data.cols() + j;
      EXPECT_EQ(A(k, 0), (int)i);
      EXPECT_EQ(A(k, 1), (int)j);
    }
  clipper::Affinity M = clipper.getAffinityMatrix();
  clipper::Constraint C = clipper.getConstraintMatrix();
  EXPECT_EQ(M.rows(), A.rows());
  EXPECT_EQ(M.cols(), A.rows());
  EXPECT_EQ(M.diagonal(), Eigen::VectorXd::Ones(M.rows()));
  EXPECT_EQ(M, M.transpose());
  EXPECT_EQ(C, C.transpose());
  EXPECT_EQ(M, C);
  Eigen::MatrixXd Mtrue(M.rows(), M.cols());
  Mtrue << 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0,
           0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0,
           0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0,
           0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0,
           1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0,
           0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0,
           0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0,
           0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0,
           1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0,
           0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0,
           0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0,
           0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1;
  EXPECT_EQ(M, Mtrue);
}

// reference test/clipper_test.cpp:15-68 (u0 fixed: a random start reaches a 2-clique ~6% of the time)
static void CLIPPER_EuclideanDistance() {
  clipper::Params params;
  clipper::CLIPPER clipper(make_invariant(), params);
  Eigen::MatrixXd model, data;
  toy(model, data);
  clipper.scorePairwiseConsistency(model, data);
  Eigen::VectorXd u0 = Eigen::VectorXd::Ones(12);
  clipper.solve(u0);
  clipper::Association Ainliers = clipper.getSelectedAssociations();
  ASSERT_EQ(Ainliers.rows(), 3);
  for (size_t i = 0; i < (size_t)Ainliers.rows(); ++i) EXPECT_EQ(Ainliers(i, 0), Ainliers(i, 1));
  clipper.solve();  // random u0 path must run
  EXPECT_EQ(clipper.getSolution().u0.size(), 12);
}

// reference test/clipper_test.cpp:72-134 (solveAsMSRCSDR needs SCS; without it the reference
// returns no nodes -- the getter/setter round trip and the PGA solve are checked instead)
static void CLIPPER_EuclideanDistance_UseGetSet() {
  clipper::Params params;
  clipper::CLIPPER clipper(make_invariant(), params);
  Eigen::MatrixXd model, data;
  toy(model, data);
  clipper.scorePairwiseConsistency(model, data);
  clipper::Affinity M = clipper.getAffinityMatrix();
  clipper::Constraint C = clipper.getConstraintMatrix();
  clipper::CLIPPER clipper2(make_invariant(), params);
  clipper2.setMatrixData(M, C);
  EXPECT_EQ(clipper2.getAffinityMatrix(), M);
  EXPECT_EQ(clipper2.getConstraintMatrix(), C);
  clipper2.solveAsMSRCSDR();
  EXPECT_EQ(clipper2.getSolution().nodes.size(), (size_t)0);
  clipper2.solve(Eigen::VectorXd::Ones(12));
  clipper::Association Ainliers = clipper::utils::selectInlierAssociations(clipper2.getSolution(), clipper.getInitialAssociations());
  ASSERT_EQ(Ainliers.rows(), 3);
  for (size_t i = 0; i < (size_t)Ainliers.rows(); ++i) EXPECT_EQ(Ainliers(i, 0), Ainliers(i, 1));
}

// reference test/clipper_test.cpp:138-207
static void CLIPPER_EuclideanDistance_UseSparseGetSet() {
  clipper::Params params;
  clipper::CLIPPER clipper(make_invariant(), params);
  Eigen::MatrixXd model, data;
  toy(model, data);
  clipper.scorePairwiseConsistency(model, data);
  clipper::Affinity M = clipper.getAffinityMatrix();
  clipper::Constraint C = clipper.getConstraintMatrix();
  std::vector<clipper::SpTriplet> tm, tc;
  for (long j = 0; j < (long)M.cols(); ++j)
    for (long i = 0; i < j; ++i) {
      if (M(i, j) != 0) tm.emplace_back(i, j, M(i, j));
      if (C(i, j) != 0) tc.emplace_back(i, j, C(i, j));
    }
  clipper::SpAffinity Ms(M.rows(), M.cols());
  clipper::SpConstraint Cs(C.rows(), C.cols());
  Ms.setFromTriplets(tm.begin(), tm.end());
  Cs.setFromTriplets(tc.begin(), tc.end());
  clipper::CLIPPER clipper2(make_invariant(), params);
  clipper2.setSparseMatrixData(Ms, Cs);
  EXPECT_EQ(clipper2.getAffinityMatrix(), M);
  EXPECT_EQ(clipper2.getConstraintMatrix(), C);
  clipper2.solve(Eigen::VectorXd::Ones(12));
  clipper::Association Ainliers = clipper::utils::selectInlierAssociations(clipper2.getSolution(), clipper.getInitialAssociations());
  ASSERT_EQ(Ainliers.rows(), 3);
  for (size_t i = 0; i < (size_t)Ainliers.rows(); ++i) EXPECT_EQ(Ainliers(i, 0), Ainliers(i, 1));
}

// reference test/sdp_test.cpp:15-64 (no assertions upstream; here: runs, and the cluster is the
// strongest pair {5,12} of the DSD answer {3,5,12,14,15})
static void CLIPPERSDR_FindGlobalMax() {
  Eigen::MatrixXd M = m20();
  Eigen::MatrixXd C(20, 20);
  for (int j = 0; j < 20; ++j) for (int i = 0; i < 20; ++i) C(i, j) = M(i, j) > 0 ? 1.0 : 0.0;
  clipper::Params params;
  clipper::CLIPPER clipper(make_invariant(), params);
  clipper.setMatrixData(M, C);
  clipper.solve(Eigen::VectorXd::Ones(20));
  const clipper::Solution s = clipper.getSolution();
  ASSERT_EQ(s.nodes.size(), (size_t)2);
  EXPECT_TRUE((s.nodes[0] == 5 && s.nodes[1] == 12) || (s.nodes[0] == 12 && s.nodes[1] == 5));
  EXPECT_TRUE(std::abs(s.score - 1.9927) < 1e-4);
  clipper.solveAsMSRCSDR();
  // Rounding::DSD through the solver
  params.rounding = clipper::Params::Rounding::DSD;
  clipper::CLIPPER clipper3(make_invariant(), params);
  clipper3.setMatrixData(M, C);
  clipper3.solve(Eigen::VectorXd::Ones(20));
  EXPECT_TRUE(clipper3.getSolution().nodes.size() >= 2);
}

// reference test/dsd_test.cpp:14-44
static void DSD_Solve() {
  const std::vector<int> dsd_nodes = {3, 5, 12, 14, 15};
  std::vector<int> nodes = clipper::dsd::solve(m20());
  ASSERT_EQ(nodes.size(), dsd_nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) EXPECT_EQ(nodes[i], dsd_nodes[i]);
}

// reference test/dsd_test.cpp:48-80
static void DSD_SolveRestrictedGraph() {
  const std::vector<int> dsd_nodes = {3, 5, 12, 14, 15};
  const std::vector<int> S = {0, 1, 3, 5, 7, 12, 14, 15, 19};
  std::vector<int> nodes = clipper::dsd::solve(m20(), S);
  ASSERT_EQ(nodes.size(), dsd_nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) EXPECT_EQ(nodes[i], dsd_nodes[i]);
  // sparse overload (upper triangle only)
  Eigen::MatrixXd M = m20();
  std::vector<clipper::SpTriplet> t;
  for (int j = 0; j < 20; ++j) for (int i = 0; i < j; ++i) if (M(i, j) != 0) t.emplace_back(i, j, M(i, j));
  clipper::SpAffinity Ms(20, 20);
  Ms.setFromTriplets(t.begin(), t.end());
  EXPECT_EQ(clipper::dsd::solve(Ms, S), dsd_nodes);
}

static void Utils() {
  size_t k = 0;
  for (size_t i = 0; i < 9; ++i)
    for (size_t j = i + 1; j < 9; ++j) {
      size_t a, b; std::tie(a, b) = clipper::utils::k2ij(k++, 9);
      EXPECT_EQ(a, i); EXPECT_EQ(b, j);
    }
  Eigen::VectorXd x(4); x << 5, 5, 5, 7;
  EXPECT_EQ(clipper::utils::findIndicesOfkLargest(x, 2), (std::vector<int>{3, 1}));
  EXPECT_EQ(clipper::utils::findIndicesWhereAboveThreshold(x, 5.0), (std::vector<int>{3}));
  clipper::Association A = clipper::utils::createAllToAll(2, 3);
  EXPECT_EQ(A.rows(), 6); EXPECT_EQ(A(4, 0), 1); EXPECT_EQ(A(4, 1), 1);
  Eigen::VectorXi ind(4); ind << 1, 0, 0, 1;
  EXPECT_EQ(clipper::utils::selectFromIndicator(x, ind).size(), 2);
  EXPECT_EQ(clipper::utils::randvec(7).size(), 7);
}

// a user-defined invariant goes through the host path and must give the same graph
struct MyInvariant : clipper::invariants::PairwiseInvariant {
  double operator()(const clipper::invariants::Datum& ai, const clipper::invariants::Datum& aj,
                    const clipper::invariants::Datum& bi, const clipper::invariants::Datum& bj) override {
    double l1 = 0, l2 = 0;
    for (int q = 0; q < 3; ++q) { l1 += (ai(q) - aj(q)) * (ai(q) - aj(q)); l2 += (bi(q) - bj(q)) * (bi(q) - bj(q)); }
    const double c = std::abs(std::sqrt(l1) - std::sqrt(l2));
    return c < 0.06 ? std::exp(-0.5 * c * c / (0.01 * 0.01)) : 0.0;
  }
};
static void CustomInvariantHostPath() {
  clipper::Params params;
  clipper::CLIPPER a(make_invariant(), params), b(std::make_shared<MyInvariant>(), params);
  Eigen::MatrixXd model, data;
  toy(model, data);
  a.scorePairwiseConsistency(model, data);
  b.scorePairwiseConsistency(model, data);
  EXPECT_EQ(a.getAffinityMatrix(), b.getAffinityMatrix());
  EXPECT_EQ(a.getInitialAssociations(), b.getInitialAssociations());
  b.solve(Eigen::VectorXd::Ones(12));
  EXPECT_EQ(b.getSelectedAssociations().rows(), 3);
}

int main(int argc, char** argv) {
  const bool cpu_only = argc > 1 && std::strcmp(argv[1], "--cpu-only") == 0;
  struct T { const char* name; void (*fn)(); bool gpu; };
  const T tests[] = {
    {"DSD.Solve", DSD_Solve, false}, {"DSD.SolveRestrictedGraph", DSD_SolveRestrictedGraph, false}, {"Utils", Utils, false},
    {"Affinity.EuclideanDistance", Affinity_EuclideanDistance, true}, {"CLIPPER.EuclideanDistance", CLIPPER_EuclideanDistance, true},
    {"CLIPPER.EuclideanDistance_UseGetSet", CLIPPER_EuclideanDistance_UseGetSet, true},
    {"CLIPPER.EuclideanDistance_UseSparseGetSet", CLIPPER_EuclideanDistance_UseSparseGetSet, true},
    {"CLIPPERSDR.FindGlobalMax", CLIPPERSDR_FindGlobalMax, true}, {"CustomInvariantHostPath", CustomInvariantHostPath, true}};
  int ran = 0;
  for (const T& t : tests) {
    if (cpu_only && t.gpu) continue;
    const int before = g_fail;
    std::printf("[ RUN  ] %s\n", t.name);
    try { t.fn(); } catch (const std::exception& e) { ++g_fail; std::printf("  EXCEPTION %s\n", e.what()); }
    std::printf("[ %s ] %s\n", g_fail == before ? " OK " : "FAIL", t.name);
    ++ran;
  }
  std::printf("%d tests, %d checks, %d failures\n", ran, g_checks, g_fail);
  return g_fail ? 1 : 0;
}

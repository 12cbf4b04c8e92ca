// clipper_shell.cpp -- bodies of the reference's C++ API (include/clipper/*.h) over the C-ABI.
//
// This is the "reference-side binding": every method of clipper::CLIPPER marshals its Eigen-typed
// arguments into plain pointers and calls include/clipper_b200.h (which launches the sm_100a
// kernels).  Nothing numerical of the hot path happens in this file; what does happen here is
//   * the custom-invariant host path: a user-defined PairwiseInvariant cannot run in a
//     CUDA kernel, so it is evaluated pair by pair on the host (like reference clipper.cpp:31-56)
//     and the resulting dense matrices are uploaded with clp_set_dense;
//   * single-pair operator() of the two built-in invariants (API completeness; the reference
//     exposes them, e.g. py_clipper.cpp:43), never used for scoring association sets.
#include "clipper/clipper.h"
#include "clipper/utils.h"
#include "clipper_b200.h"

#include <cmath>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>

namespace clipper {

namespace {

void check(void* h, int rc, const char* what) {
  if (rc != CLP_OK)
    throw std::runtime_error(std::string("clipper_b200: ") + what + ": " + clp_last_error(static_cast<clp_handle>(h)));
}

double dist3(const invariants::Datum& a, const invariants::Datum& b, int off, int d) {
  double s = 0;
  for (int q = 0; q < d; ++q) { const double t = a(off + q) - b(off + q); s = s + t * t; }
  return std::sqrt(s);
}

}  // namespace

// ---- invariants: single-pair functors ---------------------------------------------------------
namespace invariants {

// reference src/invariants/euclidean_distance.cpp:13-31
double EuclideanDistance::operator()(const Datum& ai, const Datum& aj, const Datum& bi, const Datum& bj)
{
  const int d = (int)ai.size();
  const double l1 = dist3(ai, aj, 0, d);
  const double l2 = dist3(bi, bj, 0, d);
  if (params_.mindist > 0 && (l1 < params_.mindist || l2 < params_.mindist)) return 0.0;
  const double c = std::abs(l1 - l2);
  return (c < params_.epsilon) ? std::exp(-0.5 * c * c / (params_.sigma * params_.sigma)) : 0;
}

// reference src/invariants/pointnormal_distance.cpp:13-35 (acos deliberately not clamped)
double PointNormalDistance::operator()(const Datum& ai, const Datum& aj, const Datum& bi, const Datum& bj)
{
  const double l1 = dist3(ai, aj, 0, 3);
  const double l2 = dist3(bi, bj, 0, 3);
  const double alpha1 = std::acos(ai(3) * aj(3) + ai(4) * aj(4) + ai(5) * aj(5));
  const double alpha2 = std::acos(bi(3) * bj(3) + bi(4) * bj(4) + bi(5) * bj(5));
  const double dp = std::abs(l1 - l2);
  const double dn = std::abs(alpha1 - alpha2);
  if (dp < params_.epsp && dn < params_.epsn) {
    const double sp = std::exp(-0.5 * dp * dp / (params_.sigp * params_.sigp));
    const double sn = std::exp(-0.5 * dn * dn / (params_.sign * params_.sign));
    return sp * sn;
  }
  return 0.0;
}

}  // namespace invariants

// ---- CLIPPER -----------------------------------------------------------------------------------
CLIPPER::CLIPPER(const invariants::PairwiseInvariantPtr& invariant, const Params& params)
: params_(params), invariant_(invariant)
{}

void CLIPPER::setDevice(int device, int storage)
{
  device_ = device; storage_ = storage;
  handle_.reset();
}

void* CLIPPER::handle()
{
  if (!handle_) {
    clp_handle h = nullptr;
    const int rc = clp_create(device_, storage_, &h);
    if (rc != CLP_OK) throw std::runtime_error(std::string("clipper_b200: clp_create: ") + clp_last_error(nullptr));
    handle_ = std::shared_ptr<void>(h, [](void* p) { clp_destroy(static_cast<clp_handle>(p)); });
  }
  clp_params p;
  p.tol_u = params_.tol_u; p.tol_F = params_.tol_F; p.tol_Fop = params_.tol_Fop;
  p.maxiniters = params_.maxiniters; p.maxoliters = params_.maxoliters;
  p.beta = params_.beta; p.maxlsiters = params_.maxlsiters;
  p.eps = params_.eps; p.affinityeps = params_.affinityeps;
  p.rescale_u0 = params_.rescale_u0 ? 1 : 0; p.rounding = (int)params_.rounding;
  check(handle_.get(), clp_set_params(static_cast<clp_handle>(handle_.get()), &p), "clp_set_params");
  return handle_.get();
}

void CLIPPER::scorePairwiseConsistency(const invariants::Data& D1, const invariants::Data& D2, const Association& A)
{
  clp_handle h = static_cast<clp_handle>(handle());
  const int32_t* Ap = (A.size() == 0) ? nullptr : reinterpret_cast<const int32_t*>(A.data());
  const int64_t m = (A.size() == 0) ? 0 : (int64_t)A.rows();
  have_A_ = false;

  if (auto* e = dynamic_cast<invariants::EuclideanDistance*>(invariant_.get())) {
    const auto& ip = e->params();
    check(h, clp_score_euclidean(h, D1.data(), (int32_t)D1.rows(), (int64_t)D1.cols(), D2.data(), (int64_t)D2.cols(),
                                 Ap, m, ip.sigma, ip.epsilon, ip.mindist), "clp_score_euclidean");
    return;
  }
  if (auto* pn = dynamic_cast<invariants::PointNormalDistance*>(invariant_.get())) {
    if (D1.rows() != 6 || D2.rows() != 6) throw std::runtime_error("PointNormalDistance expects 6 x n data");
    const auto& ip = pn->params();
    check(h, clp_score_pointnormal(h, D1.data(), (int64_t)D1.cols(), D2.data(), (int64_t)D2.cols(), Ap, m,
                                   ip.sigp, ip.epsp, ip.sign, ip.epsn), "clp_score_pointnormal");
    return;
  }

  // custom invariant: host loop over the pairs (reference clipper.cpp:31-56), then upload
  if (A.size() == 0) A_ = utils::createAllToAll(D1.cols(), D2.cols());
  else A_ = A;
  have_A_ = true;
  const size_t mm = A_.rows();
  Eigen::MatrixXd M = Eigen::MatrixXd::Zero(mm, mm);
  Eigen::MatrixXd C = Eigen::MatrixXd::Zero(mm, mm);
  for (size_t j = 0; j < mm; ++j) {
    const invariants::Datum d1j = D1.col(A_(j, 0)), d2j = D2.col(A_(j, 1));
    for (size_t i = 0; i < j; ++i) {
      if (A_(i, 0) == A_(j, 0) || A_(i, 1) == A_(j, 1)) continue;
      const invariants::Datum d1i = D1.col(A_(i, 0)), d2i = D2.col(A_(i, 1));
      const double scr = (*invariant_)(d1i, d1j, d2i, d2j);
      if (scr > params_.affinityeps) { M(i, j) = scr; C(i, j) = 1; }  // C_ = pattern of M_ (clipper.cpp:63-64)
    }
  }
  check(h, clp_set_dense(h, M.data(), C.data(), (int64_t)mm), "clp_set_dense");
}

void CLIPPER::solve(const Eigen::VectorXd& u0)
{
  clp_handle h = static_cast<clp_handle>(handle());
  int64_t m = 0;
  check(h, clp_num_associations(h, &m), "clp_num_associations");
  if (u0.size() != 0 && (int64_t)u0.size() != m) throw std::runtime_error("clipper_b200: u0 has the wrong length");
  clp_solution s;
  Eigen::VectorXd u(m), u0used(m);
  std::vector<int32_t> nodes((size_t)std::max<int64_t>(m, 1));
  check(h, clp_solve(h, u0.size() == 0 ? nullptr : u0.data(), &s, u.data(), nodes.data(), u0used.data()), "clp_solve");
  soln_.t = s.t;
  soln_.ifinal = s.ifinal;
  soln_.nodes.assign(nodes.begin(), nodes.begin() + s.n_nodes);
  soln_.u0 = u0used;
  soln_.u = u;
  soln_.score = s.score;
  kernel_ms_ = s.kernel_ms;
  n_evals_ = s.n_evals;
}

// reference clipper.cpp:82-97
void CLIPPER::solveAsMaximumClique(const maxclique::Params& params)
{
  Eigen::MatrixXd C = getConstraintMatrix();
  for (long i = 0; i < (long)C.rows(); ++i) C(i, i) -= 1.0;
  utils::Timer tim;
  tim.start();
  std::vector<int> nodes = maxclique::solve(C, params);
  tim.stop();
  soln_.t = tim.getElapsedSeconds();
  soln_.ifinal = 0;
  std::swap(soln_.nodes, nodes);
  soln_.u = Eigen::VectorXd::Zero(C.cols());
  soln_.score = -1;
}

// reference clipper.cpp:101-113
void CLIPPER::solveAsMSRCSDR(const sdp::Params& params)
{
  Eigen::MatrixXd M = getAffinityMatrix();
  Eigen::MatrixXd C = getConstraintMatrix();
  sdp::Solution soln = sdp::solve(M, C, params);
  soln_.t = soln.t;
  soln_.ifinal = 0;
  std::swap(soln_.nodes, soln.nodes);
  soln_.u = Eigen::VectorXd::Zero(M.cols());
  soln_.score = -1;
}

Affinity CLIPPER::getAffinityMatrix()
{
  clp_handle h = static_cast<clp_handle>(handle());
  int64_t m = 0;
  check(h, clp_num_associations(h, &m), "clp_num_associations");
  Affinity M(m, m);
  if (m > 0) check(h, clp_get_dense(h, 0, M.data()), "clp_get_dense");
  return M;
}

Constraint CLIPPER::getConstraintMatrix()
{
  clp_handle h = static_cast<clp_handle>(handle());
  int64_t m = 0;
  check(h, clp_num_associations(h, &m), "clp_num_associations");
  Constraint C(m, m);
  if (m > 0) check(h, clp_get_dense(h, 1, C.data()), "clp_get_dense");
  return C;
}

void CLIPPER::setMatrixData(const Affinity& M, const Constraint& C)
{
  clp_handle h = static_cast<clp_handle>(handle());
  if (M.rows() != M.cols() || C.rows() != M.rows() || C.cols() != M.cols())
    throw std::runtime_error("clipper_b200: setMatrixData expects square M and C of equal size");
  have_A_ = false;
  check(h, clp_set_dense(h, M.data(), C.data(), (int64_t)M.rows()), "clp_set_dense");
}

void CLIPPER::setSparseMatrixData(const SpAffinity& Min, const SpConstraint& Cin)
{
  clp_handle h = static_cast<clp_handle>(handle());
  SpAffinity M = Min; SpConstraint C = Cin;
  M.makeCompressed(); C.makeCompressed();
  const int64_t m = (int64_t)M.cols();
  auto widen = [m](const int* p) { return std::vector<int64_t>(p, p + m + 1); };
  const std::vector<int64_t> cpM = widen(M.outerIndexPtr()), cpC = widen(C.outerIndexPtr());
  have_A_ = false;
  check(h, clp_set_sparse_upper(h, m, cpM.data(), M.innerIndexPtr(), M.valuePtr(), cpC.data(), C.innerIndexPtr(),
                                C.valuePtr()), "clp_set_sparse_upper");
}

Association CLIPPER::getInitialAssociations()
{
  if (have_A_) return A_;
  clp_handle h = static_cast<clp_handle>(handle());
  int64_t m = 0;
  check(h, clp_num_associations(h, &m), "clp_num_associations");
  Association A(m, 2);
  if (m > 0) check(h, clp_get_associations(h, reinterpret_cast<int32_t*>(A.data())), "clp_get_associations");
  return A;
}

Association CLIPPER::getSelectedAssociations()
{
  return utils::selectInlierAssociations(soln_, getInitialAssociations());
}

// ---- utils -------------------------------------------------------------------------------------
namespace utils {

Eigen::VectorXd randvec(size_t n)
{
  std::random_device rd;
  std::mt19937 gen(rd());
  std::uniform_real_distribution<double> dis(0, 1);
  Eigen::VectorXd v(n);
  for (size_t i = 0; i < n; ++i) v(i) = dis(gen);
  return v;
}

std::vector<int> findIndicesOfkLargest(const Eigen::VectorXd& x, int k)
{
  if (k < 1 || x.size() == 0) return {};
  std::vector<int32_t> out((size_t)std::min<long>(k, (long)x.size()));
  const int n = clp_find_k_largest(x.data(), (int64_t)x.size(), k, out.data());
  return std::vector<int>(out.begin(), out.begin() + n);
}

std::vector<int> findIndicesWhereAboveThreshold(const Eigen::VectorXd& x, double thr)
{
  std::vector<int32_t> out((size_t)std::max<long>(1, (long)x.size()));
  const int n = clp_find_above(x.data(), (int64_t)x.size(), thr, out.data());
  return std::vector<int>(out.begin(), out.begin() + n);
}

Association createAllToAll(size_t n1, size_t n2)
{
  Association A(n1 * n2, 2);
  if (n1 * n2 > 0) clp_create_all_to_all((int64_t)n1, (int64_t)n2, reinterpret_cast<int32_t*>(A.data()));
  return A;
}

Eigen::VectorXd selectFromIndicator(const Eigen::VectorXd& x, const Eigen::VectorXi& ind)
{
  long cnt = 0;
  for (long i = 0; i < (long)ind.size(); ++i) cnt += ind(i) ? 1 : 0;
  Eigen::VectorXd y(cnt);
  long k = 0;
  for (long i = 0; i < (long)x.size(); ++i) if (ind(i)) y(k++) = x(i);
  return y;
}

Association selectInlierAssociations(const Solution& soln, const Association& A)
{
  Association Ain(soln.nodes.size(), 2);
  for (size_t i = 0; i < soln.nodes.size(); ++i) {
    Ain(i, 0) = A(soln.nodes[i], 0);
    Ain(i, 1) = A(soln.nodes[i], 1);
  }
  return Ain;
}

std::tuple<size_t,size_t> k2ij(size_t k, size_t n)
{
  uint64_t i = 0, j = 0;
  clp_k2ij(k, n, &i, &j);
  return {(size_t)i, (size_t)j};
}

}  // namespace utils

// ---- dsd / sdp / maxclique front-ends -------------------------------------------------------------
namespace dsd {

std::vector<int> solve(const Eigen::MatrixXd& A, const std::vector<int>& S)
{
  const int64_t n = (int64_t)A.rows();
  std::vector<int32_t> out((size_t)std::max<int64_t>(n, 1));
  std::vector<int32_t> Sin(S.begin(), S.end());
  const int k = clp_dsd_dense(A.data(), n, Sin.empty() ? nullptr : Sin.data(), (int32_t)Sin.size(), out.data());
  if (k < 0) throw std::runtime_error("clipper_b200: dsd::solve failed");
  return std::vector<int>(out.begin(), out.begin() + k);
}

std::vector<int> solve(const SpAffinity& Ain, const std::vector<int>& S)
{
  // the sparse overload reads the upper triangle only (reference dsd.cpp:300-302): densify it
  SpAffinity A = Ain;
  A.makeCompressed();
  const long n = (long)A.rows();
  Eigen::MatrixXd D = Eigen::MatrixXd::Zero(n, n);
  for (long j = 0; j < (long)A.cols(); ++j)
    for (int q = A.outerIndexPtr()[j]; q < A.outerIndexPtr()[j + 1]; ++q) {
      const long i = A.innerIndexPtr()[q];
      if (i < j) { D(i, j) = A.valuePtr()[q]; D(j, i) = A.valuePtr()[q]; }
    }
  return solve(D, S);
}

}  // namespace dsd

namespace sdp {
Solution solve(const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Params&)
{
  std::cout << "Warning: clipper was not built with SCS; SDR solver unavailable." << std::endl;
  return {};
}
}  // namespace sdp

namespace maxclique {
std::vector<int> solve(const Eigen::MatrixXd&, const Params&)
{
  std::cout << "Warning: clipper was not built with PMC; maximum clique solver unavailable." << std::endl;
  return {};
}
}  // namespace maxclique

}  // namespace clipper

/**
 * @file types.h
 * @brief Types of the CLIPPER public API (mirror of reference include/clipper/types.h:15-23)
 *
 * With Eigen3 installed these ARE the reference's Eigen types, so existing callers compile
 * unchanged.  Without Eigen (this repository's offline build box) a minimal same-named
 * vocabulary is used, see clipper/compat/mini_eigen.h.
 */
#pragma once

#if defined(__has_include)
#  if __has_include(<Eigen/Dense>) && !defined(CLIPPER_FORCE_MINI_EIGEN)
#    define CLIPPER_HAS_EIGEN 1
#  endif
#endif

#ifdef CLIPPER_HAS_EIGEN
#  include <Eigen/Dense>
#  include <Eigen/Sparse>
#else
#  include "clipper/compat/mini_eigen.h"
#endif

namespace clipper {

  using SpMat = Eigen::SparseMatrix<double>;
  using SpTriplet = Eigen::Triplet<double>;

  using Association = Eigen::Matrix<int, Eigen::Dynamic, 2>;   // column-major m x 2
  using Affinity = Eigen::MatrixXd;
  using Constraint = Eigen::MatrixXd;

  using SpAffinity = SpMat;
  using SpConstraint = SpMat;

} // ns clipper

/**
 * @file euclidean_distance.h
 * @brief Pairwise Euclidean-distance invariant (mirror of reference invariants/euclidean_distance.h:19-49)
 */
#pragma once

#include "clipper/invariants/abstract.h"

namespace clipper {
namespace invariants {

  class EuclideanDistance : public PairwiseInvariant
  {
  public:
    struct Params
    {
      double sigma = 0.01;   ///< spread of the exponential kernel
      double epsilon = 0.06; ///< bound on the consistency score
      double mindist = 0;    ///< minimum allowable distance between inlier points of one data set
    };
  public:
    EuclideanDistance(const Params& params) : params_(params) {}
    ~EuclideanDistance() = default;

    /// single-pair evaluation (reference euclidean_distance.cpp:13-31); scoring whole
    /// association sets goes through the GPU kernel, not through this functor
    double operator()(const Datum& ai, const Datum& aj, const Datum& bi, const Datum& bj) override;

    /// additive accessor (the reference keeps params_ private with no getter)
    const Params& params() const { return params_; }

  private:
    Params params_;
  };
  using EuclideanDistancePtr = std::shared_ptr<EuclideanDistance>;

} // ns invariants
} // ns clipper

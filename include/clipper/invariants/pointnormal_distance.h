/**
 * @file pointnormal_distance.h
 * @brief Pairwise point-normal invariant (mirror of reference invariants/pointnormal_distance.h:22-53)
 *        Data: 6 x n, datum = [point(3); normal(3)].
 */
#pragma once

#include "clipper/invariants/abstract.h"

namespace clipper {
namespace invariants {

  class PointNormalDistance : public PairwiseInvariant
  {
  public:
    struct Params
    {
      double sigp = 0.5;  ///< point - spread of exp kernel
      double epsp = 0.5;  ///< point - bound on consistency score
      double sign = 0.10; ///< normal - spread of exp kernel
      double epsn = 0.35; ///< normal - bound on consistency score
    };
  public:
    PointNormalDistance(const Params& params) : params_(params) {}
    ~PointNormalDistance() = default;

    /// single-pair evaluation (reference pointnormal_distance.cpp:13-35)
    double operator()(const Datum& ai, const Datum& aj, const Datum& bi, const Datum& bj) override;

    const Params& params() const { return params_; }

  private:
    Params params_;
  };
  using PointNormalDistancePtr = std::shared_ptr<PointNormalDistance>;

} // ns invariants
} // ns clipper

/**
 * @file sdp.h
 * @brief Semidefinite relaxation front-end (mirror of reference sdp.h:20-57).
 *        The relaxation is solved by the third-party SCS library in the reference
 *        (CMakeLists.txt:83-86); it is not part of the GPU hot path and SCS is not bundled, so
 *        solve() behaves like a reference build without CLIPPER_HAS_SCS (sdp.cpp:298-302).
 */
#pragma once
#include <vector>
#include "clipper/types.h"

namespace clipper {
namespace sdp {

  struct Solution
  {
    Eigen::MatrixXd X;
    Eigen::VectorXd lambdas;
    Eigen::VectorXd evec1;
    double thr = 0;
    std::vector<int> nodes;
    int iters = 0;
    float pobj = 0, dobj = 0;
    double t = 0, t_parse = 0, t_scs = 0, t_scs_setup = 0, t_scs_solve = 0, t_scs_linsys = 0,
           t_scs_cone = 0, t_scs_accel = 0, t_extract = 0;
  };

  struct Params
  {
    bool verbose = false;
    int max_iters = 2000;
    int acceleration_interval = 10;
    int acceleration_lookback = 10;
    float eps_abs = 1e-3;
    float eps_rel = 1e-3;
    float eps_infeas = 1e-7;
    float time_limit_secs = 0;
  };

  Solution solve(const Eigen::MatrixXd& M, const Eigen::MatrixXd& C, const Params& params = Params{});

} // ns sdp
} // ns clipper

// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Deterministic restatement of what Estimator::optimization() asks of ceres::Solve
// (vins_estimator/src/estimator.cpp:803-818): TRUST_REGION minimiser, DOGLEG (traditional) strategy,
// DENSE_SCHUR linear solver, Jacobi scaling, max_num_iterations = NUM_ITERATIONS, every other option at
// its Ceres 1.14 default.  The Ceres sources are not in /root/reference; this follows the published
// algorithm of Ceres 1.14's TrustRegionMinimizer / DoglegStrategy and is NOT
// pinned against Ceres itself ("parity unpinned": no Ceres build or golden vector is available offline).
// The reference's wall-clock cap (max_solver_time_in_seconds) is deliberately absent: it makes the
// reference's own output machine-speed dependent.
#pragma once
#include <memory>
#include <vector>

#include "be_factors.h"

namespace orc {

struct SolveSummary {
    int iterations = 0;            // trust-region step attempts (Ceres counts iteration 0 separately)
    int successful_steps = 0;
    int linear_solver_retries = 0;
    double initial_cost = 0, final_cost = 0;
    int termination = 0;           // 0 iteration cap, 1 parameter tol, 2 function tol, 3 gradient tol, 4 failure
};

class Problem {
  public:
    void AddParameterBlock(double* ptr, int size, bool is_pose);
    void SetParameterBlockConstant(double* ptr);
    void AddResidualBlock(std::shared_ptr<CostFunction> cf, const CauchyLoss* loss, std::vector<double*> params);

    struct Block {
        double* ptr;
        int size, local;
        bool is_pose, constant = false, used = false;
        int offset = -1;  // column offset in the reduced local parameter vector
    };
    struct Res {
        std::shared_ptr<CostFunction> cf;
        const CauchyLoss* loss;
        std::vector<int> blocks;
    };
    std::vector<Block> blocks;
    std::vector<Res> residuals;
    int find(double* ptr) const;
};

SolveSummary Solve(Problem& problem, int max_num_iterations);

// Cross-check hooks (tests/test_oracle_solver_minimum.py): the robustified residual vector of the whole problem,
// sqrt(rho(|r_i|^2)) r_i / |r_i| per residual block (so that its squared norm is twice the Ceres cost), at the current
// parameter values (+) delta, delta in the local coordinates of the non-constant blocks (poses 6, others their size).
// An independent optimiser (scipy) minimises it and must land on the minimum Solve converges to.
int ProbeColumns(Problem& problem);
int ProbeResiduals(Problem& problem);
void ProbeEvaluate(Problem& problem, const double* delta, double* out);

}  // namespace orc

// Device-side data layout of the sliding-window BA path (all float64, structure-of-arrays in HBM).
//
// One "BaProblem" = one sequence's window at one frame: states (current + candidate), the observation
// table grouped by landmark, the pre-integration slots, the marginalization prior in information form,
// the accumulation buffers of the normal equations and the trust-region state.  Kernels take a BaProblem
// by value (a bundle of device pointers) so a batch of sequences is just an array of these.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vb {

constexpr int BA_MAX_OBS_PER_LM = 32;  // lanes of the landmark warp (window size + 1 <= 32)

struct PreInt {  // IntegrationBase (integration_base.h:189-209) + cached sqrt_info (imu_factor.h:64)
    double sum_dt;
    double dp[3], dq[4] /* w x y z */, dv[3];
    double ba[3], bg[3];      // linearized_ba / linearized_bg
    double acc0[3], gyr0[3];  // last sample (acc_0 / gyr_0)
    double jac[225], cov[225], sqrt_info[225];
};

struct BaDims {
    int W;        // WINDOW_SIZE; frames 0..W
    int D;        // reduced (camera-side) dimension: 6(W+1) + 9(W+1) [+6] [+1]
    int L;        // landmarks in this problem
    int M;        // non-anchor observations (= visual residual blocks)
    int est_ex, est_td;
    int col_sb;   // first speed-bias column = 6(W+1)
    int col_ex;   // -1 when the extrinsic is constant
    int col_td;   // -1 when td is not estimated
    double sqrt_info_vis;  // FOCAL_LENGTH / 1.5
    double tr_over_row;    // TR / ROW
    double half_row;       // ROW / 2 (projection_td_factor.cpp:18-19)
    double G[3];
};

struct BaStates {  // one point x: Ceres parameter layouts
    double* pose;  // (W+1) x 7: p, qx qy qz qw
    double* sb;    // (W+1) x 9: v, ba, bg
    double* ex;    // 7
    double* td;    // 1
    double* lam;   // L inverse depths
};

struct BaAccum {  // normal equations of one linearization point (unscaled)
    double* Hpp;  // D x D, upper triangle (row <= col) accumulated, rest zero
    double* gp;   // D
    double* Hpl;  // L x D dense rows (J_l^T J_p)
    double* Hll;  // L
    double* gl;   // L
    double* cost; // 1
};

struct BaPrior {  // MarginalizationInfo in information form: A = J0^T J0, g0 = J0^T r0, c0 = |r0|^2
    int n;            // rows/cols
    int nblocks;
    const int* type;  // per kept block: 0 pose, 1 speed-bias, 2 ex pose, 3 td
    const int* index; // frame index for pose / speed-bias blocks
    const int* off;   // local offset in the prior vector
    const double* x0; // nblocks x 9 linearization point (global parameterisation)
    const double* A;  // n x n
    const double* g0; // n
    const double* c0; // 1
};

struct SolverState {  // trust-region / dogleg state, lives in device memory
    int iteration, max_iterations;
    int done;        // 0 running; 1 iteration cap; 2 parameter tol; 3 function tol; 4 failure
    int cur;         // which of the two state/accumulation buffers holds the current point
    int reuse, cand_valid, first, successful, retries, invalid_streak;
    int alpha_valid, pad2;
    unsigned lin_ticket;
    int pad;
    double radius, mu, x_cost, cand_cost, model_cost_change, dogleg_step_norm, alpha, x_norm, step_norm;
    double initial_cost;
    long long clk[10];  // per-phase cycle counters of the last ba_step_kernel (profiling aid)
};

struct BaProblem {
    BaDims dims;
    BaStates x[2];
    BaAccum acc[2];
    // observation table
    const int* lm_anchor;     // L: frame of the first observation
    const int* lm_start;      // L+1: offsets into the ob_* arrays
    const double* lm_pts;     // L x 2: normalised point in the anchor frame (z = 1)
    const double* lm_vel;     // L x 2
    const double* lm_td;      // L
    const double* lm_row;     // L: raw v pixel coordinate
    const int* ob_frame;      // M
    const double* ob_pts;     // M x 2
    const double* ob_vel;     // M x 2
    const double* ob_td;      // M
    const double* ob_row;     // M
    // IMU factors: factor k links frames k and k+1 through pre-integration slot imu_slot[k] (-1: skipped)
    const PreInt* preint;     // slot array
    const int* imu_slot;      // W
    BaPrior prior;
    // solver workspace
    double* S;       // D x D Schur complement (full symmetric)
    double* Spk;     // the same, packed row-major lower triangle (what the Cholesky consumes)
    double* Hfull;   // D x D symmetrised Hpp
    double* gred;    // D
    double* scale;   // D + L Jacobi scaling
    double* diag;    // D + L trust-region diagonal
    double* grad;    // D + L
    double* gn;      // D + L
    double* work;    // 4 x (D + L) scratch
    SolverState* st;
};

}  // namespace vb

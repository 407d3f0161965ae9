// Launch interface of the BA kernels (ba_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "ba_types.h"
#include "kprof.h"

namespace vb {

constexpr int BA_MAX_FRAMES = 32;

struct MargPlan {  // dense marginalisation system layout: [m_dense | n_lm landmark columns | n kept]
    int P, m_dense, n_lm, n;
    const int* lms;     // device: indices (into the problem's landmark table) of the marginalised landmarks
    const int* col_lm;  // device: their columns
    int col_pose[BA_MAX_FRAMES], col_sb[BA_MAX_FRAMES];  // -1 when the block does not take part
    int col_ex, col_td;
    int use_imu;        // include IMU factor (frames 0,1)
    double* Am;         // P x P (upper triangle accumulated)
    double* bm;         // P
    double* Aout;       // n x n   new prior A  (after the eps floor)
    double* gout;       // n       new prior g0
    double* cout;       // 1       new prior c0
    double* Araw;       // n x n   Schur complement before the eps floor (tests), may be null
    double* graw;       // n
    double* Wglobal;    // q x q scratch (q = m_dense + n) used when the reduced system does not fit shared memory
    int w_in_global;    // set by launch_marginalize
};

void launch_preint_push(PreInt* slot, int n, const double* d_samples, double acc_n, double gyr_n, double acc_w, double gyr_w,
                        cudaStream_t s);
void launch_preint_init(PreInt* slot, const double* acc0, const double* gyr0, const double* ba, const double* bg, cudaStream_t s);
void launch_sqrt_info(PreInt* slots, const int* d_which, int count, cudaStream_t s);
// full trust-region solve: linearise x[st->cur], then max_iterations x {schur, step, zero, linearise+decide}
// profile slots: 0 linearize, 1 schur, 2 step, 3 zero, 4 marg_build, 5 marg_solve, 6 preint, 7 sqrt_info
void launch_ba_solve(const BaProblem& p, int max_iterations, cudaStream_t s, int* launches, KernelProfile* prof = nullptr);
void launch_marginalize(const BaProblem& p, const MargPlan& mp, cudaStream_t s, int* launches, KernelProfile* prof = nullptr);
size_t ba_work_doubles(int D, int L);

}  // namespace vb

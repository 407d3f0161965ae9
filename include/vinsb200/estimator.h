/* vinsb200 back end — C ABI of the sliding-window estimator hot path.
 *
 * Replaces, one entry point per reference interface (vins_estimator/src/estimator.h):
 *   ve_create / ve_destroy     Estimator::Estimator + setParameter + readParameters
 *                              (estimator.cpp:3-20, parameters.cpp:42-137)
 *   ve_clear_state             Estimator::clearState                         estimator.h:39, estimator.cpp:22-82
 *   ve_process_imu             Estimator::processIMU(dt, acc, gyr)           estimator.h:34, estimator.cpp:84-118
 *   ve_process_image           Estimator::processImage(image, header)        estimator.h:35, estimator.cpp:120-217
 *                              (runs solveOdometry -> triangulate -> optimization() -> marginalisation ->
 *                               slideWindow exactly in the reference's order; optimization() is
 *                               estimator.h:47, estimator.cpp:670-1003)
 *   ve_get_states              the public state arrays Ps, Rs (as quaternions), Vs, Bas, Bgs, td   estimator.h:71-79
 *   ve_info                    solver_flag, frame_count, marginalization_flag (estimator.h:65-66) + solver summary
 *   ve_set_seed                stand-in for initialStructure() (estimator.cpp:218-362):
 *                              the first window is seeded from a caller-supplied trajectory
 *   ve_get_prior               last_marginalization_info in information form (tests)
 *
 * `image` is the feature message: n points with ids (feature_id) and 7 doubles each
 * (x, y, z, u, v, velocity_x, velocity_y), i.e. map<int, vector<pair<int, Matrix<double,7,1>>>> flattened
 * for NUM_OF_CAM = 1.  Plain C, int status (0 ok, < 0 ve_status), no exceptions, no CPU fallback.  A handle is
 * externally synchronised like the reference's m_estimator mutex (estimator_node.cpp:220).
 */
#ifndef VINSB200_ESTIMATOR_H
#define VINSB200_ESTIMATOR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ve_status { VE_OK = 0, VE_ERR_INVALID = -1, VE_ERR_NO_DEVICE = -2, VE_ERR_CUDA = -3, VE_ERR_CAPACITY = -4 } ve_status;

typedef struct ve_config {
    int window_size;         /* WINDOW_SIZE (parameters.h:12), runtime here; 3..22 (ve_create rejects larger windows) */
    int max_features;        /* NUM_OF_F (parameters.h:13): landmark capacity, bound-checked */
    int num_iterations;      /* max_num_iterations (euroc_config.yaml:55) */
    int estimate_extrinsic;  /* 0 or 1 (2 = online calibration is not part of this path) */
    int estimate_td;
    double focal_length;     /* FOCAL_LENGTH = 460 (parameters.h:11) */
    double keyframe_parallax;/* keyframe_parallax in pixels (euroc_config.yaml:56) */
    double acc_n, gyr_n, acc_w, gyr_w, g_norm;
    double init_depth;       /* INIT_DEPTH = 5.0 (parameters.cpp:114) */
    double td, tr, row;      /* td, rolling_shutter_tr (0 when global shutter), image_height */
    double tic[3];           /* extrinsicTranslation */
    double ric[9];           /* extrinsicRotation, row-major */
    int device;
} ve_config;

typedef struct ve_estimator ve_estimator;

int ve_create(const ve_config* cfg, ve_estimator** out);
void ve_destroy(ve_estimator* e);
const char* ve_last_error(const ve_estimator* e);
int ve_clear_state(ve_estimator* e);

/* rows: n x 11 doubles (t, p[3], q[wxyz], v[3]) looked up by image stamp when the window first fills. */
int ve_set_seed(ve_estimator* e, int n, const double* rows, const double* ba, const double* bg);

int ve_process_imu(ve_estimator* e, double dt, const double* acc, const double* gyr);
int ve_process_image(ve_estimator* e, int n, const int* ids, const double* xyz_uv_vel, double stamp);

/* out: (window_size + 1) x 16 doubles: p[3] q[wxyz] v[3] ba[3] bg[3]; td may be NULL */
int ve_get_states(const ve_estimator* e, double* out, double* td);
/* out10: solver_flag, frame_count, marginalization_flag, n_solves, n_reboots, landmarks, visual factors,
 *        iterations, successful steps, termination; costs2: initial, final cost of the last solve */
int ve_info(const ve_estimator* e, int* out10, double* costs2);
/* Schur complement (before the eps floor) of the last marginalisation: returns n, fills A (n x n) and b (n)
 * in the canonical block order poses ascending, speed-biases ascending, ex pose, td; blocks4 gets (type, index,
 * offset, size) per kept block.  Returns -n if cap is too small. */
int ve_get_prior(const ve_estimator* e, int cap, double* A, double* b, int* nblocks, int* blocks4);
/* Device milliseconds (CUDA events) of the last process_image: [0] pre-integration, [1] solve, [2] marginalisation,
 * [3] total; launches = kernels launched. */
int ve_last_timing(const ve_estimator* e, float* ms4, int* launches);

/* n consecutive ve_process_imu calls (same semantics; saves per-call overhead in scripting hosts). */
int ve_process_imu_batch(ve_estimator* e, int n, const double* dt, const double* acc, const double* gyr);
/* Kernel profiling (serialises the pipeline; not for timed runs): accumulated device ms and launch counts per kernel:
 * 0 ba_linearize, 1 ba_schur, 2 ba_step, 3 ba_zero, 4 marg_build, 5 marg_solve, 6 preint_push, 7 sqrt_info. */
int ve_set_profile(ve_estimator* e, int on);
int ve_kernel_times(const ve_estimator* e, double* ms8, int* count8);
/* Host<->device bytes moved by the last ve_process_image. */
int ve_last_traffic(const ve_estimator* e, double* h2d_bytes, double* d2h_bytes);

/* Solver internals of the last solve (profiling/tests): out[0] linear-solver retries, [1] mu, [2] radius,
 * [3..10] per-phase cycle counters of the step kernel summed over the iterations, [11..12] cycles of the
 * tridiagonalisation / QL phases of the last marginalisation's prior eigen-decomposition, [13..17] its phase cycle counters. */
int ve_solver_debug(const ve_estimator* e, double* out18);

#ifdef __cplusplus
}
#endif
#endif

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
 *   ve_set_seed                OPTIONAL external initialiser: a trajectory that covers the first full window replaces
 *                              initialStructure() (estimator.cpp:218-362).  Without it the estimator initialises
 *                              itself (section "Initialisation" below), also after a failureDetection reboot.
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
typedef struct ve_batch ve_batch;

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

/* WINDOW_SIZE of the handle; acc_0 / gyr_0 (the last IMU sample handed to processIMU, estimator.h:84) and g (estimator.h:69),
 * read by the node's update() (estimator_node.cpp:80-96); any pointer may be NULL. */
int ve_window_size(const ve_estimator* e);
int ve_get_latest_imu(const ve_estimator* e, double* acc0, double* gyr0, double* g3);
/* tic (3) and ric (9, row-major) as currently estimated (estimator.h:75-76); either may be NULL. */
int ve_get_extrinsic(const ve_estimator* e, double* tic3, double* ric9);

/* ---- Batched sequences (SURVEY.md 8b "Threading", BASELINE configs[2] / [4]) --------------------------------------
 * n independent estimators with one configuration that advance frame by frame TOGETHER: one host-to-device copy, one
 * kernel launch per solver stage (the member index is a grid dimension) and one device-to-host copy per frame for the
 * whole batch, the host bookkeeping of the members on a thread pool.  Within a member the serial chain of the reference
 * (last_marginalization_info feeding the next optimisation, estimator.cpp:929-930) is kept: frame k+1 of a member is
 * enqueued behind its own frame k.  A stand-alone handle from ve_create is a batch of one running the same code.
 *
 *   ve_batch_member            borrowed handle of member k for the per-sequence calls: ve_set_seed, ve_process_imu(_batch),
 *                              ve_clear_state, ve_get_states, ve_info, ve_get_prior, ve_last_traffic (NOT ve_process_image /
 *                              ve_destroy: members are driven and destroyed through the batch)
 *   ve_batch_process_image     Estimator::processImage for every member k with active[k] != 0 (NULL = all): n[k] points,
 *                              ids[k], xyz_uv_vel[k], stamps[k] as in ve_process_image.  status (may be NULL) receives each
 *                              member's ve_status; the return value is the first non-zero one.
 *   ve_batch_last_timing       device ms of the last batch frame: [0] pre-integration, [1] solve, [2] marginalisation, [3] total
 */
int ve_batch_create(const ve_config* cfg, int n, ve_batch** out);
void ve_batch_destroy(ve_batch* b);
int ve_batch_size(const ve_batch* b);
/* Launch groups: the members are split into this many contiguous groups, each with its own stream and launch chain, so that the
 * chains overlap on the GPU (default: groups of 16 members for batches of 32 and more; VINSB200_BATCH_GROUPS overrides). */
int ve_batch_groups(const ve_batch* b);
ve_estimator* ve_batch_member(ve_batch* b, int k);
const char* ve_batch_last_error(const ve_batch* b);
int ve_batch_process_image(ve_batch* b, const int* active, const int* n, const int* const* ids, const double* const* xyz_uv_vel,
                           const double* stamps, int* status);
int ve_batch_last_timing(const ve_batch* b, float* ms4, int* launches);
int ve_batch_set_profile(ve_batch* b, int on);
int ve_batch_kernel_times(const ve_batch* b, double* ms8, int* count8);
/* Waits until everything the batch has enqueued (including the last marginalisation) has finished. */
int ve_batch_sync(ve_batch* b);

/* ---- Single-factor test entries (parity tests; projection_factor.cpp:176-224 has the reference's own check()) -------
 * The device functions of the solve evaluated on caller-supplied parameter blocks, Jacobians in the 6-dof tangent
 * parameterisation (the reference's 7th column is identically zero).
 *   ve_debug_projection_factor  params23 = pose_i 7 (p, qx qy qz qw) | pose_j 7 | ex_pose 7 | inverse depth | td;
 *                               data12 = pts_i xy, pts_j xy, velocity_i xy, velocity_j xy, td_i, td_j, row_i, row_j;
 *                               out43 = residual 2 | J row 0 (20: pose_i 6, pose_j 6, ex 6, depth, td) | J row 1 | rho/2.
 *                               use_td selects ProjectionTdFactor; robust applies CauchyLoss(1) + the corrector.
 *   ve_debug_imu_factor         integrates n samples (sample 0 = acc_0 / gyr_0 of the IntegrationBase, dt[0] unused) with
 *                               noise4 = acc_n, gyr_n, acc_w, gyr_w and linearised biases ba, bg on the device;
 *                               out_preint (686) = sum_dt | delta_p 3 | delta_q wxyz | delta_v 3 | jacobian 225 | covariance 225 |
 *                               sqrt_info 225; with params32 = pose_i 7 | speedbias_i 9 | pose_j 7 | speedbias_j 9,
 *                               out_factor (465) = whitened residual 15 | whitened Jacobian 15 x 30 (pose_i 6, sb_i 9, pose_j 6, sb_j 9). */
int ve_debug_projection_factor(const double* params23, const double* data12, int use_td, double focal_length, double tr, double row,
                               int robust, double* out43);
int ve_debug_imu_factor(const double* noise4, double g_norm, const double* ba, const double* bg, int n, const double* dt,
                        const double* acc, const double* gyr, const double* params32, double* out_preint, double* out_factor);

/* ---- Initialisation (Estimator::initialStructure, vins_estimator/src/estimator.cpp:218-471) ----------
 * The estimator bootstraps itself from the first full window (relative pose -> global SfM -> PnP of every image ->
 * visual-inertial alignment) unless a seed trajectory covering the window was supplied with ve_set_seed; with
 * ve_config.estimate_extrinsic = 2 it first calibrates the camera-IMU rotation from the image / gyroscope rotation pairs
 * (CalibrationExRotation, estimator.cpp:140-156) and then continues with estimate_extrinsic = 1.  The stages are
 * host code in the reference as well; the entries below run them on caller-supplied arrays WITHOUT a handle or a GPU so
 * that they can be pinned against OpenCV / numpy twins (tests/test_host_initial.py).
 *   ve_debug_relative_rt        solve_5pts.cpp:193-227: corres4 = n x (x0 y0 x1 y1) normalised coordinates; R9 / T3 = the
 *                               reference's Rotation / Translation; returns 1 when inlier_cnt > 12, inliers = inlier_cnt.
 *   ve_debug_solve_pnp          cv::solvePnP(obj, img, I, none, rvec, t, useExtrinsicGuess = 1): R9 / t3 in-out
 *                               (world -> camera); returns 1 on success.
 *   ve_debug_sfm_construct      GlobalSFM::construct (initial_sfm.cpp:117-312): tracks = per feature id, first window
 *                               frame, number of consecutive observations, xy per observation; q_wxyz (4 per frame) / T
 *                               (3 per frame) = camera-to-world poses; pt_ids / pts = sfm_tracked_points (capacity
 *                               n_tracks); function_tolerance <= 0 selects Ceres' default 1e-6 (the bundle's stopping
 *                               rule; tests tighten it to reach the minimum); returns 1 on success.
 *   ve_debug_initial_structure  the whole of initialStructure up to VisualIMUAlignment.  Window headers[F]; all image
 *                               frames (n_all stamps; per frame its feature ids ascending + xy through pts_off, and the IMU
 *                               samples dt acc gyr (7 doubles) pre-integrated INTO that frame through imu_off, function_tolerance as above, lin6 =
 *                               the acc / gyr sample its integration starts from); tracks as above.  Outputs: frame_R
 *                               (9 per frame, already x RIC^T) / frame_T, x (3 per frame velocities | 2 | scale), g3 (before the
 *                               yaw alignment), delta_bg3, info4 = l, bundle iterations, is-keyframe count, reserved;
 *                               bundle_cost.  Returns 0 or the failing stage (1 relative pose, 2 SfM, 3 PnP, 4 alignment). */
/*   ve_debug_ex_rotation        InitialEXRotation::CalibrationExRotation (initial_ex_rotation.cpp:11-67; ve_config.estimate_extrinsic
 *                               = 2, with ric = I and tic = 0 as parameters.cpp:101-106 sets them) called n_steps times in a row:
 *                               step k gets the correspondences corres4[corres_off[k] .. corres_off[k+1]) (x0 y0 x1 y1 between the
 *                               two newest frames) and the gyroscope rotation dq_wxyz[4 k ..]; ric_out (9 per step) = the running
 *                               estimate, ok_out = its return value, cov_out (may be NULL) = the singular value it thresholds, rc_out (may be NULL,
 *                               9 per step) = the camera rotation solveRelativeR extracted from the step's correspondences; rc_in (may be
 *                               NULL, 9 per step): use these camera rotations instead (pins the quaternion system by itself). */
int ve_debug_ex_rotation(int n_steps, const int* corres_off, const double* corres4, const double* dq_wxyz, int window_size, double* ric_out,
                         int* ok_out, double* cov_out, double* rc_out, const double* rc_in);
int ve_debug_relative_rt(const double* corres4, int n, double* R9, double* T3, int* inliers);
int ve_debug_solve_pnp(const double* pts3, const double* pts2, int n, double* R9, double* t3);
int ve_debug_sfm_construct(int frame_num, int l, const double* relative_R9, const double* relative_T3, int n_tracks,
                           const int* track_ids, const int* track_start, const int* track_nobs, const double* track_xy,
                           double function_tolerance, double* q_wxyz, double* T, int* n_pts, int* pt_ids, double* pts,
                           int* iterations, double* final_cost);
int ve_debug_initial_structure(int F, const double* headers, int n_all, const double* stamps, const int* pts_off,
                               const int* pt_ids, const double* pt_xy, const int* imu_off, const double* imu7,
                               const double* lin6, int n_tracks, const int* track_ids, const int* track_start,
                               const int* track_nobs, const double* track_xy, const double* ric9, const double* tic3,
                               double g_norm, double function_tolerance, double* frame_R, double* frame_T, double* x, double* g3,
                               double* delta_bg3, int* info4, double* bundle_cost);
/* ---- Relocalisation (the only coupling to pose_graph; estimator.h:36, estimator.cpp:769-801, 598-617, 1128-1146) -------------
 *   ve_set_relo_frame       Estimator::setReloFrame(frame_stamp, frame_index, match_points, relo_t, relo_r): match_points = n x
 *                           (x, y, feature id) as decoded from /pose_graph/match_points (estimator_node.cpp:275-290; ascending
 *                           id), relo_r row-major.  When frame_stamp is the stamp of a window frame the NEXT solve carries
 *                           relo_Pose as a 12th pose block with one ProjectionFactor per matched landmark anchored at or before
 *                           that frame.  Returns 1 when the stamp was found (relocalization_info set), 0 otherwise.
 *   ve_get_relocalization   out24 = drift_correct_r 9 (row-major) | drift_correct_t 3 | relo_relative_t 3 | relo_relative_q wxyz |
 *                           relo_relative_yaw (degrees) | relocalization_info still pending | relo_frame_local_index |
 *                           relocalisation factors of the last solve | solves that carried a relocalisation block. */
/* f_manager.feature (feature_manager.h:48-72) as the publishers read it (visualization.cpp:228-296 pubPointCloud, :352-397
 * pubKeyframe): per feature its id, start_frame, solve_flag, estimated_depth and the observations of its consecutive frames
 * (obs5 = point x y z | uv) through obs_offset (cap_features + 1 entries).  Returns the number of features; call with both
 * capacities 0 for that count alone; -(count) - 1 when a capacity is too small (an observation capacity of
 * count x (WINDOW_SIZE + 1) always suffices). */
int ve_get_features(const ve_estimator* e, int cap_features, int cap_obs, int* feature_id, int* start_frame, int* solve_flag,
                    double* estimated_depth, int* obs_offset, double* obs5);
/* Headers[0 .. WINDOW_SIZE] as stamps in seconds (estimator.h:70): what setReloFrame compares frame_stamp with. */
int ve_get_headers(const ve_estimator* e, double* stamps);
int ve_set_relo_frame(ve_estimator* e, double frame_stamp, int frame_index, int n, const double* match_points, const double* relo_t,
                      const double* relo_r);
int ve_get_relocalization(const ve_estimator* e, double* out24);

/* 1 when the handle's last transition to NON_LINEAR came from its own initialisation (0: from a seed); result8 (may be
 * NULL) = l, scale, g (3, after the yaw alignment), bundle iterations, bundle cost, failed attempts so far. */
int ve_init_info(const ve_estimator* e, double* result8);

/* Solver internals of the last solve (profiling/tests): out[0] linear-solver retries, [1] mu, [2] radius,
 * [3..10] per-phase cycle counters of the step kernel summed over the iterations, [11] cycles of the tridiagonalisation of
 * the last marginalisation's A', [12] eigenpairs its eps floor separated explicitly (-1: the full decomposition ran),
 * [13..17] phase cycle counters of the marginalisation solve. */
int ve_solver_debug(const ve_estimator* e, double* out18);

#ifdef __cplusplus
}
#endif
#endif

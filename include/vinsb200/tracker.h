/* vinsb200 front end — C ABI of the feature-tracker hot path.
 *
 * Replaces, one entry point per reference interface:
 *   vt_create / vt_destroy        FeatureTracker::FeatureTracker + readIntrinsicParameter
 *                                 (feature_tracker/src/feature_tracker.h:28-41, feature_tracker.cpp:216-220)
 *                                 + readParameters (feature_tracker/src/parameters.cpp:37-74)
 *   vt_read_image(_device)        FeatureTracker::readImage(const cv::Mat&, double) (feature_tracker.h:33,
 *                                 feature_tracker.cpp:81-167) with PUB_THIS_FRAME passed explicitly
 *                                 (global at feature_tracker.cpp:130) and the node's updateID loop
 *                                 (feature_tracker_node.cpp:103-111) applied before returning
 *   vt_count / vt_get             the public result vectors cur_pts, cur_un_pts, pts_velocity, ids, track_cnt
 *                                 (feature_tracker.h:49-62)
 *   vt_node_image                 img_callback's gating (first frame, stream discontinuity -> restart,
 *                                 frequency control) around readImage (feature_tracker_node.cpp:28-111)
 *   vt_node_pack                  the sensor_msgs/PointCloud payload of img_callback
 *                                 (feature_tracker_node.cpp:113-165): only track_cnt > 1 points
 *
 * Conventions: plain C, no exceptions, int status (0 = ok, < 0 = vt_status error).  A handle is
 * externally synchronised (one call at a time) and owns a CUDA stream plus all its device memory;
 * different handles may be driven from different threads.  There is no CPU fallback: vt_create fails
 * with VT_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef VINSB200_TRACKER_H
#define VINSB200_TRACKER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vt_status {
    VT_OK = 0,
    VT_ERR_INVALID = -1,    /* bad argument / unsupported configuration */
    VT_ERR_NO_DEVICE = -2,  /* no usable CUDA device */
    VT_ERR_CUDA = -3,       /* CUDA runtime error (vt_last_error has the text) */
    VT_ERR_CAPACITY = -4    /* more features than the handle was created for */
} vt_status;

/* camera_model/src/camera_models/{PinholeCamera,CataCamera,EquidistantCamera}.cc: model_type PINHOLE, MEI, KANNALA_BRANDT */
typedef enum vt_camera_model { VT_CAMERA_PINHOLE = 0, VT_CAMERA_MEI = 1, VT_CAMERA_KANNALA_BRANDT = 2 } vt_camera_model;

typedef struct vt_config {
    int rows, cols;        /* image_height, image_width */
    int max_cnt;           /* max_cnt  (euroc_config.yaml:45) */
    int min_dist;          /* min_dist (:46) */
    int freq;              /* freq (:47), 0 -> 100 as in parameters.cpp:68-69 */
    int equalize;          /* equalize (:50) */
    int fisheye;           /* fisheye (:51): start setMask from fisheye_mask instead of 255 */
    int focal_length;      /* FOCAL_LENGTH = 460 (parameters.cpp:65) */
    double f_threshold;    /* F_threshold (:48) */
    int camera_model;      /* vt_camera_model */
    double intrinsics[8];  /* PINHOLE: fx fy cx cy k1 k2 p1 p2; MEI: gamma1 gamma2 u0 v0 k1 k2 p1 p2 (+ xi below);
                              KANNALA_BRANDT: mu mv u0 v0 k2 k3 k4 k5 */
    const uint8_t* fisheye_mask; /* rows*cols bytes, only read when fisheye != 0 */
    int device;            /* CUDA device ordinal */
    double xi;             /* MEI mirror parameter (mirror_parameters.xi); ignored by the other models */
} vt_config;

typedef struct vt_tracker vt_tracker;
typedef struct vt_batch vt_batch;

int vt_create(const vt_config* cfg, vt_tracker** out);
void vt_destroy(vt_tracker* t);
const char* vt_last_error(const vt_tracker* t);

/* img: host pointer, row_stride bytes between rows. */
int vt_read_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double cur_time, int pub_this_frame);
/* Same with the frame already resident in device memory (pointer valid on the handle's device). */
int vt_read_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double cur_time, int pub_this_frame);

/* Number of features currently held (cur_pts.size()). */
int vt_count(const vt_tracker* t);
/* Copies n = vt_count() entries; any pointer may be NULL.  Points are (x, y) pairs. */
int vt_get(const vt_tracker* t, int* ids, int* track_cnt, float* cur_pts, float* cur_un_pts, float* pts_velocity);

/* img_callback: returns 0 (frame consumed, nothing tracked), 1 (tracked, nothing to publish) or
 * 2 (tracked and a feature message is due) or < 0 on error; *restart is set to 1 when the
 * discontinuity rule fired (the node publishes /feature_tracker/restart). */
int vt_node_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double stamp, int* restart);
/* Same with the frame already resident in device memory. */
int vt_node_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double stamp, int* restart);
/* Feature message of the last vt_node_image that returned 2: per point x_un, y_un (z = 1),
 * channel values id*NUM_OF_CAM+cam (NUM_OF_CAM = 1), u, v, velocity_x, velocity_y.  Returns the point count. */
int vt_node_pack(const vt_tracker* t, int capacity, float* xy_un, float* id_of_point, float* u_of_point,
                 float* v_of_point, float* velocity_x, float* velocity_y);

/* ---- Batched sequences (SURVEY.md 8b "Threading", BASELINE configs[2] / [4]) --------------------------------------
 * n trackers with one configuration that advance image by image TOGETHER: per image step one host-to-device copy, one
 * launch per kernel stage (CLAHE, pyramid, LK, mask, Shi-Tomasi, candidate sort, selection; the member index is the last grid
 * dimension) and one device-to-host copy for the whole batch; the members' host bookkeeping (culls, F-RANSAC, setMask,
 * undistortion, ids) runs on a thread pool.  Within a member the serial chain prev/cur/forw of the reference
 * (feature_tracker.cpp:160-164) is kept.  A stand-alone handle from vt_create is a batch of one running the same code.
 *
 *   vt_batch_member       borrowed handle of member k for vt_count / vt_get / vt_node_pack / vt_last_traffic (image calls and
 *                         vt_destroy go through the batch)
 *   vt_batch_read_image   FeatureTracker::readImage for every member k with active[k] != 0 (NULL = all); imgs[k] host or device
 *                         pointers (images_on_device), all with the same row_stride
 *   vt_batch_node_image   img_callback for every active member; results[k] = 0 / 1 / 2 as vt_node_image, restarts[k] may be NULL
 */
int vt_batch_create(const vt_config* cfg, int n, vt_batch** out);
void vt_batch_destroy(vt_batch* b);
int vt_batch_size(const vt_batch* b);
vt_tracker* vt_batch_member(vt_batch* b, int k);
const char* vt_batch_last_error(const vt_batch* b);
int vt_batch_read_image(vt_batch* b, const int* active, const uint8_t* const* imgs, size_t row_stride, const double* cur_times,
                        const int* pub_this_frame, int images_on_device);
int vt_batch_node_image(vt_batch* b, const int* active, const uint8_t* const* imgs, size_t row_stride, const double* stamps,
                        int images_on_device, int* results, int* restarts);
int vt_batch_last_timing(const vt_batch* b, float* device_ms, int* kernel_launches);
int vt_batch_set_profile(vt_batch* b, int on);
int vt_batch_kernel_times(const vt_batch* b, double* ms6, int* count6);

/* Device time (ms, CUDA events on the handle's stream) and launch count of the last read_image. */
int vt_last_timing(const vt_tracker* t, float* device_ms, int* kernel_launches);

/* Kernel profiling (serialises the pipeline; not for timed runs): accumulated device ms / launch counts per group:
 * 0 clahe (lut + apply), 1 pyrdown x3, 2 lk_track, 3 mask_discs, 4 min_eig, 5 candidates + sort + select. */
int vt_set_profile(vt_tracker* t, int on);
int vt_kernel_times(const vt_tracker* t, double* ms6, int* count6);
/* Host<->device bytes moved by the last vt_read_image. */
int vt_last_traffic(const vt_tracker* t, double* h2d_bytes, double* d2h_bytes);

/* Test/benchmark access to single stages on device-resident data (all arrays are host pointers). */
int vt_debug_equalized(vt_tracker* t, int level, uint8_t* out, int* rows, int* cols);
int vt_debug_gftt(vt_tracker* t, const uint8_t* img, size_t row_stride, const uint8_t* mask, int max_corners,
                  float* corners, int* n_candidates, float* eig_out);
int vt_debug_lk(vt_tracker* t, const uint8_t* prev, const uint8_t* next, size_t row_stride, const float* pts, int n,
                float* next_pts, uint8_t* status);
/* The host side of rejectWithF (feature_tracker.cpp:191: cv::findFundamentalMat(un_cur, un_forw, FM_RANSAC, F_THRESHOLD,
 * 0.99, status)) on n correspondences of (x, y) floats; needs no handle and no device.  Returns 1 when a model was
 * found (status = inlier mask), 0 otherwise (status all zero). */
int vt_debug_fundamental_ransac(const float* pts1, const float* pts2, int n, double threshold, double confidence, uint8_t* status);
/* PinholeCamera::liftProjective (PinholeCamera.cc:450-510) as undistortedPoints() applies it: n pixel pairs -> n normalised
 * (x, y) pairs; intrinsics8 = fx fy cx cy k1 k2 p1 p2.  Host only. */
/* The same call returning the winning model as well (row-major 3x3; OpenCV does not refit it): what the estimator's
 * initialisation hands to recoverPose / decomposeE (solve_5pts.cpp:205, initial_ex_rotation.cpp:79). */
int vt_debug_fundamental_ransac_model(const float* pts1, const float* pts2, int n, double threshold, double confidence, uint8_t* status,
                                      double* F9);
int vt_debug_lift_projective(const double* intrinsics8, const double* px, int n, double* out_xy);
/* The same for any vt_camera_model: liftProjective followed by the division by z (what undistortedPoints / rejectWithF use). */
int vt_debug_lift_projective_model(int camera_model, const double* intrinsics8, double xi, const double* px, int n, double* out_xy);
/* Half widths per row offset |dy| = 0..radius of the filled cv::circle that setMask() draws (feature_tracker.cpp:66): the
 * table the mask kernel rasterises discs from.  out has radius + 1 entries.  Host only. */
int vt_debug_disc_half_widths(int radius, int* out);

#ifdef __cplusplus
}
#endif
#endif

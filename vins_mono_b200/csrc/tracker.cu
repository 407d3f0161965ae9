// vt_* C ABI: the feature-tracker pipeline (FeatureTracker::readImage and img_callback of the reference,
// feature_tracker/src/feature_tracker.cpp:81-306, feature_tracker_node.cpp:28-165) driven from the host with
// every image-sized operation on the GPU.
//
// Per image:   H2D frame -> clahe (2 kernels) -> pyramid (3 kernels) -> LK (1 kernel, all points, all levels,
//              inBorder cull fused) -> D2H points+status
// publish frames additionally:  host F-RANSAC + setMask bookkeeping (<= max_cnt points) -> mask discs ->
//              Shi-Tomasi map + masked max -> candidates -> sort -> greedy select -> D2H new corners
// Images, pyramids, the mask and all scratch stay resident in HBM; only O(max_cnt) point data crosses PCIe.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "fe_kernels.h"
#include "fm_ransac.h"
#include "kprof.h"
#include "vinsb200/tracker.h"

namespace {

struct Pt {
    float x, y;
};

struct DevicePyramid {
    uint8_t* base = nullptr;
    vb::PyramidView view{};
};

inline int cv_round(float v) { return (int)lrintf(v); }

}  // namespace

struct vt_tracker {
    vt_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // device memory
    uint8_t* d_raw = nullptr;
    int raw_pitch = 0;
    uint8_t* d_lut = nullptr;
    DevicePyramid pyr[2];
    int cur = 0;  // pyr[cur] = cur_img pyramid, pyr[cur^1] = forw_img pyramid
    uint8_t* d_mask = nullptr;
    uint8_t* d_fisheye = nullptr;
    float* d_eig = nullptr;
    unsigned long long* d_keys = nullptr;
    int key_capacity = 0;
    int* d_count = nullptr;         // [0] candidate count, [1] selected count
    unsigned* d_max = nullptr;
    int* d_cell_cnt = nullptr;
    short2* d_cell_pts = nullptr;
    float* d_pts_in = nullptr;      // LK input points
    float* d_pts_out = nullptr;     // LK output / new corners
    uint8_t* d_status = nullptr;
    int* d_centres = nullptr;
    int* d_halfw = nullptr;
    // pinned host staging
    float* h_pts = nullptr;
    uint8_t* h_status = nullptr;
    int* h_centres = nullptr;
    int* h_counts = nullptr;
    uint8_t* h_img = nullptr;
    std::vector<int> halfw;
    int capacity = 0;
    // FeatureTracker state (feature_tracker.h:45-64)
    bool have_img = false;
    std::vector<Pt> cur_pts, forw_pts, cur_un_pts, pts_velocity, n_pts;
    std::vector<int> ids, track_cnt;
    std::map<int, Pt> cur_un_pts_map, prev_un_pts_map;
    double cur_time = 0, prev_time = 0;
    int n_id = 0;
    // img_callback state (feature_tracker_node.cpp:21-26)
    double first_image_time = 0, last_image_time = 0;
    int pub_count = 1;
    bool first_image_flag = true, init_pub = false;
    // diagnostics
    float last_ms = 0;
    int last_launches = 0;
    vb::KernelProfile prof;  // 0 clahe_lut+apply, 1 pyrdown x3, 2 lk, 3 mask, 4 min_eig, 5 candidates+sort+select
    size_t h2d_bytes = 0, d2h_bytes = 0;
    double host_ms_ransac = 0, host_ms_mask = 0;
};

namespace {

#define VT_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            t->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return VT_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

int align_up(int v, int a) { return (v + a - 1) / a * a; }

int lk_levels(int rows, int cols, int win, int max_level) {
    int lv = 0, r = rows, c = cols;
    while (lv < max_level) {
        const int nr = (r + 1) / 2, nc = (c + 1) / 2;
        if (nc <= win || nr <= win) break;
        r = nr;
        c = nc;
        lv++;
    }
    return lv;
}

int alloc_pyramid(vt_tracker* t, DevicePyramid& p) {
    const int nlev = lk_levels(t->cfg.rows, t->cfg.cols, 21, 3);
    size_t total = 0, offs[vb::MAX_PYR_LEVELS];
    int r = t->cfg.rows, c = t->cfg.cols;
    for (int l = 0; l <= nlev; l++) {
        p.view.rows[l] = r;
        p.view.cols[l] = c;
        p.view.pitch[l] = align_up(c, 64);
        offs[l] = total;
        total += (size_t)align_up(p.view.pitch[l] * r, 256);
        r = (r + 1) / 2;
        c = (c + 1) / 2;
    }
    p.view.nlev = nlev;
    VT_CUDA(cudaMalloc(&p.base, total));
    for (int l = 0; l <= nlev; l++) p.view.img[l] = p.base + offs[l];
    for (int l = nlev + 1; l < vb::MAX_PYR_LEVELS; l++) p.view.img[l] = nullptr;
    return VT_OK;
}

// Half widths of cv::circle(..., FILLED) per row offset (midpoint algorithm of OpenCV's Circle()).
std::vector<int> disc_half_widths(int radius) {
    std::vector<int> hw(radius + 1, -1);
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        hw[dy] = std::max(hw[dy], dx);  // rows cy +- dy span cx +- dx
        hw[dx] = std::max(hw[dx], dy);  // rows cy +- dx span cx +- dy
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
    return hw;
}

// PinholeCamera::liftProjective (camera_model/src/camera_models/PinholeCamera.cc:450-510)
void lift_projective(const vt_config& c, double px, double py, double& X, double& Y) {
    const double fx = c.intrinsics[0], fy = c.intrinsics[1], cx = c.intrinsics[2], cy = c.intrinsics[3];
    const double k1 = c.intrinsics[4], k2 = c.intrinsics[5], p1 = c.intrinsics[6], p2 = c.intrinsics[7];
    const double mx_d = (1.0 / fx) * px + (-cx / fx), my_d = (1.0 / fy) * py + (-cy / fy);
    if (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0) {
        X = mx_d;
        Y = my_d;
        return;
    }
    double ux = mx_d, uy = my_d;
    for (int i = 0; i < 8; ++i) {
        const double x2 = ux * ux, y2 = uy * uy, xy = ux * uy, rho2 = x2 + y2;
        const double rad = k1 * rho2 + k2 * rho2 * rho2;
        const double ddx = ux * rad + 2.0 * p1 * xy + p2 * (rho2 + 2.0 * x2);
        const double ddy = uy * rad + 2.0 * p2 * xy + p1 * (rho2 + 2.0 * y2);
        ux = mx_d - ddx;
        uy = my_d - ddy;
    }
    X = ux;
    Y = uy;
}

template <class T>
void compact(std::vector<T>& v, const uint8_t* keep) {
    size_t j = 0;
    for (size_t i = 0; i < v.size(); i++)
        if (keep[i]) v[j++] = v[i];
    v.resize(j);
}

// rejectWithF (feature_tracker.cpp:169-202)
void reject_with_f(vt_tracker* t) {
    const int n = (int)t->forw_pts.size();
    if (n < 8) return;
    std::vector<float> a(2 * n), b(2 * n);
    const double f = t->cfg.focal_length, cx = t->cfg.cols / 2.0, cy = t->cfg.rows / 2.0;
    for (int i = 0; i < n; i++) {
        double X, Y;
        lift_projective(t->cfg, t->cur_pts[i].x, t->cur_pts[i].y, X, Y);
        a[2 * i] = (float)(f * X / 1.0 + cx);
        a[2 * i + 1] = (float)(f * Y / 1.0 + cy);
        lift_projective(t->cfg, t->forw_pts[i].x, t->forw_pts[i].y, X, Y);
        b[2 * i] = (float)(f * X / 1.0 + cx);
        b[2 * i + 1] = (float)(f * Y / 1.0 + cy);
    }
    std::vector<uint8_t> status(n);
    vb::fundamental_ransac_mask(a.data(), b.data(), n, t->cfg.f_threshold, 0.99, status.data());
    compact(t->cur_pts, status.data());
    compact(t->forw_pts, status.data());
    compact(t->cur_un_pts, status.data());
    compact(t->ids, status.data());
    compact(t->track_cnt, status.data());
}

// setMask (feature_tracker.cpp:36-69): same std::sort call (unstable, libstdc++ introsort) on the same
// element type layout; the "is the mask still 255 here" test is evaluated against the discs already
// drawn instead of a rasterised image.  Produces the integer disc centres for the device mask.
int set_mask(vt_tracker* t) {
    std::vector<std::pair<int, std::pair<Pt, int>>> cnt_pts_id;
    for (size_t i = 0; i < t->forw_pts.size(); i++)
        cnt_pts_id.push_back(std::make_pair(t->track_cnt[i], std::make_pair(t->forw_pts[i], t->ids[i])));
    std::sort(cnt_pts_id.begin(), cnt_pts_id.end(),
              [](const std::pair<int, std::pair<Pt, int>>& a, const std::pair<int, std::pair<Pt, int>>& b) {
                  return a.first > b.first;
              });
    t->forw_pts.clear();
    t->ids.clear();
    t->track_cnt.clear();
    const int r = t->cfg.min_dist;
    int kept = 0;
    for (auto& it : cnt_pts_id) {
        const int px = cv_round(it.second.first.x), py = cv_round(it.second.first.y);
        bool free_px = true;
        if (t->cfg.fisheye && t->cfg.fisheye_mask) free_px = t->cfg.fisheye_mask[(size_t)py * t->cfg.cols + px] == 255;
        for (int k = 0; k < kept && free_px; k++) {
            const int dy = std::abs(py - t->h_centres[2 * k + 1]), dx = std::abs(px - t->h_centres[2 * k]);
            if (dy <= r && dx <= t->halfw[dy]) free_px = false;
        }
        if (free_px) {
            t->forw_pts.push_back(it.second.first);
            t->ids.push_back(it.second.second);
            t->track_cnt.push_back(it.first);
            t->h_centres[2 * kept] = px;
            t->h_centres[2 * kept + 1] = py;
            kept++;
        }
    }
    return kept;
}

// undistortedPoints (feature_tracker.cpp:258-306)
void undistorted_points(vt_tracker* t) {
    t->cur_un_pts.clear();
    t->cur_un_pts_map.clear();
    for (size_t i = 0; i < t->cur_pts.size(); i++) {
        double X, Y;
        lift_projective(t->cfg, t->cur_pts[i].x, t->cur_pts[i].y, X, Y);
        const Pt u{(float)(X / 1.0), (float)(Y / 1.0)};
        t->cur_un_pts.push_back(u);
        t->cur_un_pts_map.insert(std::make_pair(t->ids[i], u));
    }
    if (!t->prev_un_pts_map.empty()) {
        const double dt = t->cur_time - t->prev_time;
        t->pts_velocity.clear();
        for (size_t i = 0; i < t->cur_un_pts.size(); i++) {
            Pt v{0, 0};
            if (t->ids[i] != -1) {
                auto it = t->prev_un_pts_map.find(t->ids[i]);
                if (it != t->prev_un_pts_map.end())
                    v = Pt{(float)((t->cur_un_pts[i].x - it->second.x) / dt), (float)((t->cur_un_pts[i].y - it->second.y) / dt)};
            }
            t->pts_velocity.push_back(v);
        }
    } else {
        for (size_t i = 0; i < t->cur_pts.size(); i++) t->pts_velocity.push_back(Pt{0, 0});
    }
    t->prev_un_pts_map = t->cur_un_pts_map;
}

// Builds the forw pyramid in pyr[cur^1] from the frame in d_raw.
int build_forw_pyramid(vt_tracker* t) {
    DevicePyramid& f = t->pyr[t->cur ^ 1];
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    uint8_t* l0 = const_cast<uint8_t*>(f.view.img[0]);
    if (t->cfg.equalize) {
        t->prof.begin(t->stream);
        vb::launch_clahe(t->d_raw, rows, cols, t->raw_pitch, t->d_lut, l0, f.view.pitch[0], t->stream);
        t->prof.end(0, t->stream, 2);
        t->last_launches += 2;
    } else {
        VT_CUDA(cudaMemcpy2DAsync(l0, f.view.pitch[0], t->d_raw, t->raw_pitch, cols, rows, cudaMemcpyDeviceToDevice,
                                  t->stream));
    }
    t->prof.begin(t->stream);
    for (int l = 1; l <= f.view.nlev; l++) {
        vb::launch_pyrdown(f.view.img[l - 1], f.view.rows[l - 1], f.view.cols[l - 1], f.view.pitch[l - 1],
                           const_cast<uint8_t*>(f.view.img[l]), f.view.pitch[l], t->stream);
        t->last_launches++;
    }
    t->prof.end(1, t->stream, f.view.nlev);
    return VT_OK;
}

int detect_new(vt_tracker* t, const uint8_t* d_img, int pitch, const uint8_t* d_mask, int max_corners, float* h_out,
               int* n_out, int* n_cand) {
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    VT_CUDA(cudaMemsetAsync(t->d_count, 0, 2 * sizeof(int), t->stream));
    VT_CUDA(cudaMemsetAsync(t->d_max, 0, sizeof(unsigned), t->stream));
    t->prof.begin(t->stream);
    vb::launch_min_eig(d_img, rows, cols, pitch, d_mask, cols, t->d_eig, cols, t->d_max, t->stream);
    t->prof.end(4, t->stream);
    t->prof.begin(t->stream);
    vb::launch_gftt_tail(t->d_eig, rows, cols, cols, d_mask, cols, t->d_max, 0.01, t->d_keys, t->key_capacity,
                         t->d_count, max_corners, (float)t->cfg.min_dist, t->d_cell_cnt, t->d_cell_pts, t->d_pts_out,
                         t->d_count + 1, t->stream);
    t->prof.end(5, t->stream, 3);
    t->last_launches += 4;
    t->d2h_bytes += 2 * sizeof(int) + (size_t)max_corners * 2 * sizeof(float);
    VT_CUDA(cudaMemcpyAsync(t->h_counts, t->d_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, t->stream));
    VT_CUDA(cudaMemcpyAsync(h_out, t->d_pts_out, (size_t)max_corners * 2 * sizeof(float), cudaMemcpyDeviceToHost,
                            t->stream));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    if (t->h_counts[0] > t->key_capacity) {
        t->err = "Shi-Tomasi candidate buffer overflow";
        return VT_ERR_CAPACITY;
    }
    *n_out = t->h_counts[1];
    if (n_cand) *n_cand = t->h_counts[0];
    return VT_OK;
}

int read_image_impl(vt_tracker* t, double cur_time, bool pub) {
    // precondition: the frame is in t->d_raw (enqueued on t->stream)
    t->cur_time = cur_time;
    int rc = build_forw_pyramid(t);
    if (rc) return rc;
    const bool first = !t->have_img;
    t->have_img = true;
    if (first) t->cur = t->cur ^ 1;  // prev = cur = forw = img: the just-built pyramid is also cur
    const DevicePyramid& curp = t->pyr[t->cur];
    const DevicePyramid& forwp = first ? t->pyr[t->cur] : t->pyr[t->cur ^ 1];
    t->forw_pts.clear();
    if (!t->cur_pts.empty()) {
        const int n = (int)t->cur_pts.size();
        std::memcpy(t->h_pts, t->cur_pts.data(), (size_t)n * sizeof(Pt));
        VT_CUDA(cudaMemcpyAsync(t->d_pts_in, t->h_pts, (size_t)n * sizeof(Pt), cudaMemcpyHostToDevice, t->stream));
        t->prof.begin(t->stream);
        vb::launch_lk(curp.view, forwp.view, t->d_pts_in, n, t->d_pts_out, t->d_status, t->stream);
        t->prof.end(2, t->stream);
        t->last_launches++;
        t->h2d_bytes += (size_t)n * sizeof(Pt);
        t->d2h_bytes += (size_t)n * (sizeof(Pt) + 1);
        VT_CUDA(cudaMemcpyAsync(t->h_pts, t->d_pts_out, (size_t)n * sizeof(Pt), cudaMemcpyDeviceToHost, t->stream));
        VT_CUDA(cudaMemcpyAsync(t->h_status, t->d_status, n, cudaMemcpyDeviceToHost, t->stream));
        VT_CUDA(cudaStreamSynchronize(t->stream));
        t->forw_pts.assign(reinterpret_cast<Pt*>(t->h_pts), reinterpret_cast<Pt*>(t->h_pts) + n);
        // status already includes the inBorder() cull (fused into the LK kernel)
        compact(t->cur_pts, t->h_status);
        compact(t->forw_pts, t->h_status);
        compact(t->ids, t->h_status);
        compact(t->cur_un_pts, t->h_status);
        compact(t->track_cnt, t->h_status);
    }
    for (auto& c : t->track_cnt) c++;
    if (pub) {
        reject_with_f(t);
        const int kept = set_mask(t);
        const int n_max_cnt = t->cfg.max_cnt - (int)t->forw_pts.size();
        t->n_pts.clear();
        if (n_max_cnt > 0) {
            const int rows = t->cfg.rows, cols = t->cfg.cols;
            if (t->cfg.fisheye)
                VT_CUDA(cudaMemcpyAsync(t->d_mask, t->d_fisheye, (size_t)rows * cols, cudaMemcpyDeviceToDevice, t->stream));
            else
                VT_CUDA(cudaMemsetAsync(t->d_mask, 255, (size_t)rows * cols, t->stream));
            if (kept > 0) {
                VT_CUDA(cudaMemcpyAsync(t->d_centres, t->h_centres, (size_t)kept * 2 * sizeof(int),
                                        cudaMemcpyHostToDevice, t->stream));
                t->prof.begin(t->stream);
                vb::launch_mask_discs(t->d_mask, rows, cols, cols, t->d_centres, kept, t->cfg.min_dist, t->d_halfw,
                                      t->stream);
                t->prof.end(3, t->stream);
                t->last_launches++;
                t->h2d_bytes += (size_t)kept * 2 * sizeof(int);
            }
            int n_new = 0;
            rc = detect_new(t, forwp.view.img[0], forwp.view.pitch[0], t->d_mask, n_max_cnt, t->h_pts, &n_new, nullptr);
            if (rc) return rc;
            for (int i = 0; i < n_new; i++) t->n_pts.push_back(Pt{t->h_pts[2 * i], t->h_pts[2 * i + 1]});
        }
        for (auto& p : t->n_pts) {  // addPoints (feature_tracker.cpp:71-79)
            t->forw_pts.push_back(p);
            t->ids.push_back(-1);
            t->track_cnt.push_back(1);
        }
    }
    if (!first) t->cur ^= 1;  // cur_img = forw_img (prev_img is never read again by the hot path)
    t->cur_pts = t->forw_pts;
    undistorted_points(t);
    t->prev_time = t->cur_time;
    for (size_t i = 0; i < t->ids.size(); i++)  // updateID loop of img_callback (feature_tracker_node.cpp:103-111)
        if (t->ids[i] == -1) t->ids[i] = t->n_id++;
    return VT_OK;
}

int upload_frame(vt_tracker* t, const uint8_t* img, size_t stride, bool on_device) {
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    if (on_device) {
        VT_CUDA(cudaMemcpy2DAsync(t->d_raw, t->raw_pitch, img, stride, cols, rows, cudaMemcpyDeviceToDevice, t->stream));
    } else {
        // stage through pinned memory so the copy is a true async DMA
        for (int y = 0; y < rows; y++) std::memcpy(t->h_img + (size_t)y * cols, img + (size_t)y * stride, cols);
        VT_CUDA(cudaMemcpy2DAsync(t->d_raw, t->raw_pitch, t->h_img, cols, cols, rows, cudaMemcpyHostToDevice, t->stream));
    }
    return VT_OK;
}

int read_image_common(vt_tracker* t, const uint8_t* img, size_t stride, double cur_time, int pub, bool on_device) {
    if (!t || !img || stride < (size_t)t->cfg.cols) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    t->last_launches = 0;
    t->h2d_bytes = on_device ? 0 : (size_t)t->cfg.rows * t->cfg.cols;
    t->d2h_bytes = 0;
    VT_CUDA(cudaEventRecord(t->ev0, t->stream));
    int rc = upload_frame(t, img, stride, on_device);
    if (rc) return rc;
    rc = read_image_impl(t, cur_time, pub != 0);
    if (rc) return rc;
    VT_CUDA(cudaEventRecord(t->ev1, t->stream));
    VT_CUDA(cudaEventSynchronize(t->ev1));
    VT_CUDA(cudaEventElapsedTime(&t->last_ms, t->ev0, t->ev1));
    return VT_OK;
}

}  // namespace

extern "C" {

int vt_create(const vt_config* cfg, vt_tracker** out) {
    if (!cfg || !out) return VT_ERR_INVALID;
    *out = nullptr;
    if (cfg->rows < 32 || cfg->cols < 32 || cfg->max_cnt <= 0 || cfg->min_dist < 1 ||
        cfg->camera_model != VT_CAMERA_PINHOLE)
        return VT_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
        return VT_ERR_NO_DEVICE;
    vt_tracker* t = new vt_tracker();
    t->cfg = *cfg;
    if (t->cfg.freq == 0) t->cfg.freq = 100;
    if (t->cfg.focal_length == 0) t->cfg.focal_length = 460;
    auto fail = [&](int code) {
        vt_destroy(t);
        return code;
    };
#define VT_TRY(call)                                   \
    do {                                               \
        if ((call) != cudaSuccess) return fail(VT_ERR_CUDA); \
    } while (0)
    VT_TRY(cudaSetDevice(cfg->device));
    VT_TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    VT_TRY(cudaEventCreate(&t->ev0));
    VT_TRY(cudaEventCreate(&t->ev1));
    const int rows = cfg->rows, cols = cfg->cols;
    const size_t npx = (size_t)rows * cols;
    t->raw_pitch = align_up(cols, 64);
    t->capacity = std::max(cfg->max_cnt, 16);
    VT_TRY(cudaMalloc(&t->d_raw, (size_t)t->raw_pitch * rows));
    VT_TRY(cudaMalloc(&t->d_lut, 64 * 256));
    if (alloc_pyramid(t, t->pyr[0]) || alloc_pyramid(t, t->pyr[1])) return fail(VT_ERR_CUDA);
    VT_TRY(cudaMalloc(&t->d_mask, npx));
    VT_TRY(cudaMalloc(&t->d_eig, npx * sizeof(float)));
    t->key_capacity = 1;
    while ((size_t)t->key_capacity < npx / 8) t->key_capacity <<= 1;  // local maxima of a 3x3 NMS: < npx/4 in practice
    VT_TRY(cudaMalloc(&t->d_keys, (size_t)t->key_capacity * sizeof(unsigned long long)));
    VT_TRY(cudaMalloc(&t->d_count, 2 * sizeof(int)));
    VT_TRY(cudaMalloc(&t->d_max, sizeof(unsigned)));
    const int cell = cfg->min_dist, gw = (cols + cell - 1) / cell, gh = (rows + cell - 1) / cell;
    VT_TRY(cudaMalloc(&t->d_cell_cnt, (size_t)gw * gh * sizeof(int)));
    VT_TRY(cudaMalloc(&t->d_cell_pts, (size_t)gw * gh * 8 * sizeof(short2)));
    VT_TRY(cudaMalloc(&t->d_pts_in, (size_t)t->capacity * 2 * sizeof(float)));
    VT_TRY(cudaMalloc(&t->d_pts_out, (size_t)t->capacity * 2 * sizeof(float)));
    VT_TRY(cudaMalloc(&t->d_status, t->capacity));
    VT_TRY(cudaMalloc(&t->d_centres, (size_t)t->capacity * 2 * sizeof(int)));
    t->halfw = disc_half_widths(cfg->min_dist);
    VT_TRY(cudaMalloc(&t->d_halfw, t->halfw.size() * sizeof(int)));
    VT_TRY(cudaMemcpy(t->d_halfw, t->halfw.data(), t->halfw.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (cfg->fisheye) {
        if (!cfg->fisheye_mask) return fail(VT_ERR_INVALID);
        VT_TRY(cudaMalloc(&t->d_fisheye, npx));
        VT_TRY(cudaMemcpy(t->d_fisheye, cfg->fisheye_mask, npx, cudaMemcpyHostToDevice));
        uint8_t* keep = new uint8_t[npx];
        std::memcpy(keep, cfg->fisheye_mask, npx);
        t->cfg.fisheye_mask = keep;  // own a copy: setMask reads it on the host
    }
    VT_TRY(cudaHostAlloc(&t->h_pts, (size_t)t->capacity * 2 * sizeof(float), cudaHostAllocDefault));
    VT_TRY(cudaHostAlloc(&t->h_status, t->capacity, cudaHostAllocDefault));
    VT_TRY(cudaHostAlloc(&t->h_centres, (size_t)t->capacity * 2 * sizeof(int), cudaHostAllocDefault));
    VT_TRY(cudaHostAlloc(&t->h_counts, 2 * sizeof(int), cudaHostAllocDefault));
    VT_TRY(cudaHostAlloc(&t->h_img, npx, cudaHostAllocDefault));
#undef VT_TRY
    *out = t;
    return VT_OK;
}

void vt_destroy(vt_tracker* t) {
    if (!t) return;
    cudaSetDevice(t->cfg.device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    cudaFree(t->d_raw);
    cudaFree(t->d_lut);
    cudaFree(t->pyr[0].base);
    cudaFree(t->pyr[1].base);
    cudaFree(t->d_mask);
    cudaFree(t->d_fisheye);
    cudaFree(t->d_eig);
    cudaFree(t->d_keys);
    cudaFree(t->d_count);
    cudaFree(t->d_max);
    cudaFree(t->d_cell_cnt);
    cudaFree(t->d_cell_pts);
    cudaFree(t->d_pts_in);
    cudaFree(t->d_pts_out);
    cudaFree(t->d_status);
    cudaFree(t->d_centres);
    cudaFree(t->d_halfw);
    cudaFreeHost(t->h_pts);
    cudaFreeHost(t->h_status);
    cudaFreeHost(t->h_centres);
    cudaFreeHost(t->h_counts);
    cudaFreeHost(t->h_img);
    if (t->cfg.fisheye && t->d_fisheye) delete[] t->cfg.fisheye_mask;
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

const char* vt_last_error(const vt_tracker* t) { return t ? t->err.c_str() : "null handle"; }

int vt_read_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double cur_time, int pub_this_frame) {
    return read_image_common(t, img, row_stride, cur_time, pub_this_frame, false);
}

int vt_read_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double cur_time, int pub_this_frame) {
    return read_image_common(t, d_img, row_stride, cur_time, pub_this_frame, true);
}

int vt_count(const vt_tracker* t) { return t ? (int)t->cur_pts.size() : VT_ERR_INVALID; }

int vt_get(const vt_tracker* t, int* ids, int* track_cnt, float* cur_pts, float* cur_un_pts, float* pts_velocity) {
    if (!t) return VT_ERR_INVALID;
    const size_t n = t->cur_pts.size();
    if (ids) std::memcpy(ids, t->ids.data(), n * sizeof(int));
    if (track_cnt) std::memcpy(track_cnt, t->track_cnt.data(), n * sizeof(int));
    if (cur_pts) std::memcpy(cur_pts, t->cur_pts.data(), n * sizeof(Pt));
    if (cur_un_pts) std::memcpy(cur_un_pts, t->cur_un_pts.data(), n * sizeof(Pt));
    if (pts_velocity) std::memcpy(pts_velocity, t->pts_velocity.data(), n * sizeof(Pt));
    return (int)n;
}

static int node_image_common(vt_tracker* t, const uint8_t* img, size_t row_stride, double stamp, int* restart, bool on_device) {
    if (!t) return VT_ERR_INVALID;
    if (restart) *restart = 0;
    if (t->first_image_flag) {
        t->first_image_flag = false;
        t->first_image_time = stamp;
        t->last_image_time = stamp;
        return 0;
    }
    if (stamp - t->last_image_time > 1.0 || stamp < t->last_image_time) {
        t->first_image_flag = true;
        t->last_image_time = 0;
        t->pub_count = 1;
        if (restart) *restart = 1;
        return 0;
    }
    t->last_image_time = stamp;
    bool pub;
    if (std::round(1.0 * t->pub_count / (stamp - t->first_image_time)) <= t->cfg.freq) {
        pub = true;
        if (std::abs(1.0 * t->pub_count / (stamp - t->first_image_time) - t->cfg.freq) < 0.01 * t->cfg.freq) {
            t->first_image_time = stamp;
            t->pub_count = 0;
        }
    } else
        pub = false;
    const int rc = read_image_common(t, img, row_stride, stamp, pub, on_device);
    if (rc) return rc;
    if (pub) {
        t->pub_count++;
        if (!t->init_pub) {
            t->init_pub = true;
            return 1;
        }
        return 2;
    }
    return 1;
}

int vt_node_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double stamp, int* restart) {
    return node_image_common(t, img, row_stride, stamp, restart, false);
}

int vt_node_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double stamp, int* restart) {
    return node_image_common(t, d_img, row_stride, stamp, restart, true);
}

int vt_node_pack(const vt_tracker* t, int capacity, float* xy_un, float* id_of_point, float* u_of_point,
                 float* v_of_point, float* velocity_x, float* velocity_y) {
    if (!t) return VT_ERR_INVALID;
    int k = 0;
    for (size_t j = 0; j < t->ids.size(); j++) {
        if (t->track_cnt[j] <= 1) continue;
        if (k >= capacity) return VT_ERR_CAPACITY;
        xy_un[2 * k] = t->cur_un_pts[j].x;
        xy_un[2 * k + 1] = t->cur_un_pts[j].y;
        id_of_point[k] = (float)(t->ids[j] * 1 + 0);
        u_of_point[k] = t->cur_pts[j].x;
        v_of_point[k] = t->cur_pts[j].y;
        velocity_x[k] = t->pts_velocity[j].x;
        velocity_y[k] = t->pts_velocity[j].y;
        k++;
    }
    return k;
}

int vt_debug_fundamental_ransac(const float* pts1, const float* pts2, int n, double threshold, double confidence, uint8_t* status) {
    if (!pts1 || !pts2 || !status || n < 0) return VT_ERR_INVALID;
    return vb::fundamental_ransac_mask(pts1, pts2, n, threshold, confidence, status) ? 1 : 0;
}

int vt_debug_lift_projective(const double* intrinsics8, const double* px, int n, double* out_xy) {
    if (!intrinsics8 || !px || !out_xy || n < 0) return VT_ERR_INVALID;
    vt_config c{};
    for (int i = 0; i < 8; i++) c.intrinsics[i] = intrinsics8[i];
    for (int k = 0; k < n; k++) lift_projective(c, px[2 * k], px[2 * k + 1], out_xy[2 * k], out_xy[2 * k + 1]);
    return VT_OK;
}

int vt_debug_disc_half_widths(int radius, int* out) {
    if (radius < 0 || !out) return VT_ERR_INVALID;
    const std::vector<int> hw = disc_half_widths(radius);
    for (int d = 0; d <= radius; d++) out[d] = hw[d];
    return VT_OK;
}

int vt_last_timing(const vt_tracker* t, float* device_ms, int* kernel_launches) {
    if (!t) return VT_ERR_INVALID;
    if (device_ms) *device_ms = t->last_ms;
    if (kernel_launches) *kernel_launches = t->last_launches;
    return VT_OK;
}

int vt_set_profile(vt_tracker* t, int on) {
    if (!t) return VT_ERR_INVALID;
    t->prof.enable(on != 0);
    return VT_OK;
}

int vt_kernel_times(const vt_tracker* t, double* ms6, int* count6) {
    if (!t) return VT_ERR_INVALID;
    for (int k = 0; k < 6; k++) {
        if (ms6) ms6[k] = t->prof.ms[k];
        if (count6) count6[k] = t->prof.count[k];
    }
    return VT_OK;
}

int vt_last_traffic(const vt_tracker* t, double* h2d_bytes, double* d2h_bytes) {
    if (!t) return VT_ERR_INVALID;
    if (h2d_bytes) *h2d_bytes = (double)t->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = (double)t->d2h_bytes;
    return VT_OK;
}

int vt_debug_equalized(vt_tracker* t, int level, uint8_t* out, int* rows, int* cols) {
    if (!t || !t->have_img) return VT_ERR_INVALID;
    const vb::PyramidView& v = t->pyr[t->cur].view;
    if (level < 0 || level > v.nlev) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(cudaMemcpy2D(out, v.cols[level], v.img[level], v.pitch[level], v.cols[level], v.rows[level],
                         cudaMemcpyDeviceToHost));
    if (rows) *rows = v.rows[level];
    if (cols) *cols = v.cols[level];
    return VT_OK;
}

int vt_debug_gftt(vt_tracker* t, const uint8_t* img, size_t row_stride, const uint8_t* mask, int max_corners,
                  float* corners, int* n_candidates, float* eig_out) {
    if (!t || !img || max_corners <= 0 || max_corners > t->capacity) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    VT_CUDA(cudaMemcpy2DAsync(t->d_raw, t->raw_pitch, img, row_stride, cols, rows, cudaMemcpyHostToDevice, t->stream));
    if (mask) VT_CUDA(cudaMemcpyAsync(t->d_mask, mask, (size_t)rows * cols, cudaMemcpyHostToDevice, t->stream));
    int n = 0;
    const int rc = detect_new(t, t->d_raw, t->raw_pitch, mask ? t->d_mask : nullptr, max_corners, t->h_pts, &n, n_candidates);
    if (rc) return rc;
    std::memcpy(corners, t->h_pts, (size_t)n * 2 * sizeof(float));
    if (eig_out) VT_CUDA(cudaMemcpy(eig_out, t->d_eig, (size_t)rows * cols * sizeof(float), cudaMemcpyDeviceToHost));
    return n;
}

int vt_debug_lk(vt_tracker* t, const uint8_t* prev, const uint8_t* next, size_t row_stride, const float* pts, int n,
                float* next_pts, uint8_t* status) {
    if (!t || !prev || !next || n < 0 || n > t->capacity) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    for (int k = 0; k < 2; k++) {
        DevicePyramid& p = t->pyr[k];
        VT_CUDA(cudaMemcpy2DAsync(const_cast<uint8_t*>(p.view.img[0]), p.view.pitch[0], k == 0 ? prev : next, row_stride,
                                  cols, rows, cudaMemcpyHostToDevice, t->stream));
        for (int l = 1; l <= p.view.nlev; l++)
            vb::launch_pyrdown(p.view.img[l - 1], p.view.rows[l - 1], p.view.cols[l - 1], p.view.pitch[l - 1],
                               const_cast<uint8_t*>(p.view.img[l]), p.view.pitch[l], t->stream);
    }
    if (n > 0) {
        VT_CUDA(cudaMemcpyAsync(t->d_pts_in, pts, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, t->stream));
        vb::launch_lk(t->pyr[0].view, t->pyr[1].view, t->d_pts_in, n, t->d_pts_out, t->d_status, t->stream);
        VT_CUDA(cudaMemcpyAsync(next_pts, t->d_pts_out, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
        VT_CUDA(cudaMemcpyAsync(status, t->d_status, n, cudaMemcpyDeviceToHost, t->stream));
    }
    VT_CUDA(cudaStreamSynchronize(t->stream));
    t->have_img = false;  // the pyramids no longer belong to the tracking state
    return VT_OK;
}

}  // extern "C"

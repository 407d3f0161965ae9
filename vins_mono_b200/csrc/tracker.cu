// vt_* C ABI: the feature-tracker pipeline (FeatureTracker::readImage and img_callback of the reference,
// feature_tracker/src/feature_tracker.cpp:81-306, feature_tracker_node.cpp:28-165) driven from the host with
// every image-sized operation on the GPU.
//
// Execution model: trackers are members of a batch (vt_batch; a stand-alone handle is a batch of one) that advance
// image by image together.  Per image step:
//   prepare (host, per member)  ->  one H2D copy (descriptors, points[, frames])  ->  CLAHE (2 kernels), pyramid (3), LK (1,
//   all points, all levels, inBorder cull fused) for ALL members, member = last grid dimension  ->  one D2H copy
//   (points + status)  ->  host: compaction; on publishing frames F-RANSAC + setMask bookkeeping (<= max_cnt points)
//   ->  one H2D copy (disc centres)  ->  mask, Shi-Tomasi map + masked max, candidates, sort, greedy select for all
//   publishing members  ->  one D2H copy (new corners)  ->  host: addPoints, undistortedPoints, updateID.
// Images, pyramids, masks and all scratch stay resident in HBM; only O(max_cnt) point data crosses PCIe (plus the frame
// itself when it is given as a host pointer).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fe_kernels.h"
#include "fm_ransac.h"
#include "hostpool.h"
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
    vt_batch* batch = nullptr;
    int member = 0;
    // device memory of this member
    uint8_t* d_lut = nullptr;
    DevicePyramid pyr[2];
    int cur = 0;  // pyr[cur] = cur_img pyramid, pyr[cur^1] = forw_img pyramid
    uint8_t* d_mask = nullptr;
    uint8_t* d_fisheye = nullptr;
    uint8_t* owned_fisheye_mask = nullptr;  // host copy read by setMask
    float* d_eig = nullptr;
    unsigned long long* d_keys = nullptr;
    int key_capacity = 0;
    int* d_cell_cnt = nullptr;
    short2* d_cell_pts = nullptr;
    // this member's slots in the batch arenas
    uint8_t* d_raw = nullptr;  // frame slot (host-image path)
    uint8_t* h_img = nullptr;
    float* h_pts_in = nullptr;   // pinned, up arena
    int* h_centres = nullptr;
    float* h_pts_out = nullptr;  // pinned, down arena
    uint8_t* h_status = nullptr;
    int* h_counts = nullptr;
    float* h_new_pts = nullptr;
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
    // the current image step
    int status = VT_OK;
    bool step_active = false, step_pub = false, step_first = false, step_detect = false;
    int step_n = 0, step_kept = 0, step_max_new = 0;
    size_t h2d_bytes = 0, d2h_bytes = 0;
    double host_ms_ransac = 0, host_ms_mask = 0;
};

struct vt_batch {
    vt_config cfg{};
    int S = 0;
    bool standalone = false;
    std::vector<vt_tracker*> members;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    vb::FeSeq* h_seq = nullptr;
    vb::FeSeq* d_seq = nullptr;
    int raw_pitch = 0, capacity = 0, nlev = 0;
    uint8_t* d_raw = nullptr;   // S x rows x raw_pitch
    uint8_t* h_img = nullptr;   // pinned S x rows x cols
    uint8_t* h_up = nullptr;    // pinned: pts_in (S x cap x 2 f32) | centres (S x cap x 2 i32)
    uint8_t* d_up = nullptr;
    uint8_t* h_dn = nullptr;    // pinned: pts_out (S x cap x 2 f32) | status (S x cap) | counts (S x 2 i32) | new_pts (S x cap x 2 f32)
    uint8_t* d_dn = nullptr;
    size_t up_pts_bytes = 0, up_bytes = 0, dn_track_bytes = 0, dn_counts_off = 0, dn_bytes = 0;
    int* d_halfw = nullptr;
    vb::KernelProfile prof;  // 0 clahe_lut+apply, 1 pyrdown x3, 2 lk, 3 mask, 4 min_eig, 5 candidates+sort+select
    vb::HostPool* pool = nullptr;
    float last_ms = 0;
    int last_launches = 0;
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

// Recursive inverse of the radial-tangential distortion, n = 8 (PinholeCamera.cc:487-503, CataCamera.cc:596-611).
inline void undistort_radtan(double k1, double k2, double p1, double p2, double mx_d, double my_d, double& ux, double& uy) {
    ux = mx_d;
    uy = my_d;
    for (int i = 0; i < 8; ++i) {
        const double x2 = ux * ux, y2 = uy * uy, xy = ux * uy, rho2 = x2 + y2;
        const double rad = k1 * rho2 + k2 * rho2 * rho2;
        const double ddx = ux * rad + 2.0 * p1 * xy + p2 * (rho2 + 2.0 * x2);
        const double ddy = uy * rad + 2.0 * p2 * xy + p1 * (rho2 + 2.0 * y2);
        ux = mx_d - ddx;
        uy = my_d - ddy;
    }
}

// EquidistantCamera::backprojectSymmetric (EquidistantCamera.cc:716-818): the smallest non-negative real root of
// k5 t^9 + k4 t^7 + k3 t^5 + k2 t^3 + t - r = 0, r itself when there is none.  The reference takes it from the eigenvalues of
// the companion matrix; here the first sign change of the polynomial on a fine grid is bracketed and polished by
// bisection + Newton (the same root to rounding).
double kb_theta(const double kin[4], double r) {
    // the reference lowers the polynomial degree by 2 for EVERY zero coefficient, whichever it is, and drops the terms above
    // the resulting degree (EquidistantCamera.cc:733-770)
    int npow = 9;
    for (int i = 0; i < 4; i++) npow -= kin[i] == 0.0 ? 2 : 0;
    if (npow == 1) return r;
    double k[4];
    for (int i = 0; i < 4; i++) k[i] = 2 * i + 3 <= npow ? kin[i] : 0.0;
    auto f = [&](double t) {
        const double t2 = t * t;
        return t * (1.0 + t2 * (k[0] + t2 * (k[1] + t2 * (k[2] + t2 * k[3])))) - r;
    };
    auto df = [&](double t) {
        const double t2 = t * t;
        return 1.0 + t2 * (3.0 * k[0] + t2 * (5.0 * k[1] + t2 * (7.0 * k[2] + t2 * 9.0 * k[3])));
    };
    if (r <= 1e-10) return 0.0;  // f(0) = -r: the root is within the reference's tolerance of zero
    const double step = 1.0 / 512.0;
    double lo = 0.0;
    for (int i = 1; i <= 4096; i++) {  // up to theta = 8 rad, far beyond any lens
        double hi = i * step;
        if (f(hi) >= 0.0) {
            for (int it = 0; it < 24; it++) {  // bisection narrows the bracket, Newton finishes
                const double mid = 0.5 * (lo + hi);
                if (f(mid) < 0.0) lo = mid; else hi = mid;
            }
            double t = 0.5 * (lo + hi);
            for (int it = 0; it < 20; it++) {
                const double d = df(t);
                if (d == 0.0) break;
                const double tn = t - f(t) / d;
                if (!(tn >= lo && tn <= hi)) break;
                if (tn == t) break;
                t = tn;
            }
            return t;
        }
        lo = hi;
    }
    return r;
}

// CameraPtr::liftProjective followed by the division by z that every caller applies (feature_tracker.cpp:179-186, 268-272):
//   PINHOLE  PinholeCamera.cc:450-510      intrinsics = fx fy cx cy k1 k2 p1 p2
//   MEI      CataCamera.cc:556-625         intrinsics = gamma1 gamma2 u0 v0 k1 k2 p1 p2, cfg.xi
//   KANNALA_BRANDT  EquidistantCamera.cc:428-442   intrinsics = mu mv u0 v0 k2 k3 k4 k5
void lift_projective(const vt_config& c, double px, double py, double& X, double& Y) {
    const double fx = c.intrinsics[0], fy = c.intrinsics[1], cx = c.intrinsics[2], cy = c.intrinsics[3];
    const double k1 = c.intrinsics[4], k2 = c.intrinsics[5], p1 = c.intrinsics[6], p2 = c.intrinsics[7];
    const double mx_d = (1.0 / fx) * px + (-cx / fx), my_d = (1.0 / fy) * py + (-cy / fy);
    if (c.camera_model == VT_CAMERA_KANNALA_BRANDT) {
        const double r = std::sqrt(mx_d * mx_d + my_d * my_d);
        const double phi = r < 1e-10 ? 0.0 : std::atan2(my_d, mx_d);
        const double theta = kb_theta(&c.intrinsics[4], r);
        const double Px = std::sin(theta) * std::cos(phi), Py = std::sin(theta) * std::sin(phi), Pz = std::cos(theta);
        X = Px / Pz;
        Y = Py / Pz;
        return;
    }
    double ux = mx_d, uy = my_d;
    if (!(k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0)) undistort_radtan(k1, k2, p1, p2, mx_d, my_d, ux, uy);
    if (c.camera_model == VT_CAMERA_MEI) {
        const double xi = c.xi;
        double Pz;
        if (xi == 1.0)
            Pz = (1.0 - ux * ux - uy * uy) / 2.0;
        else {
            const double rho2 = ux * ux + uy * uy;
            Pz = 1.0 - xi * (rho2 + 1.0) / (xi + std::sqrt(1.0 + (1.0 - xi * xi) * rho2));
        }
        X = ux / Pz;
        Y = uy / Pz;
        return;
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

// ---- one image step of a batch ------------------------------------------------------------------
struct ImageMsg {
    int active;
    const uint8_t* img;
    double stamp;
    int pub;
};

// Phase A: readImage up to the LK launch (feature_tracker.cpp:81-113): descriptor, points, frame staging.
void step_prepare(vt_tracker* t, const ImageMsg& m, size_t stride, bool on_device) {
    vt_batch* b = t->batch;
    vb::FeSeq& q = b->h_seq[t->member];
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    t->status = VT_OK;
    t->step_active = true;
    t->step_pub = m.pub != 0;
    t->step_detect = false;
    t->cur_time = m.stamp;
    t->h2d_bytes = on_device ? 0 : (size_t)rows * cols;
    t->d2h_bytes = 0;
    t->step_first = !t->have_img;
    t->have_img = true;
    q.track = 1;
    q.detect = 0;
    q.equalize = t->cfg.equalize ? 1 : 0;
    if (on_device) {
        q.raw = m.img;
        q.raw_pitch = (int)stride;
    } else {
        for (int y = 0; y < rows; y++) std::memcpy(t->h_img + (size_t)y * cols, m.img + (size_t)y * stride, cols);
        q.raw = t->d_raw;
        q.raw_pitch = b->raw_pitch;
    }
    q.lut = t->d_lut;
    // prev = cur = forw = img on the very first frame: the pyramid built now is also the current one
    q.forw = t->pyr[t->cur ^ 1].view;
    q.cur = t->step_first ? q.forw : t->pyr[t->cur].view;
    t->forw_pts.clear();
    const int n = t->step_first ? 0 : (int)t->cur_pts.size();
    t->step_n = n;
    if (n) std::memcpy(t->h_pts_in, t->cur_pts.data(), (size_t)n * sizeof(Pt));
    q.n_pts = n;
    q.pts_in = reinterpret_cast<const float*>(b->d_up + ((uint8_t*)t->h_pts_in - b->h_up));
    q.pts_out = reinterpret_cast<float*>(b->d_dn + ((uint8_t*)t->h_pts_out - b->h_dn));
    q.status = b->d_dn + (t->h_status - b->h_dn);
    t->h2d_bytes += (size_t)n * sizeof(Pt) + sizeof(vb::FeSeq);
    t->d2h_bytes += (size_t)n * (sizeof(Pt) + 1);
}

// Phase B: the culls after LK, and on publishing frames rejectWithF + setMask (feature_tracker.cpp:115-146).
void step_after_track(vt_tracker* t) {
    vt_batch* b = t->batch;
    vb::FeSeq& q = b->h_seq[t->member];
    q.track = 0;
    const int n = t->step_n;
    if (n) {
        t->forw_pts.assign(reinterpret_cast<Pt*>(t->h_pts_out), reinterpret_cast<Pt*>(t->h_pts_out) + n);
        // status already includes the inBorder() cull (fused into the LK kernel)
        compact(t->cur_pts, t->h_status);
        compact(t->forw_pts, t->h_status);
        compact(t->ids, t->h_status);
        compact(t->cur_un_pts, t->h_status);
        compact(t->track_cnt, t->h_status);
    }
    for (auto& c : t->track_cnt) c++;
    t->n_pts.clear();
    if (!t->step_pub) return;
    reject_with_f(t);
    const int kept = set_mask(t);
    const int n_max_cnt = t->cfg.max_cnt - (int)t->forw_pts.size();
    if (n_max_cnt <= 0) return;
    t->step_detect = true;
    t->step_kept = kept;
    t->step_max_new = n_max_cnt;
    q.detect = 1;
    q.use_mask = 1;
    q.det_img = q.forw.img[0];
    q.det_pitch = q.forw.pitch[0];
    q.n_centres = kept;
    q.mask = t->d_mask;
    q.mask_init = t->cfg.fisheye ? t->d_fisheye : nullptr;
    q.centres = reinterpret_cast<const int*>(b->d_up + ((uint8_t*)t->h_centres - b->h_up));
    q.eig = t->d_eig;
    q.keys = t->d_keys;
    q.count = reinterpret_cast<int*>(b->d_dn + ((uint8_t*)t->h_counts - b->h_dn));
    q.maxv = reinterpret_cast<unsigned*>(q.count + 2);  // the third int of the member's counter slot
    q.cell_cnt = t->d_cell_cnt;
    q.cell_pts = t->d_cell_pts;
    q.new_pts = reinterpret_cast<float*>(b->d_dn + ((uint8_t*)t->h_new_pts - b->h_dn));
    q.max_corners = n_max_cnt;
    t->h2d_bytes += (size_t)kept * 2 * sizeof(int) + sizeof(vb::FeSeq);
    t->d2h_bytes += 4 * sizeof(int) + (size_t)n_max_cnt * 2 * sizeof(float);
}

// Phase C: addPoints, rotation of the image / point sets, undistortedPoints, updateID (feature_tracker.cpp:148-167,
// feature_tracker_node.cpp:103-111).
void step_finish(vt_tracker* t) {
    if (t->step_detect) {
        if (t->h_counts[0] > t->key_capacity) {
            t->err = "Shi-Tomasi candidate buffer overflow";
            t->status = VT_ERR_CAPACITY;
        } else {
            const int n_new = t->h_counts[1];
            for (int i = 0; i < n_new; i++) t->n_pts.push_back(Pt{t->h_new_pts[2 * i], t->h_new_pts[2 * i + 1]});
        }
    }
    if (t->step_pub)
        for (auto& p : t->n_pts) {  // addPoints (feature_tracker.cpp:71-79)
            t->forw_pts.push_back(p);
            t->ids.push_back(-1);
            t->track_cnt.push_back(1);
        }
    t->cur ^= 1;  // cur_img = forw_img (prev_img is never read again by the hot path)
    t->cur_pts = t->forw_pts;
    undistorted_points(t);
    t->prev_time = t->cur_time;
    for (size_t i = 0; i < t->ids.size(); i++)  // updateID loop of img_callback (feature_tracker_node.cpp:103-111)
        if (t->ids[i] == -1) t->ids[i] = t->n_id++;
    t->step_active = false;
}

#define VTB_CUDA(call)                                                       \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            b->err = std::string(#call) + ": " + cudaGetErrorString(e_);     \
            return VT_ERR_CUDA;                                              \
        }                                                                    \
    } while (0)

vb::FeShape base_shape(const vt_batch* b) {
    vb::FeShape sh{};
    sh.S = b->S;
    sh.rows = b->cfg.rows;
    sh.cols = b->cfg.cols;
    sh.nlev = b->nlev;
    sh.min_dist = b->cfg.min_dist;
    sh.key_capacity = b->members[0]->key_capacity;
    return sh;
}

// readImage for every member with msgs[k].active: all of them advance by one image together.
int batch_read_image(vt_batch* b, const ImageMsg* msgs, size_t stride, bool on_device) {
    VTB_CUDA(cudaSetDevice(b->cfg.device));
    const int S = b->S;
    static const bool trace = std::getenv("VINSB200_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
    const auto t0 = now();
    b->last_launches = 0;
    b->pool->run(S, [&](int k) {
        vt_tracker* t = b->members[k];
        vb::FeSeq& q = b->h_seq[k];
        q.track = q.detect = 0;
        q.n_pts = 0;
        t->step_active = false;
        t->status = VT_OK;
        if (msgs[k].active) step_prepare(t, msgs[k], stride, on_device);
    });
    vb::FeShape sh = base_shape(b);
    int first_active = -1, last_active = -1;
    for (int k = 0; k < S; k++) {
        const vb::FeSeq& q = b->h_seq[k];
        if (!q.track) continue;
        if (first_active < 0) first_active = k;
        last_active = k;
        sh.any_track = 1;
        sh.any_equalize |= q.equalize;
        sh.max_pts = std::max(sh.max_pts, q.n_pts);
    }
    if (!sh.any_track) return VT_OK;
    const auto t1 = now();
    VTB_CUDA(cudaEventRecord(b->ev0, b->stream));
    if (!on_device) {  // the frames of members first_active..last_active travel in one copy
        const int rows = b->cfg.rows, cols = b->cfg.cols;
        VTB_CUDA(cudaMemcpy2DAsync(b->d_raw + (size_t)first_active * rows * b->raw_pitch, b->raw_pitch,
                                   b->h_img + (size_t)first_active * rows * cols, cols, cols, (size_t)rows * (last_active - first_active + 1),
                                   cudaMemcpyHostToDevice, b->stream));
    }
    VTB_CUDA(cudaMemcpyAsync(b->d_seq, b->h_seq, sizeof(vb::FeSeq) * S, cudaMemcpyHostToDevice, b->stream));
    if (sh.max_pts > 0) VTB_CUDA(cudaMemcpyAsync(b->d_up, b->h_up, b->up_pts_bytes, cudaMemcpyHostToDevice, b->stream));
    vb::launch_track(b->d_seq, sh, b->stream, &b->last_launches, &b->prof);
    if (sh.max_pts > 0) VTB_CUDA(cudaMemcpyAsync(b->h_dn, b->d_dn, b->dn_track_bytes, cudaMemcpyDeviceToHost, b->stream));
    VTB_CUDA(cudaStreamSynchronize(b->stream));
    const auto t2 = now();
    b->pool->run(S, [&](int k) {
        if (b->members[k]->step_active) step_after_track(b->members[k]);
    });
    const auto t3 = now();
    for (int k = 0; k < S; k++) {
        const vb::FeSeq& q = b->h_seq[k];
        if (!q.detect) continue;
        sh.any_detect = 1;
        sh.max_centres = std::max(sh.max_centres, q.n_centres);
    }
    if (sh.any_detect) {
        VTB_CUDA(cudaMemcpyAsync(b->d_seq, b->h_seq, sizeof(vb::FeSeq) * S, cudaMemcpyHostToDevice, b->stream));
        if (sh.max_centres > 0)
            VTB_CUDA(cudaMemcpyAsync(b->d_up + b->up_pts_bytes, b->h_up + b->up_pts_bytes, b->up_bytes - b->up_pts_bytes,
                                     cudaMemcpyHostToDevice, b->stream));
        vb::launch_detect(b->d_seq, sh, b->d_halfw, b->stream, &b->last_launches, &b->prof);
        VTB_CUDA(cudaMemcpyAsync(b->h_dn + b->dn_counts_off, b->d_dn + b->dn_counts_off, b->dn_bytes - b->dn_counts_off,
                                 cudaMemcpyDeviceToHost, b->stream));
    }
    VTB_CUDA(cudaEventRecord(b->ev1, b->stream));
    VTB_CUDA(cudaEventSynchronize(b->ev1));
    VTB_CUDA(cudaGetLastError());
    VTB_CUDA(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
    const auto t4 = now();
    b->pool->run(S, [&](int k) {
        if (b->members[k]->step_active) step_finish(b->members[k]);
    });
    if (trace)
        std::fprintf(stderr, "[vt_batch] prepare %.3f track(gpu) %.3f host %.3f detect(gpu) %.3f finish %.3f ms (S=%d detect=%d)\n", ms(t0, t1), ms(t1, t2),
                     ms(t2, t3), ms(t3, t4), ms(t4, now()), S, sh.any_detect);
    int rc = VT_OK;
    for (int k = 0; k < S; k++)
        if (b->members[k]->status != VT_OK && rc == VT_OK) {
            rc = b->members[k]->status;
            b->err = "member " + std::to_string(k) + ": " + b->members[k]->err;
        }
    return rc;
}

// img_callback's gating (feature_tracker_node.cpp:28-62): returns -1 when readImage is due (pub decided), else the
// node result (0) with *restart set when the discontinuity rule fired.
int node_gate(vt_tracker* t, double stamp, int* restart, bool* pub) {
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
    if (std::round(1.0 * t->pub_count / (stamp - t->first_image_time)) <= t->cfg.freq) {
        *pub = true;
        if (std::abs(1.0 * t->pub_count / (stamp - t->first_image_time) - t->cfg.freq) < 0.01 * t->cfg.freq) {
            t->first_image_time = stamp;
            t->pub_count = 0;
        }
    } else
        *pub = false;
    return -1;
}

int node_after(vt_tracker* t, bool pub) {
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

void destroy_member(vt_tracker* t) {
    if (!t) return;
    cudaFree(t->d_lut);
    cudaFree(t->pyr[0].base);
    cudaFree(t->pyr[1].base);
    cudaFree(t->d_mask);
    cudaFree(t->d_fisheye);
    cudaFree(t->d_eig);
    cudaFree(t->d_keys);
    cudaFree(t->d_cell_cnt);
    cudaFree(t->d_cell_pts);
    delete[] t->owned_fisheye_mask;
    delete t;
}

vt_tracker* create_member(vt_batch* b, int k) {
    vt_tracker* t = new vt_tracker();
    t->cfg = b->cfg;
    t->cfg.fisheye_mask = nullptr;
    t->batch = b;
    t->member = k;
    const vt_config* cfg = &b->cfg;
    const int rows = cfg->rows, cols = cfg->cols;
    const size_t npx = (size_t)rows * cols;
    t->capacity = b->capacity;
    if (cfg->fisheye) {  // own a copy first: setMask reads it on the host
        t->owned_fisheye_mask = new uint8_t[npx];
        std::memcpy(t->owned_fisheye_mask, cfg->fisheye_mask, npx);
        t->cfg.fisheye_mask = t->owned_fisheye_mask;
    }
    bool ok = cudaMalloc(&t->d_lut, 64 * 256) == cudaSuccess;
    ok = ok && alloc_pyramid(t, t->pyr[0]) == VT_OK && alloc_pyramid(t, t->pyr[1]) == VT_OK;
    ok = ok && cudaMalloc(&t->d_mask, npx) == cudaSuccess && cudaMalloc(&t->d_eig, npx * sizeof(float)) == cudaSuccess;
    t->key_capacity = 1;
    while ((size_t)t->key_capacity < npx / 8) t->key_capacity <<= 1;  // local maxima of a 3x3 NMS: < npx/4 in practice
    ok = ok && cudaMalloc(&t->d_keys, (size_t)t->key_capacity * sizeof(unsigned long long)) == cudaSuccess;
    const int cell = cfg->min_dist, gw = (cols + cell - 1) / cell, gh = (rows + cell - 1) / cell;
    ok = ok && cudaMalloc(&t->d_cell_cnt, (size_t)gw * gh * sizeof(int)) == cudaSuccess &&
         cudaMalloc(&t->d_cell_pts, (size_t)gw * gh * 8 * sizeof(short2)) == cudaSuccess;
    if (ok && cfg->fisheye)
        ok = cudaMalloc(&t->d_fisheye, npx) == cudaSuccess &&
             cudaMemcpy(t->d_fisheye, t->owned_fisheye_mask, npx, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        destroy_member(t);
        return nullptr;
    }
    t->halfw = disc_half_widths(cfg->min_dist);
    const size_t cap = (size_t)b->capacity, S = (size_t)b->S;
    t->d_raw = b->d_raw + (size_t)k * rows * b->raw_pitch;
    t->h_img = b->h_img + (size_t)k * npx;
    t->h_pts_in = reinterpret_cast<float*>(b->h_up) + (size_t)k * cap * 2;
    t->h_centres = reinterpret_cast<int*>(b->h_up + b->up_pts_bytes) + (size_t)k * cap * 2;
    t->h_pts_out = reinterpret_cast<float*>(b->h_dn) + (size_t)k * cap * 2;
    t->h_status = b->h_dn + S * cap * 2 * sizeof(float) + (size_t)k * cap;
    t->h_counts = reinterpret_cast<int*>(b->h_dn + b->dn_counts_off) + (size_t)k * 4;
    t->h_new_pts = reinterpret_cast<float*>(b->h_dn + b->dn_counts_off + S * 4 * sizeof(int)) + (size_t)k * cap * 2;
    return t;
}

}  // namespace

extern "C" {

int vt_batch_create(const vt_config* cfg, int n, vt_batch** out) {
    if (!cfg || !out || n < 1 || n > 4096) return VT_ERR_INVALID;
    *out = nullptr;
    if (cfg->rows < 32 || cfg->cols < 32 || cfg->max_cnt <= 0 || cfg->min_dist < 1 ||
        cfg->camera_model < VT_CAMERA_PINHOLE || cfg->camera_model > VT_CAMERA_KANNALA_BRANDT || (cfg->fisheye && !cfg->fisheye_mask))
        return VT_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
        return VT_ERR_NO_DEVICE;
    vt_batch* b = new vt_batch();
    b->cfg = *cfg;
    if (b->cfg.freq == 0) b->cfg.freq = 100;
    if (b->cfg.focal_length == 0) b->cfg.focal_length = 460;
    b->S = n;
    auto fail = [&](int code) {
        vt_batch_destroy(b);
        return code;
    };
#define VT_TRY(call)                                         \
    do {                                                     \
        if ((call) != cudaSuccess) return fail(VT_ERR_CUDA); \
    } while (0)
    VT_TRY(cudaSetDevice(cfg->device));
    VT_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    VT_TRY(cudaEventCreate(&b->ev0));
    VT_TRY(cudaEventCreate(&b->ev1));
    const int rows = cfg->rows, cols = cfg->cols;
    const size_t npx = (size_t)rows * cols, S = (size_t)n;
    b->raw_pitch = align_up(cols, 64);
    b->capacity = std::max(cfg->max_cnt, 16);
    b->nlev = lk_levels(rows, cols, 21, 3);
    const size_t cap = (size_t)b->capacity;
    b->up_pts_bytes = S * cap * 2 * sizeof(float);
    b->up_bytes = b->up_pts_bytes + S * cap * 2 * sizeof(int);
    b->dn_track_bytes = S * cap * 2 * sizeof(float) + S * cap;
    b->dn_counts_off = (b->dn_track_bytes + 15) & ~(size_t)15;
    b->dn_bytes = b->dn_counts_off + S * 4 * sizeof(int) + S * cap * 2 * sizeof(float);
    VT_TRY(cudaMalloc(&b->d_raw, S * rows * b->raw_pitch));
    VT_TRY(cudaHostAlloc(&b->h_img, S * npx, cudaHostAllocDefault));
    VT_TRY(cudaMalloc(&b->d_up, b->up_bytes));
    VT_TRY(cudaHostAlloc(&b->h_up, b->up_bytes, cudaHostAllocDefault));
    VT_TRY(cudaMalloc(&b->d_dn, b->dn_bytes));
    VT_TRY(cudaHostAlloc(&b->h_dn, b->dn_bytes, cudaHostAllocDefault));
    VT_TRY(cudaMalloc(&b->d_seq, sizeof(vb::FeSeq) * S));
    VT_TRY(cudaHostAlloc(&b->h_seq, sizeof(vb::FeSeq) * S, cudaHostAllocDefault));
    std::memset(b->h_seq, 0, sizeof(vb::FeSeq) * S);
    const std::vector<int> hw = disc_half_widths(cfg->min_dist);
    VT_TRY(cudaMalloc(&b->d_halfw, hw.size() * sizeof(int)));
    VT_TRY(cudaMemcpy(b->d_halfw, hw.data(), hw.size() * sizeof(int), cudaMemcpyHostToDevice));
#undef VT_TRY
    for (int k = 0; k < n; k++) {
        vt_tracker* t = create_member(b, k);
        if (!t) return fail(VT_ERR_CUDA);
        t->cfg.freq = b->cfg.freq;
        t->cfg.focal_length = b->cfg.focal_length;
        b->members.push_back(t);
    }
    b->pool = new vb::HostPool(vb::HostPool::default_workers(n));
    *out = b;
    return VT_OK;
}

void vt_batch_destroy(vt_batch* b) {
    if (!b) return;
    cudaSetDevice(b->cfg.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    delete b->pool;
    for (auto* t : b->members) destroy_member(t);
    cudaFree(b->d_raw); cudaFree(b->d_up); cudaFree(b->d_dn); cudaFree(b->d_seq); cudaFree(b->d_halfw);
    cudaFreeHost(b->h_img); cudaFreeHost(b->h_up); cudaFreeHost(b->h_dn); cudaFreeHost(b->h_seq);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int vt_batch_size(const vt_batch* b) { return b ? b->S : VT_ERR_INVALID; }
vt_tracker* vt_batch_member(vt_batch* b, int k) { return (b && k >= 0 && k < b->S) ? b->members[k] : nullptr; }
const char* vt_batch_last_error(const vt_batch* b) { return b ? b->err.c_str() : "null batch"; }

int vt_batch_read_image(vt_batch* b, const int* active, const uint8_t* const* imgs, size_t row_stride, const double* cur_times,
                        const int* pub_this_frame, int images_on_device) {
    if (!b || !imgs || !cur_times || !pub_this_frame || row_stride < (size_t)b->cfg.cols) return VT_ERR_INVALID;
    std::vector<ImageMsg> msgs(b->S);
    for (int k = 0; k < b->S; k++) {
        msgs[k] = ImageMsg{active ? (active[k] != 0) : 1, imgs[k], cur_times[k], pub_this_frame[k]};
        if (msgs[k].active && !imgs[k]) return VT_ERR_INVALID;
    }
    return batch_read_image(b, msgs.data(), row_stride, images_on_device != 0);
}

int vt_batch_node_image(vt_batch* b, const int* active, const uint8_t* const* imgs, size_t row_stride, const double* stamps,
                        int images_on_device, int* results, int* restarts) {
    if (!b || !imgs || !stamps || !results || row_stride < (size_t)b->cfg.cols) return VT_ERR_INVALID;
    std::vector<ImageMsg> msgs(b->S);
    for (int k = 0; k < b->S; k++) {
        msgs[k] = ImageMsg{0, imgs[k], stamps[k], 0};
        results[k] = 0;
        if (restarts) restarts[k] = 0;
        if (active && !active[k]) continue;
        if (!imgs[k]) return VT_ERR_INVALID;
        bool pub = false;
        const int r = node_gate(b->members[k], stamps[k], restarts ? restarts + k : nullptr, &pub);
        if (r >= 0) continue;  // first frame or restart: the image is consumed without tracking
        msgs[k].active = 1;
        msgs[k].pub = pub ? 1 : 0;
    }
    const int rc = batch_read_image(b, msgs.data(), row_stride, images_on_device != 0);
    if (rc) return rc;
    for (int k = 0; k < b->S; k++)
        if (msgs[k].active) results[k] = node_after(b->members[k], msgs[k].pub != 0);
    return VT_OK;
}

int vt_batch_last_timing(const vt_batch* b, float* device_ms, int* kernel_launches) {
    if (!b) return VT_ERR_INVALID;
    if (device_ms) *device_ms = b->last_ms;
    if (kernel_launches) *kernel_launches = b->last_launches;
    return VT_OK;
}

int vt_batch_set_profile(vt_batch* b, int on) {
    if (!b) return VT_ERR_INVALID;
    b->prof.enable(on != 0);
    return VT_OK;
}

int vt_batch_kernel_times(const vt_batch* b, double* ms6, int* count6) {
    if (!b) return VT_ERR_INVALID;
    for (int k = 0; k < 6; k++) {
        if (ms6) ms6[k] = b->prof.ms[k];
        if (count6) count6[k] = b->prof.count[k];
    }
    return VT_OK;
}

int vt_create(const vt_config* cfg, vt_tracker** out) {
    if (!cfg || !out) return VT_ERR_INVALID;
    *out = nullptr;
    vt_batch* b = nullptr;
    const int rc = vt_batch_create(cfg, 1, &b);
    if (rc) return rc;
    b->standalone = true;
    *out = b->members[0];
    return VT_OK;
}

void vt_destroy(vt_tracker* t) {
    if (t && t->batch && t->batch->standalone) vt_batch_destroy(t->batch);  // members of an explicit batch die with it
}

const char* vt_last_error(const vt_tracker* t) {
    if (!t) return "null handle";
    return t->err.empty() && t->batch ? t->batch->err.c_str() : t->err.c_str();
}

static int read_image_single(vt_tracker* t, const uint8_t* img, size_t stride, double cur_time, int pub, bool on_device) {
    if (!t || !img || stride < (size_t)t->cfg.cols) return VT_ERR_INVALID;
    if (!t->batch->standalone) {
        t->err = "per-handle image call on a member of a batch: use vt_batch_read_image / vt_batch_node_image";
        return VT_ERR_INVALID;
    }
    t->err.clear();
    ImageMsg m{1, img, cur_time, pub};
    return batch_read_image(t->batch, &m, stride, on_device);
}

int vt_read_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double cur_time, int pub_this_frame) {
    return read_image_single(t, img, row_stride, cur_time, pub_this_frame, false);
}

int vt_read_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double cur_time, int pub_this_frame) {
    return read_image_single(t, d_img, row_stride, cur_time, pub_this_frame, true);
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

static int node_image_single(vt_tracker* t, const uint8_t* img, size_t row_stride, double stamp, int* restart, bool on_device) {
    if (!t) return VT_ERR_INVALID;
    bool pub = false;
    const int r = node_gate(t, stamp, restart, &pub);
    if (r >= 0) return r;
    const int rc = read_image_single(t, img, row_stride, stamp, pub, on_device);
    if (rc) return rc;
    return node_after(t, pub);
}

int vt_node_image(vt_tracker* t, const uint8_t* img, size_t row_stride, double stamp, int* restart) {
    return node_image_single(t, img, row_stride, stamp, restart, false);
}

int vt_node_image_device(vt_tracker* t, const uint8_t* d_img, size_t row_stride, double stamp, int* restart) {
    return node_image_single(t, d_img, row_stride, stamp, restart, true);
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

int vt_debug_fundamental_ransac_model(const float* pts1, const float* pts2, int n, double threshold, double confidence, uint8_t* status, double* F9) {
    if (!pts1 || !pts2 || !status || !F9 || n < 0) return VT_ERR_INVALID;
    return vb::fundamental_ransac(pts1, pts2, n, threshold, confidence, status, F9) ? 1 : 0;
}

int vt_debug_lift_projective(const double* intrinsics8, const double* px, int n, double* out_xy) {
    if (!intrinsics8 || !px || !out_xy || n < 0) return VT_ERR_INVALID;
    vt_config c{};
    for (int i = 0; i < 8; i++) c.intrinsics[i] = intrinsics8[i];
    for (int k = 0; k < n; k++) lift_projective(c, px[2 * k], px[2 * k + 1], out_xy[2 * k], out_xy[2 * k + 1]);
    return VT_OK;
}

int vt_debug_lift_projective_model(int camera_model, const double* intrinsics8, double xi, const double* px, int n, double* out_xy) {
    if (!intrinsics8 || !px || !out_xy || n < 0 || camera_model < VT_CAMERA_PINHOLE || camera_model > VT_CAMERA_KANNALA_BRANDT)
        return VT_ERR_INVALID;
    vt_config c{};
    c.camera_model = camera_model;
    c.xi = xi;
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
    return t ? vt_batch_last_timing(t->batch, device_ms, kernel_launches) : VT_ERR_INVALID;
}

int vt_set_profile(vt_tracker* t, int on) { return t ? vt_batch_set_profile(t->batch, on) : VT_ERR_INVALID; }

int vt_kernel_times(const vt_tracker* t, double* ms6, int* count6) { return t ? vt_batch_kernel_times(t->batch, ms6, count6) : VT_ERR_INVALID; }

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
    vt_batch* b = t->batch;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    uint8_t* d_user_mask = nullptr;
    VT_CUDA(cudaMemcpy2DAsync(t->d_raw, b->raw_pitch, img, row_stride, cols, rows, cudaMemcpyHostToDevice, b->stream));
    if (mask) {
        VT_CUDA(cudaMalloc(&d_user_mask, (size_t)rows * cols));
        VT_CUDA(cudaMemcpyAsync(d_user_mask, mask, (size_t)rows * cols, cudaMemcpyHostToDevice, b->stream));
    }
    vb::FeSeq& q = b->h_seq[t->member];
    std::memset(&q, 0, sizeof(q));
    q.detect = 1;
    q.use_mask = mask ? 1 : 0;
    q.det_img = t->d_raw;
    q.det_pitch = b->raw_pitch;
    q.mask = t->d_mask;
    q.mask_init = d_user_mask;
    q.eig = t->d_eig;
    q.keys = t->d_keys;
    q.count = reinterpret_cast<int*>(b->d_dn + ((uint8_t*)t->h_counts - b->h_dn));
    q.maxv = reinterpret_cast<unsigned*>(q.count + 2);
    q.cell_cnt = t->d_cell_cnt;
    q.cell_pts = t->d_cell_pts;
    q.new_pts = reinterpret_cast<float*>(b->d_dn + ((uint8_t*)t->h_new_pts - b->h_dn));
    q.max_corners = max_corners;
    VT_CUDA(cudaMemcpyAsync(b->d_seq, b->h_seq, sizeof(vb::FeSeq) * b->S, cudaMemcpyHostToDevice, b->stream));
    vb::FeShape sh = base_shape(b);
    sh.any_detect = 1;
    vb::launch_detect(b->d_seq, sh, b->d_halfw, b->stream, nullptr, nullptr);
    VT_CUDA(cudaMemcpyAsync(b->h_dn + b->dn_counts_off, b->d_dn + b->dn_counts_off, b->dn_bytes - b->dn_counts_off,
                            cudaMemcpyDeviceToHost, b->stream));
    VT_CUDA(cudaStreamSynchronize(b->stream));
    q.detect = 0;
    if (d_user_mask) cudaFree(d_user_mask);
    if (t->h_counts[0] > t->key_capacity) {
        t->err = "Shi-Tomasi candidate buffer overflow";
        return VT_ERR_CAPACITY;
    }
    const int n = t->h_counts[1];
    if (n_candidates) *n_candidates = t->h_counts[0];
    std::memcpy(corners, t->h_new_pts, (size_t)n * 2 * sizeof(float));
    if (eig_out) VT_CUDA(cudaMemcpy(eig_out, t->d_eig, (size_t)rows * cols * sizeof(float), cudaMemcpyDeviceToHost));
    return n;
}

int vt_debug_lk(vt_tracker* t, const uint8_t* prev, const uint8_t* next, size_t row_stride, const float* pts, int n,
                float* next_pts, uint8_t* status) {
    if (!t || !prev || !next || n < 0 || n > t->capacity) return VT_ERR_INVALID;
    vt_batch* b = t->batch;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    for (int k = 0; k < 2; k++) {
        DevicePyramid& p = t->pyr[k];
        VT_CUDA(cudaMemcpy2DAsync(const_cast<uint8_t*>(p.view.img[0]), p.view.pitch[0], k == 0 ? prev : next, row_stride,
                                  cols, rows, cudaMemcpyHostToDevice, b->stream));
    }
    vb::FeSeq& q = b->h_seq[t->member];
    std::memset(&q, 0, sizeof(q));
    q.track = 1;
    q.cur = t->pyr[0].view;
    q.forw = t->pyr[1].view;
    q.n_pts = n;
    q.pts_in = reinterpret_cast<const float*>(b->d_up + ((uint8_t*)t->h_pts_in - b->h_up));
    q.pts_out = reinterpret_cast<float*>(b->d_dn + ((uint8_t*)t->h_pts_out - b->h_dn));
    q.status = b->d_dn + (t->h_status - b->h_dn);
    if (n) std::memcpy(t->h_pts_in, pts, (size_t)n * 2 * sizeof(float));
    VT_CUDA(cudaMemcpyAsync(b->d_seq, b->h_seq, sizeof(vb::FeSeq) * b->S, cudaMemcpyHostToDevice, b->stream));
    VT_CUDA(cudaMemcpyAsync(b->d_up, b->h_up, b->up_pts_bytes, cudaMemcpyHostToDevice, b->stream));
    vb::FeShape sh = base_shape(b);
    sh.any_track = 1;
    sh.max_pts = n;
    vb::launch_pyramid_only(b->d_seq, sh, 1, b->stream);
    vb::launch_pyramid_only(b->d_seq, sh, 0, b->stream);
    vb::launch_lk_only(b->d_seq, sh, b->stream);
    VT_CUDA(cudaMemcpyAsync(b->h_dn, b->d_dn, b->dn_track_bytes, cudaMemcpyDeviceToHost, b->stream));
    VT_CUDA(cudaStreamSynchronize(b->stream));
    q.track = 0;
    if (n) {
        std::memcpy(next_pts, t->h_pts_out, (size_t)n * 2 * sizeof(float));
        std::memcpy(status, t->h_status, n);
    }
    t->have_img = false;  // the pyramids no longer belong to the tracking state
    return VT_OK;
}

}  // extern "C"

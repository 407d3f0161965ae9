// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU twin of the feature_tracker node's per-image path, following the reference line by line:
//   FeatureTracker::readImage            feature_tracker/src/feature_tracker.cpp:81-167
//   inBorder / reduceVector              :5-29
//   setMask                              :36-69   (std::sort with the reference's comparator; the
//                                                  unstable tie order therefore comes from the same
//                                                  libstdc++ introsort the reference links)
//   addPoints / updateID                 :71-79, :204-214
//   rejectWithF                          :169-202
//   undistortedPoints                    :258-306
//   PinholeCamera::liftProjective        camera_model/src/camera_models/PinholeCamera.cc:450-510, :646-662
//   img_callback gating + ID loop        feature_tracker/src/feature_tracker_node.cpp:28-62, :103-111, :113-165
// The OpenCV calls go to the restatements in fe_image.cpp / fe_fundamental.cpp.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <algorithm>

extern "C" {
void orc_clahe(const uint8_t*, int, int, int, double, int, int, uint8_t*, int);
void orc_lk(const uint8_t*, const uint8_t*, int, int, int, const float*, int, int, int, int, double, double, float*,
            uint8_t*);
int orc_gftt(const uint8_t*, int, int, int, const uint8_t*, int, int, double, double, float*, int*);
void orc_circle(uint8_t*, int, int, int, int, int, int, uint8_t);
int orc_find_fundamental_ransac(const float*, const float*, int, double, double, uint8_t*, int*);

struct orc_tracker_config {
    int rows, cols;
    int max_cnt, min_dist;
    int equalize;
    int freq;
    int focal_length;  // FOCAL_LENGTH = 460 (feature_tracker/src/parameters.cpp:65)
    int fisheye;       // use fisheye_mask as the initial mask
    double f_threshold;
    double fx, fy, cx, cy, k1, k2, p1, p2;  // PINHOLE (config/euroc/euroc_config.yaml:13-22)
    // model_type: 0 PINHOLE, 1 MEI (fx fy cx cy hold gamma1 gamma2 u0 v0; xi below), 2 KANNALA_BRANDT (fx fy cx cy hold mu mv u0 v0,
    // k1 k2 p1 p2 hold k2 k3 k4 k5)
    int camera_model;
    double xi;
};
}

namespace {

struct Pt {
    float x, y;
};

inline int cv_round_f(float v) { return (int)lrintf(v); }

struct Tracker {
    orc_tracker_config cfg;
    std::vector<uint8_t> fisheye_mask;
    std::vector<uint8_t> mask;
    std::vector<uint8_t> prev_img, cur_img, forw_img;
    bool have_img = false;
    std::vector<Pt> n_pts, prev_pts, cur_pts, forw_pts, prev_un_pts, cur_un_pts, pts_velocity;
    std::vector<int> ids, track_cnt;
    std::map<int, Pt> cur_un_pts_map, prev_un_pts_map;
    double cur_time = 0, prev_time = 0;
    int n_id = 0;  // static int FeatureTracker::n_id
    // node-level state (feature_tracker_node.cpp:21-26)
    double first_image_time = 0, last_image_time = 0;
    int pub_count = 1;
    bool first_image_flag = true, init_pub = false;
    // diagnostics of the last readImage
    int last_lk_in = 0, last_lk_ok = 0, last_ransac_in = 0, last_ransac_ok = 0, last_new = 0;

    // CameraPtr::liftProjective, returned as (x / z, y / z) like every caller uses it (feature_tracker.cpp:179-186, 268-272).
    //   PinholeCamera.cc:450-510, CataCamera.cc:556-625 (recursive distortion model, n = 8), EquidistantCamera.cc:428-442
    void lift(double px, double py, double& X, double& Y) const {
        const double inv_K11 = 1.0 / cfg.fx, inv_K13 = -cfg.cx / cfg.fx;
        const double inv_K22 = 1.0 / cfg.fy, inv_K23 = -cfg.cy / cfg.fy;
        double mx_d = inv_K11 * px + inv_K13;
        double my_d = inv_K22 * py + inv_K23;
        if (cfg.camera_model == 2) {
            // backprojectSymmetric (EquidistantCamera.cc:716-818): smallest non-negative real root of
            // theta + k2 theta^3 + k3 theta^5 + k4 theta^7 + k5 theta^9 = |p_u|.  The reference reads it off the companion
            // matrix; this oracle brackets the first sign change and runs the Illinois false-position iteration
            // (tests/test_host_frontend.py pins it against numpy.roots, i.e. against the companion-matrix eigenvalues).
            const double r = std::sqrt(mx_d * mx_d + my_d * my_d);
            const double phi = r < 1e-10 ? 0.0 : std::atan2(my_d, mx_d);
            double kk[4] = {cfg.k1, cfg.k2, cfg.p1, cfg.p2};
            int npow = 9;  // lowered by 2 per zero coefficient; higher terms are dropped (EquidistantCamera.cc:733-770)
            for (double kv : kk)
                if (kv == 0.0) npow -= 2;
            for (int i = 0; i < 4; i++)
                if (2 * i + 3 > npow) kk[i] = 0.0;
            auto poly = [&](double t) {
                double acc = 0, tp = t;
                const double t2 = t * t;
                acc = tp;
                for (int i = 0; i < 4; i++) {
                    tp *= t2;
                    acc += kk[i] * tp;
                }
                return acc - r;
            };
            double theta = r;
            if (r <= 1e-10)
                theta = 0.0;
            else if (kk[0] != 0.0 || kk[1] != 0.0 || kk[2] != 0.0 || kk[3] != 0.0) {
                double a = 0.0, fa = -r;
                bool found = false;
                for (int i = 1; i <= 2400 && !found; i++) {
                    double b = i / 300.0, fb = poly(b);
                    if (fb >= 0.0) {
                        found = true;
                        int side = 0;
                        for (int it = 0; it < 200 && fb != 0.0; it++) {
                            const double c = (a * fb - b * fa) / (fb - fa), fc = poly(c);
                            if (std::fabs(b - a) <= 4e-16 * std::fabs(b)) break;
                            if (fc * fb > 0) {
                                b = c; fb = fc;
                                if (side == -1) fa *= 0.5;
                                side = -1;
                            } else if (fc * fa > 0) {
                                a = c; fa = fc;
                                if (side == 1) fb *= 0.5;
                                side = 1;
                            } else {
                                a = b = c;
                                break;
                            }
                        }
                        theta = std::fabs(poly(a)) < std::fabs(poly(b)) ? a : b;
                    } else {
                        a = b;
                        fa = fb;
                    }
                }
            }
            const double Px = std::sin(theta) * std::cos(phi), Py = std::sin(theta) * std::sin(phi), Pz = std::cos(theta);
            X = Px / Pz;
            Y = Py / Pz;
            return;
        }
        double mx_u, my_u;
        if (cfg.k1 == 0.0 && cfg.k2 == 0.0 && cfg.p1 == 0.0 && cfg.p2 == 0.0) {
            mx_u = mx_d;
            my_u = my_d;
        } else {
            auto distortion = [&](double ux, double uy, double& dx, double& dy) {
                double mx2 = ux * ux, my2 = uy * uy, mxy = ux * uy;
                double rho2 = mx2 + my2;
                double rad = cfg.k1 * rho2 + cfg.k2 * rho2 * rho2;
                dx = ux * rad + 2.0 * cfg.p1 * mxy + cfg.p2 * (rho2 + 2.0 * mx2);
                dy = uy * rad + 2.0 * cfg.p2 * mxy + cfg.p1 * (rho2 + 2.0 * my2);
            };
            double dx, dy;
            distortion(mx_d, my_d, dx, dy);
            mx_u = mx_d - dx;
            my_u = my_d - dy;
            for (int i = 1; i < 8; ++i) {
                distortion(mx_u, my_u, dx, dy);
                mx_u = mx_d - dx;
                my_u = my_d - dy;
            }
        }
        if (cfg.camera_model == 1) {
            const double xi = cfg.xi;
            double z;
            if (xi == 1.0)
                z = (1.0 - mx_u * mx_u - my_u * my_u) / 2.0;
            else {
                const double rho2_d = mx_u * mx_u + my_u * my_u;
                z = 1.0 - xi * (rho2_d + 1.0) / (xi + std::sqrt(1.0 + (1.0 - xi * xi) * rho2_d));
            }
            X = mx_u / z;
            Y = my_u / z;
            return;
        }
        X = mx_u;
        Y = my_u;  // Z = 1
    }

    bool in_border(const Pt& pt) const {
        const int B = 1;
        int ix = cv_round_f(pt.x), iy = cv_round_f(pt.y);
        return B <= ix && ix < cfg.cols - B && B <= iy && iy < cfg.rows - B;
    }

    template <class T>
    static void reduce(std::vector<T>& v, const std::vector<uint8_t>& status) {
        int j = 0;
        for (int i = 0; i < (int)v.size(); i++)
            if (status[i]) v[j++] = v[i];
        v.resize(j);
    }

    void reject_with_f() {
        last_ransac_in = (int)forw_pts.size();
        last_ransac_ok = last_ransac_in;
        if (forw_pts.size() >= 8) {
            const int n = (int)cur_pts.size();
            std::vector<float> un_cur(2 * n), un_forw(2 * n);
            for (int i = 0; i < n; i++) {
                double X, Y;
                lift(cur_pts[i].x, cur_pts[i].y, X, Y);
                un_cur[2 * i] = (float)(cfg.focal_length * X / 1.0 + cfg.cols / 2.0);
                un_cur[2 * i + 1] = (float)(cfg.focal_length * Y / 1.0 + cfg.rows / 2.0);
                lift(forw_pts[i].x, forw_pts[i].y, X, Y);
                un_forw[2 * i] = (float)(cfg.focal_length * X / 1.0 + cfg.cols / 2.0);
                un_forw[2 * i + 1] = (float)(cfg.focal_length * Y / 1.0 + cfg.rows / 2.0);
            }
            std::vector<uint8_t> status(n);
            orc_find_fundamental_ransac(un_cur.data(), un_forw.data(), n, cfg.f_threshold, 0.99, status.data(),
                                        nullptr);
            reduce(prev_pts, status);
            reduce(cur_pts, status);
            reduce(forw_pts, status);
            reduce(cur_un_pts, status);
            reduce(ids, status);
            reduce(track_cnt, status);
            last_ransac_ok = (int)forw_pts.size();
        }
    }

    void set_mask() {
        if (cfg.fisheye)
            mask = fisheye_mask;
        else
            mask.assign((size_t)cfg.rows * cfg.cols, 255);
        std::vector<std::pair<int, std::pair<Pt, int>>> cnt_pts_id;
        for (unsigned i = 0; i < forw_pts.size(); i++)
            cnt_pts_id.push_back(std::make_pair(track_cnt[i], std::make_pair(forw_pts[i], ids[i])));
        std::sort(cnt_pts_id.begin(), cnt_pts_id.end(),
                  [](const std::pair<int, std::pair<Pt, int>>& a, const std::pair<int, std::pair<Pt, int>>& b) {
                      return a.first > b.first;
                  });
        forw_pts.clear();
        ids.clear();
        track_cnt.clear();
        for (auto& it : cnt_pts_id) {
            int px = cv_round_f(it.second.first.x), py = cv_round_f(it.second.first.y);
            // mask.at<uchar>(Point2f) -> Point(cvRound(x), cvRound(y)); tracked points are inBorder so
            // the access is always inside the image.
            if (mask[(size_t)py * cfg.cols + px] == 255) {
                forw_pts.push_back(it.second.first);
                ids.push_back(it.second.second);
                track_cnt.push_back(it.first);
                orc_circle(mask.data(), cfg.rows, cfg.cols, cfg.cols, px, py, cfg.min_dist, 0);
            }
        }
    }

    void undistorted_points() {
        cur_un_pts.clear();
        cur_un_pts_map.clear();
        for (unsigned i = 0; i < cur_pts.size(); i++) {
            double X, Y;
            lift(cur_pts[i].x, cur_pts[i].y, X, Y);
            Pt u{(float)(X / 1.0), (float)(Y / 1.0)};
            cur_un_pts.push_back(u);
            cur_un_pts_map.insert(std::make_pair(ids[i], u));
        }
        if (!prev_un_pts_map.empty()) {
            double dt = cur_time - prev_time;
            pts_velocity.clear();
            for (unsigned i = 0; i < cur_un_pts.size(); i++) {
                if (ids[i] != -1) {
                    auto it = prev_un_pts_map.find(ids[i]);
                    if (it != prev_un_pts_map.end()) {
                        double vx = (cur_un_pts[i].x - it->second.x) / dt;
                        double vy = (cur_un_pts[i].y - it->second.y) / dt;
                        pts_velocity.push_back(Pt{(float)vx, (float)vy});
                    } else
                        pts_velocity.push_back(Pt{0, 0});
                } else
                    pts_velocity.push_back(Pt{0, 0});
            }
        } else {
            for (unsigned i = 0; i < cur_pts.size(); i++) pts_velocity.push_back(Pt{0, 0});
        }
        prev_un_pts_map = cur_un_pts_map;
    }

    void read_image(const uint8_t* _img, int stride, double _cur_time, bool pub_this_frame) {
        const int rows = cfg.rows, cols = cfg.cols;
        std::vector<uint8_t> img((size_t)rows * cols);
        cur_time = _cur_time;
        if (cfg.equalize)
            orc_clahe(_img, rows, cols, stride, 3.0, 8, 8, img.data(), cols);
        else
            for (int y = 0; y < rows; y++) std::memcpy(&img[(size_t)y * cols], _img + (size_t)y * stride, cols);
        if (!have_img) {
            prev_img = cur_img = forw_img = img;
            have_img = true;
        } else
            forw_img = img;
        forw_pts.clear();
        last_lk_in = (int)cur_pts.size();
        last_lk_ok = 0;
        if (cur_pts.size() > 0) {
            const int n = (int)cur_pts.size();
            std::vector<uint8_t> status(n);
            forw_pts.resize(n);
            orc_lk(cur_img.data(), forw_img.data(), rows, cols, cols, &cur_pts[0].x, n, 21, 3, 30, 0.01, 1e-4,
                   &forw_pts[0].x, status.data());
            for (int i = 0; i < n; i++)
                if (status[i] && !in_border(forw_pts[i])) status[i] = 0;
            reduce(prev_pts, status);
            reduce(cur_pts, status);
            reduce(forw_pts, status);
            reduce(ids, status);
            reduce(cur_un_pts, status);
            reduce(track_cnt, status);
            last_lk_ok = (int)forw_pts.size();
        }
        for (auto& n : track_cnt) n++;
        last_new = 0;
        if (pub_this_frame) {
            reject_with_f();
            set_mask();
            int n_max_cnt = cfg.max_cnt - (int)forw_pts.size();
            if (n_max_cnt > 0) {
                std::vector<float> c(2 * (size_t)n_max_cnt);
                int k = orc_gftt(forw_img.data(), rows, cols, cols, mask.data(), cols, n_max_cnt, 0.01,
                                 (double)cfg.min_dist, c.data(), nullptr);
                n_pts.resize(k);
                for (int i = 0; i < k; i++) n_pts[i] = Pt{c[2 * i], c[2 * i + 1]};
            } else
                n_pts.clear();
            for (auto& p : n_pts) {  // addPoints
                forw_pts.push_back(p);
                ids.push_back(-1);
                track_cnt.push_back(1);
            }
            last_new = (int)n_pts.size();
        }
        prev_img = cur_img;
        prev_pts = cur_pts;
        prev_un_pts = cur_un_pts;
        cur_img = forw_img;
        cur_pts = forw_pts;
        undistorted_points();
        prev_time = cur_time;
    }

    void update_ids() {  // feature_tracker_node.cpp:103-111 with NUM_OF_CAM = 1
        for (unsigned i = 0; i < ids.size(); i++)
            if (ids[i] == -1) ids[i] = n_id++;
    }
};

}  // namespace

extern "C" {

void* orc_tracker_create(const orc_tracker_config* cfg, const uint8_t* fisheye_mask) {
    Tracker* t = new Tracker();
    t->cfg = *cfg;
    if (cfg->fisheye && fisheye_mask) t->fisheye_mask.assign(fisheye_mask, fisheye_mask + (size_t)cfg->rows * cfg->cols);
    return t;
}
void orc_tracker_destroy(void* h) { delete (Tracker*)h; }

// FeatureTracker::readImage followed by the node's updateID loop.
void orc_tracker_read_image(void* h, const uint8_t* img, int stride, double t, int pub_this_frame) {
    Tracker* T = (Tracker*)h;
    T->read_image(img, stride, t, pub_this_frame != 0);
    T->update_ids();
}

// img_callback: returns 0 = frame consumed without tracking (first frame / restart), 1 = tracked,
// 2 = tracked and a feature message is published.  *restart is set when the stream discontinuity
// rule fires (feature_tracker_node.cpp:38-48).
int orc_tracker_node_image(void* h, const uint8_t* img, int stride, double stamp, int* restart) {
    Tracker* T = (Tracker*)h;
    if (restart) *restart = 0;
    if (T->first_image_flag) {
        T->first_image_flag = false;
        T->first_image_time = stamp;
        T->last_image_time = stamp;
        return 0;
    }
    if (stamp - T->last_image_time > 1.0 || stamp < T->last_image_time) {
        T->first_image_flag = true;
        T->last_image_time = 0;
        T->pub_count = 1;
        if (restart) *restart = 1;
        return 0;
    }
    T->last_image_time = stamp;
    bool pub;
    if (std::round(1.0 * T->pub_count / (stamp - T->first_image_time)) <= T->cfg.freq) {
        pub = true;
        if (std::abs(1.0 * T->pub_count / (stamp - T->first_image_time) - T->cfg.freq) < 0.01 * T->cfg.freq) {
            T->first_image_time = stamp;
            T->pub_count = 0;
        }
    } else
        pub = false;
    T->read_image(img, stride, stamp, pub);
    T->update_ids();
    if (pub) {
        T->pub_count++;
        if (!T->init_pub) {
            T->init_pub = true;
            return 1;  // first feature set is never published
        }
        return 2;
    }
    return 1;
}

int orc_tracker_count(void* h) { return (int)((Tracker*)h)->cur_pts.size(); }

// Copies the public result vectors (cur_pts, cur_un_pts, pts_velocity, ids, track_cnt).
void orc_tracker_get(void* h, int* ids, int* track_cnt, float* cur_pts, float* un_pts, float* velocity) {
    Tracker* T = (Tracker*)h;
    const int n = (int)T->cur_pts.size();
    for (int i = 0; i < n; i++) {
        ids[i] = T->ids[i];
        track_cnt[i] = T->track_cnt[i];
        cur_pts[2 * i] = T->cur_pts[i].x;
        cur_pts[2 * i + 1] = T->cur_pts[i].y;
        un_pts[2 * i] = T->cur_un_pts[i].x;
        un_pts[2 * i + 1] = T->cur_un_pts[i].y;
        velocity[2 * i] = T->pts_velocity[i].x;
        velocity[2 * i + 1] = T->pts_velocity[i].y;
    }
}

void orc_tracker_stats(void* h, int* out5) {
    Tracker* T = (Tracker*)h;
    out5[0] = T->last_lk_in;
    out5[1] = T->last_lk_ok;
    out5[2] = T->last_ransac_in;
    out5[3] = T->last_ransac_ok;
    out5[4] = T->last_new;
}

void orc_lift_projective_pinhole(const orc_tracker_config* cfg, const double* px, int n, double* out_xy) {
    Tracker t;
    t.cfg = *cfg;
    for (int i = 0; i < n; i++) t.lift(px[2 * i], px[2 * i + 1], out_xy[2 * i], out_xy[2 * i + 1]);
}

}  // extern "C"

// Permutation produced by std::sort with setMask's comparator (feature_tracker.cpp:50-53) on the
// given track counts: perm[k] = original index of the k-th element after sorting.  The element type
// carries the same fields as the reference's pair<int, pair<Point2f,int>>; introsort's move sequence
// depends only on comparison outcomes, so the permutation equals the reference's.
extern "C" void orc_setmask_sort_perm(const int* track_cnt, int n, int* perm) {
    std::vector<std::pair<int, std::pair<Pt, int>>> v;
    for (int i = 0; i < n; i++) v.push_back(std::make_pair(track_cnt[i], std::make_pair(Pt{0, 0}, i)));
    std::sort(v.begin(), v.end(),
              [](const std::pair<int, std::pair<Pt, int>>& a, const std::pair<int, std::pair<Pt, int>>& b) {
                  return a.first > b.first;
              });
    for (int i = 0; i < n; i++) perm[i] = v[i].second.second;
}

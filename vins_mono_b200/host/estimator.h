// Host shim with the reference's class surface: drop-in for vins_estimator/src/estimator.h:25-139 as far as
// estimator_node.cpp and utility/visualization.cpp use it: processIMU, processImage(image, header), setReloFrame,
// clearState, setParameter and the public members Ps, Vs, Rs, Bas, Bgs, tic, ric, td, acc_0, gyr_0, g, Headers,
// solver_flag, marginalization_flag, frame_count.
//
// Types: with Eigen on the include path (the reference's build) the members ARE Eigen::Vector3d / Matrix3d /
// Matrix<double, 7, 1>, so the node's expressions (`tmp_Q = estimator.Rs[WINDOW_SIZE]`, `estimator.Ps[i].x()`,
// `estimator.ric[0] * p + estimator.tic[0]`) compile unchanged; without Eigen, small stand-ins with the same accessors
// (v[k], v.x(), m(r, c)) are used so that this header and its compile test need no third-party dependency.
#pragma once
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vinsb200/estimator.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(VINSB200_NO_EIGEN)
#define VINSB200_HAVE_EIGEN 1
#include <Eigen/Dense>
#endif
#endif

namespace vinsb200 {

#ifdef VINSB200_HAVE_EIGEN
using Vector3d = Eigen::Vector3d;
using Matrix3d = Eigen::Matrix3d;
using Vector7d = Eigen::Matrix<double, 7, 1>;
using Vector2d = Eigen::Vector2d;
#else
struct Vector3d {
    double v[3] = {0, 0, 0};
    Vector3d() {}
    Vector3d(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3d operator*(double s) const { return Vector3d(v[0] * s, v[1] * s, v[2] * s); }
    double operator()(int i) const { return v[i]; }
};
struct Matrix3d {
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double& operator()(int r, int c) { return m[3 * r + c]; }
    double operator()(int r, int c) const { return m[3 * r + c]; }
    Vector3d operator*(const Vector3d& p) const {
        return Vector3d(m[0] * p[0] + m[1] * p[1] + m[2] * p[2], m[3] * p[0] + m[4] * p[1] + m[5] * p[2], m[6] * p[0] + m[7] * p[1] + m[8] * p[2]);
    }
};
struct Vector2d {
    double v[2] = {0, 0};
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
struct Vector7d {
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
#endif

// FeatureManager as far as the publishers read it (feature_manager.h:19-72; visualization.cpp:228-296, :352-397)
struct FeaturePerFrame {
    Vector3d point;
    Vector2d uv;
};
struct FeaturePerId {
    int feature_id = 0, start_frame = 0, used_num = 0, solve_flag = 0;
    double estimated_depth = -1;
    std::vector<FeaturePerFrame> feature_per_frame;
    int endFrame() const { return start_frame + (int)feature_per_frame.size() - 1; }
};
struct FeatureManager {
    std::list<FeaturePerId> feature;
    int getFeatureCount() const {
        int cnt = 0;
        for (const auto& it : feature) cnt += (int)it.feature_per_frame.size() >= 2 && it.start_frame < window_size - 2;
        return cnt;
    }
    int window_size = 10;
};

class Estimator {
  public:
    enum SolverFlag { INITIAL, NON_LINEAR };
    enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };
    // the feature message of estimator_node.cpp:296-313: feature_id -> [(camera_id, x y z u v vx vy)]
    using ImageMap = std::map<int, std::vector<std::pair<int, Vector7d>>>;

    explicit Estimator(const ve_config& cfg) : cfg_(cfg) {
        if (ve_create(&cfg, &h_) != VE_OK) throw std::runtime_error("ve_create failed (no CUDA device?)");
        const int n = cfg.window_size + 1;
        Ps.resize(n); Vs.resize(n); Bas.resize(n); Bgs.resize(n); Rs.resize(n); Headers.resize(n, 0.0);
        refresh();
    }
    ~Estimator() { ve_destroy(h_); }
    Estimator(const Estimator&) = delete;
    Estimator& operator=(const Estimator&) = delete;

    void setParameter() {}  // parameters travel in ve_config (parameters.cpp:42-137); clearState re-applies them
    void clearState() { check(ve_clear_state(h_)); refresh(); }
    void processIMU(double dt, const Vector3d& linear_acceleration, const Vector3d& angular_velocity) {
        const double a[3] = {linear_acceleration[0], linear_acceleration[1], linear_acceleration[2]};
        const double w[3] = {angular_velocity[0], angular_velocity[1], angular_velocity[2]};
        check(ve_process_imu(h_, dt, a, w));
        refresh_imu();
    }
    // Estimator::processImage(image, header): Header is std_msgs::Header (header.stamp.toSec()) or any type with that shape
    template <class Header>
    void processImage(const ImageMap& image, const Header& header) { processImageAt(image, header.stamp.toSec()); }
    void processImageAt(const ImageMap& image, double stamp) {
        std::vector<int> ids;
        std::vector<double> d;
        for (const auto& kv : image) {  // NUM_OF_CAM = 1: the first (only) camera's observation
            ids.push_back(kv.first);
            for (int k = 0; k < 7; k++) d.push_back(kv.second[0].second[k]);
        }
        check(ve_process_image(h_, (int)ids.size(), ids.data(), d.data(), stamp));
        const int i = frame_count < (int)Headers.size() ? frame_count : (int)Headers.size() - 1;
        Headers[i] = stamp;
        refresh();
    }
    // Estimator::setReloFrame (estimator.cpp:1128-1146): match_points holds (x, y, feature id) triples (vector<Vector3d> in the node,
    // estimator_node.cpp:275-290).  The next processImage optimises relo_Pose with the matched landmarks and refresh() publishes
    // drift_correct_r / drift_correct_t / relo_relative_t / relo_relative_q / relo_relative_yaw / relo_frame_index as the
    // reference does (estimator.cpp:598-617, read by visualization.cpp:326-350 and estimator_node.cpp:327-328).
    template <class Points>
    void setReloFrame(double frame_stamp, int frame_index, Points& match_points, const Vector3d& relo_t, const Matrix3d& relo_r) {
        std::vector<double> mp;
        for (const auto& q : match_points)
            for (int k = 0; k < 3; k++) mp.push_back(q[k]);
        const double t[3] = {relo_t[0], relo_t[1], relo_t[2]};
        double r[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r[3 * i + j] = relo_r(i, j);
        relo_frame_stamp = frame_stamp;
        relo_frame_index = frame_index;
        const int rc = ve_set_relo_frame(h_, frame_stamp, frame_index, (int)(mp.size() / 3), mp.data(), t, r);
        check(rc);
        if (rc == 1) relocalization_info = true;
    }

    SolverFlag solver_flag = INITIAL;
    MarginalizationFlag marginalization_flag = MARGIN_OLD;
    int frame_count = 0;
    Vector3d g;
    std::vector<Vector3d> Ps, Vs, Bas, Bgs;
    std::vector<Matrix3d> Rs;
    Matrix3d ric[1];
    Vector3d tic[1];
    Vector3d acc_0, gyr_0;
    std::vector<double> Headers;  // stamps (the reference keeps std_msgs::Header objects)
    double td = 0;
    std::vector<Vector3d> key_poses;  // Ps[0 .. WINDOW_SIZE] while NON_LINEAR (estimator.cpp:208-210)
    FeatureManager f_manager;         // refreshed after every processImage
    bool relocalization_info = false;
    double relo_frame_stamp = 0, relo_frame_index = 0, relo_relative_yaw = 0;  // (the reference declares relo_frame_index as double)
    int relo_frame_local_index = 0;
    Matrix3d drift_correct_r;
    Vector3d drift_correct_t, relo_relative_t;
    double relo_relative_q[4] = {1, 0, 0, 0};  // w x y z
    ve_estimator* handle() { return h_; }

  private:
    void check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("vinsb200: ") + ve_last_error(h_));
    }
    static void q_to_R(const double* q /* w x y z */, Matrix3d& R) {
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - z * w); R(0, 2) = 2 * (x * z + y * w);
        R(1, 0) = 2 * (x * y + z * w); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - x * w);
        R(2, 0) = 2 * (x * z - y * w); R(2, 1) = 2 * (y * z + x * w); R(2, 2) = 1 - 2 * (x * x + y * y);
    }
    void refresh_imu() {
        double a[3], w[3], gg[3];
        ve_get_latest_imu(h_, a, w, gg);
        for (int k = 0; k < 3; k++) { acc_0[k] = a[k]; gyr_0[k] = w[k]; g[k] = gg[k]; }
    }
    void refresh() {
        std::vector<double> s(16 * Ps.size());
        ve_get_states(h_, s.data(), &td);
        for (size_t i = 0; i < Ps.size(); i++) {
            const double* o = &s[16 * i];
            for (int k = 0; k < 3; k++) { Ps[i][k] = o[k]; Vs[i][k] = o[7 + k]; Bas[i][k] = o[10 + k]; Bgs[i][k] = o[13 + k]; }
            q_to_R(o + 3, Rs[i]);
        }
        double t3[3], r9[9];
        ve_get_extrinsic(h_, t3, r9);
        for (int k = 0; k < 3; k++) tic[0][k] = t3[k];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) ric[0](r, c) = r9[3 * r + c];
        refresh_imu();
        int info[10];
        double c[2];
        ve_info(h_, info, c);
        key_poses.clear();
        if (info[0])
            for (size_t i = 0; i < Ps.size(); i++) key_poses.push_back(Ps[i]);
        {
            f_manager.window_size = cfg_.window_size;
            f_manager.feature.clear();
            const int nf = ve_get_features(h_, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            if (nf > 0) {
                const int cap_obs = nf * (cfg_.window_size + 1);
                std::vector<int> id(nf), st(nf), sf(nf), off(nf + 1);
                std::vector<double> dep(nf), obs(5 * (size_t)cap_obs);
                if (ve_get_features(h_, nf, cap_obs, id.data(), st.data(), sf.data(), dep.data(), off.data(), obs.data()) == nf)
                    for (int k = 0; k < nf; k++) {
                        FeaturePerId f;
                        f.feature_id = id[k]; f.start_frame = st[k]; f.solve_flag = sf[k]; f.estimated_depth = dep[k];
                        for (int o = off[k]; o < off[k + 1]; o++) {
                            FeaturePerFrame pf;
                            for (int c = 0; c < 3; c++) pf.point[c] = obs[5 * (size_t)o + c];
                            pf.uv[0] = obs[5 * (size_t)o + 3];
                            pf.uv[1] = obs[5 * (size_t)o + 4];
                            f.feature_per_frame.push_back(pf);
                        }
                        f.used_num = (int)f.feature_per_frame.size();
                        f_manager.feature.push_back(std::move(f));
                    }
            }
        }
        double ro[24];
        ve_get_relocalization(h_, ro);
        for (int r = 0; r < 3; r++) {
            for (int cc = 0; cc < 3; cc++) drift_correct_r(r, cc) = ro[3 * r + cc];
            drift_correct_t[r] = ro[9 + r];
            relo_relative_t[r] = ro[12 + r];
        }
        for (int k = 0; k < 4; k++) relo_relative_q[k] = ro[15 + k];
        relo_relative_yaw = ro[19];
        relocalization_info = ro[20] != 0.0;
        relo_frame_local_index = (int)ro[21];
        solver_flag = info[0] ? NON_LINEAR : INITIAL;
        frame_count = info[1];
        marginalization_flag = info[2] ? MARGIN_SECOND_NEW : MARGIN_OLD;
    }
    ve_config cfg_;
    ve_estimator* h_ = nullptr;
};

}  // namespace vinsb200

// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU twin of the vins_estimator sliding-window back end, following the reference line by line:
//   Estimator::processIMU / processImage / solveOdometry     vins_estimator/src/estimator.cpp:84-217, :473-484
//   vector2double / double2vector / failureDetection         :486-667
//   Estimator::optimization (problem assembly, solve, both marginalisation branches)   :670-1003
//   slideWindow / slideWindowNew / slideWindowOld            :1005-1126
//   FeatureManager (addFeatureCheckParallax, triangulate, setDepth, getDepthVector, removeBackShiftDepth,
//                   removeBack, removeFront, removeFailures, compensatedParallax2)   feature_manager.cpp:28-388
//   Utility::R2ypr / ypr2R                                    utility/utility.h:70-112
// Not restated: the one-shot initialisation (initial/*, estimator.cpp:218-471) — the
// window is seeded from a caller-provided trajectory instead, then the reference's own
// "repropagate + triangulate" tail of visualInitialAlign runs; relocalisation (estimator.cpp:769-801) and
// ESTIMATE_EXTRINSIC == 2.  The solver is be_solver.cpp (no wall-clock cap).
#include <cstdio>
#include <array>
#include <chrono>
#include <list>
#include <map>

#include "be_factors.h"
#include "be_solver.h"

namespace orc {

static V3 R2ypr(const M3& R) {
    const V3 n = R.col(0), o = R.col(1), a = R.col(2);
    const double y = atan2(n.y, n.x);
    const double p = atan2(-n.z, n.x * cos(y) + n.y * sin(y));
    const double r = atan2(a.x * sin(y) - a.y * cos(y), -o.x * sin(y) + o.y * cos(y));
    return V3(y, p, r) / M_PI * 180.0;
}
static M3 ypr2R(const V3& ypr) {
    const double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
    M3 Rz, Ry, Rx;
    Rz(0, 0) = cos(y); Rz(0, 1) = -sin(y); Rz(1, 0) = sin(y); Rz(1, 1) = cos(y); Rz(2, 2) = 1;
    Ry(0, 0) = cos(p); Ry(0, 2) = sin(p); Ry(1, 1) = 1; Ry(2, 0) = -sin(p); Ry(2, 2) = cos(p);
    Rx(0, 0) = 1; Rx(1, 1) = cos(r); Rx(1, 2) = -sin(r); Rx(2, 1) = sin(r); Rx(2, 2) = cos(r);
    return Rz * Ry * Rx;
}

struct FeaturePerFrame {
    V3 point;
    double u, v, vx, vy, cur_td;
};
struct FeaturePerId {
    int feature_id, start_frame;
    std::vector<FeaturePerFrame> feature_per_frame;
    int used_num = 0;
    double estimated_depth = -1.0;
    int solve_flag = 0;
    FeaturePerId(int id, int sf) : feature_id(id), start_frame(sf) {}
    int endFrame() const { return start_frame + (int)feature_per_frame.size() - 1; }
};
struct Obs {  // one entry of the feature message: id -> (x, y, z, u, v, vx, vy)
    int id;
    double d[7];
};

class FeatureManager {
  public:
    std::list<FeaturePerId> feature;
    int last_track_num = 0;
    int W = 10;
    double min_parallax = 10.0 / 460.0, init_depth = 5.0;
    const M3* Rs = nullptr;

    bool usable(FeaturePerId& it) {
        it.used_num = (int)it.feature_per_frame.size();
        return it.used_num >= 2 && it.start_frame < W - 2;
    }
    int getFeatureCount() {
        int c = 0;
        for (auto& it : feature) c += usable(it);
        return c;
    }
    bool addFeatureCheckParallax(int frame_count, const std::vector<Obs>& image, double td) {
        double parallax_sum = 0;
        int parallax_num = 0;
        last_track_num = 0;
        for (const Obs& o : image) {  // std::map iteration: ascending feature id
            FeaturePerFrame f{V3(o.d[0], o.d[1], o.d[2]), o.d[3], o.d[4], o.d[5], o.d[6], td};
            auto it = std::find_if(feature.begin(), feature.end(), [&](const FeaturePerId& x) { return x.feature_id == o.id; });
            if (it == feature.end()) {
                feature.push_back(FeaturePerId(o.id, frame_count));
                feature.back().feature_per_frame.push_back(f);
            } else {
                it->feature_per_frame.push_back(f);
                last_track_num++;
            }
        }
        if (frame_count < 2 || last_track_num < 20) return true;
        for (auto& it : feature)
            if (it.start_frame <= frame_count - 2 && it.start_frame + (int)it.feature_per_frame.size() - 1 >= frame_count - 1) {
                parallax_sum += compensatedParallax2(it, frame_count);
                parallax_num++;
            }
        if (parallax_num == 0) return true;
        return parallax_sum / parallax_num >= min_parallax;
    }
    double compensatedParallax2(const FeaturePerId& it, int frame_count) {
        const FeaturePerFrame& fi = it.feature_per_frame[frame_count - 2 - it.start_frame];
        const FeaturePerFrame& fj = it.feature_per_frame[frame_count - 1 - it.start_frame];
        const double u_j = fj.point.x, v_j = fj.point.y;
        const double dep_i = fi.point.z;
        const double du = fi.point.x / dep_i - u_j, dv = fi.point.y / dep_i - v_j;
        return std::max(0.0, std::sqrt(std::min(du * du + dv * dv, du * du + dv * dv)));
    }
    void setDepth(const std::vector<double>& x) {
        int k = -1;
        for (auto& it : feature) {
            if (!usable(it)) continue;
            it.estimated_depth = 1.0 / x[++k];
            it.solve_flag = it.estimated_depth < 0 ? 2 : 1;
        }
    }
    void clearDepth() {
        for (auto& it : feature) it.estimated_depth = -1;
    }
    std::vector<double> getDepthVector() {
        std::vector<double> d;
        for (auto& it : feature)
            if (usable(it)) d.push_back(1. / it.estimated_depth);
        return d;
    }
    void removeFailures() {
        for (auto it = feature.begin(); it != feature.end();) it = it->solve_flag == 2 ? feature.erase(it) : std::next(it);
    }
    void triangulate(const V3 Ps[], const V3& tic, const M3& ric) {
        for (auto& it : feature) {
            if (!usable(it)) continue;
            if (it.estimated_depth > 0) continue;
            int imu_i = it.start_frame, imu_j = imu_i - 1;
            Mat A(2 * (int)it.feature_per_frame.size(), 4);
            int row = 0;
            const V3 t0 = Ps[imu_i] + Rs[imu_i] * tic;
            const M3 R0 = Rs[imu_i] * ric;
            for (auto& f : it.feature_per_frame) {
                imu_j++;
                const V3 t1 = Ps[imu_j] + Rs[imu_j] * tic;
                const M3 R1 = Rs[imu_j] * ric;
                const V3 t = R0.T() * (t1 - t0);
                const M3 R = R0.T() * R1;
                const M3 Rt = R.T();
                const V3 mt = -(Rt * t);
                double P[3][4];
                for (int i = 0; i < 3; i++) {
                    for (int j = 0; j < 3; j++) P[i][j] = Rt(i, j);
                    P[i][3] = mt[i];
                }
                const V3 fn = f.point.normalized();
                for (int c = 0; c < 4; c++) {
                    A(row, c) = fn.x * P[2][c] - fn.z * P[0][c];
                    A(row + 1, c) = fn.y * P[2][c] - fn.z * P[1][c];
                }
                row += 2;
            }
            double v[4];
            smallest_right_singular_vector4(A, v);
            it.estimated_depth = v[2] / v[3];
            if (it.estimated_depth < 0.1) it.estimated_depth = init_depth;
        }
    }
    void removeBackShiftDepth(const M3& marg_R, const V3& marg_P, const M3& new_R, const V3& new_P) {
        for (auto it = feature.begin(); it != feature.end();) {
            auto nx = std::next(it);
            if (it->start_frame != 0)
                it->start_frame--;
            else {
                const V3 uv_i = it->feature_per_frame[0].point;
                it->feature_per_frame.erase(it->feature_per_frame.begin());
                if (it->feature_per_frame.size() < 2) {
                    feature.erase(it);
                } else {
                    const V3 pts_i = uv_i * it->estimated_depth;
                    const V3 w_pts_i = marg_R * pts_i + marg_P;
                    const V3 pts_j = new_R.T() * (w_pts_i - new_P);
                    it->estimated_depth = pts_j.z > 0 ? pts_j.z : init_depth;
                }
            }
            it = nx;
        }
    }
    void removeBack() {
        for (auto it = feature.begin(); it != feature.end();) {
            auto nx = std::next(it);
            if (it->start_frame != 0)
                it->start_frame--;
            else {
                it->feature_per_frame.erase(it->feature_per_frame.begin());
                if (it->feature_per_frame.empty()) feature.erase(it);
            }
            it = nx;
        }
    }
    void removeFront(int frame_count) {
        for (auto it = feature.begin(); it != feature.end();) {
            auto nx = std::next(it);
            if (it->start_frame == frame_count)
                it->start_frame--;
            else {
                const int j = W - 1 - it->start_frame;
                if (it->endFrame() >= frame_count - 1) {
                    it->feature_per_frame.erase(it->feature_per_frame.begin() + j);
                    if (it->feature_per_frame.empty()) feature.erase(it);
                }
            }
            it = nx;
        }
    }
};

struct SeedState {
    double t;
    V3 P, V;
    M3 R;
};

class Estimator {
  public:
    enum SolverFlag { INITIAL, NON_LINEAR };
    enum MargFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };
    BeConfig cfg;
    int W;
    SolverFlag solver_flag = INITIAL;
    MargFlag marginalization_flag = MARGIN_OLD;
    V3 g;
    M3 ric;
    V3 tic;
    std::vector<V3> Ps, Vs, Bas, Bgs;
    std::vector<M3> Rs;
    double td = 0;
    M3 back_R0, last_R, last_R0;
    V3 back_P0, last_P, last_P0;
    std::vector<double> Headers;
    std::vector<std::unique_ptr<IntegrationBase>> pre_integrations;
    V3 acc_0, gyr_0;
    std::vector<std::vector<double>> dt_buf;
    std::vector<std::vector<V3>> linear_acceleration_buf, angular_velocity_buf;
    int frame_count = 0;
    FeatureManager f_manager;
    bool first_imu = false, failure_occur = false;
    std::vector<std::array<double, 7>> para_Pose;
    std::vector<std::array<double, 9>> para_SpeedBias;
    std::vector<std::array<double, 1>> para_Feature;
    std::array<double, 7> para_Ex_Pose;
    std::array<double, 1> para_Td;
    std::unique_ptr<MarginalizationInfo> last_marginalization_info;
    std::vector<double*> last_marginalization_parameter_blocks;
    std::vector<SeedState> seeds;
    V3 seed_ba, seed_bg;
    CauchyLoss loss{1.0};
    // relocalisation (estimator.h:125-138)
    bool relocalization_info = false;
    double relo_frame_stamp = 0;
    int relo_frame_index = 0, relo_frame_local_index = 0;
    std::vector<V3> match_points;  // x, y, feature id
    std::array<double, 7> relo_Pose{};
    M3 drift_correct_r = M3::Identity(), prev_relo_r = M3::Identity();
    V3 drift_correct_t, prev_relo_t, relo_relative_t;
    Quat relo_relative_q;
    double relo_relative_yaw = 0;
    int n_relo_factors = 0, n_relo_solves = 0;
    SolveSummary last_summary;
    int n_solves = 0, n_reboots = 0, last_landmarks = 0, last_visual = 0;
    // test hook: called with (columns, residuals) right before Solve of solve number probe_solve (-1: every solve)
    bool fast_eigen = false;  // see MarginalizationInfo::eigen_ql
    double prof_solve_s = 0, prof_marg_s = 0;  // accumulated wall time of Solve() and of the marginalisation (profiling)
    void (*probe_cb)(int, int) = nullptr;
    int probe_solve = -1;
    Problem* probe_problem = nullptr;
    double sqrt_info_scale;

    explicit Estimator(const BeConfig& c) : cfg(c), W(c.window_size) {
        const int n = W + 1;
        Ps.resize(n); Vs.resize(n); Bas.resize(n); Bgs.resize(n); Rs.resize(n); Headers.assign(n, 0.0);
        pre_integrations.resize(n);
        dt_buf.resize(n); linear_acceleration_buf.resize(n); angular_velocity_buf.resize(n);
        para_Pose.resize(n); para_SpeedBias.resize(n); para_Feature.resize(4096);
        f_manager.Rs = Rs.data();
        f_manager.W = W;
        f_manager.min_parallax = c.min_parallax;
        f_manager.init_depth = c.init_depth;
        clearState();
        setParameter();
    }
    void setParameter() {
        tic = cfg.tic;
        ric = cfg.ric;
        sqrt_info_scale = cfg.focal_length / 1.5;
        td = cfg.td;
        g = cfg.G;
    }
    void clearState() {
        for (int i = 0; i <= W; i++) {
            Rs[i] = M3::Identity();
            Ps[i] = Vs[i] = Bas[i] = Bgs[i] = V3();
            dt_buf[i].clear(); linear_acceleration_buf[i].clear(); angular_velocity_buf[i].clear();
            pre_integrations[i].reset();
        }
        tic = V3();
        ric = M3::Identity();
        solver_flag = INITIAL;
        first_imu = false;
        frame_count = 0;
        td = cfg.td;
        last_marginalization_info.reset();
        last_marginalization_parameter_blocks.clear();
        f_manager.feature.clear();
        failure_occur = false;
        relocalization_info = false;
        drift_correct_r = M3::Identity();
        drift_correct_t = V3();
    }
    // Estimator::setReloFrame (estimator.cpp:1128-1146)
    void setReloFrame(double stamp, int index, const std::vector<V3>& pts, const V3& relo_t, const M3& relo_r) {
        relo_frame_stamp = stamp;
        relo_frame_index = index;
        match_points = pts;
        prev_relo_t = relo_t;
        prev_relo_r = relo_r;
        for (int i = 0; i < W; i++)
            if (relo_frame_stamp == Headers[i]) {
                relo_frame_local_index = i;
                relocalization_info = true;
                relo_Pose = para_Pose[i];
            }
    }
    void processIMU(double dt, const V3& linear_acceleration, const V3& angular_velocity) {
        if (!first_imu) {
            first_imu = true;
            acc_0 = linear_acceleration;
            gyr_0 = angular_velocity;
        }
        if (!pre_integrations[frame_count])
            pre_integrations[frame_count].reset(new IntegrationBase(acc_0, gyr_0, Bas[frame_count], Bgs[frame_count], cfg));
        if (frame_count != 0) {
            pre_integrations[frame_count]->push_back(dt, linear_acceleration, angular_velocity);
            dt_buf[frame_count].push_back(dt);
            linear_acceleration_buf[frame_count].push_back(linear_acceleration);
            angular_velocity_buf[frame_count].push_back(angular_velocity);
            const int j = frame_count;
            const V3 un_acc_0 = Rs[j] * (acc_0 - Bas[j]) - g;
            const V3 un_gyr = 0.5 * (gyr_0 + angular_velocity) - Bgs[j];
            Rs[j] = Rs[j] * deltaQ(un_gyr * dt).R();
            const V3 un_acc_1 = Rs[j] * (linear_acceleration - Bas[j]) - g;
            const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
            Ps[j] += dt * Vs[j] + 0.5 * dt * dt * un_acc;
            Vs[j] += dt * un_acc;
        }
        acc_0 = linear_acceleration;
        gyr_0 = angular_velocity;
    }
    // Stand-in for initialStructure(): states of the window frames come from the seed trajectory; then the
    // tail of visualInitialAlign (estimator.cpp:400-470): repropagate with the gyro bias, reset depths, triangulate.
    bool initialFromSeed() {
        for (int i = 0; i <= W; i++) {
            const SeedState* s = nullptr;
            for (auto& c : seeds)
                if (std::fabs(c.t - Headers[i]) < 1e-6) s = &c;
            if (!s) return false;
            Ps[i] = s->P; Rs[i] = s->R; Vs[i] = s->V; Bas[i] = seed_ba; Bgs[i] = seed_bg;
        }
        for (int i = 0; i <= W; i++) pre_integrations[i]->repropagate(Bas[i], Bgs[i]);
        f_manager.clearDepth();
        f_manager.triangulate(Ps.data(), tic, ric);
        return true;
    }
    void processImage(const std::vector<Obs>& image, double stamp) {
        marginalization_flag = f_manager.addFeatureCheckParallax(frame_count, image, td) ? MARGIN_OLD : MARGIN_SECOND_NEW;
        Headers[frame_count] = stamp;
        if (solver_flag == INITIAL) {
            if (frame_count == W) {
                if (initialFromSeed()) {
                    solver_flag = NON_LINEAR;
                    solveOdometry();
                    slideWindow();
                    f_manager.removeFailures();
                    last_R = Rs[W]; last_P = Ps[W]; last_R0 = Rs[0]; last_P0 = Ps[0];
                } else
                    slideWindow();
            } else
                frame_count++;
        } else {
            solveOdometry();
            if (failureDetection()) {
                failure_occur = true;
                clearState();
                setParameter();
                n_reboots++;
                return;
            }
            slideWindow();
            f_manager.removeFailures();
            last_R = Rs[W]; last_P = Ps[W]; last_R0 = Rs[0]; last_P0 = Ps[0];
        }
    }
    void solveOdometry() {
        if (frame_count < W) return;
        if (solver_flag == NON_LINEAR) {
            f_manager.triangulate(Ps.data(), tic, ric);
            optimization();
        }
    }
    void vector2double() {
        for (int i = 0; i <= W; i++) {
            const Quat q = Quat::FromR(Rs[i]);
            para_Pose[i] = {Ps[i].x, Ps[i].y, Ps[i].z, q.x, q.y, q.z, q.w};
            para_SpeedBias[i] = {Vs[i].x, Vs[i].y, Vs[i].z, Bas[i].x, Bas[i].y, Bas[i].z, Bgs[i].x, Bgs[i].y, Bgs[i].z};
        }
        const Quat q = Quat::FromR(ric);
        para_Ex_Pose = {tic.x, tic.y, tic.z, q.x, q.y, q.z, q.w};
        const std::vector<double> dep = f_manager.getDepthVector();
        for (size_t i = 0; i < dep.size(); i++) para_Feature[i][0] = dep[i];
        if (cfg.estimate_td) para_Td[0] = td;
    }
    void double2vector() {
        V3 origin_R0 = R2ypr(Rs[0]);
        V3 origin_P0 = Ps[0];
        if (failure_occur) {
            origin_R0 = R2ypr(last_R0);
            origin_P0 = last_P0;
            failure_occur = false;
        }
        const M3 R00 = Quat(para_Pose[0][6], para_Pose[0][3], para_Pose[0][4], para_Pose[0][5]).R();
        const V3 origin_R00 = R2ypr(R00);
        const double y_diff = origin_R0.x - origin_R00.x;
        M3 rot_diff = ypr2R(V3(y_diff, 0, 0));
        if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0)
            rot_diff = Rs[0] * R00.T();
        for (int i = 0; i <= W; i++) {
            Rs[i] = rot_diff * Quat(para_Pose[i][6], para_Pose[i][3], para_Pose[i][4], para_Pose[i][5]).normalized().R();
            Ps[i] = rot_diff * V3(para_Pose[i][0] - para_Pose[0][0], para_Pose[i][1] - para_Pose[0][1], para_Pose[i][2] - para_Pose[0][2]) + origin_P0;
            Vs[i] = rot_diff * V3(para_SpeedBias[i][0], para_SpeedBias[i][1], para_SpeedBias[i][2]);
            Bas[i] = V3(para_SpeedBias[i][3], para_SpeedBias[i][4], para_SpeedBias[i][5]);
            Bgs[i] = V3(para_SpeedBias[i][6], para_SpeedBias[i][7], para_SpeedBias[i][8]);
        }
        tic = V3(para_Ex_Pose[0], para_Ex_Pose[1], para_Ex_Pose[2]);
        ric = Quat(para_Ex_Pose[6], para_Ex_Pose[3], para_Ex_Pose[4], para_Ex_Pose[5]).R();
        std::vector<double> dep = f_manager.getDepthVector();
        for (size_t i = 0; i < dep.size(); i++) dep[i] = para_Feature[i][0];
        f_manager.setDepth(dep);
        if (cfg.estimate_td) td = para_Td[0];
        if (relocalization_info) {  // estimator.cpp:598-617
            const M3 relo_r = rot_diff * Quat(relo_Pose[6], relo_Pose[3], relo_Pose[4], relo_Pose[5]).normalized().R();
            const V3 relo_t = rot_diff * V3(relo_Pose[0] - para_Pose[0][0], relo_Pose[1] - para_Pose[0][1], relo_Pose[2] - para_Pose[0][2]) + origin_P0;
            const double drift_correct_yaw = R2ypr(prev_relo_r).x - R2ypr(relo_r).x;
            drift_correct_r = ypr2R(V3(drift_correct_yaw, 0, 0));
            drift_correct_t = prev_relo_t - drift_correct_r * relo_t;
            relo_relative_t = relo_r.T() * (Ps[relo_frame_local_index] - relo_t);
            relo_relative_q = Quat::FromR(relo_r.T() * Rs[relo_frame_local_index]);
            double a = R2ypr(Rs[relo_frame_local_index]).x - R2ypr(relo_r).x;  // Utility::normalizeAngle (degrees)
            if (a > 180.0) a -= 360.0;
            else if (a < -180.0) a += 360.0;
            relo_relative_yaw = a;
            relocalization_info = false;
            n_relo_solves++;
        }
    }
    bool failureDetection() {
        if (Bas[W].norm() > 2.5) return true;
        if (Bgs[W].norm() > 1.0) return true;
        const V3 tmp_P = Ps[W];
        if ((tmp_P - last_P).norm() > 5) return true;
        if (std::fabs(tmp_P.z - last_P.z) > 1) return true;
        return false;
    }
    std::shared_ptr<CostFunction> make_visual(const FeaturePerId& f, const FeaturePerFrame& fj) {
        const FeaturePerFrame& f0 = f.feature_per_frame[0];
        if (cfg.estimate_td)
            return std::make_shared<ProjectionTdFactor>(f0.point, fj.point, f0.vx, f0.vy, fj.vx, fj.vy, f0.cur_td, fj.cur_td,
                                                        f0.v, fj.v, sqrt_info_scale, cfg.tr, cfg.row);
        return std::make_shared<ProjectionFactor>(f0.point, fj.point, sqrt_info_scale);
    }
    void optimization() {
        Problem problem;
        for (int i = 0; i <= W; i++) {
            problem.AddParameterBlock(para_Pose[i].data(), 7, true);
            problem.AddParameterBlock(para_SpeedBias[i].data(), 9, false);
        }
        problem.AddParameterBlock(para_Ex_Pose.data(), 7, true);
        if (!cfg.estimate_extrinsic) problem.SetParameterBlockConstant(para_Ex_Pose.data());
        if (cfg.estimate_td) problem.AddParameterBlock(para_Td.data(), 1, false);
        vector2double();
        if (last_marginalization_info)
            problem.AddResidualBlock(std::make_shared<MarginalizationFactor>(last_marginalization_info.get()), nullptr,
                                     last_marginalization_parameter_blocks);
        for (int i = 0; i < W; i++) {
            const int j = i + 1;
            if (pre_integrations[j]->sum_dt > 10.0) continue;
            problem.AddResidualBlock(std::make_shared<IMUFactor>(pre_integrations[j].get()), nullptr,
                                     {para_Pose[i].data(), para_SpeedBias[i].data(), para_Pose[j].data(), para_SpeedBias[j].data()});
        }
        int f_m_cnt = 0, feature_index = -1;
        for (auto& it : f_manager.feature) {
            if (!f_manager.usable(it)) continue;
            ++feature_index;
            const int imu_i = it.start_frame;
            int imu_j = imu_i - 1;
            for (auto& fj : it.feature_per_frame) {
                imu_j++;
                if (imu_i == imu_j) continue;
                std::vector<double*> pb = {para_Pose[imu_i].data(), para_Pose[imu_j].data(), para_Ex_Pose.data(), para_Feature[feature_index].data()};
                if (cfg.estimate_td) pb.push_back(para_Td.data());
                problem.AddResidualBlock(make_visual(it, fj), &loss, pb);
                f_m_cnt++;
            }
        }
        last_landmarks = feature_index + 1;
        last_visual = f_m_cnt;
        n_relo_factors = 0;
        if (relocalization_info) {  // estimator.cpp:769-801
            problem.AddParameterBlock(relo_Pose.data(), 7, true);
            size_t retrive_feature_index = 0;
            int fi = -1;
            for (auto& it : f_manager.feature) {
                if (!f_manager.usable(it)) continue;
                ++fi;
                const int start = it.start_frame;
                if (start <= relo_frame_local_index) {
                    // (the reference walks match_points without a bound; past its end nothing can match)
                    while (retrive_feature_index < match_points.size() && (int)match_points[retrive_feature_index].z < it.feature_id)
                        retrive_feature_index++;
                    if (retrive_feature_index < match_points.size() && (int)match_points[retrive_feature_index].z == it.feature_id) {
                        const V3 pts_j(match_points[retrive_feature_index].x, match_points[retrive_feature_index].y, 1.0);
                        const V3 pts_i = it.feature_per_frame[0].point;
                        problem.AddResidualBlock(std::make_shared<ProjectionFactor>(pts_i, pts_j, sqrt_info_scale), &loss,
                                                 {para_Pose[start].data(), relo_Pose.data(), para_Ex_Pose.data(), para_Feature[fi].data()});
                        retrive_feature_index++;
                        n_relo_factors++;
                    }
                }
            }
        }
        if (probe_cb && (probe_solve < 0 || probe_solve == n_solves)) {  // tests: hand the problem to an independent optimiser
            probe_problem = &problem;
            probe_cb(ProbeColumns(problem), ProbeResiduals(problem));
            probe_problem = nullptr;
        }
        const auto t_solve0 = std::chrono::steady_clock::now();
        last_summary = Solve(problem, cfg.num_iterations);
        prof_solve_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_solve0).count();
        n_solves++;
        double2vector();
        if (marginalization_flag == MARGIN_OLD) {
            auto mi = std::unique_ptr<MarginalizationInfo>(new MarginalizationInfo());
            mi->eigen_ql = fast_eigen;
            vector2double();
            if (last_marginalization_info) {
                std::vector<int> drop_set;
                for (int i = 0; i < (int)last_marginalization_parameter_blocks.size(); i++)
                    if (last_marginalization_parameter_blocks[i] == para_Pose[0].data() ||
                        last_marginalization_parameter_blocks[i] == para_SpeedBias[0].data())
                        drop_set.push_back(i);
                mi->addResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
                    std::make_shared<MarginalizationFactor>(last_marginalization_info.get()), nullptr,
                    last_marginalization_parameter_blocks, drop_set));
            }
            if (pre_integrations[1]->sum_dt < 10.0)
                mi->addResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
                    std::make_shared<IMUFactor>(pre_integrations[1].get()), nullptr,
                    std::vector<double*>{para_Pose[0].data(), para_SpeedBias[0].data(), para_Pose[1].data(), para_SpeedBias[1].data()},
                    std::vector<int>{0, 1}));
            {
                int fi = -1;
                for (auto& it : f_manager.feature) {
                    if (!f_manager.usable(it)) continue;
                    ++fi;
                    const int imu_i = it.start_frame;
                    int imu_j = imu_i - 1;
                    if (imu_i != 0) continue;
                    for (auto& fj : it.feature_per_frame) {
                        imu_j++;
                        if (imu_i == imu_j) continue;
                        std::vector<double*> pb = {para_Pose[imu_i].data(), para_Pose[imu_j].data(), para_Ex_Pose.data(), para_Feature[fi].data()};
                        if (cfg.estimate_td) pb.push_back(para_Td.data());
                        mi->addResidualBlockInfo(std::make_shared<ResidualBlockInfo>(make_visual(it, fj), &loss, pb, std::vector<int>{0, 3}));
                    }
                }
            }
            const auto t_m0 = std::chrono::steady_clock::now();
            mi->preMarginalize();
            mi->marginalize();
            prof_marg_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_m0).count();
            std::map<double*, double*> addr_shift;
            for (int i = 1; i <= W; i++) {
                addr_shift[para_Pose[i].data()] = para_Pose[i - 1].data();
                addr_shift[para_SpeedBias[i].data()] = para_SpeedBias[i - 1].data();
            }
            addr_shift[para_Ex_Pose.data()] = para_Ex_Pose.data();
            if (cfg.estimate_td) addr_shift[para_Td.data()] = para_Td.data();
            std::vector<double*> pbs = mi->getParameterBlocks(addr_shift);
            last_marginalization_info = std::move(mi);
            last_marginalization_parameter_blocks = pbs;
        } else {
            if (last_marginalization_info &&
                std::count(last_marginalization_parameter_blocks.begin(), last_marginalization_parameter_blocks.end(), para_Pose[W - 1].data())) {
                auto mi = std::unique_ptr<MarginalizationInfo>(new MarginalizationInfo());
                mi->eigen_ql = fast_eigen;
            mi->eigen_ql = fast_eigen;
                vector2double();
                std::vector<int> drop_set;
                for (int i = 0; i < (int)last_marginalization_parameter_blocks.size(); i++)
                    if (last_marginalization_parameter_blocks[i] == para_Pose[W - 1].data()) drop_set.push_back(i);
                mi->addResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
                    std::make_shared<MarginalizationFactor>(last_marginalization_info.get()), nullptr,
                    last_marginalization_parameter_blocks, drop_set));
                const auto t_m0 = std::chrono::steady_clock::now();
                mi->preMarginalize();
                mi->marginalize();
                prof_marg_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_m0).count();
                std::map<double*, double*> addr_shift;
                for (int i = 0; i <= W; i++) {
                    if (i == W - 1) continue;
                    if (i == W) {
                        addr_shift[para_Pose[i].data()] = para_Pose[i - 1].data();
                        addr_shift[para_SpeedBias[i].data()] = para_SpeedBias[i - 1].data();
                    } else {
                        addr_shift[para_Pose[i].data()] = para_Pose[i].data();
                        addr_shift[para_SpeedBias[i].data()] = para_SpeedBias[i].data();
                    }
                }
                addr_shift[para_Ex_Pose.data()] = para_Ex_Pose.data();
                if (cfg.estimate_td) addr_shift[para_Td.data()] = para_Td.data();
                std::vector<double*> pbs = mi->getParameterBlocks(addr_shift);
                last_marginalization_info = std::move(mi);
                last_marginalization_parameter_blocks = pbs;
            }
        }
    }
    void slideWindow() {
        if (marginalization_flag == MARGIN_OLD) {
            back_R0 = Rs[0];
            back_P0 = Ps[0];
            if (frame_count == W) {
                for (int i = 0; i < W; i++) {
                    std::swap(Rs[i], Rs[i + 1]);
                    std::swap(pre_integrations[i], pre_integrations[i + 1]);
                    dt_buf[i].swap(dt_buf[i + 1]);
                    linear_acceleration_buf[i].swap(linear_acceleration_buf[i + 1]);
                    angular_velocity_buf[i].swap(angular_velocity_buf[i + 1]);
                    Headers[i] = Headers[i + 1];
                    std::swap(Ps[i], Ps[i + 1]);
                    std::swap(Vs[i], Vs[i + 1]);
                    std::swap(Bas[i], Bas[i + 1]);
                    std::swap(Bgs[i], Bgs[i + 1]);
                }
                Headers[W] = Headers[W - 1];
                Ps[W] = Ps[W - 1]; Vs[W] = Vs[W - 1]; Rs[W] = Rs[W - 1]; Bas[W] = Bas[W - 1]; Bgs[W] = Bgs[W - 1];
                pre_integrations[W].reset(new IntegrationBase(acc_0, gyr_0, Bas[W], Bgs[W], cfg));
                dt_buf[W].clear(); linear_acceleration_buf[W].clear(); angular_velocity_buf[W].clear();
                slideWindowOld();
            }
        } else {
            if (frame_count == W) {
                for (size_t i = 0; i < dt_buf[frame_count].size(); i++) {
                    const double tmp_dt = dt_buf[frame_count][i];
                    const V3 a = linear_acceleration_buf[frame_count][i], w = angular_velocity_buf[frame_count][i];
                    pre_integrations[frame_count - 1]->push_back(tmp_dt, a, w);
                    dt_buf[frame_count - 1].push_back(tmp_dt);
                    linear_acceleration_buf[frame_count - 1].push_back(a);
                    angular_velocity_buf[frame_count - 1].push_back(w);
                }
                Headers[frame_count - 1] = Headers[frame_count];
                Ps[frame_count - 1] = Ps[frame_count]; Vs[frame_count - 1] = Vs[frame_count]; Rs[frame_count - 1] = Rs[frame_count];
                Bas[frame_count - 1] = Bas[frame_count]; Bgs[frame_count - 1] = Bgs[frame_count];
                pre_integrations[W].reset(new IntegrationBase(acc_0, gyr_0, Bas[W], Bgs[W], cfg));
                dt_buf[W].clear(); linear_acceleration_buf[W].clear(); angular_velocity_buf[W].clear();
                f_manager.removeFront(frame_count);  // slideWindowNew
            }
        }
    }
    void slideWindowOld() {
        if (solver_flag == NON_LINEAR) {
            const M3 R0 = back_R0 * ric, R1 = Rs[0] * ric;
            const V3 P0 = back_P0 + back_R0 * tic, P1 = Ps[0] + Rs[0] * tic;
            f_manager.removeBackShiftDepth(R0, P0, R1, P1);
        } else
            f_manager.removeBack();
    }
};

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C interface for tests / benchmarks (ctypes).
using namespace orc;

extern "C" {

struct orc_be_config {
    int window_size, num_iterations, estimate_extrinsic, estimate_td;
    double focal_length, min_parallax, acc_n, gyr_n, acc_w, gyr_w, g_norm, init_depth, td, tr, row;
    double tic[3], ric[9];
};

static BeConfig to_cfg(const orc_be_config* c) {
    BeConfig b;
    b.window_size = c->window_size; b.num_iterations = c->num_iterations;
    b.estimate_extrinsic = c->estimate_extrinsic; b.estimate_td = c->estimate_td;
    b.focal_length = c->focal_length; b.min_parallax = c->min_parallax;
    b.acc_n = c->acc_n; b.gyr_n = c->gyr_n; b.acc_w = c->acc_w; b.gyr_w = c->gyr_w;
    b.G = V3(0, 0, c->g_norm); b.init_depth = c->init_depth; b.td = c->td; b.tr = c->tr; b.row = c->row;
    b.tic = V3(c->tic[0], c->tic[1], c->tic[2]);
    for (int i = 0; i < 9; i++) b.ric.m[i] = c->ric[i];
    return b;
}

void* orc_est_create(const orc_be_config* c) { return new Estimator(to_cfg(c)); }
void orc_est_destroy(void* h) { delete (Estimator*)h; }
// seed trajectory rows: t, p(3), q(wxyz), v(3)
void orc_est_set_seed(void* h, int n, const double* rows, const double* ba, const double* bg) {
    Estimator* e = (Estimator*)h;
    e->seeds.clear();
    for (int i = 0; i < n; i++) {
        const double* r = rows + 11 * i;
        SeedState s;
        s.t = r[0];
        s.P = V3(r[1], r[2], r[3]);
        s.R = Quat(r[4], r[5], r[6], r[7]).normalized().R();
        s.V = V3(r[8], r[9], r[10]);
        e->seeds.push_back(s);
    }
    e->seed_ba = V3(ba[0], ba[1], ba[2]);
    e->seed_bg = V3(bg[0], bg[1], bg[2]);
}
void orc_est_process_imu(void* h, double dt, const double* acc, const double* gyr) {
    ((Estimator*)h)->processIMU(dt, V3(acc[0], acc[1], acc[2]), V3(gyr[0], gyr[1], gyr[2]));
}
// feature message: n points, ids ascending not required (sorted here like std::map), xyz_uv_vel = n x 7
void orc_est_process_image(void* h, int n, const int* ids, const double* xyz_uv_vel, double stamp) {
    std::vector<Obs> img(n);
    for (int i = 0; i < n; i++) {
        img[i].id = ids[i];
        std::memcpy(img[i].d, xyz_uv_vel + 7 * i, 7 * sizeof(double));
    }
    std::stable_sort(img.begin(), img.end(), [](const Obs& a, const Obs& b) { return a.id < b.id; });
    ((Estimator*)h)->processImage(img, stamp);
}
// out: solver_flag, frame_count, marginalization_flag, n_solves, n_reboots, landmarks, visual factors, iterations,
//      successful steps, termination
void orc_est_info(void* h, int* out10, double* costs2) {
    Estimator* e = (Estimator*)h;
    out10[0] = e->solver_flag; out10[1] = e->frame_count; out10[2] = e->marginalization_flag; out10[3] = e->n_solves;
    out10[4] = e->n_reboots; out10[5] = e->last_landmarks; out10[6] = e->last_visual; out10[7] = e->last_summary.iterations;
    out10[8] = e->last_summary.successful_steps; out10[9] = e->last_summary.termination;
    costs2[0] = e->last_summary.initial_cost; costs2[1] = e->last_summary.final_cost;
}
// window states: per frame p(3) q(wxyz) v(3) ba(3) bg(3) = 16 doubles, (W+1) frames; plus td
void orc_est_states(void* h, double* out, double* td) {
    Estimator* e = (Estimator*)h;
    for (int i = 0; i <= e->W; i++) {
        double* o = out + 16 * i;
        const Quat q = Quat::FromR(e->Rs[i]);
        o[0] = e->Ps[i].x; o[1] = e->Ps[i].y; o[2] = e->Ps[i].z; o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z;
        o[7] = e->Vs[i].x; o[8] = e->Vs[i].y; o[9] = e->Vs[i].z;
        o[10] = e->Bas[i].x; o[11] = e->Bas[i].y; o[12] = e->Bas[i].z;
        o[13] = e->Bgs[i].x; o[14] = e->Bgs[i].y; o[15] = e->Bgs[i].z;
    }
    if (td) *td = e->td;
}
// prior in information form: n, then A' (n x n) and b' (n) of the last marginalisation (tests compare these)
int orc_est_prior(void* h, int cap, double* A, double* b) {
    Estimator* e = (Estimator*)h;
    if (!e->last_marginalization_info) return 0;
    const int n = e->last_marginalization_info->n;
    if (n > cap) return -n;
    std::memcpy(A, e->last_marginalization_info->A_debug.d.data(), (size_t)n * n * sizeof(double));
    std::memcpy(b, e->last_marginalization_info->b_debug.data(), n * sizeof(double));
    return n;
}
// Estimator::clearState + setParameter, as the node does on a restart message (estimator_node.cpp:186-203)
// Estimator::setReloFrame: match_points = n x (x, y, feature id), relo_r row major
void orc_est_set_relo_frame(void* h, double stamp, int index, int n, const double* match_points, const double* relo_t, const double* relo_r) {
    Estimator* e = (Estimator*)h;
    std::vector<V3> pts(n);
    for (int i = 0; i < n; i++) pts[i] = V3(match_points[3 * i], match_points[3 * i + 1], match_points[3 * i + 2]);
    M3 R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = relo_r[3 * i + j];
    e->setReloFrame(stamp, index, pts, V3(relo_t[0], relo_t[1], relo_t[2]), R);
}
// out24: drift_correct_r 9 | drift_correct_t 3 | relo_relative_t 3 | relo_relative_q wxyz 4 | relo_relative_yaw | pending flag |
// relo_frame_local_index | factors of the last solve | solves that carried relocalisation factors
void orc_est_relo(void* h, double* out) {
    Estimator* e = (Estimator*)h;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[3 * i + j] = e->drift_correct_r(i, j);
    out[9] = e->drift_correct_t.x; out[10] = e->drift_correct_t.y; out[11] = e->drift_correct_t.z;
    out[12] = e->relo_relative_t.x; out[13] = e->relo_relative_t.y; out[14] = e->relo_relative_t.z;
    out[15] = e->relo_relative_q.w; out[16] = e->relo_relative_q.x; out[17] = e->relo_relative_q.y; out[18] = e->relo_relative_q.z;
    out[19] = e->relo_relative_yaw;
    out[20] = e->relocalization_info ? 1.0 : 0.0;
    out[21] = e->relo_frame_local_index;
    out[22] = e->n_relo_factors;
    out[23] = e->n_relo_solves;
}

void orc_est_clear_state(void* h) {
    Estimator* e = (Estimator*)h;
    e->clearState();
    e->setParameter();
}
int orc_est_feature_count(void* h) { return (int)((Estimator*)h)->f_manager.feature.size(); }

// ---- single-factor known-answer access -------------------------------------------------------------
// IntegrationBase: construct with sample 0, push n-1 samples; out: sum_dt, dp(3), dq(wxyz), dv(3), then
// jacobian 15x15, covariance 15x15, sqrt_info 15x15 (row-major)
void orc_preintegrate(const orc_be_config* c, const double* ba, const double* bg, int n, const double* dt,
                      const double* acc, const double* gyr, double* out11, double* jac, double* cov, double* sqrt_info) {
    BeConfig cfg = to_cfg(c);
    IntegrationBase ib(V3(acc[0], acc[1], acc[2]), V3(gyr[0], gyr[1], gyr[2]), V3(ba[0], ba[1], ba[2]), V3(bg[0], bg[1], bg[2]), cfg);
    for (int i = 1; i < n; i++) ib.push_back(dt[i], V3(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]), V3(gyr[3 * i], gyr[3 * i + 1], gyr[3 * i + 2]));
    out11[0] = ib.sum_dt;
    out11[1] = ib.delta_p.x; out11[2] = ib.delta_p.y; out11[3] = ib.delta_p.z;
    out11[4] = ib.delta_q.w; out11[5] = ib.delta_q.x; out11[6] = ib.delta_q.y; out11[7] = ib.delta_q.z;
    out11[8] = ib.delta_v.x; out11[9] = ib.delta_v.y; out11[10] = ib.delta_v.z;
    std::memcpy(jac, ib.jacobian.d.data(), 225 * sizeof(double));
    std::memcpy(cov, ib.covariance.d.data(), 225 * sizeof(double));
    if (sqrt_info) std::memcpy(sqrt_info, ib.sqrt_info().d.data(), 225 * sizeof(double));
}
// IMUFactor::Evaluate on the pre-integration above: params = pose_i(7) sb_i(9) pose_j(7) sb_j(9) (Ceres layouts);
// residual 15 (whitened), raw residual 15, Jacobians 15x7, 15x9, 15x7, 15x9 row-major
void orc_imu_factor(const orc_be_config* c, const double* ba, const double* bg, int n, const double* dt, const double* acc,
                    const double* gyr, const double* params32, double* res, double* raw, double* J0, double* J1, double* J2, double* J3) {
    BeConfig cfg = to_cfg(c);
    IntegrationBase ib(V3(acc[0], acc[1], acc[2]), V3(gyr[0], gyr[1], gyr[2]), V3(ba[0], ba[1], ba[2]), V3(bg[0], bg[1], bg[2]), cfg);
    for (int i = 1; i < n; i++) ib.push_back(dt[i], V3(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]), V3(gyr[3 * i], gyr[3 * i + 1], gyr[3 * i + 2]));
    IMUFactor f(&ib);
    const double* p[4] = {params32, params32 + 7, params32 + 16, params32 + 23};
    double* J[4] = {J0, J1, J2, J3};
    f.Evaluate(p, res, J);
    if (raw) {
        auto Q = [](const double* x) { return Quat(x[6], x[3], x[4], x[5]); };
        ib.evaluate(V3(p[0][0], p[0][1], p[0][2]), Q(p[0]), V3(p[1][0], p[1][1], p[1][2]), V3(p[1][3], p[1][4], p[1][5]), V3(p[1][6], p[1][7], p[1][8]),
                    V3(p[2][0], p[2][1], p[2][2]), Q(p[2]), V3(p[3][0], p[3][1], p[3][2]), V3(p[3][3], p[3][4], p[3][5]), V3(p[3][6], p[3][7], p[3][8]), raw);
    }
}
// Projection(Td)Factor::Evaluate.  params: pose_i(7) pose_j(7) ex(7) inv_depth(1) td(1); data: pts_i(3) pts_j(3)
// [vel_i(2) vel_j(2) td_i td_j row_i row_j]; J: 2x7,2x7,2x7,2x1,2x1
void orc_projection_factor(int use_td, double focal_length, double TR, double ROW, const double* params23, const double* data,
                           double* res, double* Ji, double* Jj, double* Jex, double* Jl, double* Jtd) {
    const double* p[5] = {params23, params23 + 7, params23 + 14, params23 + 21, params23 + 22};
    double* J[5] = {Ji, Jj, Jex, Jl, Jtd};
    const V3 pi(data[0], data[1], data[2]), pj(data[3], data[4], data[5]);
    if (use_td) {
        ProjectionTdFactor f(pi, pj, data[6], data[7], data[8], data[9], data[10], data[11], data[12], data[13], focal_length / 1.5, TR, ROW);
        f.Evaluate(p, res, J);
    } else {
        ProjectionFactor f(pi, pj, focal_length / 1.5);
        f.Evaluate(p, res, J);
    }
}
void orc_pose_plus(const double* x, const double* delta, double* out) { pose_plus(x, delta, out); }
// solver cross-check hooks (be_solver.h: ProbeEvaluate)
void orc_est_set_probe(void* h, void (*cb)(int, int), int solve_index) {
    Estimator* e = static_cast<Estimator*>(h);
    e->probe_cb = cb;
    e->probe_solve = solve_index;
}
int orc_est_probe_residuals(void* h, const double* delta, double* out) {
    Estimator* e = static_cast<Estimator*>(h);
    if (!e->probe_problem) return -1;
    ProbeEvaluate(*e->probe_problem, delta, out);
    return 0;
}
void orc_est_profile(void* h, double* out2) {
    Estimator* e = static_cast<Estimator*>(h);
    out2[0] = e->prof_solve_s;
    out2[1] = e->prof_marg_s;
}
void orc_est_set_fast_eigen(void* h, int on) { static_cast<Estimator*>(h)->fast_eigen = on != 0; }
void orc_est_set_iterations(void* h, int n) { static_cast<Estimator*>(h)->cfg.num_iterations = n; }
// sym_eigen / cholesky known answers
void orc_sym_eigen_ql(int n, const double* A, double* w, double* V) {
    Mat a(n, n), v;
    std::memcpy(a.d.data(), A, (size_t)n * n * sizeof(double));
    std::vector<double> ww;
    sym_eigen_ql(a, ww, v);
    std::memcpy(w, ww.data(), n * sizeof(double));
    std::memcpy(V, v.d.data(), (size_t)n * n * sizeof(double));
}
void orc_sym_eigen(int n, const double* A, double* w, double* V) {
    Mat a(n, n), v;
    std::memcpy(a.d.data(), A, (size_t)n * n * sizeof(double));
    std::vector<double> ww;
    sym_eigen(a, ww, v);
    std::memcpy(w, ww.data(), n * sizeof(double));
    std::memcpy(V, v.d.data(), (size_t)n * n * sizeof(double));
}

}  // extern "C"

// kept blocks of the last prior, (type, index after addr_shift, offset, local size) each; returns the block count
extern "C" int orc_est_prior_blocks(void* h, int* out4) {
    Estimator* e = (Estimator*)h;
    if (!e->last_marginalization_info) return 0;
    const MarginalizationInfo* mi = e->last_marginalization_info.get();
    const int nb = (int)mi->keep_block_size.size();
    for (int k = 0; k < nb; k++) {
        double* addr = e->last_marginalization_parameter_blocks[k];
        int type = -1, index = 0;
        for (int i = 0; i <= e->W; i++) {
            if (addr == e->para_Pose[i].data()) type = 0, index = i;
            if (addr == e->para_SpeedBias[i].data()) type = 1, index = i;
        }
        if (addr == e->para_Ex_Pose.data()) type = 2;
        if (addr == e->para_Td.data()) type = 3;
        out4[4 * k] = type;
        out4[4 * k + 1] = index;
        out4[4 * k + 2] = mi->keep_block_idx[k] - mi->m;
        out4[4 * k + 3] = MarginalizationInfo::localSize(mi->keep_block_size[k]);
    }
    return nb;
}

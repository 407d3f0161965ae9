// Compile/link check of the host shims against the C ABI (tests/test_abi.py builds this; it is never run on CPU boxes).
#include <cstdio>
#include <vector>

#include "estimator.h"
#include "feature_tracker.h"

namespace vinsb200 { bool PUB_THIS_FRAME = false; }

struct FakeMat {  // what cv::Mat looks like to readImage
    const unsigned char* data;
    int rows, cols;
    size_t step;
};

int main() {
    vt_config tc{};
    tc.rows = 480; tc.cols = 752; tc.max_cnt = 150; tc.min_dist = 30; tc.freq = 10; tc.equalize = 1; tc.focal_length = 460;
    tc.f_threshold = 1.0;
    const double K[8] = {461.6, 460.3, 363.0, 248.1, -0.2917, 0.08228, 5.333e-05, -1.578e-04};
    for (int i = 0; i < 8; i++) tc.intrinsics[i] = K[i];
    try {
        vinsb200::FeatureTracker trk;
        trk.configure(tc);
        std::vector<unsigned char> img(480 * 752, 128);
        FakeMat m{img.data(), 480, 752, 752};
        trk.readImage(m, 0.0);
        std::printf("tracked %zu\n", trk.ids.size());
        ve_config ec{};
        ec.window_size = 10; ec.max_features = 1000; ec.num_iterations = 8; ec.focal_length = 460; ec.keyframe_parallax = 10;
        ec.acc_n = 0.08; ec.gyr_n = 0.004; ec.acc_w = 4e-5; ec.gyr_w = 2e-6; ec.g_norm = 9.81007; ec.init_depth = 5; ec.row = 480;
        ec.ric[0] = ec.ric[4] = ec.ric[8] = 1;
        vinsb200::Estimator estimator(ec);
        // the expressions estimator_node.cpp and utility/visualization.cpp apply to the Estimator (update() :84-90, process() :242-316,
        // restart_callback :190-193, pubOdometry): written out here so that a signature drift breaks the build
        using vinsb200::Vector3d;
        using vinsb200::Matrix3d;
        const int WINDOW_SIZE = ec.window_size;
        estimator.processIMU(0.005, Vector3d(0, 0, 9.8), Vector3d(0, 0, 0));
        vinsb200::Estimator::ImageMap image;
        vinsb200::Vector7d xyz_uv_velocity;
        xyz_uv_velocity[0] = 0.1; xyz_uv_velocity[1] = -0.2; xyz_uv_velocity[2] = 1.0;
        image[7].emplace_back(0, xyz_uv_velocity);
        struct Stamp { double t; double toSec() const { return t; } };
        struct Header { Stamp stamp; } header{{0.05}};
        estimator.processImage(image, header);
        Vector3d tmp_P = estimator.Ps[WINDOW_SIZE], tmp_V = estimator.Vs[WINDOW_SIZE], tmp_Ba = estimator.Bas[WINDOW_SIZE], tmp_Bg = estimator.Bgs[WINDOW_SIZE];
        Matrix3d tmp_R = estimator.Rs[WINDOW_SIZE];
        Vector3d acc_0 = estimator.acc_0, gyr_0 = estimator.gyr_0, g = estimator.g;
        Vector3d cam = estimator.ric[0] * Vector3d(0, 0, 1);
        const double camx = cam[0] + estimator.tic[0][0];
        std::vector<Vector3d> match_points;
        estimator.setReloFrame(0.05, 3, match_points, Vector3d(0, 0, 0), tmp_R);
        // what pubRelocalization / pubOdometry read after a loop message (utility/visualization.cpp:130-131, 326-350)
        Vector3d correct_t = estimator.drift_correct_r * estimator.Ps[WINDOW_SIZE] + estimator.drift_correct_t;
        const double relo_sum = correct_t[0] + estimator.relo_relative_t[1] + estimator.relo_relative_q[0] + estimator.relo_relative_yaw +
                                estimator.relo_frame_index + estimator.relo_frame_stamp + (estimator.relocalization_info ? 1.0 : 0.0);
        (void)relo_sum;
        // pubKeyPoses / pubPointCloud / pubKeyframe (utility/visualization.cpp:176-206, 228-296, 352-397)
        double cloud = 0;
        for (size_t i = 0; i < estimator.key_poses.size(); i++) cloud += estimator.key_poses[i].x();
        for (auto& it_per_id : estimator.f_manager.feature) {
            int used_num = it_per_id.feature_per_frame.size();
            if (!(used_num >= 2 && it_per_id.start_frame < WINDOW_SIZE - 2)) continue;
            if (it_per_id.start_frame > WINDOW_SIZE * 3.0 / 4.0 || it_per_id.solve_flag != 1) continue;
            int imu_i = it_per_id.start_frame;
            Vector3d pts_i = it_per_id.feature_per_frame[0].point * it_per_id.estimated_depth;
            Vector3d w_pts_i = estimator.Rs[imu_i] * (estimator.ric[0] * pts_i + estimator.tic[0]) + estimator.Ps[imu_i];
            int imu_j = WINDOW_SIZE - 2 - it_per_id.start_frame;
            cloud += w_pts_i(0) + it_per_id.feature_per_frame[imu_j < used_num ? imu_j : 0].uv.x() + it_per_id.feature_id;
        }
        cloud += estimator.f_manager.getFeatureCount();
        (void)cloud;
        const bool nonlinear = estimator.solver_flag == vinsb200::Estimator::SolverFlag::NON_LINEAR;
        estimator.clearState();
        estimator.setParameter();
        std::printf("frame_count %d td %.3f nonlinear %d %.3f %.3f %.3f\n", estimator.frame_count, estimator.td, (int)nonlinear,
                    tmp_P.x() + tmp_V.y() + tmp_Ba.z() + tmp_Bg[0] + tmp_R(0, 0), acc_0[2] + gyr_0[0] + g[2], camx + estimator.Headers[0]);
    } catch (const std::exception& e) {
        std::printf("no device: %s\n", e.what());
    }
    return 0;
}

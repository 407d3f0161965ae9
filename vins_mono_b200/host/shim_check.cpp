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
        vinsb200::Estimator est(ec);
        est.processIMU(0.005, vinsb200::Vector3d{{0, 0, 9.8}}, vinsb200::Vector3d{{0, 0, 0}});
        std::printf("frame_count %d\n", est.frame_count);
    } catch (const std::exception& e) {
        std::printf("no device: %s\n", e.what());
    }
    return 0;
}

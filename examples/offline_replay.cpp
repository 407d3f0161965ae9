// Offline replay of one or more recorded sequences through the C ABI, without ROS or Python: the reference's two node
// loops (include/vinsb200/replay.h) around a CUDA tracker and estimator per sequence.
//
//   offline_replay <dir> [copies]
//
// <dir> holds a sequence exported by harness/export_sequence.py (or by any tool that writes the same plain files):
//   meta.txt     rows cols n_images                          (one line)
//   frames.u8    n_images * rows * cols bytes, row-major grayscale
//   stamps.txt   n_images image time stamps [s], one per line
//   imu.txt      t ax ay az gx gy gz per line (accelerometer m/s^2, gyroscope rad/s)
//   seed.txt     OPTIONAL: t px py pz qw qx qy qz vx vy vz per line + last line "bias bax bay baz bgx bgy bgz", an external
//                initial window handed to ve_set_seed; without the file the estimator runs its own initialStructure()
//                (relative pose, global SfM, visual-inertial alignment) like the reference
// [copies] > 1 runs that many replicas concurrently on the GPU (throughput experiments).
// Camera / noise parameters are the EuRoC ones of config/euroc/euroc_config.yaml.
//
// Build:  g++ -O2 -std=c++17 -I include examples/offline_replay.cpp -L vins_mono_b200/lib -lvinsb200 \
//             -Wl,-rpath,$PWD/vins_mono_b200/lib -lpthread -o offline_replay
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "vinsb200/replay.h"

static bool read_numbers(const std::string& path, std::vector<double>& out) {
    std::ifstream f(path);
    if (!f) return false;
    double v;
    std::string tok;
    while (f >> tok) {
        if (tok == "bias") continue;
        out.push_back(std::atof(tok.c_str()));
    }
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <sequence dir> [copies]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    const int copies = argc > 2 ? std::max(1, std::atoi(argv[2])) : 1;
    std::vector<double> meta, stamps, imu, seed;
    if (!read_numbers(dir + "/meta.txt", meta) || meta.size() < 3 || !read_numbers(dir + "/stamps.txt", stamps) ||
        !read_numbers(dir + "/imu.txt", imu)) {
        std::fprintf(stderr, "cannot read the sequence files in %s\n", dir.c_str());
        return 2;
    }
    const int rows = (int)meta[0], cols = (int)meta[1], n_img = (int)meta[2];
    std::vector<uint8_t> frames((size_t)rows * cols * n_img);
    {
        std::ifstream f(dir + "/frames.u8", std::ios::binary);
        if (!f.read(reinterpret_cast<char*>(frames.data()), (std::streamsize)frames.size())) {
            std::fprintf(stderr, "frames.u8 is shorter than meta.txt says\n");
            return 2;
        }
    }
    const int n_imu = (int)imu.size() / 7;
    std::vector<double> imu_t(n_imu), acc(3 * (size_t)n_imu), gyr(3 * (size_t)n_imu);
    for (int k = 0; k < n_imu; k++) {
        imu_t[k] = imu[7 * k];
        for (int c = 0; c < 3; c++) {
            acc[3 * k + c] = imu[7 * k + 1 + c];
            gyr[3 * k + c] = imu[7 * k + 4 + c];
        }
    }
    const bool have_seed = read_numbers(dir + "/seed.txt", seed) && seed.size() >= 17;
    const int n_seed = have_seed ? ((int)seed.size() - 6) / 11 : 0;
    const double* bias = have_seed ? seed.data() + 11 * (size_t)n_seed : nullptr;

    vt_config tc{};
    tc.rows = rows; tc.cols = cols; tc.max_cnt = 150; tc.min_dist = 30; tc.freq = 10; tc.equalize = 1;
    tc.focal_length = 460; tc.f_threshold = 1.0; tc.camera_model = VT_CAMERA_PINHOLE;
    const double K[8] = {461.6, 460.3, 363.0, 248.1, -0.2917, 0.08228, 5.333e-05, -1.578e-04};
    for (int i = 0; i < 8; i++) tc.intrinsics[i] = K[i];
    ve_config ec{};
    ec.window_size = 10; ec.max_features = 1000; ec.num_iterations = 8; ec.focal_length = 460; ec.keyframe_parallax = 10;
    ec.acc_n = 0.08; ec.gyr_n = 0.004; ec.acc_w = 4e-5; ec.gyr_w = 2e-6; ec.g_norm = 9.81007; ec.init_depth = 5; ec.row = rows;
    const double ric[9] = {0.0148655429818, -0.999880929698, 0.00414029679422, 0.999557249008, 0.0149672133247, 0.025715529948,
                           -0.0257744366974, 0.00375618835797, 0.999660727178};
    const double tic[3] = {-0.0216401454975, -0.064676986768, 0.00981073058949};
    for (int i = 0; i < 9; i++) ec.ric[i] = ric[i];
    for (int i = 0; i < 3; i++) ec.tic[i] = tic[i];

    std::vector<vt_tracker*> trk(copies, nullptr);
    std::vector<ve_estimator*> est(copies, nullptr);
    std::vector<vr_sequence> seqs(copies);
    for (int k = 0; k < copies; k++) {
        if (vt_create(&tc, &trk[k]) != VT_OK || ve_create(&ec, &est[k]) != VE_OK) {
            std::fprintf(stderr, "no usable CUDA device (this library has no CPU path)\n");
            return 1;
        }
        if (have_seed) ve_set_seed(est[k], n_seed, seed.data(), bias, bias + 3);
        vr_sequence& s = seqs[k];
        s.images = frames.data();
        s.row_stride = (size_t)cols;
        s.frame_stride = (size_t)rows * cols;
        s.n_images = n_img;
        s.images_on_device = 0;
        s.stamps = stamps.data();
        s.n_imu = n_imu;
        s.imu_t = imu_t.data();
        s.acc = acc.data();
        s.gyr = gyr.data();
    }
    vr_session* ses = nullptr;
    if (vr_open(copies, trk.data(), est.data(), seqs.data(), &ses) != 0) return 1;
    const auto t0 = std::chrono::steady_clock::now();
    const int frames_done = vr_advance(ses, n_img);  // until the images run out
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (frames_done < 0) {
        std::fprintf(stderr, "replay failed: %s\n", vr_last_error(ses));
        return 1;
    }
    std::printf("# %d published frames over %d sequence(s) in %.3f s (%.1f frames/s, host images)\n", frames_done, copies, sec,
                frames_done / sec);
    {
        double ii[8];
        const int self = ve_init_info(est[0], ii);
        std::printf("# initial window: %s (reference frame l = %d, scale %.4f, failed attempts %d)\n",
                    self == 1 ? "own initialStructure" : "seed.txt", (int)ii[0], ii[1], (int)ii[7]);
    }
    const int n = vr_trajectory(ses, 0, 0, nullptr, nullptr);
    std::vector<double> tt(n), pp(3 * (size_t)n);
    vr_trajectory(ses, 0, n, tt.data(), pp.data());
    for (int k = 0; k < n; k++) std::printf("%.6f %.9f %.9f %.9f\n", tt[k], pp[3 * k], pp[3 * k + 1], pp[3 * k + 2]);
    vr_close(ses);
    for (int k = 0; k < copies; k++) {
        vt_destroy(trk[k]);
        ve_destroy(est[k]);
    }
    return 0;
}

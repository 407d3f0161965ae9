// Host shim with the reference's class surface: drop-in for vins_estimator/src/estimator.h:25-139 as far as
// estimator_node.cpp and utility/visualization.cpp use it (processIMU, processImage, clearState, setParameter and
// the public window arrays).  Vector/matrix types are plain structs so that Eigen is not required; with Eigen
// available, Eigen::Map<Eigen::Vector3d>(Ps[i].v) etc. view them without copies.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vinsb200/estimator.h"

namespace vinsb200 {

struct Vector3d { double v[3]; };
struct Quaterniond { double w, x, y, z; };
using FeatureObservation = std::pair<int, std::vector<double>>;  // camera id, (x, y, z, u, v, vx, vy)

class Estimator {
  public:
    enum SolverFlag { INITIAL, NON_LINEAR };
    enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };

    explicit Estimator(const ve_config& cfg) : cfg_(cfg) {
        if (ve_create(&cfg, &h_) != VE_OK) throw std::runtime_error("ve_create failed (no CUDA device?)");
        const int n = cfg.window_size + 1;
        Ps.resize(n); Vs.resize(n); Bas.resize(n); Bgs.resize(n); Rs.resize(n);
    }
    ~Estimator() { ve_destroy(h_); }
    Estimator(const Estimator&) = delete;
    Estimator& operator=(const Estimator&) = delete;

    void setParameter() {}                       // parameters travel in ve_config
    void clearState() { check(ve_clear_state(h_)); refresh(); }
    void processIMU(double dt, const Vector3d& linear_acceleration, const Vector3d& angular_velocity) {
        check(ve_process_imu(h_, dt, linear_acceleration.v, angular_velocity.v));
    }
    // image: feature_id -> [(camera_id, xyz_uv_velocity)] exactly as estimator_node.cpp:296-313 builds it
    void processImage(const std::map<int, std::vector<FeatureObservation>>& image, double header_stamp) {
        std::vector<int> ids;
        std::vector<double> d;
        for (const auto& kv : image) {
            ids.push_back(kv.first);
            d.insert(d.end(), kv.second[0].second.begin(), kv.second[0].second.begin() + 7);
        }
        check(ve_process_image(h_, (int)ids.size(), ids.data(), d.data(), header_stamp));
        refresh();
    }

    SolverFlag solver_flag = INITIAL;
    MarginalizationFlag marginalization_flag = MARGIN_OLD;
    int frame_count = 0;
    std::vector<Vector3d> Ps, Vs, Bas, Bgs;
    std::vector<Quaterniond> Rs;  // rotations as quaternions (w, x, y, z)
    double td = 0;

  private:
    void check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("vinsb200: ") + ve_last_error(h_));
    }
    void refresh() {
        std::vector<double> s(16 * Ps.size());
        ve_get_states(h_, s.data(), &td);
        for (size_t i = 0; i < Ps.size(); i++) {
            const double* o = &s[16 * i];
            for (int k = 0; k < 3; k++) { Ps[i].v[k] = o[k]; Vs[i].v[k] = o[7 + k]; Bas[i].v[k] = o[10 + k]; Bgs[i].v[k] = o[13 + k]; }
            Rs[i] = Quaterniond{o[3], o[4], o[5], o[6]};
        }
        int info[10];
        double c[2];
        ve_info(h_, info, c);
        solver_flag = info[0] ? NON_LINEAR : INITIAL;
        frame_count = info[1];
        marginalization_flag = info[2] ? MARGIN_SECOND_NEW : MARGIN_OLD;
    }
    ve_config cfg_;
    ve_estimator* h_ = nullptr;
};

}  // namespace vinsb200

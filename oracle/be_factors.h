// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// float64 CPU restatement of the vins_estimator factor "operators", following the reference files:
//   IntegrationBase            vins_estimator/src/factor/integration_base.h:13-186
//   IMUFactor                  vins_estimator/src/factor/imu_factor.h:19-179
//   ProjectionFactor           vins_estimator/src/factor/projection_factor.cpp:21-121
//   ProjectionTdFactor         vins_estimator/src/factor/projection_td_factor.cpp:6-141
//   PoseLocalParameterization  vins_estimator/src/factor/pose_local_parameterization.cpp:3-27
//   ResidualBlockInfo / MarginalizationInfo / MarginalizationFactor
//                              vins_estimator/src/factor/marginalization_factor.cpp:3-69, :89-129, :174-319, :333-381
//   ceres::CauchyLoss(1.0)     vins_estimator/src/estimator.cpp:675 (Ceres source not in /root/reference)
// Interfaces mirror ceres::CostFunction::Evaluate(parameters, residuals, jacobians) (row-major Jacobians in
// the global 7-wide pose parameterisation, last column zero).  Parity: unpinned by the reference (it
// ships no tests); pinned here by SURVEY.md Appendix E known answers and finite differences
// (tests/test_oracle_backend.py).
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "be_math.h"

namespace orc {

enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };

struct BeConfig {
    int window_size = 10;
    int num_iterations = 8;
    int estimate_extrinsic = 0;
    int estimate_td = 0;
    double focal_length = 460.0;
    double min_parallax = 10.0 / 460.0;  // keyframe_parallax / FOCAL_LENGTH (parameters.cpp:79)
    double acc_n = 0.08, gyr_n = 0.004, acc_w = 0.00004, gyr_w = 2.0e-6;
    V3 G{0, 0, 9.81007};
    double init_depth = 5.0;  // INIT_DEPTH (parameters.cpp:114)
    double td = 0.0, tr = 0.0, row = 480.0;
    V3 tic;
    M3 ric = M3::Identity();
};

struct CostFunction {
    int num_residuals = 0;
    std::vector<int> block_sizes;
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    virtual ~CostFunction() {}
};

struct CauchyLoss {
    double b, c;
    explicit CauchyLoss(double a) : b(a * a), c(1 / (a * a)) {}
    void Evaluate(double s, double rho[3]) const;
};

class IntegrationBase {
  public:
    IntegrationBase(const V3& acc_0, const V3& gyr_0, const V3& ba, const V3& bg, const BeConfig& cfg);
    void push_back(double dt, const V3& acc, const V3& gyr);
    void repropagate(const V3& ba, const V3& bg);
    void propagate(double dt, const V3& acc_1, const V3& gyr_1);
    void evaluate(const V3& Pi, const Quat& Qi, const V3& Vi, const V3& Bai, const V3& Bgi, const V3& Pj, const Quat& Qj,
                  const V3& Vj, const V3& Baj, const V3& Bgj, double residuals[15]) const;
    const Mat& sqrt_info() const;  // LLT(covariance^-1).matrixL().transpose() (imu_factor.h:64)

    double dt = 0;
    V3 acc_0, gyr_0, acc_1, gyr_1;
    V3 linearized_acc, linearized_gyr, linearized_ba, linearized_bg;
    Mat jacobian, covariance, noise;
    double sum_dt = 0;
    V3 delta_p, delta_v;
    Quat delta_q;
    std::vector<double> dt_buf;
    std::vector<V3> acc_buf, gyr_buf;
    V3 G;

  private:
    mutable Mat sqrt_info_;
    mutable bool sqrt_info_valid_ = false;
};

struct IMUFactor : CostFunction {
    explicit IMUFactor(const IntegrationBase* p) : pre_integration(p) {
        num_residuals = 15;
        block_sizes = {7, 9, 7, 9};
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
    const IntegrationBase* pre_integration;
};

struct ProjectionFactor : CostFunction {
    ProjectionFactor(const V3& pi, const V3& pj, double sqrt_info_scale) : pts_i(pi), pts_j(pj), sqrt_info(sqrt_info_scale) {
        num_residuals = 2;
        block_sizes = {7, 7, 7, 1};
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
    V3 pts_i, pts_j;
    double sqrt_info;  // FOCAL_LENGTH / 1.5 on the diagonal (estimator.cpp:17)
};

struct ProjectionTdFactor : CostFunction {
    ProjectionTdFactor(const V3& pi, const V3& pj, double vix, double viy, double vjx, double vjy, double td_i_,
                       double td_j_, double row_i_, double row_j_, double sqrt_info_scale, double TR_, double ROW_)
        : pts_i(pi), pts_j(pj), velocity_i(vix, viy, 0), velocity_j(vjx, vjy, 0), td_i(td_i_), td_j(td_j_),
          row_i(row_i_ - ROW_ / 2), row_j(row_j_ - ROW_ / 2), sqrt_info(sqrt_info_scale), TR(TR_), ROW(ROW_) {
        num_residuals = 2;
        block_sizes = {7, 7, 7, 1, 1};
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
    V3 pts_i, pts_j, velocity_i, velocity_j;
    double td_i, td_j, row_i, row_j, sqrt_info, TR, ROW;
};

// PoseLocalParameterization::Plus
void pose_plus(const double* x, const double* delta, double* x_plus_delta);

struct ResidualBlockInfo {
    ResidualBlockInfo(std::shared_ptr<CostFunction> cf, const CauchyLoss* loss, std::vector<double*> blocks,
                      std::vector<int> drop)
        : cost_function(cf), loss_function(loss), parameter_blocks(blocks), drop_set(drop) {}
    void Evaluate();
    std::shared_ptr<CostFunction> cost_function;
    const CauchyLoss* loss_function;
    std::vector<double*> parameter_blocks;
    std::vector<int> drop_set;
    std::vector<Mat> jacobians;  // row-major num_residuals x block_size
    std::vector<double> residuals;
};

class MarginalizationInfo {
  public:
    void addResidualBlockInfo(std::shared_ptr<ResidualBlockInfo> info);
    void preMarginalize();
    void marginalize();
    std::vector<double*> getParameterBlocks(std::map<double*, double*>& addr_shift);
    static int localSize(int size) { return size == 7 ? 6 : size; }

    std::vector<std::shared_ptr<ResidualBlockInfo>> factors;
    int m = 0, n = 0;
    // The reference keys these maps by pointer value in std::unordered_map (hash order).  Here blocks are
    // ordered by first appearance, which only permutes rows/columns of A.
    std::vector<double*> block_order;
    std::map<double*, int> parameter_block_size;  // global size
    std::map<double*, int> parameter_block_idx;   // local offset
    std::map<double*, bool> dropped;
    std::map<double*, std::vector<double>> parameter_block_data;
    std::vector<int> keep_block_size, keep_block_idx;
    std::vector<std::vector<double>> keep_block_data;
    Mat linearized_jacobians;
    std::vector<double> linearized_residuals;
    Mat A_debug;                  // Schur-complemented information matrix (for parity tests)
    std::vector<double> b_debug;
    const double eps = 1e-8;
    // false: cyclic Jacobi (parity tests: independent of the product, relative accuracy); true: tridiagonal QL, the
    // algorithm class of the Eigen solver the reference calls and ~10x faster on the 200x200 A_mm (timed CPU baseline)
    bool eigen_ql = false;
};

struct MarginalizationFactor : CostFunction {
    explicit MarginalizationFactor(const MarginalizationInfo* info);
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;
    const MarginalizationInfo* marginalization_info;
};

}  // namespace orc

// ORACLE — TEST INFRASTRUCTURE ONLY.  See be_factors.h for the reference files each part follows.
#include "be_factors.h"

#include <cfloat>
#include <limits>

namespace orc {

void CauchyLoss::Evaluate(double s, double rho[3]) const {
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
}

// ------------------------------------------------------------------------------------------------
IntegrationBase::IntegrationBase(const V3& a0, const V3& g0, const V3& ba, const V3& bg, const BeConfig& cfg)
    : acc_0(a0), gyr_0(g0), linearized_acc(a0), linearized_gyr(g0), linearized_ba(ba), linearized_bg(bg),
      jacobian(Mat::Identity(15)), covariance(15, 15), noise(18, 18), G(cfg.G) {
    for (int i = 0; i < 3; i++) {
        noise(i, i) = cfg.acc_n * cfg.acc_n;
        noise(3 + i, 3 + i) = cfg.gyr_n * cfg.gyr_n;
        noise(6 + i, 6 + i) = cfg.acc_n * cfg.acc_n;
        noise(9 + i, 9 + i) = cfg.gyr_n * cfg.gyr_n;
        noise(12 + i, 12 + i) = cfg.acc_w * cfg.acc_w;
        noise(15 + i, 15 + i) = cfg.gyr_w * cfg.gyr_w;
    }
}

void IntegrationBase::push_back(double dt_, const V3& acc, const V3& gyr) {
    dt_buf.push_back(dt_);
    acc_buf.push_back(acc);
    gyr_buf.push_back(gyr);
    propagate(dt_, acc, gyr);
}

void IntegrationBase::repropagate(const V3& ba, const V3& bg) {
    sum_dt = 0.0;
    acc_0 = linearized_acc;
    gyr_0 = linearized_gyr;
    delta_p = V3();
    delta_q = Quat();
    delta_v = V3();
    linearized_ba = ba;
    linearized_bg = bg;
    jacobian = Mat::Identity(15);
    covariance = Mat(15, 15);
    for (size_t i = 0; i < dt_buf.size(); i++) propagate(dt_buf[i], acc_buf[i], gyr_buf[i]);
}

// midPointIntegration + bookkeeping of propagate() (integration_base.h:54-158)
void IntegrationBase::propagate(double _dt, const V3& _acc_1, const V3& _gyr_1) {
    dt = _dt;
    acc_1 = _acc_1;
    gyr_1 = _gyr_1;
    const V3 un_acc_0 = delta_q * (acc_0 - linearized_ba);
    const V3 un_gyr = 0.5 * (gyr_0 + gyr_1) - linearized_bg;
    Quat result_delta_q = delta_q * Quat(1, un_gyr.x * _dt / 2, un_gyr.y * _dt / 2, un_gyr.z * _dt / 2);
    const V3 un_acc_1 = result_delta_q * (acc_1 - linearized_ba);
    const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    const V3 result_delta_p = delta_p + delta_v * _dt + 0.5 * un_acc * _dt * _dt;
    const V3 result_delta_v = delta_v + un_acc * _dt;
    {
        const V3 w_x = 0.5 * (gyr_0 + gyr_1) - linearized_bg;
        const V3 a_0_x = acc_0 - linearized_ba;
        const V3 a_1_x = acc_1 - linearized_ba;
        const M3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
        const M3 Rq = delta_q.R(), Rr = result_delta_q.R(), I = M3::Identity();
        Mat F(15, 15);
        F.set_block3(0, 0, I);
        F.set_block3(0, 3, -0.25 * Rq * R_a_0_x * _dt * _dt + -0.25 * Rr * R_a_1_x * (I - R_w_x * _dt) * _dt * _dt);
        F.set_block3(0, 6, I * _dt);
        F.set_block3(0, 9, -0.25 * (Rq + Rr) * _dt * _dt);
        F.set_block3(0, 12, -0.25 * Rr * R_a_1_x * _dt * _dt * -_dt);
        F.set_block3(3, 3, I - R_w_x * _dt);
        F.set_block3(3, 12, -1.0 * I * _dt);
        F.set_block3(6, 3, -0.5 * Rq * R_a_0_x * _dt + -0.5 * Rr * R_a_1_x * (I - R_w_x * _dt) * _dt);
        F.set_block3(6, 6, I);
        F.set_block3(6, 9, -0.5 * (Rq + Rr) * _dt);
        F.set_block3(6, 12, -0.5 * Rr * R_a_1_x * _dt * -_dt);
        F.set_block3(9, 9, I);
        F.set_block3(12, 12, I);
        Mat V(15, 18);
        V.set_block3(0, 0, 0.25 * Rq * _dt * _dt);
        V.set_block3(0, 3, 0.25 * -Rr * R_a_1_x * _dt * _dt * 0.5 * _dt);
        V.set_block3(0, 6, 0.25 * Rr * _dt * _dt);
        V.set_block3(0, 9, V.block3(0, 3));
        V.set_block3(3, 3, 0.5 * I * _dt);
        V.set_block3(3, 9, 0.5 * I * _dt);
        V.set_block3(6, 0, 0.5 * Rq * _dt);
        V.set_block3(6, 3, 0.5 * -Rr * R_a_1_x * _dt * 0.5 * _dt);
        V.set_block3(6, 6, 0.5 * Rr * _dt);
        V.set_block3(6, 9, V.block3(6, 3));
        V.set_block3(9, 12, I * _dt);
        V.set_block3(12, 15, I * _dt);
        jacobian = F * jacobian;
        covariance = F * covariance * F.T() + V * noise * V.T();
    }
    delta_p = result_delta_p;
    delta_q = result_delta_q.normalized();
    delta_v = result_delta_v;
    sum_dt += dt;
    acc_0 = acc_1;
    gyr_0 = gyr_1;
    sqrt_info_valid_ = false;
}

void IntegrationBase::evaluate(const V3& Pi, const Quat& Qi, const V3& Vi, const V3& Bai, const V3& Bgi, const V3& Pj,
                               const Quat& Qj, const V3& Vj, const V3& Baj, const V3& Bgj, double r[15]) const {
    const M3 dp_dba = jacobian.block3(O_P, O_BA), dp_dbg = jacobian.block3(O_P, O_BG);
    const M3 dq_dbg = jacobian.block3(O_R, O_BG);
    const M3 dv_dba = jacobian.block3(O_V, O_BA), dv_dbg = jacobian.block3(O_V, O_BG);
    const V3 dba = Bai - linearized_ba, dbg = Bgi - linearized_bg;
    const Quat corrected_delta_q = delta_q * deltaQ(dq_dbg * dbg);
    const V3 corrected_delta_v = delta_v + dv_dba * dba + dv_dbg * dbg;
    const V3 corrected_delta_p = delta_p + dp_dba * dba + dp_dbg * dbg;
    const V3 rp = Qi.inverse() * (0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
    const V3 rq = 2 * (corrected_delta_q.inverse() * (Qi.inverse() * Qj)).vec();
    const V3 rv = Qi.inverse() * (G * sum_dt + Vj - Vi) - corrected_delta_v;
    const V3 rba = Baj - Bai, rbg = Bgj - Bgi;
    for (int i = 0; i < 3; i++) {
        r[O_P + i] = rp[i];
        r[O_R + i] = rq[i];
        r[O_V + i] = rv[i];
        r[O_BA + i] = rba[i];
        r[O_BG + i] = rbg[i];
    }
}

const Mat& IntegrationBase::sqrt_info() const {
    if (!sqrt_info_valid_) {
        Mat L;
        cholesky(inverse_lu(covariance), L);
        sqrt_info_ = L.T();
        sqrt_info_valid_ = true;
    }
    return sqrt_info_;
}

// ------------------------------------------------------------------------------------------------
namespace {

Mat Qleft(const Quat& q) {  // utility.h:51-58
    Mat a(4, 4);
    a(0, 0) = q.w;
    a(0, 1) = -q.x; a(0, 2) = -q.y; a(0, 3) = -q.z;
    a(1, 0) = q.x;  a(2, 0) = q.y;  a(3, 0) = q.z;
    const M3 s = skew(q.vec());
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a(1 + i, 1 + j) = (i == j ? q.w : 0.0) + s(i, j);
    return a;
}
Mat Qright(const Quat& p) {  // utility.h:61-68
    Mat a(4, 4);
    a(0, 0) = p.w;
    a(0, 1) = -p.x; a(0, 2) = -p.y; a(0, 3) = -p.z;
    a(1, 0) = p.x;  a(2, 0) = p.y;  a(3, 0) = p.z;
    const M3 s = skew(p.vec());
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a(1 + i, 1 + j) = (i == j ? p.w : 0.0) - s(i, j);
    return a;
}
M3 br3(const Mat& a) { return a.block3(1, 1); }

void store_rowmajor(const Mat& J, double* out) { std::memcpy(out, J.d.data(), J.d.size() * sizeof(double)); }

Quat quat_from_param(const double* p) { return Quat(p[6], p[3], p[4], p[5]); }

}  // namespace

bool IMUFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    const IntegrationBase* pre = pre_integration;
    const V3 Pi(parameters[0][0], parameters[0][1], parameters[0][2]);
    const Quat Qi = quat_from_param(parameters[0]);
    const V3 Vi(parameters[1][0], parameters[1][1], parameters[1][2]);
    const V3 Bai(parameters[1][3], parameters[1][4], parameters[1][5]);
    const V3 Bgi(parameters[1][6], parameters[1][7], parameters[1][8]);
    const V3 Pj(parameters[2][0], parameters[2][1], parameters[2][2]);
    const Quat Qj = quat_from_param(parameters[2]);
    const V3 Vj(parameters[3][0], parameters[3][1], parameters[3][2]);
    const V3 Baj(parameters[3][3], parameters[3][4], parameters[3][5]);
    const V3 Bgj(parameters[3][6], parameters[3][7], parameters[3][8]);

    double r[15];
    pre->evaluate(Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj, r);
    const Mat& sqrt_info = pre->sqrt_info();
    for (int i = 0; i < 15; i++) {
        double s = 0;
        for (int k = 0; k < 15; k++) s += sqrt_info(i, k) * r[k];
        residuals[i] = s;
    }
    if (jacobians) {
        const double sum_dt = pre->sum_dt;
        const V3& G = pre->G;
        const M3 dp_dba = pre->jacobian.block3(O_P, O_BA), dp_dbg = pre->jacobian.block3(O_P, O_BG);
        const M3 dq_dbg = pre->jacobian.block3(O_R, O_BG);
        const M3 dv_dba = pre->jacobian.block3(O_V, O_BA), dv_dbg = pre->jacobian.block3(O_V, O_BG);
        const M3 RiT = Qi.inverse().R();
        if (jacobians[0]) {
            Mat J(15, 7);
            J.set_block3(O_P, O_P, -RiT);
            J.set_block3(O_P, O_R, skew(Qi.inverse() * (0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
            const Quat corrected_delta_q = pre->delta_q * deltaQ(dq_dbg * (Bgi - pre->linearized_bg));
            J.set_block3(O_R, O_R, -br3(Qleft(Qj.inverse() * Qi) * Qright(corrected_delta_q)));
            J.set_block3(O_V, O_R, skew(Qi.inverse() * (G * sum_dt + Vj - Vi)));
            store_rowmajor(sqrt_info * J, jacobians[0]);
        }
        if (jacobians[1]) {
            Mat J(15, 9);
            J.set_block3(O_P, O_V - O_V, -RiT * sum_dt);
            J.set_block3(O_P, O_BA - O_V, -dp_dba);
            J.set_block3(O_P, O_BG - O_V, -dp_dbg);
            J.set_block3(O_R, O_BG - O_V, -br3(Qleft(Qj.inverse() * Qi * pre->delta_q)) * dq_dbg);
            J.set_block3(O_V, O_V - O_V, -RiT);
            J.set_block3(O_V, O_BA - O_V, -dv_dba);
            J.set_block3(O_V, O_BG - O_V, -dv_dbg);
            J.set_block3(O_BA, O_BA - O_V, -M3::Identity());
            J.set_block3(O_BG, O_BG - O_V, -M3::Identity());
            store_rowmajor(sqrt_info * J, jacobians[1]);
        }
        if (jacobians[2]) {
            Mat J(15, 7);
            J.set_block3(O_P, O_P, RiT);
            const Quat corrected_delta_q = pre->delta_q * deltaQ(dq_dbg * (Bgi - pre->linearized_bg));
            J.set_block3(O_R, O_R, br3(Qleft(corrected_delta_q.inverse() * Qi.inverse() * Qj)));
            store_rowmajor(sqrt_info * J, jacobians[2]);
        }
        if (jacobians[3]) {
            Mat J(15, 9);
            J.set_block3(O_V, O_V - O_V, RiT);
            J.set_block3(O_BA, O_BA - O_V, M3::Identity());
            J.set_block3(O_BG, O_BG - O_V, M3::Identity());
            store_rowmajor(sqrt_info * J, jacobians[3]);
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
namespace {

// Shared tail of ProjectionFactor / ProjectionTdFactor (identical algebra in both reference files).
struct ProjGeom {
    V3 pts_camera_i, pts_imu_i, pts_w, pts_imu_j, pts_camera_j;
    M3 Ri, Rj, ric;
    double dep_j;
};

ProjGeom project(const double* const* p, const V3& pts_i_eff) {
    ProjGeom g;
    const V3 Pi(p[0][0], p[0][1], p[0][2]), Pj(p[1][0], p[1][1], p[1][2]), tic(p[2][0], p[2][1], p[2][2]);
    const Quat Qi = quat_from_param(p[0]), Qj = quat_from_param(p[1]), qic = quat_from_param(p[2]);
    const double inv_dep_i = p[3][0];
    g.pts_camera_i = pts_i_eff / inv_dep_i;
    g.pts_imu_i = qic * g.pts_camera_i + tic;
    g.pts_w = Qi * g.pts_imu_i + Pi;
    g.pts_imu_j = Qj.inverse() * (g.pts_w - Pj);
    g.pts_camera_j = qic.inverse() * (g.pts_imu_j - tic);
    g.dep_j = g.pts_camera_j.z;
    g.Ri = Qi.R();
    g.Rj = Qj.R();
    g.ric = qic.R();
    return g;
}

// reduce (2x3) times a 3-vector / 3x3 matrix helpers
struct Reduce {
    double a[2][3];
    V3 row(int i) const { return V3(a[i][0], a[i][1], a[i][2]); }
};

void pose_jacobians(const ProjGeom& g, const Reduce& red, const double* const* p, const V3& pts_i_eff, double** jac) {
    const V3 Pi(p[0][0], p[0][1], p[0][2]), Pj(p[1][0], p[1][1], p[1][2]), tic(p[2][0], p[2][1], p[2][2]);
    const double inv_dep_i = p[3][0];
    auto store = [&](const M3& left, const M3& right, double* out) {  // 2x7 row-major = reduce * [left right], 0
        for (int r = 0; r < 2; r++) {
            for (int c = 0; c < 3; c++) {
                out[r * 7 + c] = red.row(r).dot(left.col(c));
                out[r * 7 + 3 + c] = red.row(r).dot(right.col(c));
            }
            out[r * 7 + 6] = 0;
        }
    };
    const M3 ricT = g.ric.T(), RjT = g.Rj.T();
    if (jac[0]) store(ricT * RjT, ricT * RjT * g.Ri * -skew(g.pts_imu_i), jac[0]);
    if (jac[1]) store(ricT * -RjT, ricT * skew(g.pts_imu_j), jac[1]);
    if (jac[2]) {
        const M3 tmp_r = ricT * RjT * g.Ri * g.ric;
        const M3 right = -tmp_r * skew(g.pts_camera_i) + skew(tmp_r * g.pts_camera_i) +
                         skew(ricT * (RjT * (g.Ri * tic + Pi - Pj) - tic));
        store(ricT * (RjT * g.Ri - M3::Identity()), right, jac[2]);
    }
    if (jac[3]) {
        const V3 v = ricT * RjT * g.Ri * g.ric * pts_i_eff * -1.0 / (inv_dep_i * inv_dep_i);
        jac[3][0] = red.row(0).dot(v);
        jac[3][1] = red.row(1).dot(v);
    }
}

}  // namespace

bool ProjectionFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    const ProjGeom g = project(parameters, pts_i);
    residuals[0] = sqrt_info * ((g.pts_camera_j / g.dep_j).x - pts_j.x);
    residuals[1] = sqrt_info * ((g.pts_camera_j / g.dep_j).y - pts_j.y);
    if (jacobians) {
        Reduce red;
        red.a[0][0] = sqrt_info * (1. / g.dep_j); red.a[0][1] = 0; red.a[0][2] = sqrt_info * (-g.pts_camera_j.x / (g.dep_j * g.dep_j));
        red.a[1][0] = 0; red.a[1][1] = sqrt_info * (1. / g.dep_j); red.a[1][2] = sqrt_info * (-g.pts_camera_j.y / (g.dep_j * g.dep_j));
        pose_jacobians(g, red, parameters, pts_i, jacobians);
    }
    return true;
}

bool ProjectionTdFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    const double td = parameters[4][0];
    const double inv_dep_i = parameters[3][0];
    const V3 pts_i_td = pts_i - (td - td_i + TR / ROW * row_i) * velocity_i;
    const V3 pts_j_td = pts_j - (td - td_j + TR / ROW * row_j) * velocity_j;
    const ProjGeom g = project(parameters, pts_i_td);
    residuals[0] = sqrt_info * ((g.pts_camera_j / g.dep_j).x - pts_j_td.x);
    residuals[1] = sqrt_info * ((g.pts_camera_j / g.dep_j).y - pts_j_td.y);
    if (jacobians) {
        Reduce red;
        red.a[0][0] = sqrt_info * (1. / g.dep_j); red.a[0][1] = 0; red.a[0][2] = sqrt_info * (-g.pts_camera_j.x / (g.dep_j * g.dep_j));
        red.a[1][0] = 0; red.a[1][1] = sqrt_info * (1. / g.dep_j); red.a[1][2] = sqrt_info * (-g.pts_camera_j.y / (g.dep_j * g.dep_j));
        pose_jacobians(g, red, parameters, pts_i_td, jacobians);
        if (jacobians[4]) {
            const V3 v = g.ric.T() * g.Rj.T() * g.Ri * g.ric * velocity_i / inv_dep_i * -1.0;
            jacobians[4][0] = red.row(0).dot(v) + sqrt_info * velocity_j.x;
            jacobians[4][1] = red.row(1).dot(v) + sqrt_info * velocity_j.y;
        }
    }
    return true;
}

void pose_plus(const double* x, const double* delta, double* out) {
    out[0] = x[0] + delta[0];
    out[1] = x[1] + delta[1];
    out[2] = x[2] + delta[2];
    const Quat q = (Quat(x[6], x[3], x[4], x[5]) * deltaQ(V3(delta[3], delta[4], delta[5]))).normalized();
    out[3] = q.x;
    out[4] = q.y;
    out[5] = q.z;
    out[6] = q.w;
}

// ------------------------------------------------------------------------------------------------
// ResidualBlockInfo::Evaluate: cost function + the robust-loss corrector (marginalization_factor.cpp:3-69)
void ResidualBlockInfo::Evaluate() {
    const int nr = cost_function->num_residuals;
    residuals.assign(nr, 0.0);
    const std::vector<int>& bs = cost_function->block_sizes;
    jacobians.clear();
    std::vector<double*> raw(bs.size());
    for (size_t i = 0; i < bs.size(); i++) jacobians.emplace_back(nr, bs[i]);
    for (size_t i = 0; i < bs.size(); i++) raw[i] = jacobians[i].d.data();
    cost_function->Evaluate(parameter_blocks.data(), residuals.data(), raw.data());
    if (loss_function) {
        double residual_scaling_, alpha_sq_norm_, rho[3];
        double sq_norm = 0;
        for (double v : residuals) sq_norm += v * v;
        loss_function->Evaluate(sq_norm, rho);
        const double sqrt_rho1_ = std::sqrt(rho[1]);
        if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
            residual_scaling_ = sqrt_rho1_;
            alpha_sq_norm_ = 0.0;
        } else {
            const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
            const double alpha = 1.0 - std::sqrt(D);
            residual_scaling_ = sqrt_rho1_ / (1 - alpha);
            alpha_sq_norm_ = alpha / sq_norm;
        }
        for (size_t i = 0; i < bs.size(); i++) {
            Mat& J = jacobians[i];
            std::vector<double> rtJ(J.c, 0.0);
            for (int k = 0; k < nr; k++)
                for (int c = 0; c < J.c; c++) rtJ[c] += residuals[k] * J(k, c);
            for (int k = 0; k < nr; k++)
                for (int c = 0; c < J.c; c++) J(k, c) = sqrt_rho1_ * (J(k, c) - alpha_sq_norm_ * residuals[k] * rtJ[c]);
        }
        for (double& v : residuals) v *= residual_scaling_;
    }
}

void MarginalizationInfo::addResidualBlockInfo(std::shared_ptr<ResidualBlockInfo> info) {
    factors.push_back(info);
    const std::vector<int>& sizes = info->cost_function->block_sizes;
    for (size_t i = 0; i < info->parameter_blocks.size(); i++) {
        double* addr = info->parameter_blocks[i];
        if (!parameter_block_size.count(addr)) block_order.push_back(addr);
        parameter_block_size[addr] = sizes[i];
    }
    for (int d : info->drop_set) dropped[info->parameter_blocks[d]] = true;
}

void MarginalizationInfo::preMarginalize() {
    for (auto& it : factors) {
        it->Evaluate();
        const std::vector<int>& sizes = it->cost_function->block_sizes;
        for (size_t i = 0; i < sizes.size(); i++) {
            double* addr = it->parameter_blocks[i];
            if (!parameter_block_data.count(addr))
                parameter_block_data[addr] = std::vector<double>(addr, addr + sizes[i]);
        }
    }
}

void MarginalizationInfo::marginalize() {
    int pos = 0;
    for (double* a : block_order)
        if (dropped.count(a)) {
            parameter_block_idx[a] = pos;
            pos += localSize(parameter_block_size[a]);
        }
    m = pos;
    for (double* a : block_order)
        if (!dropped.count(a)) {
            parameter_block_idx[a] = pos;
            pos += localSize(parameter_block_size[a]);
        }
    n = pos - m;
    Mat A(pos, pos);
    std::vector<double> b(pos, 0.0);
    for (auto& it : factors) {  // ThreadsConstructA (single thread here: same sums, different order)
        const int nr = it->cost_function->num_residuals;
        for (size_t i = 0; i < it->parameter_blocks.size(); i++) {
            const int idx_i = parameter_block_idx[it->parameter_blocks[i]];
            const int size_i = localSize(parameter_block_size[it->parameter_blocks[i]]);
            const Mat& Ji = it->jacobians[i];
            for (size_t j = i; j < it->parameter_blocks.size(); j++) {
                const int idx_j = parameter_block_idx[it->parameter_blocks[j]];
                const int size_j = localSize(parameter_block_size[it->parameter_blocks[j]]);
                const Mat& Jj = it->jacobians[j];
                for (int a = 0; a < size_i; a++)
                    for (int c = 0; c < size_j; c++) {
                        double s = 0;
                        for (int k = 0; k < nr; k++) s += Ji(k, a) * Jj(k, c);
                        if (i == j)
                            A(idx_i + a, idx_j + c) += s;
                        else {
                            A(idx_i + a, idx_j + c) += s;
                            A(idx_j + c, idx_i + a) = A(idx_i + a, idx_j + c);
                        }
                    }
            }
            for (int a = 0; a < size_i; a++) {
                double s = 0;
                for (int k = 0; k < nr; k++) s += Ji(k, a) * it->residuals[k];
                b[idx_i + a] += s;
            }
        }
    }
    Mat Amm(m, m);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Amm(i, j) = 0.5 * (A(i, j) + A(j, i));
    std::vector<double> w;
    Mat V;
    Mat Amm_inv(m, m);
    if (m > 0) {
        if (eigen_ql) sym_eigen_ql(Amm, w, V);
        else sym_eigen(Amm, w, V);
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) {
                double s = 0;
                for (int k = 0; k < m; k++) s += V(i, k) * (w[k] > eps ? 1.0 / w[k] : 0.0) * V(j, k);
                Amm_inv(i, j) = s;
            }
    }
    const Mat Amr = A.block(0, m, m, n), Arm = A.block(m, 0, n, m), Arr = A.block(m, m, n, n);
    const Mat T = Arm * Amm_inv;
    Mat Ap = Arr - T * Amr;
    std::vector<double> bp(n);
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int k = 0; k < m; k++) s += T(i, k) * b[k];
        bp[i] = b[m + i] - s;
    }
    A_debug = Ap;
    b_debug = bp;
    std::vector<double> S;
    Mat V2;
    if (eigen_ql) sym_eigen_ql(Ap, S, V2);
    else sym_eigen(Ap, S, V2);
    linearized_jacobians = Mat(n, n);
    linearized_residuals.assign(n, 0.0);
    for (int k = 0; k < n; k++) {
        const double s = S[k] > eps ? S[k] : 0.0, sinv = S[k] > eps ? 1.0 / S[k] : 0.0;
        const double ss = std::sqrt(s), sis = std::sqrt(sinv);
        double vb = 0;
        for (int i = 0; i < n; i++) {
            linearized_jacobians(k, i) = ss * V2(i, k);
            vb += V2(i, k) * bp[i];
        }
        linearized_residuals[k] = sis * vb;
    }
}

std::vector<double*> MarginalizationInfo::getParameterBlocks(std::map<double*, double*>& addr_shift) {
    std::vector<double*> keep_block_addr;
    keep_block_size.clear();
    keep_block_idx.clear();
    keep_block_data.clear();
    for (double* a : block_order)
        if (parameter_block_idx[a] >= m) {
            keep_block_size.push_back(parameter_block_size[a]);
            keep_block_idx.push_back(parameter_block_idx[a]);
            keep_block_data.push_back(parameter_block_data[a]);
            keep_block_addr.push_back(addr_shift[a]);
        }
    return keep_block_addr;
}

MarginalizationFactor::MarginalizationFactor(const MarginalizationInfo* info) : marginalization_info(info) {
    for (int s : info->keep_block_size) block_sizes.push_back(s);
    num_residuals = info->n;
}

bool MarginalizationFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    const MarginalizationInfo* mi = marginalization_info;
    const int n = mi->n, m = mi->m;
    std::vector<double> dx(n, 0.0);
    for (size_t i = 0; i < mi->keep_block_size.size(); i++) {
        const int size = mi->keep_block_size[i], idx = mi->keep_block_idx[i] - m;
        const double* x = parameters[i];
        const double* x0 = mi->keep_block_data[i].data();
        if (size != 7) {
            for (int k = 0; k < size; k++) dx[idx + k] = x[k] - x0[k];
        } else {
            for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - x0[k];
            const Quat dq = Quat(x0[6], x0[3], x0[4], x0[5]).inverse() * Quat(x[6], x[3], x[4], x[5]);
            V3 v = 2.0 * dq.vec();
            if (!(dq.w >= 0)) v = 2.0 * -dq.vec();
            dx[idx + 3] = v.x;
            dx[idx + 4] = v.y;
            dx[idx + 5] = v.z;
        }
    }
    for (int i = 0; i < n; i++) {
        double s = mi->linearized_residuals[i];
        for (int k = 0; k < n; k++) s += mi->linearized_jacobians(i, k) * dx[k];
        residuals[i] = s;
    }
    if (jacobians)
        for (size_t i = 0; i < mi->keep_block_size.size(); i++)
            if (jacobians[i]) {
                const int size = mi->keep_block_size[i], local = MarginalizationInfo::localSize(size);
                const int idx = mi->keep_block_idx[i] - m;
                for (int r = 0; r < n; r++)
                    for (int c = 0; c < size; c++)
                        jacobians[i][r * size + c] = c < local ? mi->linearized_jacobians(r, idx + c) : 0.0;
            }
    return true;
}

}  // namespace orc

// Host-side F-matrix RANSAC of the tracker (GPU version is row "next-2").
//
// Stands in for cv::findFundamentalMat(un_cur_pts, un_forw_pts, FM_RANSAC, F_THRESHOLD, 0.99, status)
// at feature_tracker/src/feature_tracker.cpp:191.  Same published algorithm as OpenCV's
// calib3d (7-point minimal solver, MWC random stream seeded with ~0, adaptive iteration count, symmetric
// epipolar error against a float threshold; LMedS when fewer than 15 correspondences), float64 arithmetic.
#include "fm_ransac.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace vb {
namespace {

class Mwc {  // OpenCV's multiply-with-carry generator
  public:
    explicit Mwc(uint64_t seed) : s_(seed) {}
    unsigned next() {
        s_ = (uint64_t)(unsigned)s_ * 4164903690U + (unsigned)(s_ >> 32);
        return (unsigned)s_;
    }
    int below(int n) { return n == 0 ? 0 : (int)(next() % (unsigned)n); }

  private:
    uint64_t s_;
};

using Vec9 = std::array<double, 9>;

// Orthonormal basis (two vectors) of the null space of the 7x9 epipolar system: rows are
// orthogonalised by cyclic one-sided Jacobi rotations, then the complement is grown from +-1/9 sign
// vectors by double Gram-Schmidt.
void null_space(std::array<Vec9, 9>& r, Vec9& n1, Vec9& n2) {
    const int n = 7, m = 9;
    const double eps = DBL_EPSILON * 10;
    double w[9];
    auto dot = [&](const Vec9& a, const Vec9& b) {
        double s = 0;
        for (int k = 0; k < m; k++) s += a[k] * b[k];
        return s;
    };
    for (int i = 0; i < n; i++) w[i] = dot(r[i], r[i]);
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double a = w[i], b = w[j], p = dot(r[i], r[j]);
                if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = std::sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = std::sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; k++) {
                    const double t0 = c * r[i][k] + s * r[j][k];
                    const double t1 = -s * r[i][k] + c * r[j][k];
                    r[i][k] = t0;
                    r[j][k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                w[i] = a;
                w[j] = b;
                rotated = true;
            }
        if (!rotated) break;
    }
    for (int i = 0; i < n; i++) w[i] = std::sqrt(dot(r[i], r[i]));
    for (int i = 0; i < n - 1; i++) {  // selection sort by singular value, descending
        int j = i;
        for (int k = i + 1; k < n; k++)
            if (w[j] < w[k]) j = k;
        if (i != j) {
            std::swap(w[i], w[j]);
            std::swap(r[i], r[j]);
        }
    }
    Mwc rng(0x12345678);
    for (int i = 0; i < m; i++) {
        double sd = i < n ? w[i] : 0;
        for (int attempt = 0; attempt < 100 && sd <= DBL_MIN; attempt++) {
            const double v0 = 1. / m;
            for (int k = 0; k < m; k++) r[i][k] = (rng.next() & 256) != 0 ? v0 : -v0;
            for (int pass = 0; pass < 2; pass++)
                for (int j = 0; j < i; j++) {
                    sd = dot(r[i], r[j]);
                    double asum = 0;
                    for (int k = 0; k < m; k++) {
                        const double t = r[i][k] - sd * r[j][k];
                        r[i][k] = t;
                        asum += std::abs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < m; k++) r[i][k] *= asum;
                }
            sd = std::sqrt(dot(r[i], r[i]));
        }
        const double s = sd > DBL_MIN ? 1 / sd : 0.;
        for (int k = 0; k < m; k++) r[i][k] *= s;
    }
    n1 = r[7];
    n2 = r[8];
}

int cubic_roots(double a0, double a1, double a2, double a3, double x[3]) {
    x[0] = x[1] = x[2] = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return a3 == 0 ? -1 : 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = std::sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (std::fabs(q1) > std::fabs(q2)) {
            x[0] = q1 / a1;
            x[1] = a3 / q1;
        } else {
            x[0] = q2 / a1;
            x[1] = a3 / q2;
        }
        return d > 0 ? 2 : 1;
    }
    a0 = 1. / a0;
    a1 *= a0;
    a2 *= a0;
    a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = std::acos(R / std::sqrt(Qcubed));
        const double t0 = -2 * std::sqrt(Q), t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * std::cos(t1) - t2;
        x[1] = t0 * std::cos(t1 + (2. * M_PI / 3)) - t2;
        x[2] = t0 * std::cos(t1 + (4. * M_PI / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) {
            x[0] = -2 * std::pow(R, 1. / 3) - a1 / 3;
            x[1] = std::pow(R, 1. / 3) - a1 / 3;
        } else {
            x[0] = 2 * std::pow(-R, 1. / 3) - a1 / 3;
            x[1] = -std::pow(-R, 1. / 3) - a1 / 3;
        }
        const int n = x[0] == x[1] ? 1 : 2;
        if (n == 1) x[1] = 0;
        return n;
    }
    d = std::sqrt(-d);
    double e = std::pow(d + std::fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

// Up to three fundamental matrices through 7 correspondences.
int seven_point(const float* m1, const float* m2, double F[27]) {
    std::array<Vec9, 9> rows{};
    for (int i = 0; i < 7; i++) {
        const double x0 = m1[2 * i], y0 = m1[2 * i + 1], x1 = m2[2 * i], y1 = m2[2 * i + 1];
        rows[i] = {x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1.0};
    }
    Vec9 f1, f2;
    null_space(rows, f1, f2);
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    // det(lambda f1 + f2) = c0 l^3 + c1 l^2 + c2 l + c3
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    const double c3 = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    const double c2 = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
                      f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
                      f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
                      f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7];
    t1 = f1[3] * f1[8] - f1[5] * f1[6];
    t2 = f1[3] * f1[7] - f1[4] * f1[6];
    const double c0 = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    const double c1 = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
                      f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
                      f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
                      f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    double roots[3];
    const int n = cubic_roots(c0, c1, c2, c3, roots);
    if (n < 1 || n > 3) return n;
    for (int k = 0; k < n; k++) {
        double* f = F + 9 * k;
        double lambda = roots[k], mu = 1.;
        const double s = f1[8] * roots[k] + f2[8];
        if (std::fabs(s) > DBL_EPSILON) {
            mu = 1. / s;
            lambda *= mu;
            f[8] = 1.;
        } else
            f[8] = 0.;
        for (int i = 0; i < 8; i++) f[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

void epipolar_errors(const float* m1, const float* m2, int count, const double* F, float* err) {
    for (int i = 0; i < count; i++) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6];
        b = F[1] * x2 + F[4] * y2 + F[7];
        c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        err[i] = (float)std::max(d1 * d1 * s1, d2 * d2 * s2);
    }
}

bool last_point_collinear(const float* p) {
    const int i = 6;
    for (int j = 0; j < i; j++) {
        const double dx1 = p[2 * j] - p[2 * i], dy1 = p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            const double dx2 = p[2 * k] - p[2 * i], dy2 = p[2 * k + 1] - p[2 * i + 1];
            if (std::fabs(dx2 * dy1 - dy2 * dx1) <=
                FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
                return true;
        }
    }
    return false;
}

bool draw_sample(const float* m1, const float* m2, int count, Mwc& rng, int max_attempts, float* s1, float* s2) {
    int idx[7];
    for (int attempt = 0; attempt < max_attempts; attempt++) {
        for (int i = 0; i < 7; ++i) {
            int c;
            do {
                c = rng.below(count);
            } while (std::find(idx, idx + i, c) != idx + i);
            idx[i] = c;
            std::memcpy(s1 + 2 * i, m1 + 2 * c, 2 * sizeof(float));
            std::memcpy(s2 + 2 * i, m2 + 2 * c, 2 * sizeof(float));
        }
        if (!last_point_collinear(s1) && !last_point_collinear(s2)) return true;
    }
    return false;
}

int updated_iterations(double conf, double outlier_ratio, int max_iters) {
    conf = std::min(std::max(conf, 0.), 1.);
    outlier_ratio = std::min(std::max(outlier_ratio, 0.), 1.);
    double num = std::max(1. - conf, DBL_MIN);
    double denom = 1. - std::pow(1. - outlier_ratio, 7);
    if (denom < DBL_MIN) return 0;
    num = std::log(num);
    denom = std::log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

int mark_inliers(const std::vector<float>& err, double thresh, uint8_t* mask) {
    const float t = (float)(thresh * thresh);
    int good = 0;
    for (size_t i = 0; i < err.size(); i++) {
        mask[i] = err[i] <= t;
        good += mask[i];
    }
    return good;
}

}  // namespace

bool fundamental_ransac_mask(const float* m1, const float* m2, int count, double threshold, double confidence,
                             uint8_t* status) {
    return fundamental_ransac(m1, m2, count, threshold, confidence, status, nullptr);
}

bool fundamental_ransac(const float* m1, const float* m2, int count, double threshold, double confidence, uint8_t* status,
                        double* F) {
    std::memset(status, 0, count);
    if (count < 7) return false;
    if (threshold <= 0) threshold = 3;
    if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
    double models[27], best[9];
    if (count == 7) {
        if (seven_point(m1, m2, models) <= 0) return false;
        std::memset(status, 1, count);
        if (F) std::memcpy(F, models, sizeof(best));
        return true;
    }
    std::vector<float> err(count);
    std::vector<uint8_t> mask(count);
    float s1[14], s2[14];
    Mwc rng(~0ull);
    if (count >= 15) {
        int niters = 1000, best_good = 0;
        for (int iter = 0; iter < niters; iter++) {
            if (!draw_sample(m1, m2, count, rng, 10000, s1, s2)) {
                if (iter == 0) return false;
                break;
            }
            const int nm = seven_point(s1, s2, models);
            for (int k = 0; k < nm; k++) {
                epipolar_errors(m1, m2, count, models + 9 * k, err.data());
                const int good = mark_inliers(err, threshold, mask.data());
                if (good > std::max(best_good, 6)) {
                    std::memcpy(status, mask.data(), count);
                    std::memcpy(best, models + 9 * k, sizeof(best));
                    best_good = good;
                    niters = updated_iterations(confidence, (double)(count - good) / count, niters);
                }
            }
        }
        if (best_good > 0) {
            if (F) std::memcpy(F, best, sizeof(best));
            return true;
        }
        std::memset(status, 0, count);
        return false;
    }
    // fewer than 15 correspondences: least-median-of-squares
    double min_median = DBL_MAX;
    const int niters = std::max(updated_iterations(confidence, 0.45, 1000), 3);
    std::vector<float> sorted(count);
    for (int iter = 0; iter < niters; iter++) {
        if (!draw_sample(m1, m2, count, rng, 1000, s1, s2)) {
            if (iter == 0) return false;
            break;
        }
        const int nm = seven_point(s1, s2, models);
        for (int k = 0; k < nm; k++) {
            epipolar_errors(m1, m2, count, models + 9 * k, sorted.data());
            std::nth_element(sorted.begin(), sorted.begin() + count / 2, sorted.end());
            const double median = sorted[count / 2];
            if (median < min_median) {
                min_median = median;
                std::memcpy(best, models + 9 * k, sizeof(best));
            }
        }
    }
    if (min_median == DBL_MAX) return false;
    const double sigma = std::max(2.5 * 1.4826 * (1 + 5. / (count - 7)) * std::sqrt(min_median), 0.001);
    epipolar_errors(m1, m2, count, best, err.data());
    mark_inliers(err, sigma, status);
    if (F) std::memcpy(F, best, sizeof(best));
    return true;
}

}  // namespace vb

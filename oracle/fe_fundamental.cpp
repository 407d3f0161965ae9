// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of cv::findFundamentalMat(pts1, pts2, FM_RANSAC, thr, conf, status) as called by
// FeatureTracker::rejectWithF (feature_tracker/src/feature_tracker.cpp:169-202, call at :191).
// OpenCV's source is not in /root/reference (calib3d/fundam.cpp + ptsetreg.cpp of OpenCV 3.3.1/4.x);
// this follows the published algorithm: MWC RNG seeded with 2^64-1, 7-point minimal solver
// (null space of the 7x9 epipolar system, cubic det constraint, F(3,3)=1 scaling), symmetric
// epipolar distance with float threshold, adaptive iteration count; LMedS (300 iterations) when
// fewer than 15 correspondences.  Pinned against cv2 4.13 in tests/test_oracle_frontend.py
// (inlier masks identical; 7-point F matrices to 1e-9).
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

struct CvRng {  // cv::RNG (multiply-with-carry)
    uint64_t state;
    explicit CvRng(uint64_t s) : state(s) {}
    inline unsigned next() {
        state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    inline int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// Null space (2 vectors) of a 7x9 matrix the way cv::SVDecomp(A, W, U, Vt, MODIFY_A|FULL_UV) produces
// Vt rows 7 and 8: one-sided Jacobi on the 7 rows (core/lapack.cpp JacobiSVDImpl_), then the
// orthogonal complement is completed from +-1/m sign vectors drawn from RNG(0x12345678) with two
// Gram-Schmidt passes.  Any accurate null-space basis gives the same fundamental matrices; following
// the OpenCV recipe keeps the intermediate cubic identical too.
void nullspace_7x9(double A[7][9], double f1[9], double f2[9]) {
    const int n = 7, m = 9;
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    double W[9];
    double At[9][9];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) At[i][k] = A[i][k];
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = sd;
    }
    const int max_iter = std::max(m, 30);
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double* Ai = At[i];
                double* Aj = At[j];
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = std::sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = std::sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0;
                    Aj[k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = std::sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            std::swap(W[i], W[j]);
            for (int k = 0; k < m; k++) std::swap(At[i][k], At[j][k]);
        }
    }
    CvRng rng(0x12345678);
    for (int i = 0; i < m; i++) {
        double sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / m;
            for (int k = 0; k < m; k++) At[i][k] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0;
                    for (int k = 0; k < m; k++) sd += At[i][k] * At[j][k];
                    double asum = 0;
                    for (int k = 0; k < m; k++) {
                        double t = At[i][k] - sd * At[j][k];
                        At[i][k] = t;
                        asum += std::abs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < m; k++) At[i][k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
            sd = std::sqrt(sd);
        }
        double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < m; k++) At[i][k] *= s;
    }
    for (int k = 0; k < 9; k++) {
        f1[k] = At[7][k];
        f2[k] = At[8][k];
    }
}

// cv::solveCubic for a0 x^3 + a1 x^2 + a2 x + a3 = 0 (core/mathfuncs.cpp)
int solve_cubic(const double c[4], double r[3]) {
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    double x0 = 0., x1 = 0., x2 = 0.;
    int n = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0)
                n = a3 == 0 ? -1 : 0;
            else {
                x0 = -a3 / a2;
                n = 1;
            }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = std::sqrt(d);
                double q1 = (-a2 + d) * 0.5;
                double q2 = (a2 + d) * -0.5;
                if (std::fabs(q1) > std::fabs(q2)) {
                    x0 = q1 / a1;
                    x1 = a3 / q1;
                } else {
                    x0 = q2 / a1;
                    x1 = a3 / q2;
                }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0;
        a1 *= a0;
        a2 *= a0;
        a3 *= a0;
        double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        double Qcubed = Q * Q * Q;
        double d = Qcubed - R * R;
        if (d > 0) {
            double theta = std::acos(R / std::sqrt(Qcubed));
            double sqrtQ = std::sqrt(Q);
            double t0 = -2 * sqrtQ;
            double t1 = theta * (1. / 3);
            double t2 = a1 * (1. / 3);
            x0 = t0 * std::cos(t1) - t2;
            x1 = t0 * std::cos(t1 + (2. * M_PI / 3)) - t2;
            x2 = t0 * std::cos(t1 + (4. * M_PI / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R >= 0) {
                x0 = -2 * std::pow(R, 1. / 3) - a1 / 3;
                x1 = std::pow(R, 1. / 3) - a1 / 3;
            } else {
                x0 = 2 * std::pow(-R, 1. / 3) - a1 / 3;
                x1 = -std::pow(-R, 1. / 3) - a1 / 3;
            }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = std::sqrt(-d);
            double e = std::pow(d + std::fabs(R), 1. / 3);
            if (R > 0) e = -e;
            x0 = (e + Q / e) - a1 * (1. / 3);
            n = 1;
        }
    }
    r[0] = x0;
    r[1] = x1;
    r[2] = x2;
    return n;
}

// run7Point: up to 3 fundamental matrices (row-major 3x3 each) from 7 correspondences (f32 points).
int run_7point(const float* m1, const float* m2, double* fmatrix) {
    double a[7][9], f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    for (int i = 0; i < 7; i++) {
        double x0 = m1[2 * i], y0 = m1[2 * i + 1];
        double x1 = m2[2 * i], y1 = m2[2 * i + 1];
        a[i][0] = x1 * x0;
        a[i][1] = x1 * y0;
        a[i][2] = x1;
        a[i][3] = y1 * x0;
        a[i][4] = y1 * y0;
        a[i][5] = y1;
        a[i][6] = x0;
        a[i][7] = y0;
        a[i][8] = 1;
    }
    nullspace_7x9(a, f1, f2);
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7];
    double t1 = f2[3] * f2[8] - f2[5] * f2[6];
    double t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
           f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
           f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7];
    t1 = f1[3] * f1[8] - f1[5] * f1[6];
    t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
           f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
           f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return n;
    for (int k = 0; k < n; k++, fmatrix += 9) {
        double lambda = r[k], mu = 1.;
        double s = f1[8] * r[k] + f2[8];
        if (std::fabs(s) > DBL_EPSILON) {
            mu = 1. / s;
            lambda *= mu;
            fmatrix[8] = 1.;
        } else
            fmatrix[8] = 0.;
        for (int i = 0; i < 8; i++) fmatrix[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

void compute_error(const float* m1, const float* m2, int count, const double* F, float* err) {
    for (int i = 0; i < count; i++) {
        double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2];
        double b = F[3] * x1 + F[4] * y1 + F[5];
        double c = F[6] * x1 + F[7] * y1 + F[8];
        double s2 = 1. / (a * a + b * b);
        double d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6];
        b = F[1] * x2 + F[4] * y2 + F[7];
        c = F[2] * x2 + F[5] * y2 + F[8];
        double s1 = 1. / (a * a + b * b);
        double d1 = x1 * a + y1 * b + c;
        err[i] = (float)std::max(d1 * d1 * s1, d2 * d2 * s2);
    }
}

bool have_collinear(const float* p, int count) {
    int i = count - 1;
    for (int j = 0; j < i; j++) {
        double dx1 = p[2 * j] - p[2 * i];
        double dy1 = p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            double dx2 = p[2 * k] - p[2 * i];
            double dy2 = p[2 * k + 1] - p[2 * i + 1];
            if (std::fabs(dx2 * dy1 - dy2 * dx1) <=
                FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
                return true;
        }
    }
    return false;
}

bool get_subset(const float* m1, const float* m2, int count, CvRng& rng, int max_attempts, float* ms1, float* ms2) {
    int idx[7];
    int iters = 0;
    for (; iters < max_attempts; iters++) {
        for (int i = 0; i < 7; ++i) {
            int idx_i;
            for (idx_i = rng.uniform(0, count); std::find(idx, idx + i, idx_i) != idx + i;
                 idx_i = rng.uniform(0, count)) {
            }
            idx[i] = idx_i;
            ms1[2 * i] = m1[2 * idx_i];
            ms1[2 * i + 1] = m1[2 * idx_i + 1];
            ms2[2 * i] = m2[2 * idx_i];
            ms2[2 * i + 1] = m2[2 * idx_i + 1];
        }
        if (!have_collinear(ms1, 7) && !have_collinear(ms2, 7)) break;
    }
    return iters < max_attempts;
}

int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = std::max(p, 0.);
    p = std::min(p, 1.);
    ep = std::max(ep, 0.);
    ep = std::min(ep, 1.);
    double num = std::max(1. - p, DBL_MIN);
    double denom = 1. - std::pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = std::log(num);
    denom = std::log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

int find_inliers(const float* m1, const float* m2, int count, const double* F, double thresh, float* err,
                 uint8_t* mask) {
    compute_error(m1, m2, count, F, err);
    float t = (float)(thresh * thresh);
    int nz = 0;
    for (int i = 0; i < count; i++) {
        int f = err[i] <= t;
        mask[i] = (uint8_t)f;
        nz += f;
    }
    return nz;
}

}  // namespace

extern "C" {

// 7-point solver exposed for tests (returns number of models, F's row-major in out[27]).
int orc_fm_7point(const float* m1, const float* m2, double* out) { return run_7point(m1, m2, out); }

// findFundamentalMat(FM_RANSAC, thr, conf, mask).  Returns 1 when a model was found (mask valid),
// 0 otherwise (mask all zero: the std::vector<uchar> OpenCV resized stays value-initialised).
// n < 7 -> 0; n == 7 -> all ones if the solver returns a model; 8 <= n < 15 -> LMedS.
int orc_find_fundamental_ransac(const float* m1, const float* m2, int count, double thr, double conf,
                                uint8_t* mask_out, int* iters_run) {
    std::memset(mask_out, 0, count);
    if (iters_run) *iters_run = 0;
    if (count < 7) return 0;
    if (thr <= 0) thr = 3;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    double model[27], best_model[9];
    if (count == 7) {
        int n = run_7point(m1, m2, model);
        if (n <= 0) return 0;
        std::memset(mask_out, 1, count);
        return 1;
    }
    std::vector<float> err(count);
    std::vector<uint8_t> mask(count);
    float ms1[14], ms2[14];
    CvRng rng((uint64_t)-1);
    if (count >= 15) {
        int niters = 1000, max_good = 0, iter;
        for (iter = 0; iter < niters; iter++) {
            if (!get_subset(m1, m2, count, rng, 10000, ms1, ms2)) {
                if (iter == 0) return 0;
                break;
            }
            int nmodels = run_7point(ms1, ms2, model);
            if (nmodels <= 0) continue;
            for (int i = 0; i < nmodels; i++) {
                int good = find_inliers(m1, m2, count, model + 9 * i, thr, err.data(), mask.data());
                if (good > std::max(max_good, 6)) {
                    std::memcpy(mask_out, mask.data(), count);
                    std::memcpy(best_model, model + 9 * i, sizeof(best_model));
                    max_good = good;
                    niters = ransac_update_num_iters(conf, (double)(count - good) / count, 7, niters);
                }
            }
        }
        if (iters_run) *iters_run = iter;
        if (max_good > 0) return 1;
        std::memset(mask_out, 0, count);
        return 0;
    }
    // LMedS branch (8 <= count < 15): ptsetreg.cpp LMeDSPointSetRegistrator::run
    double min_median = DBL_MAX;
    int niters = ransac_update_num_iters(conf, 0.45, 7, 1000);
    niters = std::max(niters, 3);
    std::vector<float> errs(count);
    int iter;
    for (iter = 0; iter < niters; iter++) {
        if (!get_subset(m1, m2, count, rng, 1000, ms1, ms2)) {
            if (iter == 0) return 0;
            break;
        }
        int nmodels = run_7point(ms1, ms2, model);
        if (nmodels <= 0) continue;
        for (int i = 0; i < nmodels; i++) {
            compute_error(m1, m2, count, model + 9 * i, errs.data());
            std::nth_element(errs.begin(), errs.begin() + count / 2, errs.end());
            double median = errs[count / 2];
            if (median < min_median) {
                min_median = median;
                std::memcpy(best_model, model + 9 * i, sizeof(best_model));
            }
        }
    }
    if (iters_run) *iters_run = iter;
    if (min_median < DBL_MAX) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (count - 7)) * std::sqrt(min_median);
        sigma = std::max(sigma, 0.001);
        int good = find_inliers(m1, m2, count, best_model, sigma, err.data(), mask_out);
        return good >= 7 ? 1 : 1;  // the mask is written either way; OpenCV returns F only if good >= 7
    }
    return 0;
}

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY.  Small dense linear algebra used by the back-end restatement
// (stands in for the Eigen calls of the reference; Eigen is not available offline).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return i == 0 ? x : i == 1 ? y : z; }
    double operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
    V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
    V3 operator/(double s) const { return V3(x / s, y / s, z / s); }
    V3& operator+=(const V3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    V3 cross(const V3& o) const { return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    double norm() const { return std::sqrt(dot(*this)); }
    V3 normalized() const { return *this / norm(); }
};
inline V3 operator*(double s, const V3& v) { return v * s; }

struct M3 {
    double m[9];
    M3() { std::memset(m, 0, sizeof(m)); }
    static M3 Identity() { M3 r; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
    double& operator()(int i, int j) { return m[3 * i + j]; }
    double operator()(int i, int j) const { return m[3 * i + j]; }
    M3 operator*(const M3& o) const {
        M3 r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += (*this)(i, k) * o(k, j);
                r(i, j) = s;
            }
        return r;
    }
    V3 operator*(const V3& v) const {
        return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
    }
    M3 operator*(double s) const { M3 r; for (int i = 0; i < 9; i++) r.m[i] = m[i] * s; return r; }
    M3 operator+(const M3& o) const { M3 r; for (int i = 0; i < 9; i++) r.m[i] = m[i] + o.m[i]; return r; }
    M3 operator-(const M3& o) const { M3 r; for (int i = 0; i < 9; i++) r.m[i] = m[i] - o.m[i]; return r; }
    M3 operator-() const { M3 r; for (int i = 0; i < 9; i++) r.m[i] = -m[i]; return r; }
    M3 T() const { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = (*this)(j, i); return r; }
    V3 col(int j) const { return V3(m[j], m[3 + j], m[6 + j]); }
};
inline M3 operator*(double s, const M3& a) { return a * s; }

// Utility::skewSymmetric (utility/utility.h:31-38)
inline M3 skew(const V3& q) {
    M3 r;
    r(0, 1) = -q.z; r(0, 2) = q.y;
    r(1, 0) = q.z;  r(1, 2) = -q.x;
    r(2, 0) = -q.y; r(2, 1) = q.x;
    return r;
}

// Eigen::Quaterniond semantics (w, x, y, z)
struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
    Quat() {}
    Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
    V3 vec() const { return V3(x, y, z); }
    Quat operator*(const Quat& b) const {
        return Quat(w * b.w - x * b.x - y * b.y - z * b.z, w * b.x + x * b.w + y * b.z - z * b.y,
                    w * b.y + y * b.w + z * b.x - x * b.z, w * b.z + z * b.w + x * b.y - y * b.x);
    }
    V3 operator*(const V3& v) const {  // Eigen's _transformVector
        V3 uv = vec().cross(v);
        uv += uv;
        return v + w * uv + vec().cross(uv);
    }
    double sqnorm() const { return w * w + x * x + y * y + z * z; }
    Quat inverse() const {
        double n2 = sqnorm();
        return Quat(w / n2, -x / n2, -y / n2, -z / n2);
    }
    Quat normalized() const {
        double n = std::sqrt(sqnorm());
        return Quat(w / n, x / n, y / n, z / n);
    }
    M3 R() const {  // toRotationMatrix
        M3 r;
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
        const double twx = tx * w, twy = ty * w, twz = tz * w;
        const double txx = tx * x, txy = ty * x, txz = tz * x;
        const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
        r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz;       r(0, 2) = txz + twy;
        r(1, 0) = txy + twz;       r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy;       r(2, 1) = tyz + twx;       r(2, 2) = 1 - (txx + tyy);
        return r;
    }
    static Quat FromR(const M3& m) {  // Eigen's quaternion-from-matrix
        Quat q;
        double t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > 0) {
            t = std::sqrt(t + 1.0);
            q.w = 0.5 * t;
            t = 0.5 / t;
            q.x = (m(2, 1) - m(1, 2)) * t;
            q.y = (m(0, 2) - m(2, 0)) * t;
            q.z = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
            double c[3];
            c[i] = 0.5 * t;
            t = 0.5 / t;
            q.w = (m(k, j) - m(j, k)) * t;
            c[j] = (m(j, i) + m(i, j)) * t;
            c[k] = (m(k, i) + m(i, k)) * t;
            q.x = c[0]; q.y = c[1]; q.z = c[2];
        }
        return q;
    }
};

// Utility::deltaQ (utility/utility.h:16-28): [1, theta/2], NOT normalised
inline Quat deltaQ(const V3& theta) { return Quat(1.0, theta.x / 2.0, theta.y / 2.0, theta.z / 2.0); }

// Dynamic row-major matrix.
struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    static Mat Identity(int n) { Mat m(n, n); for (int i = 0; i < n; i++) m(i, i) = 1; return m; }
    double& operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
    Mat T() const { Mat t(c, r); for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) t(j, i) = (*this)(i, j); return t; }
    Mat operator*(const Mat& o) const {
        assert(c == o.r);
        Mat m(r, o.c);
        for (int i = 0; i < r; i++)
            for (int k = 0; k < c; k++) {
                const double a = (*this)(i, k);
                if (a == 0) continue;
                for (int j = 0; j < o.c; j++) m(i, j) += a * o(k, j);
            }
        return m;
    }
    Mat operator+(const Mat& o) const { Mat m = *this; for (size_t i = 0; i < d.size(); i++) m.d[i] += o.d[i]; return m; }
    Mat operator-(const Mat& o) const { Mat m = *this; for (size_t i = 0; i < d.size(); i++) m.d[i] -= o.d[i]; return m; }
    Mat operator*(double s) const { Mat m = *this; for (auto& v : m.d) v *= s; return m; }
    Mat block(int i0, int j0, int nr, int nc) const {
        Mat m(nr, nc);
        for (int i = 0; i < nr; i++) for (int j = 0; j < nc; j++) m(i, j) = (*this)(i0 + i, j0 + j);
        return m;
    }
    void set_block(int i0, int j0, const Mat& b) { for (int i = 0; i < b.r; i++) for (int j = 0; j < b.c; j++) (*this)(i0 + i, j0 + j) = b(i, j); }
    void set_block3(int i0, int j0, const M3& b) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) (*this)(i0 + i, j0 + j) = b(i, j); }
    M3 block3(int i0, int j0) const { M3 b; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) b(i, j) = (*this)(i0 + i, j0 + j); return b; }
    double max_coeff() const { return *std::max_element(d.begin(), d.end()); }
    double min_coeff() const { return *std::min_element(d.begin(), d.end()); }
};

// Cholesky A = L L^T (lower).  Returns false if A is not positive definite (Eigen LLT::info()).
inline bool cholesky(const Mat& A, Mat& L) {
    const int n = A.r;
    L = Mat(n, n);
    for (int j = 0; j < n; j++) {
        double s = A(j, j);
        for (int k = 0; k < j; k++) s -= L(j, k) * L(j, k);
        if (!(s > 0)) return false;
        const double ljj = std::sqrt(s);
        L(j, j) = ljj;
        for (int i = j + 1; i < n; i++) {
            double t = A(i, j);
            for (int k = 0; k < j; k++) t -= L(i, k) * L(j, k);
            L(i, j) = t / ljj;
        }
    }
    return true;
}
inline void chol_solve(const Mat& L, const double* b, double* x) {
    const int n = L.r;
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L(i, k) * y[k];
        y[i] = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = y[i];
        for (int k = i + 1; k < n; k++) s -= L(k, i) * x[k];
        x[i] = s / L(i, i);
    }
}

// General inverse by LU with partial pivoting (Eigen's MatrixBase::inverse() for dynamic/large fixed sizes).
inline Mat inverse_lu(const Mat& A) {
    const int n = A.r;
    Mat a = A, inv = Mat::Identity(n);
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++) if (std::fabs(a(i, k)) > std::fabs(a(p, k))) p = i;
        if (p != k) for (int j = 0; j < n; j++) { std::swap(a(k, j), a(p, j)); std::swap(inv(k, j), inv(p, j)); }
        const double piv = a(k, k);
        for (int i = k + 1; i < n; i++) {
            const double f = a(i, k) / piv;
            if (f == 0) continue;
            for (int j = k; j < n; j++) a(i, j) -= f * a(k, j);
            for (int j = 0; j < n; j++) inv(i, j) -= f * inv(k, j);
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        const double piv = a(k, k);
        for (int j = 0; j < n; j++) inv(k, j) /= piv;
        for (int i = 0; i < k; i++) {
            const double f = a(i, k);
            if (f == 0) continue;
            for (int j = 0; j < n; j++) inv(i, j) -= f * inv(k, j);
        }
    }
    return inv;
}

// Symmetric eigen-decomposition A = V diag(w) V^T by cyclic Jacobi (ascending eigenvalues like
// Eigen::SelfAdjointEigenSolver; eigenvector signs are arbitrary there as well).
inline void sym_eigen(const Mat& A, std::vector<double>& w, Mat& V) {
    const int n = A.r;
    Mat a = A;
    V = Mat::Identity(n);
    // A pair (p, q) is rotated when |a_pq| > 1e-15 sqrt|a_pp a_qq| (relative criterion of Jacobi methods: it
    // resolves the small eigenvalues of the badly scaled information matrices, entries 1e-3 .. 1e14, this is used
    // on).  Sweeps stop when every rotation of a sweep was the identity in floating point (c == 1): the matrix
    // is then a fixed point of the iteration.
    for (int sweep = 0; sweep < 60; sweep++) {
        bool changed = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a(p, q);
                if (apq == 0) continue;
                const double app = a(p, p), aqq = a(q, q);
                if (std::fabs(apq) <= 1e-300 + 1e-15 * std::sqrt(std::fabs(app * aqq))) continue;
                const double theta = (aqq - app) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                if (c != 1.0) changed = true;
                for (int k = 0; k < n; k++) {
                    const double akp = a(k, p), akq = a(k, q);
                    a(k, p) = c * akp - s * akq;
                    a(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = a(p, k), aqk = a(q, k);
                    a(p, k) = c * apk - s * aqk;
                    a(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
        if (!changed) break;
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int i, int j) { return a(i, i) < a(j, j); });
    w.resize(n);
    Mat Vs(n, n);
    for (int k = 0; k < n; k++) {
        w[k] = a(idx[k], idx[k]);
        for (int i = 0; i < n; i++) Vs(i, k) = V(i, idx[k]);
    }
    V = Vs;
}

// Second solver, used for the TIMED CPU baseline only (bench.py cpu_baseline / --impl reference): Householder
// tridiagonalisation + implicit-shift QL (EISPACK tred2 / tql2 as published in Wilkinson & Reinsch, Handbook for Automatic
// Computation II; the algorithm family of Eigen's SelfAdjointEigenSolver which the reference calls at
// marginalization_factor.cpp:268, :283), written here from the published algorithm: independent of the product's kernel
// template.  The parity tests keep the Jacobi solver above: it resolves small eigenvalues of these graded matrices to
// relative accuracy, whereas QL/QR (the reference included) delivers them to eps*|A| only — which of the round-off
// eigenvalues of the rank-deficient prior pass the reference's 1e-8 floor is therefore implementation noise, and it
// bounds the achievable state parity at ~1e-5.
inline void sym_eigen_ql(const Mat& A, std::vector<double>& w, Mat& V) {
    const int n = A.r;
    V = Mat(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V(i, j) = 0.5 * (A(i, j) + A(j, i));
    std::vector<double> d(n), e(n, 0.0);
    // ---- tred2: reduce to tridiagonal form, accumulating the orthogonal transformation in V
    for (int j = 0; j < n; j++) d[j] = V(n - 1, j);
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) {
                d[j] = V(i - 1, j);
                V(i, j) = 0.0;
                V(j, i) = 0.0;
            }
        } else {
            for (int k = 0; k < i; k++) {
                d[k] /= scale;
                h += d[k] * d[k];
            }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j];
                V(j, i) = f;
                g = e[j] + V(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) {
                    g += V(k, j) * d[k];
                    e[k] += V(k, j) * f;
                }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) {
                e[j] /= h;
                f += e[j] * d[j];
            }
            const double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j];
                g = e[j];
                for (int k = j; k <= i - 1; k++) V(k, j) -= (f * e[k] + g * d[k]);
                d[j] = V(i - 1, j);
                V(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) {
        V(n - 1, i) = V(i, i);
        V(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = V(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += V(k, i + 1) * V(k, j);
                for (int k = 0; k <= i; k++) V(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) V(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) {
        d[j] = V(n - 1, j);
        V(n - 1, j) = 0.0;
    }
    if (n > 0) V(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
    // ---- tql2: implicit-shift QL on the tridiagonal matrix
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    if (n > 0) e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::pow(2.0, -52.0);
    for (int l = 0; l < n; l++) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n) {
            if (std::fabs(e[m]) <= eps * tst1) break;
            m++;
        }
        if (m > l) {
            int iter = 0;
            do {
                iter++;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = V(k, i + 1);
                        V(k, i + 1) = s * V(k, i) + c * h;
                        V(k, i) = c * V(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int i, int j) { return d[i] < d[j]; });
    w.resize(n);
    Mat Vs(n, n);
    for (int k = 0; k < n; k++) {
        w[k] = d[idx[k]];
        for (int i = 0; i < n; i++) Vs(i, k) = V(i, idx[k]);
    }
    V = Vs;
}

// Right singular vector of the smallest singular value of A (rows x 4): one-sided Jacobi on the columns
// (JacobiSVD(...).matrixV().rightCols<1>() in FeatureManager::triangulate).
inline void smallest_right_singular_vector4(const Mat& A, double v[4]) {
    const int m = A.r, n = 4;
    Mat U = A, V = Mat::Identity(n);
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < m; k++) { a += U(k, p) * U(k, p); b += U(k, q) * U(k, q); g += U(k, p) * U(k, q); }
                if (std::fabs(g) <= 1e-17 * std::sqrt(a * b) || g == 0) continue;
                rotated = true;
                const double zeta = (b - a) / (2 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int k = 0; k < m; k++) { const double up = U(k, p), uq = U(k, q); U(k, p) = c * up - s * uq; U(k, q) = s * up + c * uq; }
                for (int k = 0; k < n; k++) { const double vp = V(k, p), vq = V(k, q); V(k, p) = c * vp - s * vq; V(k, q) = s * vp + c * vq; }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bn = 1e300;
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < m; k++) s += U(k, j) * U(k, j);
        if (s < bn) { bn = s; best = j; }
    }
    for (int k = 0; k < 4; k++) v[k] = V(k, best);
}

}  // namespace orc

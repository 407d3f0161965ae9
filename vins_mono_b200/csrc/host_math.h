// Host-side 3-vector / rotation helpers of the estimator shim (stand in for the Eigen calls in
// vins_estimator/src/estimator.cpp and utility/utility.h:70-112; Eigen is not a dependency here).
#pragma once
#include <cmath>
#include <cstring>

namespace hm {

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    Vec3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator*(double s) const { return {x * s, y * s, z * s}; }
    Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    double norm() const { return std::sqrt(x * x + y * y + z * z); }
    double operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }

struct Mat3 {
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double& operator()(int i, int j) { return m[3 * i + j]; }
    double operator()(int i, int j) const { return m[3 * i + j]; }
    Mat3 operator*(const Mat3& o) const {
        Mat3 r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r(i, j) = (*this)(i, 0) * o(0, j) + (*this)(i, 1) * o(1, j) + (*this)(i, 2) * o(2, j);
        return r;
    }
    Vec3 operator*(const Vec3& v) const {
        return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
    }
    Mat3 T() const {
        Mat3 r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r(i, j) = (*this)(j, i);
        return r;
    }
    Vec3 col(int j) const { return {m[j], m[3 + j], m[6 + j]}; }
};

struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
    Quat() {}
    Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
    Quat normalized() const {
        const double n = std::sqrt(w * w + x * x + y * y + z * z);
        return {w / n, x / n, y / n, z / n};
    }
    Mat3 R() const {
        Mat3 r;
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
        const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
        return r;
    }
    static Quat FromR(const Mat3& m) {
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
            const int j = (i + 1) % 3, k = (j + 1) % 3;
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

// Utility::deltaQ(theta).toRotationMatrix() for the small-angle update in processIMU
inline Mat3 deltaQ_R(const Vec3& th) { return Quat(1.0, th.x / 2.0, th.y / 2.0, th.z / 2.0).R(); }

inline Vec3 R2ypr(const Mat3& R) {  // degrees (utility.h:70-85)
    const Vec3 n = R.col(0), o = R.col(1), a = R.col(2);
    const double y = std::atan2(n.y, n.x);
    const double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
    const double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
    return Vec3(y, p, r) * (1.0 / M_PI * 180.0);
}
inline Mat3 ypr2R(const Vec3& ypr) {  // utility.h:88-112
    const double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
    Mat3 Rz, Ry, Rx;
    Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y);
    Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
    Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
    return Rz * Ry * Rx;
}

// Direction of the smallest singular value of a (rows x 4) matrix: one-sided Jacobi on its columns.
inline void null_direction4(const double* A, int rows, double v[4]) {
    double U[64 * 4], V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(U, A, sizeof(double) * rows * 4);
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < rows; k++) {
                    a += U[4 * k + p] * U[4 * k + p];
                    b += U[4 * k + q] * U[4 * k + q];
                    g += U[4 * k + p] * U[4 * k + q];
                }
                if (g == 0 || std::fabs(g) <= 1e-17 * std::sqrt(a * b)) continue;
                rotated = true;
                const double zeta = (b - a) / (2 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int k = 0; k < rows; k++) {
                    const double up = U[4 * k + p], uq = U[4 * k + q];
                    U[4 * k + p] = c * up - s * uq;
                    U[4 * k + q] = s * up + c * uq;
                }
                for (int k = 0; k < 4; k++) {
                    const double vp = V[4 * k + p], vq = V[4 * k + q];
                    V[4 * k + p] = c * vp - s * vq;
                    V[4 * k + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bn = 1e300;
    for (int j = 0; j < 4; j++) {
        double s = 0;
        for (int k = 0; k < rows; k++) s += U[4 * k + j] * U[4 * k + j];
        if (s < bn) bn = s, best = j;
    }
    for (int k = 0; k < 4; k++) v[k] = V[4 * k + best];
}

}  // namespace hm

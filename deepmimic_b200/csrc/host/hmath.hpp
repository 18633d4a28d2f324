// Small double-precision 3-vector / quaternion / 3x3 helpers for the host-side asset
// layer.  Quaternions are (w, x, y, z) like DeepMimic's pose vectors
// (R/DeepMimicCore/anim/KinTree.cpp:428-441).  Euler angles follow the reference's
// convention R = Rz(z) * Ry(y) * Rx(x) (R/DeepMimicCore/util/MathUtil.cpp:159-186).
#pragma once
#include <cmath>

namespace dmh {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    double& operator[](int i) { return (&x)[i]; }
    double operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(const V3& a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
    Quat() {}
    Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
};
inline Quat operator*(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
inline Quat conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }
inline double qnorm(const Quat& q) { return std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); }
inline Quat normalized(const Quat& q) {
    double n = qnorm(q);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline V3 rotate(const Quat& q, const V3& v) {
    // v' = v + 2 w (u x v) + 2 u x (u x v)
    V3 u(q.x, q.y, q.z);
    V3 t = 2.0 * cross(u, v);
    return v + q.w * t + cross(u, t);
}
inline Quat axis_angle(const V3& axis, double theta) {
    double c = std::cos(0.5 * theta), s = std::sin(0.5 * theta);
    return {c, s * axis.x, s * axis.y, s * axis.z};
}

struct M3 {
    double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
};
inline M3 mul(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
inline M3 transpose(const M3& a) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
inline V3 mul(const M3& a, const V3& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 euler_to_mat(const V3& e) {
    double xs = std::sin(e.x), xc = std::cos(e.x), ys = std::sin(e.y), yc = std::cos(e.y), zs = std::sin(e.z),
           zc = std::cos(e.z);
    M3 r;
    r.m[0][0] = yc * zc;  r.m[0][1] = xs * ys * zc - xc * zs;  r.m[0][2] = xc * ys * zc + xs * zs;
    r.m[1][0] = yc * zs;  r.m[1][1] = xs * ys * zs + xc * zc;  r.m[1][2] = xc * ys * zs - xs * zc;
    r.m[2][0] = -ys;      r.m[2][1] = xs * yc;                 r.m[2][2] = xc * yc;
    return r;
}
inline Quat mat_to_quat(const M3& a) {
    const auto& m = a.m;
    double tr = m[0][0] + m[1][1] + m[2][2];
    Quat q;
    if (tr > 0) {
        double S = std::sqrt(tr + 1.0) * 2;
        q = {0.25 * S, (m[2][1] - m[1][2]) / S, (m[0][2] - m[2][0]) / S, (m[1][0] - m[0][1]) / S};
    } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
        double S = std::sqrt(1.0 + m[0][0] - m[1][1] - m[2][2]) * 2;
        q = {(m[2][1] - m[1][2]) / S, 0.25 * S, (m[0][1] + m[1][0]) / S, (m[0][2] + m[2][0]) / S};
    } else if (m[1][1] > m[2][2]) {
        double S = std::sqrt(1.0 + m[1][1] - m[0][0] - m[2][2]) * 2;
        q = {(m[0][2] - m[2][0]) / S, (m[0][1] + m[1][0]) / S, 0.25 * S, (m[1][2] + m[2][1]) / S};
    } else {
        double S = std::sqrt(1.0 + m[2][2] - m[0][0] - m[1][1]) * 2;
        q = {(m[1][0] - m[0][1]) / S, (m[0][2] + m[2][0]) / S, (m[1][2] + m[2][1]) / S, 0.25 * S};
    }
    return q;
}
inline Quat euler_to_quat(const V3& e) { return normalized(mat_to_quat(euler_to_mat(e))); }
inline M3 quat_to_mat(const Quat& q) {
    double sw = q.w * q.w, sx = q.x * q.x, sy = q.y * q.y, sz = q.z * q.z;
    double inv = 1.0 / (sw + sx + sy + sz);
    M3 r;
    r.m[0][0] = (sx - sy - sz + sw) * inv;
    r.m[1][1] = (-sx + sy - sz + sw) * inv;
    r.m[2][2] = (-sx - sy + sz + sw) * inv;
    r.m[1][0] = 2 * (q.x * q.y + q.z * q.w) * inv;
    r.m[0][1] = 2 * (q.x * q.y - q.z * q.w) * inv;
    r.m[2][0] = 2 * (q.x * q.z - q.y * q.w) * inv;
    r.m[0][2] = 2 * (q.x * q.z + q.y * q.w) * inv;
    r.m[2][1] = 2 * (q.y * q.z + q.x * q.w) * inv;
    r.m[1][2] = 2 * (q.y * q.z - q.x * q.w) * inv;
    return r;
}
// angle normalised to [-pi, pi] exactly like the reference (fmod based).
inline double normalize_angle(double t) {
    double n = std::fmod(t, 2 * M_PI);
    if (n > M_PI) n -= 2 * M_PI;
    else if (n < -M_PI) n += 2 * M_PI;
    return n;
}
// rotation vector (axis * angle) of a unit quaternion; zero when sin(theta/2) is tiny, as in
// cMathUtil::QuaternionToAxisAngle (MathUtil.cpp:463-481).
inline V3 quat_to_rotvec(const Quat& q_in) {
    Quat q = q_in;
    if (q.w > 1) q = normalized(q);
    double s = std::sqrt(1 - q.w * q.w);
    if (s > 0.000001) {
        double th = normalize_angle(2 * std::acos(q.w));
        return V3(q.x / s * th, q.y / s * th, q.z / s * th);
    }
    return V3(0, 0, 0);
}

}  // namespace dmh

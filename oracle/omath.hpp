// ORACLE (test infrastructure, not product): math helpers restating
// R/DeepMimicCore/util/MathUtil.cpp (double precision, Eigen conventions) and the small part of
// Bullet's LinearMath (btVector3 / btQuaternion / btMatrix3x3, single precision) that the
// restated Bullet stage needs.  Bullet 2.88 is NOT in the reference tree; everything marked
// [B288-mem] is restated from knowledge of the upstream source.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// ------------------------------------------------------------------ double side (DeepMimic own math)
struct D3 {
    double x = 0, y = 0, z = 0;
    D3() {}
    D3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return (&x)[i]; }
    double operator[](int i) const { return (&x)[i]; }
};
inline D3 operator+(const D3& a, const D3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(const D3& a, const D3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator-(const D3& a) { return {-a.x, -a.y, -a.z}; }
inline D3 operator*(double s, const D3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline D3 operator*(const D3& a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline D3 operator/(const D3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline D3& operator+=(D3& a, const D3& b) { a = a + b; return a; }
inline double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 cross(const D3& a, const D3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double sqnorm(const D3& a) { return dot(a, a); }
inline double norm(const D3& a) { return std::sqrt(dot(a, a)); }

struct DQ {  // (w,x,y,z)
    double w = 1, x = 0, y = 0, z = 0;
    DQ() {}
    DQ(double a, double b, double c, double d) : w(a), x(b), y(c), z(d) {}
};
inline DQ operator*(const DQ& a, const DQ& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
inline DQ qconj(const DQ& q) { return {q.w, -q.x, -q.y, -q.z}; }
inline double qnorm(const DQ& q) { return std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); }
inline DQ qnormalized(const DQ& q) { double n = qnorm(q); return {q.w / n, q.x / n, q.y / n, q.z / n}; }

struct DM3 {
    double m[3][3];
    DM3() { std::memset(m, 0, sizeof(m)); m[0][0] = m[1][1] = m[2][2] = 1; }
    static DM3 zero() { DM3 r; std::memset(r.m, 0, sizeof(r.m)); return r; }
};
inline DM3 operator*(const DM3& a, const DM3& b) {
    DM3 r = DM3::zero();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
inline D3 operator*(const DM3& a, const D3& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline DM3 transpose(const DM3& a) { DM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
inline DM3 crossmat(const D3& a) {  // cMathUtil::CrossMat (MathUtil.cpp:217-225)
    DM3 r = DM3::zero();
    r.m[0][1] = -a.z; r.m[0][2] = a.y; r.m[1][0] = a.z; r.m[1][2] = -a.x; r.m[2][0] = -a.y; r.m[2][1] = a.x;
    return r;
}

// rigid transform x_out = R x_in + t  (the 4x4 tMatrix of the reference, kept as R|t)
struct DT {
    DM3 R;
    D3 t;
};
inline DT operator*(const DT& a, const DT& b) { return {a.R * b.R, a.R * b.t + a.t}; }
inline DT inv_rigid(const DT& a) { DM3 Rt = transpose(a.R); return {Rt, -(Rt * a.t)}; }  // cMathUtil::InvRigidMat
inline D3 xform(const DT& a, const D3& p) { return a.R * p + a.t; }

// cMathUtil::NormalizeAngle (MathUtil.cpp:33-46)
inline double NormalizeAngle(double theta) {
    double n = std::fmod(theta, 2 * M_PI);
    if (n > M_PI) n = -2 * M_PI + n;
    else if (n < -M_PI) n = 2 * M_PI + n;
    return n;
}
// cMathUtil::RotateMat(euler) (MathUtil.cpp:134-157): R = Rz * Ry * Rx
inline DM3 RotateMatEuler(const D3& e) {
    double xs = std::sin(e.x), xc = std::cos(e.x), ys = std::sin(e.y), yc = std::cos(e.y), zs = std::sin(e.z), zc = std::cos(e.z);
    DM3 r;
    r.m[0][0] = yc * zc; r.m[1][0] = yc * zs; r.m[2][0] = -ys;
    r.m[0][1] = xs * ys * zc - xc * zs; r.m[1][1] = xs * ys * zs + xc * zc; r.m[2][1] = xs * yc;
    r.m[0][2] = xc * ys * zc + xs * zs; r.m[1][2] = xc * ys * zs - xs * zc; r.m[2][2] = xc * yc;
    return r;
}
// cMathUtil::RotateMat(axis, theta) (MathUtil.cpp:159-176)
inline DM3 RotateMatAxis(const D3& a, double theta) {
    double c = std::cos(theta), s = std::sin(theta), x = a.x, y = a.y, z = a.z;
    DM3 r;
    r.m[0][0] = c + x * x * (1 - c); r.m[0][1] = x * y * (1 - c) - z * s; r.m[0][2] = x * z * (1 - c) + y * s;
    r.m[1][0] = y * x * (1 - c) + z * s; r.m[1][1] = c + y * y * (1 - c); r.m[1][2] = y * z * (1 - c) - x * s;
    r.m[2][0] = z * x * (1 - c) - y * s; r.m[2][1] = z * y * (1 - c) + x * s; r.m[2][2] = c + z * z * (1 - c);
    return r;
}
// cMathUtil::RotateMat(quaternion) (MathUtil.cpp:178-205)
inline DM3 RotateMatQuat(const DQ& q) {
    DM3 r;
    double sqw = q.w * q.w, sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z;
    double invs = 1 / (sqx + sqy + sqz + sqw);
    r.m[0][0] = (sqx - sqy - sqz + sqw) * invs; r.m[1][1] = (-sqx + sqy - sqz + sqw) * invs; r.m[2][2] = (-sqx - sqy + sqz + sqw) * invs;
    double t1 = q.x * q.y, t2 = q.z * q.w;
    r.m[1][0] = 2.0 * (t1 + t2) * invs; r.m[0][1] = 2.0 * (t1 - t2) * invs;
    t1 = q.x * q.z; t2 = q.y * q.w;
    r.m[2][0] = 2.0 * (t1 - t2) * invs; r.m[0][2] = 2.0 * (t1 + t2) * invs;
    t1 = q.y * q.z; t2 = q.x * q.w;
    r.m[2][1] = 2.0 * (t1 + t2) * invs; r.m[1][2] = 2.0 * (t1 - t2) * invs;
    return r;
}
// cMathUtil::RotMatToQuaternion (MathUtil.cpp:267-303)
inline DQ RotMatToQuaternion(const DM3& a) {
    const auto& m = a.m;
    double tr = m[0][0] + m[1][1] + m[2][2];
    DQ q;
    if (tr > 0) {
        double S = std::sqrt(tr + 1.0) * 2;
        q.w = 0.25 * S; q.x = (m[2][1] - m[1][2]) / S; q.y = (m[0][2] - m[2][0]) / S; q.z = (m[1][0] - m[0][1]) / S;
    } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
        double S = std::sqrt(1.0 + m[0][0] - m[1][1] - m[2][2]) * 2;
        q.w = (m[2][1] - m[1][2]) / S; q.x = 0.25 * S; q.y = (m[0][1] + m[1][0]) / S; q.z = (m[0][2] + m[2][0]) / S;
    } else if (m[1][1] > m[2][2]) {
        double S = std::sqrt(1.0 + m[1][1] - m[0][0] - m[2][2]) * 2;
        q.w = (m[0][2] - m[2][0]) / S; q.x = (m[0][1] + m[1][0]) / S; q.y = 0.25 * S; q.z = (m[1][2] + m[2][1]) / S;
    } else {
        double S = std::sqrt(1.0 + m[2][2] - m[0][0] - m[1][1]) * 2;
        q.w = (m[1][0] - m[0][1]) / S; q.x = (m[0][2] + m[2][0]) / S; q.y = (m[1][2] + m[2][1]) / S; q.z = 0.25 * S;
    }
    return q;
}
// cMathUtil::AxisAngleToQuaternion (MathUtil.cpp:450-461)
inline DQ AxisAngleToQuaternion(const D3& axis, double theta) {
    double c = std::cos(theta / 2), s = std::sin(theta / 2);
    return {c, s * axis.x, s * axis.y, s * axis.z};
}
// cMathUtil::EulerToAxisAngle + EulerToQuaternion (MathUtil.cpp:305-338,419-425)
inline DQ EulerToQuaternion(const D3& e) {
    double xs = std::sin(e.x), xc = std::cos(e.x), ys = std::sin(e.y), yc = std::cos(e.y), zs = std::sin(e.z), zc = std::cos(e.z);
    double c = (yc * zc + xs * ys * zs + xc * zc + xc * yc - 1) * 0.5;
    c = std::fmin(1.0, std::fmax(-1.0, c));
    double theta = std::acos(c);
    D3 axis(0, 0, 1);
    if (!(std::fabs(theta) < 0.00001)) {
        double m21 = xs * yc - xc * ys * zs + xs * zc;
        double m02 = xc * ys * zc + xs * zs + ys;
        double m10 = yc * zs - xs * ys * zc + xc * zs;
        double denom = std::sqrt(m21 * m21 + m02 * m02 + m10 * m10);
        axis = D3(m21 / denom, m02 / denom, m10 / denom);
    }
    return AxisAngleToQuaternion(axis, theta);
}
// cMathUtil::QuaternionToAxisAngle (MathUtil.cpp:463-481)
inline void QuaternionToAxisAngle(const DQ& q, D3& out_axis, double& out_theta) {
    out_theta = 0;
    out_axis = D3(0, 0, 1);
    DQ q1 = q;
    if (q1.w > 1) q1 = qnormalized(q1);
    double sin_theta = std::sqrt(1 - q1.w * q1.w);
    if (sin_theta > 0.000001) {
        out_theta = NormalizeAngle(2 * std::acos(q1.w));
        out_axis = D3(q1.x, q1.y, q1.z) / sin_theta;
    }
}
inline DQ QuatDiff(const DQ& q0, const DQ& q1) { return q1 * qconj(q0); }  // MathUtil.cpp:522-525
// cMathUtil::QuatTheta (MathUtil.cpp:533-549)
inline double QuatTheta(const DQ& dq) {
    double theta = 0;
    DQ q1 = dq;
    if (q1.w > 1) q1 = qnormalized(q1);
    double sin_theta = std::sqrt(1 - q1.w * q1.w);
    if (sin_theta > 0.0001) theta = NormalizeAngle(2 * std::acos(q1.w));
    return theta;
}
inline double QuatDiffTheta(const DQ& q0, const DQ& q1) { return QuatTheta(QuatDiff(q0, q1)); }
// Eigen: q * v
inline D3 QuatRotVec(const DQ& q, const D3& v) {
    D3 u(q.x, q.y, q.z);
    D3 uv = cross(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross(u, uv);
}
inline DQ StandardizeQuat(const DQ& q) { return (q.w < 0) ? DQ(-q.w, -q.x, -q.y, -q.z) : q; }
// cMathUtil::CalcQuaternionVel / VelRel (MathUtil.cpp:493-512)
inline D3 CalcQuaternionVel(const DQ& q0, const DQ& q1, double dt) {
    D3 axis; double theta;
    QuaternionToAxisAngle(QuatDiff(q0, q1), axis, theta);
    return (theta / dt) * axis;
}
inline D3 CalcQuaternionVelRel(const DQ& q0, const DQ& q1, double dt) {
    D3 axis; double theta;
    QuaternionToAxisAngle(qconj(q0) * q1, axis, theta);
    return (theta / dt) * axis;
}
// Eigen::Quaterniond::slerp (Eigen 3.3.7 Geometry/Quaternion.h), used by cKinTree::LerpPoses (KinTree.cpp:1547,1564)
inline DQ EigenSlerp(const DQ& a, double t, const DQ& b) {
    const double one = 1.0 - 2.220446049250313e-16;
    double d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
    double absD = std::fabs(d);
    double s0, s1;
    if (absD >= one) { s0 = 1.0 - t; s1 = t; }
    else {
        double theta = std::acos(absD), sinTheta = std::sin(theta);
        s0 = std::sin((1.0 - t) * theta) / sinTheta;
        s1 = std::sin(t * theta) / sinTheta;
    }
    if (d < 0) s1 = -s1;
    return {s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}
// cMathUtil::ExpMapToQuaternion (MathUtil.cpp:573-599)
inline DQ ExpMapToQuaternion(const D3& em) {
    double theta = norm(em);
    D3 axis(0, 0, 1);
    double th = 0;
    if (theta > 0.000001) { axis = em / theta; th = NormalizeAngle(theta); }
    return AxisAngleToQuaternion(axis, th);
}
// cMathUtil::CheckNextInterval (MathUtil.cpp:850-857)
inline bool CheckNextInterval(double delta, double curr_val, double int_size) {
    double pad = 0.001 * delta;
    int curr_count = static_cast<int>(std::floor((curr_val + pad) / int_size));
    int prev_count = static_cast<int>(std::floor((curr_val + pad - delta) / int_size));
    return curr_count != prev_count;
}
// cKinTree::CalcHeading (KinTree.cpp:1619-1627)
inline double CalcHeading(const DQ& rot) {
    D3 d = QuatRotVec(rot, D3(1, 0, 0));
    return std::atan2(-d.z, d.x);
}

// ------------------------------------------------------------------ float side (Bullet LinearMath) [B288-mem]
struct F3 {
    float x = 0, y = 0, z = 0;
    F3() {}
    F3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline F3 operator+(const F3& a, const F3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(const F3& a, const F3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator-(const F3& a) { return {-a.x, -a.y, -a.z}; }
inline F3 operator*(float s, const F3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline F3 operator*(const F3& a, float s) { return {s * a.x, s * a.y, s * a.z}; }
inline F3& operator+=(F3& a, const F3& b) { a = a + b; return a; }
inline float dot(const F3& a, const F3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross(const F3& a, const F3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(const F3& a) { return std::sqrt(dot(a, a)); }

struct FQ {  // btQuaternion storage order (x,y,z,w)
    float x = 0, y = 0, z = 0, w = 1;
    FQ() {}
    FQ(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
inline FQ operator*(const FQ& q1, const FQ& q2) {  // btQuaternion operator*
    return {q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y, q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z,
            q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x, q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z};
}
inline FQ inverse(const FQ& q) { return {-q.x, -q.y, -q.z, q.w}; }
inline FQ normalized(const FQ& q) {
    float n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
inline FQ fq_axis_angle(const F3& axis, float angle) {  // btQuaternion(axis, angle)
    float d = length(axis);
    float s = std::sin(angle * 0.5f) / d;
    return {axis.x * s, axis.y * s, axis.z * s, std::cos(angle * 0.5f)};
}
inline F3 quatRotate(const FQ& q, const F3& v) {  // btQuaternion.h quatRotate: (q * v) * q^-1
    FQ qv(q.w * v.x + q.y * v.z - q.z * v.y, q.w * v.y + q.z * v.x - q.x * v.z, q.w * v.z + q.x * v.y - q.y * v.x,
          -q.x * v.x - q.y * v.y - q.z * v.z);
    FQ r = qv * inverse(q);
    return {r.x, r.y, r.z};
}
struct FM3 {
    float m[3][3];
    FM3() { std::memset(m, 0, sizeof(m)); m[0][0] = m[1][1] = m[2][2] = 1; }
    static FM3 zero() { FM3 r; std::memset(r.m, 0, sizeof(r.m)); return r; }
    static FM3 diag(float a, float b, float c) { FM3 r = zero(); r.m[0][0] = a; r.m[1][1] = b; r.m[2][2] = c; return r; }
};
inline FM3 fm3_from_quat(const FQ& q) {  // btMatrix3x3::setRotation
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    FM3 r;
    r.m[0][0] = 1.0f - (yy + zz); r.m[0][1] = xy - wz; r.m[0][2] = xz + wy;
    r.m[1][0] = xy + wz; r.m[1][1] = 1.0f - (xx + zz); r.m[1][2] = yz - wx;
    r.m[2][0] = xz - wy; r.m[2][1] = yz + wx; r.m[2][2] = 1.0f - (xx + yy);
    return r;
}
inline FQ fm3_get_rotation(const FM3& a) {  // btMatrix3x3::getRotation (scalar path)
    const auto& m = a.m;
    float trace = m[0][0] + m[1][1] + m[2][2];
    float temp[4];
    if (trace > 0.0f) {
        float s = std::sqrt(trace + 1.0f);
        temp[3] = s * 0.5f;
        s = 0.5f / s;
        temp[0] = (m[2][1] - m[1][2]) * s; temp[1] = (m[0][2] - m[2][0]) * s; temp[2] = (m[1][0] - m[0][1]) * s;
    } else {
        int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        float s = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0f);
        temp[i] = s * 0.5f;
        s = 0.5f / s;
        temp[3] = (m[k][j] - m[j][k]) * s; temp[j] = (m[j][i] + m[i][j]) * s; temp[k] = (m[k][i] + m[i][k]) * s;
    }
    return {temp[0], temp[1], temp[2], temp[3]};
}
inline FM3 operator*(const FM3& a, const FM3& b) {
    FM3 r = FM3::zero();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
inline F3 operator*(const FM3& a, const F3& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline FM3 transpose(const FM3& a) { FM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
inline FM3 operator+(const FM3& a, const FM3& b) { FM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
inline FM3 operator-(const FM3& a, const FM3& b) { FM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
inline FM3 fm3_cross(const F3& a) { FM3 r = FM3::zero(); r.m[0][1] = -a.z; r.m[0][2] = a.y; r.m[1][0] = a.z; r.m[1][2] = -a.x; r.m[2][0] = -a.y; r.m[2][1] = a.x; return r; }
inline FM3 fm3_outer(const F3& a, const F3& b) { FM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a[i] * b[j]; return r; }
inline FM3 fm3_inverse(const FM3& a) {  // btMatrix3x3::inverse (cofactors / determinant)
    const auto& m = a.m;
    auto cof = [&](int r1, int c1, int r2, int c2) { return m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1]; };
    F3 co(cof(1, 1, 2, 2), cof(1, 2, 2, 0), cof(1, 0, 2, 1));
    float det = m[0][0] * co.x + m[0][1] * co.y + m[0][2] * co.z;
    float s = 1.0f / det;
    FM3 r;
    r.m[0][0] = co.x * s; r.m[0][1] = cof(0, 2, 2, 1) * s; r.m[0][2] = cof(0, 1, 1, 2) * s;
    r.m[1][0] = co.y * s; r.m[1][1] = cof(0, 0, 2, 2) * s; r.m[1][2] = cof(0, 2, 1, 0) * s;
    r.m[2][0] = co.z * s; r.m[2][1] = cof(0, 1, 2, 0) * s; r.m[2][2] = cof(0, 0, 1, 1) * s;
    return r;
}
inline FM3 fm3_absolute(const FM3& a) { FM3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = std::fabs(a.m[i][j]); return r; }

}  // namespace orc

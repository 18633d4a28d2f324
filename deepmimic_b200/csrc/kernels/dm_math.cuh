// Device-side fp32 vector / quaternion / 3x3 helpers for the sm_100a kernels.
// Quaternions are stored (x, y, z, w) on the device (the articulated-body state follows the
// reference's Bullet-side convention, R/DeepMimicCore/sim/SimBodyJoint.cpp:376,474-479);
// DeepMimic pose vectors (w, x, y, z) are converted at the C-ABI boundary.
#pragma once
#include <cuda_runtime.h>

namespace dmk {

struct V3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__host__ __device__ __forceinline__ V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__host__ __device__ __forceinline__ V3 unit3(int i) { return mk3(i == 0 ? 1.f : 0.f, i == 1 ? 1.f : 0.f, i == 2 ? 1.f : 0.f); }

struct Q4 {  // (x,y,z,w)
    float x, y, z, w;
};
__host__ __device__ __forceinline__ Q4 mkq(float x, float y, float z, float w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
__host__ __device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    return mkq(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
               a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
__host__ __device__ __forceinline__ Q4 qconj(Q4 q) { return mkq(-q.x, -q.y, -q.z, q.w); }
__host__ __device__ __forceinline__ Q4 qnormalize(Q4 q) {
    float inv = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return mkq(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}
__host__ __device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    V3 u = mk3(q.x, q.y, q.z);
    V3 t = 2.0f * cross(u, v);
    return v + q.w * t + cross(u, t);
}

struct M3 {  // row major
    float m[9];
};
// rotation matrix of a unit quaternion: rotates vectors by q
__host__ __device__ __forceinline__ M3 qmat(Q4 q) {
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs, yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    M3 r;
    r.m[0] = 1.0f - (yy + zz); r.m[1] = xy - wz; r.m[2] = xz + wy;
    r.m[3] = xy + wz; r.m[4] = 1.0f - (xx + zz); r.m[5] = yz - wx;
    r.m[6] = xz - wy; r.m[7] = yz + wx; r.m[8] = 1.0f - (xx + yy);
    return r;
}
__host__ __device__ __forceinline__ V3 mul(const M3& a, V3 v) {
    return mk3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__host__ __device__ __forceinline__ V3 mulT(const M3& a, V3 v) {
    return mk3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
__host__ __device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
__host__ __device__ __forceinline__ M3 transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[j * 3 + i];
    return r;
}
__host__ __device__ __forceinline__ M3 identity3() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.f : 0.f; return r; }
// quaternion of a rotation matrix (btMatrix3x3::getRotation form), used only for observations
__host__ __device__ __forceinline__ Q4 mat_to_quat(const M3& a) {
    float tr = a.m[0] + a.m[4] + a.m[8];
    float t[4];
    if (tr > 0.0f) {
        float s = sqrtf(tr + 1.0f);
        t[3] = s * 0.5f;
        s = 0.5f / s;
        t[0] = (a.m[7] - a.m[5]) * s; t[1] = (a.m[2] - a.m[6]) * s; t[2] = (a.m[3] - a.m[1]) * s;
    } else {
        int i = a.m[0] < a.m[4] ? (a.m[4] < a.m[8] ? 2 : 1) : (a.m[0] < a.m[8] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        float s = sqrtf(a.m[i * 3 + i] - a.m[j * 3 + j] - a.m[k * 3 + k] + 1.0f);
        t[i] = s * 0.5f;
        s = 0.5f / s;
        t[3] = (a.m[k * 3 + j] - a.m[j * 3 + k]) * s; t[j] = (a.m[j * 3 + i] + a.m[i * 3 + j]) * s; t[k] = (a.m[k * 3 + i] + a.m[i * 3 + k]) * s;
    }
    return mkq(t[0], t[1], t[2], t[3]);
}
// 6-D spatial vectors in a link frame: a = angular part, l = linear part (motion: [omega; v], force: [tau; f])
struct S6 {
    V3 a, l;
};
__host__ __device__ __forceinline__ S6 mks(V3 a, V3 l) { S6 s; s.a = a; s.l = l; return s; }
__host__ __device__ __forceinline__ S6 operator+(S6 x, S6 y) { return mks(x.a + y.a, x.l + y.l); }
__host__ __device__ __forceinline__ S6 operator*(float s, S6 x) { return mks(s * x.a, s * x.l); }
__host__ __device__ __forceinline__ float sdot(S6 m, S6 f) { return dot(m.a, f.a) + dot(m.l, f.l); }
// motion transform parent frame -> child frame: R = rotation parent->child, r = parent origin -> child origin in the child frame
__host__ __device__ __forceinline__ S6 xform_motion(const M3& R, V3 r, S6 m) {
    V3 w = mul(R, m.a);
    return mks(w, mul(R, m.l) - cross(r, w));
}
// force transform child frame -> parent frame (transpose of the above)
__host__ __device__ __forceinline__ S6 xform_force_up(const M3& R, V3 r, S6 f) { return mks(mulT(R, f.a + cross(r, f.l)), mulT(R, f.l)); }
// spatial cross products
__host__ __device__ __forceinline__ S6 cross_motion(S6 v, S6 m) { return mks(cross(v.a, m.a), cross(v.l, m.a) + cross(v.a, m.l)); }

// symmetric 6x6 spatial inertia in a link frame, blocks: tau = Iww*w + Iwv*v ; f = Iwv^T*w + Ivv*v
struct SpI {
    float ww[6];  // symmetric 3x3: xx xy xz yy yz zz
    float wv[9];  // general 3x3
    float vv[6];  // symmetric 3x3
};
__host__ __device__ __forceinline__ V3 sym_mul(const float* s, V3 v) {
    return mk3(s[0] * v.x + s[1] * v.y + s[2] * v.z, s[1] * v.x + s[3] * v.y + s[4] * v.z, s[2] * v.x + s[4] * v.y + s[5] * v.z);
}
__host__ __device__ __forceinline__ S6 spi_mul(const SpI& I, S6 m) {
    V3 tau = sym_mul(I.ww, m.a) + mk3(I.wv[0] * m.l.x + I.wv[1] * m.l.y + I.wv[2] * m.l.z, I.wv[3] * m.l.x + I.wv[4] * m.l.y + I.wv[5] * m.l.z,
                                      I.wv[6] * m.l.x + I.wv[7] * m.l.y + I.wv[8] * m.l.z);
    V3 f = mk3(I.wv[0] * m.a.x + I.wv[3] * m.a.y + I.wv[6] * m.a.z, I.wv[1] * m.a.x + I.wv[4] * m.a.y + I.wv[7] * m.a.z,
               I.wv[2] * m.a.x + I.wv[5] * m.a.y + I.wv[8] * m.a.z) + sym_mul(I.vv, m.l);
    return mks(tau, f);
}

}  // namespace dmk

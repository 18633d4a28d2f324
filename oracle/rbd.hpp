// ORACLE (test infrastructure, not product): restatement of DeepMimic's own rigid-body
// dynamics -- cSpAlg (R/DeepMimicCore/sim/SpAlg.cpp), cRBDUtil (sim/RBDUtil.cpp), cRBDModel
// (sim/RBDModel.cpp) and the cKinTree kinematics they call (anim/KinTree.cpp) -- in double
// precision, used by the Stable-PD controller and by the kinematic character.
#pragma once
#include <cassert>
#include <vector>

#include "../deepmimic_b200/csrc/host/assets.hpp"
#include "omath.hpp"

namespace orc {

using dmh::CharModel;
typedef std::vector<double> VecD;

struct SV {  // spatial vector [omega; v]  (SpAlg.cpp:86-118)
    double d[6] = {0, 0, 0, 0, 0, 0};
    D3 o() const { return D3(d[0], d[1], d[2]); }
    D3 v() const { return D3(d[3], d[4], d[5]); }
    static SV make(const D3& o, const D3& v) { SV s; s.d[0] = o.x; s.d[1] = o.y; s.d[2] = o.z; s.d[3] = v.x; s.d[4] = v.y; s.d[5] = v.z; return s; }
};
inline SV operator+(const SV& a, const SV& b) { SV r; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline SV operator*(double s, const SV& a) { SV r; for (int i = 0; i < 6; ++i) r.d[i] = s * a.d[i]; return r; }
inline double svdot(const SV& a, const SV& b) { double s = 0; for (int i = 0; i < 6; ++i) s += a.d[i] * b.d[i]; return s; }

struct SpTrans { DM3 E; D3 r; };  // Plucker transform stored as [E | r] (SpAlg.cpp:120-141)
struct SpMat { double m[6][6]; SpMat() { std::memset(m, 0, sizeof(m)); } };

inline SpTrans MatToTrans(const DT& mat) { return {mat.R, -(transpose(mat.R) * mat.t)}; }      // SpAlg.cpp:143-150
inline DT TransToMat(const SpTrans& X) { return {X.E, -(X.E * X.r)}; }                           // SpAlg.cpp:152-160
inline SpTrans InvTrans(const SpTrans& X) { return {transpose(X.E), -(X.E * X.r)}; }             // SpAlg.cpp:186-192
inline SpTrans CompTrans(const SpTrans& X0, const SpTrans& X1) { return {X0.E * X1.E, X1.r + transpose(X1.E) * X0.r}; }  // :330-338
inline SV ApplyTransM(const SpTrans& X, const SV& sv) {  // SpAlg.cpp:222-233
    D3 o0 = sv.o(), v0 = sv.v();
    return SV::make(X.E * o0, X.E * (v0 - cross(X.r, o0)));
}
inline SV ApplyTransF(const SpTrans& X, const SV& sv) {  // SpAlg.cpp:235-246
    D3 o0 = sv.o(), v0 = sv.v();
    return SV::make(X.E * (o0 - cross(X.r, v0)), X.E * v0);
}
inline SV ApplyInvTransM(const SpTrans& X, const SV& sv) {  // SpAlg.cpp:274-285
    DM3 Et = transpose(X.E);
    D3 o1 = Et * sv.o();
    return SV::make(o1, Et * sv.v() + cross(X.r, o1));
}
inline SV CrossM(const SV& a, const SV& m) {  // SpAlg.cpp:47-56
    return SV::make(cross(a.o(), m.o()), cross(a.v(), m.o()) + cross(a.o(), m.v()));
}
inline SV CrossF(const SV& a, const SV& f) {  // SpAlg.cpp:70-79
    return SV::make(cross(a.o(), f.o()) + cross(a.v(), f.v()), cross(a.o(), f.v()));
}
inline SpMat BuildSpatialMatM(const SpTrans& X) {  // SpAlg.cpp:162-172
    SpMat m; DM3 Er = X.E * crossmat(X.r);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { m.m[i][j] = X.E.m[i][j]; m.m[3 + i][3 + j] = X.E.m[i][j]; m.m[3 + i][j] = -Er.m[i][j]; }
    return m;
}
inline SpMat BuildSpatialMatF(const SpTrans& X) {  // SpAlg.cpp:174-184
    SpMat m; DM3 Er = X.E * crossmat(X.r);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { m.m[i][j] = X.E.m[i][j]; m.m[3 + i][3 + j] = X.E.m[i][j]; m.m[i][3 + j] = -Er.m[i][j]; }
    return m;
}
inline SpMat operator*(const SpMat& a, const SpMat& b) {
    SpMat r;
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) { double aik = a.m[i][k]; if (aik == 0) continue; for (int j = 0; j < 6; ++j) r.m[i][j] += aik * b.m[k][j]; }
    return r;
}
inline SV operator*(const SpMat& a, const SV& v) { SV r; for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += a.m[i][j] * v.d[j]; r.d[i] = s; } return r; }
inline SV mulT(const SpMat& a, const SV& v) { SV r; for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += a.m[j][i] * v.d[j]; r.d[i] = s; } return r; }

// ---------------------------------------------------------------- cKinTree kinematics
inline DQ pose_quat(const VecD& p, int off) { return DQ(p[off], p[off + 1], p[off + 2], p[off + 3]); }
inline D3 GetRootPos(const VecD& p) { return D3(p[0], p[1], p[2]); }
inline DQ GetRootRot(const VecD& p) { return pose_quat(p, 3); }
inline D3 GetRootVel(const VecD& v) { return D3(v[0], v[1], v[2]); }
inline D3 GetRootAngVel(const VecD& v) { return D3(v[3], v[4], v[5]); }

// cKinTree::BuildAttachTrans (KinTree.cpp:1022-1032)
inline DT BuildAttachTrans(const CharModel& cm, int j) {
    const auto& jd = cm.joints[j];
    return {RotateMatEuler(D3(jd.attach_theta.x, jd.attach_theta.y, jd.attach_theta.z)), D3(jd.attach_pt.x, jd.attach_pt.y, jd.attach_pt.z)};
}
// cKinTree::ChildParentTrans* (KinTree.cpp:1034-1069,1758-1830)
inline DT ChildParentTrans(const CharModel& cm, const VecD& pose, int j) {
    const auto& jd = cm.joints[j];
    DT A = BuildAttachTrans(cm, j);
    if (jd.parent < 0) {
        DT T; T.t = GetRootPos(pose);
        DT R; R.R = RotateMatQuat(GetRootRot(pose));
        return A * T * R;
    }
    switch (jd.type) {
        case dmh::kRevolute: { DT R; R.R = RotateMatAxis(D3(0, 0, 1), pose[jd.param_offset]); return A * R; }
        case dmh::kSpherical: { DT R; R.R = RotateMatQuat(pose_quat(pose, jd.param_offset)); return A * R; }
        case dmh::kFixed: return A;
        default: assert(false && "oracle: unsupported joint type"); return A;
    }
}
// cKinTree::JointWorldTrans (KinTree.cpp:1077-1088)
inline DT JointWorldTrans(const CharModel& cm, const VecD& pose, int j) {
    DT m;
    int c = j;
    while (c >= 0) { m = ChildParentTrans(cm, pose, c) * m; c = cm.joints[c].parent; }
    return m;
}
// cKinTree::BodyJointTrans (KinTree.cpp:1104-1114)
inline DT BodyJointTrans(const CharModel& cm, int b) {
    const auto& bd = cm.bodies[b];
    DT rot; rot.R = RotateMatEuler(D3(bd.attach_theta.x, bd.attach_theta.y, bd.attach_theta.z));
    DT tr; tr.t = D3(bd.attach_pt.x, bd.attach_pt.y, bd.attach_pt.z);
    return tr * rot;
}

// ---------------------------------------------------------------- cRBDUtil inertia (RBDUtil.cpp:615-749)
inline SpMat BuildMomentInertia(const CharModel& cm, int b) {
    const auto& bd = cm.bodies[b];
    double mass = bd.mass, x = 0, y = 0, z = 0;
    switch (bd.shape) {
        case dmh::kShapeBox: {
            double sx = bd.param[0], sy = bd.param[1], sz = bd.param[2];
            x = mass / 12.0 * (sy * sy + sz * sz); y = mass / 12.0 * (sx * sx + sz * sz); z = mass / 12.0 * (sx * sx + sy * sy);
            break;
        }
        case dmh::kShapeCapsule: {
            double r = 0.5 * bd.param[0], h = bd.param[1];
            double c_vol = M_PI * r * r * h, hs_vol = M_PI * 2.0 / 3.0 * r * r * r;
            double density = mass / (c_vol + 2 * hs_vol);
            double cmass = c_vol * density, hsm = hs_vol * density;
            x = cmass * (0.25 * r * r + (1.0 / 12.0) * h * h) + 2 * hsm * (0.4 * r * r + (3.0 / 8) * r * h + 0.25 * h * h);
            y = (0.5 * cmass + 0.8 * hsm) * r * r;
            z = x;
            break;
        }
        case dmh::kShapeSphere: { double r = 0.5 * bd.param[0]; x = y = z = 0.4 * mass * r * r; break; }
        case dmh::kShapeCylinder: { double r = 0.5 * bd.param[0], h = bd.param[1]; x = z = mass / 12 * (3 * r * r + h * h); y = mass * r * r / 2; break; }
        default: assert(false && "oracle: unsupported shape");
    }
    SpMat I;
    I.m[0][0] = x; I.m[1][1] = y; I.m[2][2] = z; I.m[3][3] = I.m[4][4] = I.m[5][5] = mass;
    return I;
}
inline SpMat BuildInertiaSpatialMat(const CharModel& cm, int b) {  // RBDUtil.cpp:742-749
    SpMat Ic = BuildMomentInertia(cm, b);
    SpTrans X = MatToTrans(BodyJointTrans(cm, b));
    return BuildSpatialMatF(X) * Ic * BuildSpatialMatM(InvTrans(X));
}

// ---------------------------------------------------------------- cRBDModel (RBDModel.cpp)
struct RBDModel {
    const CharModel* cm = nullptr;
    D3 gravity;
    VecD pose, vel;
    int ndof = 0, nj = 0;
    std::vector<std::vector<SV>> S;      // joint subspace columns (6 x param_size)
    std::vector<DT> child_parent;        // 4x4 child->parent
    std::vector<SpTrans> world_joint;    // world -> joint
    std::vector<SpMat> Ij;               // constant joint-frame inertias
    std::vector<double> M;               // ndof x ndof mass matrix (row major)
    VecD C;                              // bias force

    void Init(const CharModel& c, const D3& g) {
        cm = &c; gravity = g; nj = c.num_joints(); ndof = c.pose_dim;
        S.assign(nj, {}); child_parent.resize(nj); world_joint.resize(nj); Ij.resize(nj);
        for (int j = 0; j < nj; ++j) { S[j].assign(c.joints[j].param_size, SV()); if (c.bodies[j].shape != dmh::kShapeNull) Ij[j] = BuildInertiaSpatialMat(c, j); }
        M.assign(static_cast<size_t>(ndof) * ndof, 0.0); C.assign(ndof, 0.0);
    }
    SpTrans SpChildParent(int j) const { return MatToTrans(child_parent[j]); }
    SpTrans SpParentChild(int j) const { return MatToTrans(inv_rigid(child_parent[j])); }

    // cRBDUtil::BuildJointSubspace* (RBDUtil.cpp:798-887)
    void UpdateJointSubspace() {
        for (int j = 0; j < nj; ++j) {
            const auto& jd = cm->joints[j];
            for (auto& s : S[j]) s = SV();
            if (jd.parent < 0) {
                DM3 E = RotateMatQuat(GetRootRot(pose));
                for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { S[j][c].d[3 + r] = E.m[c][r]; S[j][3 + c].d[r] = E.m[c][r]; }  // E^T blocks
            } else if (jd.type == dmh::kRevolute) S[j][0].d[2] = 1;
            else if (jd.type == dmh::kSpherical) { S[j][0].d[0] = 1; S[j][1].d[1] = 1; S[j][2].d[2] = 1; }
        }
    }
    // cRBDUtil::BuildCjRoot (RBDUtil.cpp:915-958)
    SV BuildCjRoot() const {
        DQ q = GetRootRot(pose);
        D3 vl = GetRootVel(vel), va = GetRootAngVel(vel);
        // dq = BuildQuaternionDiffMat(q) * va  (MathUtil.cpp:483-491)
        DQ dq(-0.5 * q.x * va.x - 0.5 * q.y * va.y - 0.5 * q.z * va.z, 0.5 * q.w * va.x - 0.5 * q.z * va.y + 0.5 * q.y * va.z,
              0.5 * q.z * va.x + 0.5 * q.w * va.y - 0.5 * q.x * va.z, -0.5 * q.y * va.x + 0.5 * q.x * va.y + 0.5 * q.w * va.z);
        DM3 m;
        m.m[0][0] = 4 * (q.w * dq.w + q.x * dq.x); m.m[1][1] = 4 * (q.w * dq.w + q.y * dq.y); m.m[2][2] = 4 * (q.w * dq.w + q.z * dq.z);
        m.m[1][0] = 2 * (dq.x * q.y + q.x * dq.y - dq.w * q.z - q.w * dq.z);
        m.m[0][1] = 2 * (dq.x * q.y + q.x * dq.y + dq.w * q.z + q.w * dq.z);
        m.m[2][0] = 2 * (dq.x * q.z + q.x * dq.z + dq.w * q.y + q.w * dq.y);
        m.m[0][2] = 2 * (dq.x * q.z + q.x * dq.z - dq.w * q.y - q.w * dq.y);
        m.m[2][1] = 2 * (dq.y * q.z + q.y * dq.z - dq.w * q.x - q.w * dq.x);
        m.m[1][2] = 2 * (dq.y * q.z + q.y * dq.z + dq.w * q.x + q.w * dq.x);
        return SV::make(D3(0, 0, 0), m * vl);
    }
    SV Sq(int j, const VecD& x) const {
        SV r; const auto& jd = cm->joints[j];
        for (int c = 0; c < jd.param_size; ++c) r = r + x[jd.param_offset + c] * S[j][c];
        return r;
    }
    // cRBDModel::Update (RBDModel.cpp:36-46)
    void Update(const VecD& p, const VecD& v) {
        pose = p; vel = v;
        UpdateJointSubspace();
        for (int j = 0; j < nj; ++j) child_parent[j] = ChildParentTrans(*cm, pose, j);
        for (int j = 0; j < nj; ++j) {  // cRBDUtil::CalcWorldJointTransforms (RBDUtil.cpp:751-775)
            int par = cm->joints[j].parent;
            SpTrans wp; if (par >= 0) wp = world_joint[par];
            world_joint[j] = CompTrans(SpParentChild(j), wp);
        }
        BuildMassMat();
        VecD acc(ndof, 0.0);
        SolveInvDyna(acc, C);
    }
    // cRBDUtil::BuildMassMat -- composite rigid body algorithm (RBDUtil.cpp:123-195)
    void BuildMassMat() {
        std::fill(M.begin(), M.end(), 0.0);
        std::vector<SpMat> Is = Ij, cpF(nj), pcM(nj);
        for (int j = 0; j < nj; ++j) { SpTrans X = SpChildParent(j); cpF[j] = BuildSpatialMatF(X); pcM[j] = BuildSpatialMatM(InvTrans(X)); }
        for (int j = nj - 1; j >= 0; --j) {
            if (cm->bodies[j].shape == dmh::kShapeNull) continue;
            const auto& jd = cm->joints[j];
            if (jd.parent >= 0) {
                SpMat t = cpF[j] * Is[j] * pcM[j];
                for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) Is[jd.parent].m[a][b] += t.m[a][b];
            }
            int dim = jd.param_size;
            if (dim == 0) continue;
            std::vector<SV> F(dim);
            for (int c = 0; c < dim; ++c) F[c] = Is[j] * S[j][c];
            for (int a = 0; a < dim; ++a) for (int b = 0; b < dim; ++b) M[static_cast<size_t>(jd.param_offset + a) * ndof + jd.param_offset + b] = svdot(S[j][a], F[b]);
            int cur = j;
            while (cm->joints[cur].parent >= 0) {
                for (int c = 0; c < dim; ++c) F[c] = cpF[cur] * F[c];
                cur = cm->joints[cur].parent;
                const auto& cd = cm->joints[cur];
                for (int a = 0; a < dim; ++a) for (int b = 0; b < cd.param_size; ++b) {
                    double h = svdot(F[a], S[cur][b]);
                    M[static_cast<size_t>(jd.param_offset + a) * ndof + cd.param_offset + b] = h;
                    M[static_cast<size_t>(cd.param_offset + b) * ndof + jd.param_offset + a] = h;
                }
            }
        }
    }
    // cRBDUtil::SolveInvDyna -- RNEA (RBDUtil.cpp:4-97)
    void SolveInvDyna(const VecD& acc, VecD& out_tau) const {
        SV vel0, acc0 = SV::make(D3(0, 0, 0), -gravity);
        std::vector<SV> vels(nj), accs(nj), fs(nj);
        for (int j = 0; j < nj; ++j) {
            if (cm->bodies[j].shape == dmh::kShapeNull) continue;
            const auto& jd = cm->joints[j];
            SpTrans pc = SpParentChild(j);
            SV cj = (jd.parent < 0) ? BuildCjRoot() : SV();
            SV vj = Sq(j, vel), Sddq = Sq(j, acc);
            SV vel_p = (jd.parent >= 0) ? vels[jd.parent] : vel0;
            SV acc_p = (jd.parent >= 0) ? accs[jd.parent] : acc0;
            SV cv = ApplyTransM(pc, vel_p) + vj;
            SV ca = ApplyTransM(pc, acc_p) + Sddq + cj + CrossM(cv, vj);
            fs[j] = Ij[j] * ca + CrossF(cv, Ij[j] * cv);
            vels[j] = cv; accs[j] = ca;
        }
        out_tau.assign(ndof, 0.0);
        for (int j = nj - 1; j >= 0; --j) {
            if (cm->bodies[j].shape == dmh::kShapeNull) continue;
            const auto& jd = cm->joints[j];
            for (int c = 0; c < jd.param_size; ++c) out_tau[jd.param_offset + c] = svdot(S[j][c], fs[j]);
            if (jd.parent >= 0) fs[jd.parent] = fs[jd.parent] + ApplyTransF(SpChildParent(j), fs[j]);
        }
    }
};

// Symmetric solve standing in for Eigen's M.ldlt().solve(b) (ImpPDController.cpp:188).  The matrix
// has exactly-zero rows/cols (root quaternion-w slot); Eigen's pivoted LDLT returns 0 for those
// components, so they are dropped and the remaining SPD block is solved by Cholesky.
inline VecD SolveSymmetric(const std::vector<double>& A, int n, const VecD& b) {
    std::vector<int> live;
    for (int i = 0; i < n; ++i) if (A[static_cast<size_t>(i) * n + i] != 0.0) live.push_back(i);
    int m = static_cast<int>(live.size());
    std::vector<double> L(static_cast<size_t>(m) * m, 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) {
        double s = A[static_cast<size_t>(live[i]) * n + live[j]];
        for (int k = 0; k < j; ++k) s -= L[static_cast<size_t>(i) * m + k] * L[static_cast<size_t>(j) * m + k];
        if (i == j) L[static_cast<size_t>(i) * m + i] = std::sqrt(s); else L[static_cast<size_t>(i) * m + j] = s / L[static_cast<size_t>(j) * m + j];
    }
    VecD y(m), x(n, 0.0);
    for (int i = 0; i < m; ++i) { double s = b[live[i]]; for (int k = 0; k < i; ++k) s -= L[static_cast<size_t>(i) * m + k] * y[k]; y[i] = s / L[static_cast<size_t>(i) * m + i]; }
    for (int i = m - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < m; ++k) s -= L[static_cast<size_t>(k) * m + i] * x[live[k]]; x[live[i]] = s / L[static_cast<size_t>(i) * m + i]; }
    return x;
}

// ---------------------------------------------------------------- cKinTree pose algebra used by SPD, the mocap clip and the reward
// cKinTree::VelToPoseDiff (KinTree.cpp:1581-1610)
inline void VelToPoseDiff(const CharModel& cm, const VecD& pose, const VecD& vel, VecD& out) {
    out = vel;
    auto qdiff = [](const DQ& q, const D3& w, double* o) {
        o[0] = -0.5 * q.x * w.x - 0.5 * q.y * w.y - 0.5 * q.z * w.z;
        o[1] = 0.5 * q.w * w.x - 0.5 * q.z * w.y + 0.5 * q.y * w.z;
        o[2] = 0.5 * q.z * w.x + 0.5 * q.w * w.y - 0.5 * q.x * w.z;
        o[3] = -0.5 * q.y * w.x + 0.5 * q.x * w.y + 0.5 * q.w * w.z;
    };
    qdiff(GetRootRot(pose), GetRootAngVel(vel), &out[3]);
    for (int j = 1; j < cm.num_joints(); ++j) {
        const auto& jd = cm.joints[j];
        if (jd.type == dmh::kSpherical) qdiff(pose_quat(pose, jd.param_offset), D3(vel[jd.param_offset], vel[jd.param_offset + 1], vel[jd.param_offset + 2]), &out[jd.param_offset]);
    }
}
// cKinTree::PostProcessPose (KinTree.cpp:1318-1334)
inline void PostProcessPose(const CharModel& cm, VecD& pose) {
    auto n4 = [](double* q) { double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int k = 0; k < 4; ++k) q[k] /= n; };
    n4(&pose[3]);
    for (int j = 1; j < cm.num_joints(); ++j) if (cm.joints[j].type == dmh::kSpherical) n4(&pose[cm.joints[j].param_offset]);
}
// cKinTree::CalcVel (KinTree.cpp:1281-1316)
inline void CalcVel(const CharModel& cm, const VecD& p0, const VecD& p1, double dt, VecD& out) {
    out.assign(p0.size(), 0.0);
    D3 rv = (GetRootPos(p1) - GetRootPos(p0)) / dt;
    D3 rw = CalcQuaternionVel(GetRootRot(p0), GetRootRot(p1), dt);
    out[0] = rv.x; out[1] = rv.y; out[2] = rv.z; out[3] = rw.x; out[4] = rw.y; out[5] = rw.z; out[6] = 0;
    for (int j = 1; j < cm.num_joints(); ++j) {
        const auto& jd = cm.joints[j];
        if (jd.type == dmh::kSpherical) {
            D3 w = CalcQuaternionVelRel(pose_quat(p0, jd.param_offset), pose_quat(p1, jd.param_offset), dt);
            out[jd.param_offset] = w.x; out[jd.param_offset + 1] = w.y; out[jd.param_offset + 2] = w.z; out[jd.param_offset + 3] = 0;
        } else {
            for (int k = 0; k < jd.param_size; ++k) out[jd.param_offset + k] = (p1[jd.param_offset + k] - p0[jd.param_offset + k]) / dt;
        }
    }
}
// cKinTree::LerpPoses (KinTree.cpp:1336-1378)
inline void LerpPoses(const CharModel& cm, const double* p0, const double* p1, double lerp, VecD& out) {
    out.assign(cm.pose_dim, 0.0);
    for (int k = 0; k < 3; ++k) out[k] = (1 - lerp) * p0[k] + lerp * p1[k];
    DQ r = qnormalized(EigenSlerp(DQ(p0[3], p0[4], p0[5], p0[6]), lerp, DQ(p1[3], p1[4], p1[5], p1[6])));
    out[3] = r.w; out[4] = r.x; out[5] = r.y; out[6] = r.z;
    for (int j = 1; j < cm.num_joints(); ++j) {
        const auto& jd = cm.joints[j];
        int o = jd.param_offset;
        if (jd.type == dmh::kSpherical) {
            DQ q = EigenSlerp(DQ(p0[o], p0[o + 1], p0[o + 2], p0[o + 3]), lerp, DQ(p1[o], p1[o + 1], p1[o + 2], p1[o + 3]));
            out[o] = q.w; out[o + 1] = q.x; out[o + 2] = q.y; out[o + 3] = q.z;
        } else {
            for (int k = 0; k < jd.param_size; ++k) out[o + k] = (1 - lerp) * p0[o + k] + lerp * p1[o + k];
        }
    }
}

// cRBDUtil::BuildEndEffectorJacobian + CalcWorldVel (RBDUtil.cpp:217-243,412-420): world-frame spatial velocity of joint j
inline SV CalcWorldVel(const CharModel& cm, const VecD& pose, const VecD& vel, int joint_id) {
    RBDModel tmp; tmp.cm = &cm; tmp.nj = cm.num_joints(); tmp.ndof = cm.pose_dim; tmp.pose = pose; tmp.vel = vel;
    tmp.S.assign(tmp.nj, {});
    for (int j = 0; j < tmp.nj; ++j) tmp.S[j].assign(cm.joints[j].param_size, SV());
    tmp.UpdateJointSubspace();
    SV total;
    int cur = joint_id;
    SpTrans curr_trans;  // identity
    std::vector<std::pair<int, SpTrans>> chain;
    while (cur >= 0) {
        chain.emplace_back(cur, curr_trans);
        SpTrans pc = MatToTrans(inv_rigid(ChildParentTrans(cm, pose, cur)));
        curr_trans = CompTrans(curr_trans, pc);
        cur = cm.joints[cur].parent;
    }
    // J block = ApplyInvTransM(world->joint_id, ApplyTransM(curr_trans_at_c, S_c));  sv = J * vel
    for (auto& ce : chain) {
        SV sj = tmp.Sq(ce.first, vel);
        total = total + ApplyInvTransM(curr_trans, ApplyTransM(ce.second, sj));
    }
    return total;
}
// cRBDUtil::CalcCoM (RBDUtil.cpp:572-613) for the kinematic character
inline void CalcCoM(const CharModel& cm, const VecD& pose, const VecD& vel, D3& out_com, D3& out_vel) {
    out_com = D3(); out_vel = D3();
    double total = 0;
    for (int j = 0; j < cm.num_joints(); ++j) {
        if (cm.bodies[j].shape == dmh::kShapeNull) continue;
        DT body_world = JointWorldTrans(cm, pose, j) * BodyJointTrans(cm, j);
        D3 world_com = body_world.t;
        SpTrans com_trans; com_trans.r = world_com;
        SV sv = ApplyTransM(com_trans, CalcWorldVel(cm, pose, vel, j));
        double m = cm.bodies[j].mass;
        out_com += m * world_com; out_vel += m * sv.v(); total += m;
    }
    out_com = out_com / total; out_vel = out_vel / total;
}

}  // namespace orc

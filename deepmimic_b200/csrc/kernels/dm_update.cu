// Fused per-update kernel: one launch advances every environment by n_updates x Update(dt), i.e. per
// update exactly what the reference does in cSceneSimChar::Update (R/DeepMimicCore/scenes/SceneSimChar.cpp:136-161):
//   kin clock / cycle sync  (scenes/SceneImitate.cpp:306-318,420-444)
//   Stable-PD torques       (sim/ImpPDController.cpp:136-195)  -- mass matrix by composite-rigid-body recursion,
//                           bias by recursive Newton-Euler, solve by Featherstone's sparse L^T D L factorisation
//   2 x Bullet sub-step     (sim/World.cpp:93-104): link-vs-plane manifolds, unconstrained acceleration,
//                           10 projected-Gauss-Seidel sweeps over contact / friction / joint-limit rows, integration
//   controller clock + 30 Hz "need action" edge (sim/CtController.cpp:221-227), fall / explode / timer flags.
//
// B200 mapping: one tile of W lanes (16 or 32) per environment, lane = link of the articulated tree.  Tree
// recursions run level-synchronously with warp shuffles parent<->child; the per-env mass matrix is stored
// chain-sparse (row i keeps only its ancestor dofs: no fill-in across branches) in shared memory; a contact
// row's Jacobian, its M^-1 image and the PGS dot/axpy all live on the <=24-entry dof chain of the contact
// link, one lane per chain entry.  State is read once per launch with float4 loads from env-major blocks and
// stays in registers / shared memory across the n_updates loop.  No tensor cores: there is no dense
// contraction here (34 or 70 dofs, tree-sparse).
#include <cooperative_groups.h>

#include "dm_model.cuh"

namespace dmk {

namespace {

template <int W>
struct Tile {
    static __device__ __forceinline__ float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src, W); }
    static __device__ __forceinline__ int shfli(int v, int src) { return __shfl_sync(0xffffffffu, v, src, W); }
    static __device__ __forceinline__ float sum(float v) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, W);
        return v;
    }
    static __device__ __forceinline__ int maxi(int v) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o, W));
        return v;
    }
    static __device__ __forceinline__ V3 shfl3(V3 v, int src) { return mk3(shfl(v.x, src), shfl(v.y, src), shfl(v.z, src)); }
    static __device__ __forceinline__ S6 shfl6(S6 v, int src) { return mks(shfl3(v.a, src), shfl3(v.l, src)); }
};
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// rigid-body (composite) inertia about the link origin: tau = Io*w + h x v ; f = m*v - h x w
struct Rig {
    float m;
    V3 h;
    float io[6];  // xx xy xz yy yz zz
};
__device__ __forceinline__ S6 rig_mul(const Rig& I, S6 s) { return mks(sym_mul(I.io, s.a) + cross(I.h, s.l), I.m * s.l - cross(I.h, s.a)); }
// express a composite given in a child frame (origin = child COM) in its parent frame (R: parent->child, r: parent origin -> child origin, child coords)
__device__ __forceinline__ Rig rig_to_parent(const Rig& c, const M3& R, V3 r) {
    Rig p;
    p.m = c.m;
    V3 hs = c.h + c.m * r;
    p.h = mulT(R, hs);
    float rr = dot(r, r), rh = dot(r, c.h);
    float d = c.m * rr + 2.0f * rh;
    float a[9];
    a[0] = c.io[0] + d - c.m * r.x * r.x - 2.0f * r.x * c.h.x;
    a[4] = c.io[3] + d - c.m * r.y * r.y - 2.0f * r.y * c.h.y;
    a[8] = c.io[5] + d - c.m * r.z * r.z - 2.0f * r.z * c.h.z;
    a[1] = a[3] = c.io[1] - c.m * r.x * r.y - r.x * c.h.y - c.h.x * r.y;
    a[2] = a[6] = c.io[2] - c.m * r.x * r.z - r.x * c.h.z - c.h.x * r.z;
    a[5] = a[7] = c.io[4] - c.m * r.y * r.z - r.y * c.h.z - c.h.y * r.z;
    // Rt * A * R
    M3 A; for (int i = 0; i < 9; ++i) A.m[i] = a[i];
    M3 T = mul(transpose(R), mul(A, R));
    p.io[0] = T.m[0]; p.io[1] = T.m[1]; p.io[2] = T.m[2]; p.io[3] = T.m[4]; p.io[4] = T.m[5]; p.io[5] = T.m[8];
    return p;
}

__device__ __forceinline__ float normalize_angle(float t) {  // cMathUtil::NormalizeAngle
    float n = fmodf(t, 6.283185307179586f);
    if (n > 3.14159265358979f) n -= 6.283185307179586f;
    else if (n < -3.14159265358979f) n += 6.283185307179586f;
    return n;
}
// rotation vector of a unit quaternion, same semantics as cMathUtil::QuaternionToAxisAngle (theta in [-pi,pi], zero when
// sin(theta/2) <= 1e-6) but evaluated with atan2 so small angles keep fp32 accuracy
__device__ __forceinline__ V3 quat_rotvec(Q4 q) {
    float s = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    if (!(s > 0.000001f)) return mk3(0.f, 0.f, 0.f);
    float th = normalize_angle(2.0f * atan2f(s, q.w));
    float k = th / s;
    return mk3(q.x * k, q.y * k, q.z * k);
}
// btMultiBody::stepPositionsMultiDof's exponential-map quaternion update
__device__ __forceinline__ Q4 quat_integrate(V3 omega, Q4 quat, bool base_body, float dt) {
    V3 angvel = base_body ? omega : qrot(quat, omega);
    float fAngle = sqrtf(dot(angvel, angvel));
    const float kThr = 0.5f * 1.57079632679489661923f;
    if (fAngle * dt > kThr) fAngle = kThr / dt;
    V3 axis;
    if (fAngle < 0.001f) axis = angvel * (0.5f * dt - (dt * dt * dt) * 0.020833333333f * fAngle * fAngle);
    else axis = angvel * (__sinf(0.5f * fAngle * dt) / fAngle);   // |angle| <= pi/8 (ANGULAR_MOTION_THRESHOLD): fast path error ~1e-7
    float cw = __cosf(fAngle * dt * 0.5f);
    Q4 r = base_body ? qmul(quat, mkq(-axis.x, -axis.y, -axis.z, cw)) : qmul(mkq(axis.x, axis.y, axis.z, cw), quat);
    float n = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    return mkq(r.x / n, r.y / n, r.z / n, r.w / n);
}

// per-block copy of the model fields that the chain walks index by a data-dependent link / dof (everything else is read once into registers)
struct HotModel {
    const int* meta;          // nl : parent(8) | jtype(8) | ndof(8) | depth0(8), signed bytes
    const int* meta2;         // nl : dof0(8) | last_depth(8)
    const float* axis;        // nl x 3
    const float* dvec;        // nl x 3
    const unsigned char* chain;   // nl x cs
    const unsigned char* dof_link;  // n
    const unsigned char* dof_depth; // n
};
__host__ __device__ inline int hot_model_bytes(int nl, int n, int cs) { return ((nl * 8 + nl * 24 + nl * cs + 2 * n + 15) / 16) * 16; }

struct Smem {   // per-environment shared-memory carve-up (floats)
    float* Rs;      // nl x 9   parent->link rotation
    float* rs;      // nl x 3   parent origin -> link origin, link frame
    float* Rw;      // nl x 9   world->link rotation
    float* pw;      // nl x 3   link origin, world
    float* H;       // n x cs   chain-sparse mass matrix, overwritten by its L^T D L factors
    float* vel;     // n        generalised velocity [omega_w, v_w, joint rates]
    float* tau;     // n        generalised applied force (joint torques; base 0)
    float* bias;    // n        bias force / scratch rhs
    float* z;       // n        PGS accumulator in factor space
    float* dinv;    // n        1 / D_k
    float* Y;       // maxrows x cs
    float* rrhs;    // maxrows
    float* rinv;    // maxrows  1 / (J M^-1 J^T)
    float* rlam;    // maxrows
    int* rlink;     // maxrows  link whose dof chain the row lives on
    float* ppos;    // maxpts x 4 : world position of the point on the link (xyz), distance (w)
    float* pimp;    // maxpts : warm-start normal impulse
    int* pref;      // maxpts : link * 4 + slot
    float* Rwb;     // 9 : world->base rotation
};
__host__ __device__ inline int smem_floats_per_env(int nl, int n, int cs, int maxrows) {
    int maxpts = maxrows / 3;
    int f = nl * 24 + n * cs + 5 * n + maxrows * cs + 3 * maxrows + maxrows + maxpts * 6 + 12;   // + 9 floats world->base rotation (padded)
    return ((f + 15) / 32) * 32 + 16;   // stride == 16 (mod 32 banks): the two environments of a warp (W = 16) hit disjoint bank halves
}

}  // namespace


template <int W, bool DEBUG>
__global__ void __launch_bounds__(kUpdateMaxThreads, 1) dm_update_kernel(const DevModel* __restrict__ gm, DevState st, const double* __restrict__ frame_times,
                                                                 const float* __restrict__ frames, double dt, int n_updates, int sim_substeps, int maxrows, int sync_every_stage) {
    using T = Tile<W>;
    extern __shared__ __align__(16) float smem_raw[];
    const int tiles = blockDim.x / W;
    const int BLOCK = blockDim.x;
    const int tile = threadIdx.x / W;
    const int lane = threadIdx.x % W;
    const int env = blockIdx.x * tiles + tile;   // host guarantees num_envs (padded) is a multiple of tiles
    const DevModel& M = *gm;
    const int nl = M.nl, n = M.n, cs = M.cs, maxlevel = M.maxlevel;
    const int maxpts = maxrows / 3;
    const bool act = lane < nl;
    const int li = act ? lane : nl - 1;
    const DevLink& L = M.link[li];

    // ---- shared memory: [hot model tables | per-env blocks]
    HotModel HM;
    {
        unsigned char* hb = reinterpret_cast<unsigned char*>(smem_raw);
        int* meta = reinterpret_cast<int*>(hb); int* meta2 = meta + nl;
        float* ax = reinterpret_cast<float*>(meta2 + nl); float* dv = ax + nl * 3;
        unsigned char* ch = reinterpret_cast<unsigned char*>(dv + nl * 3); unsigned char* dl = ch + nl * cs; unsigned char* dd = dl + n;
        for (int j = threadIdx.x; j < nl; j += BLOCK) {
            const DevLink& K = M.link[j];
            meta[j] = (K.parent & 0xff) | ((K.jtype & 0xff) << 8) | ((K.ndof & 0xff) << 16) | ((K.depth0 & 0xff) << 24);
            meta2[j] = (K.dof0 & 0xff) | ((K.last_depth & 0xff) << 8);
            for (int k = 0; k < 3; ++k) { ax[j * 3 + k] = K.axis[k]; dv[j * 3 + k] = K.dvec[k]; }
            for (int d = 0; d < cs; ++d) ch[j * cs + d] = M.chain_dof[j][d];
        }
        for (int k = threadIdx.x; k < n; k += BLOCK) { dl[k] = M.dof_link[k]; dd[k] = M.dof_depth[k]; }
        HM.meta = meta; HM.meta2 = meta2; HM.axis = ax; HM.dvec = dv; HM.chain = ch; HM.dof_link = dl; HM.dof_depth = dd;
    }
    __syncthreads();
    auto lk_parent = [&](int j) { return static_cast<int>(static_cast<signed char>(HM.meta[j] & 0xff)); };
    auto lk_jtype = [&](int j) { return (HM.meta[j] >> 8) & 0xff; };
    auto lk_ndof = [&](int j) { return (HM.meta[j] >> 16) & 0xff; };
    auto lk_depth0 = [&](int j) { return (HM.meta[j] >> 24) & 0xff; };
    auto lk_dof0 = [&](int j) { return HM.meta2[j] & 0xff; };
    auto lk_lastd = [&](int j) { return (HM.meta2[j] >> 8) & 0xff; };

    Smem S;
    {
        float* p = smem_raw + hot_model_bytes(nl, n, cs) / 4 + static_cast<size_t>(tile) * smem_floats_per_env(nl, n, cs, maxrows);
        S.Rs = p; p += nl * 9; S.rs = p; p += nl * 3; S.Rw = p; p += nl * 9; S.pw = p; p += nl * 3;
        S.H = p; p += n * cs; S.vel = p; p += n; S.tau = p; p += n; S.bias = p; p += n; S.z = p; p += n; S.dinv = p; p += n;
        S.Y = p; p += maxrows * cs; S.rrhs = p; p += maxrows; S.rinv = p; p += maxrows; S.rlam = p; p += maxrows;
        S.rlink = reinterpret_cast<int*>(p); p += maxrows;
        S.ppos = p; p += maxpts * 4; S.pimp = p; p += maxpts; S.pref = reinterpret_cast<int*>(p); p += maxpts;
        S.Rwb = p; p += 12;
    }

    // ---- per-lane model constants
    const int parent = L.parent, jtype = L.jtype, ndof = act ? L.ndof : 0, dof0 = L.dof0, level = act ? L.level : 1000, depth0 = L.depth0;
    const V3 dvec = mk3(L.dvec[0], L.dvec[1], L.dvec[2]), evec = mk3(L.evec[0], L.evec[1], L.evec[2]);
    const Q4 zrot = mkq(L.zrot[0], L.zrot[1], L.zrot[2], L.zrot[3]);
    const V3 axis = mk3(L.axis[0], L.axis[1], L.axis[2]);
    const float mass = L.mass;
    const int plane = parent >= 0 ? parent : 0;
    const int nchild = act ? L.nchild : 0;
    const int child_pack = (L.child[0] & 0xff) | ((L.child[1] & 0xff) << 8) | ((L.child[2] & 0xff) << 16) | ((L.child[3] & 0xff) << 24);

    // ---- state load (env-major block, float4)
    const int ss = sim_stride(nl);
    float* sim = st.sim + static_cast<size_t>(env) * ss;
    double* tm = st.time + static_cast<size_t>(env) * kTimeDoubles;
    int* fl = st.flags + static_cast<size_t>(env) * kFlagInts;
    float* mani = st.manifold + static_cast<size_t>(env) * nl * kManifoldFloats;
    V3 basePos; Q4 baseQuat; V3 baseOmega, baseVel;
    {
        float4 b0 = reinterpret_cast<const float4*>(sim)[0], b1 = reinterpret_cast<const float4*>(sim)[1], b2 = reinterpret_cast<const float4*>(sim)[2],
               b3 = reinterpret_cast<const float4*>(sim)[3];
        basePos = mk3(b0.x, b0.y, b0.z); baseQuat = mkq(b1.x, b1.y, b1.z, b1.w); baseOmega = mk3(b2.x, b2.y, b2.z); baseVel = mk3(b3.x, b3.y, b3.z);
    }
    float4 jp = reinterpret_cast<const float4*>(sim + 16)[li];
    float4 jv = reinterpret_cast<const float4*>(sim + 16 + 4 * nl)[li];
    const float4 tg = reinterpret_cast<const float4*>(sim + 16 + 8 * nl)[li];
    double kin_time = tm[kTKin], ctrl_time = tm[kTCtrl], prev_act = tm[kTPrevAct], timer = tm[kTTimer];
    const double init_off = tm[kTInitOff], timer_max = tm[kTTimerMax];
    double org_x = tm[kTOrigin], org_y = tm[kTOrigin + 1], org_z = tm[kTOrigin + 2];
    int need_action = fl[kFNeedAction];
    bool alive = fl[kFDone] == 0;
    int f_over = fl[kFRowOverflow], f_updates = fl[kFUpdates];

    const float h = static_cast<float>(dt) / static_cast<float>(sim_substeps);
    const V3 grav = mk3(M.gravity[0], M.gravity[1], M.gravity[2]);
    float* dbg = (DEBUG && st.pdbg) ? st.pdbg + static_cast<size_t>(env) * kDebugFloats : nullptr;   // test hook: stage dumps of the first update

    // registers describing the current configuration (the world-frame copies live in shared memory)
    M3 R;               // parent->link
    V3 r;               // parent origin->link origin (link frame)
    S6 vel6;            // link spatial velocity (link frame)
    bool in_contact_tol = false;

    auto own_Rw = [&]() { M3 m; for (int k = 0; k < 9; ++k) m.m[k] = S.Rw[li * 9 + k]; return m; };
    auto own_pos = [&]() { return mk3(S.pw[li * 3], S.pw[li * 3 + 1], S.pw[li * 3 + 2]); };
    auto get_Rwb = [&]() { M3 m; for (int k = 0; k < 9; ++k) m.m[k] = S.Rwb[k]; return m; };
    auto joint_twist = [&]() {
        S6 vJ = mks(mk3(0, 0, 0), mk3(0, 0, 0));
        if (jtype == kJSpherical) { V3 w = mk3(jv.x, jv.y, jv.z); vJ = mks(w, cross(w, dvec)); }
        else if (jtype == kJRevolute) { V3 w = jv.x * axis; vJ = mks(w, cross(w, dvec)); }
        return vJ;
    };
    // joint motion subspace column d of link `lk` (in its own frame), from the shared-memory tables
    auto subspace = [&](int lk, int d) -> S6 {
        V3 top = (lk_jtype(lk) == kJSpherical) ? unit3(d) : mk3(HM.axis[lk * 3], HM.axis[lk * 3 + 1], HM.axis[lk * 3 + 2]);
        return mks(top, cross(top, mk3(HM.dvec[lk * 3], HM.dvec[lk * 3 + 1], HM.dvec[lk * 3 + 2])));
    };

    bool need_kin = true, pending_flags = false;
    // block-wide synchronisation policy: 1 = every stage, 0 = once per update, k < 0 = every -k updates, -1000 = never
    const int sync_period = sync_every_stage > 0 ? 1 : (sync_every_stage == 0 ? (sim_substeps + 1) : (sync_every_stage <= -1000 ? 0 : -sync_every_stage * (sim_substeps + 1)));
    const int stages_per_upd = sim_substeps + 1;
    const int total_stages = n_updates * stages_per_upd;
    #pragma unroll 1
    for (int stage = 0; stage <= total_stages; ++stage) {
        // =================================================================== configuration-dependent quantities
        if (need_kin) {
            need_kin = false;
            Q4 cached;
            if (jtype == kJSpherical) cached = qmul(mkq(jp.x, jp.y, jp.z, -jp.w), zrot);
            else if (jtype == kJRevolute) {
                float s, c;
                __sincosf(-0.5f * jp.x, &s, &c);   // |angle| <= pi/2 + limit overshoot: fast path is accurate to ~1 ulp of the result scale
                cached = qmul(mkq(axis.x * s, axis.y * s, axis.z * s, c), zrot);
            } else cached = zrot;
            R = qmat(cached);
            r = dvec + mul(R, evec);
            const M3 Rwb = qmat(baseQuat);
            M3 Rwl; V3 pos;
            if (lane == 0) { Rwl = mul(R, Rwb); pos = basePos + mulT(Rwl, r); }
            #pragma unroll 1
            for (int lv = 1; lv <= maxlevel; ++lv) {   // world transforms, level-synchronous
                M3 pR; V3 pp;
#pragma unroll
                for (int k = 0; k < 9; ++k) pR.m[k] = T::shfl(Rwl.m[k], plane);
                pp = T::shfl3(pos, plane);
                if (level == lv) { Rwl = mul(R, pR); pos = pp + mulT(Rwl, r); }
            }
            const S6 vb = mks(mul(Rwb, baseOmega), mul(Rwb, baseVel));
            const S6 vJ = joint_twist();
            if (lane == 0) vel6 = xform_motion(R, r, vb);
            #pragma unroll 1
            for (int lv = 1; lv <= maxlevel; ++lv) {   // link velocities
                S6 pv = T::shfl6(vel6, plane);
                if (level == lv) vel6 = xform_motion(R, r, pv) + vJ;
            }
            if (act) {
#pragma unroll
                for (int k = 0; k < 9; ++k) { S.Rs[lane * 9 + k] = R.m[k]; S.Rw[lane * 9 + k] = Rwl.m[k]; }
                S.rs[lane * 3] = r.x; S.rs[lane * 3 + 1] = r.y; S.rs[lane * 3 + 2] = r.z;
                S.pw[lane * 3] = pos.x; S.pw[lane * 3 + 1] = pos.y; S.pw[lane * 3 + 2] = pos.z;
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) S.Rwb[k] = Rwb.m[k];
                S.vel[0] = baseOmega.x; S.vel[1] = baseOmega.y; S.vel[2] = baseOmega.z; S.vel[3] = baseVel.x; S.vel[4] = baseVel.y; S.vel[5] = baseVel.z;
            }
            if (act && ndof == 3) { S.vel[dof0] = jv.x; S.vel[dof0 + 1] = jv.y; S.vel[dof0 + 2] = jv.z; }
            else if (act && ndof == 1) S.vel[dof0] = jv.x;
            __syncwarp();
        }
        // =================================================================== post-update flags of the update that just finished
        if (pending_flags) {
            pending_flags = false;
            need_action = 0;
            {   // cMathUtil::CheckNextInterval(dt, ctrl_time + init_time_offset, 1/30)
                const double cur = ctrl_time + init_off, pad = 0.001 * dt, T_ = M.query_dt;
                int c0 = static_cast<int>(floor((cur + pad) / T_)), c1 = static_cast<int>(floor((cur + pad - dt) / T_));
                need_action = (c0 != c1) ? 1 : 0;
            }
            // fall: any fall-contact link with a manifold point at distance <= 0.001*scale (state of the last sub-step's collision pass)
            const unsigned fb = __ballot_sync(0xffffffffu, act && L.fall_contact && in_contact_tol);
            const unsigned fseg = (W == 32) ? fb : ((fb >> (threadIdx.x & 16)) & 0xffffu);
            const int fallen = (fseg != 0 && M.enable_contact_fall) ? 1 : 0;
            // exploded velocities: any link |v|, |w| component > 100 in world axes (cSimCharacter::HasVelExploded)
            const M3 Rwl = own_Rw();
            V3 vw = mulT(Rwl, vel6.l) * (1.0f / M.scale), ww = mulT(Rwl, vel6.a);
            float mx = fmaxf(fmaxf(fmaxf(fabsf(vw.x), fabsf(vw.y)), fabsf(vw.z)), fmaxf(fmaxf(fabsf(ww.x), fabsf(ww.y)), fabsf(ww.z)));
            const unsigned eb = __ballot_sync(0xffffffffu, act && mx > 100.f);
            const unsigned eseg = (W == 32) ? eb : ((eb >> (threadIdx.x & 16)) & 0xffffu);
            if (alive) {
                int term = (M.enable_fall_end && fallen) ? 1 : 0;
                if (!term && !M.loop_motion && kin_time >= M.motion_dur) term = 1;
                f_updates++;
                const bool end = (timer >= timer_max) || term;
                if (end || stage == total_stages) {   // commit
                    if (lane == 0) {
                        reinterpret_cast<float4*>(sim)[0] = make_float4(basePos.x, basePos.y, basePos.z, 0.f);
                        reinterpret_cast<float4*>(sim)[1] = make_float4(baseQuat.x, baseQuat.y, baseQuat.z, baseQuat.w);
                        reinterpret_cast<float4*>(sim)[2] = make_float4(baseOmega.x, baseOmega.y, baseOmega.z, 0.f);
                        reinterpret_cast<float4*>(sim)[3] = make_float4(baseVel.x, baseVel.y, baseVel.z, 0.f);
                        tm[kTKin] = kin_time; tm[kTCtrl] = ctrl_time; tm[kTPrevAct] = prev_act; tm[kTTimer] = timer;
                        tm[kTOrigin] = org_x; tm[kTOrigin + 1] = org_y; tm[kTOrigin + 2] = org_z;
                        fl[kFNeedAction] = need_action; fl[kFDone] = end ? 1 : 0; fl[kFTerminate] = term; fl[kFValid] = (eseg == 0) ? 1 : 0; fl[kFFallen] = fallen;
                        fl[kFRowOverflow] = f_over; fl[kFUpdates] = f_updates;
                    }
                    if (act) {
                        reinterpret_cast<float4*>(sim + 16)[lane] = jp;
                        reinterpret_cast<float4*>(sim + 16 + 4 * nl)[lane] = jv;
                    }
                }
                if (end) alive = false;
            }
        }
        if (stage == total_stages) break;
        // stage-synchronous execution: every warp of the block runs the same stage at the same time, so the (large) kernel streams
        // through the instruction cache once per stage instead of once per warp
        if (sync_period != 0 && (stage % sync_period) == 0) { if (__syncthreads_and(!alive)) break; }
        const int ph = stage % stages_per_upd;      // 0: Stable-PD stage, 1..sim_substeps: Bullet sub-steps
        const bool first_upd = stage < stages_per_upd;
        int P = 0;
        if (ph == 0) {
            // ---------------- clocks: cScene::Update, cSceneImitate::UpdateKinChar, cDeepMimicCharController::UpdateCalcTau
            timer += dt;
            const double dur = M.motion_dur;
            double p0 = kin_time / dur; p0 -= floor(p0);
            kin_time += dt;
            double p1 = kin_time / dur; p1 -= floor(p1);
            if (M.loop_motion && p1 < p0 && M.sync_root_pos) {
                // SyncKinCharNewCycle: snap the clip's root x,z (at the new time) onto the simulated root
                int cyc = static_cast<int>(floor(kin_time / dur));
                double tt = kin_time - cyc * dur;
                int lo = 0, hi = M.num_frames - 1;   // upper_bound - 1
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (frame_times[mid] <= tt) lo = mid; else hi = mid; }
                double bl = (tt - frame_times[lo]) / (frame_times[lo + 1] - frame_times[lo]);
                bl = fmin(fmax(bl, 0.0), 1.0);
                const float* f0 = frames + static_cast<size_t>(lo) * M.pose_dim; const float* f1 = f0 + M.pose_dim;
                double rx = (1 - bl) * f0[0] + bl * f1[0] + cyc * static_cast<double>(M.cycle_delta[0]);
                double rz = (1 - bl) * f0[2] + bl * f1[2] + cyc * static_cast<double>(M.cycle_delta[2]);
                double qw = tm[kTOriginRot], qx = tm[kTOriginRot + 1], qy = tm[kTOriginRot + 2], qz = tm[kTOriginRot + 3];
                double ry_ = (1 - bl) * f0[1] + bl * f1[1];
                double ux = qy * rz - qz * ry_, uy = qz * rx - qx * rz, uz = qx * ry_ - qy * rx;
                ux *= 2; uy *= 2; uz *= 2;
                double kx = rx + qw * ux + (qy * uz - qz * uy);
                double kz = rz + qw * uz + (qx * uy - qy * ux);
                double sx = static_cast<double>(basePos.x) / M.scale, sz = static_cast<double>(basePos.z) / M.scale;
                org_x += sx - (kx + org_x);
                org_z += sz - (kz + org_z);
                org_y = 0.0;   // kin_root.y := ground_h + (kin_root.y - origin.y)  =>  origin.y returns to 0
            }
            ctrl_time += dt;
            if (need_action) { prev_act = ctrl_time; need_action = 0; }
        } else {
            // ---------------- collision: link convex vs plane y = 0, persistent manifold of <= 4 points per link (btPersistentManifold)
            float mp[48];
            int cnt = 0;
            {
                const float4* mg = reinterpret_cast<const float4*>(mani + li * kManifoldFloats);
#pragma unroll
                for (int k = 0; k < 12; ++k) { float4 v = mg[k]; mp[4 * k] = v.x; mp[4 * k + 1] = v.y; mp[4 * k + 2] = v.z; mp[4 * k + 3] = v.w; }
            }
            #pragma unroll 1
            for (int c = 0; c < 4; ++c) if (mp[c * 12] != 0.f && cnt == c) cnt = c + 1;
            const float thr = L.break_thr;
            const M3 Rwl = own_Rw();
            const V3 pos = own_pos();
            V3 dl = mul(Rwl, mk3(0.f, -1.f, 0.f));   // support direction -n in link coordinates
            V3 vtx;
            if (L.shape == kSBox) vtx = mk3(dl.x >= 0 ? L.he[0] : -L.he[0], dl.y >= 0 ? L.he[1] : -L.he[1], dl.z >= 0 ? L.he[2] : -L.he[2]);
            else {
                V3 sup = mk3(0, 0, 0);
                if (L.shape == kSCapsule) sup = mk3(0.f, (dl.y >= 0.f) ? L.he[1] : -L.he[1], 0.f);   // first end point wins ties
                float inv = rsqrtf(dot(dl, dl));
                vtx = sup + (L.he[0] * inv) * dl;
            }
            const V3 vw = pos + mulT(Rwl, vtx);
            const float dist = vw.y;
            if (act && dist < thr) {
                float best = thr * thr; int nearest = -1;
                #pragma unroll 1
                for (int c = 0; c < cnt; ++c) {
                    float dx = mp[c * 12 + 1] - vtx.x, dy = mp[c * 12 + 2] - vtx.y, dz = mp[c * 12 + 3] - vtx.z, dd = dx * dx + dy * dy + dz * dz;
                    if (dd < best) { best = dd; nearest = c; }
                }
                int idx = nearest;
                float k7 = 0, k8 = 0, k9 = 0, k11 = 0;
                if (nearest >= 0) { k7 = mp[nearest * 12 + 7]; k8 = mp[nearest * 12 + 8]; k9 = mp[nearest * 12 + 9]; k11 = mp[nearest * 12 + 11]; }
                else if (cnt < 4) { idx = cnt; cnt++; }
                else {   // btPersistentManifold::sortCachedPoints
                    int mpi = -1; float mpen = dist;
                    #pragma unroll 1
                    for (int c = 0; c < 4; ++c) if (mp[c * 12 + 10] < mpen) { mpi = c; mpen = mp[c * 12 + 10]; }
                    auto Pt = [&](int c) { return mk3(mp[c * 12 + 1], mp[c * 12 + 2], mp[c * 12 + 3]); };
                    auto area = [&](V3 a, V3 b) { V3 c = cross(a, b); return dot(c, c); };
                    float res[4] = {0, 0, 0, 0};
                    if (mpi != 0) res[0] = area(vtx - Pt(1), Pt(3) - Pt(2));
                    if (mpi != 1) res[1] = area(vtx - Pt(0), Pt(3) - Pt(2));
                    if (mpi != 2) res[2] = area(vtx - Pt(0), Pt(3) - Pt(1));
                    if (mpi != 3) res[3] = area(vtx - Pt(0), Pt(2) - Pt(1));
                    idx = 0; float bv = fabsf(res[0]);
                    #pragma unroll 1
                    for (int c = 1; c < 4; ++c) if (fabsf(res[c]) > bv) { bv = fabsf(res[c]); idx = c; }
                }
                float* q = mp + idx * 12;
                q[0] = 1.f; q[1] = vtx.x; q[2] = vtx.y; q[3] = vtx.z; q[4] = vw.x; q[5] = 0.f; q[6] = vw.z; q[7] = k7; q[8] = k8; q[9] = k9; q[10] = dist; q[11] = k11;
            }
            // refreshContactPoints
            #pragma unroll 1
            for (int c = cnt - 1; c >= 0; --c) {
                V3 pa = pos + mulT(Rwl, mk3(mp[c * 12 + 1], mp[c * 12 + 2], mp[c * 12 + 3]));
                mp[c * 12 + 10] = pa.y - mp[c * 12 + 5];
                mp[c * 12 + 11] += 1.f;
            }
            #pragma unroll 1
            for (int c = cnt - 1; c >= 0; --c) {
                V3 pa = pos + mulT(Rwl, mk3(mp[c * 12 + 1], mp[c * 12 + 2], mp[c * 12 + 3]));
                bool rm = !(mp[c * 12 + 10] <= thr);
                if (!rm) {
                    float dx = mp[c * 12 + 4] - pa.x, dy = mp[c * 12 + 5] - (pa.y - mp[c * 12 + 10]), dz = mp[c * 12 + 6] - pa.z;
                    rm = (dx * dx + dy * dy + dz * dz) > thr * thr;
                }
                if (rm) {
                    const int last = cnt - 1;
                    if (c != last) for (int k = 0; k < 12; ++k) mp[c * 12 + k] = mp[last * 12 + k];
                    mp[last * 12] = 0.f;
                    cnt--;
                }
            }
            if (!act) cnt = 0;
            in_contact_tol = false;   // cContactManager::Update: distance <= 0.001 * scale
            #pragma unroll 1
            for (int c = 0; c < cnt; ++c) if (mp[c * 12 + 10] <= 0.001f * M.scale) in_contact_tol = true;
            if (act && alive) {
                float4* mo = reinterpret_cast<float4*>(mani + li * kManifoldFloats);
#pragma unroll
                for (int k = 0; k < 12; ++k) mo[k] = make_float4(mp[4 * k], mp[4 * k + 1], mp[4 * k + 2], mp[4 * k + 3]);
            }
            // exclusive prefix over lanes -> point indices; publish points to the solver
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < W; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o, W); if (lane >= o) incl += t; }
            const int base = incl - cnt;
            #pragma unroll 1
            for (int c = 0; c < cnt; ++c) {
                const int p = base + c;
                if (p < maxpts) {
                    V3 pa = pos + mulT(Rwl, mk3(mp[c * 12 + 1], mp[c * 12 + 2], mp[c * 12 + 3]));
                    S.ppos[p * 4] = pa.x; S.ppos[p * 4 + 1] = pa.y; S.ppos[p * 4 + 2] = pa.z; S.ppos[p * 4 + 3] = mp[c * 12 + 10];
                    S.pimp[p] = mp[c * 12 + 7];
                    S.pref[p] = lane * 4 + c;
                } else f_over = 1;
            }
            P = min(T::shfli(incl, W - 1), maxpts);
        }

        // =================================================================== mass matrix (chain-sparse) + bias force
        // ph == 0: DeepMimic's SPD model (exact-shape inertia, cRBDUtil's root cj term); ph > 0: Bullet's multibody (collision-shape inertia)
        {
            const bool bullet_inertia = ph != 0, quirk = ph == 0;
            const float i0 = bullet_inertia ? L.inertiaB[0] : L.inertiaD[0], i1 = bullet_inertia ? L.inertiaB[1] : L.inertiaD[1],
                        i2 = bullet_inertia ? L.inertiaB[2] : L.inertiaD[2];
            const M3 Rwb = get_Rwb();
            // ---- composite rigid-body inertias, leaves -> root
            Rig comp;
            comp.m = act ? mass : 0.f; comp.h = mk3(0, 0, 0);
            comp.io[0] = act ? i0 : 0.f; comp.io[1] = 0; comp.io[2] = 0; comp.io[3] = act ? i1 : 0.f; comp.io[4] = 0; comp.io[5] = act ? i2 : 0.f;
            #pragma unroll 1
            for (int lv = maxlevel; lv >= 1; --lv) {
                const Rig send = rig_to_parent(comp, R, r);   // meaningful on lanes with level == lv (their subtree is complete)
                #pragma unroll 1
                for (int c = 0; c < kMaxChildren; ++c) {
                    const int cl = (c < nchild) ? ((child_pack >> (8 * c)) & 0xff) : -1;
                    const int src = cl >= 0 ? cl : lane;
                    const float gm_ = T::shfl(send.m, src);
                    const V3 gh = T::shfl3(send.h, src);
                    float gi[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) gi[k] = T::shfl(send.io[k], src);
                    if (cl >= 0 && level == lv - 1) {
                        comp.m += gm_; comp.h += gh;
#pragma unroll
                        for (int k = 0; k < 6; ++k) comp.io[k] += gi[k];
                    }
                }
            }
            // ---- H rows of this link's dofs: walk the ancestor chain
            #pragma unroll 1
            for (int d = 0; d < ndof; ++d) {
                S6 F = rig_mul(comp, subspace(lane, d));
                const int row = (dof0 + d) * cs;
                #pragma unroll 1
                for (int d2 = 0; d2 <= d; ++d2) S.H[row + depth0 + d2] = sdot(subspace(lane, d2), F);
                int cur = lane;
                #pragma unroll 1
                while (true) {
                    M3 Rc;
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rc.m[k] = S.Rs[cur * 9 + k];
                    F = xform_force_up(Rc, mk3(S.rs[cur * 3], S.rs[cur * 3 + 1], S.rs[cur * 3 + 2]), F);
                    const int p = lk_parent(cur);
                    if (p < 0) break;
                    cur = p;
                    const int nd = lk_ndof(cur), dp = lk_depth0(cur);
                    #pragma unroll 1
                    for (int d2 = 0; d2 < nd; ++d2) S.H[row + dp + d2] = sdot(subspace(cur, d2), F);
                }
                V3 fa = mulT(Rwb, F.a), fl_ = mulT(Rwb, F.l);
                S.H[row + 0] = fa.x; S.H[row + 1] = fa.y; S.H[row + 2] = fa.z; S.H[row + 3] = fl_.x; S.H[row + 4] = fl_.y; S.H[row + 5] = fl_.z;
            }
            if (lane == 0) {   // 6x6 base block from the whole-body composite expressed at the base origin, world axes
                const Rig cb = rig_to_parent(comp, R, r);   // base frame
                const M3 Rt = transpose(Rwb);                // base -> world
                const V3 hw = mul(Rt, cb.h);
                M3 A; A.m[0] = cb.io[0]; A.m[1] = A.m[3] = cb.io[1]; A.m[2] = A.m[6] = cb.io[2]; A.m[4] = cb.io[3]; A.m[5] = A.m[7] = cb.io[4]; A.m[8] = cb.io[5];
                const M3 Iw = mul(Rt, mul(A, Rwb));
                S.H[0 * cs + 0] = Iw.m[0];
                S.H[1 * cs + 0] = Iw.m[3]; S.H[1 * cs + 1] = Iw.m[4];
                S.H[2 * cs + 0] = Iw.m[6]; S.H[2 * cs + 1] = Iw.m[7]; S.H[2 * cs + 2] = Iw.m[8];
                S.H[3 * cs + 0] = 0.f;    S.H[3 * cs + 1] = hw.z;  S.H[3 * cs + 2] = -hw.y; S.H[3 * cs + 3] = cb.m;
                S.H[4 * cs + 0] = -hw.z;  S.H[4 * cs + 1] = 0.f;   S.H[4 * cs + 2] = hw.x;  S.H[4 * cs + 3] = 0.f; S.H[4 * cs + 4] = cb.m;
                S.H[5 * cs + 0] = hw.y;   S.H[5 * cs + 1] = -hw.x; S.H[5 * cs + 2] = 0.f;   S.H[5 * cs + 3] = 0.f; S.H[5 * cs + 4] = 0.f; S.H[5 * cs + 5] = cb.m;
            }
            // ---- bias force by recursive Newton-Euler (zero generalised acceleration, gravity as base acceleration -g)
            const S6 vb = mks(mul(Rwb, baseOmega), mul(Rwb, baseVel));
            const V3 w_used = quirk ? baseOmega : vb.a;   // cRBDUtil::BuildCjRoot differentiates the root quaternion with the body-frame
                                                          // formula applied to the world-frame angular velocity (RBDUtil.cpp:915-958)
            const S6 ab = mks(mk3(0, 0, 0), mul(Rwb, -grav) - cross(w_used, vb.l));
            const S6 cor = cross_motion(vel6, joint_twist());
            S6 acc;
            if (lane == 0) acc = xform_motion(R, r, ab);
            #pragma unroll 1
            for (int lv = 1; lv <= maxlevel; ++lv) {
                S6 pa = T::shfl6(acc, plane);
                if (level == lv) acc = xform_motion(R, r, pa) + cor;
            }
            S6 f;
            {
                V3 Iw = mk3(i0 * vel6.a.x, i1 * vel6.a.y, i2 * vel6.a.z);
                f.a = mk3(i0 * acc.a.x, i1 * acc.a.y, i2 * acc.a.z) + cross(vel6.a, Iw);
                f.l = mass * (acc.l + cross(vel6.a, vel6.l));
                if (!act) f = mks(mk3(0, 0, 0), mk3(0, 0, 0));
            }
            #pragma unroll 1
            for (int lv = maxlevel; lv >= 1; --lv) {
                const S6 send = xform_force_up(R, r, f);
                #pragma unroll 1
                for (int c = 0; c < kMaxChildren; ++c) {
                    const int cl = (c < nchild) ? ((child_pack >> (8 * c)) & 0xff) : -1;
                    const S6 g = T::shfl6(send, cl >= 0 ? cl : lane);
                    if (cl >= 0 && level == lv - 1) f = f + g;
                }
            }
            #pragma unroll 1
            for (int d = 0; d < ndof; ++d) S.bias[dof0 + d] = sdot(subspace(lane, d), f);
            if (lane == 0) {
                S6 fb = xform_force_up(R, r, f);
                V3 ca = mulT(Rwb, fb.a), cl_ = mulT(Rwb, fb.l);
                S.bias[0] = ca.x; S.bias[1] = ca.y; S.bias[2] = ca.z; S.bias[3] = cl_.x; S.bias[4] = cl_.y; S.bias[5] = cl_.z;
            }
            __syncwarp();
        }
        if (DEBUG && dbg && first_upd && ph == 0) {
            #pragma unroll 1
            for (int k = lane; k < n; k += W) { dbg[k] = S.bias[k]; dbg[kMaxDofs + k] = S.H[k * cs + HM.dof_depth[k]]; }
            #pragma unroll 1
            for (int k = lane; k < n * cs && k < 1024; k += W) dbg[8 * kMaxDofs + k] = S.H[k];
            __syncwarp();
        }

        // =================================================================== right-hand side
        float pe0 = 0.f, pe1 = 0.f, pe2 = 0.f;   // Kp * pose error of this joint (Stable-PD stage)
        const float fdt = static_cast<float>(dt);
        if (ph == 0) {
            // cImpPDController::CalcControlForces (ImpPDController.cpp:136-195) in the body-frame joint coordinates of the sim state
            float e0 = 0, e1 = 0, e2 = 0;
            if (jtype == kJSpherical) {
                Q4 q = mkq(jp.x, jp.y, jp.z, jp.w);
                // pose_inc = normalize(q + dt * 0.5 * q (x) (0, w))      (cKinTree::VelToPoseDiff, KinTree.cpp:1581-1610)
                Q4 dq = qmul(q, mkq(jv.x, jv.y, jv.z, 0.f));
                Q4 qi = qnormalize(mkq(q.x + 0.5f * fdt * dq.x, q.y + 0.5f * fdt * dq.y, q.z + 0.5f * fdt * dq.z, q.w + 0.5f * fdt * dq.w));
                V3 e = quat_rotvec(qmul(qconj(qi), mkq(tg.x, tg.y, tg.z, tg.w)));   // cKinTree::CalcVel(dt = 1) -> CalcQuaternionVelRel
                e0 = e.x; e1 = e.y; e2 = e.z;
            } else if (jtype == kJRevolute) {
                e0 = tg.x - (normalize_angle(jp.x) + fdt * jv.x);
            }
            const float kp = L.kp, kd = L.kd;
            pe0 = kp * e0; pe1 = kp * e1; pe2 = kp * e2;
            if (lane == 0) for (int k = 0; k < 6; ++k) S.bias[k] = -S.bias[k];
            if (ndof >= 1) { S.bias[dof0] = pe0 - kd * jv.x - S.bias[dof0]; S.H[dof0 * cs + depth0] += fdt * kd; }
            if (ndof == 3) {
                S.bias[dof0 + 1] = pe1 - kd * jv.y - S.bias[dof0 + 1]; S.H[(dof0 + 1) * cs + depth0 + 1] += fdt * kd;
                S.bias[dof0 + 2] = pe2 - kd * jv.z - S.bias[dof0 + 2]; S.H[(dof0 + 2) * cs + depth0 + 2] += fdt * kd;
            }
        } else {
            #pragma unroll 1
            for (int k = lane; k < n; k += W) S.bias[k] = S.tau[k] - S.bias[k];
        }
        __syncwarp();

        // =================================================================== Featherstone's sparse factorisation H = L^T D L in place (lanes = chain depth)
        // (A) joint dofs, leaves first: eliminating dof k touches only the rows of its JOINT ancestors here (columns incl. the 6 base columns)
#pragma unroll 1
        for (int k = n - 1; k >= 6; --k) {
            const int lk = HM.dof_link[k], dk = HM.dof_depth[k];
            const float hk = (lane <= dk) ? S.H[k * cs + lane] : 0.f;   // row k on its chain
            const float inv = 1.0f / T::shfl(hk, dk);
            const unsigned char* chn = HM.chain + lk * cs;
#pragma unroll 1
            for (int di = dk - 1; di >= 6; --di) {
                const float a = T::shfl(hk, di) * inv;
                if (lane <= di) S.H[chn[di] * cs + lane] -= a * hk;
            }
            if (lane < dk) S.H[k * cs + lane] = hk * inv;
            if (lane == dk) S.dinv[k] = inv;
            __syncwarp();
        }
        // (B) base block: Schur complement  B -= sum_k L_kb D_k L_kb'  accumulated in registers, lanes = (b, b') pairs of the lower triangle
        {
            // pair index q in [0,21): row b, col c <= b
            int pb[2], pc[2]; float acc[2] = {0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = lane + u * W;
                int b = 0;
                while ((b + 1) * (b + 2) / 2 <= q) ++b;
                pb[u] = b; pc[u] = q - b * (b + 1) / 2;
            }
            const int npass = (21 + W - 1) / W;
#pragma unroll 1
            for (int k = 6; k < n; ++k) {
                const float dkk = 1.0f / S.dinv[k];
#pragma unroll
                for (int u = 0; u < 2; ++u) if (u < npass && lane + u * W < 21) acc[u] += S.H[k * cs + pb[u]] * dkk * S.H[k * cs + pc[u]];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) if (u < npass && lane + u * W < 21) S.H[pb[u] * cs + pc[u]] -= acc[u];
            __syncwarp();
#pragma unroll 1
            for (int k = 5; k >= 0; --k) {   // dense 6x6 L^T D L of the base block
                const float hk = (lane <= k) ? S.H[k * cs + lane] : 0.f;
                const float inv = 1.0f / T::shfl(hk, k);
#pragma unroll 1
                for (int di = k - 1; di >= 0; --di) {
                    const float a = T::shfl(hk, di) * inv;
                    if (lane <= di) S.H[di * cs + lane] -= a * hk;
                }
                if (lane < k) S.H[k * cs + lane] = hk * inv;
                if (lane == k) S.dinv[k] = inv;
                __syncwarp();
            }
        }
        // =================================================================== x = M^-1 b in place on S.bias
        #pragma unroll 1
        for (int k = n - 1; k >= 1; --k) {   // b <- L^-T b
            const int lk = HM.dof_link[k], dk = HM.dof_depth[k];
            const float bk = S.bias[k];
            if (lane < dk) S.bias[HM.chain[lk * cs + lane]] -= S.H[k * cs + lane] * bk;
            __syncwarp();
        }
        #pragma unroll 1
        for (int k = lane; k < n; k += W) S.bias[k] *= S.dinv[k];
        __syncwarp();
        #pragma unroll 1
        for (int k = 1; k < n; ++k) {        // b <- L^-1 b
            const int lk = HM.dof_link[k], dk = HM.dof_depth[k];
            float part = 0.f;
            if (lane < dk) part = S.H[k * cs + lane] * S.bias[HM.chain[lk * cs + lane]];
            part = T::sum(part);
            if (lane == 0) S.bias[k] -= part;
            __syncwarp();
        }

        if (ph == 0) {
            // ---------------- torques: tau = Kp e + Kd (edot - dt a), clamped by norm (cSimBodyJoint::ClampTotalTorque, SimBodyJoint.cpp:299-307)
            const float kd = L.kd;
            float t0 = 0, t1 = 0, t2 = 0;
            if (ndof >= 1) t0 = pe0 + kd * (-jv.x - fdt * S.bias[dof0]);
            if (ndof == 3) { t1 = pe1 + kd * (-jv.y - fdt * S.bias[dof0 + 1]); t2 = pe2 + kd * (-jv.z - fdt * S.bias[dof0 + 2]); }
            const float mag = sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
            if (mag > L.tlim) { float s = L.tlim / mag; t0 *= s; t1 *= s; t2 *= s; }
            __syncwarp();
            if (lane == 0) for (int k = 0; k < 6; ++k) S.tau[k] = 0.f;
            if (ndof >= 1) S.tau[dof0] = t0;
            if (ndof == 3) { S.tau[dof0 + 1] = t1; S.tau[dof0 + 2] = t2; }
            __syncwarp();
            if (DEBUG && dbg && first_upd) { for (int k = lane; k < n; k += W) { dbg[2 * kMaxDofs + k] = S.tau[k]; dbg[3 * kMaxDofs + k] = S.bias[k]; } }
            continue;
        }

        // =================================================================== Bullet sub-step: v += a h, constraint rows, PGS, integration
        const int sub = ph - 1;
        if (DEBUG && dbg && first_upd) { for (int k = lane; k < n; k += W) dbg[(sub == 0 ? 4 * kMaxDofs : 8 * kMaxDofs + 1024) + k] = S.bias[k]; }
        #pragma unroll 1
        for (int k = lane; k < n; k += W) { float v = S.vel[k] + S.bias[k] * h; S.vel[k] = fminf(fmaxf(v, -100.f), 100.f); S.z[k] = 0.f; }   // applyDeltaVeeMultiDof clamp
        __syncwarp();
        if (DEBUG && dbg && first_upd) { for (int k = lane; k < n; k += W) dbg[(sub == 0 ? 5 * kMaxDofs : 9 * kMaxDofs + 1024) + k] = S.vel[k]; if (lane == 0) dbg[(sub == 0 ? 7 * kMaxDofs : 11 * kMaxDofs + 1024)] = static_cast<float>(P); }
        // ---- joint-limit rows (btMultiBodyJointLimitConstraint): a lane owns at most one active row
        int lim_dir = 0; float lim_pen = 0.f;
        if (act && L.has_limit) {
            float p0 = jp.x - L.lim_lo, p1 = L.lim_hi - jp.x;
            if (!(p0 > 0.f)) { lim_dir = 1; lim_pen = p0; }
            else if (!(p1 > 0.f)) { lim_dir = -1; lim_pen = p1; }
        }
        int linc = lim_dir != 0 ? 1 : 0;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, linc, o, W); if (lane >= o) linc += t; }
        const int NLm = T::shfli(linc, W - 1);
        const int NR = 3 * P + NLm;
        const int Pmax = (W == 32) ? P : warp_max(P);
        const int Lmax = (W == 32) ? NLm : warp_max(NLm);
        const V3 basePosNow = basePos;
        // ---- contact rows: lanes = chain depth.  Rows: [0,P) normals, [P,3P) friction pairs (t1 = -x, t2 = +z), [3P,NR) limits
        #pragma unroll 1
        for (int p = 0; p < Pmax; ++p) {
            const bool pv = p < P;
            const int b = pv ? (S.pref[p] >> 2) : 0;
            const int lastd = lk_lastd(b);   // deepest dof on the chain base -> link b
            const int lastd_u = (W == 32) ? lastd : warp_max(lastd);
            const V3 pa = pv ? mk3(S.ppos[p * 4], S.ppos[p * 4 + 1], S.ppos[p * 4 + 2]) : mk3(0, 0, 0);
            const float pdist = pv ? S.ppos[p * 4 + 3] : 0.f;
            float jn = 0.f, j1 = 0.f, j2 = 0.f;
            int idof = 0;
            const bool on = pv && lane <= lastd;
            if (on) {
                idof = HM.chain[b * cs + lane];
                V3 cn;   // (rel x n), (rel x t1), (rel x t2) share rel; J = top . (rel x dir)
                V3 top, rel;
                if (idof < 6) { rel = pa - basePosNow; top = unit3(idof < 3 ? idof : 0); }
                else {
                    const int a = HM.dof_link[idof];
                    const int d = idof - lk_dof0(a);
                    const float* Ra = S.Rw + a * 9;
                    V3 topl = (lk_jtype(a) == kJSpherical) ? unit3(d) : mk3(HM.axis[a * 3], HM.axis[a * 3 + 1], HM.axis[a * 3 + 2]);
                    top = mk3(Ra[0] * topl.x + Ra[3] * topl.y + Ra[6] * topl.z, Ra[1] * topl.x + Ra[4] * topl.y + Ra[7] * topl.z, Ra[2] * topl.x + Ra[5] * topl.y + Ra[8] * topl.z);
                    V3 dl = mk3(HM.dvec[a * 3], HM.dvec[a * 3 + 1], HM.dvec[a * 3 + 2]);
                    V3 dw = mk3(Ra[0] * dl.x + Ra[3] * dl.y + Ra[6] * dl.z, Ra[1] * dl.x + Ra[4] * dl.y + Ra[7] * dl.z, Ra[2] * dl.x + Ra[5] * dl.y + Ra[8] * dl.z);
                    rel = pa - (mk3(S.pw[a * 3], S.pw[a * 3 + 1], S.pw[a * 3 + 2]) - dw);
                }
                if (idof >= 3 && idof < 6) { jn = (idof == 4) ? 1.f : 0.f; j1 = (idof == 3) ? -1.f : 0.f; j2 = (idof == 5) ? 1.f : 0.f; }
                else {
                    // rel x (0,1,0) = (-rz, 0, rx) ; rel x (-1,0,0) = (0, -rz, ry) ; rel x (0,0,1) = (ry, -rx, 0)
                    jn = -top.x * rel.z + top.z * rel.x;
                    j1 = -top.y * rel.z + top.z * rel.y;
                    j2 = top.x * rel.y - top.y * rel.x;
                }
                (void)cn;
            }
            const float vi = on ? S.vel[idof] : 0.f;
            const float rvn = T::sum(jn * vi), rv1 = T::sum(j1 * vi), rv2 = T::sum(j2 * vi);
            // y = D^-1/2 L^-T J^T restricted to the chain
            #pragma unroll 1
            for (int dk = lastd_u; dk >= 1; --dk) {
                const bool ok = pv && dk <= lastd;
                const int kd = ok ? HM.chain[b * cs + dk] : 0;
                const float tn = T::shfl(jn, dk), t1_ = T::shfl(j1, dk), t2_ = T::shfl(j2, dk);
                if (ok && lane < dk) { const float l = S.H[kd * cs + lane]; jn -= l * tn; j1 -= l * t1_; j2 -= l * t2_; }
            }
            const float sd = on ? sqrtf(S.dinv[idof]) : 0.f;
            jn *= sd; j1 *= sd; j2 *= sd;
            const float an = T::sum(jn * jn), a1 = T::sum(j1 * j1), a2 = T::sum(j2 * j2);
            if (pv) {
                const int rn = p, r1 = P + 2 * p, r2 = P + 2 * p + 1;
                if (lane < cs) { S.Y[rn * cs + lane] = on ? jn : 0.f; S.Y[r1 * cs + lane] = on ? j1 : 0.f; S.Y[r2 * cs + lane] = on ? j2 : 0.f; }
                const float l0 = S.pimp[p] * 0.85f;   // SOLVER_USE_WARMSTARTING, warmstartingFactor 0.85
                if (lane == 0) {
                    const float invn = an > 1.1920929e-7f ? 1.0f / an : 0.f, inv1 = a1 > 1.1920929e-7f ? 1.0f / a1 : 0.f, inv2 = a2 > 1.1920929e-7f ? 1.0f / a2 : 0.f;
                    // setupMultiBodyContactConstraint: erp 0.2, restitution 0, no split impulse for multibodies
                    float perr = 0.f, verr = -rvn;
                    if (pdist > 0.f) verr -= pdist / h; else perr = -pdist * 0.2f / h;
                    S.rrhs[rn] = perr * invn + verr * invn; S.rinv[rn] = invn; S.rlam[rn] = l0; S.rlink[rn] = b;
                    S.rrhs[r1] = -rv1 * inv1; S.rinv[r1] = inv1; S.rlam[r1] = 0.f; S.rlink[r1] = b;
                    S.rrhs[r2] = -rv2 * inv2; S.rinv[r2] = inv2; S.rlam[r2] = 0.f; S.rlink[r2] = b;
                }
                if (l0 != 0.f && on) S.z[idof] += jn * l0;   // warm start: z += y * lambda0
            }
            __syncwarp();
        }
        // ---- limit rows
        #pragma unroll 1
        for (int q = 0; q < Lmax; ++q) {
            const bool qv = q < NLm;
            const unsigned bal = __ballot_sync(0xffffffffu, qv && lim_dir != 0 && (linc - 1) == q);
            const unsigned seg = (W == 32) ? bal : ((bal >> ((threadIdx.x & 16))) & 0xffffu);
            const int b = seg ? (__ffs(seg) - 1) : 0;
            const float dirf = T::shfl((lim_dir == -1) ? -1.f : 1.f, b);
            const float pen = T::shfl(lim_pen, b);
            const int lastd = lk_depth0(b);   // revolute: single dof
            const int lastd_u = (W == 32) ? lastd : warp_max(lastd);
            float j = (qv && lane == lastd) ? dirf : 0.f;
            const bool on = qv && lane <= lastd;
            const int idof = on ? HM.chain[b * cs + lane] : 0;
            const float rv = T::sum(j * (on ? S.vel[idof] : 0.f));
            #pragma unroll 1
            for (int dk = lastd_u; dk >= 1; --dk) {
                const bool ok = qv && dk <= lastd;
                const int kd = ok ? HM.chain[b * cs + dk] : 0;
                const float tq = T::shfl(j, dk);
                if (ok && lane < dk) j -= S.H[kd * cs + lane] * tq;
            }
            j *= on ? sqrtf(S.dinv[idof]) : 0.f;
            const float aq = T::sum(j * j);
            if (qv) {
                const int rr = 3 * P + q;
                if (rr < maxrows) {
                    if (lane < cs) S.Y[rr * cs + lane] = on ? j : 0.f;
                    if (lane == 0) {
                        const float inv = aq > 1.1920929e-7f ? 1.0f / aq : 0.f;
                        float perr = 0.f, verr = -rv;
                        const bool combine = pen > -0.04f;   // split-impulse threshold: deeper violations lose the positional term (btMultiBodyJointLimitConstraint)
                        if (pen > 0.f) verr = -pen / h; else perr = -pen * 0.2f / h;
                        S.rrhs[rr] = combine ? (perr * inv + verr * inv) : (verr * inv);
                        S.rinv[rr] = inv; S.rlam[rr] = 0.f; S.rlink[rr] = b;
                    }
                } else f_over = 1;
            }
            __syncwarp();
        }
        // ---- projected Gauss-Seidel, 10 sweeps (btMultiBodyConstraintSolver::solveSingleIteration ordering)
        {
            // one code instance of the row update; the sweep order is generated by a flat index:  [limits | normals | frictions]
            const int per_it = Lmax + 3 * Pmax;
            #pragma unroll 1
            for (int t = 0; t < 10 * per_it; ++t) {
                const int it = t / per_it, u = t - it * per_it;
                int rr; bool valid; float lo = 0.f, hi;
                if (u < Lmax) { const int qq = (it & 1) ? u : NLm - 1 - u; rr = 3 * P + qq; valid = u < NLm && rr < maxrows && qq >= 0; hi = 100.f; }
                else if (u < Lmax + Pmax) { rr = u - Lmax; valid = rr < P; hi = 1e10f; }
                else {
                    const int p = u - Lmax - Pmax;
                    valid = p < 2 * P;
                    const float tot = valid ? S.rlam[p >> 1] : 0.f;
                    hi = M.friction * tot; lo = -hi; rr = P + p; valid = valid && tot > 0.f;
                }
                const int b = valid ? S.rlink[rr] : 0;
                const int lastd = lk_lastd(b);
                const bool on = valid && lane <= lastd;
                const int idof = on ? HM.chain[b * cs + lane] : 0;
                const float y = on ? S.Y[rr * cs + lane] : 0.f;
                const float lam = valid ? S.rlam[rr] : 0.f;
                const float dv = T::sum(y * (on ? S.z[idof] : 0.f));
                float dI = (valid ? S.rrhs[rr] : 0.f) - dv * (valid ? S.rinv[rr] : 0.f);
                float sum = lam + dI;
                if (sum < lo) { dI = lo - lam; sum = lo; } else if (sum > hi) { dI = hi - lam; sum = hi; }
                __syncwarp();
                if (valid && lane == 0) S.rlam[rr] = sum;
                if (on) S.z[idof] += y * dI;
                __syncwarp();
            }
            // write impulses back to the manifold (warm start of the next sub-step)
            #pragma unroll 1
            for (int p = lane; p < P; p += W) {
                if (alive) {
                    const int ref = S.pref[p];
                    float* mpt = mani + (ref >> 2) * kManifoldFloats + (ref & 3) * 12;
                    mpt[7] = S.rlam[p]; mpt[8] = S.rlam[P + 2 * p]; mpt[9] = S.rlam[P + 2 * p + 1];
                }
            }
            // dv = L^-1 D^-1/2 z ; v += dv (clamped)
            #pragma unroll 1
            for (int k = lane; k < n; k += W) S.z[k] *= sqrtf(S.dinv[k]);
            __syncwarp();
            #pragma unroll 1
            for (int k = 1; k < n; ++k) {
                const int lk = HM.dof_link[k], dk = HM.dof_depth[k];
                float part = 0.f;
                if (lane < dk) part = S.H[k * cs + lane] * S.z[HM.chain[lk * cs + lane]];
                part = T::sum(part);
                if (lane == 0) S.z[k] -= part;
                __syncwarp();
            }
            #pragma unroll 1
            for (int k = lane; k < n; k += W) { float v = S.vel[k] + S.z[k]; S.vel[k] = fminf(fmaxf(v, -100.f), 100.f); }
            __syncwarp();
            if (DEBUG && dbg && first_upd) { for (int k = lane; k < n; k += W) dbg[(sub == 0 ? 6 * kMaxDofs : 10 * kMaxDofs + 1024) + k] = S.vel[k]; for (int k = lane; k < NR && k < 60; k += W) dbg[(sub == 0 ? 7 * kMaxDofs : 11 * kMaxDofs + 1024) + 1 + k] = S.rlam[k]; }
        }
        // ---- integrate positions (btMultiBody::stepPositionsMultiDof)
        baseOmega = mk3(S.vel[0], S.vel[1], S.vel[2]); baseVel = mk3(S.vel[3], S.vel[4], S.vel[5]);
        if (ndof == 3) { jv.x = S.vel[dof0]; jv.y = S.vel[dof0 + 1]; jv.z = S.vel[dof0 + 2]; }
        else if (ndof == 1) jv.x = S.vel[dof0];
        basePos = basePos + h * baseVel;
        baseQuat = quat_integrate(baseOmega, baseQuat, true, h);
        if (jtype == kJRevolute) jp.x += h * jv.x;
        else if (jtype == kJSpherical) { Q4 q = quat_integrate(mk3(jv.x, jv.y, jv.z), mkq(jp.x, jp.y, jp.z, jp.w), false, h); jp = make_float4(q.x, q.y, q.z, q.w); }
        __syncwarp();
        need_kin = true;
        if (ph == sim_substeps) pending_flags = true;
    }
}

// explicit instantiations used by capi.cu: (tile width, debug dumps)
template __global__ void dm_update_kernel<16, false>(const DevModel*, DevState, const double*, const float*, double, int, int, int, int);
template __global__ void dm_update_kernel<32, false>(const DevModel*, DevState, const double*, const float*, double, int, int, int, int);
template __global__ void dm_update_kernel<16, true>(const DevModel*, DevState, const double*, const float*, double, int, int, int, int);
template __global__ void dm_update_kernel<32, true>(const DevModel*, DevState, const double*, const float*, double, int, int, int, int);

int dm_update_smem_bytes(int nl, int n, int cs, int maxrows, int tiles) {
    return hot_model_bytes(nl, n, cs) + smem_floats_per_env(nl, n, cs, maxrows) * tiles * static_cast<int>(sizeof(float));
}

}  // namespace dmk

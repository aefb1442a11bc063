// Joints on the device: per-step layout, per-substep row fill + preSolve (NGS pseudo velocities), per-iteration solve.
//
// Reference (paths under /root/reference):
//   row layout per step            src/Physecs.cpp:322-354, Joints/*.cpp getSolverDesc, Constraint1DContainer.h:133-197
//   world-space joint data         src/Joint.cpp:5-14
//   row builders                   src/JointsUtil.h:8-28 (point-to-point), src/Joints/{Fixed,Revolute,Spherical,Universal,
//                                  Prismatic,Gear,Servo}Joint.cpp makeConstraints
//   preSolve / solve (colours 0-7) src/Constraint1DW.cpp:6-115 / :118-233   (SIMD semantics, quirk Q9/Q10)
//   overflow colour (index 8)      src/Constraint1D.cpp:6-52 / :55-127      (scalar semantics, sequential)
//   greedy colouring               src/Physecs.cpp:690-710 (host side: physecs_b200/joint_colors.py / host Scene)
//
// The reference packs rows of four different joints of one colour into an SSE register; joints of one colour touch
// disjoint entities, so "one thread per joint, rows in (flag list, creation) order" is the same arithmetic per lane.
// Flag-list order is NONE, ANGULAR, SOFT, LIMITED, ANGULAR|SOFT, ANGULAR|LIMITED (Constraint1DContainer.h:117).
#include "pb_ctx.h"
#include "pb_math.cuh"
#include <algorithm>
#include <vector>

#define JF_SOFT 1
#define JF_ANGULAR 2
#define JF_LIMITED 4
#define MAXR PB_MAX_JOINT_ROWS

struct JointDev {
    int n; int nDyn;
    const int* type; const int2* rows; const float4* a0p; const float4* a0q; const float4* a1p; const float4* a1q;
    const float4* prm;     // [2*J]
    float4* state;         // [2*J] persistent: prismatic {makeUpper, makeLower}, gear {pa0, pa1, va0, va1 | init}
    // rows, index r*J + j
    float4* linC;          // linear xyz, c
    float4* a0T;           // angular0 xyz, targetVelocity
    float4* a1K;           // angular1 xyz, invEffMass
    float4* a0tMin;        // I0^-1 angular0 xyz, min
    float4* a1tMax;        // I1^-1 angular1 xyz, max
    float2* soft;          // frequency, dampingRatio
    float* lambda;         // totalLambda (persists across the substeps of a step)
};

__device__ __forceinline__ int jSolverIndex(int row, int nDyn, const int* __restrict__ kinematic) {
    return (row < nDyn && !kinematic[row]) ? row : -1;
}

// number of rows and the flag of row r in SOLVE order, given joint type and per-step state
__device__ __forceinline__ int jointRowCount(int type, float4 prm0, float4 st) {
    switch (type) {
        case PB_JOINT_FIXED: return 6;
        case PB_JOINT_REVOLUTE: return prm0.x != 0.f ? 6 : 5;
        case PB_JOINT_SPHERICAL: return 3;
        case PB_JOINT_UNIVERSAL: return 4;
        case PB_JOINT_PRISMATIC: return 5 + ((st.x != 0.f || st.y != 0.f) ? 1 : 0) + (prm0.z != 0.f ? 1 : 0);
        case PB_JOINT_GEAR: return 1;
        case PB_JOINT_SERVO: return 6;
    }
    return 0;
}
__device__ __forceinline__ int jointRowFlags(int type, int r, float4 prm0, float4 st) {
    switch (type) {
        case PB_JOINT_FIXED: return r < 3 ? 0 : JF_ANGULAR;
        case PB_JOINT_REVOLUTE: return r < 3 ? 0 : (r < 5 ? JF_ANGULAR : (JF_ANGULAR | JF_LIMITED));
        case PB_JOINT_SPHERICAL: return 0;
        case PB_JOINT_UNIVERSAL: return r < 3 ? 0 : JF_ANGULAR;
        case PB_JOINT_PRISMATIC: {
            if (r < 2) return 0;
            if (r < 5) return JF_ANGULAR;
            // solve order: SOFT list before LIMITED list
            bool drive = prm0.z != 0.f;
            if (r == 5) return drive ? JF_SOFT : JF_LIMITED;
            return JF_LIMITED;
        }
        case PB_JOINT_GEAR: return JF_ANGULAR;
        case PB_JOINT_SERVO: return r < 3 ? 0 : (r < 5 ? JF_ANGULAR : (JF_ANGULAR | JF_SOFT));
    }
    return 0;
}

struct Row { V3 lin, a0, a1; float c, target, mn, mx, freq, damp; };
__device__ __forceinline__ Row rowDefault() {
    Row r; r.lin = r.a0 = r.a1 = mk3(0.f); r.c = 0.f; r.target = 0.f; r.mn = -FLT_MAX; r.mx = FLT_MAX; r.freq = 0.f; r.damp = 0.f; return r;
}
__device__ __forceinline__ void p2pRows(V3 p0, V3 p1, V3 r0, V3 r1, Row* rows) {
    V3 d = p1 - p0;
    rows[0] = rowDefault(); rows[0].lin = mk3(1, 0, 0); rows[0].a0 = mk3(0.f, r0.z, -r0.y); rows[0].a1 = mk3(0.f, r1.z, -r1.y); rows[0].c = d.x;
    rows[1] = rowDefault(); rows[1].lin = mk3(0, 1, 0); rows[1].a0 = mk3(-r0.z, 0.f, r0.x); rows[1].a1 = mk3(-r1.z, 0.f, r1.x); rows[1].c = d.y;
    rows[2] = rowDefault(); rows[2].lin = mk3(0, 0, 1); rows[2].a0 = mk3(r0.y, -r0.x, 0.f); rows[2].a1 = mk3(r1.y, -r1.x, 0.f); rows[2].c = d.z;
}
__device__ __forceinline__ Row angRow(V3 ua, V3 ub) {   // c = dot(ua, ub), angular = cross(ub, ua)
    Row r = rowDefault(); V3 a = cross(ub, ua); r.a0 = a; r.a1 = a; r.c = dot(ua, ub); return r;
}
__device__ __forceinline__ float angleDiff(float a0, float a1) {   // GearJoint.cpp:5-8
    const float pi = 3.14159265358979323846f, twoPi = 6.28318530717958647692f;
    float diff = fmodf(a1 - a0 + pi, twoPi) - pi;
    return diff < -pi ? diff + twoPi : diff;
}

// builds the rows of joint j in SOLVE order; returns the row count
__device__ int buildJointRows(int type, float4 prm0, float4 prm1, float4& st0, float4& st1, bool advanceState,
                              V3 p0, V3 p1, V3 r0, V3 r1, const M3& u0, const M3& u1, Row* rows) {
    switch (type) {
        case PB_JOINT_SPHERICAL: p2pRows(p0, p1, r0, r1, rows); return 3;
        case PB_JOINT_FIXED:
            p2pRows(p0, p1, r0, r1, rows);
            rows[3] = angRow(u0.c[0], u1.c[1]); rows[4] = angRow(u0.c[0], u1.c[2]); rows[5] = angRow(u0.c[1], u1.c[2]);
            return 6;
        case PB_JOINT_REVOLUTE: {
            p2pRows(p0, p1, r0, r1, rows);
            rows[3] = angRow(u0.c[0], u1.c[1]); rows[4] = angRow(u0.c[0], u1.c[2]);
            if (prm0.x != 0.f) {
                Row d = rowDefault(); d.a0 = u0.c[0]; d.a1 = u0.c[0]; d.target = prm0.y; d.mx = prm0.z; d.mn = -prm0.z;
                rows[5] = d; return 6;
            }
            return 5;
        }
        case PB_JOINT_UNIVERSAL:
            p2pRows(p0, p1, r0, r1, rows);
            rows[3] = angRow(u0.c[2], u1.c[2]);
            return 4;
        case PB_JOINT_SERVO: {
            p2pRows(p0, p1, r0, r1, rows);
            rows[3] = angRow(u0.c[0], u1.c[1]); rows[4] = angRow(u0.c[0], u1.c[2]);
            Row d = rowDefault(); d.a0 = u0.c[0]; d.a1 = u0.c[0];
            // glm::orientedAngle(u0[2], u1[2], u0[0])  (gtx/vector_angle.inl:37-43)
            float ang = acosf(gclamp(dot(u0.c[2], u1.c[2]), -1.f, 1.f));
            if (dot(u0.c[0], cross(u0.c[2], u1.c[2])) < 0.f) ang = -ang;
            d.c = ang - prm0.x; d.freq = prm0.y; d.damp = prm0.z;
            rows[5] = d; return 6;
        }
        case PB_JOINT_PRISMATIC: {
            // prm0 = {upper, lower, driveEnabled, targetPosition}, prm1 = {stiffness, damping}; st0 = {makeUpper, makeLower}
            V3 d = p1 - p0;
            Row a = rowDefault(); a.lin = u0.c[1]; a.a0 = cross(r0, u0.c[1]); a.a1 = cross(r1, u0.c[1]); a.c = dot(d, u0.c[1]); rows[0] = a;
            Row b = rowDefault(); b.lin = u0.c[2]; b.a0 = cross(r0, u0.c[2]); b.a1 = cross(r1, u0.c[2]); b.c = dot(d, u0.c[2]); rows[1] = b;
            rows[2] = angRow(u0.c[0], u1.c[1]); rows[3] = angRow(u0.c[0], u1.c[2]); rows[4] = angRow(u0.c[1], u1.c[2]);
            float dx = dot(d, u0.c[0]);
            V3 r0xx = cross(r0, u0.c[0]), r1xx = cross(r1, u0.c[0]);
            int n = 5;
            bool drive = prm0.z != 0.f;
            if (drive) {   // SOFT list is solved before the LIMITED list
                Row s = rowDefault(); s.lin = u0.c[0]; s.a0 = r0xx; s.a1 = r1xx; s.c = dx - prm0.w; s.freq = prm1.x; s.damp = prm1.y;
                rows[n++] = s;
            }
            if (st0.x != 0.f) { Row l = rowDefault(); l.lin = u0.c[0]; l.a0 = r0xx; l.a1 = r1xx; l.c = dx - prm0.x; l.mn = 0.f; rows[n++] = l; }
            else if (st0.y != 0.f) { Row l = rowDefault(); l.lin = u0.c[0]; l.a0 = r0xx; l.a1 = r1xx; l.c = dx - prm0.y; l.mx = 0.f; rows[n++] = l; }
            return n;
        }
        case PB_JOINT_GEAR: {
            // st0 = {persistentAngle0, persistentAngle1, virtualAngle0, virtualAngle1}, st1.x = isInitialized; prm0.x = ratio
            float angle0, angle1;
            {
                V3 p1Proj = p1 + dot(p0 - p1, u0.c[0]) * u0.c[0];
                V3 dir = normalize(p0 - p1Proj);
                V3 n = cross(u0.c[0], dir);
                M3 m; m.c[0] = u0.c[0]; m.c[1] = -n; m.c[2] = dir;
                M3 u0t = mul(transpose(m), u0);
                angle0 = atan2f(u0t.c[1].z, u0t.c[2].z);
            }
            {
                V3 p0Proj = p0 + dot(p1 - p0, u1.c[0]) * u1.c[0];
                V3 dir = normalize(p0Proj - p1);
                V3 n = cross(u1.c[0], dir);
                M3 m; m.c[0] = u1.c[0]; m.c[1] = -n; m.c[2] = dir;
                M3 u1t = mul(transpose(u1), m);
                angle1 = atan2f(u1t.c[1].z, u1t.c[2].z);
            }
            float pa0 = st0.x, pa1 = st0.y, va0 = st0.z, va1 = st0.w;
            if (st1.x == 0.f) { pa0 = angle0; pa1 = angle1; }
            va0 += angleDiff(angle0, pa0);
            va1 += angleDiff(angle1, pa1);
            if (advanceState) { st0 = make_float4(angle0, angle1, va0, va1); st1.x = 1.f; }
            Row g = rowDefault(); g.a0 = u0.c[0] * prm0.x; g.a1 = -u1.c[0]; g.c = va0 * prm0.x - va1;
            rows[0] = g; return 1;
        }
    }
    return 0;
}

// once per step (Physecs.cpp:322-354): reset accumulated impulses, prismatic limit selection (PrismaticJoint.cpp:116-155)
__global__ void k_joint_begin(JointDev J, const float4* __restrict__ pos, const float4* __restrict__ quat) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J.n) return;
    for (int r = 0; r < MAXR; ++r) J.lambda[r * J.n + j] = 0.f;
    if (J.type[j] == PB_JOINT_PRISMATIC) {
        int2 rr = J.rows[j];
        Q4 q0 = mkq(quat[rr.x]), q1 = mkq(quat[rr.y]);
        V3 p0 = mk3(pos[rr.x]) + rotate(q0, mk3(J.a0p[j]));
        V3 p1 = mk3(pos[rr.y]) + rotate(q1, mk3(J.a1p[j]));
        V3 d = p1 - p0;
        V3 u00 = rotate(qmul(q0, mkq(J.a0q[j])), mk3(1.f, 0.f, 0.f));
        float dx = dot(d, u00);
        float4 prm0 = J.prm[2 * j];
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dx > prm0.x) st.x = 1.f;
        else if (dx < prm0.y) st.y = 1.f;
        J.state[2 * j] = st;
    }
}

// per substep and colour: fill rows (makeConstraints), effective masses, NGS pseudo-velocity pass (Constraint1DW.cpp:6-115)
__global__ void k_joint_prep(JointDev J, int start, int count, int doNgs, const int* __restrict__ kinematic,
                             const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ comInvMass,
                             const float4* __restrict__ invIW, float4* __restrict__ pseudoLin, float4* __restrict__ pseudoAng) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int j = start + i;
    int type = J.type[j];
    int2 rr = J.rows[j];
    int b0 = jSolverIndex(rr.x, J.nDyn, kinematic), b1 = jSolverIndex(rr.y, J.nDyn, kinematic);
    Q4 q0 = mkq(quat[rr.x]), q1 = mkq(quat[rr.y]);
    V3 a0p = mk3(J.a0p[j]), a1p = mk3(J.a1p[j]);
    // JointSolverData r0/r1 (Physecs.cpp:344-345) and calculateWorldSpaceData (Joint.cpp:5-14)
    V3 r0l = b0 >= 0 ? a0p - mk3(comInvMass[b0]) : a0p;
    V3 r1l = b1 >= 0 ? a1p - mk3(comInvMass[b1]) : a1p;
    M3 u0 = mat3_cast(qmul(q0, mkq(J.a0q[j]))), u1 = mat3_cast(qmul(q1, mkq(J.a1q[j])));
    V3 r0 = rotate(q0, r0l), r1 = rotate(q1, r1l);
    V3 p0 = mk3(pos[rr.x]) + rotate(q0, a0p), p1 = mk3(pos[rr.y]) + rotate(q1, a1p);
    float4 prm0 = J.prm[2 * j], prm1 = J.prm[2 * j + 1];
    float4 st0 = J.state[2 * j], st1 = J.state[2 * j + 1];
    Row rows[MAXR];
    int n = buildJointRows(type, prm0, prm1, st0, st1, true, p0, p1, r0, r1, u0, u1, rows);
    if (type == PB_JOINT_GEAR) { J.state[2 * j] = st0; J.state[2 * j + 1] = st1; }
    float im0 = 0.f, im1 = 0.f;
    M3 I0, I1; I0.c[0] = I0.c[1] = I0.c[2] = mk3(0.f); I1 = I0;
    V3 pv0 = mk3(0.f), pw0 = mk3(0.f), pv1 = mk3(0.f), pw1 = mk3(0.f);
    int cnt0 = 0, cnt1 = 0;
    if (b0 >= 0) {
        im0 = comInvMass[b0].w; I0.c[0] = mk3(invIW[3 * b0]); I0.c[1] = mk3(invIW[3 * b0 + 1]); I0.c[2] = mk3(invIW[3 * b0 + 2]);
        float4 l = pseudoLin[b0]; pv0 = mk3(l); cnt0 = __float_as_int(l.w); pw0 = mk3(pseudoAng[b0]);
    }
    if (b1 >= 0) {
        im1 = comInvMass[b1].w; I1.c[0] = mk3(invIW[3 * b1]); I1.c[1] = mk3(invIW[3 * b1 + 1]); I1.c[2] = mk3(invIW[3 * b1 + 2]);
        float4 l = pseudoLin[b1]; pv1 = mk3(l); cnt1 = __float_as_int(l.w); pw1 = mk3(pseudoAng[b1]);
    }
    for (int r = 0; r < n; ++r) {
        int flags = jointRowFlags(type, r, prm0, st0);
        const Row& R = rows[r];
        V3 a0t = mul(I0, R.a0), a1t = mul(I1, R.a1);
        float k = dot(R.a0, a0t) + dot(R.a1, a1t);
        V3 l0t = mk3(0.f), l1t = mk3(0.f);
        if (!(flags & JF_ANGULAR)) {
            k += dot(R.lin, R.lin) * (im0 + im1);
            l0t = im0 * R.lin; l1t = im1 * R.lin;
        }
        int idx = r * J.n + j;
        J.linC[idx] = f4(R.lin, R.c);
        J.a0T[idx] = f4(R.a0, R.target);
        J.a1K[idx] = f4(R.a1, k);
        J.a0tMin[idx] = f4(a0t, R.mn);
        J.a1tMax[idx] = f4(a1t, R.mx);
        J.soft[idx] = make_float2(R.freq, R.damp);
        if ((flags & JF_SOFT) || !doNgs) continue;
        if (R.c != 0.f && k != 0.f) {            // NGS correction, masked per lane (quirk Q10)
            float lambda = R.c / k;
            if (flags & JF_LIMITED) lambda = fminf(fmaxf(lambda, R.mn), R.mx);
            if (!(flags & JF_ANGULAR)) { pv0 += lambda * l0t; pv1 -= lambda * l1t; }
            pw0 += lambda * a0t; pw1 -= lambda * a1t;
            if (b0 >= 0) ++cnt0;
            if (b1 >= 0) ++cnt1;
        }
    }
    if (!doNgs) return;
    if (b0 >= 0) { pseudoLin[b0] = make_float4(pv0.x, pv0.y, pv0.z, __int_as_float(cnt0)); pseudoAng[b0] = f4(pw0); }
    if (b1 >= 0) { pseudoLin[b1] = make_float4(pv1.x, pv1.y, pv1.z, __int_as_float(cnt1)); pseudoAng[b1] = f4(pw1); }
}

// ---- overflow bucket (joint colour index 8): the reference's scalar, strictly sequential path ----------------------------
// Joints of the bucket may share bodies, so rows are visited exactly in the reference's order: flag list by flag list
// (NONE, ANGULAR, SOFT, LIMITED, ANGULAR|SOFT, ANGULAR|LIMITED -- Constraint1DContainer.h:117), joints in creation
// order inside a list, rows in creation order inside a joint.  Arithmetic is Constraint1D.cpp's, which differs from the
// SIMD path (quirk Q9): the warm start only perturbs the local velocity copy, LIMITED clamps to [min, max] without the
// time step, a row with invEffMass == 0 is skipped before the warm start, and the soft / bias terms divide by invEffMass.
__device__ __forceinline__ int flagList(int flags) {
    switch (flags) { case 0: return 0; case JF_ANGULAR: return 1; case JF_SOFT: return 2; case JF_LIMITED: return 3;
                     case JF_ANGULAR | JF_SOFT: return 4; default: return 5; }
}

// preSolve NGS pass of the bucket (Constraint1D.cpp:31-51); rows were filled by k_joint_prep(doNgs = 0)
__global__ void k_joint_ngs_seq(JointDev J, int start, int count, const int* __restrict__ kinematic, const float4* __restrict__ comInvMass,
                                float4* __restrict__ pseudoLin, float4* __restrict__ pseudoAng) {
    if (blockIdx.x || threadIdx.x) return;
    for (int list = 0; list < 6; ++list) {
        if (list == 2 || list == 4) continue;   // SOFT lists return before the correction
        for (int j = start; j < start + count; ++j) {
            int type = J.type[j];
            float4 prm0 = J.prm[2 * j], st0 = J.state[2 * j];
            int n = jointRowCount(type, prm0, st0);
            int2 rr = J.rows[j];
            int b0 = jSolverIndex(rr.x, J.nDyn, kinematic), b1 = jSolverIndex(rr.y, J.nDyn, kinematic);
            float im0 = b0 >= 0 ? comInvMass[b0].w : 0.f, im1 = b1 >= 0 ? comInvMass[b1].w : 0.f;
            for (int r = 0; r < n; ++r) {
                int flags = jointRowFlags(type, r, prm0, st0);
                if (flagList(flags) != list) continue;
                int idx = r * J.n + j;
                float4 LC = J.linC[idx], A1 = J.a1K[idx], A0t = J.a0tMin[idx], A1t = J.a1tMax[idx];
                float c = LC.w, k = A1.w;
                if (c == 0.f || k == 0.f) continue;
                float lambda = c / k;
                if (flags & JF_LIMITED) lambda = gclamp(lambda, A0t.w, A1t.w);
                V3 lin = mk3(LC);
                if (b0 >= 0) {
                    float4 l = pseudoLin[b0]; V3 pv = mk3(l); int cnt = __float_as_int(l.w);
                    if (!(flags & JF_ANGULAR)) pv += lambda * (im0 * lin);
                    pseudoLin[b0] = make_float4(pv.x, pv.y, pv.z, __int_as_float(cnt + 1));
                    pseudoAng[b0] = f4(mk3(pseudoAng[b0]) + lambda * mk3(A0t));
                }
                if (b1 >= 0) {
                    float4 l = pseudoLin[b1]; V3 pv = mk3(l); int cnt = __float_as_int(l.w);
                    if (!(flags & JF_ANGULAR)) pv -= lambda * (im1 * lin);
                    pseudoLin[b1] = make_float4(pv.x, pv.y, pv.z, __int_as_float(cnt + 1));
                    pseudoAng[b1] = f4(mk3(pseudoAng[b1]) - lambda * mk3(A1t));
                }
            }
        }
    }
}

// solve pass of the bucket (Constraint1D.cpp:55-127)
__global__ void k_joint_solve_seq(JointDev J, int start, int count, float h, int warmStart, const int* __restrict__ kinematic,
                                  const float4* __restrict__ comInvMass, float4* __restrict__ velLive, float4* __restrict__ angvelLive) {
    if (blockIdx.x || threadIdx.x) return;
    for (int list = 0; list < 6; ++list) {
        for (int j = start; j < start + count; ++j) {
            int type = J.type[j];
            float4 prm0 = J.prm[2 * j], st0 = J.state[2 * j];
            int n = jointRowCount(type, prm0, st0);
            int2 rr = J.rows[j];
            int b0 = jSolverIndex(rr.x, J.nDyn, kinematic), b1 = jSolverIndex(rr.y, J.nDyn, kinematic);
            float im0 = b0 >= 0 ? comInvMass[b0].w : 0.f, im1 = b1 >= 0 ? comInvMass[b1].w : 0.f;
            for (int r = 0; r < n; ++r) {
                int flags = jointRowFlags(type, r, prm0, st0);
                if (flagList(flags) != list) continue;
                int idx = r * J.n + j;
                float4 LC = J.linC[idx], A0 = J.a0T[idx], A1 = J.a1K[idx], A0t = J.a0tMin[idx], A1t = J.a1tMax[idx];
                float c = LC.w, k = A1.w;
                if (k == 0.f) continue;
                V3 lin = mk3(LC), a0 = mk3(A0), a1 = mk3(A1), a0t = mk3(A0t), a1t = mk3(A1t);
                V3 l0t = im0 * lin, l1t = im1 * lin;
                const bool ang = (flags & JF_ANGULAR) != 0;
                V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
                if (b0 >= 0) { if (!ang) v0 = mk3(velLive[b0]); w0 = mk3(angvelLive[b0]); }
                if (b1 >= 0) { if (!ang) v1 = mk3(velLive[b1]); w1 = mk3(angvelLive[b1]); }
                float total = J.lambda[idx];
                if (!(flags & JF_SOFT)) {
                    if (warmStart && !((double)fabsf(c) > 1e-4 || fabsf(total) > 10000.f)) {
                        total = total * 0.5f;
                        if (b0 >= 0) { if (!ang) v0 += total * l0t; w0 += total * a0t; }
                        if (b1 >= 0) { if (!ang) v1 -= total * l1t; w1 -= total * a1t; }
                    }
                }
                float rel = dot(a1, w1) - dot(a0, w0);
                if (!ang) rel += dot(lin, v1) - dot(lin, v0);
                float lambda;
                if (flags & JF_SOFT) {
                    float2 sf = J.soft[idx];
                    float af = 2.f * 3.14159265358979323846f * sf.x;
                    float stiffness = af * af / k;
                    float damping = 2.f * af * sf.y / k;
                    float gamma = 1.f / (damping + h * stiffness);
                    float beta = h * stiffness / (damping + h * stiffness);
                    lambda = (rel + beta * c / h) / (k + gamma / h);
                } else {
                    lambda = (rel - A0.w + 0.2f * c / h) / k;
                }
                if (flags & JF_LIMITED) {
                    float prev = total;
                    total += lambda;
                    total = gclamp(total, A0t.w, A1t.w);
                    lambda = total - prev;
                } else total += lambda;
                J.lambda[idx] = total;
                if (b0 >= 0) {
                    if (!ang) velLive[b0] = f4(mk3(velLive[b0]) + lambda * l0t);
                    angvelLive[b0] = f4(mk3(angvelLive[b0]) + lambda * a0t);
                }
                if (b1 >= 0) {
                    if (!ang) velLive[b1] = f4(mk3(velLive[b1]) - lambda * l1t);
                    angvelLive[b1] = f4(mk3(angvelLive[b1]) - lambda * a1t);
                }
            }
        }
    }
}

// per iteration and colour (Constraint1DW.cpp:118-233)
__global__ void k_joint_solve(JointDev J, int start, int count, float h, int warmStart, const int* __restrict__ kinematic,
                              const float4* __restrict__ comInvMass, float4* __restrict__ velLive, float4* __restrict__ angvelLive) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int j = start + i;
    int type = J.type[j];
    int2 rr = J.rows[j];
    int b0 = jSolverIndex(rr.x, J.nDyn, kinematic), b1 = jSolverIndex(rr.y, J.nDyn, kinematic);
    float4 prm0 = J.prm[2 * j], st0 = J.state[2 * j];
    int n = jointRowCount(type, prm0, st0);
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (b0 >= 0) { v0 = mk3(velLive[b0]); w0 = mk3(angvelLive[b0]); im0 = comInvMass[b0].w; }
    if (b1 >= 0) { v1 = mk3(velLive[b1]); w1 = mk3(angvelLive[b1]); im1 = comInvMass[b1].w; }
    const float biasFactor = (float)(0.2 / (double)h);
    for (int r = 0; r < n; ++r) {
        int flags = jointRowFlags(type, r, prm0, st0);
        int idx = r * J.n + j;
        float4 LC = J.linC[idx], A0 = J.a0T[idx], A1 = J.a1K[idx], A0t = J.a0tMin[idx], A1t = J.a1tMax[idx];
        V3 lin = mk3(LC), a0 = mk3(A0), a1 = mk3(A1), a0t = mk3(A0t), a1t = mk3(A1t);
        float c = LC.w, k = A1.w;
        V3 l0t = im0 * lin, l1t = im1 * lin;
        float total = J.lambda[idx];
        // velocities as this row sees them: a row whose invEffMass is 0 never writes back (Constraint1DW.cpp:215-216),
        // so its warm-start perturbation must not leak into the registers
        V3 rv0 = v0, rw0 = w0, rv1 = v1, rw1 = w1;
        if (!(flags & JF_SOFT) && warmStart) {
            if (fabsf(c) < (float)1e-4 && fabsf(total) < 10000.f) {
                float lam = total * 0.5f;
                if (!(flags & JF_ANGULAR)) { rv0 += lam * l0t; rv1 -= lam * l1t; }
                rw0 += lam * a0t; rw1 -= lam * a1t;
                total = lam;
            }
        }
        if (k == 0.f) { J.lambda[idx] = total; continue; }
        float rel = dot(a1, rw1) - dot(a0, rw0);
        if (!(flags & JF_ANGULAR)) rel += dot(lin, rv1) - dot(lin, rv0);
        float effMass = 1.f / k;
        float lambda;
        if (flags & JF_SOFT) {
            float2 s = J.soft[idx];
            float af = (2.f * 3.14159265358979323846f) * s.x;
            float stiffness = af * af * effMass;
            float damping = 2.f * af * s.y * effMass;
            float gamma = 1.f / (damping + h * stiffness);
            float beta = h * stiffness * gamma;
            lambda = (rel + beta * c / h) / (k + gamma / h);
        } else {
            lambda = (rel - A0.w + biasFactor * c) * effMass;
        }
        float prev = total;
        if (flags & JF_LIMITED) {
            total += lambda;
            total = fminf(fmaxf(total, A0t.w * h), A1t.w * h);
            lambda = total - prev;
        } else total += lambda;
        J.lambda[idx] = total;
        if (!(flags & JF_ANGULAR)) { rv0 += lambda * l0t; rv1 -= lambda * l1t; }
        rw0 += lambda * a0t; rw1 -= lambda * a1t;
        if (!(flags & JF_ANGULAR)) { v0 = rv0; v1 = rv1; }
        w0 = rw0; w1 = rw1;
    }
    if (b0 >= 0) { velLive[b0] = f4(v0); angvelLive[b0] = f4(w0); }
    if (b1 >= 0) { velLive[b1] = f4(v1); angvelLive[b1] = f4(w1); }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct JointStore {
    int n = 0;
    int* type = nullptr; int2* rows = nullptr; float4* a0p = nullptr; float4* a0q = nullptr; float4* a1p = nullptr; float4* a1q = nullptr;
    float4* prm = nullptr; float4* state = nullptr;
    float4* linC = nullptr; float4* a0T = nullptr; float4* a1K = nullptr; float4* a0tMin = nullptr; float4* a1tMax = nullptr;
    float2* soft = nullptr; float* lambda = nullptr;
    std::vector<int> order;   // device slot -> caller's joint index
    // persistent state of the previous upload (gear angles), kept so surviving joints can carry it over
    std::vector<float4> prevState; std::vector<int> prevOrder;
};
static JointStore* store(pb_ctx* ctx) { return (JointStore*)ctx->jointStore; }

void pb_joints_free(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    if (!s) return;
    cudaFree(s->type); cudaFree(s->rows); cudaFree(s->a0p); cudaFree(s->a0q); cudaFree(s->a1p); cudaFree(s->a1q); cudaFree(s->prm); cudaFree(s->state);
    cudaFree(s->linC); cudaFree(s->a0T); cudaFree(s->a1K); cudaFree(s->a0tMin); cudaFree(s->a1tMax); cudaFree(s->soft); cudaFree(s->lambda);
    delete s;
    ctx->jointStore = nullptr;
    ctx->nJoints = 0;
}

static JointDev devView(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    JointDev J;
    J.n = s->n; J.nDyn = ctx->nDyn; J.type = s->type; J.rows = s->rows; J.a0p = s->a0p; J.a0q = s->a0q; J.a1p = s->a1p; J.a1q = s->a1q;
    J.prm = s->prm; J.state = s->state; J.linC = s->linC; J.a0T = s->a0T; J.a1K = s->a1K; J.a0tMin = s->a0tMin; J.a1tMax = s->a1tMax;
    J.soft = s->soft; J.lambda = s->lambda;
    return J;
}

int pb_joints_upload(pb_ctx* ctx, int n, const int* type, const int* row0, const int* row1, const float* a0p, const float* a0q,
                     const float* a1p, const float* a1q, const float* params8, const int* color) {
    std::vector<float4> prevState; std::vector<int> prevOrder;
    if (JointStore* old = store(ctx)) {
        cudaStreamSynchronize(ctx->stream);
        prevState.resize(2 * (size_t)old->n);
        cudaMemcpy(prevState.data(), old->state, sizeof(float4) * 2 * old->n, cudaMemcpyDeviceToHost);
        prevOrder = old->order;
    }
    pb_joints_free(ctx);
    if (n == 0) return PB_OK;
    for (int j = 0; j < n; ++j) {
        if (color[j] < 0 || color[j] > 8) return pb_fail(ctx, PB_EINVAL, "joint colour out of range");
        if (type[j] < 0 || type[j] > PB_JOINT_SERVO) return pb_fail(ctx, PB_EINVAL, "joint type");
    }
    JointStore* s = new JointStore();
    ctx->jointStore = s;
    s->n = n;
    s->prevState.swap(prevState); s->prevOrder.swap(prevOrder);
    // colour-major order, stable inside a colour (== the reference's per-colour joint vectors)
    s->order.resize(n);
    for (int j = 0; j < n; ++j) s->order[j] = j;
    std::stable_sort(s->order.begin(), s->order.end(), [&](int a, int b) { return color[a] < color[b]; });
    for (int c = 0; c <= PB_JOINT_COLORS; ++c) ctx->jointColorStart[c] = 0;
    for (int j = 0; j < n; ++j) ctx->jointColorStart[color[j] + 1]++;
    for (int c = 0; c < PB_JOINT_COLORS; ++c) ctx->jointColorStart[c + 1] += ctx->jointColorStart[c];
    std::vector<int> t(n); std::vector<int2> rr(n); std::vector<float4> p0(n), q0(n), p1(n), q1(n), prm(2 * (size_t)n), st(2 * (size_t)n, make_float4(0, 0, 0, 0));
    for (int k = 0; k < n; ++k) {
        int j = s->order[k];
        t[k] = type[j]; rr[k] = make_int2(row0[j], row1[j]);
        p0[k] = make_float4(a0p[3 * j], a0p[3 * j + 1], a0p[3 * j + 2], 0.f);
        q0[k] = make_float4(a0q[4 * j], a0q[4 * j + 1], a0q[4 * j + 2], a0q[4 * j + 3]);
        p1[k] = make_float4(a1p[3 * j], a1p[3 * j + 1], a1p[3 * j + 2], 0.f);
        q1[k] = make_float4(a1q[4 * j], a1q[4 * j + 1], a1q[4 * j + 2], a1q[4 * j + 3]);
        const float* P = params8 + 8 * (size_t)j;
        prm[2 * k] = make_float4(P[0], P[1], P[2], P[3]);
        prm[2 * k + 1] = make_float4(P[4], P[5], P[6], P[7]);
    }
    int rc = 0;
    size_t R = (size_t)MAXR * n;
#define A(p, cnt) if (!rc) rc = pb_alloc(ctx, &s->p, (cnt))
    A(type, n); A(rows, n); A(a0p, n); A(a0q, n); A(a1p, n); A(a1q, n); A(prm, 2 * (size_t)n); A(state, 2 * (size_t)n);
    A(linC, R); A(a0T, R); A(a1K, R); A(a0tMin, R); A(a1tMax, R); A(soft, R); A(lambda, R);
#undef A
    if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpy(s->type, t.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->rows, rr.data(), sizeof(int2) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a0p, p0.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a0q, q0.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a1p, p1.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a1q, q1.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->prm, prm.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->state, st.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemset(s->lambda, 0, sizeof(float) * R));
    ctx->nJoints = n;
    return PB_OK;
}

// setters called on live joints (RevoluteJoint::setDriveVelocity, ServoJoint::setTargetAngle, ...): new params in the
// caller's joint order; persistent joint state (gear angles, accumulated impulses) is kept
int pb_joints_update_params(pb_ctx* ctx, int n, const float* params8) {
    JointStore* s = store(ctx);
    if (!s || n != s->n) return pb_fail(ctx, PB_EINVAL, "pb_update_joint_params: joint count mismatch");
    std::vector<float4> prm(2 * (size_t)n);
    for (int k = 0; k < n; ++k) {
        const float* P = params8 + 8 * (size_t)s->order[k];
        prm[2 * k] = make_float4(P[0], P[1], P[2], P[3]);
        prm[2 * k + 1] = make_float4(P[4], P[5], P[6], P[7]);
    }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(s->prm, prm.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    return PB_OK;
}

// oldIndex[j] = index the caller's joint j had in the PREVIOUS pb_upload_joints call, or -1 for a new joint
int pb_joints_keep_state(pb_ctx* ctx, int n, const int* oldIndex) {
    JointStore* s = store(ctx);
    if (!s || n != s->n) return n == 0 ? PB_OK : pb_fail(ctx, PB_EINVAL, "pb_keep_joint_state: joint count mismatch");
    if (s->prevOrder.empty()) return PB_OK;
    std::vector<int> oldSlot(s->prevOrder.size(), -1);
    for (size_t k = 0; k < s->prevOrder.size(); ++k) oldSlot[s->prevOrder[k]] = (int)k;
    std::vector<float4> st(2 * (size_t)n);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(st.data(), s->state, sizeof(float4) * 2 * n, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; ++k) {
        int oj = oldIndex[s->order[k]];
        if (oj < 0 || oj >= (int)oldSlot.size() || oldSlot[oj] < 0) continue;
        st[2 * k] = s->prevState[2 * (size_t)oldSlot[oj]];
        st[2 * k + 1] = s->prevState[2 * (size_t)oldSlot[oj] + 1];
    }
    PB_CUDA(ctx, cudaMemcpy(s->state, st.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    return PB_OK;
}

int pb_joint_begin_step(pb_ctx* ctx) {
    if (!ctx->nJoints) return PB_OK;
    JointDev J = devView(ctx);
    ++ctx->launches, k_joint_begin<<<pb_grid(J.n, 128), 128, 0, ctx->stream>>>(J, ctx->pos, ctx->quat);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_joint_prep(pb_ctx* ctx, float h) {
    JointDev J = devView(ctx);
    pb_prof_begin(ctx, PROF_JOINTS);
    for (int c = 0; c < 8; ++c) {
        int start = ctx->jointColorStart[c], count = ctx->jointColorStart[c + 1] - start;
        if (count <= 0) continue;
        ++ctx->launches, k_joint_prep<<<pb_grid(count, 128), 128, 0, ctx->stream>>>(J, start, count, 1, ctx->kinematic, ctx->pos, ctx->quat, ctx->comInvMass,
                                                                                     ctx->invIW, ctx->pseudoLin, ctx->pseudoAng);
    }
    {   // overflow bucket: parallel row fill, sequential NGS pass
        int start = ctx->jointColorStart[8], count = ctx->jointColorStart[9] - start;
        if (count > 0) {
            ++ctx->launches, k_joint_prep<<<pb_grid(count, 128), 128, 0, ctx->stream>>>(J, start, count, 0, ctx->kinematic, ctx->pos, ctx->quat, ctx->comInvMass,
                                                                                         ctx->invIW, ctx->pseudoLin, ctx->pseudoAng);
            ++ctx->launches, k_joint_ngs_seq<<<1, 32, 0, ctx->stream>>>(J, start, count, ctx->kinematic, ctx->comInvMass, ctx->pseudoLin, ctx->pseudoAng);
        }
    }
    pb_prof_end(ctx);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_joint_solve(pb_ctx* ctx, float h, int warmStart) {
    JointDev J = devView(ctx);
    pb_prof_begin(ctx, PROF_JOINTS);
    for (int c = 0; c < 8; ++c) {
        int start = ctx->jointColorStart[c], count = ctx->jointColorStart[c + 1] - start;
        if (count <= 0) continue;
        ++ctx->launches, k_joint_solve<<<pb_grid(count, 128), 128, 0, ctx->stream>>>(J, start, count, h, warmStart, ctx->kinematic, ctx->comInvMass,
                                                                                      ctx->velLive, ctx->angvelLive);
    }
    {
        int start = ctx->jointColorStart[8], count = ctx->jointColorStart[9] - start;
        if (count > 0)
            ++ctx->launches, k_joint_solve_seq<<<1, 32, 0, ctx->stream>>>(J, start, count, h, warmStart, ctx->kinematic, ctx->comInvMass, ctx->velLive, ctx->angvelLive);
    }
    pb_prof_end(ctx);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

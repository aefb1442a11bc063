// Joints on the device (device functions shared by joints.cu and the persistent substep kernel in solver.cu): per-step layout, per-substep row fill + preSolve (NGS pseudo velocities), per-iteration solve.
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
#pragma once
#include "pb_ctx.h"
#include "pb_math.cuh"

// L1 = true: the caller is a per-CTA island sweep (solver.cu): everything it touches is private to its CTA for the whole kernel, so
// loads may hit L1 (a dependent chain then costs ~40 ns per link instead of an L2 round trip); otherwise L2 (.cg): other SMs write
// the arrays between grid barriers and L1 is not coherent across SMs.
template <bool L1, class T> __device__ __forceinline__ T ldm(const T* p) { return L1 ? *p : __ldcg(p); }

#define JF_SOFT 1
#define JF_ANGULAR 2
#define JF_LIMITED 4
#define MAXR PB_MAX_JOINT_ROWS

struct JointDev {
    int n; int nDyn;
    const int* type; const int2* rows; const int2* bodies;   // bodies: solver index of each side or -1, resolved once per step
    const float4* a0p; const float4* a0q; const float4* a1p; const float4* a1q;
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
__device__ inline int buildJointRows(int type, float4 prm0, float4 prm1, float4& st0, float4& st1, bool advanceState,
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

// per substep and colour: fill rows (makeConstraints), effective masses, NGS pseudo-velocity pass (Constraint1DW.cpp:6-115)
// All arrays a step mutates (poses, pseudo velocities, world inertia, joint rows / state / impulses) are read with __ldcg:
// inside the persistent substep kernel they are written by other SMs between grid barriers, and L1 is not coherent.
// L1: see ldm() above (the whole-step kernel's group-by-group form)
template <bool L1 = false>
__device__ inline void jointPrepOne(const JointDev& J, int j, int doNgs, const int* __restrict__ kinematic,
                                    const float4* pos, const float4* quat, const float4* __restrict__ comInvMass,
                                    const float4* bodyRec, float4* pseudoLin, float4* pseudoAng) {
    int type = J.type[j];
    int2 rr = J.rows[j];
    int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
    Q4 q0 = mkq(ldm<L1>(&quat[rr.x])), q1 = mkq(ldm<L1>(&quat[rr.y]));
    V3 a0p = mk3(J.a0p[j]), a1p = mk3(J.a1p[j]);
    // JointSolverData r0/r1 (Physecs.cpp:344-345) and calculateWorldSpaceData (Joint.cpp:5-14)
    V3 r0l = b0 >= 0 ? a0p - mk3(comInvMass[b0]) : a0p;
    V3 r1l = b1 >= 0 ? a1p - mk3(comInvMass[b1]) : a1p;
    M3 u0 = mat3_cast(qmul(q0, mkq(J.a0q[j]))), u1 = mat3_cast(qmul(q1, mkq(J.a1q[j])));
    V3 r0 = rotate(q0, r0l), r1 = rotate(q1, r1l);
    V3 p0 = mk3(ldm<L1>(&pos[rr.x])) + rotate(q0, a0p), p1 = mk3(ldm<L1>(&pos[rr.y])) + rotate(q1, a1p);
    float4 prm0 = J.prm[2 * j], prm1 = J.prm[2 * j + 1];
    float4 st0 = ldm<L1>(&J.state[2 * j]), st1 = ldm<L1>(&J.state[2 * j + 1]);
    Row rows[MAXR];
    int n = buildJointRows(type, prm0, prm1, st0, st1, true, p0, p1, r0, r1, u0, u1, rows);
    if (type == PB_JOINT_GEAR) { J.state[2 * j] = st0; J.state[2 * j + 1] = st1; }
    float im0 = 0.f, im1 = 0.f;
    M3 I0, I1; I0.c[0] = I0.c[1] = I0.c[2] = mk3(0.f); I1 = I0;
    V3 pv0 = mk3(0.f), pw0 = mk3(0.f), pv1 = mk3(0.f), pw1 = mk3(0.f);
    int cnt0 = 0, cnt1 = 0;
    if (b0 >= 0) {
        im0 = comInvMass[b0].w;
        {   // world inverse inertia from the per-substep body record (solver.cu, BodyRec layout)
            const float4* r = bodyRec + 8 * (size_t)b0; float4 r2 = ldm<L1>(r + 2), r6 = ldm<L1>(r + 6), r7 = ldm<L1>(r + 7);
            I0.c[0] = mk3(r6); I0.c[1] = mk3(r6.w, r7.x, r7.y); I0.c[2] = mk3(r7.z, r7.w, r2.w);
        }
        float4 l = ldm<L1>(&pseudoLin[b0]); pv0 = mk3(l); cnt0 = __float_as_int(l.w); pw0 = mk3(ldm<L1>(&pseudoAng[b0]));
    }
    if (b1 >= 0) {
        im1 = comInvMass[b1].w;
        {
            const float4* r = bodyRec + 8 * (size_t)b1; float4 r2 = ldm<L1>(r + 2), r6 = ldm<L1>(r + 6), r7 = ldm<L1>(r + 7);
            I1.c[0] = mk3(r6); I1.c[1] = mk3(r6.w, r7.x, r7.y); I1.c[2] = mk3(r7.z, r7.w, r2.w);
        }
        float4 l = ldm<L1>(&pseudoLin[b1]); pv1 = mk3(l); cnt1 = __float_as_int(l.w); pw1 = mk3(ldm<L1>(&pseudoAng[b1]));
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

// NGS pseudo-velocity pass of one joint from the rows jointPrepOne(doNgs = 0) stored (Constraint1DW.cpp:57-115): the same
// lambda = c / k accumulation, row by row, as the fused version above.  Splitting it off lets the expensive row fill run once
// for all joints in a plain kernel while only this light pass is ordered by joint colour.
template <bool L1>
__device__ inline void jointNgsOne(const JointDev& J, int j, const int* __restrict__ kinematic, const float4* __restrict__ comInvMass,
                                   float4* pseudoLin, float4* pseudoAng) {
    int type = J.type[j];
    int2 rr = J.rows[j];
    int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
    float4 prm0 = J.prm[2 * j], st0 = ldm<L1>(&J.state[2 * j]);
    int n = jointRowCount(type, prm0, st0);
    float im0 = 0.f, im1 = 0.f;
    V3 pv0 = mk3(0.f), pw0 = mk3(0.f), pv1 = mk3(0.f), pw1 = mk3(0.f);
    int cnt0 = 0, cnt1 = 0;
    if (b0 >= 0) { im0 = comInvMass[b0].w; float4 l = ldm<L1>(&pseudoLin[b0]); pv0 = mk3(l); cnt0 = __float_as_int(l.w); pw0 = mk3(ldm<L1>(&pseudoAng[b0])); }
    if (b1 >= 0) { im1 = comInvMass[b1].w; float4 l = ldm<L1>(&pseudoLin[b1]); pv1 = mk3(l); cnt1 = __float_as_int(l.w); pw1 = mk3(ldm<L1>(&pseudoAng[b1])); }
    for (int r = 0; r < n; ++r) {
        int flags = jointRowFlags(type, r, prm0, st0);
        if (flags & JF_SOFT) continue;
        int idx = r * J.n + j;
        float4 LC = ldm<L1>(&J.linC[idx]), A1 = ldm<L1>(&J.a1K[idx]), A0t = ldm<L1>(&J.a0tMin[idx]), A1t = ldm<L1>(&J.a1tMax[idx]);
        float c = LC.w, k = A1.w;
        if (c != 0.f && k != 0.f) {            // masked per lane in the reference (quirk Q10)
            float lambda = c / k;
            if (flags & JF_LIMITED) lambda = fminf(fmaxf(lambda, A0t.w), A1t.w);
            V3 lin = mk3(LC);
            if (!(flags & JF_ANGULAR)) { pv0 += lambda * (im0 * lin); pv1 -= lambda * (im1 * lin); }
            pw0 += lambda * mk3(A0t); pw1 -= lambda * mk3(A1t);
            if (b0 >= 0) ++cnt0;
            if (b1 >= 0) ++cnt1;
        }
    }
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
__device__ inline void jointNgsSeq(const JointDev& J, int start, int count, const int* __restrict__ kinematic, const float4* __restrict__ comInvMass,
                                   float4* pseudoLin, float4* pseudoAng) {
    for (int list = 0; list < 6; ++list) {
        if (list == 2 || list == 4) continue;   // SOFT lists return before the correction
        for (int j = start; j < start + count; ++j) {
            int type = J.type[j];
            float4 prm0 = J.prm[2 * j], st0 = __ldcg(&J.state[2 * j]);
            int n = jointRowCount(type, prm0, st0);
            int2 rr = J.rows[j];
            int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
            float im0 = b0 >= 0 ? comInvMass[b0].w : 0.f, im1 = b1 >= 0 ? comInvMass[b1].w : 0.f;
            for (int r = 0; r < n; ++r) {
                int flags = jointRowFlags(type, r, prm0, st0);
                if (flagList(flags) != list) continue;
                int idx = r * J.n + j;
                float4 LC = __ldcg(&J.linC[idx]), A1 = __ldcg(&J.a1K[idx]), A0t = __ldcg(&J.a0tMin[idx]), A1t = __ldcg(&J.a1tMax[idx]);
                float c = LC.w, k = A1.w;
                if (c == 0.f || k == 0.f) continue;
                float lambda = c / k;
                if (flags & JF_LIMITED) lambda = gclamp(lambda, A0t.w, A1t.w);
                V3 lin = mk3(LC);
                if (b0 >= 0) {
                    float4 l = __ldcg(&pseudoLin[b0]); V3 pv = mk3(l); int cnt = __float_as_int(l.w);
                    if (!(flags & JF_ANGULAR)) pv += lambda * (im0 * lin);
                    pseudoLin[b0] = make_float4(pv.x, pv.y, pv.z, __int_as_float(cnt + 1));
                    pseudoAng[b0] = f4(mk3(__ldcg(&pseudoAng[b0])) + lambda * mk3(A0t));
                }
                if (b1 >= 0) {
                    float4 l = __ldcg(&pseudoLin[b1]); V3 pv = mk3(l); int cnt = __float_as_int(l.w);
                    if (!(flags & JF_ANGULAR)) pv -= lambda * (im1 * lin);
                    pseudoLin[b1] = make_float4(pv.x, pv.y, pv.z, __int_as_float(cnt + 1));
                    pseudoAng[b1] = f4(mk3(__ldcg(&pseudoAng[b1])) - lambda * mk3(A1t));
                }
            }
        }
    }
}

// solve pass of the bucket (Constraint1D.cpp:55-127)
__device__ inline void jointSolveSeq(const JointDev& J, int start, int count, float h, int warmStart, const int* __restrict__ kinematic,
                                     const float4* __restrict__ comInvMass, float4* velLive, float4* angvelLive) {
    for (int list = 0; list < 6; ++list) {
        for (int j = start; j < start + count; ++j) {
            int type = J.type[j];
            float4 prm0 = J.prm[2 * j], st0 = __ldcg(&J.state[2 * j]);
            int n = jointRowCount(type, prm0, st0);
            int2 rr = J.rows[j];
            int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
            float im0 = b0 >= 0 ? comInvMass[b0].w : 0.f, im1 = b1 >= 0 ? comInvMass[b1].w : 0.f;
            for (int r = 0; r < n; ++r) {
                int flags = jointRowFlags(type, r, prm0, st0);
                if (flagList(flags) != list) continue;
                int idx = r * J.n + j;
                float4 LC = __ldcg(&J.linC[idx]), A0 = __ldcg(&J.a0T[idx]), A1 = __ldcg(&J.a1K[idx]), A0t = __ldcg(&J.a0tMin[idx]), A1t = __ldcg(&J.a1tMax[idx]);
                float c = LC.w, k = A1.w;
                if (k == 0.f) continue;
                V3 lin = mk3(LC), a0 = mk3(A0), a1 = mk3(A1), a0t = mk3(A0t), a1t = mk3(A1t);
                V3 l0t = im0 * lin, l1t = im1 * lin;
                const bool ang = (flags & JF_ANGULAR) != 0;
                V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
                if (b0 >= 0) { if (!ang) v0 = mk3(__ldcg(&velLive[2 * b0])); w0 = mk3(__ldcg(&angvelLive[2 * b0])); }
                if (b1 >= 0) { if (!ang) v1 = mk3(__ldcg(&velLive[2 * b1])); w1 = mk3(__ldcg(&angvelLive[2 * b1])); }
                float total = __ldcg(&J.lambda[idx]);
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
                    float2 sf = __ldcg(&J.soft[idx]);
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
                    if (!ang) velLive[2 * b0] = f4(mk3(__ldcg(&velLive[2 * b0])) + lambda * l0t, im0);
                    angvelLive[2 * b0] = f4(mk3(__ldcg(&angvelLive[2 * b0])) + lambda * a0t);
                }
                if (b1 >= 0) {
                    if (!ang) velLive[2 * b1] = f4(mk3(__ldcg(&velLive[2 * b1])) - lambda * l1t, im1);
                    angvelLive[2 * b1] = f4(mk3(__ldcg(&angvelLive[2 * b1])) - lambda * a1t);
                }
            }
        }
    }
}

// One joint row of the SIMD-path solve (Constraint1DW.cpp:118-233): reads / updates the joint's running body velocities and
// the row's accumulated impulse.  Returns through `total`; `v0..w1` are only committed when the row has an effective mass.
__device__ __forceinline__ void jointRowSolve(int flags, float4 LC, float4 A0, float4 A1, float4 A0t, float4 A1t, float2 softp, float& total,
                                              float h, float biasFactor, int warmStart, float im0, float im1, V3& v0, V3& w0, V3& v1, V3& w1) {
    V3 lin = mk3(LC), a0 = mk3(A0), a1 = mk3(A1), a0t = mk3(A0t), a1t = mk3(A1t);
    float c = LC.w, k = A1.w;
    V3 l0t = im0 * lin, l1t = im1 * lin;
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
    if (k == 0.f) return;
    float rel = dot(a1, rw1) - dot(a0, rw0);
    if (!(flags & JF_ANGULAR)) rel += dot(lin, rv1) - dot(lin, rv0);
    float effMass = 1.f / k;
    float lambda;
    if (flags & JF_SOFT) {
        float af = (2.f * 3.14159265358979323846f) * softp.x;
        float stiffness = af * af * effMass;
        float damping = 2.f * af * softp.y * effMass;
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
    if (!(flags & JF_ANGULAR)) { rv0 += lambda * l0t; rv1 -= lambda * l1t; }
    rw0 += lambda * a0t; rw1 -= lambda * a1t;
    if (!(flags & JF_ANGULAR)) { v0 = rv0; v1 = rv1; }
    w0 = rw0; w1 = rw1;
}

// per iteration and colour, EIGHT lanes per joint: lane r holds row r, so every load of the joint is issued in one wave;
// the rows are then applied in order (the Gauss-Seidel dependence is real) by passing the 12 running velocity components
// from lane to lane with shuffles.  Same arithmetic as the one-thread version below, ~3 memory round trips instead of ~9.
template <bool L1>
__device__ inline void jointSolveOct(const JointDev& J, int j, int lane8, unsigned gmask, float h, int warmStart, const int* __restrict__ kinematic,
                                     const float4* __restrict__ comInvMass, float4* velLive, float4* angvelLive) {
    int type = J.type[j];
    int2 rr = J.rows[j];
    int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
    float4 prm0 = J.prm[2 * j], st0 = ldm<L1>(&J.state[2 * j]);
    int n = jointRowCount(type, prm0, st0);
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (b0 >= 0) { float4 t_ = ldm<L1>(&velLive[2 * b0]); v0 = mk3(t_); im0 = t_.w; w0 = mk3(ldm<L1>(&angvelLive[2 * b0])); }
    if (b1 >= 0) { float4 t_ = ldm<L1>(&velLive[2 * b1]); v1 = mk3(t_); im1 = t_.w; w1 = mk3(ldm<L1>(&angvelLive[2 * b1])); }
    const float biasFactor = (float)(0.2 / (double)h);
    const bool mine = lane8 < n;
    int flags = 0; float total = 0.f;
    float4 LC = make_float4(0, 0, 0, 0), A0 = LC, A1 = LC, A0t = LC, A1t = LC; float2 softp = make_float2(0.f, 0.f);
    int idx = lane8 * J.n + j;
    if (mine) {
        flags = jointRowFlags(type, lane8, prm0, st0);
        LC = ldm<L1>(&J.linC[idx]); A0 = ldm<L1>(&J.a0T[idx]); A1 = ldm<L1>(&J.a1K[idx]); A0t = ldm<L1>(&J.a0tMin[idx]); A1t = ldm<L1>(&J.a1tMax[idx]);
        total = ldm<L1>(&J.lambda[idx]);
        if (flags & JF_SOFT) softp = ldm<L1>(&J.soft[idx]);
    }
    for (int r = 0; r < n; ++r) {
        if (lane8 == r) jointRowSolve(flags, LC, A0, A1, A0t, A1t, softp, total, h, biasFactor, warmStart, im0, im1, v0, w0, v1, w1);
        v0.x = __shfl_sync(gmask, v0.x, r, 8); v0.y = __shfl_sync(gmask, v0.y, r, 8); v0.z = __shfl_sync(gmask, v0.z, r, 8);
        w0.x = __shfl_sync(gmask, w0.x, r, 8); w0.y = __shfl_sync(gmask, w0.y, r, 8); w0.z = __shfl_sync(gmask, w0.z, r, 8);
        v1.x = __shfl_sync(gmask, v1.x, r, 8); v1.y = __shfl_sync(gmask, v1.y, r, 8); v1.z = __shfl_sync(gmask, v1.z, r, 8);
        w1.x = __shfl_sync(gmask, w1.x, r, 8); w1.y = __shfl_sync(gmask, w1.y, r, 8); w1.z = __shfl_sync(gmask, w1.z, r, 8);
    }
    if (mine) J.lambda[idx] = total;
    if (lane8 == 0) {
        if (b0 >= 0) { velLive[2 * b0] = f4(v0, im0); angvelLive[2 * b0] = f4(w0); }
        if (b1 >= 0) { velLive[2 * b1] = f4(v1, im1); angvelLive[2 * b1] = f4(w1); }
    }
}

// per iteration and colour (Constraint1DW.cpp:118-233)
__device__ inline void jointSolveOne(const JointDev& J, int j, float h, int warmStart, const int* __restrict__ kinematic,
                                     const float4* __restrict__ comInvMass, float4* velLive, float4* angvelLive) {
    int type = J.type[j];
    int2 rr = J.rows[j];
    int2 bb_ = J.bodies[j]; int b0 = bb_.x, b1 = bb_.y;
    float4 prm0 = J.prm[2 * j], st0 = __ldcg(&J.state[2 * j]);
    int n = jointRowCount(type, prm0, st0);
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (b0 >= 0) { float4 t_ = __ldcg(&velLive[2 * b0]); v0 = mk3(t_); im0 = t_.w; w0 = mk3(__ldcg(&angvelLive[2 * b0])); }
    if (b1 >= 0) { float4 t_ = __ldcg(&velLive[2 * b1]); v1 = mk3(t_); im1 = t_.w; w1 = mk3(__ldcg(&angvelLive[2 * b1])); }
    const float biasFactor = (float)(0.2 / (double)h);
    for (int r = 0; r < n; ++r) {
        int flags = jointRowFlags(type, r, prm0, st0);
        int idx = r * J.n + j;
        float4 LC = __ldcg(&J.linC[idx]), A0 = __ldcg(&J.a0T[idx]), A1 = __ldcg(&J.a1K[idx]), A0t = __ldcg(&J.a0tMin[idx]), A1t = __ldcg(&J.a1tMax[idx]);
        float total = __ldcg(&J.lambda[idx]);
        float2 softp = (flags & JF_SOFT) ? __ldcg(&J.soft[idx]) : make_float2(0.f, 0.f);
        jointRowSolve(flags, LC, A0, A1, A0t, A1t, softp, total, h, biasFactor, warmStart, im0, im1, v0, w0, v1, w1);
        J.lambda[idx] = total;
    }
    if (b0 >= 0) { velLive[2 * b0] = f4(v0, im0); angvelLive[2 * b0] = f4(w0); }
    if (b1 >= 0) { velLive[2 * b1] = f4(v1, im1); angvelLive[2 * b1] = f4(w1); }
}


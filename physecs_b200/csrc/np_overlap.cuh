// Boolean overlap tests for trigger pairs (reference src/Overlap.cpp:19-235, called from Physecs.cpp:200-207).
//
// A pair whose contact filter answers TRIGGER never reaches collision(); it is tested with physecs::overlap and,
// when overlapping, recorded for the enter / exit diff at the end of the step (Physecs.cpp:538-552).
//   sphere / capsule / box combinations   Overlap.cpp:19-162 (closest-feature distance vs radius sum, 15-axis SAT)
//   anything vs convex mesh               boolean GJK, src/GJK.h:191-224 -- note it differs from the simplex-returning
//                                         variant used by collision(): degenerate directions answer "overlap" at once
//   box vs convex                         Overlap.cpp:198-200 builds the convex support at pos0 (quirk Q22): restated
//   triangle meshes                       no case in the dispatch (Overlap.cpp:203-233) -> never overlap
#pragma once
#include "np_geom.cuh"
#include "np_gjk.cuh"

// GJK<...>::gjk(support0, support1, startDir) without simplex output (GJK.h:191-224)
__device__ inline bool gjkBool(const Shape& s0, const Shape& s1, V3 startDir) {
    GjkV s[4];
    D3 dir = mkd(normalize(startDir));
    s[0] = minkowski(s0, s1, tof(dir));
    dir = -mkd(s[0].pos);
    if (gjkIsZero(tof(dir))) return true;
    if (!checkDirection(s0, s1, dir, s[1])) return false;
    D3 p0 = mkd(s[0].pos), p1 = mkd(s[1].pos);
    dir = dcross(dcross(p0 - p1, -p1), p0 - p1);
    if (gjkIsZero(tof(dir))) return true;
    if (!checkDirection(s0, s1, dir, s[2])) return false;
    D3 p2 = mkd(s[2].pos);
    dir = dcross(p1 - p0, p2 - p0);
    dir = dsign(ddot(-p0, dir)) * dir;
    if (gjkIsZero(tof(dir))) return true;
    if (!checkDirection(s0, s1, dir, s[3])) return false;
    for (int i = 0; i < 100; ++i) {
        D3 a = mkd(s[0].pos), b = mkd(s[1].pos), c = mkd(s[2].pos), d = mkd(s[3].pos);
        int faceIndex = 2;
        bool outside = checkFace(d, a, b, c, dir);
        if (!outside) { faceIndex = 1; outside = checkFace(d, c, a, b, dir); }
        if (!outside) { faceIndex = 0; outside = checkFace(d, b, c, a, dir); }
        if (!outside) return true;
        if (!checkDirection(s0, s1, dir, s[faceIndex])) return false;
        GjkV t = s[faceIndex]; s[faceIndex] = s[3]; s[3] = t;
    }
    return false;
}

__device__ inline bool overlapSphereSphere(V3 pos0, float r0, V3 pos1, float r1) {
    V3 d = pos1 - pos0;
    float rs = r0 + r1;
    return !(dot(d, d) > rs * rs);
}

__device__ inline bool overlapCapsuleCapsule(V3 pos0, Q4 or0, float hh0, float r0, V3 pos1, Q4 or1, float hh1, float r1) {
    V3 up = mk3(0.f, 1.f, 0.f);
    V3 v0 = rotate(or0, up), v1 = rotate(or1, up);
    V3 c0, c1;
    closestPointsSegSegUnit(pos0, v0, -hh0, hh0, pos1, v1, -hh1, hh1, c0, c1);
    V3 d = c1 - c0;
    float rs = r0 + r1;
    return !(dot(d, d) > rs * rs);
}

__device__ inline bool overlapSphereCapsule(V3 pos0, float r0, V3 pos1, Q4 or1, float hh1, float r1) {
    V3 dir = rotate(or1, mk3(0.f, 1.f, 0.f));
    V3 q = closestPointOnSegment(pos0, pos1, dir, -hh1, hh1);
    V3 v = q - pos0;
    float rs = r0 + r1;
    return !(dot(v, v) > rs * rs);
}

__device__ inline bool overlapSphereBox(V3 pos0, float r0, V3 pos1, Q4 or1, V3 he1) {
    M3 u1 = mat3_cast(or1);
    V3 d = pos0 - pos1;
    d = mk3(dot(d, u1.c[0]), dot(d, u1.c[1]), dot(d, u1.c[2]));
    V3 q = pos1 + mul(u1, gclamp(d, -he1, he1));
    V3 v = q - pos0;
    return !(dot(v, v) > r0 * r0);
}

__device__ inline bool overlapCapsuleBox(V3 pos0, Q4 or0, float hh0, float r0, V3 pos1, Q4 or1, V3 he1) {
    M3 u1 = mat3_cast(or1);
    V3 p = pos0 - pos1;
    p = mk3(dot(p, u1.c[0]), dot(p, u1.c[1]), dot(p, u1.c[2]));
    V3 dir = rotate(or0, mk3(0.f, 1.f, 0.f));
    V3 dirLc = mk3(dot(dir, u1.c[0]), dot(dir, u1.c[1]), dot(dir, u1.c[2]));
    float t; V3 q;
    float sd = sqrDistSegmentAABB(p, dirLc, -hh0, hh0, he1, t, q);
    return !(sd >= r0 * r0);
}

// Overlap.cpp:44-149: 15-axis SAT, separation d = |t.L| - ra - rb > 0 on any axis -> no overlap
__device__ inline bool overlapBoxBox(V3 pos0, Q4 or0, V3 he0, V3 pos1, Q4 or1, V3 he1) {
    M3 u0 = mat3_cast(or0), u1 = mat3_cast(or1);
    V3 t = pos1 - pos0;
    t = mk3(dot(t, u0.c[0]), dot(t, u0.c[1]), dot(t, u0.c[2]));
    float r[3][3], a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { r[i][j] = dot(u0.c[i], u1.c[j]); a[i][j] = fabsf(r[i][j]); }
    const float e0[3] = { he0.x, he0.y, he0.z }, e1[3] = { he1.x, he1.y, he1.z }, tt[3] = { t.x, t.y, t.z };
    for (int i = 0; i < 3; ++i) {
        float ra = e0[i];
        float rb = e1[0] * a[i][0] + e1[1] * a[i][1] + e1[2] * a[i][2];
        if (fabsf(tt[i]) - ra - rb > 0.f) return false;
    }
    for (int i = 0; i < 3; ++i) {
        float ra = e0[0] * a[0][i] + e0[1] * a[1][i] + e0[2] * a[2][i];
        float rb = e1[i];
        if (fabsf(tt[0] * r[0][i] + tt[1] * r[1][i] + tt[2] * r[2][i]) - ra - rb > 0.f) return false;
    }
    // edge-edge axes A_i x B_j, written out in the reference's order and with its index pattern
    // A0 x B0..B2
    if (fabsf(tt[2] * r[1][0] - tt[1] * r[2][0]) - (e0[1] * a[2][0] + e0[2] * a[1][0]) - (e1[1] * a[0][2] + e1[2] * a[0][1]) > 0.f) return false;
    if (fabsf(tt[2] * r[1][1] - tt[1] * r[2][1]) - (e0[1] * a[2][1] + e0[2] * a[1][1]) - (e1[0] * a[0][2] + e1[2] * a[0][0]) > 0.f) return false;
    if (fabsf(tt[2] * r[1][2] - tt[1] * r[2][2]) - (e0[1] * a[2][2] + e0[2] * a[1][2]) - (e1[0] * a[0][1] + e1[1] * a[0][0]) > 0.f) return false;
    // A1 x B0..B2
    if (fabsf(tt[0] * r[2][0] - tt[2] * r[0][0]) - (e0[0] * a[2][0] + e0[2] * a[0][0]) - (e1[1] * a[1][2] + e1[2] * a[1][1]) > 0.f) return false;
    if (fabsf(tt[0] * r[2][1] - tt[2] * r[0][1]) - (e0[0] * a[2][1] + e0[2] * a[0][1]) - (e1[0] * a[1][2] + e1[2] * a[1][0]) > 0.f) return false;
    if (fabsf(tt[0] * r[2][2] - tt[2] * r[0][2]) - (e0[0] * a[2][2] + e0[2] * a[0][2]) - (e1[0] * a[1][1] + e1[1] * a[1][0]) > 0.f) return false;
    // A2 x B0..B2
    if (fabsf(tt[1] * r[0][0] - tt[0] * r[1][0]) - (e0[0] * a[1][0] + e0[1] * a[0][0]) - (e1[1] * a[2][2] + e1[2] * a[2][1]) > 0.f) return false;
    if (fabsf(tt[1] * r[0][1] - tt[0] * r[1][1]) - (e0[0] * a[1][1] + e0[1] * a[0][1]) - (e1[0] * a[2][2] + e1[2] * a[2][0]) > 0.f) return false;
    if (fabsf(tt[1] * r[0][2] - tt[0] * r[1][2]) - (e0[0] * a[1][2] + e0[1] * a[0][2]) - (e1[0] * a[2][1] + e1[1] * a[2][0]) > 0.f) return false;
    return true;
}

// physecs::overlap dispatch (Overlap.cpp:203-235).  Every "X vs convex" case keeps the convex mesh as the second GJK
// operand and pos1 - pos0 of the CALL as the start direction, so swapped calls start from the other side.
__device__ inline bool overlapShapes(int t0, float4 q0, V3 pos0, Q4 or0, int mesh0, int t1, float4 q1, V3 pos1, Q4 or1, int mesh1,
                                     const PbConvexDev* convexes) {
    if (t0 > PB_CONVEX_MESH || t1 > PB_CONVEX_MESH) return false;
    if (t0 == PB_CONVEX_MESH && t1 == PB_CONVEX_MESH)
        return gjkBool(makeShape(t0, q0, pos0, or0, convexes, mesh0), makeShape(t1, q1, pos1, or1, convexes, mesh1), pos1 - pos0);
    if (t0 == PB_CONVEX_MESH || t1 == PB_CONVEX_MESH) {
        // canonical call: (other shape, convex); the reference swaps the arguments when the convex is side 0
        bool sw = t0 == PB_CONVEX_MESH;
        int ta = sw ? t1 : t0, mb = sw ? mesh0 : mesh1;
        float4 qa = sw ? q1 : q0, qb = sw ? q0 : q1;
        V3 pa = sw ? pos1 : pos0, pb = sw ? pos0 : pos1;
        Q4 oa = sw ? or1 : or0, ob = sw ? or0 : or1;
        Shape sa = makeShape(ta, qa, pa, oa, convexes, 0);
        // quirk Q22 (Overlap.cpp:199): box-vs-convex places the convex support at the BOX position
        Shape sb = makeShape(PB_CONVEX_MESH, qb, ta == PB_BOX ? pa : pb, ob, convexes, mb);
        return gjkBool(sa, sb, pb - pa);
    }
    V3 he0 = mk3(q0.x, q0.y, q0.z), he1 = mk3(q1.x, q1.y, q1.z);
    if (t0 == PB_SPHERE) {
        if (t1 == PB_SPHERE) return overlapSphereSphere(pos0, q0.x, pos1, q1.x);
        if (t1 == PB_CAPSULE) return overlapSphereCapsule(pos0, q0.x, pos1, or1, q1.x, q1.y);
        return overlapSphereBox(pos0, q0.x, pos1, or1, he1);
    }
    if (t0 == PB_CAPSULE) {
        if (t1 == PB_SPHERE) return overlapSphereCapsule(pos1, q1.x, pos0, or0, q0.x, q0.y);
        if (t1 == PB_CAPSULE) return overlapCapsuleCapsule(pos0, or0, q0.x, q0.y, pos1, or1, q1.x, q1.y);
        return overlapCapsuleBox(pos0, or0, q0.x, q0.y, pos1, or1, he1);
    }
    if (t1 == PB_SPHERE) return overlapSphereBox(pos1, q1.x, pos0, or0, he0);
    if (t1 == PB_CAPSULE) return overlapCapsuleBox(pos1, or1, q1.x, q1.y, pos0, or0, he0);
    return overlapBoxBox(pos0, or0, he0, pos1, or1, he1);
}

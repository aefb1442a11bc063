// Analytic shape-pair routines (sphere / capsule / box), one device function per reference routine.
// Each returns true on contact and fills a Manifold: unit normal from shape 0 to shape 1, world-space
// witness points p0 (on shape 0) / p1 (on shape 1).
//   collisionSphereSphere    reference src/Collision.cpp:109-118
//   collisionCapsuleCapsule  :120-133
//   collisionSphereCapsule   :428-440
//   collisionSphereBox       :442-486
//   collisionCapsuleBox      :584-692  (+ generateCapsuleBoxContacts :501-582)
//   collisionBoxBox          :135-350  (+ generateBoxBoxFaceContacts :54-107)
#pragma once
#include "np_geom.cuh"
#include "np_clip.cuh"

__device__ inline bool collideSphereSphere(V3 pos0, float r0, V3 pos1, float r1, Manifold& m) {
    V3 d = pos1 - pos0;
    float rs = r0 + r1;
    if (dot(d, d) > rs * rs) return false;
    float len = length(d);
    m.n = len ? d / len : mk3(0.f, 1.f, 0.f);
    m.np = 1;
    m.p0[0] = pos0 + m.n * r0;
    m.p1[0] = pos1 - m.n * r1;
    return true;
}

__device__ inline bool collideCapsuleCapsule(V3 pos0, Q4 or0, float hh0, float r0, V3 pos1, Q4 or1, float hh1, float r1, Manifold& m) {
    V3 up = mk3(0.f, 1.f, 0.f);
    V3 v0 = rotate(or0, up), v1 = rotate(or1, up);
    V3 c0, c1;
    closestPointsSegSegUnit(pos0, v0, -hh0, hh0, pos1, v1, -hh1, hh1, c0, c1);
    V3 d = c1 - c0;
    float rs = r0 + r1;
    if (dot(d, d) > rs * rs) return false;
    float len = length(d);
    m.n = len ? d / len : up;
    m.np = 1;
    m.p0[0] = c0 + m.n * r0;
    m.p1[0] = c1 - m.n * r1;
    return true;
}

__device__ inline bool collideSphereCapsule(V3 pos0, float r0, V3 pos1, Q4 or1, float hh1, float r1, Manifold& m) {
    V3 up = mk3(0.f, 1.f, 0.f);
    V3 dir = rotate(or1, up);
    V3 q = closestPointOnSegment(pos0, pos1, dir, -hh1, hh1);
    V3 v = q - pos0;
    float rs = r0 + r1;
    if (dot(v, v) > rs * rs) return false;
    float len = length(v);
    m.n = len ? v / len : up;
    m.np = 1;
    m.p0[0] = pos0 + m.n * r0;
    m.p1[0] = q - m.n * r1;
    return true;
}

__device__ inline bool collideSphereBox(V3 pos0, float r0, V3 pos1, Q4 or1, V3 he1, Manifold& m) {
    M3 u1 = mat3_cast(or1);
    V3 d = pos0 - pos1;
    d = mk3(dot(d, u1.c[0]), dot(d, u1.c[1]), dot(d, u1.c[2]));
    V3 q = pos1 + mul(u1, gclamp(d, -he1, he1));
    V3 v = q - pos0;
    if (dot(v, v) > r0 * r0) return false;
    float len = length(v);
    if (len) {
        m.n = v / len;
        m.np = 1;
        m.p0[0] = pos0 + m.n * r0;
        m.p1[0] = q;
    } else {
        float mn = FLT_MAX;
        int axis = 0;
        for (int i = 0; i < 3; ++i) {
            float depth = get(he1, i) - fabsf(get(d, i));
            if (depth < mn) { mn = depth; axis = i; }
        }
        float da = get(d, axis);
        float sign = da ? gsign(da) : 1.f;
        V3 boxParam = d;
        set(boxParam, axis, sign * get(he1, axis));
        m.n = -sign * u1.c[axis];
        m.np = 1;
        m.p0[0] = pos0;
        m.p1[0] = pos1 + mul(u1, boxParam);
    }
    return true;
}

// Collision.cpp:501-582
__device__ inline void capsuleBoxContacts(V3 p0L, V3 p1L, float radius, V3 normal, V3 boxCenter, V3 he, const M3& basis, Manifold& m) {
    int boxAxis = 0;
    float boxAxisSign = 0.f;
    float maxDot = 0.f;
    for (int i = 0; i < 3; ++i) {
        float d = dot(normal, basis.c[i]);
        float ad = fabsf(d);
        if (ad > maxDot) { maxDot = ad; boxAxis = i; boxAxisSign = d < 0.f ? 1.f : -1.f; }
    }
    int clipX = (boxAxis + 1) % 3, clipY = (boxAxis + 2) % 3;
    V2 line0 = mk2(get(p0L, clipX), get(p0L, clipY));
    V2 line1 = mk2(get(p1L, clipX), get(p1L, clipY));
    float hx = get(he, clipX), hy = get(he, clipY), ha = get(he, boxAxis);
    V2 clip[4] = { mk2(hx, hy), mk2(hx, -hy), mk2(-hx, -hy), mk2(-hx, hy) };
    if (!clipLine(line0, line1, clip, 4)) { m.np = 0; return; }
    float onBoxAxis = boxAxisSign * ha;
    float a0, a1;
    float p0x = get(p0L, clipX), p1x = get(p1L, clipX), p0y = get(p0L, clipY), p1y = get(p1L, clipY);
    float p0a = get(p0L, boxAxis), p1a = get(p1L, boxAxis);
    float distX = p0x - p1x, distY = p0y - p1y;
    float adx = fabsf(distX), ady = fabsf(distY);
    if (adx > ady) {
        a0 = gmix(p0a, p1a, (p0x - line0.x) / distX);
        a1 = gmix(p1a, p0a, (p1x - line1.x) / -distX);
    } else if (ady > adx) {
        a0 = gmix(p0a, p1a, (p0y - line0.y) / distY);
        a1 = gmix(p1a, p0a, (p1y - line1.y) / -distY);
    } else { a0 = p0a; a1 = p1a; }
    a0 -= boxAxisSign * radius;
    a1 -= boxAxisSign * radius;
    int np = 0;
    if (fabsf(a0) < ha) {
        V3 p = mk3(0.f); set(p, boxAxis, a0); set(p, clipX, line0.x); set(p, clipY, line0.y);
        V3 ob = p; set(ob, boxAxis, onBoxAxis);
        m.p0[np] = boxCenter + mul(basis, p); m.p1[np] = boxCenter + mul(basis, ob); ++np;
    }
    if (fabsf(a1) < ha) {
        V3 p = mk3(0.f); set(p, boxAxis, a1); set(p, clipX, line1.x); set(p, clipY, line1.y);
        V3 ob = p; set(ob, boxAxis, onBoxAxis);
        m.p0[np] = boxCenter + mul(basis, p); m.p1[np] = boxCenter + mul(basis, ob); ++np;
    }
    m.np = np;
}

__device__ inline bool collideCapsuleBox(V3 pos0, Q4 or0, float hh0, float r0, V3 pos1, Q4 or1, V3 he1, Manifold& m) {
    M3 u1 = mat3_cast(or1);
    V3 p = pos0 - pos1;
    p = mk3(dot(p, u1.c[0]), dot(p, u1.c[1]), dot(p, u1.c[2]));
    V3 dir = rotate(or0, mk3(0.f, 1.f, 0.f));
    V3 dirLc = mk3(dot(dir, u1.c[0]), dot(dir, u1.c[1]), dot(dir, u1.c[2]));
    float t; V3 q;
    float sq = sqrDistSegmentAABB(p, dirLc, -hh0, hh0, he1, t, q);
    if (sq >= r0 * r0) return false;
    V3 onSegment = pos0 + dir * t;
    V3 onBox = pos1 + mul(u1, q);
    V3 normal = onBox - onSegment;
    float nLen = length(normal);
    if (nLen) {
        m.n = normal / nLen;
    } else {
        V3 r, absR;
        for (int i = 0; i < 3; ++i) { float v = dot(dir, u1.c[i]); set(r, i, v); set(absR, i, fabsf(v)); }
        float mx = -FLT_MAX;
        bool edge = false;
        int axis = 0;
        float ra, rb, l, d;
        for (int i = 0; i < 3; ++i) {
            ra = hh0 * get(absR, i) + r0;
            rb = get(he1, i);
            l = fabsf(get(p, i));
            d = l - ra - rb;
            if (d > mx) { mx = d; edge = false; axis = i; }
        }
        const float edgeLimit = 0.999f;
        float edgeOffset = 0.1f;
        rb = he1.y * absR.z + he1.z * absR.y;
        l = fabsf(p.z * r.y - p.y * r.z);
        d = l - r0 - rb;
        if (absR.x < edgeLimit && d > mx + edgeOffset) { mx = d; edge = true; axis = 0; edgeOffset = 0.f; }
        rb = he1.x * absR.z + he1.z * absR.x;
        l = fabsf(p.x * r.z - p.z * r.x);
        d = l - r0 - rb;
        if (absR.y < edgeLimit && d > mx + edgeOffset) { mx = d; edge = true; axis = 1; edgeOffset = 0.f; }
        rb = he1.x * absR.y + he1.y * absR.x;
        l = fabsf(p.y * r.x - p.x * r.y);
        d = l - r0 - rb;
        if (absR.z < edgeLimit && d > mx + edgeOffset) { mx = d; edge = true; axis = 2; }
        V3 sep = edge ? cross(dir, u1.c[axis]) : u1.c[axis];
        m.n = normalize(dot(sep, pos1 - pos0) > 0.f ? sep : -sep);
    }
    capsuleBoxContacts(p - dirLc * hh0, p + dirLc * hh0, r0, m.n, pos1, he1, u1, m);
    if (!m.np) {
        m.np = 1;
        m.p0[0] = onSegment + m.n * r0;
        m.p1[0] = onBox;
    }
    return true;
}

// Collision.cpp:54-107: clip the incident face of `inc` against the reference face of `ref`
__device__ inline void boxBoxFaceContacts(const M3& uRef, V3 posRef, V3 heRef, const M3& uInc, V3 posInc, V3 heInc, int refAxis,
                                          V3* c0, V3* c1, int& numPoints) {
    V3 axis = uRef.c[refAxis];
    float refSign = dot(axis, posInc - posRef) < 0.f ? -1.f : 1.f;
    float maxDot = 0.f;
    int incAxis = 0;
    float incSign = 1.f;
    for (int i = 0; i < 3; ++i) {
        float d = dot(-refSign * axis, uInc.c[i]);
        float ad = fabsf(d);
        if (ad > maxDot) { maxDot = ad; incAxis = i; incSign = d < 0.f ? -1.f : 1.f; }
    }
    int ia1 = (incAxis + 1) % 3, ia2 = (incAxis + 2) % 3;
    M3 invURef = inverse(uRef);
    V3 incPlaneOrig = mul(invURef, posInc + uInc.c[incAxis] * get(heInc, incAxis) * incSign - posRef);
    V3 e1 = uInc.c[ia1] * get(heInc, ia1), e2 = uInc.c[ia2] * get(heInc, ia2);
    V3 ne1 = -uInc.c[ia1] * get(heInc, ia1);
    V3 q0 = incPlaneOrig + mul(invURef, e1 + e2);
    V3 q1 = incPlaneOrig + mul(invURef, ne1 + e2);
    V3 q2 = incPlaneOrig + mul(invURef, ne1 - e2);
    V3 q3 = incPlaneOrig + mul(invURef, e1 - e2);
    int clipX = (refAxis + 1) % 3, clipY = (refAxis + 2) % 3;
    Poly<8> poly;
    poly.n = 4; poly.overflow = false;
    poly.p[0] = mk2(get(q0, clipX), get(q0, clipY));
    poly.p[1] = mk2(get(q1, clipX), get(q1, clipY));
    poly.p[2] = mk2(get(q2, clipX), get(q2, clipY));
    poly.p[3] = mk2(get(q3, clipX), get(q3, clipY));
    float hx = get(heRef, clipX), hy = get(heRef, clipY);
    V2 clip[4] = { mk2(hx, hy), mk2(hx, -hy), mk2(-hx, -hy), mk2(-hx, hy) };
    suthHodgClip<8, 4>(poly, clip, 4);
    V3 incPlaneNormal = mul(invURef, uInc.c[incAxis]);
    contactsPolygonBoxFace<8>(posRef, uRef, refAxis, refSign, heRef, incPlaneOrig, incPlaneNormal, poly, clipX, clipY, c0, c1, numPoints);
}

__device__ inline bool collideBoxBox(V3 pos0, Q4 or0, V3 he0, V3 pos1, Q4 or1, V3 he1, Manifold& m) {
    float edgeOffset = 0.1f;
    const float edgeLimit = 0.999f;
    float ra, rb, l, d;
    bool isEdge = false;
    int i0 = 0, i1 = 0;   // FACE: (box, axis)  EDGE: (edge0, edge1)
    float mx = -FLT_MAX;
    M3 u0 = mat3_cast(or0), u1 = mat3_cast(or1);
    V3 t = pos1 - pos0;
    t = mk3(dot(t, u0.c[0]), dot(t, u0.c[1]), dot(t, u0.c[2]));
    float r[3][3], ar[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { r[i][j] = dot(u0.c[i], u1.c[j]); ar[i][j] = fabsf(r[i][j]); }
    float h0[3] = { he0.x, he0.y, he0.z }, h1[3] = { he1.x, he1.y, he1.z }, tt[3] = { t.x, t.y, t.z };
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ra = h0[i];
        rb = h1[0] * ar[i][0] + h1[1] * ar[i][1] + h1[2] * ar[i][2];
        l = fabsf(tt[i]);
        d = l - ra - rb;
        if (d > 0.f) return false;
        if (d > mx) { mx = d; isEdge = false; i0 = 0; i1 = i; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ra = h0[0] * ar[0][i] + h0[1] * ar[1][i] + h0[2] * ar[2][i];
        rb = h1[i];
        l = fabsf(tt[0] * r[0][i] + tt[1] * r[1][i] + tt[2] * r[2][i]);
        d = l - ra - rb;
        if (d > 0.f) return false;
        if (d > mx) { mx = d; isEdge = false; i0 = 1; i1 = i; }
    }
    // 9 edge axes A_i x B_j; the last one (A2 x B2) does not reset edgeOffset (Collision.cpp:304-308)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int ia = (i + 1) % 3, ib = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int ja = (j + 1) % 3, jb = (j + 2) % 3;
            // ra = h0[ia]*ar[ib][j] + h0[ib]*ar[ia][j] ; rb = h1[ja]*ar[i][jb] + h1[jb]*ar[i][ja] ; l = |t[ib]*r[ia][j] - t[ia]*r[ib][j]|
            // written in the reference's term order for each (i, j):
            if (i == 0) { ra = h0[1] * ar[2][j] + h0[2] * ar[1][j]; l = fabsf(tt[2] * r[1][j] - tt[1] * r[2][j]); }
            else if (i == 1) { ra = h0[0] * ar[2][j] + h0[2] * ar[0][j]; l = fabsf(tt[0] * r[2][j] - tt[2] * r[0][j]); }
            else { ra = h0[0] * ar[1][j] + h0[1] * ar[0][j]; l = fabsf(tt[1] * r[0][j] - tt[0] * r[1][j]); }
            if (j == 0) rb = h1[1] * ar[i][2] + h1[2] * ar[i][1];
            else if (j == 1) rb = h1[0] * ar[i][2] + h1[2] * ar[i][0];
            else rb = h1[0] * ar[i][1] + h1[1] * ar[i][0];
            d = l - ra - rb;
            if (d > 0.f) return false;
            if (ar[i][j] < edgeLimit && d > mx + edgeOffset) {
                mx = d; isEdge = true; i0 = i; i1 = j;
                if (!(i == 2 && j == 2)) edgeOffset = 0.f;
            }
            (void)ia; (void)ib; (void)ja; (void)jb;
        }
    }
    if (!isEdge) {
        V3 axis;
        V3 c0[4], c1[4];
        int np = 0;
        if (i0) {
            axis = u1.c[i1];
            boxBoxFaceContacts(u1, pos1, he1, u0, pos0, he0, i1, c0, c1, np);
            for (int i = 0; i < np; ++i) { m.p0[i] = c1[i]; m.p1[i] = c0[i]; }
        } else {
            axis = u0.c[i1];
            boxBoxFaceContacts(u0, pos0, he0, u1, pos1, he1, i1, c0, c1, np);
            for (int i = 0; i < np; ++i) { m.p0[i] = c0[i]; m.p1[i] = c1[i]; }
        }
        m.np = np;
        m.n = dot(axis, pos1 - pos0) < 0.f ? -axis : axis;
    } else {
        V3 axis = cross(u0.c[i0], u1.c[i1]);
        m.n = normalize(dot(axis, pos1 - pos0) < 0.f ? -axis : axis);
        V3 s0 = mk3(-1.f);
        set(s0, (i0 + 1) % 3, dot(u0.c[(i0 + 1) % 3], m.n) < 0.f ? -1.f : 1.f);
        set(s0, (i0 + 2) % 3, dot(u0.c[(i0 + 2) % 3], m.n) < 0.f ? -1.f : 1.f);
        V3 p0 = pos0 + s0.x * he0.x * u0.c[0] + s0.y * he0.y * u0.c[1] + s0.z * he0.z * u0.c[2];
        V3 s1 = mk3(-1.f);
        set(s1, (i1 + 1) % 3, dot(u1.c[(i1 + 1) % 3], m.n) > 0.f ? -1.f : 1.f);
        set(s1, (i1 + 2) % 3, dot(u1.c[(i1 + 2) % 3], m.n) > 0.f ? -1.f : 1.f);
        V3 p1 = pos1 + s1.x * he1.x * u1.c[0] + s1.y * he1.y * u1.c[1] + s1.z * he1.z * u1.c[2];
        m.np = 1;
        V3 c0, c1;
        closestPointsSegSegUnit(p0, u0.c[i0], 0.f, get(he0, i0) * 2.f, p1, u1.c[i1], 0.f, get(he1, i1) * 2.f, c0, c1);
        m.p0[0] = c0; m.p1[0] = c1;
    }
    return true;
}

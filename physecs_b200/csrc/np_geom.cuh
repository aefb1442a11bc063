// Closest-point geometry used by the narrowphase bins (device restatement, same branch structure and
// fp32 operation order as the reference so feature decisions agree):
//   closestPointOnSegment            reference src/GeomUtil.cpp:7-12
//   closestPointsBetweenSegments     :14-59 (unit directions)  / :61-127 (general)
//   sqrDistSegmentAABB               :455-467 (+ line/AABB cases :129-453, incl. the region-5 quirk at :273/:326)
//   sqrDistPointTriangle             :469-538
//   sqrDistSegmentTriangle           :649-660 (+ :540-647)
#pragma once
#include "pb_math.cuh"

enum { TF_FACE = 0, TF_EDGE = 1, TF_VERTEX = 2 };

__device__ __forceinline__ V3 closestPointOnSegment(V3 point, V3 orig, V3 dir, float mn, float mx) {
    V3 r = orig - point;
    float t = -dot(r, dir);
    t = gclamp(t, mn, mx);
    return orig + t * dir;
}

// unit-direction variant
__device__ inline void closestPointsSegSegUnit(V3 o0, V3 d0, float min0, float max0, V3 o1, V3 d1, float min1, float max1, V3& p0, V3& p1) {
    float v1v2 = dot(d0, d1);
    V3 r = o1 - o0;
    float rv1 = dot(r, d0);
    float rv2 = dot(r, d1);
    float t1 = fabsf(v1v2) > 1.f - 0.0001f ? 0.f : (rv1 * v1v2 - rv2) / (1.f - v1v2 * v1v2);
    float t0 = rv1 + t1 * v1v2;
    if (t0 < min0) { t0 = min0; p0 = o0 + t0 * d0; t1 = -dot(o1 - p0, d1); }
    else if (t0 > max0) { t0 = max0; p0 = o0 + t0 * d0; t1 = -dot(o1 - p0, d1); }
    else p0 = o0 + t0 * d0;
    if (t1 < min1) {
        t1 = min1; p1 = o1 + t1 * d1;
        t0 = -dot(o0 - p1, d0); t0 = gclamp(t0, min0, max0); p0 = o0 + t0 * d0;
    } else if (t1 > max1) {
        t1 = max1; p1 = o1 + t1 * d1;
        t0 = -dot(o0 - p1, d0); t0 = gclamp(t0, min0, max0); p0 = o0 + t0 * d0;
    } else p1 = o1 + t1 * d1;
}

// general (non-normalised edge vectors, parameters in [0,1])
__device__ inline void closestPointsSegSeg(V3 o0, V3 v0, V3 o1, V3 v1, V3& p0, V3& p1) {
    float v0v0 = dot(v0, v0);
    if (v0v0 < 0.0001f) {
        float v1v1 = dot(v1, v1);
        if (v1v1 < 0.0001f) { p0 = o0; p1 = o1; return; }
        float t1 = -dot(o1 - o0, v1) / v1v1;
        p0 = o0; p1 = o1 + t1 * v1; return;
    }
    float v1v1 = dot(v1, v1);
    if (v1v1 < 0.0001f) {
        float t0 = -dot(o0 - o1, v0) / v0v0;
        p0 = o0 + t0 * v0; p1 = o1; return;
    }
    float v0v1 = dot(v0, v1);
    V3 r = o1 - o0;
    float rv0 = dot(r, v0);
    float rv1 = dot(r, v1);
    float denom = v0v0 * v1v1 - v0v1 * v0v1;
    float t1 = fabsf(denom) < 0.0001f ? 0.f : (rv0 * v0v1 - rv1 * v0v0) / denom;
    float t0 = (rv0 + t1 * v0v1) / v0v0;
    if (t0 < 0.f) { t0 = 0.f; p0 = o0 + t0 * v0; t1 = -dot(o1 - p0, v1) / v1v1; }
    else if (t0 > 1.f) { t0 = 1.f; p0 = o0 + t0 * v0; t1 = -dot(o1 - p0, v1) / v1v1; }
    else p0 = o0 + t0 * v0;
    if (t1 < 0.f) {
        t1 = 0.f; p1 = o1 + t1 * v1;
        t0 = -dot(o0 - p1, v0) / v0v0; t0 = gclamp(t0, 0.f, 1.f); p0 = o0 + t0 * v0;
    } else if (t1 > 1.f) {
        t1 = 1.f; p1 = o1 + t1 * v1;
        t0 = -dot(o0 - p1, v0) / v0v0; t0 = gclamp(t0, 0.f, 1.f); p0 = o0 + t0 * v0;
    } else p1 = o1 + t1 * v1;
}

// ---- line / segment vs AABB (Eberly) ---------------------------------------------------------------------------
__device__ inline float sdCaseTwoZeros(V3 p, int axis, V3 he, float& t, V3& q) {
    float sq = 0.f;
    t = get(he, axis) - get(p, axis);
    set(q, axis, get(p, axis) + t);
    int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int a = k ? a2 : a1;
        float pa = get(p, a), ha = get(he, a);
        if (pa > ha) { float d = pa - ha; sq += d * d; set(q, a, ha); }
        else if (pa < -ha) { float d = pa + ha; sq += d * d; set(q, a, -ha); }
        else set(q, a, pa);
    }
    return sq;
}

__device__ inline float sdCaseOneZero(V3 p, int zeroAxis, V3 dir, V3 he, float& t, V3& q) {
    float sq = 0.f;
    set(q, zeroAxis, get(p, zeroAxis));
    V3 pme = p - he;
    int x = (zeroAxis + 1) % 3, y = (zeroAxis + 2) % 3;
    float dx = get(dir, x), dy = get(dir, y), px = get(p, x), py = get(p, y), hx = get(he, x), hy = get(he, y);
    float pmx = get(pme, x), pmy = get(pme, y);
    float prod0 = dy * pmx;
    float prod1 = dx * pmy;
    if (prod0 >= prod1) {
        set(q, x, hx);
        float eY = py + hy;
        float delta = prod0 - dx * eY;
        if (delta >= 0.f) {
            sq += delta * delta;
            set(q, y, -hy);
            t = -(pmx * dx + eY * dy);
        } else {
            set(q, y, py - prod0 / dx);
            t = -pmx / dx;
        }
    } else {
        set(q, y, hy);
        float eX = px + hx;
        float delta = prod1 - dy * eX;
        if (delta >= 0.f) {
            sq += delta * delta;
            set(q, x, -hx);
            t = -(eX * dx + pmy * dy);
        } else {
            set(q, x, px - prod1 / dy);
            t = -pmy / dy;
        }
    }
    float pz = get(p, zeroAxis), hz = get(he, zeroAxis);
    if (pz < -hz) { float d = pz + hz; sq += d * d; set(q, zeroAxis, -hz); }
    else if (pz > hz) { float d = pz - hz; sq += d * d; set(q, zeroAxis, hz); }
    return sq;
}

__device__ inline float sdLineAABBFace(V3 p, V3 dir, V3 he, int x, V3 pme, float& t, V3& q) {
    q = p; t = 0.f;
    float sq = 0.f;
    int y = (x + 1) % 3, z = (y + 1) % 3;
    V3 ppe = p + he;
    float dX = get(dir, x), dY = get(dir, y), dZ = get(dir, z);
    float pmX = get(pme, x), pmY = get(pme, y), pmZ = get(pme, z);
    float ppY = get(ppe, y), ppZ = get(ppe, z);
    float hX = get(he, x), hY = get(he, y), hZ = get(he, z);
    float pY = get(p, y), pZ = get(p, z);
    // shared region evaluators ------------------------------------------------------------------
    #define REGION4(nearY) { float tmp = pY - (nearY); float delta = dX * pmX + dY * tmp + dZ * ppZ; t = -delta; \
        sq = pmX * pmX + tmp * tmp + ppZ * ppZ - delta * delta; set(q, x, hX); set(q, y, (nearY)); set(q, z, -hZ); }
    #define REGION5() { float delta = dX * pmX + dY * pmY + dZ * ppZ; t = -delta; \
        sq = pmX * pmX + pmY + pmY + ppZ * ppZ - delta * delta; set(q, x, hX); set(q, y, hY); set(q, z, -hZ); }   /* sic: reference :273/:326 */
    #define REGION2(nearZ) { float tmp = pZ - (nearZ); float delta = dX * pmX + dY * ppY + dZ * tmp; t = -delta; \
        sq = pmX * pmX + ppY * ppY + tmp * tmp - delta * delta; set(q, x, hX); set(q, y, -hY); set(q, z, (nearZ)); }
    #define REGION1() { float delta = dX * pmX + dY * ppY + dZ * pmZ; t = -delta; \
        sq = pmX * pmX + ppY * ppY + pmZ * pmZ - delta * delta; set(q, x, hX); set(q, y, -hY); set(q, z, hZ); }
    if (dX * ppY >= dY * pmX) {
        if (dX * ppZ >= dZ * pmX) {
            // region 0: line intersects the face
            set(q, x, hX);
            set(q, y, get(q, y) - pmX * dY / dX);
            set(q, z, get(q, z) - pmX * dZ / dX);
            t = -pmX / dX;
        } else {
            float sqrLen = dX * dX + dZ * dZ;
            float nearY = pY - dY * (dX * pmX + dZ * ppZ) / sqrLen;
            if (nearY <= hY) REGION4(nearY) else REGION5()
        }
    } else {
        if (dX * ppZ >= dZ * pmX) {
            float sqrLen = dX * dX + dY * dY;
            float nearZ = pZ - dZ * (dX * pmX + dY * ppY) / sqrLen;
            if (nearZ <= hZ) REGION2(nearZ) else REGION1()
        } else {
            float nearY = pY - dY * (dX * pmX + dZ * ppZ) / (dX * dX + dZ * dZ);
            if (nearY >= -hY) {
                if (nearY <= hY) REGION4(nearY) else REGION5()
                return sq;
            }
            float nearZ = pZ - dZ * (dX * pmX + dY * ppY) / (dX * dX + dY * dY);
            if (nearZ >= -hZ) {
                if (nearZ <= hZ) REGION2(nearZ) else REGION1()
                return sq;
            }
            // region 3
            float delta = dX * pmX + dY * ppY + dZ * ppZ;
            t = -delta;
            sq = pmX * pmX + ppY * ppY + ppZ * ppZ - delta * delta;
            set(q, x, hX); set(q, y, -hY); set(q, z, -hZ);
        }
    }
    #undef REGION4
    #undef REGION5
    #undef REGION2
    #undef REGION1
    return sq;
}

__device__ inline float sdCaseNoZeroes(V3 p, V3 dir, V3 he, float& t, V3& q) {
    V3 pme = p - he;
    float dxEy = dir.x * pme.y;
    float dyEx = dir.y * pme.x;
    if (dyEx >= dxEy) {
        float dzEx = dir.z * pme.x;
        float dxEz = dir.x * pme.z;
        if (dzEx >= dxEz) return sdLineAABBFace(p, dir, he, 0, pme, t, q);
        return sdLineAABBFace(p, dir, he, 2, pme, t, q);
    }
    float dzEy = dir.z * pme.y;
    float dyEz = dir.y * pme.z;
    if (dzEy >= dyEz) return sdLineAABBFace(p, dir, he, 1, pme, t, q);
    return sdLineAABBFace(p, dir, he, 2, pme, t, q);
}

__device__ inline float sqrDistPointAABB(V3 p, V3 he, V3& q) {
    q = gclamp(p, -he, he);
    V3 d = p - q;
    return dot(d, d);
}

__device__ inline float sqrDistLineAABB(V3 p, V3 dir, V3 he, float& t, V3& q) {
    bool rx = dir.x < 0.f, ry = dir.y < 0.f, rz = dir.z < 0.f;
    if (rx) { p.x = -p.x; dir.x = -dir.x; }
    if (ry) { p.y = -p.y; dir.y = -dir.y; }
    if (rz) { p.z = -p.z; dir.z = -dir.z; }
    float sq;
    if (dir.x > 0.f) {
        if (dir.y > 0.f) {
            if (dir.z > 0.f) sq = sdCaseNoZeroes(p, dir, he, t, q);
            else sq = sdCaseOneZero(p, 2, dir, he, t, q);
        } else {
            if (dir.z > 0.f) sq = sdCaseOneZero(p, 1, dir, he, t, q);
            else sq = sdCaseTwoZeros(p, 0, he, t, q);
        }
    } else {
        if (dir.y > 0.f) {
            if (dir.z > 0.f) sq = sdCaseOneZero(p, 0, dir, he, t, q);
            else sq = sdCaseTwoZeros(p, 1, he, t, q);
        } else {
            if (dir.z > 0.f) sq = sdCaseTwoZeros(p, 2, he, t, q);
            else { sq = sqrDistPointAABB(p, he, q); t = 0.f; }
        }
    }
    if (rx) q.x = -q.x;
    if (ry) q.y = -q.y;
    if (rz) q.z = -q.z;
    return sq;
}

__device__ inline float sqrDistSegmentAABB(V3 p, V3 dir, float mn, float mx, V3 he, float& t, V3& q) {
    q = mk3(0.f);
    float sq = sqrDistLineAABB(p, dir, he, t, q);
    if (t < mn) { t = mn; sq = sqrDistPointAABB(p + dir * t, he, q); }
    else if (t > mx) { t = mx; sq = sqrDistPointAABB(p + dir * t, he, q); }
    return sq;
}

// ---- point / segment vs triangle -----------------------------------------------------------------------------------
__device__ inline float sqrDistPointTriangle(V3 p, V3 a, V3 b, V3 c, V3& q, int& feature, int& fidx) {
    V3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) { feature = TF_VERTEX; fidx = 0; q = a; return distance2(p, q); }
    V3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) { feature = TF_VERTEX; fidx = 1; q = b; return distance2(p, q); }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { feature = TF_EDGE; fidx = 2; q = a + d1 / (d1 - d3) * ab; return distance2(p, q); }
    V3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) { feature = TF_VERTEX; fidx = 2; q = c; return distance2(p, q); }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && d4 >= d3 && d5 >= d6) { feature = TF_EDGE; fidx = 0; q = b + (d4 - d3) / (d4 - d3 + d5 - d6) * (c - b); return distance2(p, q); }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { feature = TF_EDGE; fidx = 1; q = c + d2 / (d2 - d6) * ac; return distance2(p, q); }
    feature = TF_FACE;
    float denom = va + vb + vc;
    float u = va / denom, v = vb / denom, w = 1.f - u - v;
    q = u * a + v * b + w * c;
    return distance2(p, q);
}

// Same cases, same expressions as the reference's early-return form, evaluated without branches: the four exits (degenerate,
// s <= 0, s >= 1, interior) differ only in which (q, t) feeds the common tail, and lanes of a warp land in all four.
__device__ inline float sqrDistLineSegment(V3 p, V3 dir, V3 a, V3 b, float& t, V3& q, int& vflag) {
    V3 ab = b - a;
    float v1v2 = dot(dir, ab), v2v2 = dot(ab, ab);
    V3 r = p - a;
    float rv1 = dot(r, dir), rv2 = dot(r, ab);
    float denom = (v1v2 * v1v2 - v2v2);
    const bool degenerate = fabsf(denom) <= 1e-6f;
    float s = (rv1 * v1v2 - rv2) / denom;          // unused (and possibly inf / nan) when degenerate
    const bool atA = degenerate || s <= 0.f;
    const bool atB = !atA && s >= 1.f;
    float tB = -dot(p - b, dir);
    float tI = s * v1v2 - rv1;
    V3 qI = a + s * ab;
    t = atA ? -rv1 : (atB ? tB : tI);
    q.x = atA ? a.x : (atB ? b.x : qI.x); q.y = atA ? a.y : (atB ? b.y : qI.y); q.z = atA ? a.z : (atB ? b.z : qI.z);
    vflag = degenerate ? 0 : (atA ? 1 : (atB ? 2 : 0));
    V3 d = p + t * dir - q;
    return dot(d, d);
}

__device__ inline float det3cols(V3 c0, V3 c1, V3 c2) { M3 A; A.c[0] = c0; A.c[1] = c1; A.c[2] = c2; return det3(A); }

__device__ inline float sqrDistLineTriangle(V3 p, V3 dir, V3 a, V3 b, V3 c, float& t, V3& q, int& feature, int& fidx) {
    V3 e0 = b - a, e1 = c - a;
    V3 nd = -dir;
    float det = det3cols(e0, e1, nd);
    if (fabsf(det) > 1e-6f) {
        V3 diff = p - a;
        float invDet = 1.f / det;
        float u = invDet * det3cols(diff, e1, nd);
        float v = invDet * det3cols(e0, diff, nd);
        float w = 1.f - u - v;
        if (u >= 0.f && v >= 0.f && w >= 0.f) {
            t = invDet * det3cols(e0, e1, diff);
            q = a + u * e0 + v * e1;
            feature = TF_FACE;
            return 0.f;
        }
    }
    // edges (a,b) -> fidx 2, (b,c) -> 0, (c,a) -> 1, the later one wins only when strictly closer; one copy of the routine, the
    // vertices rotate through registers
    int vf = 0;
    float dist = 0.f;
    fidx = 2;
    V3 x = a, y = b, z = c;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        float t1; V3 q1; int vf1;
        float dist1 = sqrDistLineSegment(p, dir, x, y, t1, q1, vf1);
        if (k == 0 || dist1 < dist) { fidx = (k + 2) % 3; t = t1; q = q1; vf = vf1; dist = dist1; }
        V3 w = x; x = y; y = z; z = w;
    }
    if (vf) { feature = TF_VERTEX; fidx = (fidx + vf) % 3; }
    else feature = TF_EDGE;
    return dist;
}

__device__ inline float sqrDistSegmentTriangle(V3 p, V3 dir, float mn, float mx, V3 a, V3 b, V3 c, float& t, V3& q, int& feature, int& fidx) {
    float sq = sqrDistLineTriangle(p, dir, a, b, c, t, q, feature, fidx);
    if (t < mn || t > mx) {          // one copy of the point routine for both ends
        t = t < mn ? mn : mx;
        sq = sqrDistPointTriangle(p + dir * t, a, b, c, q, feature, fidx);
    }
    return sq;
}

__device__ __forceinline__ float distanceAABBPlane(V3 he, V3 n, float d) {
    float r = he.x * fabsf(n.x) + he.y * fabsf(n.y) + he.z * fabsf(n.z);
    return fabsf(d) - r;
}

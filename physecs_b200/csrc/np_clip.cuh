// 2-D Sutherland-Hodgman clipping and the "clipped polygon -> <=4 contact points" generators.
//   isInside / intersection / clipEdge / suthHodgClip / clipLine   reference src/Clipping.cpp:9-71
//   generateContactsPolygonBoxFace                                 reference src/ContactsUtil.cpp:9-107
//   generateContactsPolygonPolygonFace                             reference src/ContactsUtil.cpp:109-204
// Polygons live in fixed-size per-thread arrays (the reference uses thread_local vectors; its temp
// buffer is 128 points, Clipping.cpp:6 -- here the bound is the template parameter, and exceeding it
// raises the capacity status instead of writing out of bounds).
#pragma once
#include "pb_math.cuh"

struct Manifold {
    int np;
    V3 p0[4];
    V3 p1[4];
    V3 n;
    int tri;
};

__device__ __forceinline__ bool clipInside(V2 p, V2 a, V2 b) {
    return (b.y - a.y) * (p.x - a.x) - (b.x - a.x) * (p.y - a.y) > 0.f;
}
__device__ __forceinline__ V2 clipIntersection(V2 a1, V2 b1, V2 a2, V2 b2) {
    V2 r1 = b1 - a1, r2 = b2 - a2;
    float t = det2(a1 - a2, r2) / det2(-r1, r2);
    return a1 + t * r1;
}

template <int MAXP>
struct Poly {
    V2 p[MAXP];
    int n;
    bool overflow;
};

template <int MAXP>
__device__ inline void clipEdge(Poly<MAXP>& poly, V2 a, V2 b) {
    V2 out[MAXP];
    int cnt = 0;
    int n = poly.n;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        V2 pi = poly.p[i], pj = poly.p[j];
        bool iIn = clipInside(pi, a, b), jIn = clipInside(pj, a, b);
        if (iIn && jIn) { if (cnt < MAXP) out[cnt] = pj; else poly.overflow = true; ++cnt; }
        else if (iIn) { if (cnt < MAXP) out[cnt] = clipIntersection(pi, pj, a, b); else poly.overflow = true; ++cnt; }
        else if (jIn) {
            if (cnt < MAXP) out[cnt] = clipIntersection(pi, pj, a, b); else poly.overflow = true; ++cnt;
            if (cnt < MAXP) out[cnt] = pj; else poly.overflow = true; ++cnt;
        }
    }
    if (cnt > MAXP) cnt = MAXP;
    poly.n = cnt;
    for (int i = 0; i < cnt; ++i) poly.p[i] = out[i];
}

template <int MAXP, int MAXC>
__device__ inline void suthHodgClip(Poly<MAXP>& poly, const V2* clip, int nclip) {
    for (int i = 0; i < nclip; ++i) {
        int j = (i + 1) % nclip;
        clipEdge<MAXP>(poly, clip[i], clip[j]);
    }
}

__device__ inline bool clipLine(V2& p0, V2& p1, const V2* clip, int nclip) {
    for (int i = 0; i < nclip; ++i) {
        int j = (i + 1) % nclip;
        bool in0 = clipInside(p0, clip[i], clip[j]);
        bool in1 = clipInside(p1, clip[i], clip[j]);
        if (in0 && in1) continue;
        if (in0) p1 = clipIntersection(p0, p1, clip[i], clip[j]);
        else if (in1) p0 = clipIntersection(p0, p1, clip[i], clip[j]);
        else return false;
    }
    return true;
}

// Reduce penetrating points to <= 4 (first / farthest from first / max signed area / min signed area)
// `pen` are 3-D points in the reference frame; `onRef(i)` gives the matching point on the reference plane.
template <int MAXP, class MakePair>
__device__ inline void reduceAndEmit(const V3* pen, int cnt, int clipX, int clipY, MakePair makePair, V3* c0, V3* c1, int& numPoints) {
    if (cnt > 4) {
        makePair(pen[0], c0[0], c1[0]);
        float maxD = 0.f; int maxDi = 0;
        for (int i = 0; i < cnt; ++i) {
            float d = distance2(pen[0], pen[i]);
            if (d > maxD) { maxD = d; maxDi = i; }
        }
        makePair(pen[maxDi], c0[1], c1[1]);
        float maxA = 0.f; int maxAi = 0;
        for (int i = 0; i < cnt; ++i) {
            V3 ca = pen[0] - pen[i], cb = pen[maxDi] - pen[i];
            float area = det2(mk2(get(ca, clipX), get(ca, clipY)), mk2(get(cb, clipX), get(cb, clipY)));
            if (area > maxA) { maxA = area; maxAi = i; }
        }
        makePair(pen[maxAi], c0[2], c1[2]);
        float minA = 0.f; int minAi = 0;
        for (int i = 0; i < cnt; ++i) {
            V3 da = pen[0] - pen[i], db = pen[maxDi] - pen[i];
            float area = det2(mk2(get(da, clipX), get(da, clipY)), mk2(get(db, clipX), get(db, clipY)));
            if (area < minA) { minA = area; minAi = i; }
        }
        makePair(pen[minAi], c0[3], c1[3]);
        numPoints = 4;
    } else {
        for (int i = 0; i < cnt; ++i) makePair(pen[i], c0[i], c1[i]);
        numPoints = cnt;
    }
}

// ContactsUtil.cpp:9-107.  Output: c0 = point on the box face, c1 = projected polygon point (both world space).
template <int MAXP>
__device__ inline void contactsPolygonBoxFace(V3 boxCenter, const M3& boxBasis, int boxAxis, float boxAxisSign, V3 boxHE,
                                              V3 incPlaneOrig, V3 incPlaneNormal, const Poly<MAXP>& poly, int clipX, int clipY,
                                              V3* c0, V3* c1, int& numPoints) {
    const float epsilon = 1e-4f;
    V3 pen[MAXP];
    int cnt = 0;
    V3 dir = mk3(0.f); set(dir, boxAxis, boxAxisSign);
    float heA = get(boxHE, boxAxis);
    float denom = dot(dir, incPlaneNormal);
    for (int i = 0; i < poly.n; ++i) {
        V3 p3 = mk3(0.f);
        set(p3, clipX, poly.p[i].x);
        set(p3, clipY, poly.p[i].y);
        float distance = dot(incPlaneOrig - p3, incPlaneNormal) / denom;
        if (distance < heA + epsilon) {
            set(p3, boxAxis, boxAxisSign * distance);
            pen[cnt++] = p3;
        }
    }
    auto makePair = [&](V3 q, V3& o0, V3& o1) {
        V3 onFace = q; set(onFace, boxAxis, boxAxisSign * heA);
        o0 = boxCenter + mul(boxBasis, onFace);
        o1 = boxCenter + mul(boxBasis, q);
    };
    reduceAndEmit<MAXP>(pen, cnt, clipX, clipY, makePair, c0, c1, numPoints);
}

// ContactsUtil.cpp:109-204.  Reference frame has the face normal along +Y (clipX=2, clipY=0 at every call site).
template <int MAXP>
__device__ inline void contactsPolygonPolygonFace(V3 refPos, const M3& refToWorld, V3 refPlaneOrigin, V3 refPlaneNormal,
                                                  V3 incPlaneOrigin, V3 incPlaneNormal, const Poly<MAXP>& poly, int clipX, int clipY,
                                                  V3* c0, V3* c1, int& numPoints) {
    const float epsilon = 1e-4f;
    V3 pen[MAXP];
    int cnt = 0;
    float distRef = dot(refPlaneOrigin, refPlaneNormal);
    float denom = dot(mk3(0.f, 1.f, 0.f), incPlaneNormal);
    for (int i = 0; i < poly.n; ++i) {
        V3 p3 = mk3(poly.p[i].y, 0.f, poly.p[i].x);
        float distance = dot(incPlaneOrigin - p3, incPlaneNormal) / denom;
        if (distance < distRef + epsilon) {
            p3.y = distance;
            pen[cnt++] = p3;
        }
    }
    auto makePair = [&](V3 q, V3& o0, V3& o1) {
        V3 onRef = q; onRef.y = distRef;
        o0 = refPos + mul(refToWorld, onRef);
        o1 = refPos + mul(refToWorld, q);
    };
    reduceAndEmit<MAXP>(pen, cnt, clipX, clipY, makePair, c0, c1, numPoints);
}

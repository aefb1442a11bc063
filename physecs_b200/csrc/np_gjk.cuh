// GJK / EPA bin: every pair that involves a convex mesh (and convex / box vs mesh triangles).
//
// Restates, with the same fp32 / fp64 split and the same container semantics (swap-remove faces, loose-edge list),
//   support functions            reference src/GJK.h:13-128  (convex: 4-lane max with strict '>', quirk Q13)
//   GJK with simplex return      src/GJK.h:226-292 (+ helpers :144-189)
//   EPA                          src/EPA.h:22-186 (incl. the witness-point quirk at :133, Q11, and float narrowing at :70-71)
//   manifold generators          src/Collision.cpp:352-426 (convex-convex), :488-499 (sphere-convex),
//                                :694-807 (capsule-convex), :809-887 (box-convex)
// One thread per pair; simplex / polytope live in per-thread local memory (<= 104 vertices -- 4 + the 100 iterations EPA.h:172 allows --,
// <= 256 faces, <= 64 horizon edges, clip polygons of <= 24 points: `LimFast`).  The reference's containers are unbounded
// (std::vector polytope, EPA.h:22-124; 128-point clip buffers, Clipping.cpp:6): a pair that outgrows the fast limits reports the
// cause through `ovf` (no result is stored) and is redone by a spill kernel with `LimSpill`: the polytope in global-memory scratch
// (2048 faces, 1024 horizon edges), 128-point polygons.  EPA stops after 100 insertions (EPA.h:178); a polytope that DEGENERATES --
// inconsistent visibility, unmatched horizon edges: seen once in a 250 k-body soak -- can pass even that.  The bound stays: the
// insertion is O(faces x horizon edges) on one thread, and with 8192 / 8192 one such pair held a step for 1.9 s (the reference grinds
// through the same work on a CPU core).  Past the bound the pair keeps the manifold of the truncated polytope for that step and
// PB_CAUSE_SPILL_SCRATCH is set.  Same routines, same arithmetic -- only the container bounds differ.
#pragma once
#include "np_clip.cuh"
#include "pb_ctx.h"

#define PB_STATUS_UNSUPPORTED_SHAPE 0x100
#define EPA_MAX_VERTS 104

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 mkd(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 mkd(V3 v) { return mkd((double)v.x, (double)v.y, (double)v.z); }
__device__ __forceinline__ V3 tof(D3 v) { return mk3((float)v.x, (float)v.y, (float)v.z); }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return mkd(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return mkd(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return mkd(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return mkd(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return mkd(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return mkd(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dcross(D3 x, D3 y) { return mkd(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ D3 dnormalize(D3 v) { return v * (1.0 / sqrt(ddot(v, v))); }
__device__ __forceinline__ double dsign(double x) { return (double)((0.0 < x) - (x < 0.0)); }

// ---- support functions ---------------------------------------------------------------------------------------------
struct Shape {
    int type;          // PB_SPHERE / PB_CAPSULE / PB_BOX / PB_CONVEX_MESH / 5 = triangle
    V3 pos;
    M3 basis;          // box / convex localToWorld; capsule: c[0] = axis; triangle: c[0..2] = a, b, c
    V3 prm;            // sphere r | capsule hh, r | box half extents | convex scale
    const float4* verts; int nVertsPadded;
};

// __noinline__ on the big routines of this file: GJK and EPA call support() from a dozen places, and the X-vs-convex kernels inline
// GJK, EPA and a clipping routine per shape combination -- k_np_prim<GJK> was 474 KB of SASS, whose divergent lanes (5 of 32 threads
// active per instruction) spent most of their time on instruction-cache misses (ncu: stall_no_instruction dominant).  One shared
// copy of each routine keeps the kernel inside the instruction cache; the arithmetic is unchanged.
static __device__ __noinline__ V3 support(const Shape& s, V3 dir) {
    switch (s.type) {
        case PB_SPHERE: return s.pos + dir * s.prm.x;
        case PB_CAPSULE: return s.pos + s.basis.c[0] * gsign(dot(dir, s.basis.c[0])) * s.prm.x + dir * s.prm.y;
        case PB_BOX: {
            V3 r = s.pos;
            for (int i = 0; i < 3; ++i) r += gsign(dot(dir, s.basis.c[i])) * s.basis.c[i] * get(s.prm, i);
            return r;
        }
        case PB_CONVEX_MESH: {
            V3 d = s.prm * mulT(s.basis, dir);
            float maxV[4] = { -FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
            int maxI[4] = { 0, 1, 2, 3 };
            for (int i = 0; i < s.nVertsPadded; i += 4) {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    float4 v = s.verts[i + l];
                    float dt = v.x * d.x + v.y * d.y;
                    dt = dt + v.z * d.z;
                    if (dt > maxV[l]) { maxV[l] = dt; maxI[l] = i + l; }
                }
            }
            int bi = maxI[0]; float bv = maxV[0];
#pragma unroll
            for (int l = 1; l < 4; ++l) if (maxV[l] > bv) { bv = maxV[l]; bi = maxI[l]; }
            return s.pos + mul(s.basis, s.prm * mk3(s.verts[bi]));
        }
        default: {  // triangle
            V3 a = s.basis.c[0], b = s.basis.c[1], c = s.basis.c[2];
            if (dot(dir, a) > dot(dir, b)) { if (dot(dir, a) > dot(dir, c)) return a; return c; }
            if (dot(dir, b) > dot(dir, c)) return b;
            return c;
        }
    }
}

struct GjkV { V3 pos, sp0, sp1; };   // pos is a float difference promoted to double on use (GJK.h:130-134, :185-189)

__device__ __forceinline__ GjkV minkowski(const Shape& s0, const Shape& s1, V3 dir) {
    GjkV v; v.sp0 = support(s0, dir); v.sp1 = support(s1, -dir); v.pos = v.sp0 - v.sp1; return v;
}
__device__ __forceinline__ bool gjkIsZero(V3 v) { return (double)length2(v) < 1e-6; }
__device__ __forceinline__ bool checkDirection(const Shape& s0, const Shape& s1, D3 dir, GjkV& v) {
    dir = dnormalize(dir);
    v = minkowski(s0, s1, tof(dir));
    return ddot(mkd(v.pos), dir) >= 0.0;
}
__device__ __forceinline__ bool checkFace(D3 v0, D3 v1, D3 v2, D3 opposite, D3& n) {
    n = dcross(v1 - v0, v2 - v0);
    n = dsign(ddot(v0 - opposite, n)) * n;
    return ddot(n, -v0) > (double)0.00001f;
}
__device__ __forceinline__ float sqrDistPointToLineF(V3 p, V3 a, V3 b) {
    V3 r = p - a, d = b - a;
    V3 q = a + dot(r, d) / dot(d, d) * d;
    V3 pq = p - q;
    return dot(pq, pq);
}
__device__ __forceinline__ float distPointToPlaneF(V3 p, V3 o, V3 n) { return fabsf(dot(p - o, n)); }

// arbitrary tetrahedron around a degenerate start (GJK.h:232-245 / :253-264 / :270-277): fills s[2] (optionally) and s[3]
static __device__ __noinline__ void gjkDegenerate(const Shape& s0, const Shape& s1, GjkV* s, bool needS2) {
    D3 p0 = mkd(s[0].pos), p1 = mkd(s[1].pos);
    if (needS2) {
        D3 dir = dcross(p0 - p1, mkd(0.0, 1.0, 0.0));
        if (gjkIsZero(tof(dir))) dir = dcross(p0 - p1, mkd(0.0, 0.0, 1.0));
        s[2] = minkowski(s0, s1, tof(dnormalize(dir)));
        if (sqrDistPointToLineF(s[2].pos, s[0].pos, s[1].pos) < 0.00001f) s[2] = minkowski(s0, s1, tof(-dir));
    }
    D3 p2 = mkd(s[2].pos);
    D3 dir = dnormalize(dcross(p1 - p0, p2 - p0));
    s[3] = minkowski(s0, s1, tof(dir));
    if (distPointToPlaneF(s[3].pos, s[0].pos, tof(dir)) < 0.0001f) s[3] = minkowski(s0, s1, tof(-dir));
}

static __device__ __noinline__ bool gjk(const Shape& s0, const Shape& s1, V3 startDir, GjkV* s) {
    D3 dir = mkd(normalize(startDir));
    s[0] = minkowski(s0, s1, tof(dir));
    if (gjkIsZero(s[0].pos)) {
        dir = -dir;
        s[1] = minkowski(s0, s1, tof(dir));
        gjkDegenerate(s0, s1, s, true);
        return true;
    }
    dir = -mkd(s[0].pos);
    if (!checkDirection(s0, s1, dir, s[1])) return false;
    D3 p0 = mkd(s[0].pos), p1 = mkd(s[1].pos);
    dir = dcross(dcross(p0 - p1, -p1), p0 - p1);
    if (gjkIsZero(tof(dir))) { gjkDegenerate(s0, s1, s, true); return true; }
    if (!checkDirection(s0, s1, dir, s[2])) return false;
    D3 p2 = mkd(s[2].pos);
    dir = dcross(p1 - p0, p2 - p0);
    double sign = dsign(ddot(-p0, dir));
    if (sign == 0.0) {
        dir = dnormalize(dir);
        s[3] = minkowski(s0, s1, tof(dir));
        if (distPointToPlaneF(s[3].pos, s[0].pos, tof(dir)) < 0.0001f) s[3] = minkowski(s0, s1, tof(-dir));
        return true;
    }
    dir = sign * dir;
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

// ---- EPA -------------------------------------------------------------------------------------------------------------
struct EpaFace { unsigned char i0, i1, i2; D3 n; };
template <int NF, int NL>
struct EpaT {
    static constexpr int MAX_FACES = NF, MAX_LOOSE = NL;
    GjkV v[EPA_MAX_VERTS]; int nv;
    EpaFace f[NF]; int nf;
    unsigned char le0[NL], le1[NL];     // horizon ("loose") edges of the vertex being inserted
    int overflow;                       // PB_CAUSE_SPILLED_* bits
};
// container bounds of the per-pair routines: the per-thread fast path and the spill kernels' global-memory scratch
struct LimFast  { typedef EpaT<256, 64> Epa;    static constexpr int POLY = 24; };
struct LimSpill { typedef EpaT<2048, 1024> Epa; static constexpr int POLY = 128; };   // POLY == the reference's own bound (Clipping.cpp:6)
typedef LimFast::Epa Epa;

template <class E>
__device__ __forceinline__ void epaCreateFace(E& p, int i0, int i1, int i2) {
    D3 a = mkd(p.v[i0].pos), b = mkd(p.v[i1].pos), c = mkd(p.v[i2].pos);
    D3 n = dcross(a - b, c - b);
    double len = sqrt(ddot(n, n));
    if (len) n = n / len;
    if (p.nf < E::MAX_FACES) { EpaFace& F = p.f[p.nf++]; F.i0 = (unsigned char)i0; F.i1 = (unsigned char)i1; F.i2 = (unsigned char)i2; F.n = n; }
    else p.overflow |= PB_CAUSE_SPILLED_EPA_FACES;
}

// returns the EPA normal; cp0 / cp1 = witness points (with the reference's projection quirk).  *ovf collects PB_CAUSE_SPILLED_* bits
// when a container bound was hit (the result is then not the reference's and must not be used).
template <class E>
static __device__ __noinline__ V3 epa(const Shape& s0, const Shape& s1, const GjkV* s, V3& cp0, V3& cp1, E& p, int* ovf) {
    p.nv = 4; p.nf = 0; p.overflow = 0;
    for (int i = 0; i < 4; ++i) p.v[i] = s[i];
    {
        D3 a = mkd(p.v[0].pos), b = mkd(p.v[1].pos), c = mkd(p.v[2].pos), d = mkd(p.v[3].pos);
        D3 n = dcross(a - b, c - b);
        if (ddot(d - a, n) < 0.0) { epaCreateFace(p, 0, 1, 2); epaCreateFace(p, 3, 1, 0); epaCreateFace(p, 3, 2, 1); epaCreateFace(p, 3, 0, 2); }
        else { epaCreateFace(p, 0, 2, 1); epaCreateFace(p, 3, 2, 0); epaCreateFace(p, 3, 1, 2); epaCreateFace(p, 3, 0, 1); }
    }
    float minDist = 0.f;
    D3 normal = mkd(0.0, 1.0, 0.0);
    int faceIndex = 0;
    for (int it = 0; it < 100; ++it) {
        // findClosestFace: the running minimum is a float (EPA.h:70-80)
        minDist = (float)DBL_MAX;
        for (int i = 0; i < p.nf; ++i) {
            double dist = ddot(mkd(p.v[p.f[i].i0].pos), p.f[i].n);
            if (dist < (double)minDist) { minDist = (float)dist; normal = p.f[i].n; faceIndex = i; }
        }
        GjkV v = minkowski(s0, s1, tof(normal));
        if (ddot(mkd(v.pos), normal) - (double)minDist < (double)0.00001f) break;
        // insertVertex (EPA.h:83-124)
        unsigned char* le0 = p.le0; unsigned char* le1 = p.le1;
        int nl = 0;
        D3 vp = mkd(v.pos);
        for (int i = p.nf - 1; i >= 0; --i) {
            EpaFace F = p.f[i];
            if (ddot(F.n, vp - mkd(p.v[F.i0].pos)) > 0.0) {
                unsigned char e0[3] = { F.i0, F.i1, F.i2 }, e1[3] = { F.i1, F.i2, F.i0 };
                bool found[3] = { false, false, false };
                for (int j = 0; j < 3; ++j) {
                    for (int k = nl - 1; k >= 0; --k) {
                        if (le0[k] == e1[j] && le1[k] == e0[j]) {
                            found[j] = true;
                            le0[k] = le0[nl - 1]; le1[k] = le1[nl - 1]; --nl;
                            break;
                        }
                    }
                }
                for (int j = 0; j < 3; ++j) {
                    if (found[j]) continue;
                    if (nl < E::MAX_LOOSE) { le0[nl] = e0[j]; le1[nl] = e1[j]; ++nl; } else p.overflow |= PB_CAUSE_SPILLED_EPA_LOOSE;
                }
                p.f[i] = p.f[p.nf - 1]; --p.nf;
            }
        }
        if (p.nv >= EPA_MAX_VERTS) { p.overflow |= PB_CAUSE_SPILLED_EPA_VERTS; break; }     // unreachable: 4 + 100 iterations
        int vi = p.nv;
        p.v[p.nv++] = v;
        for (int e = 0; e < nl; ++e) epaCreateFace(p, vi, le0[e], le1[e]);
    }
    *ovf |= p.overflow;
    // getClosestPoints (EPA.h:126-168)
    if (faceIndex >= p.nf) faceIndex = p.nf - 1;
    if (faceIndex < 0) { cp0 = s0.pos; cp1 = s1.pos; return tof(normal); }
    EpaFace F = p.f[faceIndex];
    GjkV A = p.v[F.i0], B = p.v[F.i1], C = p.v[F.i2];
    double dF = ddot(mkd(A.pos), F.n);
    D3 proj = mkd(F.n.x + dF, F.n.y + dF, F.n.z + dF);               // sic (quirk Q11)
    V3 v0 = tof(mkd(B.pos) - mkd(A.pos)), v1 = tof(mkd(C.pos) - mkd(A.pos)), v2 = tof(proj - mkd(A.pos));
    double d00 = (double)dot(v0, v0), d01 = (double)dot(v0, v1), d11 = (double)dot(v1, v1), d20 = (double)dot(v2, v0), d21 = (double)dot(v2, v1);
    double denom = d00 * d11 - d01 * d01;
    double u, vv, w;
    if (denom) { vv = (d20 * d11 - d21 * d01) / denom; w = (d00 * d21 - d01 * d20) / denom; u = (double)1.0f - vv - w; }
    else { w = 0; if (d00) { vv = d20 / d00; u = (double)1.0f - vv; } else { vv = 0; u = 1.0; } }
    cp0 = tof(u * mkd(A.sp0) + vv * mkd(B.sp0) + w * mkd(C.sp0));
    cp1 = tof(u * mkd(A.sp1) + vv * mkd(B.sp1) + w * mkd(C.sp1));
    return tof(normal);
}

// ---- shapes from collider rows ------------------------------------------------------------------------------------------
__device__ inline Shape makeShape(int type, float4 prm, V3 pos, Q4 ori, const PbConvexDev* convexes, int mesh) {
    Shape s; s.type = type; s.pos = pos; s.verts = nullptr; s.nVertsPadded = 0;
    s.prm = mk3(prm.x, prm.y, prm.z);
    if (type == PB_CAPSULE) { s.basis.c[0] = rotate(ori, mk3(0.f, 1.f, 0.f)); s.basis.c[1] = s.basis.c[2] = mk3(0.f); }
    else s.basis = mat3_cast(ori);
    if (type == PB_CONVEX_MESH) { s.verts = convexes[mesh].verts; s.nVertsPadded = convexes[mesh].nVertsPadded; }
    return s;
}

// face of `cm` whose scaled world normal is most aligned (wantMax) / anti-aligned with n; defaults to face 0
__device__ inline int pickConvexFace(const PbConvexDev& cm, const M3& toWorld, Q4 ori, bool useQuat, V3 scale, V3 n, bool wantMax) {
    int best = 0; float bd = 0.f;
    for (int i = 0; i < cm.nFaces; ++i) {
        V3 fn = mk3(cm.faceNormal[i]) / scale;
        V3 wn = useQuat ? rotate(ori, fn) : mul(toWorld, fn);
        float d = dot(n, normalize(wn));
        if (wantMax ? (d > bd) : (d < bd)) { bd = d; best = i; }
    }
    return best;
}

// Collision.cpp:352-412
template <class L>
static __device__ __noinline__ void convexConvexContacts(V3 pos0, Q4 or0, const PbConvexDev& m0, V3 sc0, V3 pos1, Q4 or1, const PbConvexDev& m1, V3 sc1,
                                            V3 normal, Manifold& m, int* ovf) {
    constexpr int GJK_POLY = L::POLY;
    M3 dummy;
    int f0 = pickConvexFace(m0, dummy, or0, true, sc0, normal, true);
    int f1 = pickConvexFace(m1, dummy, or1, true, sc1, normal, false);
    M3 c0ToWorld = mat3_cast(or0);
    M3 worldToC0 = transpose(c0ToWorld);
    V3 refOrigin = sc0 * mk3(m0.faceCentroid[f0]);
    int o0 = m0.faceOffsets[f0], n0 = m0.faceOffsets[f0 + 1] - o0;
    int o1 = m1.faceOffsets[f1], n1 = m1.faceOffsets[f1 + 1] - o1;
    V3 u0 = normalize(sc0 * mk3(m0.verts[m0.faceIndices[o0]]) - refOrigin);
    V3 u1 = normalize(mk3(m0.faceNormal[f0]) / sc0);
    V3 u2 = cross(u0, u1);
    M3 basis; basis.c[0] = u0; basis.c[1] = u1; basis.c[2] = u2;
    M3 c0ToRef = transpose(basis);
    M3 worldToRef = mul(c0ToRef, worldToC0);
    V2 clip[GJK_POLY];
    if (n0 > GJK_POLY || n1 > GJK_POLY) { *ovf |= PB_CAUSE_SPILLED_CLIP; m.np = 0; return; }
    for (int i = 0; i < n0; ++i) {
        V3 v = mul(c0ToRef, sc0 * mk3(m0.verts[m0.faceIndices[o0 + i]]));
        clip[i] = mk2(v.z, v.x);
    }
    Poly<GJK_POLY> poly; poly.n = n1; poly.overflow = false;
    for (int i = 0; i < n1; ++i) {
        V3 v = mul(worldToRef, pos1 + rotate(or1, sc1 * mk3(m1.verts[m1.faceIndices[o1 + i]])) - pos0);
        poly.p[i] = mk2(v.z, v.x);
    }
    suthHodgClip<GJK_POLY, GJK_POLY>(poly, clip, n0);
    if (poly.overflow) *ovf |= PB_CAUSE_SPILLED_CLIP;
    V3 incOrigin = mul(worldToRef, pos1 + rotate(or1, sc1 * mk3(m1.faceCentroid[f1])) - pos0);
    V3 incNormal = mul(worldToRef, rotate(or1, normalize(mk3(m1.faceNormal[f1]) / sc1)));
    M3 refToWorld = transpose(worldToRef);
    contactsPolygonPolygonFace<GJK_POLY>(pos0, refToWorld, refOrigin, u1, incOrigin, incNormal, poly, 2, 0, m.p0, m.p1, m.np);
}

// Collision.cpp:694-785
template <class L>
static __device__ __noinline__ void capsuleConvexContacts(V3 p0L, V3 p1L, float radius, V3 meshPos, const M3& convexToWorld, const PbConvexDev& cm, V3 sc,
                                             Manifold& m, int* ovf) {
    constexpr int GJK_POLY = L::POLY;
    Q4 qdummy;
    int f = pickConvexFace(cm, convexToWorld, qdummy, false, sc, m.n, false);
    V3 refOrigin = sc * mk3(cm.faceCentroid[f]);
    int o = cm.faceOffsets[f], n = cm.faceOffsets[f + 1] - o;
    V3 u0 = normalize(sc * mk3(cm.verts[cm.faceIndices[o]]) - refOrigin);
    V3 u1 = normalize(mk3(cm.faceNormal[f]) / sc);
    V3 u2 = cross(u0, u1);
    M3 basis; basis.c[0] = u0; basis.c[1] = u1; basis.c[2] = u2;
    M3 convexToRef = transpose(basis);
    V2 clip[GJK_POLY];
    if (n > GJK_POLY) { *ovf |= PB_CAUSE_SPILLED_CLIP; m.np = 0; return; }
    for (int i = 0; i < n; ++i) {
        V3 v = mul(convexToRef, sc * mk3(cm.verts[cm.faceIndices[o + i]]));
        clip[i] = mk2(v.z, v.x);
    }
    V3 p0r = mul(convexToRef, p0L), p1r = mul(convexToRef, p1L);
    V2 line0 = mk2(p0r.z, p0r.x), line1 = mk2(p1r.z, p1r.x);
    if (!clipLine(line0, line1, clip, n)) { m.np = 0; return; }
    M3 refToWorld = mul(convexToWorld, transpose(convexToRef));
    float distToPlane = dot(refOrigin, u1);
    float a0, a1;
    float distX = p0r.z - p1r.z, distY = p0r.x - p1r.x;
    float adx = fabsf(distX), ady = fabsf(distY);
    if (adx && adx >= ady) { a0 = gmix(p0r.y, p1r.y, (p0r.z - line0.x) / distX); a1 = gmix(p1r.y, p0r.y, (p1r.z - line1.x) / -distX); }
    else if (ady > adx) { a0 = gmix(p0r.y, p1r.y, (p0r.x - line0.y) / distY); a1 = gmix(p1r.y, p0r.y, (p1r.x - line1.y) / -distY); }
    else { a0 = p0r.y; a1 = p1r.y; }
    a0 -= radius; a1 -= radius;
    int np = 0;
    if (a0 < distToPlane) {
        V3 p = mk3(line0.y, a0, line0.x); V3 onFace = p; onFace.y = distToPlane;
        m.p0[np] = meshPos + mul(refToWorld, p); m.p1[np] = meshPos + mul(refToWorld, onFace); ++np;
    }
    if (a1 < distToPlane) {
        V3 p = mk3(line1.y, a1, line1.x); V3 onFace = p; onFace.y = distToPlane;
        m.p0[np] = meshPos + mul(refToWorld, p); m.p1[np] = meshPos + mul(refToWorld, onFace); ++np;
    }
    m.np = np;
}

// Collision.cpp:809-871
template <class L>
static __device__ __noinline__ void boxConvexContacts(V3 boxCenter, const M3& boxBasis, V3 he, V3 cPos, Q4 cOr, const PbConvexDev& cm, V3 sc, V3 normal,
                                         Manifold& m, int* ovf) {
    constexpr int GJK_POLY = L::POLY;
    int boxAxis = 0; float boxAxisSign = 0.f, maxDot = 0.f;
    for (int i = 0; i < 3; ++i) {
        float d = dot(normal, boxBasis.c[i]);
        float ad = fabsf(d);
        if (ad > maxDot) { maxDot = ad; boxAxis = i; boxAxisSign = d < 0.f ? -1.f : 1.f; }
    }
    M3 dummy;
    int f = pickConvexFace(cm, dummy, cOr, true, sc, normal, false);
    int o = cm.faceOffsets[f], n = cm.faceOffsets[f + 1] - o;
    int clipX = (boxAxis + 1) % 3, clipY = (boxAxis + 2) % 3;
    M3 worldToBox = transpose(boxBasis);
    if (n > GJK_POLY) { *ovf |= PB_CAUSE_SPILLED_CLIP; m.np = 0; return; }
    Poly<GJK_POLY> poly; poly.n = n; poly.overflow = false;
    for (int i = 0; i < n; ++i) {
        V3 v = mul(worldToBox, cPos + rotate(cOr, sc * mk3(cm.verts[cm.faceIndices[o + i]])) - boxCenter);
        poly.p[i] = mk2(get(v, clipX), get(v, clipY));
    }
    float hx = get(he, clipX), hy = get(he, clipY);
    V2 clip[4] = { mk2(hx, hy), mk2(hx, -hy), mk2(-hx, -hy), mk2(-hx, hy) };
    suthHodgClip<GJK_POLY, 4>(poly, clip, 4);
    if (poly.overflow) *ovf |= PB_CAUSE_SPILLED_CLIP;
    V3 incOrig = mul(worldToBox, cPos + rotate(cOr, sc * mk3(cm.faceCentroid[f])) - boxCenter);
    V3 incNormal = normalize(mul(worldToBox, rotate(cOr, mk3(cm.faceNormal[f]) / sc)));
    contactsPolygonBoxFace<GJK_POLY>(boxCenter, boxBasis, boxAxis, boxAxisSign, he, incOrig, incNormal, poly, clipX, clipY, m.p0, m.p1, m.np);
}

// Dispatch for the GJK bin (Collision.cpp:926-1013), in two stages so that a step can run them as two launches with the
// intersecting pairs compacted in between (narrowphase.cu): in a pile most candidate pairs do not intersect, and lanes that left
// after GJK would otherwise idle through the EPA expansion and the face clipping of their warp mates.
struct GjkPair {
    int ta, tb, ma, mb; float4 qa, qb; V3 pa, pb; Q4 oa, ob; bool flip;
};
__device__ __forceinline__ GjkPair gjkPairSetup(int t0, float4 q0, V3 pos0, Q4 or0, int mesh0, int t1, float4 q1, V3 pos1, Q4 or1, int mesh1) {
    GjkPair g;
    // the convex mesh is always argument 1 of the reference routine unless both are convex
    g.flip = (t0 == PB_CONVEX_MESH && t1 != PB_CONVEX_MESH);
    g.ta = g.flip ? t1 : t0; g.tb = g.flip ? t0 : t1;
    g.qa = g.flip ? q1 : q0; g.qb = g.flip ? q0 : q1;
    g.pa = g.flip ? pos1 : pos0; g.pb = g.flip ? pos0 : pos1;
    g.oa = g.flip ? or1 : or0; g.ob = g.flip ? or0 : or1;
    g.ma = g.flip ? mesh1 : mesh0; g.mb = g.flip ? mesh0 : mesh1;
    return g;
}
// stage 1: do the shapes intersect?  (simplex out)
__device__ inline bool gjkPairIntersect(const GjkPair& g, const PbConvexDev* convexes, GjkV* simplex) {
    Shape sa = makeShape(g.ta, g.qa, g.pa, g.oa, convexes, g.ma);
    Shape sb = makeShape(g.tb, g.qb, g.pb, g.ob, convexes, g.mb);
    return gjk(sa, sb, g.pb - g.pa, simplex);
}
// stage 2: penetration by EPA, then the contact patch of the shape combination.  `poly` = polytope scratch (local memory for LimFast,
// global-memory scratch for LimSpill); *ovf != 0 afterwards: a container bound was hit, m is not to be used.
template <class L>
__device__ inline void gjkPairManifold(const GjkPair& g, const PbConvexDev* convexes, const GjkV* simplex, Manifold& m, typename L::Epa& poly, int* ovf) {
    Shape sa = makeShape(g.ta, g.qa, g.pa, g.oa, convexes, g.ma);
    Shape sb = makeShape(g.tb, g.qb, g.pb, g.ob, convexes, g.mb);
    const int ta = g.ta, ma = g.ma, mb = g.mb;
    const float4 qa = g.qa, qb = g.qb;
    const V3 pa = g.pa, pb = g.pb;
    const Q4 oa = g.oa, ob = g.ob;
    V3 cp0, cp1;
    m.n = epa(sa, sb, simplex, cp0, cp1, poly, ovf);
    m.p0[0] = cp0; m.p1[0] = cp1;
    V3 scb = mk3(qb.x, qb.y, qb.z);
    const PbConvexDev& cb = convexes[mb];
    if (ta == PB_SPHERE) { m.np = 1; return; }
    if (ta == PB_CAPSULE) {
        M3 convexToWorld = mat3_cast(ob);
        M3 worldToConvex = inverse(convexToWorld);
        V3 capsuleAxis = rotate(oa, mk3(0.f, 1.f, 0.f));
        V3 p = mul(worldToConvex, pa - pb);
        V3 axisLc = mul(worldToConvex, capsuleAxis);
        Manifold t = m;
        capsuleConvexContacts<L>(p + axisLc * qa.x, p - axisLc * qa.x, qa.y, pb, convexToWorld, cb, scb, t, ovf);
        if (t.np) { m = t; } else m.np = 1;
        return;
    }
    if (ta == PB_BOX) {
        Manifold t = m;
        boxConvexContacts<L>(pa, mat3_cast(oa), mk3(qa.x, qa.y, qa.z), pb, ob, cb, scb, m.n, t, ovf);
        if (t.np) { m = t; } else m.np = 1;
        return;
    }
    // convex - convex
    Manifold t = m;
    convexConvexContacts<L>(pa, oa, convexes[ma], mk3(qa.x, qa.y, qa.z), pb, ob, cb, scb, m.n, t, ovf);
    if (t.np) { m = t; } else m.np = 1;
}

// both stages in one call (scene queries, where the handful of pairs does not warrant two launches; the spill kernel)
template <class L>
__device__ inline bool collideGjkPair(int t0, float4 q0, V3 pos0, Q4 or0, int mesh0, int t1, float4 q1, V3 pos1, Q4 or1, int mesh1,
                                      const PbConvexDev* convexes, Manifold& m, bool& flip, typename L::Epa& poly, int* ovf) {
    GjkPair g = gjkPairSetup(t0, q0, pos0, or0, mesh0, t1, q1, pos1, or1, mesh1);
    flip = g.flip;
    GjkV simplex[4];
    if (!gjkPairIntersect(g, convexes, simplex)) return false;
    gjkPairManifold<L>(g, convexes, simplex, m, poly, ovf);
    return true;
}

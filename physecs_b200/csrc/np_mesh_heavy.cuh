// Box and convex mesh vs static triangle mesh: per-triangle tests and manifold generators.
//   collisionBoxTriangle                     reference src/CollisionTriangleMesh.cpp:221-456 (13-axis SAT, edge axes first, quirk Q15)
//   generateContactsBoxTriangleFace          :458-530     generateContactsBoxFaceTriangle       :532-571
//   generateContactsBoxEdgeTriangleEdge      :573-601     generateContactsBoxTriangleEdge       :603-610
//   collisionConvexMeshTriangle              :614-699 (GJK/EPA vs TriangleSupportFunction + barycentric feature classification)
//   generateContactsConvexMeshTriangleFace   :701-755     generateContactsConvexFaceTriangle    :757-824
// (the two convex generators leave `convexFaceIndex` uninitialised in the reference, quirk Q24: face 0 is used here)
#pragma once
#include "np_mesh.cuh"
#include "np_gjk.cuh"

enum { BOXF_FACE = 0, BOXF_EDGE = 1, BOXF_UNKNOWN = 2 };

__device__ __forceinline__ float gmax3(float a, float b, float c) { return gmax(gmax(a, b), c); }
__device__ __forceinline__ float gmin3(float a, float b, float c) { return gmin(gmin(a, b), c); }

__device__ inline bool boxTriangle(V3 pos, Q4 ori, V3 he, V3 a, V3 b, V3 c, V3 n, TriContact& tc) {
    float edgeOffset = 0.f;
    const float edgeLimit = 0.999f;
    M3 u = mat3_cast(ori);
    M3 invU = transpose(u);
    V3 v0 = mul(invU, a - pos), v1 = mul(invU, b - pos), v2 = mul(invU, c - pos);
    V3 f0 = v1 - v0, f1 = v2 - v1, f2 = v0 - v2;
    float f0Limit = edgeLimit * length2(f0), f1Limit = edgeLimit * length2(f1), f2Limit = edgeLimit * length2(f2);
    float mx = -FLT_MAX;
    float p0, p1, r, d;
#define EDGE_AXIS(P0, P1, R, COND, FIDX, AXIS) \
    p0 = (P0); p1 = (P1); r = (R); d = gmax(-gmax(p0, p1), gmin(p0, p1)) - r; \
    if (d > 0.f) return false; \
    if ((COND) && d > mx) { mx = d; tc.feature = TF_EDGE; tc.fidx = (FIDX); tc.dist = d; tc.boxFeature = BOXF_EDGE; tc.boxAxis = (AXIS); edgeOffset = 0.1f; }
    // u0 x f0, f1, f2
    EDGE_AXIS(v0.z * f0.y - v0.y * f0.z, v2.z * f0.y - v2.y * f0.z, he.y * fabsf(f0.z) + he.z * fabsf(f0.y), f0.x * f0.x < f0Limit, 2, 0)
    EDGE_AXIS(v0.z * f1.y - v0.y * f1.z, v1.z * f1.y - v1.y * f1.z, he.y * fabsf(f1.z) + he.z * fabsf(f1.y), f1.x * f1.x < f1Limit, 0, 0)
    EDGE_AXIS(v1.z * f2.y - v1.y * f2.z, v2.z * f2.y - v2.y * f2.z, he.y * fabsf(f2.z) + he.z * fabsf(f2.y), f2.x * f2.x < f2Limit, 1, 0)
    // u1 x f0, f1, f2
    EDGE_AXIS(v0.x * f0.z - v0.z * f0.x, v2.x * f0.z - v2.z * f0.x, he.x * fabsf(f0.z) + he.z * fabsf(f0.x), f0.y * f0.y < f0Limit, 2, 1)
    EDGE_AXIS(v0.x * f1.z - v0.z * f1.x, v1.x * f1.z - v1.z * f1.x, he.x * fabsf(f1.z) + he.z * fabsf(f1.x), f1.y * f1.y < f1Limit, 0, 1)
    EDGE_AXIS(v1.x * f2.z - v1.z * f2.x, v2.x * f2.z - v2.z * f2.x, he.x * fabsf(f2.z) + he.z * fabsf(f2.x), f2.y * f2.y < f2Limit, 1, 1)
    // u2 x f0, f1, f2
    EDGE_AXIS(v0.y * f0.x - v0.x * f0.y, v2.y * f0.x - v2.x * f0.y, he.x * fabsf(f0.y) + he.y * fabsf(f0.x), f0.z * f0.z < f0Limit, 2, 2)
    EDGE_AXIS(v0.y * f1.x - v0.x * f1.y, v1.y * f1.x - v1.x * f1.y, he.x * fabsf(f1.y) + he.y * fabsf(f1.x), f1.z * f1.z < f1Limit, 0, 2)
    EDGE_AXIS(v1.y * f2.x - v1.x * f2.y, v2.y * f2.x - v2.x * f2.y, he.x * fabsf(f2.y) + he.y * fabsf(f2.x), f2.z * f2.z < f2Limit, 1, 2)
#undef EDGE_AXIS
    // box face normals
    for (int i = 0; i < 3; ++i) {
        float a0 = get(v0, i), a1 = get(v1, i), a2 = get(v2, i), h = get(he, i);
        d = gmax(-gmax3(a0, a1, a2), gmin3(a0, a1, a2)) - h;
        if (d > 0.f) return false;
        if (d > mx - edgeOffset) {
            mx = d;
            bool in0 = fabsf(a0) < h, in1 = fabsf(a1) < h, in2 = fabsf(a2) < h;
            if (in0) {
                if (in1) { if (in2) tc.feature = TF_FACE; else { tc.feature = TF_EDGE; tc.fidx = 2; } }
                else { if (in2) { tc.feature = TF_EDGE; tc.fidx = 1; } else { tc.feature = TF_VERTEX; tc.fidx = 0; } }
            } else {
                if (in1) { if (in2) { tc.feature = TF_EDGE; tc.fidx = 0; } else { tc.feature = TF_VERTEX; tc.fidx = 1; } }
                else { tc.feature = TF_VERTEX; tc.fidx = 2; }
            }
            tc.boxFeature = BOXF_FACE; tc.boxAxis = i;
            tc.dist = d;
            edgeOffset = 0.f;
        }
    }
    // triangle normal
    V3 nLocal = mul(invU, n);
    d = distanceAABBPlane(he, nLocal, dot(nLocal, v0));
    if (d > 0.f) return false;
    if (d > mx - edgeOffset) { tc.feature = TF_FACE; tc.dist = d; tc.boxFeature = BOXF_UNKNOWN; }
    return true;
}

struct TriFrame { V3 refOrigin, u0, u1, u2; M3 meshToRef; V3 va[3]; int4 ti; V3 n; };
__device__ inline TriFrame triFrame(const PbTriMeshDev& mesh, int tri, bool normalizeU1) {
    TriFrame f;
    f.ti = mesh.tris[tri];
    f.refOrigin = mk3(mesh.triCentroid[tri]);
    f.n = mk3(mesh.triNormal[tri]);
    f.va[0] = mk3(mesh.verts[f.ti.x]); f.va[1] = mk3(mesh.verts[f.ti.y]); f.va[2] = mk3(mesh.verts[f.ti.z]);
    f.u0 = normalize(f.va[0] - f.refOrigin);
    f.u1 = normalizeU1 ? normalize(f.n) : f.n;
    f.u2 = cross(f.u0, f.u1);
    M3 basis; basis.c[0] = f.u0; basis.c[1] = f.u1; basis.c[2] = f.u2;
    f.meshToRef = transpose(basis);
    return f;
}

// CTM.cpp:458-530 (triangle is the reference face, incident box face clipped against it)
__device__ inline void boxTriangleFaceManifold(V3 boxPos, Q4 boxOr, V3 he, V3 meshPos, Q4 meshOr, const PbTriMeshDev& mesh, const TriContact& tc, Manifold& m) {
    M3 boxBasis = mat3_cast(boxOr);
    TriFrame f = triFrame(mesh, tc.tri, true);
    V2 clip[3];
    for (int i = 0; i < 3; ++i) { V3 v = mul(f.meshToRef, f.va[i]); clip[3 - i - 1] = mk2(v.z, v.x); }
    int incAxis = 0; float incSign = 1.f;
    if (tc.boxFeature == BOXF_FACE) {
        incAxis = tc.boxAxis;
        incSign = dot(f.n, boxBasis.c[incAxis]) > 0.f ? -1.f : 1.f;
    } else {
        float maxDot = 0.f;
        for (int i = 0; i < 3; ++i) {
            float d = dot(f.n, boxBasis.c[i]);
            float ad = fabsf(d);
            if (ad > maxDot) { maxDot = ad; incAxis = i; incSign = d > 0.f ? -1.f : 1.f; }
        }
    }
    int ia1 = (incAxis + 1) % 3, ia2 = (incAxis + 2) % 3;
    V3 incPlaneOrig = mul(f.meshToRef, boxPos + boxBasis.c[incAxis] * get(he, incAxis) * incSign);
    V3 e1 = boxBasis.c[ia1] * get(he, ia1), e2 = boxBasis.c[ia2] * get(he, ia2), ne1 = -boxBasis.c[ia1] * get(he, ia1);
    V3 q0 = incPlaneOrig + mul(f.meshToRef, e1 + e2);
    V3 q1 = incPlaneOrig + mul(f.meshToRef, ne1 + e2);
    V3 q2 = incPlaneOrig + mul(f.meshToRef, ne1 - e2);
    V3 q3 = incPlaneOrig + mul(f.meshToRef, e1 - e2);
    Poly<8> poly; poly.n = 4; poly.overflow = false;
    poly.p[0] = mk2(q0.z, q0.x); poly.p[1] = mk2(q1.z, q1.x); poly.p[2] = mk2(q2.z, q2.x); poly.p[3] = mk2(q3.z, q3.x);
    suthHodgClip<8, 3>(poly, clip, 3);
    V3 incPlaneNormal = mul(f.meshToRef, boxBasis.c[incAxis]);
    M3 meshToWorld = mat3_cast(meshOr);
    M3 refToWorld = mul(meshToWorld, transpose(f.meshToRef));
    m.n = mul(-1.f * meshToWorld, f.n);
    V3 c0[4], c1[4];
    contactsPolygonPolygonFace<8>(meshPos, refToWorld, f.refOrigin, f.u1, incPlaneOrig, incPlaneNormal, poly, 2, 0, c0, c1, m.np);
    for (int i = 0; i < m.np; ++i) { m.p0[i] = c1[i]; m.p1[i] = c0[i]; }
}

// CTM.cpp:532-571 (box face is the reference, triangle clipped against it)
__device__ inline void boxFaceTriangleManifold(V3 boxPos, Q4 boxOr, V3 he, V3 meshPos, Q4 meshOr, const PbTriMeshDev& mesh, const TriContact& tc, Manifold& m) {
    M3 boxBasis = mat3_cast(boxOr);
    M3 meshToBox = transpose(boxBasis);
    int4 ti = mesh.tris[tc.tri];
    V3 centroid = mk3(mesh.triCentroid[tc.tri]), triN = mk3(mesh.triNormal[tc.tri]);
    int refAxis = tc.boxAxis;
    float refSign = dot(boxBasis.c[refAxis], centroid - boxPos) < 0.f ? -1.f : 1.f;
    int clipX = (refAxis + 1) % 3, clipY = (refAxis + 2) % 3;
    Poly<8> poly; poly.n = 3; poly.overflow = false;
    int idx[3] = { ti.x, ti.y, ti.z };
    for (int i = 0; i < 3; ++i) {
        V3 v = mul(meshToBox, mk3(mesh.verts[idx[i]]) - boxPos);
        poly.p[i] = mk2(get(v, clipX), get(v, clipY));
    }
    float hx = get(he, clipX), hy = get(he, clipY);
    V2 clip[4] = { mk2(hx, hy), mk2(hx, -hy), mk2(-hx, -hy), mk2(-hx, hy) };
    suthHodgClip<8, 4>(poly, clip, 4);
    V3 incOrigin = mul(meshToBox, centroid - boxPos);
    V3 incNormal = mul(meshToBox, triN);
    M3 meshToWorld = mat3_cast(meshOr);
    m.n = mul(meshToWorld, refSign * boxBasis.c[refAxis]);
    V3 c0[4], c1[4];
    contactsPolygonBoxFace<8>(boxPos, boxBasis, refAxis, refSign, he, incOrigin, incNormal, poly, clipX, clipY, c0, c1, m.np);
    for (int i = 0; i < m.np; ++i) { m.p0[i] = meshPos + mul(meshToWorld, c0[i]); m.p1[i] = meshPos + mul(meshToWorld, c1[i]); }
}

// CTM.cpp:573-601
__device__ inline void boxEdgeTriangleEdgeManifold(V3 boxPos, Q4 boxOr, V3 he, V3 meshPos, Q4 meshOr, const PbTriMeshDev& mesh, const TriContact& tc, Manifold& m) {
    M3 boxBasis = mat3_cast(boxOr);
    int4 ti = mesh.tris[tc.tri];
    int idx[3] = { ti.x, ti.y, ti.z };
    V3 centroid = mk3(mesh.triCentroid[tc.tri]);
    int boxAxis = tc.boxAxis, triEdgeI = tc.fidx;
    V3 triA = mk3(mesh.verts[idx[(triEdgeI + 1) % 3]]), triB = mk3(mesh.verts[idx[(triEdgeI + 2) % 3]]);
    V3 triEdge = triB - triA;
    V3 axis = cross(boxBasis.c[boxAxis], triEdge);
    M3 meshToWorld = mat3_cast(meshOr);
    m.n = mul(meshToWorld, normalize(dot(axis, centroid - boxPos) < 0.f ? -axis : axis));
    m.np = 1;
    V3 sb = mk3(-1.f);
    set(sb, (boxAxis + 1) % 3, dot(boxBasis.c[(boxAxis + 1) % 3], m.n) < 0.f ? -1.f : 1.f);
    set(sb, (boxAxis + 2) % 3, dot(boxBasis.c[(boxAxis + 2) % 3], m.n) < 0.f ? -1.f : 1.f);
    V3 boxA = boxPos + sb.x * he.x * boxBasis.c[0] + sb.y * he.y * boxBasis.c[1] + sb.z * he.z * boxBasis.c[2];
    V3 boxEdge = boxBasis.c[boxAxis] * get(he, boxAxis) * 2.f;
    V3 p0, p1;
    closestPointsSegSeg(boxA, boxEdge, triA, triEdge, p0, p1);
    m.p0[0] = meshPos + mul(meshToWorld, p0); m.p1[0] = meshPos + mul(meshToWorld, p1);
}

// ---- convex mesh vs triangle ---------------------------------------------------------------------------------------------
template <class E>
__device__ inline bool convexTriangle(const Shape& convex, V3 a, V3 b, V3 c, V3 centroid, TriContact& tc, E& scratch, int* ovf) {
    Shape tri; tri.type = 5; tri.pos = centroid; tri.basis.c[0] = a; tri.basis.c[1] = b; tri.basis.c[2] = c; tri.prm = mk3(0.f); tri.verts = nullptr; tri.nVertsPadded = 0;
    GjkV s[4];
    if (!gjk(convex, tri, centroid - convex.pos, s)) return false;
    tc.normal = epa(convex, tri, s, tc.cpBody, tc.cpTri, scratch, ovf);
    tc.dist = dot(tc.cpTri - tc.cpBody, tc.normal);
    V3 v0 = b - a, v1 = c - a, v2 = tc.cpTri - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    float u, v, w;
    if (denom) { v = (d20 * d11 - d21 * d01) / denom; w = (d00 * d21 - d01 * d20) / denom; u = 1.0f - v - w; }
    else { w = 0.f; if (d00) { v = d20 / d00; u = 1.0f - v; } else { v = 0.f; u = 1.0f; } }
    const float eps = (float)1e-4;
    if (u < eps) {
        if (v < eps) { tc.feature = TF_VERTEX; tc.fidx = 2; }
        else if (w < eps) { tc.feature = TF_VERTEX; tc.fidx = 1; }
        else { tc.feature = TF_EDGE; tc.fidx = 0; }
    } else {
        if (v < eps) {
            if (w < eps) { tc.feature = TF_VERTEX; tc.fidx = 0; } else { tc.feature = TF_EDGE; tc.fidx = 1; }
        } else if (w < eps) {
            if (u < eps) { tc.feature = TF_VERTEX; tc.fidx = 0; } else { tc.feature = TF_EDGE; tc.fidx = 2; }
        } else tc.feature = TF_FACE;
    }
    return true;
}

// CTM.cpp:701-755
template <class L>
__device__ inline void convexTriangleFaceManifold(V3 cPos, Q4 cOr, const PbConvexDev& cm, V3 sc, V3 meshPos, Q4 meshOr, const PbTriMeshDev& mesh,
                                                  const TriContact& tc, Manifold& m, int* ovf) {
    constexpr int GJK_POLY = L::POLY;
    TriFrame f = triFrame(mesh, tc.tri, false);
    V2 clip[3];
    for (int i = 0; i < 3; ++i) { V3 v = mul(f.meshToRef, f.va[i]); clip[3 - i - 1] = mk2(v.z, v.x); }
    M3 dummy;
    int face = pickConvexFace(cm, dummy, cOr, true, sc, f.n, false);
    int o = cm.faceOffsets[face], n = cm.faceOffsets[face + 1] - o;
    if (n > GJK_POLY) { *ovf |= PB_CAUSE_SPILLED_CLIP; m.np = 0; return; }
    Poly<GJK_POLY> poly; poly.n = n; poly.overflow = false;
    for (int i = 0; i < n; ++i) {
        V3 v = mul(f.meshToRef, cPos + rotate(cOr, sc * mk3(cm.verts[cm.faceIndices[o + i]])));
        poly.p[i] = mk2(v.z, v.x);
    }
    suthHodgClip<GJK_POLY, 3>(poly, clip, 3);
    if (poly.overflow) *ovf |= PB_CAUSE_SPILLED_CLIP;
    V3 incOrigin = mul(f.meshToRef, cPos + rotate(cOr, sc * mk3(cm.faceCentroid[face])));
    V3 incNormal = mul(f.meshToRef, rotate(cOr, normalize(mk3(cm.faceNormal[face]) / sc)));
    M3 meshToWorld = mat3_cast(meshOr);
    M3 refToWorld = mul(meshToWorld, transpose(f.meshToRef));
    m.n = mul(-1.f * meshToWorld, f.n);
    V3 c0[4], c1[4];
    contactsPolygonPolygonFace<GJK_POLY>(meshPos, refToWorld, f.refOrigin, f.u1, incOrigin, incNormal, poly, 2, 0, c0, c1, m.np);
    for (int i = 0; i < m.np; ++i) { m.p0[i] = c1[i]; m.p1[i] = c0[i]; }
}

// CTM.cpp:757-824
template <class L>
__device__ inline void convexFaceTriangleManifold(V3 cPos, Q4 cOr, const PbConvexDev& cm, V3 sc, V3 meshPos, Q4 meshOr, const PbTriMeshDev& mesh,
                                                  const TriContact& tc, Manifold& m, int* ovf) {
    constexpr int GJK_POLY = L::POLY;
    int4 ti = mesh.tris[tc.tri];
    int idx[3] = { ti.x, ti.y, ti.z };
    V3 centroid = mk3(mesh.triCentroid[tc.tri]), triN = mk3(mesh.triNormal[tc.tri]);
    M3 convexToMesh = mat3_cast(cOr);
    M3 meshToConvex = transpose(convexToMesh);
    M3 dummy;
    int face = pickConvexFace(cm, dummy, cOr, true, sc, tc.normal, true);
    int o = cm.faceOffsets[face], n = cm.faceOffsets[face + 1] - o;
    V3 refOrigin = sc * mk3(cm.faceCentroid[face]);
    V3 u0 = normalize(sc * mk3(cm.verts[cm.faceIndices[o]]) - refOrigin);
    V3 u1 = normalize(mk3(cm.faceNormal[face]) / sc);
    V3 u2 = cross(u0, u1);
    M3 basis; basis.c[0] = u0; basis.c[1] = u1; basis.c[2] = u2;
    M3 convexToRef = transpose(basis);
    M3 meshToRef = mul(convexToRef, meshToConvex);
    if (n > GJK_POLY) { *ovf |= PB_CAUSE_SPILLED_CLIP; m.np = 0; return; }
    V2 clip[GJK_POLY];
    for (int i = 0; i < n; ++i) { V3 v = mul(convexToRef, sc * mk3(cm.verts[cm.faceIndices[o + i]])); clip[i] = mk2(v.z, v.x); }
    Poly<GJK_POLY> poly; poly.n = 3; poly.overflow = false;
    for (int i = 0; i < 3; ++i) { V3 v = mul(meshToRef, mk3(mesh.verts[idx[i]]) - cPos); poly.p[i] = mk2(v.z, v.x); }
    suthHodgClip<GJK_POLY, GJK_POLY>(poly, clip, n);
    if (poly.overflow) *ovf |= PB_CAUSE_SPILLED_CLIP;
    V3 incOrigin = mul(meshToRef, centroid - cPos);
    V3 incNormal = mul(meshToRef, triN);
    M3 meshToWorld = mat3_cast(meshOr);
    M3 refToMesh = transpose(meshToRef);
    V3 c0[4], c1[4];
    contactsPolygonPolygonFace<GJK_POLY>(cPos, refToMesh, refOrigin, u1, incOrigin, incNormal, poly, 2, 0, c0, c1, m.np);
    if (m.np) {
        for (int i = 0; i < m.np; ++i) { m.p0[i] = meshPos + mul(meshToWorld, c0[i]); m.p1[i] = meshPos + mul(meshToWorld, c1[i]); }
        m.n = mul(mul(meshToWorld, convexToMesh), u1);
    } else {
        m.n = mul(meshToWorld, tc.normal);
        m.np = 1;
        m.p0[0] = meshPos + mul(meshToWorld, tc.cpBody); m.p1[0] = meshPos + mul(meshToWorld, tc.cpTri);
    }
}

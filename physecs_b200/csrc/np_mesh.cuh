// Shape vs static triangle mesh: mesh-BVH cull, per-triangle tests, the reference's two-pass feature
// filter ("voided vertices") and one manifold per surviving triangle.
//   collisionTriangleMesh            reference src/CollisionTriangleMesh.cpp:882-956
//   TriangleMesh::overlapBvh         reference src/TriangleMesh.cpp:166-192 (stack order: left child first)
//   collisionSphereTriangle          CTM.cpp:44-71     generateContactsSphereTriangle       :73-81
//   collisionCapsuleTriangle         CTM.cpp:85-119    generateContactsCapsuleTriangleFace  :121-217
// Everything is evaluated in the mesh's local frame and moved to world space at the end, as the reference does.
#pragma once
#include "np_geom.cuh"
#include "np_clip.cuh"
#include "np_bounds.cuh"
#include "pb_ctx.h"

struct TriContact {
    int tri;
    V3 normal, cpBody, cpTri;
    int feature, fidx;
    float dist;
    int boxFeature, boxAxis;   // box only
};

__device__ inline bool sphereTriangle(V3 pos0, float r0, V3 a, V3 b, V3 c, V3 n, TriContact& tc) {
    V3 u; int feature, fidx = 0;
    float sq = sqrDistPointTriangle(pos0, a, b, c, u, feature, fidx);
    if (sq > r0 * r0) return false;
    if (feature == TF_FACE) tc.normal = -n;
    else {
        V3 normal = u - pos0;
        float len = length(normal);
        if (len) tc.normal = normal / len;
        else tc.normal = -n;
    }
    tc.cpBody = pos0 + tc.normal * r0;
    tc.cpTri = u;
    tc.feature = feature; tc.fidx = fidx; tc.dist = sq;
    return true;
}

__device__ inline bool capsuleTriangle(V3 pos, Q4 ori, float hh, float radius, V3 a, V3 b, V3 c, V3 n, TriContact& tc) {
    V3 dir = rotate(ori, mk3(0.f, 1.f, 0.f));
    float t; V3 u; int feature, fidx = 0;
    float sq = sqrDistSegmentTriangle(pos, dir, -hh, hh, a, b, c, t, u, feature, fidx);
    if (sq > radius * radius) return false;
    V3 onSeg = pos + dir * t;
    if (feature == TF_FACE) tc.normal = -n;
    else {
        V3 normal = u - onSeg;
        float len = length(normal);
        if (len) tc.normal = normal / len;
        else tc.normal = -n;
    }
    tc.cpBody = onSeg + tc.normal * radius;
    tc.cpTri = u;
    tc.feature = feature; tc.fidx = fidx; tc.dist = sq;
    return true;
}

// world-space manifold from a single closest-point contact (sphere: all features; capsule: edge / vertex)
__device__ inline void manifoldFromClosest(V3 pos1, Q4 or1, const TriContact& tc, Manifold& m) {
    m.n = rotate(or1, tc.normal);
    m.np = 1;
    m.p0[0] = pos1 + rotate(or1, tc.cpBody);
    m.p1[0] = pos1 + rotate(or1, tc.cpTri);
    m.tri = tc.tri;
}

// CTM.cpp:121-217; returns false when the clipped capsule segment misses the triangle (no manifold is produced)
__device__ inline bool capsuleTriangleFaceManifold(V3 pos0, Q4 or0, float hh0, float r0, V3 pos1, Q4 or1, const PbTriMeshDev& mesh,
                                                   const TriContact& tc, Manifold& m) {
    V3 dir = rotate(or0, mk3(0.f, 1.f, 0.f));
    V3 p0l = pos0 - dir * hh0;
    V3 p1l = pos0 + dir * hh0;
    int4 ti = mesh.tris[tc.tri];
    V3 refOrigin = mk3(mesh.triCentroid[tc.tri]);
    V3 triN = mk3(mesh.triNormal[tc.tri]);
    V3 va[3] = { mk3(mesh.verts[ti.x]), mk3(mesh.verts[ti.y]), mk3(mesh.verts[ti.z]) };
    V3 u0 = normalize(va[0] - refOrigin);
    V3 u1 = normalize(triN);
    V3 u2 = cross(u0, u1);
    M3 basis; basis.c[0] = u0; basis.c[1] = u1; basis.c[2] = u2;
    M3 meshToRef = transpose(basis);
    V2 clip[3];
    for (int i = 0; i < 3; ++i) {
        V3 v = mul(meshToRef, va[i]);
        clip[3 - i - 1] = mk2(v.z, v.x);
    }
    V3 p0r = mul(meshToRef, p0l), p1r = mul(meshToRef, p1l);
    V2 line0 = mk2(p0r.z, p0r.x), line1 = mk2(p1r.z, p1r.x);
    if (!clipLine(line0, line1, clip, 3)) return false;
    M3 meshToWorld = mat3_cast(or1);
    M3 refToWorld = mul(meshToWorld, transpose(meshToRef));
    float distToPlane = dot(refOrigin, u1);
    float a0, a1;
    float distX = p0r.z - p1r.z, distY = p0r.x - p1r.x;
    float adx = fabsf(distX), ady = fabsf(distY);
    if (adx && adx >= ady) {
        a0 = gmix(p0r.y, p1r.y, (p0r.z - line0.x) / distX);
        a1 = gmix(p1r.y, p0r.y, (p1r.z - line1.x) / -distX);
    } else if (ady > adx) {
        a0 = gmix(p0r.y, p1r.y, (p0r.x - line0.y) / distY);
        a1 = gmix(p1r.y, p0r.y, (p1r.x - line1.y) / -distY);
    } else { a0 = p0r.y; a1 = p1r.y; }
    a0 -= r0; a1 -= r0;
    int np = 0;
    m.tri = tc.tri;
    m.n = mul(meshToWorld, tc.normal);
    if (a0 < distToPlane) {
        V3 p = mk3(line0.y, a0, line0.x);
        V3 onFace = p; onFace.y = distToPlane;
        m.p0[np] = pos1 + mul(refToWorld, p); m.p1[np] = pos1 + mul(refToWorld, onFace); ++np;
    }
    if (a1 < distToPlane) {
        V3 p = mk3(line1.y, a1, line1.x);
        V3 onFace = p; onFace.y = distToPlane;
        m.p0[np] = pos1 + mul(refToWorld, p); m.p1[np] = pos1 + mul(refToWorld, onFace); ++np;
    }
    if (!np) {
        np = 1;
        m.p0[0] = pos1 + mul(meshToWorld, tc.cpBody);
        m.p1[0] = pos1 + mul(meshToWorld, tc.cpTri);
    }
    m.np = np;
    return true;
}

__device__ __forceinline__ bool voided(const unsigned int* set, int n, unsigned int v) {
    for (int i = 0; i < n; ++i) if (set[i] == v) return true;
    return false;
}
__device__ __forceinline__ void voidInsert(unsigned int* set, int& n, unsigned int v) {
    if (!voided(set, n, v) && n < 3 * PB_MAX_TRI_CONTACTS) set[n++] = v;
}


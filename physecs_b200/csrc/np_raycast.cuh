// Ray vs collider geometry for Scene::raycastClosest (reference src/Raycast.cpp:5-151), same branch structure and fp32 order.
//   intersectRayAABB      :5-29   slab test; an axis with |dir| < 0.001 counts as parallel (origin must lie inside the slab)
//   intersectRaySphere    :31-46  ; intersectRayCapsule :48-104 (two hemispheres + cylinder in the capsule frame)
//   intersectRayBox       :106-111; intersectRayConvexMesh :113-139 (half-space clipping over the scaled faces)
//   triangle meshes have no ray routine in the reference (intersectRayGeometry :141-149 answers false)
#pragma once
#include "pb_math.cuh"
#include "pb_ctx.h"

__device__ inline bool rayAABB(V3 o, V3 d, V3 bmin, V3 bmax, float& t) {
    float tMin = 0.f, tMax = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float di = get(d, i), oi = get(o, i), lo = get(bmin, i), hi = get(bmax, i);
        if (fabsf(di) < 0.001f) {
            if (oi < lo || oi > hi) return false;
        } else {
            float ood = 1.f / di;
            float t1 = (lo - oi) * ood, t2 = (hi - oi) * ood;
            if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
            if (t1 > tMin) tMin = t1;
            if (t2 < tMax) tMax = t2;
            if (tMin > tMax) return false;
        }
    }
    t = tMin;
    return true;
}

__device__ inline bool raySphere(V3 o, V3 d, V3 pos, float radius, float& t) {
    V3 m = o - pos;
    float b = dot(m, d);
    float c = dot(m, m) - radius * radius;
    if (c > 0.f && b > 0.f) return false;
    float discr = b * b - c;
    if (discr < 0.f) return false;
    t = -b - sqrtf(discr);
    if (t < 0.f) t = 0.f;
    return true;
}

__device__ inline bool rayCapsule(V3 o, V3 dir, V3 pos, Q4 ori, float hh, float radius, float& t) {
    Q4 inv = qinverse(ori);
    V3 p = rotate(inv, o - pos);
    V3 d = rotate(inv, dir);
    float pd = dot(p, d);
    float hdy = hh * d.y;
    float thpy = 2 * hh * p.y;
    float phr = dot(p, p) + hh * hh - radius * radius;
    float tMin = FLT_MAX;
    bool hit = false;
    float b = pd - hdy, c = phr - thpy;
    float discr = b * b - c;
    if (discr >= 0.f) {                                   // top hemisphere
        t = -b - sqrtf(discr);
        if (t < 0.f) t = 0.f;
        if ((p + t * d).y >= hh) { hit = true; tMin = t; }
    }
    float a = d.x * d.x + d.z * d.z;                      // cylinder
    b = p.x * d.x + p.z * d.z;
    c = p.x * p.x + p.z * p.z - radius * radius;
    discr = b * b - a * c;
    if (discr >= 0.f) {
        t = (-b - sqrtf(discr)) / a;
        if (t < 0.f) t = 0.f;
        float y = (p + t * d).y;
        if (-hh <= y && y < hh && t < tMin) { hit = true; tMin = t; }
    }
    b = pd + hdy; c = phr + thpy;                         // bottom hemisphere
    discr = b * b - c;
    if (discr >= 0.f) {
        t = -b - sqrtf(discr);
        if (t < 0.f) t = 0.f;
        if ((p + t * d).y < -hh && t < tMin) { hit = true; tMin = t; }
    }
    t = tMin;
    return hit;
}

__device__ inline bool rayBox(V3 o, V3 dir, V3 pos, Q4 ori, V3 he, float& t) {
    Q4 inv = qinverse(ori);
    return rayAABB(rotate(inv, o - pos), rotate(inv, dir), -he, he, t);
}

__device__ inline bool rayConvex(V3 o, V3 dir, V3 pos, Q4 ori, const PbConvexDev& cm, V3 scale, float& t) {
    float tMin = 0.f, tMax = FLT_MAX;
    for (int f = 0; f < cm.nFaces; ++f) {
        V3 planeOrig = pos + rotate(ori, scale * mk3(cm.faceCentroid[f]));
        V3 planeNormal = rotate(ori, (mk3(1.f) / scale) * mk3(cm.faceNormal[f]));
        float denom = dot(dir, planeNormal);
        float dist = dot(planeOrig - o, planeNormal);
        if (denom == 0.f) {
            if (dist > 0) return false;
        } else {
            t = dist / denom;
            if (denom < 0) { if (t > tMin) tMin = t; }
            else { if (t < tMax) tMax = t; }
            if (tMin > tMax) return false;
        }
    }
    t = tMin;
    return true;
}

__device__ inline bool rayGeometry(V3 o, V3 d, int type, float4 prm, V3 pos, Q4 ori, const PbConvexDev* convexes, int mesh, float& t) {
    switch (type) {
        case PB_SPHERE: return raySphere(o, d, pos, prm.x, t);
        case PB_CAPSULE: return rayCapsule(o, d, pos, ori, prm.x, prm.y, t);
        case PB_BOX: return rayBox(o, d, pos, ori, mk3(prm.x, prm.y, prm.z), t);
        case PB_CONVEX_MESH: return rayConvex(o, d, pos, ori, convexes[mesh], mk3(prm.x, prm.y, prm.z), t);
        default: return false;
    }
}

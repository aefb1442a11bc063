// World-space AABB of one collider shape (shared by the broadphase bounds kernel and the mesh-local cull).
//   getBoundsSphere / Capsule / Box / ConvexMesh   reference src/BoundsUtil.cpp:35-69
#pragma once
#include "pb_math.cuh"
#include "pb_ctx.h"

struct Aabb { V3 mn, mx; };

// getBoundsSphere / Capsule / Box / ConvexMesh  (BoundsUtil.cpp:35-69)
__device__ inline Aabb shapeBounds(V3 pos, Q4 ori, int type, float4 prm, const PbConvexDev* convexes, int mesh) {
    Aabb b;
    switch (type) {
        case PB_SPHERE: {
            b.mn = pos - mk3(prm.x); b.mx = pos + mk3(prm.x);
        } break;
        case PB_CAPSULE: {
            V3 p0 = pos + rotate(ori, mk3(0.f, prm.x, 0.f));
            V3 p1 = pos + rotate(ori, mk3(0.f, -prm.x, 0.f));
            b.mn = mk3(gmin(p0.x, p1.x), gmin(p0.y, p1.y), gmin(p0.z, p1.z)) - mk3(prm.y);
            b.mx = mk3(gmax(p0.x, p1.x), gmax(p0.y, p1.y), gmax(p0.z, p1.z)) + mk3(prm.y);
        } break;
        case PB_BOX: {
            M3 u = mat3_cast(ori);
            M3 a; a.c[0] = vabs(u.c[0]); a.c[1] = vabs(u.c[1]); a.c[2] = vabs(u.c[2]);
            V3 w = mul(a, mk3(prm.x, prm.y, prm.z));
            b.mn = pos - w; b.mx = pos + w;
        } break;
        case PB_CONVEX_MESH: {
            const PbConvexDev& cm = convexes[mesh];
            V3 mn = mk3(FLT_MAX), mx = mk3(-FLT_MAX);
            V3 scale = mk3(prm.x, prm.y, prm.z);
            for (int i = 0; i < cm.nVertsPadded; ++i) {
                V3 p = pos + rotate(ori, scale * mk3(cm.verts[i]));
                mn = mk3(gmin(mn.x, p.x), gmin(mn.y, p.y), gmin(mn.z, p.z));
                mx = mk3(gmax(mx.x, p.x), gmax(mx.y, p.y), gmax(mx.z, p.z));
            }
            b.mn = mn; b.mx = mx;
        } break;
        default: b.mn = pos; b.mx = pos;
    }
    return b;
}


// Scene queries on the device tree: the callers next to the step (the reference's CharacterController uses them).
//   Scene::raycastClosest   reference src/Physecs.cpp:571-626 + src/Raycast.cpp
//   Scene::overlap          :628-650 + src/Overlap.cpp (np_overlap.cuh)
// The reference walks its incremental query BVH (BVH.cpp); inner-node boxes there and here are unions of the same leaf
// bounds (BroadPhaseEntry::bounds), so a leaf passes its ancestors' tests whenever it passes its own: both walks visit
// exactly the colliders whose own bounds the ray / query box hits.  The walk here is over the step's LBVH, rebuilt on demand
// when bounds or poses changed since it was built.  Built -fmad=false like the narrowphase.
#include "pb_ctx.h"
#include "pb_math.cuh"
#include "np_bounds.cuh"
#include "np_raycast.cuh"
#include "np_overlap.cuh"

__device__ __forceinline__ bool boxesIntersect(V3 amn, V3 amx, float4 bmn, float4 bmx) {
    return !(amx.x < bmn.x || amn.x > bmx.x) && !(amx.y < bmn.y || amn.y > bmx.y) && !(amx.z < bmn.z || amn.z > bmx.z);
}

// every collider the ray hits within maxDist: rows (ray, collider, t).  One thread per ray.
__global__ void k_query_raycast(int nRays, const float* __restrict__ orig3, const float* __restrict__ dir3, float maxDist, int nCol,
                                const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax,
                                const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                const float4* __restrict__ wpos, const float4* __restrict__ wquat, const PbConvexDev* __restrict__ convexes,
                                int* __restrict__ out, int cap, int* __restrict__ count) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRays) return;
    V3 o = mk3(orig3[3 * r], orig3[3 * r + 1], orig3[3 * r + 2]);
    V3 d = mk3(dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]);
    auto leaf = [&](int c) {
        float t;
        // the leaf's own bounds first (Physecs.cpp:574-575), then the geometry at its current world pose (:579-580)
        if (!rayAABB(o, d, mk3(aabbMin[c]), mk3(aabbMax[c]), t) || t > maxDist) return;
        if (!rayGeometry(o, d, colType[c], colParams[c], mk3(wpos[c]), mkq(wquat[c]), convexes, colMesh[c], t) || t > maxDist) return;
        int slot = atomicAdd(count, 1);
        if (slot < cap) { out[3 * slot] = r; out[3 * slot + 1] = c; out[3 * slot + 2] = __float_as_int(t); }
    };
    if (nCol == 1) { leaf(0); return; }
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        int node = stack[--sp];
        float4 lmn = nodeMin[2 * node], lmx = nodeMax[2 * node], rmn = nodeMin[2 * node + 1], rmx = nodeMax[2 * node + 1];
        int lc = __float_as_int(lmn.w), rc = __float_as_int(lmx.w);
        float t;
        if (rayAABB(o, d, mk3(lmn), mk3(lmx), t) && t <= maxDist) { if (lc >= 0) { if (sp < 64) stack[sp++] = lc; } else leaf(~lc); }
        if (rayAABB(o, d, mk3(rmn), mk3(rmx), t) && t <= maxDist) { if (rc >= 0) { if (sp < 64) stack[sp++] = rc; } else leaf(~rc); }
    }
}

// colliders overlapping a query shape: (collider) rows.  filter != 0 keeps colliders with (data & filter) != 0 (Physecs.cpp:634).
__global__ void k_query_overlap(int qType, float4 qPrm, float4 qPos, float4 qQuat, int qMesh, int filter, int nCol,
                                const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax,
                                const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                const int* __restrict__ colData, const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                const PbConvexDev* __restrict__ convexes, int* __restrict__ out, int cap, int* __restrict__ count) {
    if (blockIdx.x || threadIdx.x) return;
    V3 pos = mk3(qPos); Q4 ori = mkq(qQuat);
    Aabb qb = shapeBounds(pos, ori, qType, qPrm, convexes, qMesh);      // getBounds(pos, ori, geometry), no margin (:642)
    auto leaf = [&](int c) {
        if (!boxesIntersect(qb.mn, qb.mx, aabbMin[c], aabbMax[c])) return;
        if (filter && !(colData[c] & filter)) return;
        if (!overlapShapes(qType, qPrm, pos, ori, qMesh, colType[c], colParams[c], mk3(wpos[c]), mkq(wquat[c]), colMesh[c], convexes)) return;
        int slot = atomicAdd(count, 1);
        if (slot < cap) out[slot] = c;
    };
    if (nCol == 1) { leaf(0); return; }
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        int node = stack[--sp];
        float4 lmn = nodeMin[2 * node], lmx = nodeMax[2 * node], rmn = nodeMin[2 * node + 1], rmx = nodeMax[2 * node + 1];
        int lc = __float_as_int(lmn.w), rc = __float_as_int(lmx.w);
        if (boxesIntersect(qb.mn, qb.mx, lmn, lmx)) { if (lc >= 0) { if (sp < 64) stack[sp++] = lc; } else leaf(~lc); }
        if (boxesIntersect(qb.mn, qb.mx, rmn, rmx)) { if (rc >= 0) { if (sp < 64) stack[sp++] = rc; } else leaf(~rc); }
    }
}

// overlapWithMinTranslationalDistance, stage 1: the query shape becomes collider slot nCol; every collider whose bounds meet the
// query bounds (Physecs.cpp:654-655) is paired with it, collider first (collision(collider, query), :659).
__global__ void k_query_candidates(int qType, float4 qPrm, float4 qPos, float4 qQuat, int qMesh, int nCol,
                                   const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax,
                                   const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                   int* __restrict__ colType, float4* __restrict__ colParams, int* __restrict__ colMesh,
                                   float4* __restrict__ wpos, float4* __restrict__ wquat, const PbConvexDev* __restrict__ convexes,
                                   int2* __restrict__ pairs, int cap, int* __restrict__ counters) {
    if (blockIdx.x || threadIdx.x) return;
    colType[nCol] = qType; colParams[nCol] = qPrm; colMesh[nCol] = qMesh; wpos[nCol] = qPos; wquat[nCol] = qQuat;
    V3 pos = mk3(qPos); Q4 ori = mkq(qQuat);
    Aabb qb = shapeBounds(pos, ori, qType, qPrm, convexes, qMesh);      // getBounds(pos, ori, geometry), no margin (:681)
    int n = 0;
    auto leaf = [&](int c) {
        if (!boxesIntersect(qb.mn, qb.mx, aabbMin[c], aabbMax[c])) return;
        if (n < cap) pairs[n] = make_int2(c, nCol);
        ++n;
    };
    if (nCol == 1) leaf(0);
    else {
        int stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            int node = stack[--sp];
            float4 lmn = nodeMin[2 * node], lmx = nodeMax[2 * node], rmn = nodeMin[2 * node + 1], rmx = nodeMax[2 * node + 1];
            int lc = __float_as_int(lmn.w), rc = __float_as_int(lmx.w);
            if (boxesIntersect(qb.mn, qb.mx, lmn, lmx)) { if (lc >= 0) { if (sp < 64) stack[sp++] = lc; } else leaf(~lc); }
            if (boxesIntersect(qb.mn, qb.mx, rmn, rmx)) { if (rc >= 0) { if (sp < 64) stack[sp++] = rc; } else leaf(~rc); }
        }
    }
    counters[CNT_PAIRS] = n;          // may exceed cap: the host retries with a larger arena
}

// tree + world collider poses for the current device state (no-op while nothing changed since the last build)
static int prepareQueries(pb_ctx* ctx, int cap) {
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }      // a pending pb_set_state upload (copy stream) is part of the state
    if (!ctx->queryTreeValid) {
        int rc = pb_world_poses(ctx); if (rc) return rc;
        if (ctx->nCol >= 2) { rc = pb_build_tree(ctx); if (rc) return rc; }
        ctx->queryTreeValid = true;
    }
    if (cap > ctx->queryCap) {
        int rc = pb_alloc(ctx, &ctx->queryOut, (size_t)cap + 4); if (rc) return rc;
        ctx->queryCap = cap;
    }
    return PB_OK;
}

extern "C" {

int pb_query_raycast(pb_ctx* ctx, int nRays, const float* orig3, const float* dir3, float maxDist, int cap, int* outRay, int* outEntity,
                     int* outColIdx, float* outT, int* nHits) {
    cudaSetDevice(ctx->device);
    *nHits = 0;
    if (nRays <= 0 || ctx->nCol == 0) return PB_OK;
    int rc = prepareQueries(ctx, 3 * cap + 6 * nRays); if (rc) return rc;
    int* count = ctx->queryOut + 3 * cap;
    float* dRays = (float*)(count + 1);
    PB_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(dRays, orig3, sizeof(float) * 3 * nRays, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(dRays + 3 * nRays, dir3, sizeof(float) * 3 * nRays, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->launches, k_query_raycast<<<pb_grid(nRays, 64), 64, 0, ctx->stream>>>(nRays, dRays, dRays + 3 * nRays, maxDist, ctx->nCol, ctx->nodeMin, ctx->nodeMax,
        ctx->aabbMin, ctx->aabbMax, ctx->colType, ctx->colParams, ctx->colMesh, ctx->colWPos, ctx->colWQuat, ctx->convexDev, ctx->queryOut, cap, count);
    std::vector<int> h((size_t)3 * cap + 1);
    PB_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->queryOut, sizeof(int) * (3 * (size_t)cap + 1), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int n = h[3 * (size_t)cap];
    *nHits = n;                       // may exceed cap: the caller retries with a larger buffer
    for (int i = 0; i < n && i < cap; ++i) {
        int c = h[3 * i + 1];
        outRay[i] = h[3 * i];
        outEntity[i] = ctx->hRowEntity[ctx->hColRow[c]];
        outColIdx[i] = ctx->hColIndex[c];
        memcpy(&outT[i], &h[3 * i + 2], sizeof(float));
    }
    return PB_OK;
}

int pb_query_overlap(pb_ctx* ctx, const float* pos3, const float* quat4, int type, const float* params4, int mesh, int filter, int cap,
                     int* outEntity, int* outColIdx, int* nHits) {
    cudaSetDevice(ctx->device);
    *nHits = 0;
    if (ctx->nCol == 0) return PB_OK;
    if (type == PB_CONVEX_MESH && (mesh < 0 || mesh >= (int)ctx->convexes.size())) return pb_fail(ctx, PB_EINVAL, "pb_query_overlap: bad convex handle");
    if (type == PB_TRIANGLE_MESH) return PB_OK;      // physecs::overlap has no triangle-mesh case (Overlap.cpp:203-233): nothing overlaps
    int rc = prepareQueries(ctx, cap + 1); if (rc) return rc;
    int* count = ctx->queryOut + cap;
    PB_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
    ++ctx->launches, k_query_overlap<<<1, 32, 0, ctx->stream>>>(type, make_float4(params4[0], params4[1], params4[2], params4[3]),
        make_float4(pos3[0], pos3[1], pos3[2], 0.f), make_float4(quat4[0], quat4[1], quat4[2], quat4[3]), mesh, filter, ctx->nCol, ctx->nodeMin, ctx->nodeMax,
        ctx->aabbMin, ctx->aabbMax, ctx->colType, ctx->colParams, ctx->colMesh, ctx->colData, ctx->colWPos, ctx->colWQuat, ctx->convexDev, ctx->queryOut, cap, count);
    std::vector<int> h((size_t)cap + 1);
    PB_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->queryOut, sizeof(int) * ((size_t)cap + 1), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int n = h[cap];
    *nHits = n;
    for (int i = 0; i < n && i < cap; ++i) { int c = h[i]; outEntity[i] = ctx->hRowEntity[ctx->hColRow[c]]; outColIdx[i] = ctx->hColIndex[c]; }
    return PB_OK;
}

int pb_get_tree(pb_ctx* ctx, int cap, float* childBoxes12, int* childLinks2, int* nInternal) {
    cudaSetDevice(ctx->device);
    *nInternal = 0;
    if (ctx->nCol < 2) return PB_OK;
    int rc = prepareQueries(ctx, 1); if (rc) return rc;
    int n = ctx->nCol - 1;
    *nInternal = n;
    if (cap < n) return PB_OK;
    std::vector<float4> mn((size_t)2 * n), mx((size_t)2 * n);
    PB_CUDA(ctx, cudaMemcpyAsync(mn.data(), ctx->nodeMin, sizeof(float4) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(mx.data(), ctx->nodeMax, sizeof(float4) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        for (int side = 0; side < 2; ++side) {
            const float4 &a = mn[2 * (size_t)i + side], &b = mx[2 * (size_t)i + side];
            float* o = childBoxes12 + 12 * (size_t)i + 6 * side;
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z;
        }
        // links live in the .w words of the LEFT child's min / max (broadphase.cu): >= 0 internal node, < 0 ~collider
        int lc, rcx;
        memcpy(&lc, &mn[2 * (size_t)i].w, 4); memcpy(&rcx, &mx[2 * (size_t)i].w, 4);
        for (int side = 0; side < 2; ++side) {
            int link = side ? rcx : lc;
            if (link >= 0) childLinks2[2 * i + side] = link;
            else { int c = ~link; childLinks2[2 * i + side] = -1 - c; }
        }
    }
    return PB_OK;
}

int pb_collider_ids(pb_ctx* ctx, int n, const int* cols, int* outEntity, int* outColIdx) {
    for (int i = 0; i < n; ++i) {
        int c = cols[i];
        if (c < 0 || c >= ctx->nCol) return pb_fail(ctx, PB_EINVAL, "pb_collider_ids: collider out of range");
        outEntity[i] = ctx->hRowEntity[ctx->hColRow[c]];
        outColIdx[i] = ctx->hColIndex[c];
    }
    return PB_OK;
}

int pb_query_overlap_mtd(pb_ctx* ctx, const float* pos3, const float* quat4, int type, const float* params4, int mesh, int cap,
                         int* outEntity, int* outColIdx, float* outNormal3, float* outMtd, int* nHits) {
    cudaSetDevice(ctx->device);
    *nHits = 0;
    if (ctx->nCol == 0) return PB_OK;
    if (type == PB_CONVEX_MESH && (mesh < 0 || mesh >= (int)ctx->convexes.size())) return pb_fail(ctx, PB_EINVAL, "pb_query_overlap_mtd: bad convex handle");
    if (type == PB_TRIANGLE_MESH) return pb_fail(ctx, PB_EINVAL, "pb_query_overlap_mtd: a triangle mesh cannot be the query shape");
    int rc = prepareQueries(ctx, 1); if (rc) return rc;
    const int outCap = cap;
    if (cap < 64) cap = 64;
    if (cap > ctx->qArenaCap) {
        // candidates and manifolds share one capacity: a mesh collider yields up to PB_MAX_TRI_CONTACTS manifolds, the rest at most one
        if ((rc = pb_alloc(ctx, &ctx->qCounters, CNT_TOTAL)) || (rc = pb_alloc(ctx, &ctx->qPairs, (size_t)cap)) || (rc = pb_alloc(ctx, &ctx->qPairOrder, 2 * (size_t)cap)) ||
            (rc = pb_alloc(ctx, &ctx->qmKey, (size_t)cap)) || (rc = pb_alloc(ctx, &ctx->qmNormal, (size_t)cap)) || (rc = pb_alloc(ctx, &ctx->qmPts, 8 * (size_t)cap))) return rc;
        ctx->qArenaCap = cap;
    }
    cap = ctx->qArenaCap;
    PB_CUDA(ctx, cudaMemsetAsync(ctx->qCounters, 0, sizeof(int) * CNT_TOTAL, ctx->stream));
    ++ctx->launches, k_query_candidates<<<1, 32, 0, ctx->stream>>>(type, make_float4(params4[0], params4[1], params4[2], params4[3]),
        make_float4(pos3[0], pos3[1], pos3[2], 0.f), make_float4(quat4[0], quat4[1], quat4[2], quat4[3]), mesh, ctx->nCol, ctx->nodeMin, ctx->nodeMax,
        ctx->aabbMin, ctx->aabbMax, ctx->colType, ctx->colParams, ctx->colMesh, ctx->colWPos, ctx->colWQuat, ctx->convexDev, ctx->qPairs, cap, ctx->qCounters);
    rc = pb_narrowphase_query(ctx, ctx->qCounters, ctx->qPairs, ctx->qPairOrder, cap, ctx->qmKey, ctx->qmNormal, ctx->qmPts); if (rc) return rc;
    int hc[CNT_TOTAL];
    PB_CUDA(ctx, cudaMemcpyAsync(hc, ctx->qCounters, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int nPairs = hc[CNT_PAIRS], nM = hc[CNT_RAWM];
    if (nPairs > cap || nM > cap || nM > outCap) { *nHits = (nPairs > nM ? nPairs : nM) + PB_MAX_TRI_CONTACTS; return PB_OK; }   // caller retries with that capacity
    // (a query shape that meets more triangles / needs a larger polytope than the per-thread containers hold went through the spill kernels)
    if (hc[CNT_STATUS] & PB_ECAPACITY) { *nHits = 2 * cap + PB_MAX_TRI_CONTACTS; return PB_OK; }      // a manifold slot past the arena: retry larger
    if (nM == 0) return PB_OK;
    std::vector<int4> key((size_t)nM); std::vector<float4> nrm((size_t)nM), pts((size_t)8 * nM);
    PB_CUDA(ctx, cudaMemcpyAsync(key.data(), ctx->qmKey, sizeof(int4) * nM, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(nrm.data(), ctx->qmNormal, sizeof(float4) * nM, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(pts.data(), ctx->qmPts, sizeof(float4) * 8 * nM, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int w = 0;
    for (int i = 0; i < nM; ++i) {
        if (key[i].w == 0) continue;      // manifolds without points are dropped (Physecs.cpp:661)
        int c = key[i].x;
        // mtd = the deepest point's dot(position0 - position1, normal), never below 0 (Physecs.cpp:662-669); glm::dot sums x, y, z in order
        float mtd = 0.f;
        for (int k = 0; k < key[i].w; ++k) {
            const float4 &p0 = pts[8 * (size_t)i + 2 * k], &p1 = pts[8 * (size_t)i + 2 * k + 1];
            volatile float tx = (p0.x - p1.x) * nrm[i].x, ty = (p0.y - p1.y) * nrm[i].y, tz = (p0.z - p1.z) * nrm[i].z;
            volatile float s = tx + ty;
            float d = s + tz;
            if (d > mtd) mtd = d;
        }
        if (w < outCap) {
            outEntity[w] = ctx->hRowEntity[ctx->hColRow[c]];
            outColIdx[w] = ctx->hColIndex[c];
            outNormal3[3 * w] = nrm[i].x; outNormal3[3 * w + 1] = nrm[i].y; outNormal3[3 * w + 2] = nrm[i].z;
            outMtd[w] = mtd;
        }
        ++w;
    }
    *nHits = w;
    return PB_OK;
}

} // extern "C"

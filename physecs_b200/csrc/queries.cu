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

// tree + world collider poses for the current device state (no-op while nothing changed since the last build)
static int prepareQueries(pb_ctx* ctx, int cap) {
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

} // extern "C"

// Narrowphase: candidate pairs -> contact manifolds.
//
// Stage layout (one grid-stride launch each, counts stay on the device):
//   k_pair_classify   filter (trigger / nonCollidingPairs) + shape-pair bin histogram   Physecs.cpp:200-209
//   k_bin_starts      exclusive scan of the bin histogram
//   k_pair_scatter    pair indices grouped by bin
//   k_np_prim<BIN>    thread-per-pair analytic routines                                  Collision.cpp:895-1016 dispatch
//   k_np_mesh         thread-per-(shape, mesh) pair: BVH cull + filter + manifolds        CollisionTriangleMesh.cpp:882-956
// Each warp appends its manifolds with ONE atomic (ballot + prefix), so the manifold arena stays dense.
// Built with -fmad=false: feature / face decisions then follow the reference's fp32 arithmetic.
#include "pb_ctx.h"
#include "pb_math.cuh"
#include "np_prims.cuh"
#include "np_mesh.cuh"
#include "np_gjk.cuh"
#include "np_mesh_heavy.cuh"
#include "np_overlap.cuh"
#include <mutex>
#include <unordered_map>
#include <algorithm>

#define COLF_TRIGGER 1
#define COLF_ENABLE 2
#define COLF_DYNAMIC 4

// mesh bins are split by the dynamic shape's type so every warp of a mesh launch runs one triangle routine
enum { BIN_SS = 0, BIN_SC, BIN_CC, BIN_SB, BIN_CB, BIN_BB, BIN_GJK, BIN_MESH_S, BIN_MESH_C, BIN_MESH_B, BIN_MESH_X, BIN_TRIGGER, BIN_COUNT };
#define PB_MAX_TRI_CAND 128   // BVH leaf triangles visited per (shape, mesh) pair before the per-triangle tests

__device__ __forceinline__ int binOf(int t0, int t1) {
    if (t0 == PB_TRIANGLE_MESH || t1 == PB_TRIANGLE_MESH) {
        if (t0 == t1) return -1;
        int other = t0 == PB_TRIANGLE_MESH ? t1 : t0;
        return BIN_MESH_S + other;   // sphere, capsule, box, convex (box / convex need GJK-sized scratch)
    }
    if (t0 == PB_CONVEX_MESH || t1 == PB_CONVEX_MESH) return BIN_GJK;
    int lo = min(t0, t1), hi = max(t0, t1);
    if (lo == PB_SPHERE) return hi == PB_SPHERE ? BIN_SS : (hi == PB_CAPSULE ? BIN_SC : BIN_SB);
    if (lo == PB_CAPSULE) return hi == PB_CAPSULE ? BIN_CC : BIN_CB;
    return BIN_BB;
}

__device__ __forceinline__ bool isNonColliding(const unsigned long long* __restrict__ keys, int n, unsigned long long k) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        unsigned long long v = keys[mid];
        if (v == k) return true;
        if (v < k) lo = mid + 1; else hi = mid - 1;
    }
    return false;
}

__global__ void k_pair_classify(const int2* __restrict__ pairs, int* __restrict__ pairBin, int* __restrict__ counters, int maxPairs,
                                const int* __restrict__ colType, const int* __restrict__ colFlags, const int* __restrict__ colRow,
                                const int* __restrict__ rowEntity, const unsigned long long* __restrict__ nonColl, int nNonColl,
                                const int* __restrict__ colClass, const unsigned char* __restrict__ filterLut, int nClasses) {
    __shared__ int sh[BIN_COUNT];
    if (threadIdx.x < BIN_COUNT) sh[threadIdx.x] = 0;
    __syncthreads();
    int n = min(counters[CNT_PAIRS], maxPairs);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int2 p = pairs[i];
        int bin;
        // contactFilter(col0.isTrigger, col0.data, col1.isTrigger, col1.data) (Physecs.cpp:200): the default filter
        // (Physecs.cpp:20-23) reads the trigger flags; a custom one is tabulated per (isTrigger, data) class on the host
        bool trigger;
        if (filterLut) trigger = filterLut[colClass[p.x] * nClasses + colClass[p.y]] != 0;
        else trigger = ((colFlags[p.x] | colFlags[p.y]) & COLF_TRIGGER) != 0;
        if (trigger) bin = BIN_TRIGGER;
        else {
            bin = binOf(colType[p.x], colType[p.y]);
            if (bin >= 0 && nNonColl) {
                unsigned long long k = ((unsigned long long)(unsigned int)rowEntity[colRow[p.x]] << 32) | (unsigned int)rowEntity[colRow[p.y]];
                if (isNonColliding(nonColl, nNonColl, k)) bin = -1;
            }
        }
        if (bin >= 0) atomicAdd(&sh[bin], 1);
        pairBin[i] = bin;
    }
    __syncthreads();
    if (threadIdx.x < BIN_COUNT && sh[threadIdx.x]) atomicAdd(&counters[CNT_BIN0 + threadIdx.x], sh[threadIdx.x]);
}

__global__ void k_bin_starts(int* counters) {
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < BIN_COUNT; ++b) { counters[CNT_BINSTART + b] = run; run += counters[CNT_BIN0 + b]; counters[CNT_BIN0 + b] = 0; }
        counters[CNT_BINSTART + BIN_COUNT] = run;
        counters[CNT_MESH_PAIRS] = counters[CNT_BINSTART + BIN_TRIGGER] - counters[CNT_BINSTART + BIN_MESH_S];
    }
}

// Pairs grouped by bin.  One CTA ranks a tile of blockDim pairs: warps rank their lanes per bin (match_any), the CTA adds the warp
// counts up in shared memory and reserves ONE range per bin and tile (the per-warp atomics on three or four hot counters were
// most of this kernel's time at 2 M pairs).
#define SCATTER_THREADS 1024
__global__ void __launch_bounds__(SCATTER_THREADS) k_pair_scatter(const int* __restrict__ pairBin, int* __restrict__ pairOrder, int* __restrict__ counters, int maxPairs) {
    __shared__ int warpCount[SCATTER_THREADS / 32][BIN_COUNT];
    __shared__ int binStart[BIN_COUNT];
    const int n = min(counters[CNT_PAIRS], maxPairs);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < BIN_COUNT) binStart[threadIdx.x] = counters[CNT_BINSTART + threadIdx.x];
    for (int tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x) {
        for (int b = lane; b < BIN_COUNT; b += 32) warpCount[w][b] = 0;
        __syncthreads();
        int i = tile + threadIdx.x;
        int bin = i < n ? pairBin[i] : -1;
        unsigned int act = __ballot_sync(0xffffffffu, bin >= 0);
        int rank = 0;
        if (bin >= 0) {
            unsigned int peers = __match_any_sync(act, bin);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) warpCount[w][bin] = __popc(peers);
        }
        __syncthreads();
        if (threadIdx.x < BIN_COUNT) {
            // exclusive prefix over the warps of this bin, then one reservation for the whole tile
            int run = 0;
            for (int k = 0; k < SCATTER_THREADS / 32; ++k) { int c = warpCount[k][threadIdx.x]; warpCount[k][threadIdx.x] = run; run += c; }
            int base = run ? atomicAdd(&counters[CNT_BIN0 + threadIdx.x], run) : 0;
            for (int k = 0; k < SCATTER_THREADS / 32; ++k) warpCount[k][threadIdx.x] += base;
        }
        __syncthreads();
        if (bin >= 0) pairOrder[binStart[bin] + warpCount[w][bin] + rank] = i;
        __syncthreads();
    }
}

// ---- manifold arena append ------------------------------------------------------------------------------------
__device__ __forceinline__ int warpReserve(int want, int* counter) {
    // every lane of the warp must call; returns this lane's base slot for `want` entries
    int lane = threadIdx.x & 31;
    int inc = want;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    int total = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - want;
}

// A pair outgrew the per-thread containers of its bin kernel (`cause` = PB_CAUSE_SPILLED_* bits): no result is stored, the pair goes
// on the bin family's spill list and k_np_gjk_spill / k_np_mesh_spill redo it on global-memory scratch.
__device__ __forceinline__ void spillAppend(int* counters, int which, int* __restrict__ list, int pairIndex, int cause) {
    atomicOr(&counters[CNT_CAUSE], cause);
    int s = atomicAdd(&counters[which], 1);
    if (s < PB_SPILL_CAP) list[s] = pairIndex;
    else atomicOr(&counters[CNT_CAUSE], PB_CAUSE_SPILL_LIST);
}

__device__ __forceinline__ void storeManifold(int slot, int maxManifolds, int colA, int colB, const Manifold& m, bool flip,
                                              int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int* counters) {
    if (slot >= maxManifolds) { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_MANIFOLDS); return; }
    mKey[slot] = make_int4(colA, colB, m.tri, m.np);
    V3 n = flip ? -m.n : m.n;
    mNormal[slot] = f4(n);
    for (int k = 0; k < m.np; ++k) {
        mPts[8 * (size_t)slot + 2 * k] = f4(flip ? m.p1[k] : m.p0[k]);
        mPts[8 * (size_t)slot + 2 * k + 1] = f4(flip ? m.p0[k] : m.p1[k]);
    }
}

// ---- analytic bins ------------------------------------------------------------------------------------------------
template <int BIN>
__device__ __forceinline__ void npPrimBody(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                           const int* __restrict__ colType, const float4* __restrict__ colParams,
                                           const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                           const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                           int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                           int* __restrict__ spillList, const int bx, const int gx) {
    int start = counters[CNT_BINSTART + BIN], end = counters[CNT_BINSTART + BIN + 1];
    int lane = threadIdx.x & 31;
    for (int base = start + ((bx * blockDim.x + threadIdx.x) & ~31); base < end; base += gx * blockDim.x) {
        int idx = base + lane;
        bool hit = false, flip = false;
        Manifold m; m.np = 0; m.tri = -1;
        int a = 0, b = 0;
        if (idx < end) {
            const int pi = pairOrder[idx];
            int2 p = pairs[pi];
            a = p.x; b = p.y;
            int t0 = colType[a], t1 = colType[b];
            float4 q0 = colParams[a], q1 = colParams[b];
            V3 pos0 = mk3(wpos[a]), pos1 = mk3(wpos[b]);
            Q4 or0 = mkq(wquat[a]), or1 = mkq(wquat[b]);
            if (BIN == BIN_SS) hit = collideSphereSphere(pos0, q0.x, pos1, q1.x, m);
            else if (BIN == BIN_CC) hit = collideCapsuleCapsule(pos0, or0, q0.x, q0.y, pos1, or1, q1.x, q1.y, m);
            else if (BIN == BIN_BB) hit = collideBoxBox(pos0, or0, mk3(q0.x, q0.y, q0.z), pos1, or1, mk3(q1.x, q1.y, q1.z), m);
            else if (BIN == BIN_SC) {
                if (t0 == PB_SPHERE) hit = collideSphereCapsule(pos0, q0.x, pos1, or1, q1.x, q1.y, m);
                else { hit = collideSphereCapsule(pos1, q1.x, pos0, or0, q0.x, q0.y, m); flip = true; }
            } else if (BIN == BIN_SB) {
                if (t0 == PB_SPHERE) hit = collideSphereBox(pos0, q0.x, pos1, or1, mk3(q1.x, q1.y, q1.z), m);
                else { hit = collideSphereBox(pos1, q1.x, pos0, or0, mk3(q0.x, q0.y, q0.z), m); flip = true; }
            } else if (BIN == BIN_CB) {
                if (t0 == PB_CAPSULE) hit = collideCapsuleBox(pos0, or0, q0.x, q0.y, pos1, or1, mk3(q1.x, q1.y, q1.z), m);
                else { hit = collideCapsuleBox(pos1, or1, q1.x, q1.y, pos0, or0, mk3(q0.x, q0.y, q0.z), m); flip = true; }
            } else if (BIN == BIN_GJK) {
                Epa poly; int ovf = 0;
                hit = collideGjkPair<LimFast>(t0, q0, pos0, or0, colMesh[a], t1, q1, pos1, or1, colMesh[b], convexes, m, flip, poly, &ovf);
                if (ovf) { hit = false; spillAppend(counters, CNT_SPILL_GJK, spillList, pi, ovf); }
            }
            hit = hit && m.np > 0;
        }
        int slot = warpReserve(hit ? 1 : 0, &counters[CNT_RAWM]);
        if (hit) storeManifold(slot, maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
    }
}

template <int BIN>
__global__ void __launch_bounds__(128) k_np_prim(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                 const int* __restrict__ colType, const float4* __restrict__ colParams,
                                                 const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                 const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                                 int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                                 int* __restrict__ spillList) {
    npPrimBody<BIN>(pairs, pairOrder, counters, colType, colParams, wpos, wquat, convexes, colMesh, mKey, mNormal, mPts, maxManifolds, spillList, blockIdx.x, gridDim.x);
}

// Small scenes: the six analytic bins in ONE launch (blockIdx.y = bin) -- they then overlap instead of queueing behind one another, and a
// step saves five launches; each bin of such a scene holds a few thousand pairs at most and cannot fill the device on its own.
__global__ void __launch_bounds__(128) k_np_prim_all(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                     const int* __restrict__ colType, const float4* __restrict__ colParams,
                                                     const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                     const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                                     int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                                     int* __restrict__ spillList) {
#define PB_BIN_CASE(B) case B: npPrimBody<B>(pairs, pairOrder, counters, colType, colParams, wpos, wquat, convexes, colMesh, mKey, mNormal, mPts, maxManifolds, spillList, blockIdx.x, gridDim.x); break;
    switch (blockIdx.y) { PB_BIN_CASE(BIN_SS) PB_BIN_CASE(BIN_SC) PB_BIN_CASE(BIN_CC) PB_BIN_CASE(BIN_SB) PB_BIN_CASE(BIN_CB) PB_BIN_CASE(BIN_BB) }
#undef PB_BIN_CASE
}

// ---- GJK bin in two launches (the step; scene queries use k_np_prim<BIN_GJK>) ------------------------------------------------------
// k_np_gjk_hits: GJK on every pair of the bin; the intersecting ones are appended, with their simplex, to a dense hit list.
// k_np_gjk_manifolds: EPA + contact patch on the hit list: every lane of every warp has a penetrating pair.
__global__ void __launch_bounds__(128) k_np_gjk_hits(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                     const int* __restrict__ colType, const float4* __restrict__ colParams,
                                                     const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                     const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                                     int* __restrict__ hitPair, float4* __restrict__ hitSimplex, int maxHits) {
    int start = counters[CNT_BINSTART + BIN_GJK], end = counters[CNT_BINSTART + BIN_GJK + 1];
    int lane = threadIdx.x & 31;
    for (int base = start + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); base < end; base += gridDim.x * blockDim.x) {
        int idx = base + lane;
        bool hit = false;
        GjkV simplex[4];
        int pi = 0;
        if (idx < end) {
            pi = pairOrder[idx];
            int2 p = pairs[pi];
            GjkPair g = gjkPairSetup(colType[p.x], colParams[p.x], mk3(wpos[p.x]), mkq(wquat[p.x]), colMesh[p.x],
                                     colType[p.y], colParams[p.y], mk3(wpos[p.y]), mkq(wquat[p.y]), colMesh[p.y]);
            hit = gjkPairIntersect(g, convexes, simplex);
        }
        int slot = warpReserve(hit ? 1 : 0, &counters[CNT_GJK_HITS]);
        if (hit) {
            if (slot < maxHits) {
                hitPair[slot] = pi;
                float4* o = hitSimplex + 9 * (size_t)slot;
                const float* f = (const float*)simplex;        // 4 x {pos, sp0, sp1} = 36 floats
#pragma unroll
                for (int k = 0; k < 9; ++k) o[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
            } else { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_MANIFOLDS); }    // hit list capacity == the manifold arena's
        }
    }
}

__global__ void __launch_bounds__(128) k_np_gjk_manifolds(const int2* __restrict__ pairs, int* __restrict__ counters,
                                                          const int* __restrict__ colType, const float4* __restrict__ colParams,
                                                          const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                          const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                                          const int* __restrict__ hitPair, const float4* __restrict__ hitSimplex, int maxHits,
                                                          int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                                          int* __restrict__ spillList) {
    int n = min(counters[CNT_GJK_HITS], maxHits);
    int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
        int h = base + lane;
        bool hit = false, flip = false;
        Manifold m; m.np = 0; m.tri = -1;
        int a = 0, b = 0;
        if (h < n) {
            int2 p = pairs[hitPair[h]];
            a = p.x; b = p.y;
            GjkPair g = gjkPairSetup(colType[a], colParams[a], mk3(wpos[a]), mkq(wquat[a]), colMesh[a],
                                     colType[b], colParams[b], mk3(wpos[b]), mkq(wquat[b]), colMesh[b]);
            flip = g.flip;
            GjkV simplex[4];
            float* f = (float*)simplex;
            const float4* in = hitSimplex + 9 * (size_t)h;
#pragma unroll
            for (int k = 0; k < 9; ++k) { float4 v = in[k]; f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w; }
            Epa poly; int ovf = 0;
            gjkPairManifold<LimFast>(g, convexes, simplex, m, poly, &ovf);
            hit = m.np > 0;
            if (ovf) { hit = false; spillAppend(counters, CNT_SPILL_GJK, spillList, hitPair[h], ovf); }
        }
        int slot = warpReserve(hit ? 1 : 0, &counters[CNT_RAWM]);
        if (hit) storeManifold(slot, maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
    }
}

// ---- mesh bins -------------------------------------------------------------------------------------------------------
// Collect triangle contacts of one (shape, mesh) pair in the reference's traversal order (TriangleMesh.cpp:166-192:
// explicit stack, left child popped first; CollisionTriangleMesh.cpp:895-907).
template <int TYPE>
__device__ inline int meshCollect(float4 prm, V3 localPos, Q4 localOr, const PbTriMeshDev& mesh,
                                  const PbConvexDev* convexes, int convexId, TriContact* contacts, int* ovf, Epa* scratch) {
    constexpr bool HEAVY = TYPE >= PB_BOX;
    Aabb lb = shapeBounds(localPos, localOr, TYPE, prm, convexes, convexId);
    Shape convexShape;
    if (TYPE == PB_CONVEX_MESH) convexShape = makeShape(TYPE, prm, localPos, localOr, convexes, convexId);
    // pass A: BVH cull only.  Leaf triangles are recorded in visiting order; keeping the tests out of this loop lets the
    // lanes of a warp walk their (different-depth) trees without dragging the long triangle routines through every branch.
    int cand[PB_MAX_TRI_CAND];
    int nc = 0;
    {
        int stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            int node = stack[--sp];
            float4 nmn = mesh.nodeMin[node], nmx = mesh.nodeMax[node];
            if (lb.mx.x < nmn.x || lb.mn.x > nmx.x) continue;      // physecs::intersects (BoundsUtil.cpp:87-92)
            if (lb.mx.y < nmn.y || lb.mn.y > nmx.y) continue;
            if (lb.mx.z < nmn.z || lb.mn.z > nmx.z) continue;
            int triCount = __float_as_int(nmn.w), index = __float_as_int(nmx.w);
            if (triCount) {
                for (int k = 0; k < triCount; ++k) {
                    if (nc < PB_MAX_TRI_CAND) cand[nc++] = index + k;
                    else *ovf |= PB_CAUSE_SPILLED_TRI_CAND;
                }
            } else {
                if (sp + 2 <= 64) { stack[sp++] = index + 1; stack[sp++] = index; }
                else *ovf |= PB_CAUSE_SPILLED_MESH_STACK;
            }
        }
    }
    // pass B: per-triangle tests, in the visiting order (== the reference's overlapBvh order, which fixes contact order)
    int cnt = 0;
    for (int i = 0; i < nc; ++i) {
        int tri = cand[i];
        int4 ti = mesh.tris[tri];
        V3 a = mk3(mesh.verts[ti.x]), b = mk3(mesh.verts[ti.y]), c = mk3(mesh.verts[ti.z]);
        V3 n = mk3(mesh.triNormal[tri]);
        TriContact tc;
        tc.boxFeature = 0; tc.boxAxis = 0; tc.fidx = 0; tc.feature = TF_FACE; tc.dist = 0.f;
        tc.normal = tc.cpBody = tc.cpTri = mk3(0.f);
        bool hit;
        if (TYPE == PB_SPHERE) hit = sphereTriangle(localPos, prm.x, a, b, c, n, tc);
        else if (TYPE == PB_CAPSULE) hit = capsuleTriangle(localPos, localOr, prm.x, prm.y, a, b, c, n, tc);
        else if (TYPE == PB_BOX) hit = boxTriangle(localPos, localOr, mk3(prm.x, prm.y, prm.z), a, b, c, n, tc);
        else hit = convexTriangle(convexShape, a, b, c, mk3(mesh.triCentroid[tri]), tc, *scratch, ovf);
        if (hit) {
            tc.tri = tri;
            if (cnt < PB_MAX_TRI_CONTACTS) contacts[cnt++] = tc;
            else *ovf |= PB_CAUSE_SPILLED_TRI_CONTACTS;
        }
    }
    (void)HEAVY;
    return cnt;
}

// std::sort(contacts.begin(), contacts.end()) of the delayed edge / vertex contacts (CTM.cpp:932; operator< compares distance,
// :32-34), on an index array.  The oracle is the reference built with libstdc++, whose std::sort is: introsort partitioning
// (median of three to the front, unguarded Hoare partition) while a range holds more than 16 elements, then one insertion sort
// over everything -- restated here move for move so that contacts at EQUAL distance end up in the same order (up to 16 elements
// it is a plain insertion sort).  `less(i, j)`: contact i sorts strictly before contact j.
template <class IdxT, class Less>
__device__ inline void sortLikeStdSort(IdxT* v, int n, Less less) {
    if (n > 16) {
        int depth = 2 * (31 - __clz(n));
        int stackLo[64], stackHi[64], stackD[64];      // pending right-hand ranges (the recursion of __introsort_loop)
        int sp = 0;
        int first = 0, last = n;
        for (;;) {
            while (last - first > 16 && depth > 0) {
                --depth;
                int mid = first + (last - first) / 2;
                // __move_median_to_first(first, first + 1, mid, last - 1)
                int a = first + 1, b = mid, c = last - 1, r;
                if (less(v[a], v[b])) r = less(v[b], v[c]) ? b : (less(v[a], v[c]) ? c : a);
                else r = less(v[a], v[c]) ? a : (less(v[b], v[c]) ? c : b);
                { IdxT t = v[first]; v[first] = v[r]; v[r] = t; }
                // __unguarded_partition(first + 1, last, pivot = first)
                int i = first + 1, j = last;
                for (;;) {
                    while (less(v[i], v[first])) ++i;
                    --j;
                    while (less(v[first], v[j])) --j;
                    if (!(i < j)) break;
                    IdxT t = v[i]; v[i] = v[j]; v[j] = t;
                    ++i;
                }
                if (sp < 64) { stackLo[sp] = i; stackHi[sp] = last; stackD[sp] = depth; ++sp; }     // __introsort_loop(cut, last, depth_limit)
                last = i;
            }
            // (depth exhausted: libstdc++ heap-sorts the range; the final insertion sort below orders it as well, equal keys may differ)
            if (!sp) break;
            --sp; first = stackLo[sp]; last = stackHi[sp]; depth = stackD[sp];
        }
    }
    for (int i = 1; i < n; ++i) {          // __final_insertion_sort
        IdxT x = v[i];
        int j = i;
        while (j > 0 && less(x, v[j - 1])) { v[j] = v[j - 1]; --j; }
        v[j] = x;
    }
}

// The reference's two-pass feature filter over the triangle contacts of one (shape, mesh) pair (CTM.cpp:913-953): writes the
// generation order (indices into contacts) and returns how many contacts survive.
template <class Contact>
__device__ inline int meshFilter(const PbTriMeshDev& mesh, const Contact* contacts, int cnt, unsigned char* order) {
    int nGen = 0;
    // pass 1 (CTM.cpp:913-929): face contacts, reverse order, swap-remove; they void their vertices
    unsigned int voidSet[3 * PB_MAX_TRI_CONTACTS];
    int nVoid = 0;
    unsigned char live[PB_MAX_TRI_CONTACTS];
    int nLive = cnt;
    for (int i = 0; i < cnt; ++i) live[i] = (unsigned char)i;
    for (int i = nLive - 1; i >= 0; --i) {
        int ci = live[i];
        if (contacts[ci].feature == TF_FACE) {
            int4 ti = mesh.tris[contacts[ci].tri];
            voidInsert(voidSet, nVoid, (unsigned)ti.x); voidInsert(voidSet, nVoid, (unsigned)ti.y); voidInsert(voidSet, nVoid, (unsigned)ti.z);
            order[nGen++] = (unsigned char)ci;
            live[i] = live[nLive - 1];
            --nLive;
        }
    }
    sortLikeStdSort(live, nLive, [&](unsigned char x, unsigned char y) { return contacts[x].dist < contacts[y].dist; });
    // pass 2 (CTM.cpp:934-953)
    for (int i = 0; i < nLive; ++i) {
        int ci = live[i];
        const Contact& tc = contacts[ci];
        int4 ti = mesh.tris[tc.tri];
        unsigned int vi[3] = { (unsigned)ti.x, (unsigned)ti.y, (unsigned)ti.z };
        if (tc.feature == TF_EDGE) {
            if (voided(voidSet, nVoid, vi[(tc.fidx + 1) % 3]) && voided(voidSet, nVoid, vi[(tc.fidx + 2) % 3])) continue;
        } else {
            if (voided(voidSet, nVoid, vi[tc.fidx])) continue;
        }
        order[nGen++] = (unsigned char)ci;
        voidInsert(voidSet, nVoid, vi[0]); voidInsert(voidSet, nVoid, vi[1]); voidInsert(voidSet, nVoid, vi[2]);
    }
    return nGen;
}

template <int TYPE>
__global__ void __launch_bounds__(128) k_np_mesh(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                 const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                                 const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                 const PbTriMeshDev* __restrict__ meshes, const PbConvexDev* __restrict__ convexes,
                                                 int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                                 int* __restrict__ spillList) {
    constexpr bool HEAVY = TYPE >= PB_BOX;
    const int BIN = BIN_MESH_S + TYPE;
    int start = counters[CNT_BINSTART + BIN], end = counters[CNT_BINSTART + BIN + 1];
    int lane = threadIdx.x & 31;
    for (int base = start + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); base < end; base += gridDim.x * blockDim.x) {
        int idx = base + lane;
        TriContact contacts[PB_MAX_TRI_CONTACTS];
        unsigned char order[PB_MAX_TRI_CONTACTS];   // generation order: >=0 index into contacts
        int nGen = 0;
        int a = 0, b = 0, shape = 0;
        const int type = TYPE;
        bool flip = false;
        V3 pos1 = mk3(0.f), localPos = mk3(0.f); Q4 or1 = mkq(make_float4(0, 0, 0, 1)), localOr = or1;
        float4 prm = make_float4(0, 0, 0, 0);
        int meshId = 0;
        int ovf = 0, pi = 0;
        if (idx < end) {
            pi = pairOrder[idx];
            int2 p = pairs[pi];
            a = p.x; b = p.y;
            // Collision.cpp:897-907: mesh on side 0 -> collide (shape1, mesh0) and flip
            flip = colType[a] == PB_TRIANGLE_MESH;
            shape = flip ? b : a;
            int meshCol = flip ? a : b;
            prm = colParams[shape];
            meshId = colMesh[meshCol];
            V3 pos0 = mk3(wpos[shape]); Q4 or0 = mkq(wquat[shape]);
            pos1 = mk3(wpos[meshCol]); or1 = mkq(wquat[meshCol]);
            Q4 invOr1 = qinverse(or1);
            localPos = rotate(invOr1, pos0 - pos1);
            localOr = qmul(invOr1, or0);
            {
                Epa scratch;
                int cnt = meshCollect<TYPE>(prm, localPos, localOr, meshes[meshId], convexes, colMesh[shape], contacts, &ovf, HEAVY ? &scratch : nullptr);
                if (!ovf) nGen = meshFilter(meshes[meshId], contacts, cnt, order);
            }
        }
        // the convex generators clip a face against a triangle (or the reverse): at most (face vertices + 3) points.  Decided before the
        // arena slots are reserved, so a spilled pair stores nothing here.
        if (TYPE == PB_CONVEX_MESH && nGen && convexes[colMesh[shape]].maxFaceVerts + 3 > LimFast::POLY) ovf |= PB_CAUSE_SPILLED_CLIP;
        if (ovf) { nGen = 0; spillAppend(counters, CNT_SPILL_MESH, spillList, pi, ovf); }
        int slot = warpReserve(nGen, &counters[CNT_RAWM]);
        for (int g = 0; g < nGen; ++g) {
            const TriContact& tc = contacts[order[g]];
            Manifold m; m.np = 0; m.tri = tc.tri;
            int ovfGen = 0;
            if (!HEAVY) {
                if (type == PB_CAPSULE && tc.feature == TF_FACE) {
                    if (!capsuleTriangleFaceManifold(localPos, localOr, prm.x, prm.y, pos1, or1, meshes[meshId], tc, m)) { m.np = 0; m.n = mk3(0.f, 1.f, 0.f); }
                } else {
                    manifoldFromClosest(pos1, or1, tc, m);
                }
            } else if (type == PB_BOX) {
                // CTM.cpp:843-880: FACE -> box vs triangle face; EDGE -> by box feature; VERTEX -> box face vs triangle
                V3 he = mk3(prm.x, prm.y, prm.z);
                if (tc.feature == TF_FACE) boxTriangleFaceManifold(localPos, localOr, he, pos1, or1, meshes[meshId], tc, m);
                else if (tc.feature == TF_EDGE && tc.boxFeature != BOXF_FACE) boxEdgeTriangleEdgeManifold(localPos, localOr, he, pos1, or1, meshes[meshId], tc, m);
                else boxFaceTriangleManifold(localPos, localOr, he, pos1, or1, meshes[meshId], tc, m);
            } else {
                const PbConvexDev& cm = convexes[colMesh[shape]];
                V3 sc = mk3(prm.x, prm.y, prm.z);
                if (tc.feature == TF_FACE) convexTriangleFaceManifold<LimFast>(localPos, localOr, cm, sc, pos1, or1, meshes[meshId], tc, m, &ovfGen);
                else convexFaceTriangleManifold<LimFast>(localPos, localOr, cm, sc, pos1, or1, meshes[meshId], tc, m, &ovfGen);
                if (ovfGen) atomicOr(&counters[CNT_CAUSE], ovfGen | PB_CAUSE_SPILL_SCRATCH);      // unreachable: bounded by the maxFaceVerts test above
            }
            m.tri = tc.tri;
            // a == side 0 of the pair; manifolds are computed shape->mesh, flip when the mesh is side 0
            storeManifold(slot + g, maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
        }
    }
}

// ---- sphere / capsule vs mesh: the 1 M-body terrain scene's bins -------------------------------------------------------------
// Same result as k_np_mesh<PB_SPHERE|PB_CAPSULE> (same node / triangle predicates, same candidate and contact order), laid out for
// latency: (1) the cull walk tests BOTH children of a node per step -- siblings are adjacent in nodeMin / nodeMax, so one round
// of four independent loads replaces two dependent visits and boxes that miss are never popped; a right child whose left sibling
// still has a subtree to walk waits on the stack (leaf: as ~node), which keeps the reference's left-first order
// (TriangleMesh.cpp:166-192); (2) a triangle is one 64-byte record (vertices + normal) instead of index -> three vertices +
// normal, and the next candidate's record is in flight while the current one is tested.
// (Measured and dropped: laying the (pair, candidate) items of a warp's 32 pairs end to end and testing 32 at a time through shared
// memory -- bit-identical, but 1.46 vs 1.07 ms of narrowphase at 1 M bodies; and a 4-wide collapse of the tree -- half the dependent
// rounds, twice the L1 wavefronts per round: 0.89 vs 0.90 ms.)
#define ML_WARPS 4

__device__ __forceinline__ bool aabbHits(const Aabb& lb, float4 mn, float4 mx) {      // physecs::intersects (BoundsUtil.cpp:87-92)
    if (lb.mx.x < mn.x || lb.mn.x > mx.x) return false;
    if (lb.mx.y < mn.y || lb.mn.y > mx.y) return false;
    if (lb.mx.z < mn.z || lb.mn.z > mx.z) return false;
    return true;
}

__device__ inline int meshCullDual(const Aabb& lb, const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax, int* cand, int* ovf) {
    int nc = 0;
#define ML_ADD_LEAF(cnt_, idx_) do { for (int k_ = 0; k_ < (cnt_); ++k_) { if (nc < PB_MAX_TRI_CAND) cand[nc++] = (idx_) + k_; else *ovf |= PB_CAUSE_SPILLED_TRI_CAND; } } while (0)
#define ML_PUSH(v_) do { if (sp < 64) stack[sp++] = (v_); else *ovf |= PB_CAUSE_SPILLED_MESH_STACK; } while (0)
    float4 mn = nodeMin[0], mx = nodeMax[0];
    if (!aabbHits(lb, mn, mx)) return 0;
    if (__float_as_int(mn.w)) { ML_ADD_LEAF(__float_as_int(mn.w), __float_as_int(mx.w)); return nc; }
    int stack[64];
    int sp = 0;
    stack[sp++] = __float_as_int(mx.w);
    while (sp > 0) {
        int e = stack[--sp];
        if (e < 0) {                                    // a leaf that had to wait for its left sibling's subtree
            int cnt = __float_as_int(nodeMin[~e].w), idx = __float_as_int(nodeMax[~e].w);
            ML_ADD_LEAF(cnt, idx);
            continue;
        }
        float4 lmn = nodeMin[e], lmx = nodeMax[e], rmn = nodeMin[e + 1], rmx = nodeMax[e + 1];
        bool ol = aabbHits(lb, lmn, lmx), orr = aabbHits(lb, rmn, rmx);
        int lcnt = __float_as_int(lmn.w), lidx = __float_as_int(lmx.w), rcnt = __float_as_int(rmn.w), ridx = __float_as_int(rmx.w);
        bool leftSubtree = ol && !lcnt;
        if (ol && lcnt) ML_ADD_LEAF(lcnt, lidx);
        if (orr) {
            if (rcnt && !leftSubtree) ML_ADD_LEAF(rcnt, ridx);
            else ML_PUSH(rcnt ? ~(e + 1) : ridx);
        }
        if (leftSubtree) ML_PUSH(lidx);
    }
#undef ML_ADD_LEAF
#undef ML_PUSH
    return nc;
}

template <int TYPE>
__device__ __forceinline__ bool lightTriTest(V3 localPos, Q4 localOr, float rA, float rB, float4 r0, float4 r1, float4 r2, TriContact& tc) {
    V3 a = mk3(r0), b = mk3(r1), c = mk3(r2), n = mk3(r0.w, r1.w, r2.w);
    tc.boxFeature = 0; tc.boxAxis = 0; tc.fidx = 0; tc.feature = TF_FACE; tc.dist = 0.f;
    tc.normal = tc.cpBody = tc.cpTri = mk3(0.f);
    if (TYPE == PB_SPHERE) return sphereTriangle(localPos, rA, a, b, c, n, tc);
    return capsuleTriangle(localPos, localOr, rA, rB, a, b, c, n, tc);
}

template <int TYPE>
__global__ void __launch_bounds__(32 * ML_WARPS) k_np_mesh_light(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                 const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                                 const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                 const PbTriMeshDev* __restrict__ meshes,
                                                 int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds,
                                                 int* __restrict__ spillList) {
    const int BIN = BIN_MESH_S + TYPE;
    int start = counters[CNT_BINSTART + BIN], end = counters[CNT_BINSTART + BIN + 1];
    const int lane = threadIdx.x & 31;
    for (int base = start + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); base < end; base += gridDim.x * blockDim.x) {
        int idx = base + lane;
        TriContact contacts[PB_MAX_TRI_CONTACTS];
        unsigned char order[PB_MAX_TRI_CONTACTS];
        int cand[PB_MAX_TRI_CAND];
        int nc = 0, cnt = 0, nGen = 0;
        int a = 0, b = 0;
        bool flip = false;
        int ovf = 0, pi = 0;
        V3 pos1 = mk3(0.f), localPos = mk3(0.f); Q4 or1 = mkq(make_float4(0, 0, 0, 1)), localOr = or1;
        float4 prm = make_float4(0, 0, 0, 0);
        int meshId = 0;
        const float4* rec = nullptr;
        if (idx < end) {
            pi = pairOrder[idx];
            int2 p = pairs[pi];
            a = p.x; b = p.y;
            flip = colType[a] == PB_TRIANGLE_MESH;      // Collision.cpp:897-907: mesh on side 0 -> collide (shape1, mesh0) and flip
            int shape = flip ? b : a;
            int meshCol = flip ? a : b;
            prm = colParams[shape];
            meshId = colMesh[meshCol];
            V3 pos0 = mk3(wpos[shape]); Q4 or0 = mkq(wquat[shape]);
            pos1 = mk3(wpos[meshCol]); or1 = mkq(wquat[meshCol]);
            Q4 invOr1 = qinverse(or1);
            localPos = rotate(invOr1, pos0 - pos1);
            localOr = qmul(invOr1, or0);
            Aabb lb = shapeBounds(localPos, localOr, TYPE, prm, nullptr, 0);
            const PbTriMeshDev& mesh = meshes[meshId];
            rec = mesh.triRec;
            nc = meshCullDual(lb, mesh.nodeMin, mesh.nodeMax, cand, &ovf);
            if (ovf) nc = 0;         // goes to the spill kernel: no point in testing the candidates that fit
        }
        {
            // per-pair loop in candidate order, the next triangle's record in flight while this one is tested
            float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;
            if (nc > 0) { const float4* r = rec + 4 * (size_t)cand[0]; q0 = r[0]; q1 = r[1]; q2 = r[2]; }
            for (int i = 0; i < nc; ++i) {
                float4 r0 = q0, r1 = q1, r2 = q2;
                if (i + 1 < nc) { const float4* r = rec + 4 * (size_t)cand[i + 1]; q0 = r[0]; q1 = r[1]; q2 = r[2]; }
                TriContact tc;
                if (lightTriTest<TYPE>(localPos, localOr, prm.x, prm.y, r0, r1, r2, tc)) {
                    tc.tri = cand[i];
                    if (cnt < PB_MAX_TRI_CONTACTS) contacts[cnt++] = tc;
                    else ovf |= PB_CAUSE_SPILLED_TRI_CONTACTS;
                }
            }
        }
        if (ovf) spillAppend(counters, CNT_SPILL_MESH, spillList, pi, ovf);
        else if (idx < end) nGen = meshFilter(meshes[meshId], contacts, cnt, order);
        int slot = warpReserve(nGen, &counters[CNT_RAWM]);
        for (int g = 0; g < nGen; ++g) {
            const TriContact& tc = contacts[order[g]];
            Manifold m; m.np = 0; m.tri = tc.tri;
            if (TYPE == PB_CAPSULE && tc.feature == TF_FACE) {
                if (!capsuleTriangleFaceManifold(localPos, localOr, prm.x, prm.y, pos1, or1, meshes[meshId], tc, m)) { m.np = 0; m.n = mk3(0.f, 1.f, 0.f); }
            } else {
                manifoldFromClosest(pos1, or1, tc, m);
            }
            m.tri = tc.tri;
            storeManifold(slot + g, maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
        }
    }
}

// ---- spill kernels: pairs that outgrew the per-thread containers ----------------------------------------------------------------
// The reference keeps per-pair work in std::vectors (EPA polytope EPA.h:22-124, overlapBvh's triangle list TriangleMesh.cpp:166-192,
// the contact list CTM.cpp:893-907) and in 128-point clip buffers (Clipping.cpp:6).  The bin kernels above hold them in per-thread
// local memory, sized for the common case; a pair that needs more lands on a spill list and is redone here from scratch with the
// same routines on global-memory scratch sized far beyond anything observed (LimSpill; MS_* below).  Exceeding even those sets
// PB_CAUSE_SPILL_SCRATCH and keeps what fits -- a step never fails because of one pair.

// GJK / EPA bin: one thread per spilled pair, polytope in epaScratch[global thread].
__global__ void __launch_bounds__(PB_SPILL_GJK_THREADS) k_np_gjk_spill(const int2* __restrict__ pairs, const int* __restrict__ spillList, int* __restrict__ counters,
                                                     const int* __restrict__ colType, const float4* __restrict__ colParams,
                                                     const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                     const PbConvexDev* __restrict__ convexes, const int* __restrict__ colMesh,
                                                     LimSpill::Epa* __restrict__ epaScratch,
                                                     int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds) {
    const int n = min(counters[CNT_SPILL_GJK], PB_SPILL_CAP);
    if (n == 0) return;
    if (threadIdx.x == 0) atomicAdd(&counters[CNT_SPILLED], n);
    LimSpill::Epa& poly = epaScratch[threadIdx.x];
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        int2 p = pairs[spillList[e]];
        const int a = p.x, b = p.y;
        Manifold m; m.np = 0; m.tri = -1;
        bool flip = false;
        int ovf = 0;
        bool hit = collideGjkPair<LimSpill>(colType[a], colParams[a], mk3(wpos[a]), mkq(wquat[a]), colMesh[a],
                                            colType[b], colParams[b], mk3(wpos[b]), mkq(wquat[b]), colMesh[b], convexes, m, flip, poly, &ovf);
        if (ovf) atomicOr(&counters[CNT_CAUSE], PB_CAUSE_SPILL_SCRATCH);
        if (hit && m.np > 0) storeManifold(atomicAdd(&counters[CNT_RAWM], 1), maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
    }
}

// Mesh bins (all four shape kinds): one WARP per spilled pair.  Lane 0 walks the mesh BVH in the reference's order into the
// candidate list; the lanes test 32 candidates at a time and append the hits in candidate order (ballot prefix); the feature
// filter runs warp-wide with identical control flow in every lane (the voided-vertex lookups are strided over the lanes); the
// surviving contacts' manifolds are generated one per lane.
#define MS_CAND 65536
#define MS_CONTACTS 8192
#define MS_STACK 4096
struct MeshSpillScratch {
    int cand[MS_CAND];
    TriContact contacts[MS_CONTACTS];
    unsigned short live[MS_CONTACTS], order[MS_CONTACTS];
    unsigned int voidSet[3 * MS_CONTACTS];
    int stack[MS_STACK];
};

__device__ __forceinline__ bool voidedW(const unsigned int* set, int n, unsigned int v, int lane) {
    bool f = false;
    for (int i = lane; i < n; i += 32) f |= set[i] == v;
    return __any_sync(0xffffffffu, f);
}
__device__ __forceinline__ void voidInsertW(unsigned int* set, int& n, unsigned int v, int lane) {
    if (voidedW(set, n, v, lane)) return;
    if (lane == 0) set[n] = v;
    ++n;
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_np_mesh_spill(const int2* __restrict__ pairs, const int* __restrict__ spillList, int* __restrict__ counters,
                                                       const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                                       const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                       const PbTriMeshDev* __restrict__ meshes, const PbConvexDev* __restrict__ convexes,
                                                       MeshSpillScratch* __restrict__ scratch, LimSpill::Epa* __restrict__ epaScratch,
                                                       int4* __restrict__ mKey, float4* __restrict__ mNormal, float4* __restrict__ mPts, int maxManifolds) {
    const int n = min(counters[CNT_SPILL_MESH], PB_SPILL_CAP);
    if (n == 0) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    if (warp == 0 && lane == 0) atomicAdd(&counters[CNT_SPILLED], n);
    MeshSpillScratch& S = scratch[warp];
    LimSpill::Epa* myEpa = epaScratch ? epaScratch + PB_SPILL_GJK_THREADS + (size_t)warp * 32 + lane : nullptr;
    for (int e = warp; e < n; e += nWarps) {
        int2 p = pairs[spillList[e]];
        const int a = p.x, b = p.y;
        const bool flip = colType[a] == PB_TRIANGLE_MESH;      // Collision.cpp:897-907
        const int shape = flip ? b : a, meshCol = flip ? a : b;
        const int type = colType[shape];
        const float4 prm = colParams[shape];
        const PbTriMeshDev& mesh = meshes[colMesh[meshCol]];
        const V3 pos0 = mk3(wpos[shape]); const Q4 or0 = mkq(wquat[shape]);
        const V3 pos1 = mk3(wpos[meshCol]); const Q4 or1 = mkq(wquat[meshCol]);
        const Q4 invOr1 = qinverse(or1);
        const V3 localPos = rotate(invOr1, pos0 - pos1);
        const Q4 localOr = qmul(invOr1, or0);
        const int convexId = colMesh[shape];
        int ovf = 0;
        // ---- cull (TriangleMesh.cpp:166-192), lane 0
        int nc = 0;
        if (lane == 0) {
            Aabb lb = shapeBounds(localPos, localOr, type, prm, convexes, convexId);
            int sp = 0;
            S.stack[sp++] = 0;
            while (sp > 0) {
                int node = S.stack[--sp];
                float4 nmn = mesh.nodeMin[node], nmx = mesh.nodeMax[node];
                if (!aabbHits(lb, nmn, nmx)) continue;
                int triCount = __float_as_int(nmn.w), index = __float_as_int(nmx.w);
                if (triCount) {
                    for (int k = 0; k < triCount; ++k) { if (nc < MS_CAND) S.cand[nc++] = index + k; else ovf |= PB_CAUSE_SPILL_SCRATCH; }
                } else if (sp + 2 <= MS_STACK) { S.stack[sp++] = index + 1; S.stack[sp++] = index; }
                else ovf |= PB_CAUSE_SPILL_SCRATCH;
            }
        }
        nc = __shfl_sync(0xffffffffu, nc, 0);
        __syncwarp();
        // ---- per-triangle tests, 32 candidates at a time, hits appended in candidate order (CTM.cpp:895-907)
        Shape convexShape;
        if (type == PB_CONVEX_MESH) convexShape = makeShape(type, prm, localPos, localOr, convexes, convexId);
        int cnt = 0;
        for (int base = 0; base < nc; base += 32) {
            const int i = base + lane;
            bool hit = false;
            TriContact tc;
            if (i < nc) {
                const int tri = S.cand[i];
                int4 ti = mesh.tris[tri];
                V3 va = mk3(mesh.verts[ti.x]), vb = mk3(mesh.verts[ti.y]), vc = mk3(mesh.verts[ti.z]);
                V3 nn = mk3(mesh.triNormal[tri]);
                tc.boxFeature = 0; tc.boxAxis = 0; tc.fidx = 0; tc.feature = TF_FACE; tc.dist = 0.f;
                tc.normal = tc.cpBody = tc.cpTri = mk3(0.f);
                if (type == PB_SPHERE) hit = sphereTriangle(localPos, prm.x, va, vb, vc, nn, tc);
                else if (type == PB_CAPSULE) hit = capsuleTriangle(localPos, localOr, prm.x, prm.y, va, vb, vc, nn, tc);
                else if (type == PB_BOX) hit = boxTriangle(localPos, localOr, mk3(prm.x, prm.y, prm.z), va, vb, vc, nn, tc);
                else hit = convexTriangle(convexShape, va, vb, vc, mk3(mesh.triCentroid[tri]), tc, *myEpa, &ovf);
                tc.tri = tri;
            }
            const unsigned int mask = __ballot_sync(0xffffffffu, hit);
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (hit) { if (pos < MS_CONTACTS) S.contacts[pos] = tc; else ovf |= PB_CAUSE_SPILL_SCRATCH; }
            cnt += __popc(mask);
        }
        if (cnt > MS_CONTACTS) cnt = MS_CONTACTS;
        __syncwarp();
        // ---- feature filter (CTM.cpp:913-953), every lane in lock step
        int nGen = 0, nVoid = 0, nLive = cnt;
        for (int i = lane; i < cnt; i += 32) S.live[i] = (unsigned short)i;
        __syncwarp();
        for (int i = nLive - 1; i >= 0; --i) {
            const int ci = S.live[i];
            if (S.contacts[ci].feature == TF_FACE) {
                int4 ti = mesh.tris[S.contacts[ci].tri];
                voidInsertW(S.voidSet, nVoid, (unsigned)ti.x, lane); voidInsertW(S.voidSet, nVoid, (unsigned)ti.y, lane); voidInsertW(S.voidSet, nVoid, (unsigned)ti.z, lane);
                const unsigned short lastLive = S.live[nLive - 1];
                __syncwarp();
                if (lane == 0) { S.order[nGen] = (unsigned short)ci; S.live[i] = lastLive; }
                ++nGen; --nLive;
                __syncwarp();
            }
        }
        if (lane == 0) sortLikeStdSort(S.live, nLive, [&](unsigned short x, unsigned short y) { return S.contacts[x].dist < S.contacts[y].dist; });
        __syncwarp();
        for (int i = 0; i < nLive; ++i) {
            const int ci = S.live[i];
            const int feature = S.contacts[ci].feature, fidx = S.contacts[ci].fidx;
            int4 ti = mesh.tris[S.contacts[ci].tri];
            unsigned int vi[3] = { (unsigned)ti.x, (unsigned)ti.y, (unsigned)ti.z };
            if (feature == TF_EDGE) {
                if (voidedW(S.voidSet, nVoid, vi[(fidx + 1) % 3], lane) && voidedW(S.voidSet, nVoid, vi[(fidx + 2) % 3], lane)) continue;
            } else {
                if (voidedW(S.voidSet, nVoid, vi[fidx], lane)) continue;
            }
            if (lane == 0) S.order[nGen] = (unsigned short)ci;
            ++nGen;
            voidInsertW(S.voidSet, nVoid, vi[0], lane); voidInsertW(S.voidSet, nVoid, vi[1], lane); voidInsertW(S.voidSet, nVoid, vi[2], lane);
        }
        __syncwarp();
        // ---- manifolds, one per lane (CTM.cpp:826-880 dispatch)
        int slot = 0;
        if (lane == 0 && nGen) slot = atomicAdd(&counters[CNT_RAWM], nGen);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        for (int g = lane; g < nGen; g += 32) {
            const TriContact tc = S.contacts[S.order[g]];
            Manifold m; m.np = 0; m.tri = tc.tri;
            if (type == PB_SPHERE || type == PB_CAPSULE) {
                if (type == PB_CAPSULE && tc.feature == TF_FACE) {
                    if (!capsuleTriangleFaceManifold(localPos, localOr, prm.x, prm.y, pos1, or1, mesh, tc, m)) { m.np = 0; m.n = mk3(0.f, 1.f, 0.f); }
                } else manifoldFromClosest(pos1, or1, tc, m);
            } else if (type == PB_BOX) {
                V3 he = mk3(prm.x, prm.y, prm.z);
                if (tc.feature == TF_FACE) boxTriangleFaceManifold(localPos, localOr, he, pos1, or1, mesh, tc, m);
                else if (tc.feature == TF_EDGE && tc.boxFeature != BOXF_FACE) boxEdgeTriangleEdgeManifold(localPos, localOr, he, pos1, or1, mesh, tc, m);
                else boxFaceTriangleManifold(localPos, localOr, he, pos1, or1, mesh, tc, m);
            } else {
                const PbConvexDev& cm = convexes[convexId];
                V3 sc = mk3(prm.x, prm.y, prm.z);
                if (tc.feature == TF_FACE) convexTriangleFaceManifold<LimSpill>(localPos, localOr, cm, sc, pos1, or1, mesh, tc, m, &ovf);
                else convexFaceTriangleManifold<LimSpill>(localPos, localOr, cm, sc, pos1, or1, mesh, tc, m, &ovf);
            }
            m.tri = tc.tri;
            storeManifold(slot + g, maxManifolds, a, b, m, flip, mKey, mNormal, mPts, counters);
        }
        if (ovf) atomicOr(&counters[CNT_CAUSE], PB_CAUSE_SPILL_SCRATCH);
        __syncwarp();
    }
}

// scratch of the spill kernels, allocated the first time a scene could need it (convex meshes / triangle meshes registered)
static int ensureSpillScratch(pb_ctx* ctx) {
    int rc;
    if (!ctx->spillList) {
        int* p = nullptr;
        if ((rc = pb_alloc(ctx, &p, 2 * (size_t)PB_SPILL_CAP))) return rc;
        ctx->spillList = p;
    }
    if (!ctx->convexes.empty() && !ctx->spillEpa) {
        LimSpill::Epa* p = nullptr;
        if ((rc = pb_alloc(ctx, &p, (size_t)PB_SPILL_GJK_THREADS + 32 * (size_t)PB_SPILL_MESH_WARPS))) return rc;
        ctx->spillEpa = p;
    }
    if (!ctx->triMeshes.empty() && !ctx->spillMesh) {
        MeshSpillScratch* p = nullptr;
        if ((rc = pb_alloc(ctx, &p, (size_t)PB_SPILL_MESH_WARPS))) return rc;
        ctx->spillMesh = p;
    }
    return PB_OK;
}

static void launchSpill(pb_ctx* ctx, const int2* pairs, int* counters, int4* mKey, float4* mNormal, float4* mPts, int cap) {
    if (!ctx->convexes.empty())
        ++ctx->launches, k_np_gjk_spill<<<1, PB_SPILL_GJK_THREADS, 0, ctx->stream>>>(pairs, ctx->spillList, counters, ctx->colType, ctx->colParams, ctx->colWPos, ctx->colWQuat,
            ctx->convexDev, ctx->colMesh, (LimSpill::Epa*)ctx->spillEpa, mKey, mNormal, mPts, cap);
    if (!ctx->triMeshes.empty())
        ++ctx->launches, k_np_mesh_spill<<<PB_SPILL_MESH_WARPS / 4, 128, 0, ctx->stream>>>(pairs, ctx->spillList + PB_SPILL_CAP, counters, ctx->colType, ctx->colParams, ctx->colMesh,
            ctx->colWPos, ctx->colWQuat, ctx->triMeshDev, ctx->convexDev, (MeshSpillScratch*)ctx->spillMesh, (LimSpill::Epa*)ctx->spillEpa, mKey, mNormal, mPts, cap);
}

// ---- trigger pairs (Physecs.cpp:200-207) --------------------------------------------------------------------------------
// One thread per TRIGGER-classified pair: physecs::overlap on the world collider poses; overlapping pairs are appended
// (collider indices, side 0 = lower entity) for the enter / exit diff the host runs after the step (Physecs.cpp:538-552).
__global__ void __launch_bounds__(128) k_np_trigger(const int2* __restrict__ pairs, const int* __restrict__ pairOrder, int* __restrict__ counters,
                                                    const int* __restrict__ colType, const float4* __restrict__ colParams, const int* __restrict__ colMesh,
                                                    const float4* __restrict__ wpos, const float4* __restrict__ wquat,
                                                    const PbConvexDev* __restrict__ convexes, int2* __restrict__ trigPairs, int maxTrig) {
    int start = counters[CNT_BINSTART + BIN_TRIGGER], end = counters[CNT_BINSTART + BIN_TRIGGER + 1];
    int lane = threadIdx.x & 31;
    for (int base = start + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31); base < end; base += gridDim.x * blockDim.x) {
        int idx = base + lane;
        bool hit = false;
        int2 p = make_int2(0, 0);
        if (idx < end) {
            p = pairs[pairOrder[idx]];
            hit = overlapShapes(colType[p.x], colParams[p.x], mk3(wpos[p.x]), mkq(wquat[p.x]), colMesh[p.x],
                                colType[p.y], colParams[p.y], mk3(wpos[p.y]), mkq(wquat[p.y]), colMesh[p.y], convexes);
        }
        int slot = warpReserve(hit ? 1 : 0, &counters[CNT_TRIGGERS]);
        if (hit) {
            if (slot < maxTrig) trigPairs[slot] = p;
            else { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_TRIGGERS); }
        }
    }
}

// ---- query mode (Scene::overlapWithMinTranslationalDistance, Physecs.cpp:652-688) ------------------------------------------------
// physecs::collision(collider, query shape) for every candidate pair: the SAME bin kernels as the step, pointed at a private
// arena.  No trigger / nonColliding filter: the reference's query calls collision() directly (:659).
__global__ void k_query_classify(const int2* __restrict__ pairs, int* __restrict__ pairBin, int* __restrict__ counters, int cap,
                                 const int* __restrict__ colType) {
    int n = min(counters[CNT_PAIRS], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int2 p = pairs[i];
        int bin = binOf(colType[p.x], colType[p.y]);
        if (bin >= 0) atomicAdd(&counters[CNT_BIN0 + bin], 1);
        pairBin[i] = bin;
    }
}

// Grid of a grid-stride bin kernel: a whole number of waves of co-resident CTAs (occupancy x SMs x waves).  numSMs * 8 CTAs of a
// kernel that fits 7 per SM ran as one full wave plus a 1/7 wave: the mesh bins lost a third of their time to that tail.
// `heavy`: the GJK / EPA kernels, whose per-pair cost varies by an order of magnitude, get twice the waves (250 k convex pile: 4.08 -> 3.91 ms).
template <typename K>
static int npGrid(pb_ctx* ctx, K kernel, int threads, bool heavy = false) {
    if (ctx->npWaves <= 0) return ctx->numSMs * 8;
    // occupancy per kernel, asked once per process (every device of a box is the same part)
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    int occ = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find((const void*)kernel);
        if (it != cache.end()) occ = it->second;
        else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 4; }
            cache[(const void*)kernel] = occ;
        }
    }
    int grid = ctx->numSMs * occ * ctx->npWaves * (heavy ? 2 : 1);
    if (ctx->pairsHint >= 0) {            // a bin holds at most every pair: no point in launching CTAs far beyond that on small scenes
        long long need = (2LL * ctx->pairsHint + 4096) / threads + ctx->numSMs;
        if (need < grid) grid = (int)need;
    }
    return grid;
}

static void launchMeshLight(pb_ctx* ctx, const int2* pairs, const int* pairOrder, int* counters, int4* mKey, float4* mNormal, float4* mPts, int cap, int blocksOverride) {
#define LAUNCH_LIGHT(TYPE) ++ctx->launches, k_np_mesh_light<TYPE><<<blocksOverride ? blocksOverride : npGrid(ctx, k_np_mesh_light<TYPE>, 32 * ML_WARPS), 32 * ML_WARPS, 0, ctx->stream>>>( \
        pairs, pairOrder, counters, ctx->colType, ctx->colParams, ctx->colMesh, ctx->colWPos, ctx->colWQuat, ctx->triMeshDev, mKey, mNormal, mPts, cap, ctx->spillList + PB_SPILL_CAP)
    LAUNCH_LIGHT(PB_CAPSULE); LAUNCH_LIGHT(PB_SPHERE);
#undef LAUNCH_LIGHT
}

int pb_narrowphase_query(pb_ctx* ctx, int* counters, const int2* pairs, int* pairOrder, int cap, int4* mKey, float4* mNormal, float4* mPts) {
    const int blocks = 8;
    int* pairBin = pairOrder + cap;
    { int rc = ensureSpillScratch(ctx); if (rc) return rc; }
    ++ctx->launches, k_query_classify<<<blocks, 256, 0, ctx->stream>>>(pairs, pairBin, counters, cap, ctx->colType);
    ++ctx->launches, k_bin_starts<<<1, 32, 0, ctx->stream>>>(counters);
    ++ctx->launches, k_pair_scatter<<<1, SCATTER_THREADS, 0, ctx->stream>>>(pairBin, pairOrder, counters, cap);
#define LAUNCH_PRIM(BIN) ++ctx->launches, k_np_prim<BIN><<<blocks, 128, 0, ctx->stream>>>(pairs, pairOrder, counters, ctx->colType, ctx->colParams, \
        ctx->colWPos, ctx->colWQuat, ctx->convexDev, ctx->colMesh, mKey, mNormal, mPts, cap, ctx->spillList)
    LAUNCH_PRIM(BIN_SS); LAUNCH_PRIM(BIN_SC); LAUNCH_PRIM(BIN_CC); LAUNCH_PRIM(BIN_SB); LAUNCH_PRIM(BIN_CB); LAUNCH_PRIM(BIN_BB);
    if (!ctx->convexes.empty()) LAUNCH_PRIM(BIN_GJK);
#undef LAUNCH_PRIM
#define LAUNCH_MESH(TYPE) ++ctx->launches, k_np_mesh<TYPE><<<blocks, 128, 0, ctx->stream>>>(pairs, pairOrder, counters, ctx->colType, ctx->colParams, ctx->colMesh, \
        ctx->colWPos, ctx->colWQuat, ctx->triMeshDev, ctx->convexDev, mKey, mNormal, mPts, cap, ctx->spillList + PB_SPILL_CAP)
    if (!ctx->triMeshes.empty()) {
        if (!ctx->convexes.empty()) LAUNCH_MESH(PB_CONVEX_MESH);
        LAUNCH_MESH(PB_BOX);
        if (ctx->meshLightMode == 0) { LAUNCH_MESH(PB_CAPSULE); LAUNCH_MESH(PB_SPHERE); }
        else launchMeshLight(ctx, pairs, pairOrder, counters, mKey, mNormal, mPts, cap, blocks);
    }
#undef LAUNCH_MESH
    launchSpill(ctx, pairs, counters, mKey, mNormal, mPts, cap);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_narrowphase(pb_ctx* ctx) {
    if (ctx->nCol < 2) return PB_OK;
    const int blocks = pb_hint_grid(ctx->pairsHint, 256, ctx->numSMs * 8);
    int* pairBin = ctx->pairOrder + ctx->caps.max_pairs;   // second half of the pairOrder allocation
    { int rc = ensureSpillScratch(ctx); if (rc) return rc; }
    ++ctx->launches, k_pair_classify<<<blocks, 256, 0, ctx->stream>>>((const int2*)ctx->pairs, pairBin, ctx->counters, ctx->caps.max_pairs, ctx->colType, ctx->colFlags,
                                                      ctx->colRow, ctx->rowEntity, ctx->nonColliding, ctx->nNonColliding,
                                                      ctx->colClass, ctx->filterLut, ctx->nFilterClasses);
    ++ctx->launches, k_bin_starts<<<1, 32, 0, ctx->stream>>>(ctx->counters);
    ++ctx->launches, k_pair_scatter<<<pb_hint_grid(ctx->pairsHint, SCATTER_THREADS, ctx->numSMs * 2), SCATTER_THREADS, 0, ctx->stream>>>(pairBin, ctx->pairOrder, ctx->counters, ctx->caps.max_pairs);
    const int2* pairs = (const int2*)ctx->pairs;
#define LAUNCH_PRIM(BIN) ++ctx->launches, k_np_prim<BIN><<<npGrid(ctx, k_np_prim<BIN>, 128), 128, 0, ctx->stream>>>(pairs, ctx->pairOrder, ctx->counters, ctx->colType, ctx->colParams, \
        ctx->colWPos, ctx->colWQuat, ctx->convexDev, ctx->colMesh, ctx->mKey, ctx->mNormal, ctx->mPts, ctx->caps.max_manifolds, ctx->spillList)
    if (ctx->pairsHint >= 0 && ctx->pairsHint <= 131072 && ctx->npFuseSmall) {
        const int gx = std::max(1, std::min(ctx->numSMs * 2, (ctx->pairsHint + 127) / 128 + 1));
        ++ctx->launches, k_np_prim_all<<<dim3(gx, 6), 128, 0, ctx->stream>>>(pairs, ctx->pairOrder, ctx->counters, ctx->colType, ctx->colParams, ctx->colWPos, ctx->colWQuat,
            ctx->convexDev, ctx->colMesh, ctx->mKey, ctx->mNormal, ctx->mPts, ctx->caps.max_manifolds, ctx->spillList);
    } else {
        LAUNCH_PRIM(BIN_SS); LAUNCH_PRIM(BIN_SC); LAUNCH_PRIM(BIN_CC); LAUNCH_PRIM(BIN_SB); LAUNCH_PRIM(BIN_CB); LAUNCH_PRIM(BIN_BB);
    }
    if (!ctx->convexes.empty()) {
        // hit list: one entry per intersecting pair, i.e. per manifold of the bin -> the manifold capacity bounds it
        int want = ctx->caps.max_manifolds < ctx->caps.max_pairs ? ctx->caps.max_manifolds : ctx->caps.max_pairs;
        if (ctx->gjkHitCap < want) {
            int rc;
            if ((rc = pb_alloc(ctx, &ctx->gjkHitPair, (size_t)want)) || (rc = pb_alloc(ctx, &ctx->gjkHitSimplex, 9 * (size_t)want))) return rc;
            ctx->gjkHitCap = want;
        }
        ++ctx->launches, k_np_gjk_hits<<<npGrid(ctx, k_np_gjk_hits, 128, true), 128, 0, ctx->stream>>>(pairs, ctx->pairOrder, ctx->counters, ctx->colType, ctx->colParams, ctx->colWPos, ctx->colWQuat,
                                                                     ctx->convexDev, ctx->colMesh, ctx->gjkHitPair, ctx->gjkHitSimplex, ctx->gjkHitCap);
        ++ctx->launches, k_np_gjk_manifolds<<<npGrid(ctx, k_np_gjk_manifolds, 128, true), 128, 0, ctx->stream>>>(pairs, ctx->counters, ctx->colType, ctx->colParams, ctx->colWPos, ctx->colWQuat, ctx->convexDev,
                                                                          ctx->colMesh, ctx->gjkHitPair, ctx->gjkHitSimplex, ctx->gjkHitCap, ctx->mKey, ctx->mNormal, ctx->mPts,
                                                                          ctx->caps.max_manifolds, ctx->spillList);
    }
#undef LAUNCH_PRIM
#define LAUNCH_MESH(TYPE) ++ctx->launches, k_np_mesh<TYPE><<<npGrid(ctx, k_np_mesh<TYPE>, 128), 128, 0, ctx->stream>>>(pairs, ctx->pairOrder, ctx->counters, ctx->colType, ctx->colParams, ctx->colMesh, \
        ctx->colWPos, ctx->colWQuat, ctx->triMeshDev, ctx->convexDev, ctx->mKey, ctx->mNormal, ctx->mPts, ctx->caps.max_manifolds, ctx->spillList + PB_SPILL_CAP)
    if (!ctx->triMeshes.empty()) {
        // heavy shapes first: their long threads overlap with the tail of nothing else, the light bins fill in after
        if (!ctx->convexes.empty()) LAUNCH_MESH(PB_CONVEX_MESH);
        LAUNCH_MESH(PB_BOX);
        if (ctx->meshLightMode == 0) { LAUNCH_MESH(PB_CAPSULE); LAUNCH_MESH(PB_SPHERE); }
        else launchMeshLight(ctx, pairs, ctx->pairOrder, ctx->counters, ctx->mKey, ctx->mNormal, ctx->mPts, ctx->caps.max_manifolds, 0);
    }
#undef LAUNCH_MESH
    launchSpill(ctx, pairs, ctx->counters, ctx->mKey, ctx->mNormal, ctx->mPts, ctx->caps.max_manifolds);
    if (ctx->triggersPossible)
        ++ctx->launches, k_np_trigger<<<npGrid(ctx, k_np_trigger, 128), 128, 0, ctx->stream>>>(pairs, ctx->pairOrder, ctx->counters, ctx->colType, ctx->colParams, ctx->colMesh, ctx->colWPos,
                                                    ctx->colWQuat, ctx->convexDev, ctx->trigPairs, ctx->caps.max_pairs);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// Broadphase: per-collider world AABBs, 30-bit Morton keys, radix sort, Karras LBVH build, bottom-up
// refit and a stack traversal that emits overlapping collider pairs through warp-aggregated appends.
//
// Replaces (reference, /root/reference):
//   Scene::updateBounds            src/Physecs.cpp:79-90   + getBounds*  src/BoundsUtil.cpp:23-85
//   SAP insertion sort + sweep      src/Physecs.cpp:119-173 (pair predicates :138-154, ordering :158-168)
// The pair SET is identical to the sweep's: two closed AABBs overlap on all three axes, both colliders
// have enableSimulation, they belong to different entities and at least one owner is a non-kinematic
// dynamic body.  Built with -fmad=false so bounds equal the reference's fp32 values bit for bit.
#include "pb_ctx.h"
#include "pb_math.cuh"
#include "np_bounds.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define COLF_TRIGGER 1
#define COLF_ENABLE 2
#define COLF_DYNAMIC 4
#define COLF_BIG 8            // set per tree build: the collider sits on the step's big-static list instead of in the tree
#define PB_BIG_MAX 32

// ---- ordered-int float encoding for atomic min/max -------------------------------------------------------
__device__ __forceinline__ int floatToOrdered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float orderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// mode 0: all colliders; mode 1: colliders of non-kinematic dynamic bodies; mode 2: colliders whose row is flagged in rowMark
__global__ void k_update_bounds(int n, int mode, const int* __restrict__ rowMark, float margin,
                                const int* __restrict__ colRow, const int* __restrict__ colType, const int* __restrict__ colFlags,
                                const int* __restrict__ colMesh, const float4* __restrict__ colLPos, const float4* __restrict__ colLQuat,
                                const float4* __restrict__ colParams, const float4* __restrict__ pos, const float4* __restrict__ quat,
                                const PbConvexDev* __restrict__ convexes, float4* __restrict__ aabbMin, float4* __restrict__ aabbMax,
                                const int* __restrict__ skipStatus) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (skipStatus && (*skipStatus & (PB_ECAPACITY | 0x100))) return;      // end-of-step refresh of a step whose solve was skipped (capi.cu collectStep)
    int type = colType[i];
    if (type == PB_TRIANGLE_MESH) return;  // separate reduction kernel
    int flags = colFlags[i];
    int row = colRow[i];
    if (mode == 1 && !(flags & COLF_DYNAMIC)) return;
    if (mode == 2 && !rowMark[row]) return;
    V3 bp = mk3(pos[row]); Q4 bq = mkq(quat[row]);
    V3 wp = bp + rotate(bq, mk3(colLPos[i]));
    Q4 wq = qmul(bq, mkq(colLQuat[i]));
    Aabb b = shapeBounds(wp, wq, type, colParams[i], convexes, colMesh[i]);
    b.mx = b.mx + mk3(margin);
    b.mn = b.mn - mk3(margin);
    aabbMin[i] = f4(b.mn); aabbMax[i] = f4(b.mx);
}

// getBoundsTriangleMesh (BoundsUtil.cpp:77-85): min/max over every transformed vertex
__global__ void k_trimesh_bounds(int col, const int* __restrict__ colRow, const float4* __restrict__ colLPos,
                                 const float4* __restrict__ colLQuat, const float4* __restrict__ pos, const float4* __restrict__ quat,
                                 const float4* __restrict__ verts, int nVerts, int* __restrict__ acc) {
    int row = colRow[col];
    V3 bp = mk3(pos[row]); Q4 bq = mkq(quat[row]);
    V3 wp = bp + rotate(bq, mk3(colLPos[col]));
    Q4 wq = qmul(bq, mkq(colLQuat[col]));
    V3 mn = mk3(FLT_MAX), mx = mk3(-FLT_MAX);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nVerts; i += gridDim.x * blockDim.x) {
        V3 p = wp + rotate(wq, mk3(verts[i]));
        mn = vmin(mn, p); mx = vmax(mx, p);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, d)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, d));
        mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, d));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, d)); mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, d));
        mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&acc[0], floatToOrdered(mn.x)); atomicMin(&acc[1], floatToOrdered(mn.y)); atomicMin(&acc[2], floatToOrdered(mn.z));
        atomicMax(&acc[3], floatToOrdered(mx.x)); atomicMax(&acc[4], floatToOrdered(mx.y)); atomicMax(&acc[5], floatToOrdered(mx.z));
    }
}
__global__ void k_trimesh_bounds_init(int* acc) {
    if (threadIdx.x < 3) acc[threadIdx.x] = floatToOrdered(FLT_MAX);
    else if (threadIdx.x < 6) acc[threadIdx.x] = floatToOrdered(-FLT_MAX);
}
__global__ void k_trimesh_bounds_store(int col, float margin, const int* __restrict__ acc, float4* aabbMin, float4* aabbMax, const int* __restrict__ skipStatus) {
    if (skipStatus && (*skipStatus & (PB_ECAPACITY | 0x100))) return;
    V3 mn = mk3(orderedToFloat(acc[0]), orderedToFloat(acc[1]), orderedToFloat(acc[2]));
    V3 mx = mk3(orderedToFloat(acc[3]), orderedToFloat(acc[4]), orderedToFloat(acc[5]));
    mx = mx + mk3(margin); mn = mn - mk3(margin);
    aabbMin[col] = f4(mn); aabbMax[col] = f4(mx);
}

static int launchTrimeshBounds(pb_ctx* ctx, int col, float margin, const int* skipStatus = nullptr) {
    int mesh = ctx->hColMesh[col];
    const PbTriMesh& tm = ctx->triMeshes[mesh];
    int* acc = (int*)ctx->sceneBounds + 8;
    ++ctx->launches, k_trimesh_bounds_init<<<1, 32, 0, ctx->stream>>>(acc);
    int blocks = pb_grid(tm.nVerts, 256); if (blocks > 1024) blocks = 1024;
    ++ctx->launches, k_trimesh_bounds<<<blocks, 256, 0, ctx->stream>>>(col, ctx->colRow, ctx->colLPos, ctx->colLQuat, ctx->pos, ctx->quat, tm.verts, tm.nVerts, acc);
    ++ctx->launches, k_trimesh_bounds_store<<<1, 1, 0, ctx->stream>>>(col, margin, acc, ctx->aabbMin, ctx->aabbMax, skipStatus);
    return PB_OK;
}

// onlyDynamic: the colliders of non-kinematic dynamic bodies (updateBounds(entity) for every moving body, Physecs.cpp:556-559) --
// triangle-mesh colliders riding on such a body included (one vertex reduction each; found through the host mirrors).
// skipStatus: device status word of the step this refresh ends (nullptr outside pb_step).
int pb_update_bounds_all(pb_ctx* ctx, float margin, int onlyDynamic, const int* skipStatus) {
    if (ctx->nCol == 0) return PB_OK;
    ++ctx->launches, k_update_bounds<<<pb_grid(ctx->nCol, 256), 256, 0, ctx->stream>>>(ctx->nCol, onlyDynamic ? 1 : 0, nullptr, margin, ctx->colRow, ctx->colType,
        ctx->colFlags, ctx->colMesh, ctx->colLPos, ctx->colLQuat, ctx->colParams, ctx->pos, ctx->quat, ctx->convexDev, ctx->aabbMin, ctx->aabbMax, skipStatus);
    {
        for (int c : ctx->hTrimeshCols) {
            const int row = ctx->hColRow[c];
            const bool moving = row < ctx->nDyn && row < (int)ctx->hKinematic.size() && !ctx->hKinematic[row];
            if (!onlyDynamic || moving) launchTrimeshBounds(ctx, c, margin, skipStatus);
        }
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// rows moved through registry.patch: dRowMark[row] != 0 marks them (device array of nRows ints)
int pb_update_bounds_rows(pb_ctx* ctx, const int* dRowMark, int n, float margin) {
    if (ctx->nCol == 0) return PB_OK;
    ++ctx->launches, k_update_bounds<<<pb_grid(ctx->nCol, 256), 256, 0, ctx->stream>>>(ctx->nCol, 2, dRowMark, margin, ctx->colRow, ctx->colType,
        ctx->colFlags, ctx->colMesh, ctx->colLPos, ctx->colLQuat, ctx->colParams, ctx->pos, ctx->quat, ctx->convexDev, ctx->aabbMin, ctx->aabbMax, nullptr);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}
int pb_update_bounds_trimesh_col(pb_ctx* ctx, int col, float margin) { return launchTrimeshBounds(ctx, col, margin); }

// world pose of every collider (Physecs.cpp:194-198), refreshed once per step for the narrowphase
__global__ void k_world_pose(int n, const int* __restrict__ colRow, const float4* __restrict__ colLPos, const float4* __restrict__ colLQuat,
                             const float4* __restrict__ pos, const float4* __restrict__ quat, float4* __restrict__ wpos, float4* __restrict__ wquat) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int row = colRow[i];
    V3 bp = mk3(pos[row]); Q4 bq = mkq(quat[row]);
    wpos[i] = f4(bp + rotate(bq, mk3(colLPos[i])));
    wquat[i] = f4(qmul(bq, mkq(colLQuat[i])));
}
int pb_world_poses(pb_ctx* ctx) {
    if (ctx->nCol == 0) return PB_OK;
    ++ctx->launches, k_world_pose<<<pb_grid(ctx->nCol, 256), 256, 0, ctx->stream>>>(ctx->nCol, ctx->colRow, ctx->colLPos, ctx->colLQuat, ctx->pos, ctx->quat, ctx->colWPos, ctx->colWQuat);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// ---- Morton keys --------------------------------------------------------------------------------------------------
__global__ void k_scene_bounds_init(int* sb, int* bigList) {
    if (threadIdx.x < 3) sb[threadIdx.x] = floatToOrdered(FLT_MAX);
    else if (threadIdx.x < 6) sb[threadIdx.x] = floatToOrdered(-FLT_MAX);
    else if (threadIdx.x == 6 && bigList) bigList[0] = 0;
}
__global__ void k_scene_bounds(int n, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, int* __restrict__ sb) {
    V3 mn = mk3(FLT_MAX), mx = mk3(-FLT_MAX);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 c = (mk3(aabbMin[i]) + mk3(aabbMax[i])) * 0.5f;
        // clamp so a single huge static (terrain, ground) does not flatten the key space
        mn = vmin(mn, c); mx = vmax(mx, c);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, d)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, d));
        mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, d));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, d)); mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, d));
        mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, d));
    }
    // block-level reduction first: six atomics per CTA, not per warp (57 k atomics on six addresses were most of this kernel's 45 us)
    __shared__ float red[8][6];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[w][0] = mn.x; red[w][1] = mn.y; red[w][2] = mn.z; red[w][3] = mx.x; red[w][4] = mx.y; red[w][5] = mx.z; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) v = threadIdx.x < 3 ? fminf(v, red[k][threadIdx.x]) : fmaxf(v, red[k][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&sb[threadIdx.x], floatToOrdered(v)); else atomicMax(&sb[threadIdx.x], floatToOrdered(v));
    }
}
__device__ __forceinline__ unsigned int expandBits(unsigned int v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
// Big statics (step trees only, bigList != nullptr): an enabled collider that never issues a query (static / kinematic owner) and
// spans more than 1/8 of the scene -- a terrain, a ground box, the walls of a bin -- would put a chain of scene-sized boxes from the
// root down to its leaf, and EVERY query packet would walk that chain (~25 node visits of ~160 per warp at 1 M bodies).  Up to
// PB_BIG_MAX of them go on a side list that every querying collider tests directly; in the tree their leaf box is empty.  Which
// colliders are listed does not change the pair set.
__global__ void k_morton(int n, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const int* __restrict__ sb,
                         unsigned int* __restrict__ keys, int* __restrict__ ids, int* __restrict__ colFlags, int* __restrict__ bigList,
                         const int* __restrict__ colRow, const int* __restrict__ rowEntity, int4* __restrict__ colInfo, int isotropic) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 lo = mk3(orderedToFloat(sb[0]), orderedToFloat(sb[1]), orderedToFloat(sb[2]));
    V3 hi = mk3(orderedToFloat(sb[3]), orderedToFloat(sb[4]), orderedToFloat(sb[5]));
    V3 c = (mk3(aabbMin[i]) + mk3(aabbMax[i])) * 0.5f;
    V3 ext = hi - lo;
    {
        const int old = colFlags[i];
        int f = old & ~COLF_BIG;
        if (bigList && (f & (COLF_ENABLE | COLF_DYNAMIC)) == COLF_ENABLE) {
            V3 e = mk3(aabbMax[i]) - mk3(aabbMin[i]);
            if (fmaxf(e.x, fmaxf(e.y, e.z)) > 0.125f * fmaxf(ext.x, fmaxf(ext.y, ext.z))) {
                int slot = atomicAdd(&bigList[0], 1);
                if (slot < PB_BIG_MAX) { bigList[1 + slot] = i; f |= COLF_BIG; }
            }
        }
        if (f != old) colFlags[i] = f;
        // everything the pair walk needs to know about a collider in ONE 16-byte word (flags, body row, entity): the walk's leaf test
        // was three dependent loads (flags -> row -> entity of the row)
        const int row = colRow[i];
        colInfo[i] = make_int4(f, row, rowEntity[row], 0);
    }
    float sx = ext.x > 0.f ? 1024.f / ext.x : 0.f, sy = ext.y > 0.f ? 1024.f / ext.y : 0.f, sz = ext.z > 0.f ? 1024.f / ext.z : 0.f;
    // Cubic cells: one scale for the three axes.  A scene that is wide and flat (a million bodies on a 1 km terrain, 6 m of height)
    // otherwise spends a third of its key bits on millimetres of height, and leaves that are neighbours in the order lie scattered
    // along a height contour of a 30 m cell instead of side by side -- the packet walks live on that neighbourhood.
    if (isotropic) { const float m = fmaxf(ext.x, fmaxf(ext.y, ext.z)); sx = sy = sz = m > 0.f ? 1024.f / m : 0.f; }
    unsigned int x = (unsigned int)fminf(fmaxf((c.x - lo.x) * sx, 0.f), 1023.f);
    unsigned int y = (unsigned int)fminf(fmaxf((c.y - lo.y) * sy, 0.f), 1023.f);
    unsigned int z = (unsigned int)fminf(fmaxf((c.z - lo.z) * sz, 0.f), 1023.f);
    keys[i] = (expandBits(x) << 2) | (expandBits(y) << 1) | expandBits(z);
    ids[i] = i;
}

// ---- Karras LBVH ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int deltaKey(const unsigned int* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned int a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((unsigned int)i ^ (unsigned int)j);
    return __clz(a ^ b);
}

// child encoding: >= 0 internal node index, < 0 leaf: ~(sorted leaf position)
__global__ void k_lbvh_build(int n, const unsigned int* __restrict__ keys, int* __restrict__ left, int* __restrict__ right,
                             int* __restrict__ parent, int* __restrict__ leafParent, int* __restrict__ flag, int2* __restrict__ nodeRange) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    flag[i] = 0;
    int d = (deltaKey(keys, n, i, i + 1) - deltaKey(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = deltaKey(keys, n, i, i - d);
    int lmax = 2;
    while (deltaKey(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (deltaKey(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = deltaKey(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (deltaKey(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc;
    nodeRange[i] = make_int2(gamma, hi);     // last sorted-leaf position under the left child / under the node
    if (lc >= 0) parent[lc] = i; else leafParent[~lc] = i;
    if (rc >= 0) parent[rc] = i; else leafParent[~rc] = i;
    if (i == 0) parent[0] = -1;
}

// node layout for traversal: nodeMin[2*i] = left child box min (w = left child), nodeMax[2*i] = left box max (w = right child)
//                            nodeMin[2*i+1] = right child box min (w = last sorted-leaf position under the LEFT child,
//                                             bit 31 = the left subtree holds a non-dynamic collider),
//                            nodeMax[2*i+1] = right box max (w = last position under the RIGHT child, bit 31 likewise)
// leaf children are re-encoded as ~colliderIndex so the traversal needs no indirection.
// The position / flag words let a query at sorted position i skip every subtree that only holds dynamic colliders at
// positions <= i: those pairs are reported by the other collider's query (each dynamic pair is found exactly once).
__global__ void k_lbvh_refit(int n, const int* __restrict__ leafId, const int* __restrict__ left, const int* __restrict__ right,
                             const int* __restrict__ parent, const int* __restrict__ leafParent, int* __restrict__ flag,
                             const int2* __restrict__ nodeRange, const int* __restrict__ colFlags,
                             const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                             float4* __restrict__ nodeMin, float4* __restrict__ nodeMax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = leafParent[i];
    while (p >= 0) {
        // the node's topology words are fetched BEFORE the fence / arrival counter: they do not depend on the sibling, and behind the
        // atomic they were a fourth dependent memory round trip per level of a ~30-level chain
        const int lc = left[p], rc = right[p], up = parent[p];
        const int2 rg = nodeRange[p];
        __threadfence();
        if (atomicAdd(&flag[p], 1) == 0) return;  // first arrival waits for the sibling subtree
        float4 lmn, lmx, rmn, rmx;
        int lenc, renc;
        unsigned int lstat, rstat;   // subtree holds a collider that never issues a query (static / kinematic / disabled)
        if (lc < 0) {
            int c = leafId[~lc]; int fc = colFlags[c];
            lmn = aabbMin[c]; lmx = aabbMax[c]; lenc = ~c; lstat = (fc & (COLF_ENABLE | COLF_DYNAMIC)) != (COLF_ENABLE | COLF_DYNAMIC);
            if (fc & COLF_BIG) { lmn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f); lmx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f); lstat = 0; }
        }
        else {
            float4 a0 = __ldcg(&nodeMin[2 * lc]), a1 = __ldcg(&nodeMax[2 * lc]), b0 = __ldcg(&nodeMin[2 * lc + 1]), b1 = __ldcg(&nodeMax[2 * lc + 1]);
            lstat = ((unsigned int)__float_as_int(b0.w) | (unsigned int)__float_as_int(b1.w)) >> 31;
            lmn = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.f);
            lmx = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.f);
            lenc = lc;
        }
        if (rc < 0) {
            int c = leafId[~rc]; int fc = colFlags[c];
            rmn = aabbMin[c]; rmx = aabbMax[c]; renc = ~c; rstat = (fc & (COLF_ENABLE | COLF_DYNAMIC)) != (COLF_ENABLE | COLF_DYNAMIC);
            if (fc & COLF_BIG) { rmn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f); rmx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f); rstat = 0; }
        }
        else {
            float4 a0 = __ldcg(&nodeMin[2 * rc]), a1 = __ldcg(&nodeMax[2 * rc]), b0 = __ldcg(&nodeMin[2 * rc + 1]), b1 = __ldcg(&nodeMax[2 * rc + 1]);
            rstat = ((unsigned int)__float_as_int(b0.w) | (unsigned int)__float_as_int(b1.w)) >> 31;
            rmn = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.f);
            rmx = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.f);
            renc = rc;
        }
        lmn.w = __int_as_float(lenc); lmx.w = __int_as_float(renc);
        rmn.w = __int_as_float((int)((unsigned int)rg.x | (lstat << 31)));
        rmx.w = __int_as_float((int)((unsigned int)rg.y | (rstat << 31)));
        nodeMin[2 * p] = lmn; nodeMax[2 * p] = lmx; nodeMin[2 * p + 1] = rmn; nodeMax[2 * p + 1] = rmx;
        p = up;
    }
}

// closed-interval overlap == !(a.max < b.min || a.min > b.max) per axis (Physecs.cpp:152-154)
__device__ __forceinline__ bool overlaps(V3 amn, V3 amx, float4 bmn, float4 bmx) {
    return !(amx.x < bmn.x || amn.x > bmx.x) && !(amx.y < bmn.y || amn.y > bmx.y) && !(amx.z < bmn.z || amn.z > bmx.z);
}

// Append one pair per calling lane: the lanes that arrive together (__activemask) share one atomic.  (The cooperative-groups
// coalesced_threads() version of this cost ~40 instructions per call -- a quarter of the whole walk at ~66 emissions per warp.)
__device__ __forceinline__ void emitPair(int a, int b, const int* __restrict__ colRow, const int* __restrict__ rowEntity,
                                         int2* __restrict__ pairs, int* __restrict__ counters, int maxPairs) {
    const unsigned int m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[CNT_PAIRS], __popc(m));
    base = __shfl_sync(m, base, leader);
    int slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < maxPairs) {
        unsigned int ea = (unsigned int)rowEntity[colRow[a]], eb = (unsigned int)rowEntity[colRow[b]];
        pairs[slot] = (ea < eb) ? make_int2(a, b) : make_int2(b, a);   // lower entity id first (Physecs.cpp:158-168)
    } else {
        atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_PAIRS);
    }
}

// emitPair with the entity ids already at hand (k_lbvh_pairs reads them from colInfo)
__device__ __forceinline__ void emitPairEnt(int a, int b, unsigned int ea, unsigned int eb, int2* __restrict__ pairs, int* __restrict__ counters, int maxPairs) {
    const unsigned int m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[CNT_PAIRS], __popc(m));
    base = __shfl_sync(m, base, leader);
    int slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < maxPairs) pairs[slot] = (ea < eb) ? make_int2(a, b) : make_int2(b, a);   // lower entity id first (Physecs.cpp:158-168)
    else { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_PAIRS); }
}

// One WARP per 32 consecutive sorted leaves, walking the tree as a packet: the warp keeps one stack (shared memory), pops one
// node at a time, every lane tests its own box against the two child boxes, and a child is descended when ANY lane overlaps it.
// Morton-sorted neighbours take nearly the same path, so the union of the 32 paths is a small multiple of one path, while every
// node is fetched once per warp from one address (a broadcast: 4 L1 wavefronts) instead of 32 lanes fetching 32 different
// 64-byte nodes (128 wavefronts per step of a per-thread walk -- what bounded the per-thread version at 0.8 ms for 1 M leaves).
// Only colliders of non-kinematic dynamic bodies issue queries (a pair needs one, Physecs.cpp:147), which keeps huge static
// boxes / terrain from walking the whole tree; the other lanes carry an empty box.
#define PAIRS_WARPS 4
__global__ void __launch_bounds__(32 * PAIRS_WARPS) k_lbvh_pairs(int n, const int* __restrict__ leafId, const int4* __restrict__ colInfo,
                             const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                             const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax,
                             int2* __restrict__ pairs, int* __restrict__ counters, int maxPairs, const int* __restrict__ bigList) {
    __shared__ int sstack[PAIRS_WARPS][64];
    const unsigned FULL = 0xffffffffu;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int a = -1, rowA = -1;
    unsigned int entA = 0;
    bool active = false;
    V3 amn = mk3(FLT_MAX), amx = mk3(-FLT_MAX);     // empty box: overlaps nothing
    if (i < n) {
        a = leafId[i];
        const int4 ia = colInfo[a];
        active = (ia.x & (COLF_ENABLE | COLF_DYNAMIC)) == (COLF_ENABLE | COLF_DYNAMIC);
        if (active) { amn = mk3(aabbMin[a]); amx = mk3(aabbMax[a]); rowA = ia.y; entA = (unsigned int)ia.z; }
    }
    if (!__any_sync(FULL, active)) return;
    if (bigList) {
        // the big statics kept out of the tree (k_morton): enabled, never querying, so only the entity test is left of the leaf checks
        const int nb = min(bigList[0], PB_BIG_MAX);
        for (int k = 0; k < nb; ++k) {
            const int b = bigList[1 + k];
            const int4 ib = colInfo[b];
            if (active && ib.y != rowA && overlaps(amn, amx, aabbMin[b], aabbMax[b])) emitPairEnt(a, b, entA, (unsigned int)ib.z, pairs, counters, maxPairs);
        }
    }
    int* st = sstack[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    int sp = 0;                        // warp-uniform
    if (lane == 0) st[0] = 0;
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        int node = st[--sp];
        __syncwarp();                  // everyone has read the slot before lane 0 may overwrite it
        float4 lmn = nodeMin[2 * node], lmx = nodeMax[2 * node], rmn = nodeMin[2 * node + 1], rmx = nodeMax[2 * node + 1];
        int lc = __float_as_int(lmn.w), rc = __float_as_int(lmx.w);
        unsigned int lw = (unsigned int)__float_as_int(rmn.w), rw = (unsigned int)__float_as_int(rmx.w);
        int llast = (int)(lw & 0x7fffffffu), rlast = (int)(rw & 0x7fffffffu);
        // a subtree of dynamic colliders that all sort at or before this query is the other side's job
        bool ol = active && ((lw >> 31) || llast > i) && overlaps(amn, amx, lmn, lmx);
        bool orr = active && ((rw >> 31) || rlast > i) && overlaps(amn, amx, rmn, rmx);
        unsigned int bl = __ballot_sync(FULL, ol), br = __ballot_sync(FULL, orr);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            unsigned int any = side ? br : bl;
            int c = side ? rc : lc;
            if (!any) continue;                          // warp-uniform
            if (c >= 0) {                                // warp-uniform: the link is the same word for every lane
                // 64 slots cannot run out: a root-to-leaf path has at most 62 internal nodes (deltaKey takes 62 distinct values: 2..31 on
                // 30-bit keys, 32..63 on the index tie-break) and a depth-first walk holds one pending sibling per level.  Should a
                // future key layout break that, the step fails loudly instead of dropping the subtree.
                if (sp < 64) { if (lane == 0) st[sp] = c; ++sp; }
                else if (lane == 0) { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_WALK_STACK); }
                continue;
            }
            if (!(side ? orr : ol)) continue;
            int b = ~c;
            if (b == a) continue;
            const int4 ib = colInfo[b];
            if (!(ib.x & COLF_ENABLE)) continue;
            // a leaf child sits at position llast (left) / llast + 1 (right): dynamic-dynamic pairs are emitted by the earlier one
            if ((ib.x & COLF_DYNAMIC) && (side ? llast + 1 : llast) < i) continue;
            if (ib.y == rowA) continue;                     // same entity (Physecs.cpp:145)
            emitPairEnt(a, b, entA, (unsigned int)ib.z, pairs, counters, maxPairs);
        }
        __syncwarp();
    }
}

// Small scenes (n <= PB_BRUTE_FORCE_MAX colliders): every dynamic collider against every collider, tile by tile through shared
// memory.  Same predicates, same pair set; it replaces a Morton sort (4 radix passes), tree build, refit and walk whose ~25 tiny
// launches and dependent-latency chains cost ~0.2 ms however small the scene is (6 k colliders: 19 M box tests, a few microseconds).
// The tree is still built on demand for the scene queries (queries.cu).
#define PB_BRUTE_FORCE_MAX 8192
#define BF_TILE 128
// Tile culling: colliders are stored in creation order, and creation order is spatially coherent in the scenes this path exists for
// (a batch of little scenes laid out side by side: 128 consecutive colliders are ~10 neighbouring scenes).  k_tile_bounds keeps the
// union box of every tile of 128 colliders (enabled ones only); a CTA skips, as a whole, every tile whose box misses the union box of
// its own 128 queries, and a thread skips a tile its own box misses.  Pure culling of tests that cannot succeed: same pair set.
__global__ void __launch_bounds__(BF_TILE) k_tile_bounds(int n, const int* __restrict__ colFlags, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                                        float4* __restrict__ tileMin, float4* __restrict__ tileMax) {
    __shared__ float red[BF_TILE / 32][6];
    const int i = blockIdx.x * BF_TILE + threadIdx.x;
    V3 mn = mk3(FLT_MAX), mx = mk3(-FLT_MAX);
    if (i < n && (colFlags[i] & COLF_ENABLE)) { mn = mk3(aabbMin[i]); mx = mk3(aabbMax[i]); }
    for (int d = 16; d > 0; d >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, d)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, d)); mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, d));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, d)); mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, d)); mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, d));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[w][0] = mn.x; red[w][1] = mn.y; red[w][2] = mn.z; red[w][3] = mx.x; red[w][4] = mx.y; red[w][5] = mx.z; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < BF_TILE / 32; ++k) {
            mn = vmin(mn, mk3(red[k][0], red[k][1], red[k][2])); mx = vmax(mx, mk3(red[k][3], red[k][4], red[k][5]));
        }
        tileMin[blockIdx.x] = f4(mn); tileMax[blockIdx.x] = f4(mx);
    }
}

// 32 queries per CTA, four threads per query: thread (q = lane, part = warp) tests query q against candidates [32 part, 32 part + 32) of
// every staged tile.  (128 queries per CTA with one thread walking all 128 candidates left each warp a stream of thousands of dependent
// instructions at ~20 cycles apiece and ~11 warps per SM: ncu showed the kernel bound by exactly that, whatever the culling did.)
#define BF_QUERIES 32
__global__ void __launch_bounds__(BF_TILE) k_pairs_bruteforce(int n, int tilesPerSlice, const int* __restrict__ colFlags, const int* __restrict__ colRow,
                                                             const int* __restrict__ rowEntity, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                                             const float4* __restrict__ tileMin, const float4* __restrict__ tileMax,
                                                             int2* __restrict__ pairs, int* __restrict__ counters, int maxPairs) {
    __shared__ float4 smn[BF_TILE], smx[BF_TILE];
    __shared__ int sflag[BF_TILE], srow[BF_TILE];
    __shared__ unsigned int sHit[BF_TILE / 32];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int a = blockIdx.x * BF_QUERIES + lane;
    bool active = false;
    V3 amn = mk3(FLT_MAX), amx = mk3(-FLT_MAX);       // empty box: meets nothing
    int rowA = -1;
    if (a < n) {
        active = (colFlags[a] & (COLF_ENABLE | COLF_DYNAMIC)) == (COLF_ENABLE | COLF_DYNAMIC);
        if (active) { amn = mk3(aabbMin[a]); amx = mk3(aabbMax[a]); rowA = colRow[a]; }
    }
    // union box of the CTA's 32 queries (every warp holds the same 32: a warp reduction is the CTA's)
    V3 qmn = amn, qmx = amx;
    for (int d = 16; d > 0; d >>= 1) {
        qmn.x = fminf(qmn.x, __shfl_xor_sync(0xffffffffu, qmn.x, d)); qmn.y = fminf(qmn.y, __shfl_xor_sync(0xffffffffu, qmn.y, d)); qmn.z = fminf(qmn.z, __shfl_xor_sync(0xffffffffu, qmn.z, d));
        qmx.x = fmaxf(qmx.x, __shfl_xor_sync(0xffffffffu, qmx.x, d)); qmx.y = fmaxf(qmx.y, __shfl_xor_sync(0xffffffffu, qmx.y, d)); qmx.z = fmaxf(qmx.z, __shfl_xor_sync(0xffffffffu, qmx.z, d));
    }
    if (qmn.x > qmx.x) return;                 // no querying collider in this CTA (uniform)
    // Tiles (128 consecutive colliders) in reach: 128 tile boxes at a time, one per thread, against the union box -- the boxes of a
    // batch of thousands of scenes cost a few coalesced loads instead of one dependent round trip per tile; the CTA then walks the hits
    // in tile order, and a query skips a tile its own box misses.
    const int firstTile = blockIdx.y * tilesPerSlice;
    const int nTiles = (n + BF_TILE - 1) / BF_TILE;
    const int lastTile = min(firstTile + tilesPerSlice, nTiles);
    int met = 0;
    for (int chunk = firstTile; chunk < lastTile; chunk += BF_TILE) {
        const int tt = chunk + threadIdx.x;
        const bool reach = tt < lastTile && overlaps(qmn, qmx, tileMin[tt], tileMax[tt]);
        const unsigned int bal = __ballot_sync(0xffffffffu, reach);
        if (lane == 0) sHit[part] = bal;
        __syncthreads();
        unsigned int masks[BF_TILE / 32];
        for (int w = 0; w < BF_TILE / 32; ++w) masks[w] = sHit[w];
        __syncthreads();
#pragma unroll
        for (int w = 0; w < BF_TILE / 32; ++w) {
            unsigned int tm = masks[w];
            while (tm) {                            // uniform across the CTA
                const int t = chunk + 32 * w + __ffs(tm) - 1;
                tm &= tm - 1;
                ++met;
                const int base = t * BF_TILE;
                const float4 tmn = tileMin[t], tmx = tileMax[t];
                const int b0 = base + threadIdx.x;
                if (b0 < n) { smn[threadIdx.x] = aabbMin[b0]; smx[threadIdx.x] = aabbMax[b0]; sflag[threadIdx.x] = colFlags[b0]; srow[threadIdx.x] = colRow[b0]; }
                else sflag[threadIdx.x] = 0;
                __syncthreads();
                // this thread's hits among its 32 candidates of the tile (branch-free, unrolled: the shared-memory loads of a run of
                // iterations go out together); the pairs are appended afterwards with ONE atomic per warp and tile
                unsigned int hit = 0;
                if (active && overlaps(amn, amx, tmn, tmx)) {
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) {
                        const int j = 32 * part + k;
                        const int fb = sflag[j];
                        // enabled; a pair of two querying colliders is emitted by the lower index; not the same entity (Physecs.cpp:145)
                        const bool ok = (fb & COLF_ENABLE) && !((fb & COLF_DYNAMIC) && base + j <= a) && srow[j] != rowA && overlaps(amn, amx, smn[j], smx[j]);
                        hit |= (ok ? 1u : 0u) << k;
                    }
                }
                {
                    const int mine = __popc(hit);
                    int inc = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int t_ = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t_; }
                    const int total = __shfl_sync(0xffffffffu, inc, 31);
                    if (total) {                                          // warp-uniform
                        int wbase = 0;
                        if (lane == 31) wbase = atomicAdd(&counters[CNT_PAIRS], total);
                        wbase = __shfl_sync(0xffffffffu, wbase, 31);
                        int slot = wbase + inc - mine;
                        if (mine) {
                            const unsigned int ea = (unsigned int)rowEntity[rowA];
                            unsigned int m = hit;
                            while (m) {
                                const int j = 32 * part + __ffs(m) - 1;
                                m &= m - 1;
                                const int b = base + j;
                                if (slot < maxPairs) {
                                    const unsigned int eb = (unsigned int)rowEntity[srow[j]];
                                    pairs[slot] = (ea < eb) ? make_int2(a, b) : make_int2(b, a);   // lower entity id first (Physecs.cpp:158-168)
                                } else { atomicOr(&counters[CNT_STATUS], PB_ECAPACITY); atomicOr(&counters[CNT_CAUSE], PB_CAUSE_PAIRS); }
                                ++slot;
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0 && met) atomicAdd(&counters[CNT_TILE_HITS], met);
}

// How many (group of 32 queries, tile) pairs meet -- what the all-pairs kernel would have to walk, counted the way it counts: the union
// box of a group's querying colliders against every tile box.  Run beside the tree broadphase for mid-sized scenes, so the host can
// tell (one step late) a batch of scenes laid out side by side -- a handful of tiles in reach of every group: all-pairs wins -- from a
// pile, where every group meets most tiles.
__global__ void __launch_bounds__(BF_TILE) k_tile_probe(int n, int nTiles, const int* __restrict__ colFlags, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                                                       const float4* __restrict__ tileMin, const float4* __restrict__ tileMax, int* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * BF_TILE + threadIdx.x;
    V3 qmn = mk3(FLT_MAX), qmx = mk3(-FLT_MAX);
    if (a < n && (colFlags[a] & (COLF_ENABLE | COLF_DYNAMIC)) == (COLF_ENABLE | COLF_DYNAMIC)) { qmn = mk3(aabbMin[a]); qmx = mk3(aabbMax[a]); }
    for (int d = 16; d > 0; d >>= 1) {
        qmn.x = fminf(qmn.x, __shfl_xor_sync(0xffffffffu, qmn.x, d)); qmn.y = fminf(qmn.y, __shfl_xor_sync(0xffffffffu, qmn.y, d)); qmn.z = fminf(qmn.z, __shfl_xor_sync(0xffffffffu, qmn.z, d));
        qmx.x = fmaxf(qmx.x, __shfl_xor_sync(0xffffffffu, qmx.x, d)); qmx.y = fmaxf(qmx.y, __shfl_xor_sync(0xffffffffu, qmx.y, d)); qmx.z = fmaxf(qmx.z, __shfl_xor_sync(0xffffffffu, qmx.z, d));
    }
    if (qmn.x > qmx.x) return;                 // no querying collider in this group (warp-uniform)
    int met = 0;
    for (int t = lane; t < nTiles; t += 32) if (overlaps(qmn, qmx, tileMin[t], tileMax[t])) ++met;
    for (int d = 16; d > 0; d >>= 1) met += __shfl_xor_sync(0xffffffffu, met, d);
    if (lane == 0 && met) atomicAdd(&counters[CNT_TILE_HITS], met);
}

// Morton sort + LBVH build + refit over the current collider bounds (n >= 2).  Shared by the step's pair search and the
// scene queries (queries.cu); ctx->treeLeafIds is the sorted leaf -> collider table of the tree just built.
int pb_build_tree(pb_ctx* ctx, bool forStep) {
    int n = ctx->nCol;
    if (n < 2) return PB_OK;
    int* sb = (int*)ctx->sceneBounds;
    int* bigList = (forStep && ctx->bigListMode) ? ctx->bigList : nullptr;     // scene queries walk a tree that holds every collider
    ++ctx->launches, k_scene_bounds_init<<<1, 32, 0, ctx->stream>>>(sb, bigList);
    int blocks = pb_grid(n, 256); if (blocks > ctx->numSMs * 8) blocks = ctx->numSMs * 8;
    ++ctx->launches, k_scene_bounds<<<blocks, 256, 0, ctx->stream>>>(n, ctx->aabbMin, ctx->aabbMax, sb);
    ++ctx->launches, k_morton<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ctx->aabbMin, ctx->aabbMax, sb, ctx->mortonA, ctx->leafIdA, ctx->colFlags, bigList,
                                                                       ctx->colRow, ctx->rowEntity, ctx->colInfo, ctx->mortonIso);
    bool inA = true;
    int rc = pb_radix_sort_pairs(ctx, ctx->mortonA, ctx->leafIdA, ctx->mortonB, ctx->leafIdB, n, 30, ctx->radixHist, ctx->radixTiles, &inA);
    if (rc) return rc;
    unsigned int* keys = inA ? ctx->mortonA : ctx->mortonB;
    int* ids = inA ? ctx->leafIdA : ctx->leafIdB;
    ++ctx->launches, k_lbvh_build<<<pb_grid(n - 1, 256), 256, 0, ctx->stream>>>(n, keys, ctx->nodeLeft, ctx->nodeRight, ctx->nodeParent, ctx->leafParent, ctx->nodeFlag, ctx->nodeRange);
    ++ctx->launches, k_lbvh_refit<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ids, ctx->nodeLeft, ctx->nodeRight, ctx->nodeParent, ctx->leafParent, ctx->nodeFlag,
                                                            ctx->nodeRange, ctx->colFlags, ctx->aabbMin, ctx->aabbMax, ctx->nodeMin, ctx->nodeMax);
    ctx->treeLeafIds = ids;
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// n == 2..: general path.  n < 2: no pairs.
int pb_broadphase(pb_ctx* ctx) {
    int n = ctx->nCol;
    const bool wasBrute = ctx->stepBrute;
    ctx->stepBrute = false; ctx->pendingTiles = 0;
    if (n < 2) return PB_OK;
    const int tiles = (n + BF_TILE - 1) / BF_TILE;
    // Mid-sized scenes: all pairs when creation order is spatially coherent -- a group of 32 consecutive colliders meets a handful of
    // tiles (a batch of scenes side by side; the statistics are the previous step's) --, the tree otherwise.  The tree's ~25 dependent launches cost ~0.2 ms however
    // small the scene; the all-pairs kernel of 4096 ragdoll scenes walks ~3 tiles per tile.
    const bool midSized = n > ctx->bruteForceMax && n <= ctx->bruteForceBigMax;
    const bool coherent = midSized && ctx->lastTileHits >= 0 && ctx->lastTiles == tiles && (long long)ctx->lastTileHits <= (wasBrute ? 32ll : 24ll) * tiles;     // <= 6 tiles in reach of a group of 32 queries, on average (8 to stay: no flapping at the threshold)
    if (n <= ctx->bruteForceMax || coherent) {
        const int groups = (n + BF_QUERIES - 1) / BF_QUERIES;     // CTAs of 32 queries
        int slices = (4 * ctx->numSMs + groups - 1) / groups;      // enough CTAs to fill the device a few times over
        if (slices > tiles) slices = tiles;
        if (slices < 1) slices = 1;
        const int tilesPerSlice = (tiles + slices - 1) / slices;
        // tile boxes live in the (otherwise idle) tree node arrays: one entry of nodeMin / nodeMax per tile
        ++ctx->launches, k_tile_bounds<<<tiles, BF_TILE, 0, ctx->stream>>>(n, ctx->colFlags, ctx->aabbMin, ctx->aabbMax, ctx->nodeMin, ctx->nodeMax);
        ++ctx->launches, k_pairs_bruteforce<<<dim3(groups, slices), BF_TILE, 0, ctx->stream>>>(n, tilesPerSlice, ctx->colFlags, ctx->colRow, ctx->rowEntity, ctx->aabbMin, ctx->aabbMax,
                                                                                          ctx->nodeMin, ctx->nodeMax, (int2*)ctx->pairs, ctx->counters, ctx->caps.max_pairs);
        ctx->stepBrute = true; ctx->pendingTiles = tiles;
        PB_CUDA(ctx, cudaGetLastError());
        return PB_OK;
    }
    if (midSized) {
        // the tile statistics for the next step's choice (tile boxes in the tail of the pair arena: the tree owns the node arrays)
        float4* tb = (float4*)ctx->pairs;
        if ((size_t)ctx->caps.max_pairs * sizeof(int2) >= 2 * (size_t)tiles * sizeof(float4)) {
            ++ctx->launches, k_tile_bounds<<<tiles, BF_TILE, 0, ctx->stream>>>(n, ctx->colFlags, ctx->aabbMin, ctx->aabbMax, tb, tb + tiles);
            ++ctx->launches, k_tile_probe<<<tiles, BF_TILE, 0, ctx->stream>>>(n, tiles, ctx->colFlags, ctx->aabbMin, ctx->aabbMax, tb, tb + tiles, ctx->counters);
            ctx->pendingTiles = tiles;
        }
    }
    int rc = pb_build_tree(ctx, true);
    if (rc) return rc;
    const int* ids = ctx->treeLeafIds;
    ++ctx->launches, k_lbvh_pairs<<<pb_grid(n, 32 * PAIRS_WARPS), 32 * PAIRS_WARPS, 0, ctx->stream>>>(n, ids, ctx->colInfo, ctx->aabbMin, ctx->aabbMax,
                                                            ctx->nodeMin, ctx->nodeMax, (int2*)ctx->pairs, ctx->counters, ctx->caps.max_pairs,
                                                            ctx->bigListMode ? ctx->bigList : nullptr);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

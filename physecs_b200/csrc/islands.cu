// Simulation islands: connected components of the body / constraint graph, so that small islands (a ragdoll, a separate
// little pile) are solved entirely inside ONE CTA of the persistent substep kernel, colour after colour with CTA barriers,
// instead of paying a device-wide barrier per colour (solver.cu, "local mode").
//
// The reference solves every contact, then every joint colour, in one global order (src/Physecs.cpp:482-490).  Constraints of
// different islands share no dynamic body, so running the islands' sub-sequences concurrently is the same arithmetic as any
// interleaving of them -- in particular as the (group, colour, slot) order the parity tap reports and the oracle is fed.
//
//   k_island_init / k_island_hook_*   lock-free union-find over the solver bodies (hook the larger root under the smaller)
//   k_island_compress                 root of every body
//   k_island_count_*                  constraints per island (an overflow-colour joint marks its island as not local)
//   k_island_group                    island -> group: local islands are spread over G CTAs by the position of their root
//                                     body (neighbouring bodies -> neighbouring groups), large islands go to group G = "global"
#include "pb_ctx.h"
#include "joints.cuh"
#include <algorithm>

bool pb_joint_view(pb_ctx* ctx, JointDev* out);

__device__ __forceinline__ int islandFind(int* parent, int x) {
    while (true) {
        int p = ((volatile int*)parent)[x];
        if (p == x) return x;
        int gp = ((volatile int*)parent)[p];
        if (gp != p) ((volatile int*)parent)[x] = gp;      // path halving; a benign race: any ancestor is a valid parent
        x = p;
    }
}

__device__ __forceinline__ void islandUnion(int* parent, int a, int b) {
    while (true) {
        a = islandFind(parent, a); b = islandFind(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }            // a = larger root, hooked under the smaller one
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

__global__ void k_island_init(int n, int* __restrict__ parent, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; cnt[i] = 0; }
}

__device__ __forceinline__ int islandSolverIndex(int row, int nDyn, const int* __restrict__ kinematic) {
    return (row < nDyn && !kinematic[row]) ? row : -1;
}

__global__ void k_island_hook_contacts(const int* __restrict__ counters, int maxManifolds, const int4* __restrict__ mKey, const int* __restrict__ colRow,
                                       int nDyn, const int* __restrict__ kinematic, int* parent) {
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 key = mKey[i];
        if (key.w <= 0) continue;
        int b0 = islandSolverIndex(colRow[key.x], nDyn, kinematic), b1 = islandSolverIndex(colRow[key.y], nDyn, kinematic);
        if (b0 >= 0 && b1 >= 0 && b0 != b1) islandUnion(parent, b0, b1);
    }
}

__global__ void k_island_hook_joints(int nJ, const int2* __restrict__ bodies, int* parent) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nJ) return;
    int2 b = bodies[j];
    if (b.x >= 0 && b.y >= 0 && b.x != b.y) islandUnion(parent, b.x, b.y);
}
// both hooks in one launch (grid-stride over manifolds, then over joints): a step of a small scene is a chain of few-microsecond launches
__global__ void k_island_hook(const int* __restrict__ counters, int maxManifolds, const int4* __restrict__ mKey, const int* __restrict__ colRow,
                              int nDyn, const int* __restrict__ kinematic, int nJ, const int2* __restrict__ bodies, int* parent) {
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 key = mKey[i];
        if (key.w <= 0) continue;
        int b0 = islandSolverIndex(colRow[key.x], nDyn, kinematic), b1 = islandSolverIndex(colRow[key.y], nDyn, kinematic);
        if (b0 >= 0 && b1 >= 0 && b0 != b1) islandUnion(parent, b0, b1);
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nJ; j += gridDim.x * blockDim.x) {
        int2 b = bodies[j];
        if (b.x >= 0 && b.y >= 0 && b.x != b.y) islandUnion(parent, b.x, b.y);
    }
}
__global__ void k_island_count(const int* __restrict__ counters, int maxManifolds, const int4* __restrict__ mKey, const int* __restrict__ colRow,
                               int nDyn, const int* __restrict__ kinematic, int nJ, const int2* __restrict__ bodies, int overflowStart, int localMax,
                               const int* __restrict__ root, int* __restrict__ cnt) {
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 key = mKey[i];
        if (key.w <= 0) continue;
        int b = islandSolverIndex(colRow[key.x], nDyn, kinematic);
        if (b < 0) b = islandSolverIndex(colRow[key.y], nDyn, kinematic);
        if (b >= 0) atomicAdd(&cnt[root[b]], 1);
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nJ; j += gridDim.x * blockDim.x) {
        int2 bb = bodies[j];
        int b = bb.x >= 0 ? bb.x : bb.y;
        if (b >= 0) atomicAdd(&cnt[root[b]], j >= overflowStart ? localMax + 1 : 1);     // overflow-bucket joints keep their island in the global sweep
    }
}
// group of every body + the island statistics + (optionally) the per-group body histogram, one launch
__global__ void __launch_bounds__(256) k_island_group_stats(int n, int* rootThenGroup, const int* __restrict__ cnt, int G, int localMax, int* __restrict__ stats,
                                                            int* __restrict__ bodyHist) {
    extern __shared__ int sh[];          // [G + 1] body histogram (bodyHist != nullptr)
    __shared__ int red[2][8];
    if (bodyHist) { for (int i = threadIdx.x; i <= G; i += blockDim.x) sh[i] = 0; __syncthreads(); }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0, loc = 0;
    if (i < n) {
        int r = rootThenGroup[i];
        int cr = cnt[r];
        int g = cr <= localMax ? (int)min((long long)G - 1, (long long)r * G / n) : G;
        rootThenGroup[i] = g;
        if (bodyHist) atomicAdd(&sh[g], 1);
        c = min(cnt[i], 1 << 20);            // cnt[] is non-zero exactly at the roots
        loc = (c > 0 && g != G) ? c : 0;     // (a root's group is its island's group)
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, d); loc += __shfl_xor_sync(0xffffffffu, loc, d); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = c; red[1][threadIdx.x >> 5] = loc; }
    __syncthreads();
    if (threadIdx.x < 2) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        if (t) atomicAdd(&stats[1 - threadIdx.x], t);
    }
    if (bodyHist) for (int k = threadIdx.x; k <= G; k += blockDim.x) if (sh[k]) atomicAdd(&bodyHist[k], sh[k]);
}

// roots go to a separate array: a concurrent path-halving write of another thread may still land on parent[i] after this one
__global__ void k_island_compress(int n, int* parent, int* __restrict__ root) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) root[i] = islandFind(parent, i);
}

__global__ void k_island_count_contacts(const int* __restrict__ counters, int maxManifolds, const int4* __restrict__ mKey, const int* __restrict__ colRow,
                                        int nDyn, const int* __restrict__ kinematic, const int* __restrict__ root, int* __restrict__ cnt) {
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 key = mKey[i];
        if (key.w <= 0) continue;
        int b = islandSolverIndex(colRow[key.x], nDyn, kinematic);
        if (b < 0) b = islandSolverIndex(colRow[key.y], nDyn, kinematic);
        if (b >= 0) atomicAdd(&cnt[root[b]], 1);
    }
}

__global__ void k_island_count_joints(int nJ, const int2* __restrict__ bodies, int overflowStart, int localMax, const int* __restrict__ root, int* __restrict__ cnt) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nJ) return;
    int2 bb = bodies[j];
    int b = bb.x >= 0 ? bb.x : bb.y;
    // joints of the sequential overflow bucket (scalar semantics, creation order) stay in the global sweep, with their whole island
    if (b >= 0) atomicAdd(&cnt[root[b]], j >= overflowStart ? localMax + 1 : 1);
}

// group of every body
__global__ void k_island_group(int n, int* rootThenGroup, const int* __restrict__ cnt, int G, int localMax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int* group = rootThenGroup;
    int r = rootThenGroup[i];
    int c = cnt[r];
    bool local = c <= localMax;
    group[i] = local ? (int)min((long long)G - 1, (long long)r * G / n) : G;
}

// per-group body lists: bodies counted per group (CTA-local histogram first), then placed with one atomic per distinct group per warp
__global__ void __launch_bounds__(256) k_island_body_hist(int n, const int* __restrict__ group, int G, int* __restrict__ hist) {
    extern __shared__ int sh[];
    for (int i = threadIdx.x; i <= G; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[group[i]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i <= G; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void __launch_bounds__(256) k_island_body_scatter(int n, const int* __restrict__ group, const int* __restrict__ start, int* __restrict__ fill, int* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        const bool valid = i < n;
        const int g = valid ? group[i] : -1;
        const unsigned int act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned int peers = __match_any_sync(act, g);
            const int leader = __ffs(peers) - 1;
            int slot = 0;
            if (lane == leader) slot = start[g] + atomicAdd(&fill[g], __popc(peers));
            slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1u));
            order[slot] = i;
        }
    }
}

// stats[0] = constraints of local islands, stats[1] = constraints of all islands (one atomic pair per CTA: 31 k warps adding to two
// words were most of this kernel's 45 us at 1 M bodies)
__global__ void __launch_bounds__(256) k_island_stats(int n, const int* __restrict__ group, const int* __restrict__ cnt, int G, int* __restrict__ stats) {
    __shared__ int red[2][8];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // after k_island_group the roots are gone from `group`, but cnt[] is non-zero exactly at the roots
    int c = i < n ? min(cnt[i], 1 << 20) : 0;
    int loc = (i < n && c > 0 && group[i] != G) ? c : 0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, d); loc += __shfl_xor_sync(0xffffffffu, loc, d); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = c; red[1][threadIdx.x >> 5] = loc; }
    __syncthreads();
    if (threadIdx.x < 2) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        if (t) atomicAdd(&stats[1 - threadIdx.x], t);
    }
}

// buffers of the island search; the per-group body lists exist once a scene small enough for the whole-step kernel's group-by-group form asks
int pb_islands_alloc(pb_ctx* ctx) {
    int rc;
    if (!ctx->islandParent) {
        if ((rc = pb_alloc(ctx, &ctx->islandParent, (size_t)ctx->caps.max_bodies)) || (rc = pb_alloc(ctx, &ctx->islandCount, (size_t)ctx->caps.max_bodies)) ||
            (rc = pb_alloc(ctx, &ctx->bodyGroup, (size_t)ctx->caps.max_bodies)) || (rc = pb_alloc(ctx, &ctx->islandStats, 4))) return rc;
    }
    const bool lists = ctx->nDyn <= ctx->fusedLocalMax && ctx->fusedMode != 0;
    if (lists && !ctx->bodyOrder) {
        const int G = ctx->islandGroups;
        if ((rc = pb_alloc(ctx, &ctx->bodyOrder, (size_t)ctx->caps.max_bodies)) || (rc = pb_alloc(ctx, &ctx->bodyStart, (size_t)G + 2)) || (rc = pb_alloc(ctx, &ctx->bodyCursor, (size_t)G + 2))) return rc;
    }
    return PB_OK;
}

// (islandStats, bodyStart and the body fill counters arrive zeroed: contacts.cu k_build_clear)
int pb_islands_build(pb_ctx* ctx) {
    const int n = ctx->nDyn;
    if (n <= 0) return PB_OK;
    { int rc = pb_islands_alloc(ctx); if (rc) return rc; }
    const int blocks = ctx->numSMs * 8;
    const int G = ctx->islandGroups;
    ++ctx->launches, k_island_init<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ctx->islandParent, ctx->islandCount);
    JointDev J;
    const bool joints = pb_joint_view(ctx, &J);
    const int nJ = joints ? J.n : 0;
    const int2* jb = joints ? J.bodies : nullptr;
    const int hookBlocks = pb_hint_grid(ctx->rawHint < 0 ? -1 : std::max(ctx->rawHint, nJ), 256, blocks);
    ++ctx->launches, k_island_hook<<<hookBlocks, 256, 0, ctx->stream>>>(ctx->counters, ctx->caps.max_manifolds, ctx->mKey, ctx->colRow, n, ctx->kinematic, nJ, jb, ctx->islandParent);
    ++ctx->launches, k_island_compress<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ctx->islandParent, ctx->bodyGroup);
    ++ctx->launches, k_island_count<<<hookBlocks, 256, 0, ctx->stream>>>(ctx->counters, ctx->caps.max_manifolds, ctx->mKey, ctx->colRow, n, ctx->kinematic, nJ, jb,
                                                                         ctx->jointColorStart[8], ctx->islandLocalMax, ctx->bodyGroup, ctx->islandCount);
    // body lists per group for the whole-step kernel's group-by-group form (solver.cu k_step_solve_small): only scenes small enough to take it
    const bool lists = n <= ctx->fusedLocalMax && ctx->fusedMode != 0;
    ctx->bodyListsBuilt = false;
    // bodyGroup holds the roots up to here and is converted in place (thread i reads and writes entry i only)
    ++ctx->launches, k_island_group_stats<<<pb_grid(n, 256), 256, lists ? sizeof(int) * (G + 1) : 0, ctx->stream>>>(n, ctx->bodyGroup, ctx->islandCount, G, ctx->islandLocalMax, ctx->islandStats,
                                                                                                                 lists ? ctx->bodyStart : nullptr);
    if (lists) {
        int rc = pb_exclusive_scan(ctx, ctx->bodyStart, ctx->bodyStart, G + 2, (int*)ctx->radixHist); if (rc) return rc;
        const int hb = std::min(pb_grid(n, 256), ctx->numSMs * 4);
        ++ctx->launches, k_island_body_scatter<<<hb, 256, 0, ctx->stream>>>(n, ctx->bodyGroup, ctx->bodyStart, ctx->bodyCursor, ctx->bodyOrder);
        ctx->bodyListsBuilt = true;
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

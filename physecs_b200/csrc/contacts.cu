// Contact constraint build: raw manifolds -> colour-batched, solve-ordered constraint SoA.
//
//   k_color          lock-free greedy colouring of the body/manifold graph (64-bit colour set per dynamic body).
//                    The reference solves contacts strictly sequentially (src/Physecs.cpp:484-486); within one
//                    colour no two manifolds share a dynamic body, so a parallel colour == a sequential sub-sweep
//                    and the whole step equals the reference fed the (colour, slot) order (north_star gate 3).
//   k_scatter_by_key manifolds placed in solve order (counting sort over the run table of (group, colour, single | multi) keys),
//                    exclusive scan of point counts -> point offsets
//   k_contact_build  material mix, body indices, body-local arms, restitution target with the previous step's
//                    contact cache (src/Physecs.cpp:215-315; cache semantics :237, :291-300, :313)
#include "pb_ctx.h"
#include "pb_math.cuh"
#include <utility>
#include <algorithm>

#define DISCARD_COLOR 255u

__device__ __forceinline__ int solverIndex(int row, int nDyn, const int* __restrict__ kinematic) {
    return (row < nDyn && !kinematic[row]) ? row : -1;
}

__global__ void k_color(const int4* __restrict__ mKey, int* __restrict__ counters, int maxManifolds, const int* __restrict__ colRow,
                        int nDyn, const int* __restrict__ kinematic, unsigned long long* __restrict__ colorMask,
                        unsigned int* __restrict__ sortKey) {
    __shared__ int hist[2 * PB_MAX_COLORS];     // [colour] all manifolds, [PB_MAX_COLORS + colour] the single-point ones
    if (threadIdx.x < 2 * PB_MAX_COLORS) hist[threadIdx.x] = 0;
    __syncthreads();
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 key = mKey[i];
        unsigned int color = DISCARD_COLOR;
        if (key.w > 0) {
            int b0 = solverIndex(colRow[key.x], nDyn, kinematic), b1 = solverIndex(colRow[key.y], nDyn, kinematic);
            int lo = b0, hi = b1;
            if (lo < 0 || (hi >= 0 && hi < lo)) { int t = lo; lo = hi; hi = t; }   // lo = smallest valid index first (or -1 if none)
            if (lo < 0) { lo = hi; hi = -1; }
            if (lo < 0) color = 0;
            else {
                while (true) {
                    unsigned long long m0 = *((volatile unsigned long long*)&colorMask[lo]);
                    unsigned long long m1 = hi >= 0 ? *((volatile unsigned long long*)&colorMask[hi]) : 0ull;
                    unsigned long long freeSet = ~(m0 | m1) & ~(1ull << PB_OVERFLOW_COLOR);
                    if (!freeSet) { color = PB_OVERFLOW_COLOR; break; }
                    int c = __ffsll((long long)freeSet) - 1;
                    unsigned long long bit = 1ull << c;
                    unsigned long long old = atomicOr(&colorMask[lo], bit);
                    if (old & bit) continue;
                    if (hi >= 0 && hi != lo) {
                        old = atomicOr(&colorMask[hi], bit);
                        if (old & bit) { atomicAnd(&colorMask[lo], ~bit); continue; }
                    }
                    color = (unsigned int)c;
                    break;
                }
            }
            atomicAdd(&hist[color], 1);
            if (key.w == 1) atomicAdd(&hist[PB_MAX_COLORS + color], 1);
            // inside a colour the single-point manifolds sort first: the solver gives them one thread each and the multi-point
            // ones four lanes each, so neither path waits on the other's dependent loads
            color = 2u * color + (key.w > 1 ? 1u : 0u);
        }
        sortKey[i] = color;
    }
    __syncthreads();
    if (threadIdx.x < PB_MAX_COLORS && hist[threadIdx.x]) atomicAdd(&counters[CNT_COLORSTART + threadIdx.x], hist[threadIdx.x]);
    if (threadIdx.x < PB_MAX_COLORS && hist[PB_MAX_COLORS + threadIdx.x]) atomicAdd(&counters[CNT_MULTISTART + threadIdx.x], hist[PB_MAX_COLORS + threadIdx.x]);
}

// ---- deterministic colouring (PB_DETERMINISTIC=1) -------------------------------------------------------------------------------
// k_color's claims race, so colours -- and with them the solve order and the trajectory -- differ from run to run; the reference with
// numThreads = 0 is reproducible (ThreadPool.cpp:30-46 runs everything on the caller).  This variant is a Jones-Plassmann colouring
// with fixed priorities: every manifold's priority is a 64-bit hash of its key (collider pair + triangle), a manifold is coloured in
// the round in which it holds the smallest priority among the uncoloured manifolds of BOTH its dynamic bodies, and it takes the
// lowest colour free on both.  Winners of a round share no body, so their colour sets do not race; the result is a function of the
// manifold SET alone (not of arena order, thread timing or grid shape).  One cooperative launch, two grid barriers per round;
// rounds ~ the longest priority-decreasing chain (tens).  Slower than k_color (every round walks all manifolds): a debugging mode.
__device__ __forceinline__ unsigned long long hashKey(int a, int b, int tri);
struct JpBarrier {
    unsigned int* counter; unsigned int target;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            unsigned int seen;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(seen) : "l"(counter) : "memory");
            ++seen;
            while (seen < target) { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); }
        }
        __syncthreads();
    }
};
#define JP_UNCOLOURED 0xFFFFFFFEu
__global__ void __launch_bounds__(256) k_color_jp(const int4* __restrict__ mKey, int* __restrict__ counters, int maxManifolds, const int* __restrict__ colRow,
                                                  int nDyn, const int* __restrict__ kinematic, unsigned long long* __restrict__ colorMask,
                                                  unsigned long long* __restrict__ bodyBest, unsigned int* __restrict__ sortKey,
                                                  unsigned int* __restrict__ barrier, int* __restrict__ remaining) {
    __shared__ int hist[2 * PB_MAX_COLORS];
    if (threadIdx.x < 2 * PB_MAX_COLORS) hist[threadIdx.x] = 0;
    JpBarrier bar; bar.counter = barrier; bar.target = 0;
    const int n = min(counters[CNT_RAWM], maxManifolds);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int b = tid; b < nDyn; b += nth) bodyBest[b] = ~0ull;
    for (int i = tid; i < n; i += nth) {
        int4 key = mKey[i];
        unsigned int k = DISCARD_COLOR;
        if (key.w > 0) {
            int b0 = solverIndex(colRow[key.x], nDyn, kinematic), b1 = solverIndex(colRow[key.y], nDyn, kinematic);
            if (b0 < 0 && b1 < 0) { k = key.w > 1 ? 1u : 0u; atomicAdd(&hist[0], 1); if (key.w == 1) atomicAdd(&hist[PB_MAX_COLORS], 1); }     // colour 0
            else k = JP_UNCOLOURED;
        }
        sortKey[i] = k;
    }
    if (tid == 0) { remaining[0] = 0; remaining[1] = 0; }
    bar.sync();
    for (int round = 0; ; ++round) {
        // post: every uncoloured manifold offers its priority to its bodies
        for (int i = tid; i < n; i += nth) {
            if (sortKey[i] != JP_UNCOLOURED) continue;
            int4 key = mKey[i];
            int b0 = solverIndex(colRow[key.x], nDyn, kinematic), b1 = solverIndex(colRow[key.y], nDyn, kinematic);
            unsigned long long pr = hashKey(key.x, key.y, key.z);
            if (b0 >= 0) atomicMin(&bodyBest[b0], pr);
            if (b1 >= 0 && b1 != b0) atomicMin(&bodyBest[b1], pr);
        }
        bar.sync();
        // (every CTA is past its read of the previous round's counter -- the read sits before the barrier above -- so the slot the next
        // round will count into can be cleared now)
        if (tid == 0) remaining[(round + 1) & 1] = 0;
        // decide: the holder of the smallest priority on both bodies colours itself and re-opens its bodies
        int lost = 0;
        for (int i = tid; i < n; i += nth) {
            if (sortKey[i] != JP_UNCOLOURED) continue;
            int4 key = mKey[i];
            int b0 = solverIndex(colRow[key.x], nDyn, kinematic), b1 = solverIndex(colRow[key.y], nDyn, kinematic);
            unsigned long long pr = hashKey(key.x, key.y, key.z);
            const bool win = (b0 < 0 || __ldcg(&bodyBest[b0]) == pr) && (b1 < 0 || __ldcg(&bodyBest[b1]) == pr);
            if (!win) { ++lost; continue; }
            unsigned long long m0 = b0 >= 0 ? __ldcg(&colorMask[b0]) : 0ull, m1 = b1 >= 0 ? __ldcg(&colorMask[b1]) : 0ull;
            unsigned long long freeSet = ~(m0 | m1) & ~(1ull << PB_OVERFLOW_COLOR);
            unsigned int color = PB_OVERFLOW_COLOR;
            if (freeSet) {
                color = (unsigned int)(__ffsll((long long)freeSet) - 1);
                if (b0 >= 0) __stcg(&colorMask[b0], m0 | (1ull << color));
                if (b1 >= 0 && b1 != b0) __stcg(&colorMask[b1], m1 | (1ull << color));
            }
            if (b0 >= 0) __stcg(&bodyBest[b0], ~0ull);
            if (b1 >= 0) __stcg(&bodyBest[b1], ~0ull);
            atomicAdd(&hist[color], 1);
            if (key.w == 1) atomicAdd(&hist[PB_MAX_COLORS + color], 1);
            sortKey[i] = 2u * color + (key.w > 1 ? 1u : 0u);
        }
        if (lost) atomicAdd(&remaining[round & 1], lost);
        bar.sync();
        if (__ldcg(&remaining[round & 1]) == 0) break;       // uniform over the grid
    }
    __syncthreads();
    if (threadIdx.x < PB_MAX_COLORS && hist[threadIdx.x]) atomicAdd(&counters[CNT_COLORSTART + threadIdx.x], hist[threadIdx.x]);
    if (threadIdx.x < PB_MAX_COLORS && hist[PB_MAX_COLORS + threadIdx.x]) atomicAdd(&counters[CNT_MULTISTART + threadIdx.x], hist[PB_MAX_COLORS + threadIdx.x]);
}

// deterministic mode: the sequential bucket (colour 63, manifolds that may share bodies) is solved in slot order, and the counting-sort
// scatter fills slots in arrival order -- re-order that one run by key hash.  One thread: the bucket holds the few manifolds of bodies
// with more than 63 contacts, usually none.
__global__ void k_sort_overflow_run(int nGroups, const int* __restrict__ keyStart, const int4* __restrict__ mKey, int* __restrict__ outRaw) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per group's run table (islands off: only the last one is filled)
    if (g >= nGroups) return;
    const int* keyStartG = keyStart + (size_t)g * PB_KEY_COLORS;
    const int start = keyStartG[2 * PB_OVERFLOW_COLOR], end = keyStartG[2 * PB_OVERFLOW_COLOR + 2];
    for (int i = start + 1; i < end; ++i) {
        int x = outRaw[i]; int4 kx = mKey[x]; unsigned long long hx = hashKey(kx.x, kx.y, kx.z);
        int j = i;
        while (j > start) { int4 kp = mKey[outRaw[j - 1]]; if (hashKey(kp.x, kp.y, kp.z) <= hx) break; outRaw[j] = outRaw[j - 1]; --j; }
        outRaw[j] = x;
    }
}

// colour starts (taps, counters) and, for the plain colour-major order (islands off), the run table of group G: entry c * 2 = first
// single-point slot of colour c, c * 2 + 1 = first multi-point slot, [PB_KEY_COLORS] = end
__global__ void k_color_starts(int* counters, int* __restrict__ keyStartG) {
    // one warp, two colours per lane: exclusive scan of the colour counts
    const int lane = threadIdx.x;
    int cnt[2], singles[2];
    for (int k = 0; k < 2; ++k) { cnt[k] = counters[CNT_COLORSTART + 2 * lane + k]; singles[k] = counters[CNT_MULTISTART + 2 * lane + k]; }
    int sum = cnt[0] + cnt[1], inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    int run = inc - sum;
    int top = cnt[1] > 0 ? 2 * lane + 2 : (cnt[0] > 0 ? 2 * lane + 1 : 0);      // colours in use = highest non-empty colour + 1
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, d));
    for (int k = 0; k < 2; ++k) {
        int c = 2 * lane + k;
        counters[CNT_COLORSTART + c] = run;
        counters[CNT_MULTISTART + c] = run + singles[k];
        if (keyStartG) { keyStartG[2 * c] = run; keyStartG[2 * c + 1] = run + singles[k]; }
        if (c == PB_OVERFLOW_COLOR) counters[CNT_OVERFLOW] = cnt[k];
        run += cnt[k];
    }
    if (lane == 31) {
        counters[CNT_COLORSTART + PB_MAX_COLORS] = run;
        if (keyStartG) keyStartG[PB_KEY_COLORS] = run;
        counters[CNT_MANIFOLDS] = run;
        counters[CNT_NCOLORS] = top;
    }
}

// islands on: solve-order key = group * 128 + colour * 2 + multi; histogram of the keys (scanned into the run table)
__global__ void k_island_keys(const int* __restrict__ counters, int maxManifolds, const int4* __restrict__ mKey, const int* __restrict__ colRow,
                              int nDyn, const int* __restrict__ kinematic, const int* __restrict__ bodyGroup, int G,
                              unsigned int* __restrict__ sortKey, int* __restrict__ keyHist) {
    int n = min(counters[CNT_RAWM], maxManifolds);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned int k = sortKey[i];
        if (k >= PB_KEY_COLORS) { sortKey[i] = 0xFFFFFFFFu; continue; }      // discarded (no points): sorts behind every group
        int4 key = mKey[i];
        int b = solverIndex(colRow[key.x], nDyn, kinematic);
        if (b < 0) b = solverIndex(colRow[key.y], nDyn, kinematic);
        int g = b >= 0 ? bodyGroup[b] : G;
        unsigned int full = (unsigned int)g * PB_KEY_COLORS + k;
        sortKey[i] = full;
        atomicAdd(&keyHist[full], 1);
    }
}

// Every small table a step's build zeroes, in ONE launch (they were a dozen memsets / copies of a few microseconds each, a tenth of the
// step of a 512-scene batch): up to 8 (pointer, word count) ranges.
struct ClearList { int* p[8]; int n[8]; };
__global__ void __launch_bounds__(256) k_build_clear(ClearList L) {
    for (int k = 0; k < 8; ++k) {
        int* p = L.p[k]; const int n = L.n[k];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
    }
}

// counting-sort scatter: sortKey[i] (< keyLimit; anything else is a discarded manifold) + keyBase indexes the run cursors
__global__ void __launch_bounds__(256) k_scatter_by_key(const int* __restrict__ counters, int maxManifolds, const unsigned int* __restrict__ sortKey,
                                                        unsigned int keyBase, unsigned int keyLimit, const int* __restrict__ runStart, int* __restrict__ fill,
                                                        int* __restrict__ outRaw, unsigned int* __restrict__ outKey) {
    const int n = min(counters[CNT_RAWM], maxManifolds);
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        const unsigned int k = i < n ? sortKey[i] : 0xFFFFFFFFu;
        const bool valid = k < keyLimit;
        const unsigned int act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned int peers = __match_any_sync(act, k);
            const int leader = __ffs(peers) - 1;
            int slot = 0;
            if (lane == leader) slot = runStart[keyBase + k] + atomicAdd(&fill[keyBase + k], __popc(peers));
            slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1u));
            outRaw[slot] = i;
            outKey[slot] = k;
        }
    }
}

__global__ void k_gather_np(const int* __restrict__ counters, const int* __restrict__ mSorted, const int4* __restrict__ mKey, int* __restrict__ np) {
    int n = counters[CNT_MANIFOLDS];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) np[i] = mKey[mSorted[i]].w;
}

__global__ void k_count_points(int* counters, const int* __restrict__ pointOfs, const int* __restrict__ np) {
    int n = counters[CNT_MANIFOLDS];
    counters[CNT_POINTS] = n > 0 ? pointOfs[n - 1] + np[n - 1] : 0;
}

__device__ __forceinline__ unsigned long long hashKey(int a, int b, int tri) {
    unsigned long long h = ((unsigned long long)(unsigned int)a << 32) | (unsigned int)b;
    h ^= (unsigned long long)(unsigned int)tri * 0x9E3779B97F4A7C15ull;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h | 1ull;   // 0 = empty slot
}

__global__ void __launch_bounds__(128) k_contact_build(
    int* __restrict__ counters, const int* __restrict__ mSorted, const int4* __restrict__ mKey, const float4* __restrict__ mNormal,
    const float4* __restrict__ mPts, const int* __restrict__ pointOfs, const int* __restrict__ colRow, const float4* __restrict__ colMat,
    int nDyn, const int* __restrict__ kinematic, const float4* __restrict__ pos, const float4* __restrict__ quat,
    const float4* __restrict__ vel, const float4* __restrict__ angvel, const float4* __restrict__ comInvMass,
    int4* __restrict__ cHead, int2* __restrict__ cBodies, int2* __restrict__ cRowsT, float4* __restrict__ cNormal, float4* __restrict__ cSoft, int* __restrict__ cNp,
    float4* __restrict__ pR0T, float4* __restrict__ pR1,
    // previous step (contact cache)
    const unsigned long long* __restrict__ prevTag, const int4* __restrict__ prevVal, const int* __restrict__ prevPointOfs,
    const int* __restrict__ prevNp, const float4* __restrict__ prevR0T,
    unsigned long long* __restrict__ curTag, int4* __restrict__ curVal, int cacheMask) {
    int n = counters[CNT_MANIFOLDS];
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CNT_POINTS] = n > 0 ? pointOfs[n - 1] + mKey[mSorted[n - 1]].w : 0;      // contact points of the step (taps)
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        int raw = mSorted[s];
        int4 key = mKey[raw];
        int np = key.w;
        V3 nrm = mk3(mNormal[raw]);
        int row0 = colRow[key.x], row1 = colRow[key.y];
        float4 mat0 = colMat[key.x], mat1 = colMat[key.y];
        float friction = (mat0.x + mat1.x) * 0.5f;
        float isSoft = 0.f, frequency = 0.f, damping = 0.f, restitution;
        if (mat0.z != 0.f || mat1.z != 0.f) {
            isSoft = 1.f;
            if (mat0.z != 0.f && mat1.z != 0.f) { frequency = gmin(mat0.y, mat1.y); damping = gmin(mat0.z, mat1.z); }
            else if (mat0.z != 0.f) { frequency = mat0.y; damping = mat0.z; }
            else { frequency = mat1.y; damping = mat1.z; }
            restitution = 0.f;
        } else restitution = (mat0.y + mat1.y) * 0.5f;
        int b0 = solverIndex(row0, nDyn, kinematic), b1 = solverIndex(row1, nDyn, kinematic);
        Q4 q0 = mkq(quat[row0]), q1 = mkq(quat[row1]);
        V3 com0 = mk3(0.f), v0 = mk3(0.f), w0 = mk3(0.f), com1 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
        if (b0 >= 0) { com0 = mk3(pos[row0]) + rotate(q0, mk3(comInvMass[b0])); v0 = mk3(vel[2 * b0]); w0 = mk3(angvel[2 * b0]); }
        if (b1 >= 0) { com1 = mk3(pos[row1]) + rotate(q1, mk3(comInvMass[b1])); v1 = mk3(vel[2 * b1]); w1 = mk3(angvel[2 * b1]); }
        Q4 iq0 = qinverse(q0), iq1 = qinverse(q1);
        // previous manifold of the same (pair, triangle) key
        int prevSlot = -1;
        unsigned long long tag = hashKey(key.x, key.y, key.z);
        bool useCache = restitution != 0.f;
        if (useCache && prevTag) {
            unsigned int h = (unsigned int)(tag >> 1) & cacheMask;
            while (true) {
                unsigned long long t = prevTag[h];
                if (t == 0ull) break;
                if (t == tag) { int4 v = prevVal[h]; if (v.x == key.x && v.y == key.y && v.z == key.z) { prevSlot = v.w; break; } }
                h = (h + 1) & cacheMask;
            }
        }
        int pofs = pointOfs[s];
        for (int k = 0; k < np; ++k) {
            V3 r0 = mk3(mPts[8 * (size_t)raw + 2 * k]) - com0;
            V3 r1 = mk3(mPts[8 * (size_t)raw + 2 * k + 1]) - com1;
            V3 rel = v1 + cross(w1, r1) - v0 - cross(w0, r0);
            float relN = dot(rel, nrm);
            r0 = rotate(iq0, r0);
            r1 = rotate(iq1, r1);
            float target = 0.f;
            bool found = false;
            if (prevSlot >= 0) {
                int pn = prevNp[prevSlot], po = prevPointOfs[prevSlot];
                for (int i = 0; i < pn; ++i) {
                    float4 pr = prevR0T[po + i];
                    if ((double)distance(r0, mk3(pr)) < 0.1) { found = true; target = pr.w; break; }
                }
            }
            if (!found) target = -restitution * relN;
            pR0T[pofs + k] = f4(r0, target);
            pR1[pofs + k] = f4(r1, 0.f);
        }
        cHead[s] = make_int4(b0, b1, pofs, np | (isSoft != 0.f ? 0x100 : 0));
        cBodies[s] = make_int2(b0, b1);
        cRowsT[s] = make_int2(row0, row1);
        cNormal[s] = f4(nrm, friction);
        cSoft[s] = make_float4(isSoft, frequency, damping, 0.f);
        cNp[s] = np;
        if (useCache) {
            unsigned int h = (unsigned int)(tag >> 1) & cacheMask;
            while (true) {
                unsigned long long old = atomicCAS(&curTag[h], 0ull, tag);
                if (old == 0ull) { curVal[h] = make_int4(key.x, key.y, key.z, s); break; }
                h = (h + 1) & cacheMask;
            }
        }
    }
}

// Re-key the last step's contact cache after a collider re-upload (collider indices are part of the key): entries whose
// colliders survive are re-inserted under their new indices, the rest are dropped.  Payload slots stay valid because the
// per-point arrays of the previous step are untouched by an upload.
__global__ void k_cache_remap(int size, const unsigned long long* __restrict__ oldTag, const int4* __restrict__ oldVal, const int* __restrict__ oldToNew,
                              int nOld, unsigned long long* __restrict__ newTag, int4* __restrict__ newVal) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    if (oldTag[i] == 0ull) return;
    int4 v = oldVal[i];
    if (v.x < 0 || v.x >= nOld || v.y < 0 || v.y >= nOld) return;
    int a = oldToNew[v.x], b = oldToNew[v.y];
    if (a < 0 || b < 0) return;
    unsigned long long tag = hashKey(a, b, v.z);
    unsigned int h = (unsigned int)(tag >> 1) & (unsigned int)(size - 1);
    while (true) {
        unsigned long long old = atomicCAS(&newTag[h], 0ull, tag);
        if (old == 0ull) { newVal[h] = make_int4(a, b, v.z, v.w); break; }
        h = (h + 1) & (unsigned int)(size - 1);
    }
}

// same entries, larger table (arena growth)
__global__ void k_cache_rehash(int oldSize, const unsigned long long* __restrict__ oldTag, const int4* __restrict__ oldVal,
                               int newSize, unsigned long long* __restrict__ newTag, int4* __restrict__ newVal) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= oldSize) return;
    unsigned long long tag = oldTag[i];
    if (tag == 0ull) return;
    unsigned int h = (unsigned int)(tag >> 1) & (unsigned int)(newSize - 1);
    while (true) {
        unsigned long long old = atomicCAS(&newTag[h], 0ull, tag);
        if (old == 0ull) { newVal[h] = oldVal[i]; break; }
        h = (h + 1) & (unsigned int)(newSize - 1);
    }
}
void pb_contact_cache_rehash(pb_ctx* ctx, int oldSize, const unsigned long long* oldTag, const int4* oldVal, int newSize, unsigned long long* newTag, int4* newVal) {
    cudaMemsetAsync(newTag, 0, sizeof(unsigned long long) * (size_t)newSize, ctx->stream);
    ++ctx->launches, k_cache_rehash<<<pb_grid(oldSize, 256), 256, 0, ctx->stream>>>(oldSize, oldTag, oldVal, newSize, newTag, newVal);
}

int pb_contact_cache_remap(pb_ctx* ctx, int nOld, const int* dOldToNew) {
    int cur = ctx->curBuf, other = cur ^ 1;
    cudaMemsetAsync(ctx->cacheTag[other], 0, sizeof(unsigned long long) * (size_t)ctx->cacheSize, ctx->stream);
    ++ctx->launches, k_cache_remap<<<pb_grid(ctx->cacheSize, 256), 256, 0, ctx->stream>>>(ctx->cacheSize, ctx->cacheTag[cur], ctx->cacheVal[cur], dOldToNew, nOld,
                                                                                         ctx->cacheTag[other], ctx->cacheVal[other]);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->cacheTag[0], ctx->cacheTag[1]);
    std::swap(ctx->cacheVal[0], ctx->cacheVal[1]);
    return PB_OK;
}

// Nothing here needs a count on the host: every kernel reads the manifold count from the device counters, launch shapes follow
// ctx->rawHint (the previous step's count, a guess that only shapes grids) or the arena capacity.  When the narrowphase overflowed an
// arena (CNT_STATUS), the kernels still run over the clamped counts -- everything they write is per-step scratch or the "current"
// half of a double buffer that the host flips back when it collects the step's status (capi.cu collectStep).
int pb_islands_alloc(pb_ctx* ctx);
int pb_joint_lists_alloc(pb_ctx* ctx);
int pb_contact_build(pb_ctx* ctx) {
    const int blocks = pb_hint_grid(ctx->rawHint, 256, ctx->numSMs * 8);
    int maxM = ctx->caps.max_manifolds;
    const int G = ctx->islandGroups, nKeys = (G + 1) * PB_KEY_COLORS;
    int rc;
    // one launch zeroes the colour sets, the run-fill counters and (islands on) the key histogram, the island statistics, the body and
    // joint list tables.  Islands off: the run table of group G is written whole by k_color_starts, nothing to clear.
    {
        ClearList L{};
        int k = 0;
        auto add = [&](void* p, size_t words) { if (p && words) { L.p[k] = (int*)p; L.n[k] = (int)words; ++k; } };
        add(ctx->colorMask, 2 * (size_t)(ctx->nDyn > 0 ? ctx->nDyn : 1));
        add(ctx->keyCursor, (size_t)nKeys + 1);
        if (ctx->islandsOn) {
            if ((rc = pb_islands_alloc(ctx))) return rc;
            add(ctx->keyStart, (size_t)nKeys + 1);
            add(ctx->islandStats, 4);
            add(ctx->bodyStart, ctx->bodyStart ? (size_t)G + 2 : 0);
            add(ctx->bodyCursor, ctx->bodyCursor ? (size_t)G + 2 : 0);
            if (ctx->nJoints && (rc = pb_joint_lists_alloc(ctx))) return rc;
            if (ctx->nJoints) { add(ctx->jointStart, (size_t)G * 8 + 10); add(ctx->jointSortTmp[0], (size_t)G * 8 + 10); }
        }
        size_t most = 0;
        for (int i = 0; i < k; ++i) most = std::max(most, (size_t)L.n[i]);
        ++ctx->launches, k_build_clear<<<std::max(1, std::min(ctx->numSMs * 4, (int)((most + 1023) / 1024))), 256, 0, ctx->stream>>>(L);
    }
    // Two independent strands meet in the manifold order: the COLOURS (k_color) and the GROUPS (island search).  With islands on they run
    // side by side -- the colouring on a second stream behind the clear kernel -- and so do, further down, the joint lists (they need the
    // groups only) and the ordering of the manifolds.  A step of a small scene is a chain of few-microsecond kernels: what runs beside
    // another is off that chain (512 ragdoll scenes: ~30 us of 420).
    struct StreamSwap {        // launches of the enclosed calls go to the side stream (everything here launches on ctx->stream)
        pb_ctx* c; cudaStream_t keep;
        StreamSwap(pb_ctx* c_, cudaStream_t s) : c(c_), keep(c_->stream) { c->stream = s; }
        ~StreamSwap() { c->stream = keep; }
    };
    const bool fork = ctx->islandsOn && !ctx->deterministic && ctx->buildFork && ctx->sideStream && G <= 4000;      // (G: the joint lists' scan stays a one-CTA scan without scratch)
    cudaStream_t mainStream = ctx->stream;
    if (fork) { PB_CUDA(ctx, cudaEventRecord(ctx->evFork, mainStream)); PB_CUDA(ctx, cudaStreamWaitEvent(ctx->sideStream, ctx->evFork, 0)); }
    if (ctx->deterministic) {
        if (!ctx->jpBest) {
            if ((rc = pb_alloc(ctx, &ctx->jpBest, (size_t)ctx->caps.max_bodies)) || (rc = pb_alloc(ctx, &ctx->jpScratch, 64))) return rc;
            int perSM = 0;
            PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_color_jp, 256, 0));
            ctx->jpGrid = ctx->numSMs * std::max(1, std::min(perSM, 4));
        }
        PB_CUDA(ctx, cudaMemsetAsync(ctx->jpScratch, 0, sizeof(int) * 64, ctx->stream));
        const int4* mKey = ctx->mKey; int* counters = ctx->counters; const int* colRow = ctx->colRow; int nDyn = ctx->nDyn; const int* kin = ctx->kinematic;
        unsigned long long* mask = ctx->colorMask; unsigned long long* best = ctx->jpBest; unsigned int* sk = ctx->mSortKeyA;
        unsigned int* bar = (unsigned int*)ctx->jpScratch; int* rem = ctx->jpScratch + 16;
        void* args[] = { &mKey, &counters, &maxM, &colRow, &nDyn, &kin, &mask, &best, &sk, &bar, &rem };
        ++ctx->launches;
        PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_color_jp, dim3(ctx->jpGrid), dim3(256), args, 0, ctx->stream));
        ++ctx->launches, k_color_starts<<<1, 32, 0, ctx->stream>>>(ctx->counters, ctx->islandsOn ? nullptr : ctx->keyStart + (size_t)G * PB_KEY_COLORS);
    } else {
        cudaStream_t cs = fork ? ctx->sideStream : mainStream;
        ++ctx->launches, k_color<<<blocks, 256, 0, cs>>>(ctx->mKey, ctx->counters, maxM, ctx->colRow, ctx->nDyn, ctx->kinematic, ctx->colorMask, ctx->mSortKeyA);
        ++ctx->launches, k_color_starts<<<1, 32, 0, cs>>>(ctx->counters, ctx->islandsOn ? nullptr : ctx->keyStart + (size_t)G * PB_KEY_COLORS);
        if (fork) PB_CUDA(ctx, cudaEventRecord(ctx->evColour, cs));
    }
    if (ctx->islandsOn) {
        if ((rc = pb_islands_build(ctx))) return rc;
        if (fork && ctx->nJoints) {
            // the joint lists need the groups and nothing else: side stream, joined in front of the solver (end of this function)
            PB_CUDA(ctx, cudaEventRecord(ctx->evGroups, mainStream));
            PB_CUDA(ctx, cudaStreamWaitEvent(ctx->sideStream, ctx->evGroups, 0));
            { StreamSwap sw(ctx, ctx->sideStream); rc = pb_joint_lists(ctx); }
            if (rc) return rc;
            PB_CUDA(ctx, cudaEventRecord(ctx->evJoints, ctx->sideStream));
        } else if ((rc = pb_joint_lists(ctx))) return rc;
        if (fork) PB_CUDA(ctx, cudaStreamWaitEvent(mainStream, ctx->evColour, 0));      // the manifold keys below need the colours
    }
    // Solve order = a counting sort by (group, colour, single | multi): the run table holds the first slot of every key (scanned key
    // histogram), so one scatter pass places every manifold (slot = run start + arrival rank, one atomic per distinct key per warp).
    // The order inside a run is arrival order: a run's manifolds share no dynamic body (the overflow bucket is solved in slot order,
    // whatever that is), and the taps report the order that was used.  (This replaced a stable radix sort: 2 passes, ~0.13 ms at 2 M manifolds.)
    unsigned int keyBase = 0, keyLimit = PB_KEY_COLORS;
    if (ctx->islandsOn) {
        // group-major order: every local group's manifolds are contiguous (colour by colour inside), the global group comes last
        ++ctx->launches, k_island_keys<<<blocks, 256, 0, ctx->stream>>>(ctx->counters, maxM, ctx->mKey, ctx->colRow, ctx->nDyn, ctx->kinematic, ctx->bodyGroup, G,
                                                                       ctx->mSortKeyA, ctx->keyStart);
        if ((rc = pb_exclusive_scan(ctx, ctx->keyStart, ctx->keyStart, nKeys + 1, (int*)ctx->radixHist))) return rc;
        keyLimit = (unsigned int)nKeys;
    } else {
        keyBase = (unsigned int)G * PB_KEY_COLORS;     // plain colour-major order: the run table of group G (k_color_starts)
    }
    ++ctx->launches, k_scatter_by_key<<<blocks, 256, 0, ctx->stream>>>(ctx->counters, maxM, ctx->mSortKeyA, keyBase, keyLimit, ctx->keyStart, ctx->keyCursor, ctx->mSortValB, ctx->mSortKeyB);
    // (islands off: only the run table of group G is this step's)
    if (ctx->deterministic) {
        const int nTables = ctx->islandsOn ? G + 1 : 1;
        ++ctx->launches, k_sort_overflow_run<<<pb_grid(nTables, 128), 128, 0, ctx->stream>>>(nTables, ctx->keyStart + (ctx->islandsOn ? 0 : (size_t)G * PB_KEY_COLORS), ctx->mKey, ctx->mSortValB);
    }
    ctx->mSorted = ctx->mSortValB;
    ctx->mSortedKeys = ctx->mSortKeyB;
    int cur = ctx->curBuf, prev = cur ^ 1;
    int* pointOfs = ctx->cPointOfsBuf[cur];
    ++ctx->launches, k_gather_np<<<blocks, 256, 0, ctx->stream>>>(ctx->counters, ctx->mSorted, ctx->mKey, ctx->cNpBuf[cur]);
    rc = pb_exclusive_scan_dev(ctx, ctx->cNpBuf[cur], pointOfs, ctx->counters + CNT_MANIFOLDS, maxM, ctx->rawHint < 0 ? -1 : 2 * ctx->rawHint + 4096, (int*)ctx->radixHist);
    if (rc) return rc;
    // (the contact cache only carries restitution targets: a scene without restitution neither fills nor reads it)
    if (ctx->anyRestitution) cudaMemsetAsync(ctx->cacheTag[cur], 0, sizeof(unsigned long long) * (size_t)ctx->cacheSize, ctx->stream);
    ++ctx->launches, k_contact_build<<<blocks, 128, 0, ctx->stream>>>(ctx->counters, ctx->mSorted, ctx->mKey, ctx->mNormal, ctx->mPts, pointOfs, ctx->colRow, ctx->colMat,
        ctx->nDyn, ctx->kinematic, ctx->pos, ctx->quat, ctx->vel, ctx->angvel, ctx->comInvMass,
        ctx->cHead, ctx->cBodies, ctx->cRowsT, ctx->cNormal, ctx->cSoft, ctx->cNpBuf[cur], ctx->pR0T[cur], ctx->pR1,
        (ctx->cacheValid && ctx->anyRestitution) ? ctx->cacheTag[prev] : nullptr, ctx->cacheVal[prev], ctx->cPointOfsBuf[prev], ctx->cNpBuf[prev], ctx->pR0T[prev],
        ctx->cacheTag[cur], ctx->cacheVal[cur], ctx->cacheSize - 1);
    if (fork && ctx->nJoints) PB_CUDA(ctx, cudaStreamWaitEvent(mainStream, ctx->evJoints, 0));
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

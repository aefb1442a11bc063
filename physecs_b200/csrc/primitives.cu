// Device-wide primitives used by several stages: exclusive scan and a stable LSD radix sort of
// (32-bit key, 32-bit payload) pairs.  They replace the reference's persistent insertion sort
// (src/Physecs.cpp:121-133) as the ordering machinery of the broadphase and also order manifolds by colour.
//
// Radix pass = 3 launches: per-warp-tile digit histogram -> scan of the (digit-major) histogram ->
// stable scatter where each warp ranks its keys round by round with __match_any_sync.
#include "pb_ctx.h"

#define SCAN_THREADS 1024
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int warpInclusiveScan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// inclusive scan over a block of SCAN_THREADS; returns inclusive value, total via *total
__device__ __forceinline__ int blockInclusiveScan(int v, int* total) {
    __shared__ int warpSums[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warpInclusiveScan(v, lane);
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = (lane < (blockDim.x >> 5)) ? warpSums[lane] : 0;
        s = warpInclusiveScan(s, lane);
        warpSums[lane] = s;
    }
    __syncthreads();
    int offset = w ? warpSums[w - 1] : 0;
    if (total) *total = warpSums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return inc + offset;
}

__global__ void k_scan_reduce(const int* __restrict__ in, int* __restrict__ blockSums, int n) {
    int base = blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) s += in[idx];
    }
    int total;
    blockInclusiveScan(s, &total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

__global__ void k_scan_blocksums(int* __restrict__ blockSums, int nb) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int idx = base + threadIdx.x;
        int v = idx < nb ? blockSums[idx] : 0;
        int total;
        int inc = blockInclusiveScan(v, &total);
        int c = carry;
        if (idx < nb) blockSums[idx] = c + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

__global__ void k_scan_final(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ blockSums, int n) {
    int base = blockIdx.x * SCAN_TILE;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    int inc = blockInclusiveScan(s, nullptr);
    int run = blockSums[blockIdx.x] + inc - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
}

// small inputs: one block walks the array tile by tile with a running carry (one launch instead of three)
__global__ void k_scan_small(const int* in, int* out, int n) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_TILE) {
        int v[SCAN_ITEMS];
        int s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int idx = base + threadIdx.x * SCAN_ITEMS + i;
            v[i] = idx < n ? in[idx] : 0;
            s += v[i];
        }
        int total;
        int inc = blockInclusiveScan(s, &total);
        int run = carry + inc - s;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int idx = base + threadIdx.x * SCAN_ITEMS + i;
            if (idx < n) out[idx] = run;
            run += v[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

// out[i] = sum(in[0..i-1]); in and out may alias.  scratch: at least ceil(n/4096)+1 ints.
int pb_exclusive_scan(pb_ctx* ctx, const int* in, int* out, int n, int* scratch) {
    if (n <= 0) return PB_OK;
    if (n <= 8 * SCAN_TILE) {
        ++ctx->launches, k_scan_small<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, n);
        PB_CUDA(ctx, cudaGetLastError());
        return PB_OK;
    }
    int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    ++ctx->launches, k_scan_reduce<<<nb, SCAN_THREADS, 0, ctx->stream>>>(in, scratch, n);
    ++ctx->launches, k_scan_blocksums<<<1, SCAN_THREADS, 0, ctx->stream>>>(scratch, nb);
    ++ctx->launches, k_scan_final<<<nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, scratch, n);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// The same scan over a DEVICE-side element count (*nDev, clamped to cap): the launch shape follows the capacity, blocks past the
// count only publish a zero block sum.  Lets a step order its manifolds without the host ever reading how many there are.
__global__ void k_scan_reduce_dev(const int* __restrict__ in, int* __restrict__ blockSums, const int* __restrict__ nDev, int cap) {
    const int n = min(*nDev, cap);
    if (blockIdx.x * SCAN_TILE >= n) { if (threadIdx.x == 0) blockSums[blockIdx.x] = 0; return; }      // uniform per CTA
    int base = blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) s += in[idx];
    }
    int total;
    blockInclusiveScan(s, &total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}
__global__ void k_scan_final_dev(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ blockSums, const int* __restrict__ nDev, int cap) {
    const int n = min(*nDev, cap);
    if (blockIdx.x * SCAN_TILE >= n) return;
    int base = blockIdx.x * SCAN_TILE;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    int inc = blockInclusiveScan(s, nullptr);
    int run = blockSums[blockIdx.x] + inc - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int idx = base + threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
}
__global__ void k_scan_small_dev(const int* in, int* out, const int* __restrict__ nDev, int cap) {
    __shared__ int carry;
    const int n = min(*nDev, cap);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_TILE) {
        int v[SCAN_ITEMS];
        int s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int idx = base + threadIdx.x * SCAN_ITEMS + i;
            v[i] = idx < n ? in[idx] : 0;
            s += v[i];
        }
        int total;
        int inc = blockInclusiveScan(s, &total);
        int run = carry + inc - s;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            int idx = base + threadIdx.x * SCAN_ITEMS + i;
            if (idx < n) out[idx] = run;
            run += v[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}
// `bound`: host-side guess of the count (shapes the launch only; any value is correct).  One CTA walks the array while the guess is
// small, the three-launch form covers the whole capacity otherwise.
int pb_exclusive_scan_dev(pb_ctx* ctx, const int* in, int* out, const int* nDev, int cap, int bound, int* scratch) {
    if (cap <= 0) return PB_OK;
    if (bound < 0 || bound > cap) bound = cap;
    if (bound <= 8 * SCAN_TILE) {
        ++ctx->launches, k_scan_small_dev<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, nDev, cap);
        PB_CUDA(ctx, cudaGetLastError());
        return PB_OK;
    }
    int nb = (cap + SCAN_TILE - 1) / SCAN_TILE;
    ++ctx->launches, k_scan_reduce_dev<<<nb, SCAN_THREADS, 0, ctx->stream>>>(in, scratch, nDev, cap);
    ++ctx->launches, k_scan_blocksums<<<1, SCAN_THREADS, 0, ctx->stream>>>(scratch, nb);
    ++ctx->launches, k_scan_final_dev<<<nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, scratch, nDev, cap);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------------------
#define RS_WARPS 8
// keys per warp tile = 32 * RS_ITEMS: 16 rounds per warp for large inputs; 4 for small ones, where a 512-key tile would leave
// most SMs idle (a 45 k-key sort had 11 CTAs and took ~20 us per kernel, ncu)
template <int RS_ITEMS>
__global__ void k_radix_hist(const unsigned int* __restrict__ keys, unsigned int* __restrict__ hist, int n, int shift, int numTiles) {
    __shared__ unsigned int sh[RS_WARPS][256];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 256; i += 32) sh[w][i] = 0;
    __syncwarp();
    int tile = blockIdx.x * RS_WARPS + w;
    if (tile < numTiles) {
        int base = tile * (32 * RS_ITEMS);
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r) {
            int idx = base + r * 32 + lane;
            if (idx < n) atomicAdd(&sh[w][(keys[idx] >> shift) & 255u], 1u);
        }
        __syncwarp();
        for (int i = lane; i < 256; i += 32) hist[(size_t)i * numTiles + tile] = sh[w][i];
    }
}

template <int RS_ITEMS>
__global__ void k_radix_scatter(const unsigned int* __restrict__ keys, const int* __restrict__ vals,
                                unsigned int* __restrict__ keysOut, int* __restrict__ valsOut,
                                const unsigned int* __restrict__ hist, int n, int shift, int numTiles) {
    __shared__ unsigned int sh[RS_WARPS][256];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x * RS_WARPS + w;
    if (tile >= numTiles) return;
    for (int i = lane; i < 256; i += 32) sh[w][i] = hist[(size_t)i * numTiles + tile];
    __syncwarp();
    int base = tile * (32 * RS_ITEMS);
    unsigned int ltMask = (1u << lane) - 1u;
#pragma unroll 1
    for (int r = 0; r < RS_ITEMS; ++r) {
        int idx = base + r * 32 + lane;
        bool valid = idx < n;
        unsigned int active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned int k = keys[idx];
            int v = vals[idx];
            unsigned int d = (k >> shift) & 255u;
            unsigned int peers = __match_any_sync(active, d);
            unsigned int rank = __popc(peers & ltMask);
            unsigned int b = sh[w][d];
            __syncwarp(active);
            if (rank == 0) sh[w][d] = b + __popc(peers);
            __syncwarp(active);
            keysOut[b + rank] = k;
            valsOut[b + rank] = v;
        }
    }
}

// Small inputs (n <= SORT_SMALL_MAX): every pass of the LSD sort inside ONE CTA, keys and payloads ping-ponging in shared memory
// -- one launch instead of (histogram + scan + scatter) x passes.  Same stable 8-bit passes: each warp owns a contiguous chunk,
// ranks its keys round by round with match_any, and the per-(digit, warp) counts are scanned digit-major by the whole CTA.
#define SORT_SMALL_MAX 8192
#define SORT_SMALL_THREADS 1024
__global__ void __launch_bounds__(SORT_SMALL_THREADS) k_sort_small(unsigned int* __restrict__ keys, int* __restrict__ vals, int n, int bits) {
    extern __shared__ unsigned int smem[];
    unsigned int* k0 = smem; unsigned int* k1 = k0 + SORT_SMALL_MAX;
    int* v0 = (int*)(k1 + SORT_SMALL_MAX); int* v1 = v0 + SORT_SMALL_MAX;
    unsigned int* cnt = (unsigned int*)(v1 + SORT_SMALL_MAX);        // [256 digits][32 warps], digit-major
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nWarps = SORT_SMALL_THREADS / 32;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { k0[i] = keys[i]; v0[i] = vals[i]; }
    const int chunk = ((n + nWarps - 1) / nWarps + 31) & ~31;          // keys per warp, a multiple of 32
    const int begin = min(w * chunk, n), end = min(begin + chunk, n);
    const unsigned int ltMask = (1u << lane) - 1u;
    for (int shift = 0; shift < bits; shift += 8) {
        for (int i = threadIdx.x; i < 256 * nWarps; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        // per-warp digit counts
        for (int base = begin; base < end; base += 32) {
            int idx = base + lane;
            if (idx < end) atomicAdd(&cnt[((k0[idx] >> shift) & 255u) * nWarps + w], 1u);
        }
        __syncthreads();
        // exclusive scan over the 256 * nWarps counts (digit-major == output order), 8 entries per thread
        {
            const int per = 256 * nWarps / SORT_SMALL_THREADS;      // 8
            unsigned int loc[8]; unsigned int s = 0;
#pragma unroll
            for (int q = 0; q < per; ++q) { loc[q] = cnt[threadIdx.x * per + q]; s += loc[q]; }
            int inc = warpInclusiveScan((int)s, lane);
            __shared__ int warpTot[32];
            if (lane == 31) warpTot[w] = inc;
            __syncthreads();
            if (w == 0) { int t = warpTot[lane]; t = warpInclusiveScan(t, lane); warpTot[lane] = t; }
            __syncthreads();
            unsigned int run = (unsigned int)(inc - (int)s + (w ? warpTot[w - 1] : 0));
#pragma unroll
            for (int q = 0; q < per; ++q) { cnt[threadIdx.x * per + q] = run; run += loc[q]; }
        }
        __syncthreads();
        // stable scatter: each warp walks its chunk in order
        for (int base = begin; base < end; base += 32) {
            int idx = base + lane;
            bool valid = idx < end;
            unsigned int active = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                unsigned int k = k0[idx];
                int v = v0[idx];
                unsigned int d = (k >> shift) & 255u;
                unsigned int peers = __match_any_sync(active, d);
                unsigned int rank = __popc(peers & ltMask);
                unsigned int b = cnt[d * nWarps + w];
                __syncwarp(active);
                if (rank == 0) cnt[d * nWarps + w] = b + __popc(peers);
                __syncwarp(active);
                k1[b + rank] = k; v1[b + rank] = v;
            }
        }
        __syncthreads();
        unsigned int* tk = k0; k0 = k1; k1 = tk;
        int* tv = v0; v0 = v1; v1 = tv;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { keys[i] = k0[i]; vals[i] = v0[i]; }
}

// Large inputs: every pass of the LSD sort inside ONE cooperative launch.  The multi-launch version above costs three launches and a
// 524 k-entry histogram scan per pass (56 us per pass at 1 M keys, of which the keys themselves are ~3 us of traffic); here each CTA
// owns a contiguous chunk, histograms are per CTA (256 x grid entries), the columns are scanned by the first 256 CTAs, and the three
// phases of a pass are separated by device-wide barriers (atom.acq_rel + ld.acquire on one counter, as in the substep solver).
// Same stable 8-bit passes: chunk order, then warp order inside the chunk, then lane order inside a round (match_any ranks).
// Everything another CTA may have written is read through L2 (__ldcg): L1 is not coherent across SMs and the buffers ping-pong.
#define RC_THREADS 256
#define RC_WARPS (RC_THREADS / 32)
__device__ __forceinline__ void coopBarrier(unsigned int* counter, unsigned int& target) {
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

__global__ void __launch_bounds__(RC_THREADS) k_radix_sort_coop(unsigned int* keysA, int* valsA, unsigned int* keysB, int* valsB, int n, int bits, int chunk,
                                                                unsigned int* hist, unsigned int* barrier) {
    __shared__ unsigned int sh[RC_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, cta = blockIdx.x, nCta = gridDim.x;
    unsigned int* totals = hist + 256 * (size_t)nCta;
    const int wchunk = chunk / RC_WARPS;                 // a multiple of 32
    const long long first = (long long)cta * chunk + (long long)w * wchunk;
    const int begin = (int)(first < n ? first : n), end = min(begin + wchunk, n);
    const unsigned int ltMask = (1u << lane) - 1u;
    unsigned int target = 0;
    unsigned int* src = keysA; int* srcV = valsA; unsigned int* dst = keysB; int* dstV = valsB;
    for (int shift = 0; shift < bits; shift += 8) {
        // phase 1: digit counts of this CTA's chunk (per warp, then summed; the per-warp exclusive prefixes stay in shared memory)
        for (int i = lane; i < 256; i += 32) sh[w][i] = 0;
        __syncwarp();
        for (int base = begin; base < end; base += 32) {
            int idx = base + lane;
            if (idx < end) atomicAdd(&sh[w][(__ldcg(&src[idx]) >> shift) & 255u], 1u);
        }
        __syncthreads();
        {
            const int d = threadIdx.x;
            unsigned int run = 0;
#pragma unroll
            for (int k = 0; k < RC_WARPS; ++k) { unsigned int t = sh[k][d]; sh[k][d] = run; run += t; }
            hist[(size_t)d * nCta + cta] = run;
        }
        coopBarrier(barrier, target);
        // phase 2: exclusive scan of every digit's column over the CTAs, digit totals
        for (int d = cta; d < 256; d += nCta) {
            unsigned int carry = 0;
            for (int base = 0; base < nCta; base += RC_THREADS) {
                int j = base + threadIdx.x;
                int v = j < nCta ? (int)__ldcg(&hist[(size_t)d * nCta + j]) : 0;
                int total;
                int inc = blockInclusiveScan(v, &total);
                if (j < nCta) hist[(size_t)d * nCta + j] = carry + (unsigned int)(inc - v);
                carry += (unsigned int)total;
            }
            if (threadIdx.x == 0) totals[d] = carry;
        }
        coopBarrier(barrier, target);
        // phase 3: first output slot of (digit, this CTA, warp), then the stable scatter
        {
            const int d = threadIdx.x;
            int tot = (int)__ldcg(&totals[d]);
            int inc = blockInclusiveScan(tot, nullptr);
            unsigned int base = (unsigned int)(inc - tot) + __ldcg(&hist[(size_t)d * nCta + cta]);
#pragma unroll
            for (int k = 0; k < RC_WARPS; ++k) sh[k][d] += base;
        }
        __syncthreads();
        for (int base = begin; base < end; base += 32) {
            int idx = base + lane;
            bool valid = idx < end;
            unsigned int active = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                unsigned int k = __ldcg(&src[idx]);
                int v = __ldcg(&srcV[idx]);
                unsigned int d = (k >> shift) & 255u;
                unsigned int peers = __match_any_sync(active, d);
                unsigned int rank = __popc(peers & ltMask);
                unsigned int b = sh[w][d];
                __syncwarp(active);
                if (rank == 0) sh[w][d] = b + __popc(peers);
                __syncwarp(active);
                dst[b + rank] = k;
                dstV[b + rank] = v;
            }
        }
        coopBarrier(barrier, target);
        unsigned int* tk = src; src = dst; dst = tk;
        int* tv = srcV; srcV = dstV; dstV = tv;
    }
}

// Sort (keysA, valsA) by the low `bits` bits of the key, 8 bits per pass, ping-ponging with (keysB, valsB).
// hist must hold 256 * ceil(n/512) (+ scan scratch of ceil(that/4096)+1) uints.
int pb_radix_sort_pairs(pb_ctx* ctx, unsigned int* keysA, int* valsA, unsigned int* keysB, int* valsB, int n, int bits,
                        unsigned int* hist, int histCapTiles, bool* resultInA) {
    *resultInA = true;
    if (n <= 1) return PB_OK;
    if (n <= SORT_SMALL_MAX) {
        const size_t smemBytes = sizeof(unsigned int) * (4 * SORT_SMALL_MAX + 256 * (SORT_SMALL_THREADS / 32));
        if (!ctx->sortSmallOptIn) { PB_CUDA(ctx, cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes)); ctx->sortSmallOptIn = true; }
        ++ctx->launches, k_sort_small<<<1, SORT_SMALL_THREADS, smemBytes, ctx->stream>>>(keysA, valsA, n, bits);    // sorted in place: result in A
        PB_CUDA(ctx, cudaGetLastError());
        return PB_OK;
    }
    if (ctx->sortCoopMode) {
        if (!ctx->sortCoopGrid) {
            int occ = 0;
            PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_radix_sort_coop, RC_THREADS, 0));
            if (occ > 4) occ = 4;
            ctx->sortCoopGrid = occ * ctx->numSMs;
            if (ctx->sortCoopGrid > 0) { int rc = pb_alloc(ctx, &ctx->sortBarrier, 64); if (rc) return rc; }
        }
        int nCta = (n + 1535) / 1536;
        if (nCta > ctx->sortCoopGrid) nCta = ctx->sortCoopGrid;
        if (nCta >= 1 && nCta + 2 <= histCapTiles) {
            int chunk = (((n + nCta - 1) / nCta) + RC_THREADS - 1) / RC_THREADS * RC_THREADS;
            int passes = (bits + 7) / 8;
            PB_CUDA(ctx, cudaMemsetAsync(ctx->sortBarrier, 0, sizeof(unsigned int), ctx->stream));
            void* args[] = { &keysA, &valsA, &keysB, &valsB, &n, &bits, &chunk, &hist, &ctx->sortBarrier };
            ++ctx->launches;
            PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_radix_sort_coop, dim3(nCta), dim3(RC_THREADS), args, 0, ctx->stream));
            *resultInA = (passes & 1) == 0;
            return PB_OK;
        }
    }
    int items = 16;
    if ((n + 127) / 128 <= histCapTiles && n <= (1 << 19)) items = 4;
    const int tileKeys = 32 * items;
    int numTiles = (n + tileKeys - 1) / tileKeys;
    if (numTiles > histCapTiles) return pb_fail(ctx, PB_ECAPACITY, "radix sort histogram capacity");
    int blocks = (numTiles + RS_WARPS - 1) / RS_WARPS;
    int histN = 256 * numTiles;
    int* scanScratch = (int*)(hist + (size_t)256 * histCapTiles);
    unsigned int* src = keysA; int* srcV = valsA; unsigned int* dst = keysB; int* dstV = valsB;
    for (int shift = 0; shift < bits; shift += 8) {
        if (items == 4) ++ctx->launches, k_radix_hist<4><<<blocks, RS_WARPS * 32, 0, ctx->stream>>>(src, hist, n, shift, numTiles);
        else ++ctx->launches, k_radix_hist<16><<<blocks, RS_WARPS * 32, 0, ctx->stream>>>(src, hist, n, shift, numTiles);
        int rc = pb_exclusive_scan(ctx, (const int*)hist, (int*)hist, histN, scanScratch);
        if (rc) return rc;
        if (items == 4) ++ctx->launches, k_radix_scatter<4><<<blocks, RS_WARPS * 32, 0, ctx->stream>>>(src, srcV, dst, dstV, hist, n, shift, numTiles);
        else ++ctx->launches, k_radix_scatter<16><<<blocks, RS_WARPS * 32, 0, ctx->stream>>>(src, srcV, dst, dstV, hist, n, shift, numTiles);
        unsigned int* t = src; src = dst; dst = t;
        int* tv = srcV; srcV = dstV; dstV = tv;
        *resultInA = !*resultInA;
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

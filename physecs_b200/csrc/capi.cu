// C ABI (include/physecs_b200.h): context, scene upload, per-step orchestration and parity taps.
// Step orchestration == physecs::Scene::simulate (reference src/Physecs.cpp:112-561), device side.
#include "pb_ctx.h"
#include "pb_math.cuh"
#include "trimesh_build.h"
#include <algorithm>
#include <cstring>
#include <cuda_profiler_api.h>

int pb_fail(pb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int pb_joints_upload(pb_ctx* ctx, int n, const int* type, const int* row0, const int* row1, const float* a0p, const float* a0q,
                     const float* a1p, const float* a1q, const float* params8, const int* color);
void pb_joints_free(pb_ctx* ctx);
int pb_joints_update_params(pb_ctx* ctx, int n, const float* params8);
int pb_joints_keep_state(pb_ctx* ctx, int n, const int* oldIndex);

// ---- packed host layout <-> float4 SoA ----------------------------------------------------------------------------
__global__ void k_unpack3(int n, const float* __restrict__ src, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}
__global__ void k_unpack4(int n, const float* __restrict__ src, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
__global__ void k_pack3(int n, const float4* __restrict__ src, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = src[i]; dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; }
}
__global__ void k_pack4(int n, const float4* __restrict__ src, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = src[i]; dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w; }
}
// interleaved velocity buffer: dst[2*i] = {v, invMass}, dst[2*i + 1] = {w, 0}; either source may be null (left untouched)
__global__ void k_unpack_vel(int n, const float* __restrict__ v3, const float* __restrict__ w3, const float4* __restrict__ comInvMass, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (v3) dst[2 * i] = make_float4(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2], comInvMass[i].w);
    if (w3) dst[2 * i + 1] = make_float4(w3[3 * i], w3[3 * i + 1], w3[3 * i + 2], 0.f);
}
__global__ void k_pack_vel(int n, const float4* __restrict__ src, float* __restrict__ v3, float* __restrict__ w3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (v3) { float4 v = src[2 * i]; v3[3 * i] = v.x; v3[3 * i + 1] = v.y; v3[3 * i + 2] = v.z; }
    if (w3) { float4 w = src[2 * i + 1]; w3[3 * i] = w.x; w3[3 * i + 1] = w.y; w3[3 * i + 2] = w.z; }
}
__global__ void k_refresh_invmass(int n, const float4* __restrict__ comInvMass, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[2 * i].w = comInvMass[i].w;
}
__global__ void k_com_invmass(int n, const float* __restrict__ com, const float* __restrict__ invMass, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(com[3 * i], com[3 * i + 1], com[3 * i + 2], invMass[i]);
}
__global__ void k_unpack_m3(int n, const float* __restrict__ src, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int c = 0; c < 3; ++c) dst[3 * i + c] = make_float4(src[9 * i + 3 * c], src[9 * i + 3 * c + 1], src[9 * i + 3 * c + 2], 0.f);
}
__global__ void k_scatter_rows(int n, const int* __restrict__ rows, const float* __restrict__ p3, const float* __restrict__ q4,
                               float4* __restrict__ pos, float4* __restrict__ quat, int* __restrict__ mark) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = rows[i];
    pos[r] = make_float4(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], 0.f);
    quat[r] = make_float4(q4[4 * i], q4[4 * i + 1], q4[4 * i + 2], q4[4 * i + 3]);
    mark[r] = 1;
}

__global__ void k_scatter_bounds(int n, const int* __restrict__ cols, const float* __restrict__ b6, float4* __restrict__ mn, float4* __restrict__ mx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cols[i];
    mn[c] = make_float4(b6[6 * i], b6[6 * i + 1], b6[6 * i + 2], 0.f);
    mx[c] = make_float4(b6[6 * i + 3], b6[6 * i + 4], b6[6 * i + 5], 0.f);
}
// bounds kept across a collider re-upload: old collider o lives on as collider oldToNew[o] (-1 = gone, or it starts from creation bounds)
__global__ void k_restore_bounds(int nOld, const int* __restrict__ oldToNew, int nNew, const float4* __restrict__ keptMin, const float4* __restrict__ keptMax,
                                 float4* __restrict__ mn, float4* __restrict__ mx) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nOld) return;
    int c = oldToNew[o];
    if (c < 0 || c >= nNew) return;
    mn[c] = keptMin[o];
    mx[c] = keptMax[o];
}
// BroadPhaseEntry::isDynamic (Physecs.cpp:56-77, :753-770): bit2 of the device collider flags
__global__ void k_col_dynamic_flag(int n, const int* __restrict__ colRow, int nDyn, const int* __restrict__ kinematic, int* __restrict__ colFlags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int row = colRow[i];
    bool dyn = row < nDyn && !kinematic[row];
    colFlags[i] = (colFlags[i] & 3) | (dyn ? 4 : 0);
}

// A call that changes bounds / colliders / body kinds: a broadphase already enqueued by pb_step_begin is stale now.  If the narrowphase
// went out as well (pb_step_narrowphase), its host-side bookkeeping is taken back -- the device's results are scratch that the next
// pb_step_begin resets -- and pb_step starts the step over.
static void staleBroadphase(pb_ctx* ctx) {
    if (ctx->stepNarrowed) {
        ctx->stepNarrowed = false;
        ctx->stepPending = false;        // nobody collects the abandoned narrowphase
        ctx->curBuf ^= 1;
        ctx->cacheValid = ctx->undoCacheValid; ctx->cacheBuilt = ctx->undoCacheBuilt;
    }
    ctx->queryTreeValid = false; ctx->stepBegun = false; ctx->mainMarked = false;
}
#define PB_STALE_BROADPHASE(ctx) staleBroadphase(ctx)

static int ensureStage(pb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stageBytes) return PB_OK;
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr; ctx->stageBytes = 0;
    cudaError_t e = cudaMalloc((void**)&ctx->stage, bytes);
    if (e != cudaSuccess) return pb_fail(ctx, PB_ECUDA, std::string("cudaMalloc stage: ") + cudaGetErrorString(e));
    ctx->stageBytes = bytes;
    return PB_OK;
}

// upload a packed host array (n*width floats) and expand to float4 rows at dst
static int uploadVec(pb_ctx* ctx, const float* host, int n, int width, float4* dst) {
    if (n <= 0) return PB_OK;
    size_t bytes = sizeof(float) * (size_t)n * width;
    int rc = ensureStage(ctx, bytes); if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (width == 3) ++ctx->launches, k_unpack3<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ctx->stage, dst);
    else ++ctx->launches, k_unpack4<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, ctx->stage, dst);
    // the staging buffer is reused by the next upload on the same stream: stream order keeps this safe
    return PB_OK;
}
template <class T> static int uploadRaw(pb_ctx* ctx, const T* host, size_t n, T* dst) {
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaMemcpyAsync(dst, host, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
    return PB_OK;
}

int pb_wait_poses(pb_ctx* ctx) {
    if (!ctx->posePending) return PB_OK;
    PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evPoseReady, 0));
    ctx->posePending = false;
    return PB_OK;
}
int pb_wait_velocities(pb_ctx* ctx) {
    int rc = pb_wait_poses(ctx); if (rc) return rc;
    if (!ctx->velPending) return PB_OK;
    PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evVelReady, 0));
    ctx->velPending = false;
    return PB_OK;
}

extern "C" {

// enable/disable per-phase timing inside the persistent substep kernel (CTA 0 stamps %globaltimer at every grid barrier);
// enabling resets the accumulators
int pb_set_profile(pb_ctx* ctx, int on) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->profile = on != 0;
    unsigned long long tmp[16];
    return pb_solve_profile(ctx, tmp, true);
}
// accumulated phase times since pb_set_profile(1): ms8 / count8 indexed by phase kind
// (0 = integrate velocities, 1 = contact + joint prep, 2 = contact solve colour phases, 3 = joint solve, 4 = integrate positions,
// 5 = per-CTA island sweeps);
// count = number of phases (grid barriers) of that kind
int pb_get_profile(pb_ctx* ctx, double* ms8, long long* count8) {
    cudaSetDevice(ctx->device);
    unsigned long long raw[16] = {0};
    int rc = pb_solve_profile(ctx, raw, false);
    if (rc) return rc;
    for (int k = 0; k < 8; ++k) { ms8[k] = k < 6 ? raw[k] * 1e-6 : 0.0; count8[k] = k < 6 ? (long long)raw[6 + k] : 0; }
    // slot 6: the k_substep_solve launches of the LAST step, timed by CUDA events on the context's stream (ms summed, launches)
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < ctx->evSubCount; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evSub[2 * k], ctx->evSub[2 * k + 1]) == cudaSuccess) { ms8[6] += ms; ++count8[6]; }
    }
    return PB_OK;
}
int pb_get_profile_colors(pb_ctx* ctx, double* ms64, long long* count64) {
    cudaSetDevice(ctx->device);
    unsigned long long raw[2 * PB_MAX_COLORS + 6];
    int rc = pb_solve_profile_colors(ctx, raw);
    if (rc) return rc;
    for (int k = 0; k < PB_MAX_COLORS; ++k) { ms64[k] = raw[k] * 1e-6; count64[k] = (long long)raw[PB_MAX_COLORS + k]; }
    // colours 61..63 never hold timed phases of their own here (63 is the sequential bucket): the last three slots carry CTA 0's
    // local sweep split -- NGS, contact, joint colour phases
    for (int k = 0; k < 3; ++k) { ms64[61 + k] = raw[2 * PB_MAX_COLORS + k] * 1e-6; count64[61 + k] = (long long)raw[2 * PB_MAX_COLORS + 3 + k]; }
    return PB_OK;
}
int pb_set_islands(pb_ctx* ctx, int mode) {
    if (mode < 0 || mode > 2) return pb_fail(ctx, PB_EINVAL, "pb_set_islands: mode 0 (off), 1 (on) or 2 (auto)");
    ctx->islandsMode = mode; ctx->islandsHold = 0;
    return PB_OK;
}
int pb_get_island_stats(pb_ctx* ctx, int* out3) {
    out3[0] = ctx->islandsOn ? 1 : 0; out3[1] = ctx->lastIslandLocal; out3[2] = ctx->lastIslandTotal;
    return PB_OK;
}
int pb_get_broadphase_info(pb_ctx* ctx, int* out3) {
    out3[0] = ctx->stepBrute ? 1 : 0; out3[1] = ctx->lastTileHits; out3[2] = ctx->lastTiles;
    return PB_OK;
}
int pb_set_deterministic(pb_ctx* ctx, int on) { ctx->deterministic = on != 0; return PB_OK; }
unsigned long long pb_get_launches(pb_ctx* ctx) { return ctx->launches; }
void pb_profiler_range(int start) { if (start) cudaProfilerStart(); else cudaProfilerStop(); }

int pb_host_alloc(void** ptr, unsigned long long bytes) { return cudaMallocHost(ptr, bytes) == cudaSuccess ? PB_OK : PB_ECUDA; }
void pb_host_free(void* ptr) { cudaFreeHost(ptr); }
const char* pb_last_error(pb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* pb_stream(pb_ctx* ctx) { return (void*)ctx->stream; }

int pb_ctx_create(int device, const pb_caps* caps, pb_ctx** out) {
    if (!caps || !out) return PB_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) return PB_ECUDA;   // no CPU fallback
    pb_ctx* ctx = new pb_ctx();
    ctx->device = device;
    ctx->caps = *caps;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PB_ECUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->numSMs = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PB_ECUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PB_ECUDA; }
    if (cudaStreamCreateWithFlags(&ctx->readStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PB_ECUDA; }
    if (cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PB_ECUDA; }
    for (cudaEvent_t* e : { &ctx->evFork, &ctx->evColour, &ctx->evGroups, &ctx->evJoints }) cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evPacked, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evMainAtSet, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->evVelReady, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evPoseReady, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evCounters, cudaEventDisableTiming);
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    const size_t R = caps->max_bodies, C = caps->max_colliders, P = caps->max_pairs, M = caps->max_manifolds;
    int rc = 0;
#define A(p, n) if (!rc) rc = pb_alloc(ctx, &ctx->p, (n))
    A(rowEntity, R); A(pos, R); A(quat, R); A(velBuf[0], 2 * R); A(velBuf[1], 2 * R); A(bodyRec, 8 * R);
    if (!rc) {
        ctx->vel = ctx->velBuf[0]; ctx->angvel = ctx->velBuf[0] + 1; ctx->velLive = ctx->velBuf[1]; ctx->angvelLive = ctx->velBuf[1] + 1;
    }
    A(comInvMass, R); A(invIL, 3 * R); A(kinematic, R); A(pseudoLin, R); A(pseudoAng, R); A(colorMask, R); A(rowMark, R);
    // + 1: slot nCol holds the query shape of pb_query_overlap_mtd while the bin kernels run on it
    A(colRow, C); A(colIndex, C); A(colType, C + 1); A(colFlags, C); A(colInfo, C + 1); A(colData, C); A(colMesh, C + 1);
    A(colLPos, C); A(colLQuat, C); A(colParams, C + 1); A(colMat, C); A(colWPos, C + 1); A(colWQuat, C + 1); A(aabbMin, C); A(aabbMax, C);
    A(mortonA, C); A(mortonB, C); A(leafIdA, C); A(leafIdB, C);
    size_t sortMax = std::max(C, M);
    ctx->radixTiles = (int)((sortMax + 511) / 512);
    A(radixHist, (size_t)256 * ctx->radixTiles + (size_t)256 * ctx->radixTiles / 4096 + 1024);
    A(sceneBounds, 32); A(bigList, 64);
    A(nodeLeft, C); A(nodeRight, C); A(nodeParent, C); A(leafParent, C); A(nodeFlag, C); A(nodeRange, C); A(nodeMin, 2 * C); A(nodeMax, 2 * C);
    A(pairs, P); A(pairOrder, 2 * P); A(trigPairs, P); A(colClass, C);
    A(mKey, M); A(mNormal, M); A(mPts, 8 * M); A(mSortKeyA, M); A(mSortKeyB, M); A(mSortValB, M);
    A(cHead, M); A(cBodies, M); A(cRowsT, M); A(cNormal, M); A(cSoft, M);
    const size_t PT = 4 * M;
    for (int b = 0; b < 2; ++b) { A(pR0T[b], PT); A(cPointOfsBuf[b], M + 1); A(cNpBuf[b], M + 1); }
    A(pR1, PT); A(rowA, PT); A(rowB, PT); A(rowC, PT); A(rowD, PT); A(rowE, PT); A(rowF, PT); A(rowG, PT); A(rowL, PT);
    size_t cs = 1; while (cs < 2 * M) cs <<= 1;
    ctx->cacheSize = (int)cs;
    for (int b = 0; b < 2; ++b) { A(cacheTag[b], cs); A(cacheVal[b], cs); }
    A(counters, CNT_TOTAL);
    ctx->islandGroups = ctx->numSMs * 3;     // co-resident 256-thread CTAs of the persistent substep kernel
    if (const char* e = getenv("PB_ISLAND_GROUPS")) if (atoi(e) > 0) ctx->islandGroups = atoi(e);      // experiments: more groups than CTAs (each CTA walks several)
    A(keyStart, (size_t)(ctx->islandGroups + 1) * PB_KEY_COLORS + 1); A(keyCursor, (size_t)(ctx->islandGroups + 1) * PB_KEY_COLORS + 1);
    A(triMeshDev, 64); A(convexDev, 256);
#undef A
    if (!rc && cudaMallocHost((void**)&ctx->hCounters, sizeof(int) * (CNT_TOTAL + 4)) != cudaSuccess) rc = PB_ECUDA;
    if (const char* e = getenv("PB_ISLANDS")) ctx->islandsMode = atoi(e);
    if (const char* e = getenv("PB_BRUTE_FORCE_MAX")) ctx->bruteForceMax = atoi(e);
    if (const char* e = getenv("PB_BRUTE_FORCE_BIG_MAX")) ctx->bruteForceBigMax = atoi(e);
    if (const char* e = getenv("PB_BUILD_FORK")) ctx->buildFork = atoi(e);
    if (const char* e = getenv("PB_FUSED")) ctx->fusedMode = atoi(e);
    if (const char* e = getenv("PB_SORT_COOP")) ctx->sortCoopMode = atoi(e);
    if (const char* e = getenv("PB_BIG_LIST")) ctx->bigListMode = atoi(e);
    if (const char* e = getenv("PB_MESH_LIGHT")) ctx->meshLightMode = atoi(e);
    if (const char* e = getenv("PB_NP_WAVES")) ctx->npWaves = atoi(e) >= 0 ? atoi(e) : 4;
    if (const char* e = getenv("PB_CLUSTER")) if (atoi(e) == 0) ctx->clusterSize = 0;
    if (const char* e = getenv("PB_FUSED_NARROW")) ctx->fusedNarrowMax = atoi(e);
    if (const char* e = getenv("PB_MORTON_ISO")) ctx->mortonIso = atoi(e);
    if (const char* e = getenv("PB_NP_FUSE")) ctx->npFuseSmall = atoi(e);
    if (const char* e = getenv("PB_FUSED_LOCAL_MAX")) ctx->fusedLocalMax = atoi(e);
    if (const char* e = getenv("PB_DETERMINISTIC")) ctx->deterministic = atoi(e) != 0;
    if (const char* e = getenv("PB_ISLAND_LOCAL_MAX")) ctx->islandLocalMax = atoi(e) > 0 ? atoi(e) : 1;
    if (rc) { std::string e = ctx->err; pb_ctx_destroy(ctx); return rc; }
    cudaMemsetAsync(ctx->counters, 0, sizeof(int) * CNT_TOTAL, ctx->stream);
    *out = ctx;
    return PB_OK;
}

// Grow the per-step arenas of a live context.  Everything sized by max_pairs / max_manifolds is scratch that a step
// rewrites from the start, EXCEPT the previous step's contact-cache payload (point arrays + hash table of buffer curBuf),
// which is copied / re-hashed so a step that overflowed can simply be run again.
int pb_grow_arenas(pb_ctx* ctx, int maxPairs, int maxManifolds) {
    cudaSetDevice(ctx->device);
    pb_collect_step(ctx);        // (an overflow it reports is what the caller is here for)
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t oldM = ctx->caps.max_manifolds;
    const size_t P = std::max(maxPairs, ctx->caps.max_pairs), M = std::max((size_t)maxManifolds, oldM), C = ctx->caps.max_colliders;
    int rc = 0;
#define A(p, n) if (!rc) rc = pb_alloc(ctx, &ctx->p, (n))
    if ((int)P > ctx->caps.max_pairs) { A(pairs, P); A(pairOrder, 2 * P); A(trigPairs, P); }
    if (M > oldM) {
        size_t sortMax = std::max(C, M);
        ctx->radixTiles = (int)((sortMax + 511) / 512);
        A(radixHist, (size_t)256 * ctx->radixTiles + (size_t)256 * ctx->radixTiles / 4096 + 1024);
        A(mKey, M); A(mNormal, M); A(mPts, 8 * M); A(mSortKeyA, M); A(mSortKeyB, M); A(mSortValB, M);
        A(cHead, M); A(cBodies, M); A(cRowsT, M); A(cNormal, M); A(cSoft, M);
        const size_t PT = 4 * M, oldPT = 4 * oldM;
        A(pR1, PT); A(rowA, PT); A(rowB, PT); A(rowC, PT); A(rowD, PT); A(rowE, PT); A(rowF, PT); A(rowG, PT); A(rowL, PT);
        const int keep = ctx->curBuf, other = keep ^ 1;
        A(pR0T[other], PT); A(cPointOfsBuf[other], M + 1); A(cNpBuf[other], M + 1);
        // previous-step payload: allocate, copy, swap in
        float4* nR = nullptr; int* nOfs = nullptr; int* nNp = nullptr;
        if (!rc) rc = pb_alloc(ctx, &nR, PT);
        if (!rc) rc = pb_alloc(ctx, &nOfs, M + 1);
        if (!rc) rc = pb_alloc(ctx, &nNp, M + 1);
        size_t cs = 1; while (cs < 2 * M) cs <<= 1;
        unsigned long long* nTag[2] = { nullptr, nullptr }; int4* nVal[2] = { nullptr, nullptr };
        for (int b = 0; b < 2 && !rc; ++b) { rc = pb_alloc(ctx, &nTag[b], cs); if (!rc) rc = pb_alloc(ctx, &nVal[b], cs); }
        if (rc) {       // allocation failed half way: the live context keeps its old arenas
            cudaFree(nR); cudaFree(nOfs); cudaFree(nNp);
            for (int b = 0; b < 2; ++b) { cudaFree(nTag[b]); cudaFree(nVal[b]); }
            return rc;
        }
        PB_CUDA(ctx, cudaMemcpyAsync(nR, ctx->pR0T[keep], sizeof(float4) * oldPT, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(nOfs, ctx->cPointOfsBuf[keep], sizeof(int) * (oldM + 1), cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(nNp, ctx->cNpBuf[keep], sizeof(int) * (oldM + 1), cudaMemcpyDeviceToDevice, ctx->stream));
        pb_contact_cache_rehash(ctx, ctx->cacheSize, ctx->cacheTag[keep], ctx->cacheVal[keep], (int)cs, nTag[keep], nVal[keep]);
        PB_CUDA(ctx, cudaMemsetAsync(nTag[other], 0, sizeof(unsigned long long) * cs, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->pR0T[keep]); cudaFree(ctx->cPointOfsBuf[keep]); cudaFree(ctx->cNpBuf[keep]);
        ctx->pR0T[keep] = nR; ctx->cPointOfsBuf[keep] = nOfs; ctx->cNpBuf[keep] = nNp;
        for (int b = 0; b < 2; ++b) { cudaFree(ctx->cacheTag[b]); cudaFree(ctx->cacheVal[b]); ctx->cacheTag[b] = nTag[b]; ctx->cacheVal[b] = nVal[b]; }
        ctx->cacheSize = (int)cs;
    }
#undef A
    if (rc) return rc;
    ctx->caps.max_pairs = (int)P;
    ctx->caps.max_manifolds = (int)M;
    // the solve-order taps pointed into buffers that may just have been replaced
    ctx->mSorted = nullptr; ctx->mSortedKeys = nullptr; ctx->lastCounts.n_manifolds = 0; ctx->countsStale = false;
    return PB_OK;
}

void pb_ctx_destroy(pb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->readStream) { cudaStreamSynchronize(ctx->readStream); cudaStreamDestroy(ctx->readStream); }
    if (ctx->sideStream) { cudaStreamSynchronize(ctx->sideStream); cudaStreamDestroy(ctx->sideStream); }
    for (cudaEvent_t e : { ctx->evFork, ctx->evColour, ctx->evGroups, ctx->evJoints }) if (e) cudaEventDestroy(e);
    if (ctx->evPacked) cudaEventDestroy(ctx->evPacked);
    if (ctx->stageRead) cudaFree(ctx->stageRead);
    pb_joints_free(ctx);
    if (ctx->stageVel) cudaFree(ctx->stageVel);
    if (ctx->evMainAtSet) cudaEventDestroy(ctx->evMainAtSet);
    if (ctx->evVelReady) cudaEventDestroy(ctx->evVelReady);
    if (ctx->evPoseReady) cudaEventDestroy(ctx->evPoseReady);
    if (ctx->evCounters) cudaEventDestroy(ctx->evCounters);
    for (auto& e : ctx->evRead) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evReadPose) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evSub) if (e) cudaEventDestroy(e);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
#define F(p) if (ctx->p) cudaFree(ctx->p)
    F(rowEntity); F(pos); F(quat); F(velBuf[0]); F(velBuf[1]); F(bodyRec); F(comInvMass); F(invIL);
    F(kinematic); F(pseudoLin); F(pseudoAng); F(colorMask); F(rowMark); F(stage);
    F(colRow); F(colIndex); F(colType); F(colFlags); F(colInfo); F(colData); F(colMesh); F(colLPos); F(colLQuat); F(colParams); F(colMat); F(colWPos);
    F(colWQuat); F(aabbMin); F(aabbMax); F(keptMin); F(keptMax); F(mortonA); F(mortonB); F(leafIdA); F(leafIdB); F(radixHist); F(sceneBounds); F(bigList); F(sortBarrier);
    F(nodeLeft); F(nodeRight); F(nodeParent); F(leafParent); F(nodeFlag); F(nodeRange); F(nodeMin); F(nodeMax); F(pairs); F(pairOrder);
    F(mKey); F(mNormal); F(mPts); F(mSortKeyA); F(mSortKeyB); F(mSortValB); F(cHead); F(cBodies); F(cRowsT); F(cNormal); F(cSoft);
    F(pR0T[0]); F(pR0T[1]); F(cPointOfsBuf[0]); F(cPointOfsBuf[1]); F(cNpBuf[0]); F(cNpBuf[1]); F(pR1);
    F(rowA); F(rowB); F(rowC); F(rowD); F(rowE); F(rowF); F(rowG); F(rowL);
    F(cacheTag[0]); F(cacheTag[1]); F(cacheVal[0]); F(cacheVal[1]); F(counters); F(triMeshDev); F(convexDev); F(nonColliding); F(trigPairs); F(colClass); F(filterLut); F(solveBarrier); F(solveProfNs); F(queryOut);
    F(jpBest); F(jpScratch); F(bodyOrder); F(bodyStart); F(bodyCursor); F(gjkHitPair); F(gjkHitSimplex); F(spillList); F(spillEpa); F(spillMesh); F(qCounters); F(qPairs); F(qPairOrder); F(qmKey); F(qmNormal); F(qmPts);
    F(islandParent); F(islandCount); F(bodyGroup); F(islandStats); F(keyStart); F(keyCursor); F(jointKey); F(jointStart); F(jointSortTmp[0]); F(jointSortTmp[1]); F(jointSortTmp[2]);
#undef F
    for (auto& m : ctx->triMeshes) { cudaFree(m.verts); cudaFree(m.tris); cudaFree(m.triNormal); cudaFree(m.triCentroid); cudaFree(m.triRec); cudaFree(m.nodeMin); cudaFree(m.nodeMax); }
    for (auto& m : ctx->convexes) { cudaFree(m.verts); cudaFree(m.faceOffsets); cudaFree(m.faceIndices); cudaFree(m.faceNormal); cudaFree(m.faceCentroid); }
    if (ctx->hCounters) cudaFreeHost(ctx->hCounters);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pb_sync(pb_ctx* ctx) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->readStream));      // a read-back begun with pb_get_state_begin has landed, too
    return pb_collect_step(ctx);
}

int pb_upload_bodies(pb_ctx* ctx, int nDyn, int nStatic, const int* entity, const float* pos3, const float* quat4, const int* kinematic,
                     const float* vel3, const float* angvel3, const float* invMass, const float* com3, const float* invI9) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    int rows = nDyn + nStatic;
    if (rows > ctx->caps.max_bodies) return pb_fail(ctx, PB_ECAPACITY, "max_bodies");
    ctx->nDyn = nDyn; ctx->nStatic = nStatic; ctx->nRows = rows;
    ctx->hRowEntity.assign(entity, entity + rows);
    int rc;
    if ((rc = uploadRaw(ctx, entity, rows, ctx->rowEntity))) return rc;
    if ((rc = uploadVec(ctx, pos3, rows, 3, ctx->pos))) return rc;
    if ((rc = uploadVec(ctx, quat4, rows, 4, ctx->quat))) return rc;
    ctx->hKinematic.assign(kinematic, kinematic + nDyn);
    if (nDyn) {
        if ((rc = uploadRaw(ctx, kinematic, nDyn, ctx->kinematic))) return rc;
        size_t bytes = sizeof(float) * (size_t)nDyn * 19;
        if ((rc = ensureStage(ctx, bytes))) return rc;
        float* s = ctx->stage;
        const size_t n = (size_t)nDyn;
        PB_CUDA(ctx, cudaMemcpyAsync(s, com3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(s + 3 * n, invMass, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(s + 4 * n, invI9, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(s + 13 * n, vel3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(s + 16 * n, angvel3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
        ++ctx->launches, k_com_invmass<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, s, s + 3 * n, ctx->comInvMass);
        ++ctx->launches, k_unpack_m3<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, s + 4 * n, ctx->invIL);
        ++ctx->launches, k_unpack_vel<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, s + 13 * n, s + 16 * n, ctx->comInvMass, ctx->vel);
    }
    ctx->cacheValid = false;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_upload_colliders(pb_ctx* ctx, int n, const int* bodyRow, const int* colIndex, const float* lpos3, const float* lquat4, const int* type,
                        const float* params4, const int* mesh, const float* material3, const int* flags, const int* data) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    if (n > ctx->caps.max_colliders) return pb_fail(ctx, PB_ECAPACITY, "max_colliders");
    ctx->nCol = n;
    ctx->hColType.assign(type, type + n); ctx->hColMesh.assign(mesh, mesh + n); ctx->hColRow.assign(bodyRow, bodyRow + n);
    ctx->hColIndex.assign(colIndex, colIndex + n);
    ctx->hTrimeshCols.clear();
    for (int i = 0; i < n; ++i) if (type[i] == PB_TRIANGLE_MESH) ctx->hTrimeshCols.push_back(i);
    // a custom filter table is indexed by collider: it has to be set again after every collider upload
    if (ctx->filterLut) { cudaFree(ctx->filterLut); ctx->filterLut = nullptr; ctx->nFilterClasses = 0; }
    // device flags: bit0 trigger, bit1 enableSimulation, bit2 owner is a non-kinematic dynamic body (BroadPhaseEntry::isDynamic)
    const std::vector<int>& kin = ctx->hKinematic;
    std::vector<int> f(n);
    std::vector<float> mat4((size_t)4 * n);
    bool anyRest = false;      // restitution of a pair = mean of the two colliders' (contacts.cu k_contact_build): zero everywhere -> the contact cache is idle
    bool trig = false;
    for (int i = 0; i < n; ++i) {
        int row = bodyRow[i];
        if (row < 0 || row >= ctx->nRows) return pb_fail(ctx, PB_EINVAL, "collider body_row out of range");
        bool dyn = row < ctx->nDyn && !kin[row];
        f[i] = (flags[i] & 3) | (dyn ? 4 : 0);
        trig |= (flags[i] & PB_COL_TRIGGER) != 0;
        if (type[i] == PB_TRIANGLE_MESH && (mesh[i] < 0 || mesh[i] >= (int)ctx->triMeshes.size())) return pb_fail(ctx, PB_EINVAL, "bad trimesh handle");
        if (type[i] == PB_CONVEX_MESH && (mesh[i] < 0 || mesh[i] >= (int)ctx->convexes.size())) return pb_fail(ctx, PB_EINVAL, "bad convex handle");
        mat4[4 * i] = material3[3 * i]; mat4[4 * i + 1] = material3[3 * i + 1]; mat4[4 * i + 2] = material3[3 * i + 2]; mat4[4 * i + 3] = 0.f;
        if (material3[3 * i + 1] != 0.f) anyRest = true;
    }
    ctx->anyTriggerFlag = trig;
    if (!ctx->filterLut) ctx->triggersPossible = trig;
    int rc;
    if ((rc = uploadRaw(ctx, bodyRow, n, ctx->colRow))) return rc;
    if ((rc = uploadRaw(ctx, colIndex, n, ctx->colIndex))) return rc;
    if ((rc = uploadRaw(ctx, type, n, ctx->colType))) return rc;
    if ((rc = uploadRaw(ctx, f.data(), n, ctx->colFlags))) return rc;
    if ((rc = uploadRaw(ctx, data, n, ctx->colData))) return rc;
    if ((rc = uploadRaw(ctx, mesh, n, ctx->colMesh))) return rc;
    if ((rc = uploadVec(ctx, lpos3, n, 3, ctx->colLPos))) return rc;
    if ((rc = uploadVec(ctx, lquat4, n, 4, ctx->colLQuat))) return rc;
    if ((rc = uploadVec(ctx, params4, n, 4, ctx->colParams))) return rc;
    if ((rc = uploadVec(ctx, mat4.data(), n, 4, ctx->colMat))) return rc;
    ctx->anyRestitution = anyRest;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    // creation-time bounds: no margin (Physecs.cpp:32)
    if ((rc = pb_update_bounds_all(ctx, 0.f, 0))) return rc;
    ctx->cacheValid = false;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

static int syncMeshTables(pb_ctx* ctx) {
    std::vector<PbTriMeshDev> t(ctx->triMeshes.size());
    for (size_t i = 0; i < t.size(); ++i) {
        auto& m = ctx->triMeshes[i];
        t[i] = { m.verts, m.tris, m.triNormal, m.triCentroid, m.nodeMin, m.nodeMax, m.nTris, m.nNodes,
                 { m.bmin[0], m.bmin[1], m.bmin[2] }, { m.bmax[0], m.bmax[1], m.bmax[2] }, m.triRec };
    }
    std::vector<PbConvexDev> c(ctx->convexes.size());
    for (size_t i = 0; i < c.size(); ++i) {
        auto& m = ctx->convexes[i];
        c[i] = { m.verts, m.faceOffsets, m.faceIndices, m.faceNormal, m.faceCentroid, m.nVerts, m.nVertsPadded, m.nFaces, m.maxFaceVerts };
    }
    if (t.size() > 64 || c.size() > 256) return pb_fail(ctx, PB_ECAPACITY, "too many meshes");
    if (!t.empty()) PB_CUDA(ctx, cudaMemcpy(ctx->triMeshDev, t.data(), sizeof(PbTriMeshDev) * t.size(), cudaMemcpyHostToDevice));
    if (!c.empty()) PB_CUDA(ctx, cudaMemcpy(ctx->convexDev, c.data(), sizeof(PbConvexDev) * c.size(), cudaMemcpyHostToDevice));
    return PB_OK;
}

int pb_register_convex(pb_ctx* ctx, const float* verts3, int nVerts, const int* faceOffsets, const int* faceIndices, int nFaces,
                       const float* faceNormals3, const float* faceCentroids3, int* handle) {
    cudaSetDevice(ctx->device);
    PbConvex m;
    m.nVerts = nVerts; m.nVertsPadded = (nVerts + 3) / 4 * 4; m.nFaces = nFaces;
    std::vector<float4> v(m.nVertsPadded);
    for (int i = 0; i < m.nVertsPadded; ++i) {
        int s = i < nVerts ? i : nVerts - 1;   // padded by repeating the last vertex (ConvexMesh.cpp:8-10)
        v[i] = make_float4(verts3[3 * s], verts3[3 * s + 1], verts3[3 * s + 2], 0.f);
    }
    std::vector<float4> fn(nFaces), fc(nFaces);
    for (int i = 0; i < nFaces; ++i) {
        fn[i] = make_float4(faceNormals3[3 * i], faceNormals3[3 * i + 1], faceNormals3[3 * i + 2], 0.f);
        fc[i] = make_float4(faceCentroids3[3 * i], faceCentroids3[3 * i + 1], faceCentroids3[3 * i + 2], 0.f);
    }
    int nIdx = faceOffsets[nFaces];
    for (int i = 0; i < nFaces; ++i) m.maxFaceVerts = std::max(m.maxFaceVerts, faceOffsets[i + 1] - faceOffsets[i]);
    int rc = 0;
    if (!rc) rc = pb_alloc(ctx, &m.verts, v.size());
    if (!rc) rc = pb_alloc(ctx, &m.faceOffsets, (size_t)nFaces + 1);
    if (!rc) rc = pb_alloc(ctx, &m.faceIndices, (size_t)nIdx);
    if (!rc) rc = pb_alloc(ctx, &m.faceNormal, (size_t)nFaces);
    if (!rc) rc = pb_alloc(ctx, &m.faceCentroid, (size_t)nFaces);
    if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpy(m.verts, v.data(), sizeof(float4) * v.size(), cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.faceOffsets, faceOffsets, sizeof(int) * (nFaces + 1), cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.faceIndices, faceIndices, sizeof(int) * nIdx, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.faceNormal, fn.data(), sizeof(float4) * nFaces, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.faceCentroid, fc.data(), sizeof(float4) * nFaces, cudaMemcpyHostToDevice));
    ctx->convexes.push_back(m);
    *handle = (int)ctx->convexes.size() - 1;
    return syncMeshTables(ctx);
}

int pb_register_trimesh(pb_ctx* ctx, const float* verts3, int nVerts, const unsigned* indices, int nIndices, int* handle, int* triOrderOut) {
    cudaSetDevice(ctx->device);
    PbHostTriMesh h;
    pb_build_trimesh_host(verts3, nVerts, indices, nIndices, h);
    PbTriMesh m;
    m.nVerts = nVerts; m.nTris = h.nTris; m.nNodes = h.nNodes;
    std::vector<float4> v(nVerts), tn(h.nTris), tc(h.nTris), nmn(h.nNodes), nmx(h.nNodes);
    std::vector<int4> ti(h.nTris);
    std::vector<float4> rec(4 * (size_t)h.nTris);
    for (int k = 0; k < 3; ++k) { m.bmin[k] = 3.4e38f; m.bmax[k] = -3.4e38f; }
    for (int i = 0; i < nVerts; ++i) {
        v[i] = make_float4(verts3[3 * i], verts3[3 * i + 1], verts3[3 * i + 2], 0.f);
        for (int k = 0; k < 3; ++k) { m.bmin[k] = std::min(m.bmin[k], verts3[3 * i + k]); m.bmax[k] = std::max(m.bmax[k], verts3[3 * i + k]); }
    }
    for (int i = 0; i < h.nTris; ++i) {
        ti[i] = make_int4((int)h.triIdx[3 * i], (int)h.triIdx[3 * i + 1], (int)h.triIdx[3 * i + 2], 0);
        tn[i] = make_float4(h.triNormal[3 * i], h.triNormal[3 * i + 1], h.triNormal[3 * i + 2], 0.f);
        tc[i] = make_float4(h.triCentroid[3 * i], h.triCentroid[3 * i + 1], h.triCentroid[3 * i + 2], 0.f);
        for (int k = 0; k < 3; ++k) {
            const float4& vk = v[h.triIdx[3 * i + k]];
            rec[4 * (size_t)i + k] = make_float4(vk.x, vk.y, vk.z, h.triNormal[3 * i + k]);
        }
        float ix[3];
        for (int k = 0; k < 3; ++k) { int id = (int)h.triIdx[3 * i + k]; memcpy(&ix[k], &id, 4); }
        rec[4 * (size_t)i + 3] = make_float4(ix[0], ix[1], ix[2], 0.f);
    }
    for (int i = 0; i < h.nNodes; ++i) {
        int cnt = h.nodeCountIndex[2 * i], idx = h.nodeCountIndex[2 * i + 1];
        float fc, fi; memcpy(&fc, &cnt, 4); memcpy(&fi, &idx, 4);
        nmn[i] = make_float4(h.nodeBounds[6 * i], h.nodeBounds[6 * i + 1], h.nodeBounds[6 * i + 2], fc);
        nmx[i] = make_float4(h.nodeBounds[6 * i + 3], h.nodeBounds[6 * i + 4], h.nodeBounds[6 * i + 5], fi);
    }
    int rc = 0;
    if (!rc) rc = pb_alloc(ctx, &m.verts, (size_t)nVerts);
    if (!rc) rc = pb_alloc(ctx, &m.tris, (size_t)h.nTris);
    if (!rc) rc = pb_alloc(ctx, &m.triNormal, (size_t)h.nTris);
    if (!rc) rc = pb_alloc(ctx, &m.triCentroid, (size_t)h.nTris);
    if (!rc) rc = pb_alloc(ctx, &m.triRec, 4 * (size_t)h.nTris);
    if (!rc) rc = pb_alloc(ctx, &m.nodeMin, (size_t)h.nNodes);
    if (!rc) rc = pb_alloc(ctx, &m.nodeMax, (size_t)h.nNodes);
    if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpy(m.verts, v.data(), sizeof(float4) * nVerts, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.tris, ti.data(), sizeof(int4) * h.nTris, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.triNormal, tn.data(), sizeof(float4) * h.nTris, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.triCentroid, tc.data(), sizeof(float4) * h.nTris, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.triRec, rec.data(), sizeof(float4) * 4 * (size_t)h.nTris, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.nodeMin, nmn.data(), sizeof(float4) * h.nNodes, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(m.nodeMax, nmx.data(), sizeof(float4) * h.nNodes, cudaMemcpyHostToDevice));
    ctx->triMeshes.push_back(m);
    *handle = (int)ctx->triMeshes.size() - 1;
    if (triOrderOut) memcpy(triOrderOut, h.triOrig.data(), sizeof(int) * h.nTris);
    return syncMeshTables(ctx);
}

int pb_build_trimesh(const float* verts3, int nVerts, const unsigned* indices, int nIndices, unsigned* triIdx, int* triOrig,
                     float* nodeBounds6, int* nodeCountIndex2, int* nNodes) {
    PbHostTriMesh h;
    pb_build_trimesh_host(verts3, nVerts, indices, nIndices, h);
    if (triIdx) memcpy(triIdx, h.triIdx.data(), sizeof(unsigned) * h.triIdx.size());
    if (triOrig) memcpy(triOrig, h.triOrig.data(), sizeof(int) * h.triOrig.size());
    if (nodeBounds6) memcpy(nodeBounds6, h.nodeBounds.data(), sizeof(float) * h.nodeBounds.size());
    if (nodeCountIndex2) memcpy(nodeCountIndex2, h.nodeCountIndex.data(), sizeof(int) * h.nodeCountIndex.size());
    if (nNodes) *nNodes = h.nNodes;
    return PB_OK;
}

int pb_upload_joints(pb_ctx* ctx, int n, const int* type, const int* row0, const int* row1, const float* a0p, const float* a0q,
                     const float* a1p, const float* a1q, const float* params8, const int* color) {
    cudaSetDevice(ctx->device);
    return pb_joints_upload(ctx, n, type, row0, row1, a0p, a0q, a1p, a1q, params8, color);
}

int pb_update_joint_params(pb_ctx* ctx, int n, const float* params8) {
    cudaSetDevice(ctx->device);
    return pb_joints_update_params(ctx, n, params8);
}

int pb_set_noncolliding_pairs(pb_ctx* ctx, int n, const int* pairs2) {
    cudaSetDevice(ctx->device);
    std::vector<unsigned long long> keys(n);
    for (int i = 0; i < n; ++i) {
        unsigned int a = (unsigned int)pairs2[2 * i], b = (unsigned int)pairs2[2 * i + 1];
        if (a > b) std::swap(a, b);
        keys[i] = ((unsigned long long)a << 32) | b;
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    int rc = pb_alloc(ctx, &ctx->nonColliding, keys.size());
    if (rc) return rc;
    ctx->nNonColliding = (int)keys.size();
    if (!keys.empty()) PB_CUDA(ctx, cudaMemcpy(ctx->nonColliding, keys.data(), sizeof(unsigned long long) * keys.size(), cudaMemcpyHostToDevice));
    return PB_OK;
}

// rows [first, first + count) of the dynamic state; the pointers address the first row of the range
static int setStateRange(pb_ctx* ctx, int first, int count, const float* pos3, const float* quat4, const float* vel3, const float* angvel3) {
    ctx->queryTreeValid = false;
    if (count <= 0) return PB_OK;
    // One packed H2D burst (13 floats / body) on the COPY stream, unpacked into the float4 SoA there: ordered after everything
    // already queued on the main stream (or after the mark pb_step_begin left), and every reader on the main stream waits for
    // evPoseReady / evVelReady (pb_ctx.h).
    const size_t N = (size_t)ctx->nDyn, f = (size_t)first, n = (size_t)count;
    if (ctx->stageVelBytes < sizeof(float) * 13 * N) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream));
        if (ctx->stageVel) cudaFree(ctx->stageVel);
        ctx->stageVel = nullptr; ctx->stageVelBytes = 0;
        PB_CUDA(ctx, cudaMalloc((void**)&ctx->stageVel, sizeof(float) * 13 * N));
        ctx->stageVelBytes = sizeof(float) * 13 * N;
    }
    float* s = ctx->stageVel;
    cudaStream_t cs = ctx->copyStream;
    if (!ctx->mainMarked) PB_CUDA(ctx, cudaEventRecord(ctx->evMainAtSet, ctx->stream));
    PB_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->evMainAtSet, 0));
    int g = pb_grid(count, 256);
    if (pos3 || quat4) {
        if (pos3) PB_CUDA(ctx, cudaMemcpyAsync(s + 3 * f, pos3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs));
        if (quat4) PB_CUDA(ctx, cudaMemcpyAsync(s + 3 * N + 4 * f, quat4, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, cs));
        if (pos3) ++ctx->launches, k_unpack3<<<g, 256, 0, cs>>>(count, s + 3 * f, ctx->pos + f);
        if (quat4) ++ctx->launches, k_unpack4<<<g, 256, 0, cs>>>(count, s + 3 * N + 4 * f, ctx->quat + f);
        PB_CUDA(ctx, cudaEventRecord(ctx->evPoseReady, cs));
        ctx->posePending = true;
    }
    if (vel3 || angvel3) {
        if (vel3) PB_CUDA(ctx, cudaMemcpyAsync(s + 7 * N + 3 * f, vel3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs));
        if (angvel3) PB_CUDA(ctx, cudaMemcpyAsync(s + 10 * N + 3 * f, angvel3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs));
        ++ctx->launches, k_unpack_vel<<<g, 256, 0, cs>>>(count, vel3 ? s + 7 * N + 3 * f : nullptr, angvel3 ? s + 10 * N + 3 * f : nullptr, ctx->comInvMass + f, ctx->vel + 2 * f);
        PB_CUDA(ctx, cudaEventRecord(ctx->evVelReady, cs));
        ctx->velPending = true;
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_set_state(pb_ctx* ctx, int nDyn, const float* pos3, const float* quat4, const float* vel3, const float* angvel3) {
    cudaSetDevice(ctx->device);
    if (nDyn != ctx->nDyn) return pb_fail(ctx, PB_EINVAL, "pb_set_state: n_dynamic mismatch");
    return setStateRange(ctx, 0, nDyn, pos3, quat4, vel3, angvel3);
}

int pb_set_state_rows(pb_ctx* ctx, int first, int count, const float* pos3, const float* quat4, const float* vel3, const float* angvel3) {
    cudaSetDevice(ctx->device);
    if (first < 0 || count < 0 || first + count > ctx->nDyn) return pb_fail(ctx, PB_EINVAL, "pb_set_state_rows: range outside the dynamic rows");
    return setStateRange(ctx, first, count, pos3, quat4, vel3, angvel3);
}

int pb_move_rows(pb_ctx* ctx, int n, const int* rows, const float* pos3, const float* quat4) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    if (n <= 0) return PB_OK;
    size_t bytes = sizeof(float) * 8 * (size_t)n;
    int rc = ensureStage(ctx, bytes); if (rc) return rc;
    float* s = ctx->stage;
    PB_CUDA(ctx, cudaMemsetAsync(ctx->rowMark, 0, sizeof(int) * ctx->nRows, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s, rows, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + n, pos3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + 4 * (size_t)n, quat4, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->launches, k_scatter_rows<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, (const int*)s, s + n, s + 4 * (size_t)n, ctx->pos, ctx->quat, ctx->rowMark);
    if ((rc = pb_update_bounds_rows(ctx, ctx->rowMark, n, 0.01f))) return rc;
    std::vector<char> moved(ctx->nRows, 0);
    for (int i = 0; i < n; ++i) if (rows[i] >= 0 && rows[i] < ctx->nRows) moved[rows[i]] = 1;
    for (int c : ctx->hTrimeshCols)
        if (moved[ctx->hColRow[c]]) pb_update_bounds_trimesh_col(ctx, c, 0.01f);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_refresh_bounds(pb_ctx* ctx) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    return pb_update_bounds_all(ctx, 0.01f, 1);
}

static int readCounters(pb_ctx* ctx) {
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->hCounters, ctx->counters, sizeof(int) * CNT_TOTAL, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// Local (per-CTA) sweeps pay off when a worthwhile share of the constraints sits in small islands; finding the islands costs a few
// kernels per step, so in auto mode a scene that turned out to be one big pile is only looked at again every 64 steps.
static void chooseIslands(pb_ctx* ctx) {
    const bool wasOn = ctx->islandsOn;
    if (ctx->islandsMode == 0) { ctx->islandsOn = false; return; }
    if (ctx->islandsMode == 1) { ctx->islandsOn = true; return; }
    if (wasOn && ctx->lastIslandTotal > 0 && 2ll * ctx->lastIslandLocal < ctx->lastIslandTotal) { ctx->islandsOn = false; ctx->islandsHold = 63; return; }
    if (!wasOn && ctx->islandsHold > 0) { --ctx->islandsHold; return; }
    ctx->islandsOn = true;
}

// The outcome of the last enqueued pb_step.  pb_step never waits for the device: the arena checks run there (a step whose
// narrowphase overflowed skips its solve, integration and bounds refresh -- solver.cu stepSkipped), the counters are copied to
// pinned memory right after the narrowphase, and the host looks at them here, at the next call that needs the step to have
// happened (pb_sync, pb_get_*, the taps, the next pb_step).  On overflow the host-side bookkeeping of the step is rolled back, so
// the scene is exactly what it was before the failed step: grow the arenas (pb_grow_arenas) and call pb_step again.
int pb_collect_step(pb_ctx* ctx) {
    if (!ctx->stepPending) return PB_OK;
    ctx->stepPending = false;
    PB_CUDA(ctx, cudaEventSynchronize(ctx->evCounters));
    const int* h = ctx->hCounters;
    const int nPairs = h[CNT_PAIRS], nRaw = h[CNT_RAWM], status = h[CNT_STATUS], cause = h[CNT_CAUSE];
    if (ctx->statsCopied) { ctx->lastIslandLocal = h[CNT_TOTAL]; ctx->lastIslandTotal = h[CNT_TOTAL + 1]; }    // of the step before (rode along)
    ctx->lastCounts = pb_counts{};
    ctx->lastCounts.n_pairs = nPairs;
    ctx->lastCounts.n_mesh_pairs = h[CNT_MESH_PAIRS];
    ctx->lastCounts.n_triggers = h[CNT_TRIGGERS];
    ctx->lastCounts.cause = cause;
    ctx->lastCounts.n_spilled = h[CNT_SPILLED];
    if (ctx->pendingTiles) { ctx->lastTileHits = h[CNT_TILE_HITS]; ctx->lastTiles = ctx->pendingTiles; }      // tile statistics: the next step's broadphase choice
    ctx->pairsHint = std::min(nPairs, ctx->caps.max_pairs);
    ctx->rawHint = std::min(nRaw, ctx->caps.max_manifolds);
    const bool overflow = nPairs > ctx->caps.max_pairs || nRaw > ctx->caps.max_manifolds || (status & PB_ECAPACITY);
    if (overflow || (status & 0x100)) {
        // roll the host side back: the double buffers the build wrote into become "other" again, the velocity pointers return to the
        // buffer that still holds the pre-step velocities (the device skipped every kernel that writes scene state)
        ctx->curBuf ^= 1;
        ctx->stepNarrowed = false; ctx->stepBegun = false;      // (collected between pb_step_narrowphase and pb_step: that step starts over)
        if (ctx->undoVelSwaps & 1) { std::swap(ctx->vel, ctx->velLive); std::swap(ctx->angvel, ctx->angvelLive); }
        ctx->cacheValid = ctx->undoCacheValid; ctx->cacheBuilt = ctx->undoCacheBuilt;
        ctx->countsStale = false;             // nothing was built: a later pb_get_counts must not replace n_manifolds by the post-build counter
        ctx->mSorted = nullptr; ctx->mSortedKeys = nullptr;
        if (!overflow) return pb_fail(ctx, PB_EUNSUPPORTED, "a candidate pair involves a shape combination not implemented on the device path (the step was not applied)");
        ctx->lastCounts.status = PB_ECAPACITY;
        ctx->lastCounts.n_manifolds = nRaw;   // what the arenas would have needed (counters keep counting past the capacity)
        std::string why;
        if (nPairs > ctx->caps.max_pairs || (cause & PB_CAUSE_PAIRS)) why += " candidate pairs";
        if (nRaw > ctx->caps.max_manifolds || (cause & PB_CAUSE_MANIFOLDS)) why += " manifolds";
        if (cause & PB_CAUSE_TRIGGERS) why += " trigger pairs";
        if (cause & PB_CAUSE_WALK_STACK) why += " broadphase walk stack";
        char hex[16]; snprintf(hex, sizeof hex, "%x", cause);
        return pb_fail(ctx, PB_ECAPACITY, "per-step arena overflow (" + (why.empty() ? std::string(" unknown") : why) + " ): pairs=" + std::to_string(nPairs) + "/" +
                       std::to_string(ctx->caps.max_pairs) + " manifolds=" + std::to_string(nRaw) + "/" + std::to_string(ctx->caps.max_manifolds) +
                       " triggers=" + std::to_string(h[CNT_TRIGGERS]) + " cause=0x" + hex + "; the step was not applied: pb_grow_arenas, then pb_step again");
    }
    ctx->countsStale = true;     // manifold / colour / point counts stay on the device until someone asks (pb_get_counts)
    return PB_OK;
}

// The head of a step that needs nothing from the host: counters reset + broadphase.  The reference's sweep runs on the bounds the
// previous simulate left behind (Physecs.cpp:119-173 reads BroadPhaseEntry::bounds, refreshed at :556-559 and by registry.patch), not
// on the transforms of this step -- so it can run while the host still gathers / uploads the new state.
int pb_step_begin(pb_ctx* ctx) {
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = pb_collect_step(ctx))) return rc;          // the PREVIOUS step overflowed and was not applied: nothing new is enqueued
    ctx->queryTreeValid = false;
    // uploads issued from here on (pb_set_state / pb_set_state_rows, copy stream) wait for what the main stream holds NOW -- the
    // previous step -- and not for the broadphase below, which reads neither poses nor velocities
    PB_CUDA(ctx, cudaEventRecord(ctx->evMainAtSet, ctx->stream));
    ctx->mainMarked = true;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    PB_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(int) * CNT_TOTAL, ctx->stream));
    if ((rc = pb_broadphase(ctx))) return rc;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->stepBegun = true;
    return PB_OK;
}

// the middle of a step: world poses + narrowphase behind the broadphase, the counters on their way to the host
static int stepNarrow(pb_ctx* ctx) {
    int rc;
    if ((rc = pb_wait_poses(ctx))) return rc;          // the broadphase above ran on the bounds of the previous step's end; from here on poses are read
    if ((rc = pb_world_poses(ctx))) return rc;
    if ((rc = pb_narrowphase(ctx))) return rc;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    // counters of the narrowphase (pairs, raw manifolds, status, triggers) to pinned memory; the island statistics of the PREVIOUS
    // step ride along.  Nobody waits here: pb_collect_step looks at them later.
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->hCounters, ctx->counters, sizeof(int) * CNT_TOTAL, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->statsCopied = ctx->islandsOn && ctx->islandStats;
    if (ctx->statsCopied) PB_CUDA(ctx, cudaMemcpyAsync(ctx->hCounters + CNT_TOTAL, ctx->islandStats, sizeof(int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaEventRecord(ctx->evCounters, ctx->stream));
    ctx->undoCacheValid = ctx->cacheValid; ctx->undoCacheBuilt = ctx->cacheBuilt; ctx->undoVelSwaps = 0;
    ctx->stepPending = true;
    ctx->curBuf ^= 1;
    ctx->stepNarrowed = true;
    return PB_OK;
}

// Optional middle of a step: everything that needs the new POSES but not the new velocities (world poses, narrowphase).  A caller
// that uploads poses first calls this, then uploads the velocities while the narrowphase runs, then pb_step.
int pb_step_narrowphase(pb_ctx* ctx) {
    cudaSetDevice(ctx->device);
    int rc;
    if (ctx->stepNarrowed) return PB_OK;
    if (!ctx->stepBegun && (rc = pb_step_begin(ctx))) return rc;
    return stepNarrow(ctx);
}

int pb_step(pb_ctx* ctx, float dt, int substeps, int iterations, float gravity) {
    cudaSetDevice(ctx->device);
    if (substeps < 1 || iterations < 0) return pb_fail(ctx, PB_EINVAL, "substeps/iterations");
    int rc;
    if (!ctx->stepNarrowed) {
        if (!ctx->stepBegun && (rc = pb_step_begin(ctx))) return rc;
        if ((rc = stepNarrow(ctx))) return rc;
    }
    ctx->stepBegun = false; ctx->mainMarked = false; ctx->stepNarrowed = false;
    if ((rc = pb_wait_velocities(ctx))) return rc;      // first readers of the velocities: contact build, then the solver
    chooseIslands(ctx);
    if ((rc = pb_joint_begin_step(ctx))) return rc;      // solver body indices of the joints: the island search hooks through them
    if ((rc = pb_contact_build(ctx))) return rc;
    cudaEventRecord(ctx->ev[3], ctx->stream);
    if ((rc = pb_solve(ctx, dt, substeps, iterations, gravity))) return rc;
    ctx->cacheValid = ctx->anyRestitution;      // (a scene without restitution leaves the cache tables alone: nothing valid to look up next step)
    ctx->cacheBuilt = ctx->anyRestitution;
    // bounds of every non-kinematic dynamic body for the next step, +0.01 margin (Physecs.cpp:556-559)
    if ((rc = pb_update_bounds_all(ctx, 0.01f, 1, ctx->counters + CNT_STATUS))) return rc;
    cudaEventRecord(ctx->ev[4], ctx->stream);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_get_state(pb_ctx* ctx, float* pos3, float* quat4, float* vel3, float* angvel3) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }      // the step before overflowed: its status instead of a state it did not produce
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    int nDyn = ctx->nDyn;
    if (!nDyn) return PB_OK;
    size_t n = (size_t)nDyn;
    int rc = ensureStage(ctx, sizeof(float) * 13 * n); if (rc) return rc;
    float* s = ctx->stage;
    int g = pb_grid(nDyn, 256);
    if (pos3) ++ctx->launches, k_pack3<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->pos, s);
    if (quat4) ++ctx->launches, k_pack4<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->quat, s + 3 * n);
    if (vel3 || angvel3)
        ++ctx->launches, k_pack_vel<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->vel, vel3 ? s + 7 * n : nullptr, angvel3 ? s + 10 * n : nullptr);
    if (pos3) PB_CUDA(ctx, cudaMemcpyAsync(pos3, s, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (quat4) PB_CUDA(ctx, cudaMemcpyAsync(quat4, s + 3 * n, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (vel3) PB_CUDA(ctx, cudaMemcpyAsync(vel3, s + 7 * n, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (angvel3) PB_CUDA(ctx, cudaMemcpyAsync(angvel3, s + 10 * n, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// The read-back in chunks: the pack kernels run once, then every chunk's four copies are followed by an event.  The caller scatters
// chunk c into its own data structures while chunk c + 1 is still crossing the bus (host/Scene.cpp).
int pb_get_state_begin(pb_ctx* ctx, float* pos3, float* quat4, float* vel3, float* angvel3, int nChunks) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    if (nChunks < 1) nChunks = 1;
    if (nChunks > PB_MAX_READ_CHUNKS) nChunks = PB_MAX_READ_CHUNKS;
    const int nDyn = ctx->nDyn;
    const int before = ctx->readChunks;
    ctx->readChunks = 0;
    if (!nDyn) return PB_OK;
    const size_t n = (size_t)nDyn;
    if (ctx->stageReadBytes < sizeof(float) * 13 * n) {
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->readStream));
        if (ctx->stageRead) cudaFree(ctx->stageRead);
        ctx->stageRead = nullptr; ctx->stageReadBytes = 0;
        PB_CUDA(ctx, cudaMalloc((void**)&ctx->stageRead, sizeof(float) * 13 * n));
        ctx->stageReadBytes = sizeof(float) * 13 * n;
    } else if (before > 0) {
        // the copies of the previous read-back leave the staging buffer before it is packed again
        PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evRead[before - 1], 0));
    }
    // The state is PACKED on the main stream (a snapshot in device memory: microseconds) and COPIED OUT on the read stream: whatever
    // the main stream is given next -- the upload and the kernels of the following step -- runs beside the PCIe transfer, and only
    // pb_get_state_wait waits for it.  A caller that reads step k's result after enqueueing step k + 1 hides the read-back entirely.
    float* s = ctx->stageRead;
    const int g = pb_grid(nDyn, 256);
    if (pos3) ++ctx->launches, k_pack3<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->pos, s);
    if (quat4) ++ctx->launches, k_pack4<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->quat, s + 3 * n);
    if (vel3 || angvel3)
        ++ctx->launches, k_pack_vel<<<g, 256, 0, ctx->stream>>>(nDyn, ctx->vel, vel3 ? s + 7 * n : nullptr, angvel3 ? s + 10 * n : nullptr);
    PB_CUDA(ctx, cudaEventRecord(ctx->evPacked, ctx->stream));
    cudaStream_t rs = ctx->readStream;
    PB_CUDA(ctx, cudaStreamWaitEvent(rs, ctx->evPacked, 0));
    int rc = PB_OK;
    // chunk by chunk (poses, then velocities of the chunk), or -- pb_set_readback_order -- the poses of every chunk first: whoever
    // sends the result back up for the next step needs the poses first (narrowphase) and the velocities a stage later (contact build)
    const int per = (nDyn + nChunks - 1) / nChunks;
    auto copyPoses = [&](int c, size_t f, size_t m) -> int {
        if (pos3) PB_CUDA(ctx, cudaMemcpyAsync(pos3 + 3 * f, s + 3 * f, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, rs));
        if (quat4) PB_CUDA(ctx, cudaMemcpyAsync(quat4 + 4 * f, s + 3 * n + 4 * f, sizeof(float) * 4 * m, cudaMemcpyDeviceToHost, rs));
        if (!ctx->evReadPose[c]) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evReadPose[c], cudaEventDisableTiming));
        PB_CUDA(ctx, cudaEventRecord(ctx->evReadPose[c], rs));
        ctx->readFirst[c] = (int)f; ctx->readCount[c] = (int)m;
        ctx->readChunks = c + 1;
        return PB_OK;
    };
    auto copyVels = [&](int c, size_t f, size_t m) -> int {
        if (vel3) PB_CUDA(ctx, cudaMemcpyAsync(vel3 + 3 * f, s + 7 * n + 3 * f, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, rs));
        if (angvel3) PB_CUDA(ctx, cudaMemcpyAsync(angvel3 + 3 * f, s + 10 * n + 3 * f, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, rs));
        if (!ctx->evRead[c]) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evRead[c], cudaEventDisableTiming));
        PB_CUDA(ctx, cudaEventRecord(ctx->evRead[c], rs));
        return PB_OK;
    };
    for (int pass = 0; pass < 2; ++pass) {
        for (int c = 0; c < nChunks; ++c) {
            const size_t f = (size_t)c * per;
            if (f >= n) break;
            const size_t m = std::min(n - f, (size_t)per);
            if (ctx->readPosesFirst) { if ((rc = pass == 0 ? copyPoses(c, f, m) : copyVels(c, f, m))) return rc; }
            else if (pass == 0) { if ((rc = copyPoses(c, f, m)) || (rc = copyVels(c, f, m))) return rc; }
        }
    }
    return PB_OK;
}

int pb_set_readback_order(pb_ctx* ctx, int poses_first) { ctx->readPosesFirst = poses_first ? 1 : 0; return PB_OK; }

int pb_get_state_wait_poses(pb_ctx* ctx, int chunk, int* first, int* count) {
    cudaSetDevice(ctx->device);
    if (chunk < 0 || chunk >= ctx->readChunks) { if (first) *first = 0; if (count) *count = 0; return chunk < 0 ? PB_EINVAL : PB_OK; }
    PB_CUDA(ctx, cudaEventSynchronize(ctx->evReadPose[chunk]));
    if (first) *first = ctx->readFirst[chunk];
    if (count) *count = ctx->readCount[chunk];
    return PB_OK;
}

int pb_get_state_wait(pb_ctx* ctx, int chunk, int* first, int* count) {
    cudaSetDevice(ctx->device);
    if (chunk < 0 || chunk >= ctx->readChunks) { if (first) *first = 0; if (count) *count = 0; return chunk < 0 ? PB_EINVAL : PB_OK; }
    PB_CUDA(ctx, cudaEventSynchronize(ctx->evRead[chunk]));
    if (first) *first = ctx->readFirst[chunk];
    if (count) *count = ctx->readCount[chunk];
    return PB_OK;
}

int pb_set_static_poses(pb_ctx* ctx, int nStatic, const float* pos3, const float* quat4) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    ctx->queryTreeValid = false;
    if (nStatic != ctx->nStatic) return pb_fail(ctx, PB_EINVAL, "pb_set_static_poses: n_static mismatch");
    if (!nStatic) return PB_OK;
    size_t n = (size_t)nStatic;
    int rc = ensureStage(ctx, sizeof(float) * 13 * (size_t)std::max(ctx->nDyn, nStatic)); if (rc) return rc;
    float* s = ctx->stage;
    PB_CUDA(ctx, cudaMemcpyAsync(s, pos3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + 3 * n, quat4, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    int g = pb_grid(nStatic, 256);
    ++ctx->launches, k_unpack3<<<g, 256, 0, ctx->stream>>>(nStatic, s, ctx->pos + ctx->nDyn);
    ++ctx->launches, k_unpack4<<<g, 256, 0, ctx->stream>>>(nStatic, s + 3 * n, ctx->quat + ctx->nDyn);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

int pb_set_bounds(pb_ctx* ctx, int n, const int* cols, const float* bounds6) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    if (n <= 0) return PB_OK;
    for (int i = 0; i < n; ++i) if (cols[i] < 0 || cols[i] >= ctx->nCol) return pb_fail(ctx, PB_EINVAL, "pb_set_bounds: collider out of range");
    int rc = ensureStage(ctx, sizeof(float) * 7 * (size_t)n); if (rc) return rc;
    float* s = ctx->stage;
    PB_CUDA(ctx, cudaMemcpyAsync(s, cols, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + n, bounds6, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->launches, k_scatter_bounds<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, (const int*)s, s + n, ctx->aabbMin, ctx->aabbMax);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// The same carry-over without the round trip through the host (pb_get_bounds + pb_set_bounds move 52 B per collider over PCIe and
// through two host loops): a device copy before the re-upload, one scatter after it.
int pb_keep_bounds_begin(pb_ctx* ctx) {
    cudaSetDevice(ctx->device);
    ctx->keptN = 0;
    if (ctx->nCol <= 0) return PB_OK;
    if (ctx->keptCap < ctx->nCol) {
        const int cap = ctx->nCol + ctx->nCol / 4 + 256;
        ctx->keptCap = 0;
        int rc = pb_alloc(ctx, &ctx->keptMin, (size_t)cap); if (rc) return rc;
        rc = pb_alloc(ctx, &ctx->keptMax, (size_t)cap); if (rc) return rc;
        ctx->keptCap = cap;
    }
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->keptMin, ctx->aabbMin, sizeof(float4) * (size_t)ctx->nCol, cudaMemcpyDeviceToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->keptMax, ctx->aabbMax, sizeof(float4) * (size_t)ctx->nCol, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->keptN = ctx->nCol;
    return PB_OK;
}

int pb_keep_bounds(pb_ctx* ctx, int nOld, const int* oldToNew) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    if (nOld != ctx->keptN) return pb_fail(ctx, PB_EINVAL, "pb_keep_bounds: n_old is not the collider count pb_keep_bounds_begin saw");
    ctx->keptN = 0;
    if (nOld <= 0) return PB_OK;
    int rc = ensureStage(ctx, sizeof(int) * (size_t)nOld); if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->stage, oldToNew, sizeof(int) * (size_t)nOld, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->launches, k_restore_bounds<<<pb_grid(nOld, 256), 256, 0, ctx->stream>>>(nOld, (const int*)ctx->stage, ctx->nCol, ctx->keptMin, ctx->keptMax, ctx->aabbMin, ctx->aabbMax);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_set_kinematic(pb_ctx* ctx, int nDyn, const int* kinematic) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }
    PB_STALE_BROADPHASE(ctx);
    if (nDyn != ctx->nDyn) return pb_fail(ctx, PB_EINVAL, "pb_set_kinematic: n_dynamic mismatch");
    if (!nDyn) return PB_OK;
    ctx->hKinematic.assign(kinematic, kinematic + nDyn);
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->kinematic, kinematic, sizeof(int) * nDyn, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->nCol) ++ctx->launches, k_col_dynamic_flag<<<pb_grid(ctx->nCol, 256), 256, 0, ctx->stream>>>(ctx->nCol, ctx->colRow, ctx->nDyn, ctx->kinematic, ctx->colFlags);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_set_mass(pb_ctx* ctx, int nDyn, const float* invMass, const float* com3, const float* invI9) {
    cudaSetDevice(ctx->device);
    { int rcw = pb_wait_velocities(ctx); if (rcw) return rcw; }      // the pending unpack reads invMass and writes the same records
    if (nDyn != ctx->nDyn) return pb_fail(ctx, PB_EINVAL, "pb_set_mass: n_dynamic mismatch");
    if (!nDyn) return PB_OK;
    int rc = ensureStage(ctx, sizeof(float) * 13 * (size_t)nDyn); if (rc) return rc;
    float* s = ctx->stage;
    PB_CUDA(ctx, cudaMemcpyAsync(s, com3, sizeof(float) * 3 * nDyn, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + 3 * (size_t)nDyn, invMass, sizeof(float) * nDyn, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(s + 4 * (size_t)nDyn, invI9, sizeof(float) * 9 * nDyn, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->launches, k_com_invmass<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, s, s + 3 * (size_t)nDyn, ctx->comInvMass);
    ++ctx->launches, k_unpack_m3<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, s + 4 * (size_t)nDyn, ctx->invIL);
    ++ctx->launches, k_refresh_invmass<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(nDyn, ctx->comInvMass, ctx->vel);   // invMass rides in v.w
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_keep_contact_cache(pb_ctx* ctx, int nOld, const int* oldToNew) {
    cudaSetDevice(ctx->device);
    if (!ctx->cacheBuilt || nOld <= 0) return PB_OK;   // nothing to keep (fresh context)
    int rc = ensureStage(ctx, sizeof(int) * (size_t)nOld); if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpyAsync(ctx->stage, oldToNew, sizeof(int) * (size_t)nOld, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = pb_contact_cache_remap(ctx, nOld, (const int*)ctx->stage))) return rc;
    ctx->cacheValid = true;
    return PB_OK;
}

int pb_keep_joint_state(pb_ctx* ctx, int n, const int* oldIndex) {
    cudaSetDevice(ctx->device);
    return pb_joints_keep_state(ctx, n, oldIndex);
}

int pb_set_contact_filter(pb_ctx* ctx, int nColliders, const int* colliderClass, int nClasses, const unsigned char* lut) {
    cudaSetDevice(ctx->device);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->filterLut) { cudaFree(ctx->filterLut); ctx->filterLut = nullptr; }
    ctx->nFilterClasses = 0;
    if (nClasses <= 0 || !lut) { ctx->triggersPossible = ctx->anyTriggerFlag; return PB_OK; }   // defaultContactFilter
    if (nColliders != ctx->nCol) return pb_fail(ctx, PB_EINVAL, "pb_set_contact_filter: collider count mismatch (upload colliders first)");
    for (int i = 0; i < nColliders; ++i) if (colliderClass[i] < 0 || colliderClass[i] >= nClasses) return pb_fail(ctx, PB_EINVAL, "collider class out of range");
    int rc = pb_alloc(ctx, &ctx->filterLut, (size_t)nClasses * nClasses); if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpy(ctx->filterLut, lut, (size_t)nClasses * nClasses, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(ctx->colClass, colliderClass, sizeof(int) * nColliders, cudaMemcpyHostToDevice));
    ctx->nFilterClasses = nClasses;
    bool any = false;
    for (size_t i = 0; i < (size_t)nClasses * nClasses; ++i) any |= lut[i] != 0;
    ctx->triggersPossible = any;
    return PB_OK;
}

int pb_get_triggers(pb_ctx* ctx, int* out4, int cap, int* n) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int nt = std::min(ctx->lastCounts.n_triggers, ctx->caps.max_pairs);
    *n = nt;
    if (!out4 || nt == 0 || cap == 0) return PB_OK;
    std::vector<int2> p(nt);
    PB_CUDA(ctx, cudaMemcpy(p.data(), ctx->trigPairs, sizeof(int2) * nt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nt && i < cap; ++i) {
        out4[4 * i] = ctx->hRowEntity[ctx->hColRow[p[i].x]]; out4[4 * i + 1] = ctx->hColIndex[p[i].x];
        out4[4 * i + 2] = ctx->hRowEntity[ctx->hColRow[p[i].y]]; out4[4 * i + 3] = ctx->hColIndex[p[i].y];
    }
    return PB_OK;
}

// post-build counters (manifolds with points, colours, points) are produced on the device after the last host sync of the
// step; they are fetched on demand so a plain simulate loop never waits for them
static int refreshCounts(pb_ctx* ctx) {
    if (!ctx->countsStale) return PB_OK;
    int rc = readCounters(ctx); if (rc) return rc;
    ctx->lastCounts.n_manifolds = ctx->hCounters[CNT_MANIFOLDS];
    ctx->lastCounts.n_colors = ctx->hCounters[CNT_NCOLORS];
    ctx->lastCounts.n_overflow = ctx->hCounters[CNT_OVERFLOW];
    ctx->lastCounts.n_points = ctx->hCounters[CNT_POINTS];
    ctx->countsStale = false;
    return PB_OK;
}

int pb_get_counts(pb_ctx* ctx, pb_counts* out) {
    cudaSetDevice(ctx->device);
    int rc = pb_collect_step(ctx);
    if (!rc) rc = refreshCounts(ctx);
    *out = ctx->lastCounts;
    return rc;
}

// candidate pairs per narrowphase bin of the last step (narrowphase.cu: sphere-sphere, sphere-capsule, capsule-capsule, sphere-box,
// capsule-box, box-box, GJK/EPA, mesh-sphere, mesh-capsule, mesh-box, mesh-convex, trigger)
int pb_get_bin_counts(pb_ctx* ctx, int* out12) {
    cudaSetDevice(ctx->device);
    int rc = pb_collect_step(ctx);
    if (!rc) { ctx->countsStale = true; rc = refreshCounts(ctx); }
    if (rc) return rc;
    for (int b = 0; b < 12; ++b) out12[b] = ctx->hCounters[CNT_BINSTART + b + 1] - ctx->hCounters[CNT_BINSTART + b];
    return PB_OK;
}

int pb_get_timings(pb_ctx* ctx, pb_timings* out) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pb_timings t{};
    cudaEventElapsedTime(&t.broadphase, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&t.narrowphase, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t.contact_build, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&t.solve, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&t.total, ctx->ev[0], ctx->ev[4]);
    if (ctx->nDyn > 0) cudaEventElapsedTime(&t.solve_kernel, ctx->ev[5], ctx->ev[6]);
    *out = t;
    return PB_OK;
}

int pb_get_pairs(pb_ctx* ctx, int* out4, int cap, int* n) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int np = std::min(ctx->lastCounts.n_pairs, ctx->caps.max_pairs);
    *n = np;
    if (!out4 || np == 0) return PB_OK;
    std::vector<int2> p(np);
    PB_CUDA(ctx, cudaMemcpy(p.data(), ctx->pairs, sizeof(int2) * np, cudaMemcpyDeviceToHost));
    std::vector<int> rowEnt(ctx->nRows), colIdx(ctx->nCol);
    PB_CUDA(ctx, cudaMemcpy(rowEnt.data(), ctx->rowEntity, sizeof(int) * ctx->nRows, cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(colIdx.data(), ctx->colIndex, sizeof(int) * ctx->nCol, cudaMemcpyDeviceToHost));
    for (int i = 0; i < np && i < cap; ++i) {
        out4[4 * i] = rowEnt[ctx->hColRow[p[i].x]]; out4[4 * i + 1] = colIdx[p[i].x];
        out4[4 * i + 2] = rowEnt[ctx->hColRow[p[i].y]]; out4[4 * i + 3] = colIdx[p[i].y];
    }
    return PB_OK;
}

int pb_get_bounds(pb_ctx* ctx, float* out6) {
    cudaSetDevice(ctx->device);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<float4> mn(ctx->nCol), mx(ctx->nCol);
    if (!ctx->nCol) return PB_OK;
    PB_CUDA(ctx, cudaMemcpy(mn.data(), ctx->aabbMin, sizeof(float4) * ctx->nCol, cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(mx.data(), ctx->aabbMax, sizeof(float4) * ctx->nCol, cudaMemcpyDeviceToHost));
    for (int i = 0; i < ctx->nCol; ++i) {
        out6[6 * i] = mn[i].x; out6[6 * i + 1] = mn[i].y; out6[6 * i + 2] = mn[i].z;
        out6[6 * i + 3] = mx[i].x; out6[6 * i + 4] = mx[i].y; out6[6 * i + 5] = mx[i].z;
    }
    return PB_OK;
}

int pb_get_manifolds(pb_ctx* ctx, int cap, int* keys5, int* numPoints, float* normal3, float* points24, int* color, int* n) {
    cudaSetDevice(ctx->device);
    { int rcc = pb_collect_step(ctx); if (rcc) return rcc; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    { int rc = refreshCounts(ctx); if (rc) return rc; }
    int nm = ctx->mSorted ? ctx->lastCounts.n_manifolds : 0;     // no solve order: the last step failed or the arenas were just replaced
    *n = nm;
    if (nm == 0 || cap == 0) return PB_OK;
    std::vector<int> sorted(nm);
    PB_CUDA(ctx, cudaMemcpy(sorted.data(), ctx->mSorted, sizeof(int) * nm, cudaMemcpyDeviceToHost));
    int nRaw = ctx->hCounters[CNT_RAWM];
    std::vector<int4> key(nRaw); std::vector<float4> nrm(nRaw), pts((size_t)8 * nRaw);
    PB_CUDA(ctx, cudaMemcpy(key.data(), ctx->mKey, sizeof(int4) * nRaw, cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(nrm.data(), ctx->mNormal, sizeof(float4) * nRaw, cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(pts.data(), ctx->mPts, sizeof(float4) * 8 * (size_t)nRaw, cudaMemcpyDeviceToHost));
    std::vector<int> rowEnt(ctx->nRows), colIdx(ctx->nCol);
    PB_CUDA(ctx, cudaMemcpy(rowEnt.data(), ctx->rowEntity, sizeof(int) * ctx->nRows, cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(colIdx.data(), ctx->colIndex, sizeof(int) * ctx->nCol, cudaMemcpyDeviceToHost));
    std::vector<unsigned int> skeys(nm);
    PB_CUDA(ctx, cudaMemcpy(skeys.data(), ctx->mSortedKeys, sizeof(unsigned int) * nm, cudaMemcpyDeviceToHost));
    for (int s = 0; s < nm && s < cap; ++s) {
        int raw = sorted[s];
        int4 k = key[raw];
        keys5[5 * s] = rowEnt[ctx->hColRow[k.x]]; keys5[5 * s + 1] = colIdx[k.x];
        keys5[5 * s + 2] = rowEnt[ctx->hColRow[k.y]]; keys5[5 * s + 3] = colIdx[k.y]; keys5[5 * s + 4] = k.z;
        numPoints[s] = k.w;
        normal3[3 * s] = nrm[raw].x; normal3[3 * s + 1] = nrm[raw].y; normal3[3 * s + 2] = nrm[raw].z;
        for (int p = 0; p < 4; ++p) for (int side = 0; side < 2; ++side) {
            float4 v = p < k.w ? pts[8 * (size_t)raw + 2 * p + side] : make_float4(0, 0, 0, 0);
            points24[24 * s + 6 * p + 3 * side] = v.x; points24[24 * s + 6 * p + 3 * side + 1] = v.y; points24[24 * s + 6 * p + 3 * side + 2] = v.z;
        }
        if (color) color[s] = (int)((skeys[s] % PB_KEY_COLORS) >> 1);      // key = group * 128 + colour * 2 + multi
    }
    return PB_OK;
}

int pb_debug_sort_pairs(pb_ctx* ctx, int n, int bits, const unsigned int* keysIn, const int* valsIn, unsigned int* keysOut, int* valsOut) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return PB_OK;
    unsigned int *kA = nullptr, *kB = nullptr, *hist = nullptr; int *vA = nullptr, *vB = nullptr;
    const int tiles = (n + 127) / 128 + 4;
    int rc = pb_alloc(ctx, &kA, (size_t)n);
    if (!rc) rc = pb_alloc(ctx, &kB, (size_t)n);
    if (!rc) rc = pb_alloc(ctx, &vA, (size_t)n);
    if (!rc) rc = pb_alloc(ctx, &vB, (size_t)n);
    if (!rc) rc = pb_alloc(ctx, &hist, (size_t)256 * tiles + (size_t)256 * tiles / 4096 + 1024);
    if (!rc) {
        cudaMemcpyAsync(kA, keysIn, sizeof(unsigned int) * n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(vA, valsIn, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream);
        bool inA = true;
        rc = pb_radix_sort_pairs(ctx, kA, vA, kB, vB, n, bits, hist, tiles, &inA);
        if (!rc) {
            cudaMemcpyAsync(keysOut, inA ? kA : kB, sizeof(unsigned int) * n, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(valsOut, inA ? vA : vB, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = pb_fail(ctx, PB_ECUDA, "pb_debug_sort_pairs");
        }
    }
    cudaFree(kA); cudaFree(kB); cudaFree(vA); cudaFree(vB); cudaFree(hist);
    return rc;
}

} // extern "C"

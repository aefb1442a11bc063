// Internal context: every device-resident SoA array of a scene, per-step arenas and scratch.
// All arrays are float4 / int aligned SoA so each kernel's loads coalesce (DESIGN.md "Data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/physecs_b200.h"

#define PB_MAX_COLORS 64          // contact colours: 0..62 parallel, 63 = sequential overflow bucket
#define PB_OVERFLOW_COLOR 63
#define PB_JOINT_COLORS 9         // reference JointGraph: 8 colours + overflow (Physecs.h:137-152)
#define PB_MAX_TRI_CONTACTS 24    // triangle contacts kept per (shape, mesh) pair
#define PB_NUM_BINS 16            // narrowphase shape-pair bins
#define PB_MAX_JOINT_ROWS 8
#define PB_ISLAND_LOCAL_MAX 1024  // constraints (manifolds + joints) an island may hold and still be solved inside one CTA
#define PB_KEY_COLORS 128         // solve-order key of a manifold: group * 128 + colour * 2 + (numPoints > 1)
#define PB_MAX_READ_CHUNKS 32     // chunks of a pb_get_state_begin read-back
#define PB_SPILL_CAP 65536        // pairs per spill list (GJK / EPA bin, mesh bins) and step
#define PB_SPILL_GJK_THREADS 128  // threads of k_np_gjk_spill == polytope scratch slots it owns
#define PB_SPILL_MESH_WARPS 32    // warps of k_np_mesh_spill == mesh scratch slots (each lane of each warp owns a polytope slot too)

struct PbTriMesh {
    int nVerts = 0, nTris = 0, nNodes = 0;
    float4* verts = nullptr;       // xyz
    int4* tris = nullptr;          // i0,i1,i2 (post-build order), w unused
    float4* triNormal = nullptr;   // xyz
    float4* triCentroid = nullptr; // xyz
    float4* triRec = nullptr;      // [4*nTris] packed triangle record for the sphere / capsule bins: {a.xyz, n.x} {b.xyz, n.y} {c.xyz, n.z} {i0, i1, i2 (int bits), 0}
    float4* nodeMin = nullptr;     // xyz, w = triCount (int bits)
    float4* nodeMax = nullptr;     // xyz, w = index    (int bits)
    float bmin[3], bmax[3];        // local-space bounds of all vertices (BoundsUtil.cpp:77-85 needs every vertex)
};

struct PbConvex {
    int nVerts = 0, nVertsPadded = 0, nFaces = 0, maxFaceVerts = 0;
    float4* verts = nullptr;        // padded to x4 by repeating the last vertex (ConvexMesh.cpp:8-10)
    int* faceOffsets = nullptr;     // nFaces+1
    int* faceIndices = nullptr;
    float4* faceNormal = nullptr;
    float4* faceCentroid = nullptr;
};

// device-side view of registered meshes (array of these lives in device memory)
struct PbTriMeshDev {
    const float4* verts; const int4* tris; const float4* triNormal; const float4* triCentroid;
    const float4* nodeMin; const float4* nodeMax; int nTris; int nNodes;
    float bmin[3]; float bmax[3];
    const float4* triRec;          // 64-byte record per triangle (vertices + normal + vertex indices): one gather instead of index -> vertex -> normal
};
struct PbConvexDev {
    const float4* verts; const int* faceOffsets; const int* faceIndices; const float4* faceNormal;
    const float4* faceCentroid; int nVerts; int nVertsPadded; int nFaces; int maxFaceVerts;
};

// header of the per-step device counters block (one int each, zeroed at step start)
enum {
    CNT_PAIRS = 0, CNT_MANIFOLDS, CNT_POINTS, CNT_STATUS, CNT_MESH_PAIRS, CNT_TRIGGERS, CNT_OVERFLOW, CNT_NCOLORS,
    CNT_RAWM,                           // raw manifold arena entries (incl. 0-point holes); CNT_MANIFOLDS = solve count
    CNT_GJK_HITS,                       // GJK-bin pairs whose shapes intersect (stage 2 of the split GJK / EPA launch works on these)
    CNT_SPILL_GJK, CNT_SPILL_MESH,      // pairs that outgrew the per-thread containers of their bin kernel (redone by the spill kernels)
    CNT_CAUSE,                          // PB_CAUSE_* bits (include/physecs_b200.h)
    CNT_SPILLED,                        // pairs the spill kernels redid
    CNT_BIN0 = 16,                      // PB_NUM_BINS bin counters
    CNT_BINSTART = 32,                  // PB_NUM_BINS+1 bin starts
    CNT_COLORSTART = 64,                // PB_MAX_COLORS+1 manifold start per colour
    CNT_MULTISTART = 130,               // PB_MAX_COLORS: where the multi-point manifolds of each colour start (singles come first)
    CNT_TILE_HITS = 202,                // (query tile, collider tile) pairs whose union boxes meet, from k_pairs_bruteforce / k_tile_probe: how coherent creation order is
    CNT_TOTAL = 256
};

struct pb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    pb_caps caps{};
    std::string err;
    int numSMs = 148;

    // ---- bodies -------------------------------------------------------------------------------
    int nDyn = 0, nStatic = 0, nRows = 0;
    int* rowEntity = nullptr;        // [rows] entt::entity integer
    float4* pos = nullptr;           // [rows] xyz
    float4* quat = nullptr;          // [rows] xyzw
    // velocity buffers interleave {v.xyz, invMass}, {w.xyz, 0} per body: X[2*i] is v of body i, angX == X + 1 so angX[2*i] is w
    float4* velBuf[2] = {nullptr, nullptr};   // owning allocations (2 float4 per body)
    float4* vel = nullptr;           // [dyn] component velocity (substep-start value)
    float4* angvel = nullptr;
    float4* velLive = nullptr;       // [dyn] velocityTemp the solver iterates on
    float4* angvelLive = nullptr;
    float4* comInvMass = nullptr;    // [dyn] com xyz, invMass w
    float4* invIL = nullptr;         // [3*dyn] local inverse inertia columns
    float4* bodyRec = nullptr;       // [8*dyn] per-substep body record for the prep kernels (solver.cu): q, world COM, invMass, v, w, vPre, wPre, world inverse inertia
    int* kinematic = nullptr;        // [dyn]
    float4* pseudoLin = nullptr;     // [dyn] xyz, w = constraintCount (int bits)
    float4* pseudoAng = nullptr;     // [dyn]
    unsigned long long* colorMask = nullptr; // [dyn] contact colours in use per body
    float* stage = nullptr;          // device staging for packed host uploads/downloads
    size_t stageBytes = 0;
    // pb_set_state uploads on a second stream.  The broadphase of the pb_step that follows reads neither poses (it works on the bounds
    // refreshed at the end of the previous step, like the reference: Physecs.cpp:556-559) nor velocities, the narrowphase needs the
    // poses, the contact build the velocities: the step waits for evPoseReady after the broadphase and for evVelReady before the
    // build, so the whole H2D copy hides behind the broadphase.  Every other entry point waits for both (pb_wait_velocities).
    // read-back of pb_get_state_begin: packed on the main stream into its own staging buffer, copied out on the READ stream, so the copies
    // of step k run beside the upload and the kernels of step k + 1 (pb_get_state_wait is the only thing that waits for them)
    cudaStream_t readStream = nullptr; cudaEvent_t evPacked = nullptr; float* stageRead = nullptr; size_t stageReadBytes = 0;
    // build stage: the colouring runs beside the island search, the joint lists beside the manifold ordering (contacts.cu pb_contact_build)
    cudaStream_t sideStream = nullptr; cudaEvent_t evFork = nullptr, evColour = nullptr, evGroups = nullptr, evJoints = nullptr; int buildFork = 1;   // env PB_BUILD_FORK=0: one stream
    cudaStream_t copyStream = nullptr; cudaEvent_t evMainAtSet = nullptr, evVelReady = nullptr, evPoseReady = nullptr; bool velPending = false, posePending = false;
    float* stageVel = nullptr; size_t stageVelBytes = 0;   // staging of the copy stream: 13 floats per body (pos 3, quat 4, vel 3, angvel 3)

    // ---- colliders ------------------------------------------------------------------------------
    int nCol = 0;
    int* colRow = nullptr; int* colIndex = nullptr; int* colType = nullptr; int* colFlags = nullptr; int* colData = nullptr;
    int* colMesh = nullptr;
    float4* colLPos = nullptr; float4* colLQuat = nullptr; float4* colParams = nullptr; float4* colMat = nullptr;
    float4* colWPos = nullptr; float4* colWQuat = nullptr;   // world pose, refreshed each step
    float4* aabbMin = nullptr; float4* aabbMax = nullptr;     // persistent bounds (BroadPhaseEntry::bounds)
    float4* keptMin = nullptr; float4* keptMax = nullptr; int keptCap = 0, keptN = 0;   // pb_keep_bounds_begin's copy of the bounds (carried across a collider re-upload)
    std::vector<int> hColType, hColMesh, hColRow, hColIndex, hRowEntity;   // host mirrors (trimesh colliders are found on the host)
    int* rowMark = nullptr;          // [rows] scratch marks for pb_move_rows

    // ---- meshes ------------------------------------------------------------------------------------
    std::vector<PbTriMesh> triMeshes; std::vector<PbConvex> convexes;
    PbTriMeshDev* triMeshDev = nullptr; PbConvexDev* convexDev = nullptr;

    // ---- non-colliding entity pairs (sorted u64 keys) ----------------------------------------------
    unsigned long long* nonColliding = nullptr; int nNonColliding = 0;

    // ---- broadphase scratch ---------------------------------------------------------------------------
    unsigned int* mortonA = nullptr; unsigned int* mortonB = nullptr; int* leafIdA = nullptr; int* leafIdB = nullptr;
    unsigned int* radixHist = nullptr; int radixTiles = 0;
    float* sceneBounds = nullptr;    // 6 floats (ordered-int encoded) min/max of AABB centres
    int* nodeLeft = nullptr; int* nodeRight = nullptr; int* nodeParent = nullptr; int* leafParent = nullptr;
    int2* nodeRange = nullptr; int* nodeFlag = nullptr;
    const int* treeLeafIds = nullptr;   // sorted leaf -> collider of the last pb_build_tree
    int fusedMode = -1;                 // -1 size rule, 0 never, 1 always: whole-step kernel for tiny scenes (env PB_FUSED)
    int sortCoopMode = 1;               // 1: sorts above 8192 keys run as one cooperative launch (k_radix_sort_coop); 0: three launches per pass (env PB_SORT_COOP)
    int sortCoopGrid = 0; unsigned int* sortBarrier = nullptr;
    bool sortSmallOptIn = false;        // k_sort_small's dynamic shared memory opt-in done on this context's device
    int meshLightMode = 1;              // sphere / capsule vs mesh bins: 0 = k_np_mesh, 1 = k_np_mesh_light (dual-child cull walk + packed
                                        // triangle records) (env PB_MESH_LIGHT)
    int mortonIso = 1;                  // Morton keys over cubic cells (one scale for the three axes); 0 = each axis scaled to its own extent (env PB_MORTON_ISO)
    int4* colInfo = nullptr;            // [colliders] (flags, body row, entity, 0) per collider, rewritten by k_morton for the pair walk
    int* bigList = nullptr;             // [1 + 32] count + colliders of the step's big-static side list (broadphase.cu k_morton)
    int bigListMode = 1;                // 0: every collider stays in the step's tree (env PB_BIG_LIST)
    int pairsHint = -1;                 // candidate pairs of the previous step (-1: none yet): bounds the bin kernels' grids on small scenes
    int npFuseSmall = 1;                // small scenes: the six analytic bins in one launch (env PB_NP_FUSE=0: one launch per bin)
    int npWaves = 4;                    // narrowphase bin kernels: grid = SMs x co-resident CTAs x npWaves (env PB_NP_WAVES; 0 = the former numSMs * 8)
    int bruteForceBigMax = 131072;      // ... and up to this many when creation order is spatially coherent (a batch of scenes side by side): decided from the
                                        // previous step's tile statistics (env PB_BRUTE_FORCE_BIG_MAX; 0 = off)
    int lastTileHits = -1, lastTiles = 0;   // tile pairs that met / tiles, last step that looked (-1: never)
    int pendingTiles = 0;               // tiles of the step in flight if it produces tile statistics (0: it does not)
    bool stepBrute = false;             // this step's broadphase was the all-pairs kernel
    int bruteForceMax = 8192;           // colliders up to which the step tests all pairs directly instead of building the tree (env PB_BRUTE_FORCE_MAX)
    bool queryTreeValid = false;        // tree + world poses match the current bounds / poses (scene queries)
    int* queryOut = nullptr; int queryCap = 0;   // device result buffer of the scene queries (queries.cu)
    // overlapWithMinTranslationalDistance: a private pair / manifold arena the narrowphase bin kernels run on in query mode
    int* qCounters = nullptr; int2* qPairs = nullptr; int* qPairOrder = nullptr; int4* qmKey = nullptr; float4* qmNormal = nullptr;
    float4* qmPts = nullptr; int qArenaCap = 0;
    float4* nodeMin = nullptr; float4* nodeMax = nullptr;
    int2* pairs = nullptr;           // [maxPairs] (colA, colB); A is the lower-entity side
    int* pairOrder = nullptr;        // [2*maxPairs] pair indices grouped by bin | bin of each pair

    // ---- manifolds (raw narrowphase output) -----------------------------------------------------------
    int4* mKey = nullptr;            // (colA, colB, tri, numPoints)
    float4* mNormal = nullptr;       // xyz
    float4* mPts = nullptr;          // [8*maxManifolds]: slot 2k = position0, 2k+1 = position1
    int* mColor = nullptr;           // colour per raw manifold
    int* gjkHitPair = nullptr; float4* gjkHitSimplex = nullptr; int gjkHitCap = 0;   // intersecting GJK-bin pairs + their simplices (9 float4 each)
    // spill path (narrowphase.cu): pair lists [2][PB_SPILL_CAP] (GJK / EPA bin, mesh bins) + global-memory scratch of the spill kernels
    int* spillList = nullptr; void* spillEpa = nullptr; void* spillMesh = nullptr;
    int* mSorted = nullptr;          // [maxManifolds] raw index per solve slot
    unsigned int* mSortKeyA = nullptr; unsigned int* mSortKeyB = nullptr; int* mSortValB = nullptr;

    // ---- contact constraints (solve order) ---------------------------------------------------------------
    int4* cHead = nullptr;           // packed solve header: b0, b1, first point, numPoints | isSoft << 8 (one 16-byte load)
    int2* cBodies = nullptr;         // solver body index or -1 (b0, b1)
    int2* cRowsT = nullptr;          // transform rows (row0, row1)
    float4* cNormal = nullptr;       // n xyz, friction w
    float4* cSoft = nullptr;         // isSoft, frequency, dampingRatio, unused
    // per point, double-buffered (prev step kept for the contact cache)
    float4* pR0T[2] = {nullptr, nullptr};   // local r0 xyz, targetVelocity w
    float4* pR1 = nullptr;                  // local r1 xyz
    int curBuf = 0;
    // per-substep rows per point
    float4* rowA = nullptr;          // r0xn xyz, c
    float4* rowB = nullptr;          // r1xn xyz, effMassN (1/k or 0)
    float4* rowC = nullptr;          // I0^-1 (r0xn) xyz, targetVelocity
    float4* rowD = nullptr;          // I1^-1 (r1xn) xyz, lambdaT0 (friction increment, constant within a substep)
    float4* rowE = nullptr;          // t xyz, kT!=0 flag
    float4* rowF = nullptr;          // I0^-1 (r0xt) xyz
    float4* rowG = nullptr;          // I1^-1 (r1xt) xyz
    float2* rowL = nullptr;          // totalLambdaN, totalLambdaT
    // contact cache (hash table over previous step's manifolds)
    unsigned long long* cacheTag[2] = {nullptr, nullptr}; int4* cacheVal[2] = {nullptr, nullptr}; int cacheSize = 0;
    bool cacheValid = false;
    bool cacheBuilt = false;         // a step has run on this context: the previous-step tables hold real data
    int* cPointOfsBuf[2] = {nullptr, nullptr}; int* cNpBuf[2] = {nullptr, nullptr};

    // ---- joints (arrays owned by joints.cu) ----------------------------------------------------------------------
    int nJoints = 0;
    int jointColorStart[PB_JOINT_COLORS + 1] = {0};
    void* jointStore = nullptr;

    // ---- counters / host mirrors ---------------------------------------------------------------------------
    int* counters = nullptr;         // CNT_TOTAL ints on device
    int* hCounters = nullptr;        // pinned mirror
    pb_counts lastCounts{};
    pb_timings lastTimings{};
    cudaEvent_t ev[8] = {nullptr};
    // contact filter (Physecs.cpp:200): optional K x K table over (isTrigger, data) classes; nullptr = defaultContactFilter
    int* colClass = nullptr; unsigned char* filterLut = nullptr; int nFilterClasses = 0;
    bool anyTriggerFlag = false, triggersPossible = false;
    bool anyRestitution = true;      // some collider has restitution != 0: only then does the contact cache (restitution targets of persisting contacts) do anything
    int2* trigPairs = nullptr;       // [maxPairs] overlapping TRIGGER pairs of the last step (collider indices)
    // simulation islands (islands.cu): group of every solver body; local groups 0..islandGroups-1 (one CTA each), group islandGroups = global
    int* islandParent = nullptr; int* islandCount = nullptr; int* bodyGroup = nullptr; int* islandStats = nullptr;
    // per-group body lists (built when the whole-step kernel may run group by group): bodyOrder = bodies sorted by group, bodyStart[g] their runs
    int* bodyOrder = nullptr; int* bodyStart = nullptr; int* bodyCursor = nullptr; bool bodyListsBuilt = false;
    int clusterSize = -1;            // CTAs of the thread-block cluster small one-pile scenes run their whole-step kernel as (-1: not asked yet, 0: off; env PB_CLUSTER=0)
    int fusedNarrowMax = 48;         // all-local scenes with at most this many constraints per group take the 128-thread whole-step kernel (env PB_FUSED_NARROW; 0 = never)
    int fusedLocalMax = 65536;       // bodies up to which an all-local scene takes the one-launch whole-step kernel (env PB_FUSED_LOCAL_MAX)
    int islandGroups = 0;            // G: fixed per context (the co-resident CTA count of the persistent kernel)
    int islandsMode = 2;             // 0 off, 1 on, 2 auto (on while a worthwhile share of the constraints sits in small islands)
    int islandLocalMax = PB_ISLAND_LOCAL_MAX;   // env PB_ISLAND_LOCAL_MAX overrides (tests: force a mix of local and device-wide sweeps)
    bool islandsOn = false;          // this step
    int islandsHold = 0;             // auto: steps left before small islands are looked for again
    int lastIslandLocal = 0, lastIslandTotal = 0;   // constraints in small islands / in all islands, last step that looked
    int* keyStart = nullptr;         // [(G + 1) * PB_KEY_COLORS + 1] first solve slot of every (group, colour, single | multi) run
    int* keyCursor = nullptr;              // [(G+1)*128+1] fill counters of the runs for the counting-sort scatter (contacts.cu; zeroed by k_build_clear)
    unsigned int* mSortedKeys = nullptr;   // solve-order keys of the last step (taps: colour of a slot)
    // per-group joint lists of the step (joints.cu): jointOrder = joints sorted by (group, colour), jointStart[g * 8 + c] their runs (g = G: the global group, colours 0..8)
    int* jointKey = nullptr; int* jointOrder = nullptr; int* jointStart = nullptr; int* jointSortTmp[3] = {nullptr, nullptr, nullptr}; int jointListCap = 0;
    // persistent substep kernel (solver.cu)
    int solveGrid = 0; unsigned int* solveBarrier = nullptr; unsigned long long* solveProfNs = nullptr;
    bool countsStale = false;        // lastCounts lacks the post-build numbers until the counters are read back
    // A step is enqueued without a host sync: the arena checks happen on the device (a step that overflowed skips its solve and leaves
    // the scene untouched) and the host COLLECTS the outcome -- counters snapshot, status -- at the next call that synchronises with the
    // step (capi.cu collectStep).  stepPending: a step's outcome has not been collected yet; undo*: host bookkeeping to roll back then.
    cudaEvent_t evCounters = nullptr; bool stepPending = false; bool statsCopied = false;   // statsCopied: the island statistics rode along with the counters
    bool stepNarrowed = false;       // pb_step_narrowphase ran: the next pb_step continues behind the narrowphase
    bool stepBegun = false;          // pb_step_begin ran: the next pb_step continues behind its broadphase
    bool mainMarked = false;         // evMainAtSet already holds "the main stream before this step's broadphase": uploads wait for that
    int readPosesFirst = 0;          // pb_set_readback_order: 1 = the poses of every chunk go out before any velocity
    cudaEvent_t evReadPose[PB_MAX_READ_CHUNKS] = {nullptr};      // ... the chunk's poses have arrived (they are copied out first)
    cudaEvent_t evRead[PB_MAX_READ_CHUNKS] = {nullptr}; int readFirst[PB_MAX_READ_CHUNKS] = {0}, readCount[PB_MAX_READ_CHUNKS] = {0}, readChunks = 0;   // pb_get_state_begin / _wait
    int rawHint = -1;                // raw manifold count of the last collected step (-1: none yet): shapes grids only
    bool undoCacheValid = false, undoCacheBuilt = false; int undoVelSwaps = 0;
    // PB_DETERMINISTIC=1 (or pb_set_deterministic): colours by fixed priorities (contacts.cu k_color_jp) -- two runs of the same scene give
    // bit-identical trajectories, like the reference with numThreads = 0 (ThreadPool.cpp:30-46)
    bool deterministic = false; unsigned long long* jpBest = nullptr; int* jpScratch = nullptr; int jpGrid = 0;
    std::vector<int> hTrimeshCols;   // colliders of type PB_TRIANGLE_MESH (their bounds are a vertex reduction each)
    std::vector<int> hKinematic;     // host mirror of `kinematic` (which trimesh colliders ride on moving bodies)
    unsigned long long launches = 0; // kernels launched by this context since creation (bench: gpu_launches)
    bool profile = false;            // per-phase device timing inside the persistent substep kernel (pb_set_profile)
    cudaEvent_t evSub[16] = {nullptr}; int evSubCount = 0;     // profiling: events around the k_substep_solve launches of the last step
};


// ---- helpers --------------------------------------------------------------------------------------------------
int pb_fail(pb_ctx* ctx, int code, const std::string& msg);
#define PB_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return pb_fail((ctx), PB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T> static inline int pb_alloc(pb_ctx* ctx, T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return pb_fail(ctx, PB_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return PB_OK;
}

static inline int pb_grid(long long n, int block) { long long g = (n + block - 1) / block; return (int)(g < 1 ? 1 : g); }
// Grid of a grid-stride kernel over a count only the device knows: `hint` = the host's guess (the previous step's count, -1 = none),
// with headroom; never more than `most` CTAs.  Only shapes the launch -- a step of a small scene is a chain of few-microsecond kernels,
// and a thousand empty CTAs per launch are most of what such a kernel costs.
static inline int pb_hint_grid(int hint, int block, int most) {
    if (hint < 0) return most;
    long long g = (2ll * hint + 2048 + block - 1) / block;
    return (int)(g < 8 ? 8 : (g > most ? most : g));
}

// stage launches (implemented in the .cu files)
int pb_wait_velocities(pb_ctx* ctx);   // main stream waits for a pending pb_set_state upload, poses and velocities (capi.cu)
int pb_wait_poses(pb_ctx* ctx);        // ... for its pose half only
int pb_broadphase(pb_ctx* ctx);
int pb_build_tree(pb_ctx* ctx, bool forStep = false);
int pb_update_bounds_all(pb_ctx* ctx, float margin, int onlyDynamic, const int* skipStatus = nullptr);
int pb_update_bounds_rows(pb_ctx* ctx, const int* dRowMark, int n, float margin);
int pb_update_bounds_trimesh_col(pb_ctx* ctx, int col, float margin);
int pb_world_poses(pb_ctx* ctx);
int pb_narrowphase(pb_ctx* ctx);
// the same bin kernels on a private arena: pairs (collider, query collider slot) -> manifolds, no filters (queries.cu)
int pb_narrowphase_query(pb_ctx* ctx, int* counters, const int2* pairs, int* pairOrder, int cap, int4* mKey, float4* mNormal, float4* mPts);
int pb_contact_build(pb_ctx* ctx);
int pb_solve(pb_ctx* ctx, float dt, int substeps, int iterations, float gravity);
int pb_solve_profile(pb_ctx* ctx, unsigned long long* out, bool reset);
int pb_solve_profile_colors(pb_ctx* ctx, unsigned long long* out128);
int pb_joint_begin_step(pb_ctx* ctx);
int pb_islands_build(pb_ctx* ctx);
int pb_joint_lists(pb_ctx* ctx);   // per-group joint lists of the step (joints.cu), needs pb_islands_build
int pb_contact_cache_remap(pb_ctx* ctx, int nOld, const int* dOldToNew);
void pb_contact_cache_rehash(pb_ctx* ctx, int oldSize, const unsigned long long* oldTag, const int4* oldVal, int newSize, unsigned long long* newTag, int4* newVal);

// generic device primitives (primitives.cu)
int pb_radix_sort_pairs(pb_ctx* ctx, unsigned int* keysA, int* valsA, unsigned int* keysB, int* valsB, int n, int bits,
                        unsigned int* hist, int histCapTiles, bool* resultInA);
int pb_exclusive_scan(pb_ctx* ctx, const int* in, int* out, int n, int* scratch);
int pb_exclusive_scan_dev(pb_ctx* ctx, const int* in, int* out, const int* nDev, int cap, int bound, int* scratch);

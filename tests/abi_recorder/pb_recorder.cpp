// TEST DOUBLE of the C ABI (include/physecs_b200.h) for the `-m "not gpu"` tests of the host layer (physecs_b200/host/Scene.cpp).
//
// It is NOT a CPU implementation of anything: no broadphase, no narrowphase, no solver, no bounds arithmetic.  It records what the host
// layer uploads (rows, collider tables, joints, carry-over maps) so a test can check the host's bookkeeping -- row <-> entity maps,
// collider order, what survives a structural edit -- without a device, and its "step" adds 1 to every non-kinematic dynamic row's
// pos.x so a test can see that the read-back lands on the entities the rows stand for.  Bounds are opaque tags: upload number k gives
// collider i the tag (k, i); pb_move_rows re-tags with k = -1.  Only tests/ builds or loads this file (tests/abi_recorder/build.py);
// the product libraries never see it.
#include "../../include/physecs_b200.h"
#include "../../physecs_b200/csrc/trimesh_build.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct pb_ctx {
    pb_caps caps{};
    std::string error;
    int nDyn = 0, nStatic = 0, nCol = 0, uploads = 0, steps = 0, reuploads = 0;
    std::vector<int> entity, kinematic;
    std::vector<float> pos, quat, vel, ang;                  // rows x 3 / 4, dynamic rows x 3
    std::vector<int> colRow, colIdx, colType, colFlags, colData, colMesh;
    std::vector<float> bounds, kept;                         // 6 per collider: (tag k, tag i, 0, 0, 0, 0)
    int keptN = 0;
    std::vector<int> lastBoundsMap, lastCacheMap, lastJointMap;
    std::vector<int> jointType, jointRow0, jointRow1, jointColor;
    std::vector<int> noCollide, filterClass;
    int nConvex = 0, nTrimesh = 0, nFilterClasses = 0;
    float *outPos = nullptr, *outQuat = nullptr, *outVel = nullptr, *outAng = nullptr;
    int chunks = 1;
    int failSteps = 0; bool failed = false; int needPairs = 0, needManifolds = 0;     // pbr_fail_next_steps: steps that overflow an arena
};

static int fail(pb_ctx* c, int code, const char* msg) { c->error = msg; return code; }

extern "C" {

int pb_ctx_create(int, const pb_caps* caps, pb_ctx** out) {
    if (std::getenv("PB_RECORDER_NO_DEVICE")) { *out = nullptr; return PB_ECUDA; }
    auto* c = new pb_ctx; c->caps = *caps; *out = c; return PB_OK;
}
void pb_ctx_destroy(pb_ctx* c) { delete c; }
const char* pb_last_error(pb_ctx* c) { return c ? c->error.c_str() : ""; }
int pb_grow_arenas(pb_ctx* c, int p, int m) { c->caps.max_pairs = p; c->caps.max_manifolds = m; return PB_OK; }
int pb_host_alloc(void** p, unsigned long long bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? PB_OK : PB_ECUDA; }
void pb_host_free(void* p) { std::free(p); }

int pb_upload_bodies(pb_ctx* c, int nDyn, int nStatic, const int* entity, const float* pos3, const float* quat4, const int* kin,
                     const float* vel3, const float* ang3, const float*, const float*, const float*) {
    const int rows = nDyn + nStatic;
    if (rows > c->caps.max_bodies) return fail(c, PB_ECAPACITY, "max_bodies");
    c->nDyn = nDyn; c->nStatic = nStatic;
    c->entity.assign(entity, entity + rows);
    c->pos.assign(pos3, pos3 + 3 * (size_t)rows); c->quat.assign(quat4, quat4 + 4 * (size_t)rows);
    c->kinematic.assign(kin, kin + nDyn);
    c->vel.assign(vel3, vel3 + 3 * (size_t)nDyn); c->ang.assign(ang3, ang3 + 3 * (size_t)nDyn);
    ++c->reuploads;
    return PB_OK;
}
int pb_upload_colliders(pb_ctx* c, int n, const int* row, const int* idx, const float*, const float*, const int* type, const float*, const int* mesh,
                        const float*, const int* flags, const int* data) {
    if (n > c->caps.max_colliders) return fail(c, PB_ECAPACITY, "max_colliders");
    for (int i = 0; i < n; ++i) if (row[i] < 0 || row[i] >= c->nDyn + c->nStatic) return fail(c, PB_EINVAL, "collider body_row out of range");
    c->nCol = n; ++c->uploads;
    c->colRow.assign(row, row + n); c->colIdx.assign(idx, idx + n); c->colType.assign(type, type + n);
    c->colFlags.assign(flags, flags + n); c->colData.assign(data, data + n); c->colMesh.assign(mesh, mesh + n);
    c->bounds.assign(6 * (size_t)n, 0.f);
    for (int i = 0; i < n; ++i) { c->bounds[6 * (size_t)i] = (float)c->uploads; c->bounds[6 * (size_t)i + 1] = (float)i; }
    c->nFilterClasses = 0; c->filterClass.clear();
    return PB_OK;
}
int pb_register_convex(pb_ctx* c, const float*, int, const int*, const int*, int, const float*, const float*, int* h) { *h = c->nConvex++; return PB_OK; }
int pb_register_trimesh(pb_ctx* c, const float*, int, const unsigned*, int, int* h, int*) { *h = c->nTrimesh++; return PB_OK; }
// setup-time HOST code of the product (csrc/trimesh_build.cpp, no device involved): the double forwards to it like capi.cu does
int pb_build_trimesh(const float* verts3, int nVerts, const unsigned* indices, int nIndices, unsigned* triIdx, int* triOrig, float* nodeBounds6, int* nodeCountIndex2, int* nNodes) {
    PbHostTriMesh h;
    pb_build_trimesh_host(verts3, nVerts, indices, nIndices, h);
    if (triIdx) std::memcpy(triIdx, h.triIdx.data(), sizeof(unsigned) * h.triIdx.size());
    if (triOrig) std::memcpy(triOrig, h.triOrig.data(), sizeof(int) * h.triOrig.size());
    if (nodeBounds6) std::memcpy(nodeBounds6, h.nodeBounds.data(), sizeof(float) * h.nodeBounds.size());
    if (nodeCountIndex2) std::memcpy(nodeCountIndex2, h.nodeCountIndex.data(), sizeof(int) * h.nodeCountIndex.size());
    if (nNodes) *nNodes = h.nNodes;
    return PB_OK;
}
int pb_upload_joints(pb_ctx* c, int n, const int* type, const int* r0, const int* r1, const float*, const float*, const float*, const float*, const float*, const int* color) {
    if (n > c->caps.max_joints) return fail(c, PB_ECAPACITY, "max_joints");
    c->jointType.assign(type, type + n); c->jointRow0.assign(r0, r0 + n); c->jointRow1.assign(r1, r1 + n); c->jointColor.assign(color, color + n);
    return PB_OK;
}
int pb_update_joint_params(pb_ctx*, int, const float*) { return PB_OK; }
int pb_keep_joint_state(pb_ctx* c, int n, const int* old) { c->lastJointMap.assign(old, old + n); return PB_OK; }
int pb_set_noncolliding_pairs(pb_ctx* c, int n, const int* p) { c->noCollide.assign(p, p + 2 * (size_t)n); return PB_OK; }
int pb_set_contact_filter(pb_ctx* c, int n, const int* cls, int K, const unsigned char*) {
    c->nFilterClasses = K > 0 ? K : 0;
    if (K > 0 && cls) c->filterClass.assign(cls, cls + n); else c->filterClass.clear();
    return PB_OK;
}
int pb_set_kinematic(pb_ctx* c, int nDyn, const int* k) { if (nDyn != c->nDyn) return fail(c, PB_EINVAL, "n_dynamic"); c->kinematic.assign(k, k + nDyn); return PB_OK; }
int pb_set_mass(pb_ctx* c, int nDyn, const float*, const float*, const float*) { return nDyn == c->nDyn ? PB_OK : fail(c, PB_EINVAL, "n_dynamic"); }

int pb_get_bounds(pb_ctx* c, float* out6) { std::copy(c->bounds.begin(), c->bounds.end(), out6); return PB_OK; }
int pb_set_bounds(pb_ctx* c, int n, const int* cols, const float* b6) {
    for (int i = 0; i < n; ++i) {
        if (cols[i] < 0 || cols[i] >= c->nCol) return fail(c, PB_EINVAL, "pb_set_bounds: collider out of range");
        std::copy(b6 + 6 * (size_t)i, b6 + 6 * (size_t)i + 6, c->bounds.begin() + 6 * (size_t)cols[i]);
    }
    return PB_OK;
}
int pb_keep_bounds_begin(pb_ctx* c) { c->kept = c->bounds; c->keptN = c->nCol; return PB_OK; }
int pb_keep_bounds(pb_ctx* c, int nOld, const int* map) {
    if (nOld != c->keptN) return fail(c, PB_EINVAL, "pb_keep_bounds: n_old is not the collider count pb_keep_bounds_begin saw");
    c->lastBoundsMap.assign(map, map + nOld);
    for (int o = 0; o < nOld; ++o)
        if (map[o] >= 0 && map[o] < c->nCol) std::copy(c->kept.begin() + 6 * (size_t)o, c->kept.begin() + 6 * (size_t)o + 6, c->bounds.begin() + 6 * (size_t)map[o]);
    c->keptN = 0;
    return PB_OK;
}
int pb_keep_contact_cache(pb_ctx* c, int nOld, const int* map) { c->lastCacheMap.assign(map, map + nOld); return PB_OK; }
int pb_move_rows(pb_ctx* c, int n, const int* rows, const float* p, const float* q) {
    for (int i = 0; i < n; ++i) {
        const int r = rows[i];
        if (r < 0 || r >= c->nDyn + c->nStatic) return fail(c, PB_EINVAL, "row");
        std::copy(p + 3 * i, p + 3 * i + 3, c->pos.begin() + 3 * (size_t)r); std::copy(q + 4 * i, q + 4 * i + 4, c->quat.begin() + 4 * (size_t)r);
        for (int k = 0; k < c->nCol; ++k) if (c->colRow[k] == r) c->bounds[6 * (size_t)k] = -1.f;
    }
    return PB_OK;
}

int pb_set_state_rows(pb_ctx* c, int first, int count, const float* p, const float* q, const float* v, const float* w) {
    if (first < 0 || first + count > c->nDyn) return fail(c, PB_EINVAL, "rows");
    if (p) std::copy(p, p + 3 * (size_t)count, c->pos.begin() + 3 * (size_t)first);
    if (q) std::copy(q, q + 4 * (size_t)count, c->quat.begin() + 4 * (size_t)first);
    if (v) std::copy(v, v + 3 * (size_t)count, c->vel.begin() + 3 * (size_t)first);
    if (w) std::copy(w, w + 3 * (size_t)count, c->ang.begin() + 3 * (size_t)first);
    return PB_OK;
}
int pb_set_state(pb_ctx* c, int nDyn, const float* p, const float* q, const float* v, const float* w) {
    if (nDyn != c->nDyn) return fail(c, PB_EINVAL, "n_dynamic");
    return pb_set_state_rows(c, 0, nDyn, p, q, v, w);
}
int pb_set_static_poses(pb_ctx* c, int nStatic, const float* p, const float* q) {
    if (nStatic != c->nStatic) return fail(c, PB_EINVAL, "n_static");
    std::copy(p, p + 3 * (size_t)nStatic, c->pos.begin() + 3 * (size_t)c->nDyn); std::copy(q, q + 4 * (size_t)nStatic, c->quat.begin() + 4 * (size_t)c->nDyn);
    return PB_OK;
}
int pb_step_begin(pb_ctx*) { return PB_OK; }
int pb_step_narrowphase(pb_ctx*) { return PB_OK; }
int pb_step(pb_ctx* c, float, int, int, float) {
    // an arena overflow as the device reports it: the step leaves the scene as it was, the status arrives with the next call that
    // synchronises with it (the read-back), the counters say how much room the step needs (needPairs < 0: a failure no arena explains)
    c->failed = false;
    if (c->failSteps > 0 && (c->needPairs < 0 || c->caps.max_pairs < c->needPairs || c->caps.max_manifolds < c->needManifolds)) { --c->failSteps; c->failed = true; return PB_OK; }
    for (int r = 0; r < c->nDyn; ++r) if (!c->kinematic[r]) c->pos[3 * (size_t)r] += 1.f;       // the visible trace of a "step"
    ++c->steps;
    return PB_OK;
}
int pb_get_state_begin(pb_ctx* c, float* p, float* q, float* v, float* w, int chunks) {
    if (c->failed) return fail(c, PB_ECAPACITY, "per-step arena overflow (recorded)");
    std::copy(c->pos.begin(), c->pos.begin() + 3 * (size_t)c->nDyn, p); std::copy(c->quat.begin(), c->quat.begin() + 4 * (size_t)c->nDyn, q);
    std::copy(c->vel.begin(), c->vel.end(), v); std::copy(c->ang.begin(), c->ang.end(), w);
    c->chunks = std::max(1, std::min(chunks, 32));
    return PB_OK;
}
int pb_get_state_wait(pb_ctx* c, int chunk, int* first, int* count) {
    const int per = (c->nDyn + c->chunks - 1) / c->chunks;
    const int f = chunk * per;
    if (chunk >= c->chunks || f >= c->nDyn) { *first = 0; *count = 0; return PB_OK; }
    *first = f; *count = std::min(per, c->nDyn - f);
    return PB_OK;
}
int pb_sync(pb_ctx*) { return PB_OK; }
int pb_get_state(pb_ctx* c, float* p, float* q, float* v, float* w) {
    if (c->failed) return fail(c, PB_ECAPACITY, "per-step arena overflow (recorded)");
    if (p) std::copy(c->pos.begin(), c->pos.begin() + 3 * (size_t)c->nDyn, p);
    if (q) std::copy(c->quat.begin(), c->quat.begin() + 4 * (size_t)c->nDyn, q);
    if (v) std::copy(c->vel.begin(), c->vel.end(), v);
    if (w) std::copy(c->ang.begin(), c->ang.end(), w);
    return PB_OK;
}
int pb_get_counts(pb_ctx* c, pb_counts* out) {
    std::memset(out, 0, sizeof *out);
    if (c->failed) { out->n_pairs = std::max(c->needPairs, 0); out->n_manifolds = std::max(c->needManifolds, 0); out->status = PB_ECAPACITY; }
    return PB_OK;
}
int pb_get_timings(pb_ctx*, pb_timings* out) { std::memset(out, 0, sizeof *out); return PB_OK; }
int pb_get_manifolds(pb_ctx*, int, int*, int*, float*, float*, int*, int* n) { *n = 0; return PB_OK; }
int pb_get_triggers(pb_ctx*, int*, int, int* n) { *n = 0; return PB_OK; }
int pb_get_tree(pb_ctx*, int, float*, int*, int* n) { *n = 0; return PB_OK; }
int pb_collider_ids(pb_ctx* c, int n, const int* cols, int* e, int* idx) {
    for (int i = 0; i < n; ++i) { e[i] = c->entity[c->colRow[cols[i]]]; idx[i] = c->colIdx[cols[i]]; }
    return PB_OK;
}
int pb_query_overlap(pb_ctx*, const float*, const float*, int, const float*, int, int, int, int*, int*, int* n) { *n = 0; return PB_OK; }
int pb_query_overlap_mtd(pb_ctx*, const float*, const float*, int, const float*, int, int, int*, int*, float*, float*, int* n) { *n = 0; return PB_OK; }
int pb_query_raycast(pb_ctx*, int, const float*, const float*, float, int, int*, int*, int*, float*, int* n) { *n = 0; return PB_OK; }

// ---- what the tests read back ------------------------------------------------------------------------------------------------------
int pbr_counts(pb_ctx* c, int* out8) {
    out8[0] = c->nDyn; out8[1] = c->nStatic; out8[2] = c->nCol; out8[3] = c->uploads; out8[4] = c->steps; out8[5] = (int)c->jointType.size();
    out8[6] = (int)c->noCollide.size() / 2; out8[7] = c->nFilterClasses;
    return PB_OK;
}
void pbr_fail_next_steps(pb_ctx* c, int n, int needPairs, int needManifolds) { c->failSteps = n; c->needPairs = needPairs; c->needManifolds = needManifolds; }
void pbr_caps(pb_ctx* c, int* out5) { out5[0] = c->caps.max_bodies; out5[1] = c->caps.max_colliders; out5[2] = c->caps.max_pairs; out5[3] = c->caps.max_manifolds; out5[4] = c->caps.max_joints; }
void pbr_rows(pb_ctx* c, int* entity) { std::copy(c->entity.begin(), c->entity.end(), entity); }
void pbr_colliders(pb_ctx* c, int* row, int* idx, int* type, float* tag2) {
    for (int i = 0; i < c->nCol; ++i) { row[i] = c->colRow[i]; idx[i] = c->colIdx[i]; type[i] = c->colType[i]; tag2[2 * i] = c->bounds[6 * (size_t)i]; tag2[2 * i + 1] = c->bounds[6 * (size_t)i + 1]; }
}
int pbr_map(pb_ctx* c, int which, int cap, int* out) {      // 0 bounds map, 1 contact-cache map, 2 joint-state map of the last carry-over
    const std::vector<int>& m = which == 0 ? c->lastBoundsMap : which == 1 ? c->lastCacheMap : c->lastJointMap;
    std::copy(m.begin(), m.begin() + std::min((size_t)cap, m.size()), out);
    return (int)m.size();
}
void pbr_joints(pb_ctx* c, int* r0, int* r1, int* color) {
    std::copy(c->jointRow0.begin(), c->jointRow0.end(), r0); std::copy(c->jointRow1.begin(), c->jointRow1.end(), r1); std::copy(c->jointColor.begin(), c->jointColor.end(), color);
}

} // extern "C"

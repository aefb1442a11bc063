// physecs::Scene -- host side of the B200 path (declared in include/Physecs/detail/b200_scene.hpp).
//
// Mirrors the reference's Scene (src/Physecs.cpp): same EnTT signal hooks (:92-98), same per-step contract (:112-561),
// same joint-graph colouring (:690-710), nonCollidingPairs (:788-793), trigger enter / exit diff (:538-552).  The
// arithmetic of the step is not here: this file walks the registry, keeps the device scene description in sync and calls
// the C ABI of include/physecs_b200.h.  Registry walks are the only O(bodies) host work per step; they run on the
// Scene's worker threads and go through pinned staging buffers.
#include <Physecs.h>
#include <MassUtil.h>
#include "../../include/physecs_b200.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace physecs {

ContactType defaultContactFilter(bool isTrigger0, int, bool isTrigger1, int) {
    return (isTrigger0 || isTrigger1) ? TRIGGER : COLLISION;
}

namespace {

// ---- fork-join helper for the gather / scatter loops ---------------------------------------------------------------------
class Workers {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cvStart, cvDone;
    std::function<void(size_t, size_t)> job;
    size_t total = 0, chunk = 0;
    std::atomic<size_t> next{0};
    int generation = 0, running = 0;
    bool quit = false;

    void drain() {
        for (;;) {
            size_t b = next.fetch_add(chunk);
            if (b >= total) return;
            job(b, std::min(total, b + chunk));
        }
    }
    void loop() {
        int seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cvStart.wait(lk, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
            lk.unlock();
            drain();
            lk.lock();
            if (--running == 0) cvDone.notify_one();
        }
    }

public:
    explicit Workers(int n) { for (int i = 0; i < n; ++i) threads.emplace_back([this] { loop(); }); }
    ~Workers() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cvStart.notify_all();
        for (auto& t : threads) t.join();
    }
    // fn(begin, end) over [0, n); the caller participates
    void parallelFor(size_t n, const std::function<void(size_t, size_t)>& fn) {
        if (n == 0) return;
        if (threads.empty() || n < 4096) { fn(0, n); return; }
        {
            std::lock_guard<std::mutex> lk(m);
            job = fn; total = n; next = 0;
            chunk = std::max<size_t>(1024, n / ((threads.size() + 1) * 8));
            running = (int)threads.size();
            ++generation;
        }
        cvStart.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(m);
        cvDone.wait(lk, [&] { return running == 0; });
    }
};

template <class T> struct Pinned {
    T* p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) pb_host_free(p);
        p = nullptr; cap = 0;
        void* q = nullptr;
        size_t want = std::max<size_t>(n + n / 4, 256);
        if (pb_host_alloc(&q, sizeof(T) * want) != PB_OK) throw std::runtime_error("physecs_b200: pinned host allocation failed");
        p = (T*)q; cap = want;
    }
    ~Pinned() { if (p) pb_host_free(p); }
    // enough of std::vector's face for the host mirrors (contents are NOT kept when the buffer grows: every user refills it)
    void resize(size_t n) { reserve(n); }
    T* data() { return p; }
    T& operator[](size_t i) { return p[i]; }
};

inline unsigned long long colKey(entt::entity e, int idx) { return ((unsigned long long)entt::to_integral(e) << 20) ^ (unsigned long long)(unsigned)idx; }

using Clock = std::chrono::steady_clock;
inline double msSince(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

} // namespace

struct Scene::Impl {
    bool queryShape(const Geometry& geometry, float prm[4], int& mesh);
    explicit Impl(int threads) : workers(std::max(threads, 0)) {}

    Workers workers;
    int device = 0;
    pb_ctx* ctx = nullptr;
    pb_caps caps{};
    SyncMode syncMode = SYNC_FULL;
    int wantPairs = 0, wantManifolds = 0;       // user-chosen initial arena sizes (0 = scale with the collider count)

    // rows: [0, nDyn) = packed order of the RigidBodyDynamicComponent storage (the reference's b0 / b1 index space,
    // Physecs.cpp:116-117), then the static rows (collision component, no dynamic component)
    std::vector<entt::entity> rowEntity;
    std::vector<int> entityRow;                 // entt::to_entity(e) -> row, -1 = none
    std::vector<unsigned> rowTransformIdx;      // cached packed index into the TransformComponent storage, validated on use
    int nDyn = 0, nStatic = 0;
    struct ColRef { entt::entity e; int idx; };
    std::vector<ColRef> cols;                   // device collider order

    bool topologyDirty = true, jointsDirty = true, pairsDirty = true, filterDirty = true, ctxFresh = true;
    bool stagingValid = false;                  // pinned buffers hold the current state of every dynamic row
    std::unordered_set<unsigned long long> freshCols;   // colliders created since the last upload: creation-time bounds (no margin)
    struct Move { entt::entity e; glm::vec3 p; glm::quat q; };
    std::vector<Move> moved;                    // registry.patch<TransformComponent> since the last step
    std::unordered_set<unsigned> touched;       // device-authoritative mode: dynamic rows to re-read

    std::unordered_map<const ConvexMesh*, int> convexHandle;
    std::unordered_map<const TriangleMesh*, int> trimeshHandle;

    // pinned staging
    Pinned<float> hPos, hQuat, hVel, hAng, hSPos, hSQuat;
    Pinned<int> upI; Pinned<float> upF;         // tables of a scene re-upload (prepareDevice)
    // host mirrors for change detection of fields the reference reads live every step
    Pinned<int> kin;                            // (pinned: they are upload sources of every re-upload)
    Pinned<float> invMass, com, invI;

    // joints
    std::vector<Joint*> joints;                 // creation order
    std::vector<Joint*> uploadedJoints;         // joint list of the last pb_upload_joints (state carry-over)
    std::unordered_map<entt::entity, unsigned char> jointBits;   // JointGraph::bitsets (Physecs.h:150)
    std::set<std::pair<unsigned, unsigned>> nonColliding;         // (lower, higher) entity integers

    // triggers
    std::vector<std::array<int, 4>> triggerCache;
    std::vector<OnTriggerEnterListener*> enterListeners;
    std::vector<OnTriggerExitListener*> exitListeners;
    ContactType (*filter)(bool, int, bool, int) = defaultContactFilter;

    std::vector<glm::vec3> contactPoints;
    std::vector<BVHNode> bvhSnapshot; int bvhRoot = -1;
    StepStats stats{};

    [[noreturn]] void fail(const char* what, int rc) {
        std::string msg = std::string("physecs_b200: ") + what + " failed (" + std::to_string(rc) + ")";
        if (ctx) { msg += ": "; msg += pb_last_error(ctx); }
        throw std::runtime_error(msg);
    }
    void check(int rc, const char* what) { if (rc != PB_OK) fail(what, rc); }
    int rowOf(entt::entity e) const {
        auto i = (size_t)entt::to_entity(e);
        if (i >= entityRow.size()) return -1;
        int r = entityRow[i];
        return (r >= 0 && rowEntity[r] == e) ? r : -1;
    }
};

// ---- construction / signals (reference Physecs.cpp:25-98, :815-821) ------------------------------------------------------------
Scene::Scene(entt::registry& registry, int numThreads) : registry(registry), impl(new Impl(numThreads)) {
    registry.on_construct<RigidBodyCollisionComponent>().connect<&Scene::onRigidBodyCreate>(this);
    registry.on_destroy<RigidBodyCollisionComponent>().connect<&Scene::onRigidBodyDelete>(this);
    registry.on_update<RigidBodyCollisionComponent>().connect<&Scene::onRigidBodyUpdate>(this);
    registry.on_update<TransformComponent>().connect<&Scene::onRigidBodyMove>(this);
    registry.on_construct<RigidBodyDynamicComponent>().connect<&Scene::onDynamicCreate>(this);
    registry.on_destroy<RigidBodyDynamicComponent>().connect<&Scene::onDynamicDelete>(this);
    // bodies that already exist are picked up by the first rebuild (the reference would miss them; harmless superset)
}

Scene::~Scene() {
    registry.on_construct<RigidBodyCollisionComponent>().disconnect<&Scene::onRigidBodyCreate>(this);
    registry.on_destroy<RigidBodyCollisionComponent>().disconnect<&Scene::onRigidBodyDelete>(this);
    registry.on_update<RigidBodyCollisionComponent>().disconnect<&Scene::onRigidBodyUpdate>(this);
    registry.on_update<TransformComponent>().disconnect<&Scene::onRigidBodyMove>(this);
    registry.on_construct<RigidBodyDynamicComponent>().disconnect<&Scene::onDynamicCreate>(this);
    registry.on_destroy<RigidBodyDynamicComponent>().disconnect<&Scene::onDynamicDelete>(this);
    if (impl->ctx) pb_ctx_destroy(impl->ctx);
    // joints still alive are leaked exactly like the reference leaks them (Physecs.cpp:815-821 never deletes them)
}

void Scene::onRigidBodyCreate(entt::registry& reg, entt::entity e) {
    impl->topologyDirty = true;
    if (impl->ctxFresh) return;      // nothing is on the device yet: no history a new collider could inherit (a 1 M-entity initial fill skips 1 M set inserts)
    auto& c = reg.get<RigidBodyCollisionComponent>(e);
    for (int i = 0; i < (int)c.colliders.size(); ++i) impl->freshCols.insert(colKey(e, i));
}
void Scene::onRigidBodyDelete(entt::registry&, entt::entity) { impl->topologyDirty = true; }
// colliders edited in place and announced with registry.patch / replace<RigidBodyCollisionComponent>: the reference reads
// collider geometry live in its narrowphase; the device copy is refreshed instead (bounds keep their history)
void Scene::onRigidBodyUpdate(entt::registry&, entt::entity) { impl->topologyDirty = true; }
void Scene::onDynamicCreate(entt::registry&, entt::entity) { impl->topologyDirty = true; }
void Scene::onDynamicDelete(entt::registry&, entt::entity) { impl->topologyDirty = true; }

// registry.patch / replace<TransformComponent> (reference onRigidBodyMove -> updateBounds, Physecs.cpp:51-54, :79-90):
// the bounds of the entity's colliders are recomputed from the transform AT PATCH TIME with the 0.01 margin
void Scene::onRigidBodyMove(entt::registry& reg, entt::entity e) {
    if (!reg.any_of<RigidBodyCollisionComponent>(e)) return;
    auto& t = reg.get<TransformComponent>(e);
    impl->moved.push_back({ e, t.position, t.orientation });
    int r = impl->rowOf(e);
    if (r >= 0 && r < impl->nDyn) impl->touched.insert((unsigned)r);
}

void Scene::setDevice(int cudaDevice) {
    if (impl->ctx) throw std::runtime_error("physecs_b200: setDevice after the first simulate()");
    impl->device = cudaDevice;
}
void Scene::setSyncMode(SyncMode mode) { impl->syncMode = mode; }
void Scene::notifyBodyChanged(entt::entity e) {
    int r = impl->rowOf(e);
    if (r >= 0 && r < impl->nDyn) impl->touched.insert((unsigned)r);
}
void Scene::setArenaCapacity(int maxPairs, int maxManifolds) { impl->wantPairs = maxPairs; impl->wantManifolds = maxManifolds; }
pb_ctx* Scene::nativeContext() { return impl->ctx; }
Scene::StepStats Scene::getLastStepStats() const { return impl->stats; }

// ---- joints (reference addJoint / destroyJoint, Physecs.cpp:690-723) -----------------------------------------------------------
void Scene::addJoint(Joint* joint) {
    auto e0 = joint->getEntity0(), e1 = joint->getEntity1();
    unsigned a = entt::to_integral(e0), b = entt::to_integral(e1);
    impl->nonColliding.insert(a < b ? std::make_pair(a, b) : std::make_pair(b, a));
    impl->pairsDirty = true;
    unsigned char& c0 = impl->jointBits[e0];
    unsigned char& c1 = impl->jointBits[e1];
    // lowest colour free on both entities; the union promotes to int so index 8 (the overflow bucket) is always
    // available and consumes no bit (quirk Q21)
    unsigned un = ~(unsigned)(c0 | c1);
    int i = __builtin_ctz(un);
    c0 = (unsigned char)(c0 | (1u << i));
    c1 = (unsigned char)(c1 | (1u << i));
    joint->setColor(i);
    impl->joints.push_back(joint);
    impl->jointsDirty = true;
}

void Scene::destroyJoint(Joint* joint) {
    auto it = std::find(impl->joints.begin(), impl->joints.end(), joint);
    if (it == impl->joints.end()) return;
    impl->joints.erase(it);
    std::replace(impl->uploadedJoints.begin(), impl->uploadedJoints.end(), joint, (Joint*)nullptr);   // the address may be reused
    auto e0 = joint->getEntity0(), e1 = joint->getEntity1();
    unsigned a = entt::to_integral(e0), b = entt::to_integral(e1);
    impl->nonColliding.erase(a < b ? std::make_pair(a, b) : std::make_pair(b, a));
    impl->jointBits[e0] &= (unsigned char)~(1u << joint->getColor());
    impl->jointBits[e1] &= (unsigned char)~(1u << joint->getColor());
    impl->pairsDirty = impl->jointsDirty = true;
    delete joint;
}

// ---- collider / body edits (reference Physecs.cpp:725-770, :788-797) -----------------------------------------------------------
void Scene::clearColliders(entt::entity entity) {
    auto& col = registry.get<RigidBodyCollisionComponent>(entity);
    col.colliders.clear();
    impl->topologyDirty = true;
}

void Scene::addCollider(entt::entity entity, const Collider& collider) {
    auto& col = registry.get<RigidBodyCollisionComponent>(entity);
    if (!impl->ctxFresh) impl->freshCols.insert(colKey(entity, (int)col.colliders.size()));
    col.colliders.push_back(collider);
    impl->topologyDirty = true;
}

void Scene::setIsKinematic(entt::entity entity, bool isKinematic) {
    auto& d = registry.get<RigidBodyDynamicComponent>(entity);
    d.isKinematic = isKinematic;
    if (isKinematic) { d.velocity = glm::vec3(0); d.angularVelocity = glm::vec3(0); }
    notifyBodyChanged(entity);   // the kinematic flag itself is picked up by the per-step mirror comparison
}

void Scene::addOnTriggerEnterCallback(OnTriggerEnterListener* cb) { impl->enterListeners.push_back(cb); }
void Scene::addOnTriggerExitCallback(OnTriggerExitListener* cb) { impl->exitListeners.push_back(cb); }
void Scene::removeOnTriggerEnterCallback(OnTriggerEnterListener* cb) {
    auto it = std::find(impl->enterListeners.begin(), impl->enterListeners.end(), cb);
    if (it != impl->enterListeners.end()) impl->enterListeners.erase(it);
}
void Scene::removeOnTriggerExitCallback(OnTriggerExitListener* cb) {
    auto it = std::find(impl->exitListeners.begin(), impl->exitListeners.end(), cb);
    if (it != impl->exitListeners.end()) impl->exitListeners.erase(it);
}

void Scene::setCanCollide(entt::entity entity0, entt::entity entity1, bool canCollide) {
    unsigned a = entt::to_integral(entity0), b = entt::to_integral(entity1);
    auto key = a < b ? std::make_pair(a, b) : std::make_pair(b, a);
    if (canCollide) impl->nonColliding.erase(key); else impl->nonColliding.insert(key);
    impl->pairsDirty = true;
}

void Scene::setContactFilter(ContactType (*filter)(bool, int, bool, int)) {
    impl->filter = filter ? filter : defaultContactFilter;
    impl->filterDirty = true;
}

const std::vector<glm::vec3>& Scene::getContactPoints() {
    auto& out = impl->contactPoints;
    out.clear();
    if (!impl->ctx) return out;
    pb_counts c{};
    pb_get_counts(impl->ctx, &c);
    int n = c.n_manifolds;
    if (n <= 0) return out;
    std::vector<int> keys((size_t)5 * n), np(n), color(n);
    std::vector<float> nrm((size_t)3 * n), pts((size_t)24 * n);
    int got = 0;
    impl->check(pb_get_manifolds(impl->ctx, n, keys.data(), np.data(), nrm.data(), pts.data(), color.data(), &got), "pb_get_manifolds");
    for (int m = 0; m < got; ++m)
        for (int k = 0; k < np[m]; ++k) {
            const float* p = &pts[(size_t)24 * m + 6 * k];
            out.emplace_back(p[0], p[1], p[2]);
            out.emplace_back(p[3], p[4], p[5]);
        }
    return out;
}

const std::vector<BVHNode>& Scene::getBVH() {
    Impl& S = *impl;
    auto& out = S.bvhSnapshot;
    out.clear();
    S.bvhRoot = BVH::null;
    prepareDevice();
    if (!S.ctx) return out;
    int nInt = 0;
    S.check(pb_get_tree(S.ctx, 0, nullptr, nullptr, &nInt), "pb_get_tree");
    auto leafNode = [&](int col, const float* box, int parent) {
        BVHNode n;
        n.bounds.min = glm::vec3(box[0], box[1], box[2]); n.bounds.max = glm::vec3(box[3], box[4], box[5]);
        n.parent = parent; n.isLeaf = true;
        int e = 0, ci = 0;
        S.check(pb_collider_ids(S.ctx, 1, &col, &e, &ci), "pb_collider_ids");
        n.leaf.entity = (entt::entity)(unsigned)e; n.leaf.colliderIndex = ci;
        return n;
    };
    if (nInt == 0) {
        // zero or one collider: the tree is that leaf
        int n = (int)S.cols.size();
        if (n == 1) {
            std::vector<float> b(6);
            S.check(pb_get_bounds(S.ctx, b.data()), "pb_get_bounds");
            out.push_back(leafNode(0, b.data(), BVH::null));
            S.bvhRoot = 0;
        }
        return out;
    }
    std::vector<float> boxes((size_t)12 * nInt); std::vector<int> links((size_t)2 * nInt);
    S.check(pb_get_tree(S.ctx, nInt, boxes.data(), links.data(), &nInt), "pb_get_tree");
    out.resize(nInt);
    for (int i = 0; i < nInt; ++i) {
        BVHNode& n = out[i];
        n.isLeaf = false;
        if (i == 0) n.parent = BVH::null;
        const float* b = &boxes[(size_t)12 * i];
        n.bounds.min = glm::min(glm::vec3(b[0], b[1], b[2]), glm::vec3(b[6], b[7], b[8]));
        n.bounds.max = glm::max(glm::vec3(b[3], b[4], b[5]), glm::vec3(b[9], b[10], b[11]));
    }
    for (int i = 0; i < nInt; ++i)
        for (int side = 0; side < 2; ++side) {
            int link = links[2 * i + side], id;
            if (link >= 0) { id = link; out[link].parent = i; }
            else { id = (int)out.size(); out.push_back(leafNode(-1 - link, &boxes[(size_t)12 * i + 6 * side], i)); }
            if (side == 0) out[i].internal.left = id; else out[i].internal.right = id;
        }
    S.bvhRoot = 0;
    return out;
}

const int Scene::getBHVRootId() {
    if (impl->bvhSnapshot.empty()) getBVH();
    return impl->bvhRoot;
}

// ---- scene queries (reference Physecs.cpp:571-688) -------------------------------------------------------------------------------
// They run on the device tree over the state of the last simulate() plus everything announced since (structural edits,
// patched transforms): the per-body registry gather of simulate() is not repeated for a query.
entt::entity Scene::raycastClosest(glm::vec3 rayOrig, glm::vec3 rayDir, float maxDistance, const std::function<bool(entt::entity)>& filter, glm::vec3* hitPos) {
    Impl& S = *impl;
    prepareDevice();
    int cap = 256;
    std::vector<int> ray, ent, col; std::vector<float> t;
    int n = 0;
    for (;;) {
        ray.resize(cap); ent.resize(cap); col.resize(cap); t.resize(cap);
        S.check(pb_query_raycast(S.ctx, 1, &rayOrig.x, &rayDir.x, maxDistance, cap, ray.data(), ent.data(), col.data(), t.data(), &n), "pb_query_raycast");
        if (n <= cap) break;
        cap = n + 64;
    }
    // the closest hit whose entity passes the filter (the reference evaluates the filter per leaf and keeps the nearer subtree hit)
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return t[a] != t[b] ? t[a] < t[b] : (unsigned)ent[a] < (unsigned)ent[b]; });
    for (int i : order) {
        entt::entity e = (entt::entity)(unsigned)ent[i];
        if (filter && !filter(e)) continue;
        if (hitPos) *hitPos = rayOrig + rayDir * t[i];
        return e;
    }
    return entt::null;
}

// The reference's unfiltered overload hands an EMPTY std::function to the walk, whose `filter && filter(e)` test then rejects
// every hit (Physecs.cpp:582, :612): it can only return entt::null.  Here the overload accepts every entity, which is what its
// signature promises; use the filtered overload for bit-for-bit reference behaviour.
entt::entity Scene::raycastClosest(glm::vec3 rayOrig, glm::vec3 rayDir, float maxDistance, glm::vec3* hitPos) {
    return raycastClosest(rayOrig, rayDir, maxDistance, std::function<bool(entt::entity)>(), hitPos);
}

// query shape -> (params, convex handle) of the C ABI; false for a triangle mesh (no query routine takes one)
bool Scene::Impl::queryShape(const Geometry& geometry, float prm[4], int& mesh) {
    Impl& S = *this;
    prm[0] = prm[1] = prm[2] = prm[3] = 0.f;
    mesh = -1;
    switch (geometry.type) {
        case SPHERE: prm[0] = geometry.sphere.radius; break;
        case CAPSULE: prm[0] = geometry.capsule.halfHeight; prm[1] = geometry.capsule.radius; break;
        case BOX: std::memcpy(prm, &geometry.box.halfExtents, sizeof(float) * 3); break;
        case CONVEX_MESH: {
            std::memcpy(prm, &geometry.convex.scale, sizeof(float) * 3);
            auto it = S.convexHandle.find(geometry.convex.mesh);
            if (it == S.convexHandle.end()) {
                // a convex mesh no collider uses yet: register it with the context for the query
                const ConvexMesh* m = geometry.convex.mesh;
                std::vector<float> v((size_t)3 * m->vertices.size()), fn((size_t)3 * m->faces.size()), fc((size_t)3 * m->faces.size());
                for (int i = 0; i < m->vertices.size(); ++i) std::memcpy(&v[3 * i], &m->vertices[i], sizeof(float) * 3);
                std::vector<int> off(m->faces.size() + 1, 0), idx;
                for (size_t f = 0; f < m->faces.size(); ++f) {
                    idx.insert(idx.end(), m->faces[f].indices.begin(), m->faces[f].indices.end());
                    off[f + 1] = (int)idx.size();
                    std::memcpy(&fn[3 * f], &m->faces[f].normal, sizeof(float) * 3);
                    std::memcpy(&fc[3 * f], &m->faces[f].centroid, sizeof(float) * 3);
                }
                int h = -1;
                S.check(pb_register_convex(S.ctx, v.data(), m->vertices.size(), off.data(), idx.data(), (int)m->faces.size(), fn.data(), fc.data(), &h), "pb_register_convex");
                it = S.convexHandle.emplace(m, h).first;
            }
            mesh = it->second;
        } break;
        case TRIANGLE_MESH: return false;
    }
    return true;
}

std::vector<OverlapHit> Scene::overlap(glm::vec3 pos, glm::quat ori, Geometry geometry, int filter) {
    Impl& S = *impl;
    prepareDevice();
    float prm[4];
    int mesh;
    if (!S.queryShape(geometry, prm, mesh)) return {};      // physecs::overlap has no triangle-mesh case
    float q[4] = { ori.x, ori.y, ori.z, ori.w };
    int cap = 256, n = 0;
    std::vector<int> ent, col;
    for (;;) {
        ent.resize(cap); col.resize(cap);
        S.check(pb_query_overlap(S.ctx, &pos.x, q, (int)geometry.type, prm, mesh, filter, cap, ent.data(), col.data(), &n), "pb_query_overlap");
        if (n <= cap) break;
        cap = n + 64;
    }
    std::vector<OverlapHit> out((size_t)n);
    for (int i = 0; i < n; ++i) out[i] = { (entt::entity)(unsigned)ent[i], col[i] };
    std::sort(out.begin(), out.end(), [](const OverlapHit& a, const OverlapHit& b) {
        return a.entity != b.entity ? entt::to_integral(a.entity) < entt::to_integral(b.entity) : a.colIndex < b.colIndex; });
    return out;
}

// == reference Scene::overlapWithMinTranslationalDistance (Physecs.cpp:652-688): physecs::collision(collider, query shape) on the
// device for every collider whose bounds meet the query's; one hit per manifold that has points (a triangle-mesh collider gives
// one per touched triangle).  Sorted by (entity, collider index); the hits of one mesh collider keep their generation order.
std::vector<OverlapMtdHit> Scene::overlapWithMinTranslationalDistance(glm::vec3 pos, glm::quat ori, Geometry geometry) {
    Impl& S = *impl;
    prepareDevice();
    float prm[4];
    int mesh;
    if (!S.queryShape(geometry, prm, mesh)) throw std::runtime_error("physecs_b200: a triangle mesh cannot be the query shape of overlapWithMinTranslationalDistance");
    float q[4] = { ori.x, ori.y, ori.z, ori.w };
    int cap = 256, n = 0;
    std::vector<int> ent, col; std::vector<float> nrm, mtd;
    for (;;) {
        ent.resize(cap); col.resize(cap); nrm.resize((size_t)3 * cap); mtd.resize(cap);
        S.check(pb_query_overlap_mtd(S.ctx, &pos.x, q, (int)geometry.type, prm, mesh, cap, ent.data(), col.data(), nrm.data(), mtd.data(), &n), "pb_query_overlap_mtd");
        if (n <= cap) break;
        cap = n + 64;
    }
    std::vector<OverlapMtdHit> out((size_t)n);
    for (int i = 0; i < n; ++i) out[i] = { (entt::entity)(unsigned)ent[i], col[i], glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), mtd[i] };
    std::stable_sort(out.begin(), out.end(), [](const OverlapMtdHit& a, const OverlapMtdHit& b) {
        return a.entity != b.entity ? entt::to_integral(a.entity) < entt::to_integral(b.entity) : a.colIndex < b.colIndex; });
    return out;
}

// ---- the step ----------------------------------------------------------------------------------------------------------------
namespace {

template <class T>
inline T& packedAt(entt::storage_for_t<T>& st, size_t pos) {
    constexpr size_t page = entt::component_traits<T>::page_size;
    return st.raw()[pos / page][pos % page];
}

} // namespace

// Bring the device scene description up to date with everything the signals / API calls recorded since the last step:
// structural changes, joints, non-colliding pairs, contact filter, patched transforms.  Shared by simulate() and the queries.
void Scene::prepareDevice() {
    Impl& S = *impl;
    auto& dynStore = registry.storage<RigidBodyDynamicComponent>();
    auto& colStore = registry.storage<RigidBodyCollisionComponent>();
    auto& trStore = registry.storage<TransformComponent>();

    // The body rows follow the packed order of the dynamic-body pool (the reference indexes bodies by their position in it, Physecs.cpp:116-117).
    // No signal tells of a registry.sort on it: compare the pool's entity list with the rows' (4 B per body, on the workers).
    if (!S.topologyDirty && S.ctx) {
        if (dynStore.size() != (size_t)S.nDyn) S.topologyDirty = true;
        else {
            std::atomic<int> differs{0};
            const entt::entity* pool = dynStore.data();
            S.workers.parallelFor((size_t)S.nDyn, [&](size_t b, size_t e) {
                if (std::memcmp(pool + b, S.rowEntity.data() + b, sizeof(entt::entity) * (e - b))) differs.store(1, std::memory_order_relaxed);
            });
            if (differs.load()) S.topologyDirty = true;
        }
    }

    // ---- (re)build the device scene description after structural changes -----------------------------------------------------
    auto rebuild = [&](bool growCaps) {
        // PB_TRACE_EDIT=1: where a re-upload spends its time, one line on stderr
        static const bool trace = std::getenv("PB_TRACE_EDIT") != nullptr;
        auto tPhase = Clock::now();
        std::string traceLine;
        auto phase = [&](const char* name) {
            if (!trace) return;
            char buf[64]; std::snprintf(buf, sizeof buf, " %s %.1f", name, msSince(tPhase));
            traceLine += buf; tPhase = Clock::now();
        };
        // What the device holds now.  Colliders that survive the re-upload keep their bounds history (creation bounds have no margin,
        // refreshed ones do: quirk Q6) and their contact-cache entries.  The colliders of an entity are contiguous in device order, so
        // "where was (entity, index) before" is one flat table over entity indices (entity versions are compared through oldCols) --
        // at 1 M colliders a hash map per lookup was most of a structural edit's cost.
        std::vector<Impl::ColRef> oldCols;
        std::vector<int> oldFirst;                  // entt::to_entity(e) -> first collider of e in oldCols, -1 = none
        if (S.ctx && !S.ctxFresh && !S.cols.empty()) {
            oldCols.swap(S.cols);
            size_t top = 0;
            for (auto& c : oldCols) top = std::max(top, (size_t)entt::to_entity(c.e));
            oldFirst.assign(top + 1, -1);
            for (size_t i = oldCols.size(); i-- > 0;) oldFirst[(size_t)entt::to_entity(oldCols[i].e)] = (int)i;
        }
        auto oldIndexOf = [&](entt::entity e, int idx) -> int {
            const size_t k = (size_t)entt::to_entity(e);
            if (k >= oldFirst.size() || oldFirst[k] < 0) return -1;
            const size_t i = (size_t)oldFirst[k] + (size_t)idx;
            if (i >= oldCols.size() || oldCols[i].e != e || oldCols[i].idx != idx) return -1;
            return (int)i;
        };
        bool carryCache = !oldCols.empty();
        phase("old-table");

        // ---- rows -----------------------------------------------------------------------------------------------------------------
        const size_t nDyn = dynStore.size();
        S.rowEntity.clear();
        S.rowEntity.reserve(nDyn + colStore.size());
        const entt::entity* dynEnts = dynStore.data();
        S.rowEntity.insert(S.rowEntity.end(), dynEnts, dynEnts + nDyn);
        const entt::entity* colEnts = colStore.data();
        for (size_t i = 0; i < colStore.size(); ++i)
            if (!dynStore.contains(colEnts[i])) S.rowEntity.push_back(colEnts[i]);
        S.nDyn = (int)nDyn;
        S.nStatic = (int)S.rowEntity.size() - S.nDyn;
        const int rows = (int)S.rowEntity.size();
        size_t maxIdx = 0;
        for (auto e : S.rowEntity) maxIdx = std::max(maxIdx, (size_t)entt::to_entity(e));
        S.entityRow.assign(rows ? maxIdx + 1 : 0, -1);
        S.rowTransformIdx.resize(rows);
        phase("rows");
        // ---- colliders, row-major: a count per row, offsets, then the list -- on the worker threads, like every other walk over the
        // registry here; the meshes the colliders reference are noted on the way (with the row that named them first)
        std::vector<int> colStart((size_t)rows + 1, 0);
        std::vector<std::pair<size_t, const ConvexMesh*>> convexSeen;
        std::vector<std::pair<size_t, const TriangleMesh*>> trimeshSeen;
        std::mutex seenLock;
        S.workers.parallelFor((size_t)rows, [&](size_t rb, size_t re) {
            std::vector<std::pair<size_t, const ConvexMesh*>> cv;
            std::vector<std::pair<size_t, const TriangleMesh*>> tm;
            for (size_t r = rb; r < re; ++r) {
                const auto e = S.rowEntity[r];
                S.entityRow[(size_t)entt::to_entity(e)] = (int)r;
                S.rowTransformIdx[r] = (unsigned)trStore.index(e);
                if (r < nDyn && !colStore.contains(e)) continue;
                const auto& cc = colStore.get(e).colliders;
                colStart[r + 1] = (int)cc.size();
                for (const Collider& c : cc) {
                    if (c.geometry.type == CONVEX_MESH) {
                        const ConvexMesh* m = c.geometry.convex.mesh;
                        if (std::find_if(cv.begin(), cv.end(), [m](auto& s) { return s.second == m; }) == cv.end()) cv.emplace_back(r, m);
                    } else if (c.geometry.type == TRIANGLE_MESH) {
                        const TriangleMesh* m = c.geometry.triangleMesh.mesh;
                        if (std::find_if(tm.begin(), tm.end(), [m](auto& s) { return s.second == m; }) == tm.end()) tm.emplace_back(r, m);
                    }
                }
            }
            if (cv.empty() && tm.empty()) return;
            std::lock_guard<std::mutex> lk(seenLock);
            convexSeen.insert(convexSeen.end(), cv.begin(), cv.end());
            trimeshSeen.insert(trimeshSeen.end(), tm.begin(), tm.end());
        });
        for (int r = 0; r < rows; ++r) colStart[r + 1] += colStart[r];
        const int nCol = colStart[rows];
        S.cols.resize((size_t)nCol);
        S.workers.parallelFor((size_t)rows, [&](size_t rb, size_t re) {
            for (size_t r = rb; r < re; ++r)
                for (int i = 0, n = colStart[r + 1] - colStart[r]; i < n; ++i) S.cols[(size_t)colStart[r] + i] = { S.rowEntity[r], i };
        });

        phase("collider-list");
        // capacities: context is sized with headroom and re-created when the scene outgrows it.  The bounds history travels on the
        // device while the context lives (pb_keep_bounds_begin / pb_keep_bounds) and through the host when it is replaced.
        auto roomy = [](int n) { return std::max(1024, n + n / 2); };
        bool need = !S.ctx || rows > S.caps.max_bodies || nCol > S.caps.max_colliders || (int)S.joints.size() > S.caps.max_joints || growCaps;
        std::vector<float> oldBounds;
        if (need) {
            if (!oldCols.empty()) {
                oldBounds.resize((size_t)6 * oldCols.size());
                S.check(pb_get_bounds(S.ctx, oldBounds.data()), "pb_get_bounds");
            }
            pb_caps c{};
            c.max_bodies = roomy(rows);
            c.max_colliders = roomy(nCol);
            c.max_pairs = S.wantPairs > 0 ? S.wantPairs : 8 * c.max_colliders + 4096;
            c.max_manifolds = S.wantManifolds > 0 ? S.wantManifolds : 6 * c.max_colliders + 4096;
            c.max_joints = roomy((int)S.joints.size());
            if (S.ctx) { pb_ctx_destroy(S.ctx); S.ctx = nullptr; }
            S.convexHandle.clear(); S.trimeshHandle.clear();
            int rc = pb_ctx_create(S.device, &c, &S.ctx);
            if (rc != PB_OK) { S.ctx = nullptr; S.fail("pb_ctx_create (a CUDA device is required: there is no CPU fallback)", rc); }
            S.caps = c;
            S.jointsDirty = S.pairsDirty = S.filterDirty = true;
            S.uploadedJoints.clear();
            carryCache = false;      // a new context starts with an empty contact cache
        } else if (!oldCols.empty()) {
            S.check(pb_keep_bounds_begin(S.ctx), "pb_keep_bounds_begin");
        }

        phase("context");
        // meshes referenced by colliders, registered in the order the colliders name them
        std::sort(convexSeen.begin(), convexSeen.end());
        std::sort(trimeshSeen.begin(), trimeshSeen.end());
        for (auto& seen : convexSeen) {
            const ConvexMesh* m = seen.second;
            if (S.convexHandle.count(m)) continue;
            std::vector<float> v((size_t)3 * m->vertices.size()), fn((size_t)3 * m->faces.size()), fc((size_t)3 * m->faces.size());
            for (int i = 0; i < m->vertices.size(); ++i) std::memcpy(&v[3 * i], &m->vertices[i], sizeof(float) * 3);
            std::vector<int> off(m->faces.size() + 1, 0), idx;
            for (size_t f = 0; f < m->faces.size(); ++f) {
                idx.insert(idx.end(), m->faces[f].indices.begin(), m->faces[f].indices.end());
                off[f + 1] = (int)idx.size();
                std::memcpy(&fn[3 * f], &m->faces[f].normal, sizeof(float) * 3);
                std::memcpy(&fc[3 * f], &m->faces[f].centroid, sizeof(float) * 3);
            }
            int h = -1;
            S.check(pb_register_convex(S.ctx, v.data(), m->vertices.size(), off.data(), idx.data(), (int)m->faces.size(), fn.data(), fc.data(), &h), "pb_register_convex");
            S.convexHandle[m] = h;
        }
        for (auto& seen : trimeshSeen) {
            const TriangleMesh* m = seen.second;
            if (S.trimeshHandle.count(m)) continue;
            int h = -1;
            S.check(pb_register_trimesh(S.ctx, (const float*)m->vertices.data(), (int)m->vertices.size(), m->getSourceIndices().data(),
                                        (int)m->getSourceIndices().size(), &h, nullptr), "pb_register_trimesh");
            S.trimeshHandle[m] = h;
        }

        phase("meshes");
        // The body and collider tables are written into pinned memory that stays allocated between edits (fresh pageable vectors cost a
        // page fault per 4 KB and a staged copy: at 1 M bodies that was a third of an edit); each upload call returns after its copy, so
        // the two tables share the buffers.
        S.upI.reserve(std::max((size_t)rows, (size_t)6 * nCol));
        S.upF.reserve(std::max((size_t)7 * rows + 6 * nDyn, (size_t)14 * nCol));

        // bodies
        {
            int* ent = S.upI.p;
            float* pos = S.upF.p; float* quat = pos + (size_t)3 * rows; float* vel = quat + (size_t)4 * rows; float* ang = vel + 3 * nDyn;
            S.kin.resize(nDyn); S.invMass.resize(nDyn); S.com.resize(3 * nDyn); S.invI.resize(9 * nDyn);
            S.workers.parallelFor((size_t)rows, [&](size_t rb, size_t re) {
                for (size_t r = rb; r < re; ++r) {
                    ent[r] = (int)entt::to_integral(S.rowEntity[r]);
                    const TransformComponent& t = packedAt<TransformComponent>(trStore, S.rowTransformIdx[r]);
                    std::memcpy(&pos[3 * r], &t.position, sizeof(float) * 3);
                    quat[4 * r] = t.orientation.x; quat[4 * r + 1] = t.orientation.y; quat[4 * r + 2] = t.orientation.z; quat[4 * r + 3] = t.orientation.w;
                    if (r < nDyn) {
                        const RigidBodyDynamicComponent& d = packedAt<RigidBodyDynamicComponent>(dynStore, r);
                        S.kin[r] = d.isKinematic ? 1 : 0;
                        std::memcpy(&vel[3 * r], &d.velocity, sizeof(float) * 3);
                        std::memcpy(&ang[3 * r], &d.angularVelocity, sizeof(float) * 3);
                        S.invMass[r] = d.invMass;
                        std::memcpy(&S.com[3 * r], &d.com, sizeof(float) * 3);
                        std::memcpy(&S.invI[9 * r], &d.invInertiaTensor, sizeof(float) * 9);
                    }
                }
            });
            phase("body-table");
            S.check(pb_upload_bodies(S.ctx, (int)nDyn, S.nStatic, ent, pos, quat, S.kin.data(), vel, ang, S.invMass.data(), S.com.data(), S.invI.data()), "pb_upload_bodies");
            phase("pb_upload_bodies");
        }

        // colliders
        {
            const size_t C = (size_t)nCol;
            int* cRow = S.upI.p; int* cIdx = cRow + C; int* cType = cIdx + C; int* cMesh = cType + C; int* cFlags = cMesh + C; int* cData = cFlags + C;
            float* lp = S.upF.p; float* lq = lp + 3 * C; float* prm = lq + 4 * C; float* mat = prm + 4 * C;
            S.workers.parallelFor(C, [&](size_t cb, size_t ce) {
                for (size_t i = cb; i < ce; ++i) {
                    const Collider& c = colStore.get(S.cols[i].e).colliders[S.cols[i].idx];
                    cRow[i] = S.entityRow[(size_t)entt::to_entity(S.cols[i].e)];
                    cIdx[i] = S.cols[i].idx;
                    cType[i] = (int)c.geometry.type;
                    cMesh[i] = -1;
                    std::memcpy(&lp[3 * i], &c.position, sizeof(float) * 3);
                    lq[4 * i] = c.orientation.x; lq[4 * i + 1] = c.orientation.y; lq[4 * i + 2] = c.orientation.z; lq[4 * i + 3] = c.orientation.w;
                    float* p = &prm[4 * i];
                    p[0] = p[1] = p[2] = p[3] = 0.f;
                    switch (c.geometry.type) {
                        case SPHERE: p[0] = c.geometry.sphere.radius; break;
                        case CAPSULE: p[0] = c.geometry.capsule.halfHeight; p[1] = c.geometry.capsule.radius; break;
                        case BOX: std::memcpy(p, &c.geometry.box.halfExtents, sizeof(float) * 3); break;
                        case CONVEX_MESH: std::memcpy(p, &c.geometry.convex.scale, sizeof(float) * 3); cMesh[i] = S.convexHandle.at(c.geometry.convex.mesh); break;
                        case TRIANGLE_MESH: cMesh[i] = S.trimeshHandle.at(c.geometry.triangleMesh.mesh); break;
                    }
                    mat[3 * i] = c.material.friction; mat[3 * i + 1] = c.material.restitution; mat[3 * i + 2] = c.material.damping;
                    cFlags[i] = (c.isTrigger ? PB_COL_TRIGGER : 0) | (c.enableSimulation ? PB_COL_ENABLE_SIM : 0);
                    cData[i] = c.data;
                }
            });
            phase("collider-table");
            S.check(pb_upload_colliders(S.ctx, nCol, cRow, cIdx, lp, lq, cType, prm, cMesh, mat, cFlags, cData), "pb_upload_colliders");
            phase("pb_upload_colliders");
        }

        if (!oldCols.empty()) {
            // boundsMap / cacheMap[o] = where old collider o is now (-1 = gone).  Persisting contacts keep their cached restitution targets
            // across the re-upload: the reference's contactCache is keyed by (entity, collider index) pairs (Physecs.cpp:237), the device
            // table by collider rows, so it is re-keyed -- by NAME, so a collider cleared and added again finds its entries, while its
            // bounds start over from creation bounds like the reference's new BroadPhaseEntry (Physecs.cpp:739-748).
            std::vector<int> boundsMap(oldCols.size(), -1), cacheMap(oldCols.size(), -1);
            S.workers.parallelFor((size_t)nCol, [&](size_t cb, size_t ce) {
                for (size_t i = cb; i < ce; ++i) {
                    const int o = oldIndexOf(S.cols[i].e, S.cols[i].idx);
                    if (o < 0) continue;
                    cacheMap[o] = (int)i;
                    if (S.freshCols.empty() || !S.freshCols.count(colKey(S.cols[i].e, S.cols[i].idx))) boundsMap[o] = (int)i;
                }
            });
            if (!need) S.check(pb_keep_bounds(S.ctx, (int)boundsMap.size(), boundsMap.data()), "pb_keep_bounds");
            else {
                std::vector<int> which; std::vector<float> b;
                for (size_t o = 0; o < boundsMap.size(); ++o) {
                    if (boundsMap[o] < 0) continue;
                    which.push_back(boundsMap[o]);
                    b.insert(b.end(), &oldBounds[6 * o], &oldBounds[6 * o] + 6);
                }
                if (!which.empty()) S.check(pb_set_bounds(S.ctx, (int)which.size(), which.data(), b.data()), "pb_set_bounds");
            }
            if (carryCache) S.check(pb_keep_contact_cache(S.ctx, (int)cacheMap.size(), cacheMap.data()), "pb_keep_contact_cache");
        }
        phase("carry-over");
        if (trace) std::fprintf(stderr, "physecs_b200 re-upload (%d rows, %d colliders), ms:%s\n", rows, nCol, traceLine.c_str());
        S.freshCols.clear();
        S.touched.clear();
        S.stagingValid = false;
        S.topologyDirty = false;
        S.ctxFresh = false;
        S.jointsDirty = true;        // body rows may have moved
        S.filterDirty = true;        // the filter table is per collider
    };

    auto uploadJoints = [&] {
        const int n = (int)S.joints.size();
        std::vector<int> type(n), r0(n), r1(n), color(n);
        std::vector<float> a0p((size_t)3 * n), a0q((size_t)4 * n), a1p((size_t)3 * n), a1q((size_t)4 * n), prm((size_t)8 * n);
        for (int j = 0; j < n; ++j) {
            Joint* J = S.joints[j];
            type[j] = J->kind; color[j] = J->getColor();
            r0[j] = S.rowOf(J->getEntity0()); r1[j] = S.rowOf(J->getEntity1());
            if (r0[j] < 0 || r1[j] < 0) throw std::runtime_error("physecs_b200: a joint references an entity that is not a rigid body");
            glm::vec3 p0 = J->getAnchor0Pos(), p1 = J->getAnchor1Pos();
            glm::quat q0 = J->getAnchor0Or(), q1 = J->getAnchor1Or();
            std::memcpy(&a0p[3 * (size_t)j], &p0, sizeof(float) * 3); std::memcpy(&a1p[3 * (size_t)j], &p1, sizeof(float) * 3);
            a0q[4 * (size_t)j] = q0.x; a0q[4 * (size_t)j + 1] = q0.y; a0q[4 * (size_t)j + 2] = q0.z; a0q[4 * (size_t)j + 3] = q0.w;
            a1q[4 * (size_t)j] = q1.x; a1q[4 * (size_t)j + 1] = q1.y; a1q[4 * (size_t)j + 2] = q1.z; a1q[4 * (size_t)j + 3] = q1.w;
            std::memcpy(&prm[8 * (size_t)j], J->params, sizeof(float) * 8);
            J->paramsDirty = false;
        }
        S.check(pb_upload_joints(S.ctx, n, type.data(), r0.data(), r1.data(), a0p.data(), a0q.data(), a1p.data(), a1q.data(), prm.data(), color.data()), "pb_upload_joints");
        if (!S.uploadedJoints.empty() && n) {
            // joints that were already on the device keep their persistent state (gear angle tracking)
            std::unordered_map<const Joint*, int> was;
            for (size_t j = 0; j < S.uploadedJoints.size(); ++j) was.emplace(S.uploadedJoints[j], (int)j);
            std::vector<int> oldIndex(n, -1);
            for (int j = 0; j < n; ++j) { auto it = was.find(S.joints[j]); if (it != was.end()) oldIndex[j] = it->second; }
            S.check(pb_keep_joint_state(S.ctx, n, oldIndex.data()), "pb_keep_joint_state");
        }
        S.uploadedJoints = S.joints;
        S.jointsDirty = false;
    };

    auto uploadFilter = [&] {
        if (S.filter == defaultContactFilter) { S.check(pb_set_contact_filter(S.ctx, 0, nullptr, 0, nullptr), "pb_set_contact_filter"); S.filterDirty = false; return; }
        // tabulate the user's function over the (isTrigger, data) classes present in the scene
        std::vector<std::pair<bool, int>> classes;
        std::vector<int> cls(S.cols.size());
        std::unordered_map<long long, int> index;
        for (size_t i = 0; i < S.cols.size(); ++i) {
            const Collider& c = colStore.get(S.cols[i].e).colliders[S.cols[i].idx];
            long long key = ((long long)c.data << 1) | (c.isTrigger ? 1 : 0);
            auto it = index.find(key);
            if (it == index.end()) { it = index.emplace(key, (int)classes.size()).first; classes.emplace_back(c.isTrigger, c.data); }
            cls[i] = it->second;
        }
        const int K = (int)classes.size();
        if ((long long)K * K > (1ll << 26)) throw std::runtime_error("physecs_b200: custom contact filter over too many distinct collider data values");
        std::vector<unsigned char> lut((size_t)K * K);
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b)
                lut[(size_t)a * K + b] = S.filter(classes[a].first, classes[a].second, classes[b].first, classes[b].second) == TRIGGER ? 1 : 0;
        S.check(pb_set_contact_filter(S.ctx, (int)cls.size(), cls.data(), K, lut.data()), "pb_set_contact_filter");
        S.filterDirty = false;
    };

    if (S.topologyDirty) rebuild(false);
    if (S.jointsDirty) uploadJoints();
    else {
        bool dirty = false;
        for (Joint* J : S.joints) dirty |= J->paramsDirty;
        if (dirty) {
            std::vector<float> prm((size_t)8 * S.joints.size());
            for (size_t j = 0; j < S.joints.size(); ++j) { std::memcpy(&prm[8 * j], S.joints[j]->params, sizeof(float) * 8); S.joints[j]->paramsDirty = false; }
            S.check(pb_update_joint_params(S.ctx, (int)S.joints.size(), prm.data()), "pb_update_joint_params");
        }
    }
    if (S.pairsDirty) {
        std::vector<int> p; p.reserve(2 * S.nonColliding.size());
        for (auto& pr : S.nonColliding) { p.push_back((int)pr.first); p.push_back((int)pr.second); }
        S.check(pb_set_noncolliding_pairs(S.ctx, (int)S.nonColliding.size(), p.data()), "pb_set_noncolliding_pairs");
        S.pairsDirty = false;
    }
    if (S.filterDirty) uploadFilter();

    // ---- bodies announced through registry.patch<TransformComponent>: pose + bounds (+0.01) from the patch-time transform -----
    if (!S.moved.empty()) {
        // an entity patched several times between two steps: the LAST patch wins, as in the reference (every patch calls
        // updateBounds, Physecs.cpp:51-54).  One row per entity goes to the device (duplicate rows would race in the scatter).
        std::vector<int> rows; std::vector<float> p, q;
        std::unordered_map<int, size_t> slotOfRow;
        for (auto& mv : S.moved) {
            int r = S.rowOf(mv.e);
            if (r < 0) continue;
            auto it = slotOfRow.find(r);
            if (it != slotOfRow.end()) {
                const size_t k = it->second;
                p[3 * k] = mv.p.x; p[3 * k + 1] = mv.p.y; p[3 * k + 2] = mv.p.z;
                q[4 * k] = mv.q.x; q[4 * k + 1] = mv.q.y; q[4 * k + 2] = mv.q.z; q[4 * k + 3] = mv.q.w;
                continue;
            }
            slotOfRow[r] = rows.size();
            rows.push_back(r);
            p.insert(p.end(), { mv.p.x, mv.p.y, mv.p.z });
            q.insert(q.end(), { mv.q.x, mv.q.y, mv.q.z, mv.q.w });
        }
        if (!rows.empty()) S.check(pb_move_rows(S.ctx, (int)rows.size(), rows.data(), p.data(), q.data()), "pb_move_rows");
        S.moved.clear();
    }
}

void Scene::simulate(float timeStep) {
    Impl& S = *impl;
    auto tStart = Clock::now();
    prepareDevice();
    const double prepareMs = msSince(tStart);
    auto& dynStore = registry.storage<RigidBodyDynamicComponent>();
    auto& trStore = registry.storage<TransformComponent>();
    const int nDyn = S.nDyn, nStatic = S.nStatic;

    // ---- gather: registry -> pinned SoA ------------------------------------------------------------------------------------------
    auto tGather = Clock::now();
    S.hPos.reserve((size_t)3 * nDyn); S.hQuat.reserve((size_t)4 * nDyn); S.hVel.reserve((size_t)3 * nDyn); S.hAng.reserve((size_t)3 * nDyn);
    S.hSPos.reserve((size_t)3 * nStatic); S.hSQuat.reserve((size_t)4 * nStatic);
    std::atomic<int> kinChanged{0}, massChanged{0};
    const entt::entity* trEnts = trStore.data();
    const size_t trSize = trStore.size();
    auto transformOf = [&](int r) -> TransformComponent& {
        unsigned idx = S.rowTransformIdx[r];
        if (idx >= trSize || trEnts[idx] != S.rowEntity[r]) {      // storage was reordered behind our back: re-resolve
            idx = (unsigned)trStore.index(S.rowEntity[r]);
            S.rowTransformIdx[r] = idx;
        }
        return packedAt<TransformComponent>(trStore, idx);
    };
    const bool full = S.syncMode == SYNC_FULL;
    // what: 1 = poses (TransformComponent), 2 = velocities + the fields the reference reads live (RigidBodyDynamicComponent), 3 = both
    auto gatherDyn = [&](size_t b, size_t e, int what) {
        for (size_t r = b; r < e; ++r) {
            if (what & 1) {
                const TransformComponent& t = transformOf((int)r);
                float* p = S.hPos.p + 3 * r; float* q = S.hQuat.p + 4 * r;
                p[0] = t.position.x; p[1] = t.position.y; p[2] = t.position.z;
                q[0] = t.orientation.x; q[1] = t.orientation.y; q[2] = t.orientation.z; q[3] = t.orientation.w;
            }
            if (!(what & 2)) continue;
            const RigidBodyDynamicComponent& d = packedAt<RigidBodyDynamicComponent>(dynStore, r);
            float* v = S.hVel.p + 3 * r; float* w = S.hAng.p + 3 * r;
            v[0] = d.velocity.x; v[1] = d.velocity.y; v[2] = d.velocity.z;
            w[0] = d.angularVelocity.x; w[1] = d.angularVelocity.y; w[2] = d.angularVelocity.z;
            // fields the reference reads live every step (Physecs.cpp:219-226, :446-467): detect direct writes
            int k = d.isKinematic ? 1 : 0;
            if (k != S.kin[r]) { S.kin[r] = k; kinChanged.store(1, std::memory_order_relaxed); }
            if (d.invMass != S.invMass[r] || std::memcmp(&d.com, &S.com[3 * r], sizeof(float) * 3) || std::memcmp(&d.invInertiaTensor, &S.invI[9 * r], sizeof(float) * 9)) {
                S.invMass[r] = d.invMass;
                std::memcpy(&S.com[3 * r], &d.com, sizeof(float) * 3);
                std::memcpy(&S.invI[9 * r], &d.invInertiaTensor, sizeof(float) * 9);
                massChanged.store(1, std::memory_order_relaxed);
            }
        }
    };
    // The step's head goes to the device first: counters reset + broadphase, which run on the bounds the previous step left (as the
    // reference's sweep does, Physecs.cpp:119-173) and need nothing from the registry.  The gather below overlaps it.
    S.check(pb_step_begin(S.ctx), "pb_step_begin");
    // Dynamic rows in chunks: each chunk's upload (copy stream) starts as soon as the worker threads have gathered it.  Poses go
    // first, for every chunk; the narrowphase, which needs nothing else, is enqueued behind them (pb_step_narrowphase), and the
    // velocities are gathered and uploaded while it runs -- only the contact build waits for them.
    const int nChunks = nDyn >= 65536 ? 8 : 1;
    const size_t perChunk = ((size_t)nDyn + nChunks - 1) / nChunks;
    auto gatherAndUpload = [&]() {
        const bool split = nDyn >= 65536;
        for (int pass = 0; pass < (split ? 2 : 1); ++pass) {
            const int what = split ? (pass == 0 ? 1 : 2) : 3;
            for (int c = 0; c < nChunks; ++c) {
                const size_t first = (size_t)c * perChunk;
                if (first >= (size_t)nDyn) break;
                const size_t count = std::min(perChunk, (size_t)nDyn - first);
                S.workers.parallelFor(count, [&](size_t b, size_t e) { gatherDyn(first + b, first + e, what); });
                S.check(pb_set_state_rows(S.ctx, (int)first, (int)count, (what & 1) ? S.hPos.p + 3 * first : nullptr, (what & 1) ? S.hQuat.p + 4 * first : nullptr,
                                          (what & 2) ? S.hVel.p + 3 * first : nullptr, (what & 2) ? S.hAng.p + 3 * first : nullptr), "pb_set_state_rows");
            }
            if (split && pass == 0) S.check(pb_step_narrowphase(S.ctx), "pb_step_narrowphase");
        }
    };
    if (full) {
        // static poses first: pb_set_static_poses runs on the main stream and waits for any pending state upload, so issued after the
        // dynamic rows it would serialise their H2D (copy stream, hidden behind the broadphase) in front of the rest of the step
        S.workers.parallelFor((size_t)nStatic, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                const TransformComponent& t = transformOf(nDyn + (int)i);
                float* p = S.hSPos.p + 3 * i; float* q = S.hSQuat.p + 4 * i;
                p[0] = t.position.x; p[1] = t.position.y; p[2] = t.position.z;
                q[0] = t.orientation.x; q[1] = t.orientation.y; q[2] = t.orientation.z; q[3] = t.orientation.w;
            }
        });
        S.check(pb_set_static_poses(S.ctx, nStatic, S.hSPos.p, S.hSQuat.p), "pb_set_static_poses");
        gatherAndUpload();
    } else if (!S.touched.empty()) {
        // device-authoritative: the staging buffers still hold the previous step's result; refresh only announced rows
        if (!S.stagingValid) gatherAndUpload();
        else {
            for (unsigned r : S.touched) if ((int)r < nDyn) gatherDyn(r, r + 1, 3);
            S.check(pb_set_state(S.ctx, nDyn, S.hPos.p, S.hQuat.p, S.hVel.p, S.hAng.p), "pb_set_state");
        }
    }
    S.touched.clear();
    // direct writes to isKinematic / mass properties seen during the gather (the reference reads them live): these calls wait for the
    // uploads; a kinematic flip also changes which colliders query the broadphase, so the head of the step is redone (pb_step does it)
    if (kinChanged.load()) S.check(pb_set_kinematic(S.ctx, nDyn, S.kin.data()), "pb_set_kinematic");
    if (massChanged.load()) S.check(pb_set_mass(S.ctx, nDyn, S.invMass.data(), S.com.data(), S.invI.data()), "pb_set_mass");
    const double gatherMs = msSince(tGather);

    // ---- device step ---------------------------------------------------------------------------------------------------------------
    // pb_step only enqueues: an arena that overflows is noticed on the device (the step then skips its solve and leaves the scene as it
    // was) and reported by the next call that synchronises with the step -- the read-back here.  Nothing persistent was touched, so the
    // arenas are enlarged in place and the step is run again.  The device counters keep counting past the capacity, so they tell
    // how much room the step needs.
    auto stepAndFetch = [&]() {
        int r = pb_step(S.ctx, timeStep, numSubSteps, numIterations, g);
        if (r == PB_OK) r = pb_get_state_begin(S.ctx, S.hPos.p, S.hQuat.p, S.hVel.p, S.hAng.p, nChunks);
        return r;
    };
    int rc = stepAndFetch();
    for (int attempt = 0; rc == PB_ECAPACITY && attempt < 6; ++attempt) {
        pb_counts need{};
        pb_get_counts(S.ctx, &need);
        int wantP = std::max(S.caps.max_pairs, need.n_pairs + need.n_pairs / 2 + 1024);
        int wantM = std::max(S.caps.max_manifolds, need.n_manifolds + need.n_manifolds / 2 + 1024);
        if (need.n_pairs <= S.caps.max_pairs && need.n_manifolds <= S.caps.max_manifolds) break;   // not a pair / manifold arena (pb_last_error names the cause)
        S.check(pb_grow_arenas(S.ctx, wantP, wantM), "pb_grow_arenas");
        S.caps.max_pairs = S.wantPairs = wantP; S.caps.max_manifolds = S.wantManifolds = wantM;
        rc = stepAndFetch();
    }
    S.check(rc, "pb_step / pb_get_state_begin");
    S.stagingValid = true;

    // ---- scatter: pinned SoA -> registry, chunk by chunk as the read-back arrives (kinematic bodies are not integrated,
    // Physecs.cpp:446, :497)
    double scatterMs = 0.0;
    for (int c = 0; ; ++c) {
        int first = 0, count = 0;
        S.check(pb_get_state_wait(S.ctx, c, &first, &count), "pb_get_state_wait");
        if (count <= 0) break;
        auto tScatter = Clock::now();
        S.workers.parallelFor((size_t)count, [&](size_t b, size_t e) {
            for (size_t r = (size_t)first + b; r < (size_t)first + e; ++r) {
                if (S.kin[r]) continue;
                TransformComponent& t = transformOf((int)r);
                RigidBodyDynamicComponent& d = packedAt<RigidBodyDynamicComponent>(dynStore, r);
                const float* p = S.hPos.p + 3 * r; const float* q = S.hQuat.p + 4 * r; const float* v = S.hVel.p + 3 * r; const float* w = S.hAng.p + 3 * r;
                t.position = glm::vec3(p[0], p[1], p[2]);
                t.orientation = glm::quat(q[3], q[0], q[1], q[2]);
                d.velocity = glm::vec3(v[0], v[1], v[2]);
                d.angularVelocity = glm::vec3(w[0], w[1], w[2]);
            }
        });
        scatterMs += msSince(tScatter);
    }

    // ---- triggers: enter / exit diff against the previous step (Physecs.cpp:538-553) ---------------------------------------------
    pb_counts counts{};
    pb_get_counts(S.ctx, &counts);
    if (counts.n_triggers > 0 || !S.triggerCache.empty()) {
        std::vector<std::array<int, 4>> cur((size_t)std::max(counts.n_triggers, 0));
        int got = 0;
        if (counts.n_triggers > 0) S.check(pb_get_triggers(S.ctx, &cur[0][0], counts.n_triggers, &got), "pb_get_triggers");
        cur.resize((size_t)got);
        std::sort(cur.begin(), cur.end());
        std::vector<std::array<int, 4>> entered, exited;
        std::set_difference(cur.begin(), cur.end(), S.triggerCache.begin(), S.triggerCache.end(), std::back_inserter(entered));
        std::set_difference(S.triggerCache.begin(), S.triggerCache.end(), cur.begin(), cur.end(), std::back_inserter(exited));
        S.triggerCache.swap(cur);
        auto listenersIn = S.enterListeners;     // listeners may unregister themselves from inside a callback
        auto listenersOut = S.exitListeners;
        for (auto& t : entered) for (auto* l : listenersIn) l->onTriggerEnter((entt::entity)(unsigned)t[0], t[1], (entt::entity)(unsigned)t[2], t[3]);
        for (auto& t : exited) for (auto* l : listenersOut) l->onTriggerExit((entt::entity)(unsigned)t[0], t[1], (entt::entity)(unsigned)t[2], t[3]);
    }

    pb_timings tm{};
    pb_get_timings(S.ctx, &tm);
    S.stats = { counts.n_pairs, counts.n_manifolds, counts.n_points, counts.n_colors, counts.n_triggers, tm.total, gatherMs, scatterMs, msSince(tStart), prepareMs };
}

} // namespace physecs

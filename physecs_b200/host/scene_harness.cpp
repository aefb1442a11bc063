// Flat C entry points over physecs::Scene (the host C++ layer of this repo) so Python tests and bench.py can drive the
// same public API an application uses: fill an entt::registry with TransformComponent / RigidBodyCollisionComponent /
// RigidBodyDynamicComponent in the reference's emplacement order (demo/Demo.cpp:32-41), create joints through
// Scene::createJoint<T>, call Scene::simulate and read the registry back.  Shaped like oracle/ref_harness.cpp on purpose:
// the parity tests run one scene description through both and compare.
#include <Physecs.h>
#include <MassUtil.h>
#include <FixedJoint.h>
#include <RevoluteJoint.h>
#include <SphericalJoint.h>
#include <UniversalJoint.h>
#include <PrismaticJoint.h>
#include <GearJoint.h>
#include <ServoJoint.h>
#include "../../include/physecs_b200.h"

#include <array>
#include <chrono>
#include <memory>
#include <string>

namespace {

struct Recorder : physecs::OnTriggerEnterListener, physecs::OnTriggerExitListener {
    std::vector<std::array<int, 5>> events;
    void onTriggerEnter(entt::entity e0, int c0, entt::entity e1, int c1) override { events.push_back({ 0, (int)e0, c0, (int)e1, c1 }); }
    void onTriggerExit(entt::entity e0, int c0, entt::entity e1, int c1) override { events.push_back({ 1, (int)e0, c0, (int)e1, c1 }); }
};

physecs::ContactType filterParity(bool t0, int d0, bool t1, int d1) {
    return ((t0 || t1) && ((d0 + d1) % 2 == 0)) ? physecs::TRIGGER : physecs::COLLISION;
}
physecs::ContactType filterAsymmetric(bool t0, int d0, bool t1, int) {
    if (t0 && !t1) return physecs::TRIGGER;
    if (t1 && d0 == 1) return physecs::TRIGGER;
    return physecs::COLLISION;
}

struct Harness {
    Recorder recorder;
    entt::registry registry;
    std::unique_ptr<physecs::Scene> scene;
    std::vector<entt::entity> entities;
    std::vector<std::unique_ptr<physecs::ConvexMesh>> convex;
    std::vector<std::unique_ptr<physecs::TriangleMesh>> trimesh;
    std::vector<physecs::Joint*> joints;
    std::string error;
    ~Harness() { scene.reset(); }
};

glm::vec3 v3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
glm::quat q4(const float* p) { return glm::quat(p[3], p[0], p[1], p[2]); }

physecs::Geometry makeGeometry(Harness* h, int type, const float* p, int mesh) {
    physecs::Geometry g{};
    g.type = (physecs::GeometryType)type;
    switch (type) {
        case physecs::SPHERE: g.sphere = { p[0] }; break;
        case physecs::CAPSULE: g.capsule = { p[0], p[1] }; break;
        case physecs::BOX: g.box = { glm::vec3(p[0], p[1], p[2]) }; break;
        case physecs::CONVEX_MESH: g.convex = { h->convex[mesh].get(), glm::vec3(p[0], p[1], p[2]) }; break;
        case physecs::TRIANGLE_MESH: g.triangleMesh = { h->trimesh[mesh].get() }; break;
    }
    return g;
}

physecs::Collider makeCollider(Harness* h, const float* lpos, const float* lquat, int type, const float* params, int mesh, const float* material,
                               int flags, int data) {
    physecs::Collider col{};
    col.position = v3(lpos);
    col.orientation = q4(lquat);
    col.geometry = makeGeometry(h, type, params, mesh);
    col.material = { material[0], material[1], material[2] };
    col.isTrigger = flags & 1;
    col.enableSimulation = (flags >> 1) & 1;
    col.data = data;
    return col;
}

} // namespace

#define GUARD(h, body) try { body; return 0; } catch (const std::exception& e) { (h)->error = e.what(); return -1; }

extern "C" {

void* psh_create(int numThreads, int device) {
    auto* h = new Harness();
    h->scene = std::make_unique<physecs::Scene>(h->registry, numThreads);
    h->scene->setDevice(device);
    return h;
}
void psh_destroy(void* hp) { delete (Harness*)hp; }
const char* psh_last_error(void* hp) { return ((Harness*)hp)->error.c_str(); }

int psh_add_convex(void* hp, const float* verts, int nv, const int* faceOffsets, const int* faceIndices, int nf, const float* normals, const float* centroids) {
    auto* h = (Harness*)hp;
    std::vector<glm::vec3> v(nv);
    for (int i = 0; i < nv; ++i) v[i] = v3(verts + 3 * i);
    std::vector<physecs::ConvexMeshFace> faces(nf);
    for (int f = 0; f < nf; ++f) {
        faces[f].indices.assign(faceIndices + faceOffsets[f], faceIndices + faceOffsets[f + 1]);
        faces[f].normal = v3(normals + 3 * f);
        faces[f].centroid = v3(centroids + 3 * f);
    }
    h->convex.push_back(std::make_unique<physecs::ConvexMesh>(std::move(v), std::move(faces)));
    return (int)h->convex.size() - 1;
}

int psh_add_trimesh(void* hp, const float* verts, int nv, const unsigned* indices, int ni) {
    auto* h = (Harness*)hp;
    std::vector<glm::vec3> v(nv);
    for (int i = 0; i < nv; ++i) v[i] = v3(verts + 3 * i);
    h->trimesh.push_back(std::make_unique<physecs::TriangleMesh>(v, std::vector<unsigned>(indices, indices + ni)));
    return (int)h->trimesh.size() - 1;
}

// flags bit0: RigidBodyCollisionComponent, bit1: RigidBodyDynamicComponent, bit2: isKinematic; colFlags bit0 trigger, bit1 enableSimulation
int psh_add_entities(void* hp, int n, const float* pos, const float* quat, const int* flags, const float* vel, const float* angvel,
                     const float* invMass, const float* com, const float* invI, const int* colOffsets, const float* colLPos,
                     const float* colLQuat, const int* colType, const float* colParams, const int* colMesh, const float* colMaterial,
                     const int* colFlags, const int* colData) {
    auto* h = (Harness*)hp;
    int first = (int)h->entities.size();
    for (int i = 0; i < n; ++i) {
        auto e = h->registry.create();
        h->entities.push_back(e);
        h->registry.emplace<TransformComponent>(e, v3(pos + 3 * i), q4(quat + 4 * i), glm::vec3(1));
        if (flags[i] & 1) {
            std::vector<physecs::Collider> cols;
            for (int c = colOffsets[i]; c < colOffsets[i + 1]; ++c)
                cols.push_back(makeCollider(h, colLPos + 3 * c, colLQuat + 4 * c, colType[c], colParams + 4 * c, colMesh[c], colMaterial + 3 * c, colFlags[c], colData[c]));
            h->registry.emplace<physecs::RigidBodyCollisionComponent>(e, std::move(cols));
        }
        if (flags[i] & 2) {
            glm::mat3 I;
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) I[c][r] = invI[9 * i + 3 * c + r];
            h->registry.emplace<physecs::RigidBodyDynamicComponent>(e, (flags[i] & 4) != 0, v3(vel + 3 * i), v3(angvel + 3 * i), invMass[i], v3(com + 3 * i), I);
        }
    }
    return first;
}

int psh_destroy_entity(void* hp, int e) {
    auto* h = (Harness*)hp;
    h->registry.destroy(h->entities[e]);
    return 0;
}

// registry.sort on the dynamic-body pool: the Scene has to notice that its rows no longer follow the pool
void psh_sort_dynamic(void* hp, int greaterFirst) {     // EnTT iterates a pool back to front: a > b keeps a creation-ordered pool's packed order, a < b reverses it
    auto* h = (Harness*)hp;
    if (greaterFirst) h->registry.sort<physecs::RigidBodyDynamicComponent>([](const entt::entity a, const entt::entity b) { return a > b; });
    else h->registry.sort<physecs::RigidBodyDynamicComponent>([](const entt::entity a, const entt::entity b) { return a < b; });
}

int psh_add_collider(void* hp, int e, const float* lpos, const float* lquat, int type, const float* params, int mesh, const float* material, int flags, int data) {
    auto* h = (Harness*)hp;
    h->scene->addCollider(h->entities[e], makeCollider(h, lpos, lquat, type, params, mesh, material, flags, data));
    return 0;
}
int psh_clear_colliders(void* hp, int e) { auto* h = (Harness*)hp; h->scene->clearColliders(h->entities[e]); return 0; }

int psh_add_joint(void* hp, int type, int e0, const float* a0p, const float* a0q, int e1, const float* a1p, const float* a1q, const float* prm) {
    auto* h = (Harness*)hp;
    auto E0 = h->entities[e0], E1 = h->entities[e1];
    physecs::Joint* j = nullptr;
    switch (type) {
        case 0: j = h->scene->createJoint<physecs::FixedJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q)); break;
        case 1: {
            auto* r = h->scene->createJoint<physecs::RevoluteJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q));
            if (prm) { r->setDriveEnabled(prm[0] != 0); r->setDriveVelocity(prm[1]); r->setDriveMaxTorque(prm[2]); }
            j = r;
        } break;
        case 2: j = h->scene->createJoint<physecs::SphericalJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q)); break;
        case 3: j = h->scene->createJoint<physecs::UniversalJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q)); break;
        case 4: {
            auto* p = h->scene->createJoint<physecs::PrismaticJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q));
            if (prm) {
                p->setUpperLimit(prm[0]); p->setLowerLimit(prm[1]); p->setDriveEnabled(prm[2] != 0);
                p->setTargetPosition(prm[3]); p->setDriveStiffness(prm[4]); p->setDriveDamping(prm[5]);
            }
            j = p;
        } break;
        case 5: {
            auto* g = h->scene->createJoint<physecs::GearJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q));
            if (prm) g->setGearRatio(prm[0]);
            j = g;
        } break;
        case 6: {
            auto* s = h->scene->createJoint<physecs::ServoJoint>(E0, v3(a0p), q4(a0q), E1, v3(a1p), q4(a1q));
            if (prm) { s->setTargetAngle(prm[0]); s->setDriveStiffness(prm[1]); s->setDriveDamping(prm[2]); }
            j = s;
        } break;
    }
    h->joints.push_back(j);
    return j ? j->getColor() : -1;
}
// RevoluteJoint::setDriveVelocity on a live joint (exercises the joint-parameter update path)
int psh_set_revolute_drive(void* hp, int joint, int enabled, float velocity, float maxTorque) {
    auto* h = (Harness*)hp;
    auto* r = dynamic_cast<physecs::RevoluteJoint*>(h->joints[joint]);
    if (!r) return -1;
    r->setDriveEnabled(enabled != 0); r->setDriveVelocity(velocity); r->setDriveMaxTorque(maxTorque);
    return 0;
}
int psh_destroy_joint(void* hp, int joint) {
    auto* h = (Harness*)hp;
    h->scene->destroyJoint(h->joints[joint]);
    h->joints[joint] = nullptr;
    return 0;
}

void psh_set_params(void* hp, int substeps, int iterations, float gravity) {
    auto* h = (Harness*)hp;
    h->scene->setNumSubSteps(substeps);
    h->scene->setNumIterations(iterations);
    h->scene->setGravity(gravity);
}
void psh_set_can_collide(void* hp, int e0, int e1, int can) { auto* h = (Harness*)hp; h->scene->setCanCollide(h->entities[e0], h->entities[e1], can != 0); }
void psh_set_kinematic(void* hp, int e, int kin) { auto* h = (Harness*)hp; h->scene->setIsKinematic(h->entities[e], kin != 0); }
void psh_set_sync_mode(void* hp, int deviceAuthoritative) {
    ((Harness*)hp)->scene->setSyncMode(deviceAuthoritative ? physecs::Scene::SYNC_DEVICE_AUTHORITATIVE : physecs::Scene::SYNC_FULL);
}
void psh_set_arena_capacity(void* hp, int maxPairs, int maxManifolds) { ((Harness*)hp)->scene->setArenaCapacity(maxPairs, maxManifolds); }
void psh_set_contact_filter(void* hp, int mode) {
    ((Harness*)hp)->scene->setContactFilter(mode == 1 ? filterParity : mode == 2 ? filterAsymmetric : physecs::defaultContactFilter);
}
void psh_record_trigger_events(void* hp) {
    auto* h = (Harness*)hp;
    h->scene->addOnTriggerEnterCallback(&h->recorder);
    h->scene->addOnTriggerExitCallback(&h->recorder);
}
int psh_take_trigger_events(void* hp, int* out, int cap) {
    auto* h = (Harness*)hp;
    int n = (int)h->recorder.events.size();
    for (int i = 0; i < n && i < cap; ++i) for (int k = 0; k < 5; ++k) out[5 * i + k] = h->recorder.events[i][k];
    h->recorder.events.clear();
    return n;
}

// overwrite registry state; patch != 0 goes through registry.patch<TransformComponent> (Scene::onRigidBodyMove)
void psh_set_state(void* hp, int n, const int* ents, const float* pos, const float* quat, const float* vel, const float* angvel, int patch) {
    auto* h = (Harness*)hp;
    for (int i = 0; i < n; ++i) {
        auto e = h->entities[ents[i]];
        auto apply = [&](TransformComponent& t) { t.position = v3(pos + 3 * i); t.orientation = q4(quat + 4 * i); };
        if (patch) h->registry.patch<TransformComponent>(e, apply);
        else apply(h->registry.get<TransformComponent>(e));
        if (auto* d = h->registry.try_get<physecs::RigidBodyDynamicComponent>(e)) {
            if (vel) d->velocity = v3(vel + 3 * i);
            if (angvel) d->angularVelocity = v3(angvel + 3 * i);
        }
    }
}

// one Scene::simulate; returns wall milliseconds, < 0 on error (psh_last_error)
double psh_simulate(void* hp, float dt) {
    auto* h = (Harness*)hp;
    auto t0 = std::chrono::high_resolution_clock::now();
    try { h->scene->simulate(dt); } catch (const std::exception& e) { h->error = e.what(); return -1.0; }
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
}

int psh_num_entities(void* hp) { return (int)((Harness*)hp)->entities.size(); }

void psh_get_state(void* hp, float* pos, float* quat, float* vel, float* angvel) {
    auto* h = (Harness*)hp;
    for (size_t i = 0; i < h->entities.size(); ++i) {
        auto e = h->entities[i];
        if (!h->registry.valid(e)) continue;
        auto& t = h->registry.get<TransformComponent>(e);
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = t.position[k];
        quat[4 * i + 0] = t.orientation.x; quat[4 * i + 1] = t.orientation.y; quat[4 * i + 2] = t.orientation.z; quat[4 * i + 3] = t.orientation.w;
        auto* d = h->registry.try_get<physecs::RigidBodyDynamicComponent>(e);
        for (int k = 0; k < 3; ++k) { vel[3 * i + k] = d ? d->velocity[k] : 0.f; angvel[3 * i + k] = d ? d->angularVelocity[k] : 0.f; }
    }
}

// Scene::raycastClosest with a filter that accepts entities whose integer id % mod != skip (mod == 0: accept all);
// returns the entity (or -1) and writes the hit position
int psh_raycast(void* hp, const float* orig, const float* dir, float maxDist, int mod, int skip, float* hitPos3) {
    auto* h = (Harness*)hp;
    glm::vec3 hit(0);
    try {
        auto e = h->scene->raycastClosest(v3(orig), v3(dir), maxDist, [=](entt::entity x) { return mod == 0 || (int)((unsigned)x % (unsigned)mod) != skip; }, &hit);
        for (int k = 0; k < 3; ++k) hitPos3[k] = hit[k];
        return e == entt::null ? -1 : (int)e;
    } catch (const std::exception& ex) { h->error = ex.what(); return -2; }
}
// Scene::overlap with a query shape; rows (entity, colIndex); returns the count (or < 0 on error)
int psh_overlap(void* hp, const float* pos, const float* quat, int type, const float* params, int mesh, int filter, int cap, int* out2) {
    auto* h = (Harness*)hp;
    try {
        auto hits = h->scene->overlap(v3(pos), q4(quat), makeGeometry(h, type, params, mesh), filter);
        for (size_t i = 0; i < hits.size() && (int)i < cap; ++i) { out2[2 * i] = (int)hits[i].entity; out2[2 * i + 1] = hits[i].colIndex; }
        return (int)hits.size();
    } catch (const std::exception& ex) { h->error = ex.what(); return -1; }
}

// Scene::overlapWithMinTranslationalDistance; rows (entity, colIndex), (normal xyz, mtd); returns the count (or < 0 on error)
int psh_overlap_mtd(void* hp, const float* pos, const float* quat, int type, const float* params, int mesh, int cap, int* out2, float* out4) {
    auto* h = (Harness*)hp;
    try {
        auto hits = h->scene->overlapWithMinTranslationalDistance(v3(pos), q4(quat), makeGeometry(h, type, params, mesh));
        for (size_t i = 0; i < hits.size() && (int)i < cap; ++i) {
            out2[2 * i] = (int)hits[i].entity; out2[2 * i + 1] = hits[i].colIndex;
            for (int k = 0; k < 3; ++k) out4[4 * i + k] = hits[i].normal[k];
            out4[4 * i + 3] = hits[i].mtd;
        }
        return (int)hits.size();
    } catch (const std::exception& ex) { h->error = ex.what(); return -1; }
}

// Scene::getBVH / getBHVRootId: checks the snapshot's structure and returns the number of leaves (< 0: malformed); every leaf's
// (entity, colIndex, bounds6) is written while room lasts
int psh_bvh_leaves(void* hp, int cap, int* ids2, float* bounds6) {
    auto* h = (Harness*)hp;
    try {
        const auto& nodes = h->scene->getBVH();
        int root = h->scene->getBHVRootId();
        if (nodes.empty()) return root == physecs::BVH::null ? 0 : -2;
        if (root < 0 || root >= (int)nodes.size() || nodes[root].parent != physecs::BVH::null) return -3;
        int leaves = 0;
        std::vector<int> stack{ root };
        size_t visited = 0;
        while (!stack.empty()) {
            int id = stack.back(); stack.pop_back();
            ++visited;
            const auto& n = nodes[id];
            if (n.isLeaf) {
                if (leaves < cap) {
                    ids2[2 * leaves] = (int)n.leaf.entity; ids2[2 * leaves + 1] = n.leaf.colliderIndex;
                    for (int k = 0; k < 3; ++k) { bounds6[6 * leaves + k] = n.bounds.min[k]; bounds6[6 * leaves + 3 + k] = n.bounds.max[k]; }
                }
                ++leaves;
                continue;
            }
            for (int child : { n.internal.left, n.internal.right }) {
                if (child < 0 || child >= (int)nodes.size() || nodes[child].parent != id) return -4;
                const auto& c = nodes[child];
                for (int k = 0; k < 3; ++k) if (c.bounds.min[k] < n.bounds.min[k] || c.bounds.max[k] > n.bounds.max[k]) return -5;   // a parent box holds its children
                stack.push_back(child);
            }
        }
        if (visited != nodes.size()) return -6;
        return leaves;
    } catch (const std::exception& ex) { h->error = ex.what(); return -1; }
}

// the Scene's C-ABI context (pb_ctx*) for the parity taps, and the statistics of the last step
void* psh_native_context(void* hp) { return ((Harness*)hp)->scene->nativeContext(); }
void psh_get_stats(void* hp, double* out10) {
    auto s = ((Harness*)hp)->scene->getLastStepStats();
    out10[0] = s.pairs; out10[1] = s.manifolds; out10[2] = s.points; out10[3] = s.colors; out10[4] = s.triggers;
    out10[5] = s.deviceMs; out10[6] = s.gatherMs; out10[7] = s.scatterMs; out10[8] = s.totalMs; out10[9] = s.prepareMs;
}

// physecs::computeCOMAndInvInertiaTensor on entity e's colliders (MassUtil parity check)
void psh_mass_props(void* hp, int e, float mass, float* com3, float* invI9) {
    auto* h = (Harness*)hp;
    glm::vec3 com; glm::mat3 inv;
    physecs::computeCOMAndInvInertiaTensor(h->registry.get<physecs::RigidBodyCollisionComponent>(h->entities[e]), mass, com, inv);
    for (int k = 0; k < 3; ++k) com3[k] = com[k];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) invI9[3 * c + r] = inv[c][r];
}

} // extern "C"

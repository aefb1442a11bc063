// Oracle harness: a flat C ABI over the reference's own C++ API (physecs::Scene on an
// entt::registry), so Python tests / bench.py can drive the UNMODIFIED reference
// implementation (compiled by oracle/build_ref.py) with numpy arrays.
//
// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the
// product path (physecs_b200/).  It is the checker for tests/, smoke() and the
// cpu_baseline / --impl reference legs of bench.py.
//
// Reference interfaces exercised (paths relative to /root/reference):
//   physecs::Scene ctor/simulate/setters      include/Physecs/Physecs.h:199-228, src/Physecs.cpp:92-561
//   component emplacement order (Transform -> Collision -> Dynamic)  demo/Demo.cpp:32-41
//   physecs::collision (isolated narrowphase)  src/Collision.h:8, src/Collision.cpp:895
//   Scene::createJoint<T>                      include/Physecs/Physecs.h:209-214
//   TriangleMesh ctor (binned-SAH BVH)         src/TriangleMesh.cpp:144-164

#include <Physecs.h>
#include <Components.h>
#include <Transform.h>
#include <Collision.h>
#include <ContactConstraints.h>
#include <ContactManifold.h>
#include <SolverData.h>
#include <BoundsUtil.h>
#include <FixedJoint.h>
#include <RevoluteJoint.h>
#include <SphericalJoint.h>
#include <UniversalJoint.h>
#include <PrismaticJoint.h>
#include <GearJoint.h>
#include <ServoJoint.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <tuple>

extern std::vector<std::array<int, 2>> physecs_oracle_manifold_keys;
extern void (*physecs_oracle_presolve_hook)(physecs::Scene*);

namespace {

struct Key5 {
    int e0, c0, e1, c1, tri;
    bool operator<(const Key5& o) const {
        return std::tie(e0, c0, e1, c1, tri) < std::tie(o.e0, o.c0, o.e1, o.c1, o.tri);
    }
};

struct TriggerRecorder : physecs::OnTriggerEnterListener, physecs::OnTriggerExitListener {
    std::vector<std::array<int, 5>> events;   // (0 = enter | 1 = exit, e0, c0, e1, c1)
    void onTriggerEnter(entt::entity e0, int c0, entt::entity e1, int c1) override { events.push_back({ 0, (int)e0, c0, (int)e1, c1 }); }
    void onTriggerExit(entt::entity e0, int c0, entt::entity e1, int c1) override { events.push_back({ 1, (int)e0, c0, (int)e1, c1 }); }
};

// custom contact filters for the setContactFilter tests (Physecs.h:224); the same predicates are tabulated in
// tests/parity.py for the device path
physecs::ContactType filterParity(bool t0, int d0, bool t1, int d1) {
    if ((t0 || t1) && ((d0 + d1) % 2 == 0)) return physecs::TRIGGER;
    return physecs::COLLISION;
}
physecs::ContactType filterAsymmetric(bool t0, int d0, bool t1, int d1) {
    if (t0 && !t1) return physecs::TRIGGER;
    if (t1 && d0 == 1) return physecs::TRIGGER;
    return physecs::COLLISION;
}

struct Harness {
    TriggerRecorder recorder;
    entt::registry registry;
    std::unique_ptr<physecs::Scene> scene;
    std::vector<entt::entity> entities;
    std::vector<std::unique_ptr<physecs::ConvexMesh>> convex;
    std::vector<std::unique_ptr<physecs::TriangleMesh>> trimesh;
    std::vector<physecs::Joint*> joints;
    // gate-3 order injection
    std::vector<Key5> wantedOrder;
    bool useOrder = false;
    int orderMatched = 0, orderMissing = 0, orderExtra = 0;
    // manifold keys of the last simulate (in solve order)
    std::vector<Key5> lastKeys;
    ~Harness() { scene.reset(); }
};

Harness* g_hooked = nullptr;

void presolveHook(physecs::Scene* s) {
    Harness* h = g_hooked;
    if (!h || h->scene.get() != s) return;
    auto& cc = s->contactConstraints;
    const size_t n = cc.size();
    std::vector<Key5> keys(n);
    for (size_t i = 0; i < n; ++i) {
        auto& pr = s->potentialContacts[physecs_oracle_manifold_keys[i][0]];
        keys[i] = { (int)pr.entity0, pr.colliderIndex0, (int)pr.entity1, pr.colliderIndex1, physecs_oracle_manifold_keys[i][1] };
    }
    if (h->useOrder) {
        std::map<Key5, int> where;
        for (size_t i = 0; i < n; ++i) where[keys[i]] = (int)i;
        std::vector<int> perm;
        perm.reserve(n);
        std::vector<char> used(n, 0);
        h->orderMatched = h->orderMissing = h->orderExtra = 0;
        for (auto& k : h->wantedOrder) {
            auto it = where.find(k);
            if (it == where.end() || used[it->second]) { ++h->orderMissing; continue; }
            used[it->second] = 1;
            perm.push_back(it->second);
            ++h->orderMatched;
        }
        for (size_t i = 0; i < n; ++i) if (!used[i]) { perm.push_back((int)i); ++h->orderExtra; }
        std::vector<physecs::ContactConstraints> tmp;
        tmp.reserve(n);
        std::vector<Key5> k2;
        k2.reserve(n);
        for (int i : perm) { tmp.push_back(cc[i]); k2.push_back(keys[i]); }
        cc.swap(tmp);
        keys.swap(k2);
    }
    h->lastKeys.swap(keys);
}

glm::vec3 v3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
glm::quat q4(const float* p) { return glm::quat(p[3], p[0], p[1], p[2]); }  // input xyzw -> glm ctor (w,x,y,z)

} // namespace

extern "C" {

void* ph_create(int numThreads) {
    auto* h = new Harness();
    h->scene = std::make_unique<physecs::Scene>(h->registry, numThreads);
    physecs_oracle_presolve_hook = presolveHook;
    return h;
}

void ph_destroy(void* hp) {
    auto* h = (Harness*)hp;
    if (g_hooked == h) g_hooked = nullptr;
    delete h;
}

int ph_add_convex(void* hp, const float* verts, int nv, const int* faceOffsets, const int* faceIndices, int nf,
                  const float* normals, const float* centroids) {
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

int ph_add_trimesh(void* hp, const float* verts, int nv, const unsigned* indices, int ni) {
    auto* h = (Harness*)hp;
    std::vector<glm::vec3> v(nv);
    for (int i = 0; i < nv; ++i) v[i] = v3(verts + 3 * i);
    std::vector<unsigned> idx(indices, indices + ni);
    h->trimesh.push_back(std::make_unique<physecs::TriangleMesh>(v, idx));
    return (int)h->trimesh.size() - 1;
}

void ph_trimesh_sizes(void* hp, int id, int* ntri, int* nnodes) {
    auto* h = (Harness*)hp;
    *ntri = (int)h->trimesh[id]->triangles.size();
    *nnodes = (int)h->trimesh[id]->bvh.size();
}

// triIdx[ntri*3] post-build order; triNormal[ntri*3]; nodeBounds[nnodes*6]; nodeCI[nnodes*2] = (triCount, index)
void ph_trimesh_get(void* hp, int id, unsigned* triIdx, float* triNormal, float* nodeBounds, int* nodeCI) {
    auto* h = (Harness*)hp;
    auto& m = *h->trimesh[id];
    for (size_t i = 0; i < m.triangles.size(); ++i) {
        for (int k = 0; k < 3; ++k) {
            triIdx[3 * i + k] = m.triangles[i].indices[k];
            triNormal[3 * i + k] = m.triangles[i].normal[k];
        }
    }
    for (size_t i = 0; i < m.bvh.size(); ++i) {
        for (int k = 0; k < 3; ++k) {
            nodeBounds[6 * i + k] = m.bvh[i].bounds.min[k];
            nodeBounds[6 * i + 3 + k] = m.bvh[i].bounds.max[k];
        }
        nodeCI[2 * i] = m.bvh[i].triCount;
        nodeCI[2 * i + 1] = m.bvh[i].index;
    }
}

static physecs::Geometry makeGeometry(Harness* h, int type, const float* p, int mesh) {
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

// flags bit0: has RigidBodyCollisionComponent, bit1: has RigidBodyDynamicComponent, bit2: isKinematic
// colFlags bit0: isTrigger, bit1: enableSimulation
int ph_add_entities(void* hp, int n, const float* pos, const float* quat, const int* flags, const float* vel,
                    const float* angvel, const float* invMass, const float* com, const float* invI,
                    const int* colOffsets, const float* colLPos, const float* colLQuat, const int* colType,
                    const float* colParams, const int* colMesh, const float* colMaterial, const int* colFlags,
                    const int* colData) {
    auto* h = (Harness*)hp;
    int first = (int)h->entities.size();
    for (int i = 0; i < n; ++i) {
        auto e = h->registry.create();
        h->entities.push_back(e);
        h->registry.emplace<TransformComponent>(e, v3(pos + 3 * i), q4(quat + 4 * i), glm::vec3(1));
        if (flags[i] & 1) {
            std::vector<physecs::Collider> cols;
            for (int c = colOffsets[i]; c < colOffsets[i + 1]; ++c) {
                physecs::Collider col{};
                col.position = v3(colLPos + 3 * c);
                col.orientation = q4(colLQuat + 4 * c);
                col.geometry = makeGeometry(h, colType[c], colParams + 4 * c, colMesh[c]);
                col.material = { colMaterial[3 * c], colMaterial[3 * c + 1], colMaterial[3 * c + 2] };
                col.isTrigger = colFlags[c] & 1;
                col.enableSimulation = (colFlags[c] >> 1) & 1;
                col.data = colData[c];
                cols.push_back(col);
            }
            h->registry.emplace<physecs::RigidBodyCollisionComponent>(e, std::move(cols));
        }
        if (flags[i] & 2) {
            glm::mat3 I;
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) I[c][r] = invI[9 * i + 3 * c + r];
            h->registry.emplace<physecs::RigidBodyDynamicComponent>(e, (flags[i] & 4) != 0, v3(vel + 3 * i), v3(angvel + 3 * i),
                                                                     invMass[i], v3(com + 3 * i), I);
        }
    }
    return first;
}

// type: 0 fixed, 1 revolute, 2 spherical, 3 universal, 4 prismatic, 5 gear, 6 servo
int ph_add_joint(void* hp, int type, int e0, const float* a0p, const float* a0q, int e1, const float* a1p,
                 const float* a1q, const float* prm) {
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

void ph_set_params(void* hp, int substeps, int iterations, float gravity) {
    auto* h = (Harness*)hp;
    h->scene->setNumSubSteps(substeps);
    h->scene->setNumIterations(iterations);
    h->scene->setGravity(gravity);
}

void ph_set_can_collide(void* hp, int e0, int e1, int can) {
    auto* h = (Harness*)hp;
    h->scene->setCanCollide(h->entities[e0], h->entities[e1], can != 0);
}

void ph_set_kinematic(void* hp, int e, int kin) {
    auto* h = (Harness*)hp;
    h->scene->setIsKinematic(h->entities[e], kin != 0);
}

// Overwrite transforms (+ velocities of dynamic entities).  patch!=0 announces the move through
// registry.patch<TransformComponent> (-> Scene::onRigidBodyMove -> updateBounds, Physecs.cpp:51-54).
void ph_set_state(void* hp, int n, const int* ents, const float* pos, const float* quat, const float* vel,
                  const float* angvel, int patch) {
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

double ph_simulate(void* hp, float dt) {
    auto* h = (Harness*)hp;
    g_hooked = h;
    auto t0 = std::chrono::high_resolution_clock::now();
    h->scene->simulate(dt);
    auto t1 = std::chrono::high_resolution_clock::now();
    h->useOrder = false;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

int ph_num_entities(void* hp) { return (int)((Harness*)hp)->entities.size(); }

void ph_get_state(void* hp, float* pos, float* quat, float* vel, float* angvel) {
    auto* h = (Harness*)hp;
    for (size_t i = 0; i < h->entities.size(); ++i) {
        auto e = h->entities[i];
        if (!h->registry.valid(e)) continue;
        auto& t = h->registry.get<TransformComponent>(e);
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = t.position[k];
        quat[4 * i + 0] = t.orientation.x; quat[4 * i + 1] = t.orientation.y;
        quat[4 * i + 2] = t.orientation.z; quat[4 * i + 3] = t.orientation.w;
        auto* d = h->registry.try_get<physecs::RigidBodyDynamicComponent>(e);
        for (int k = 0; k < 3; ++k) {
            vel[3 * i + k] = d ? d->velocity[k] : 0.f;
            angvel[3 * i + k] = d ? d->angularVelocity[k] : 0.f;
        }
    }
}

// potentialContacts of the last simulate (Physecs.cpp:134-172): rows of (e0, c0, e1, c1)
int ph_num_pairs(void* hp) { return (int)((Harness*)hp)->scene->potentialContacts.size(); }
void ph_get_pairs(void* hp, int* out) {
    auto* h = (Harness*)hp;
    auto& pc = h->scene->potentialContacts;
    for (size_t i = 0; i < pc.size(); ++i) {
        out[4 * i + 0] = (int)pc[i].entity0; out[4 * i + 1] = pc[i].colliderIndex0;
        out[4 * i + 2] = (int)pc[i].entity1; out[4 * i + 3] = pc[i].colliderIndex1;
    }
}

// broadphase entry bounds as the Scene currently holds them: rows (entity, colIdx), bounds[6]
int ph_num_bounds(void* hp) { return (int)((Harness*)hp)->scene->broadPhaseEntries.size(); }
void ph_get_bounds(void* hp, int* ids, float* bounds) {
    auto* h = (Harness*)hp;
    auto& be = h->scene->broadPhaseEntries;
    for (size_t i = 0; i < be.size(); ++i) {
        ids[2 * i] = (int)be[i].entity; ids[2 * i + 1] = be[i].colliderIndex;
        for (int k = 0; k < 3; ++k) { bounds[6 * i + k] = be[i].bounds.min[k]; bounds[6 * i + 3 + k] = be[i].bounds.max[k]; }
    }
}

// manifold keys of the last simulate in the order the solver visited them: rows (e0,c0,e1,c1,tri)
int ph_num_manifolds(void* hp) { return (int)((Harness*)hp)->lastKeys.size(); }
void ph_get_manifold_keys(void* hp, int* out) {
    auto* h = (Harness*)hp;
    for (size_t i = 0; i < h->lastKeys.size(); ++i) {
        auto& k = h->lastKeys[i];
        out[5 * i] = k.e0; out[5 * i + 1] = k.c0; out[5 * i + 2] = k.e1; out[5 * i + 3] = k.c1; out[5 * i + 4] = k.tri;
    }
}

// Impose a contact-constraint order on the NEXT simulate (gate 3).  stats = (matched, missing, extra).
void ph_set_manifold_order(void* hp, const int* keys, int n) {
    auto* h = (Harness*)hp;
    h->wantedOrder.resize(n);
    for (int i = 0; i < n; ++i) h->wantedOrder[i] = { keys[5 * i], keys[5 * i + 1], keys[5 * i + 2], keys[5 * i + 3], keys[5 * i + 4] };
    h->useOrder = true;
}
void ph_get_order_stats(void* hp, int* out) {
    auto* h = (Harness*)hp;
    out[0] = h->orderMatched; out[1] = h->orderMissing; out[2] = h->orderExtra;
}

// Isolated narrowphase on the CURRENT registry state for given pairs (e0,c0,e1,c1), world collider poses
// formed exactly as Physecs.cpp:194-198.  Output rows per manifold with numPoints>0:
//   keys[5] = (pair index, tri, numPoints, 0, 0); normal[3]; points[4][2][3]
int ph_narrowphase(void* hp, const int* pairs, int npairs, int cap, int* keys, float* normal, float* points) {
    auto* h = (Harness*)hp;
    int m = 0;
    std::vector<physecs::ContactManifold> buf;
    for (int i = 0; i < npairs; ++i) {
        auto e0 = h->entities[pairs[4 * i]], e1 = h->entities[pairs[4 * i + 2]];
        auto& col0 = h->registry.get<physecs::RigidBodyCollisionComponent>(e0).colliders[pairs[4 * i + 1]];
        auto& col1 = h->registry.get<physecs::RigidBodyCollisionComponent>(e1).colliders[pairs[4 * i + 3]];
        auto& t0 = h->registry.get<TransformComponent>(e0);
        auto& t1 = h->registry.get<TransformComponent>(e1);
        auto pos0 = t0.position + t0.orientation * col0.position;
        auto or0 = t0.orientation * col0.orientation;
        auto pos1 = t1.position + t1.orientation * col1.position;
        auto or1 = t1.orientation * col1.orientation;
        // same filters as the step: trigger pairs and nonCollidingPairs never reach collision() (Physecs.cpp:200-209)
        if (h->scene->contactFilter(col0.isTrigger, col0.data, col1.isTrigger, col1.data) == physecs::TRIGGER) continue;
        if (h->scene->nonCollidingPairs.count({ e0, e1 })) continue;
        buf.clear();
        if (!physecs::collision(pos0, or0, col0.geometry, pos1, or1, col1.geometry, buf)) continue;
        for (auto& r : buf) {
            if (!r.numPoints) continue;
            if (m < cap) {
                keys[5 * m] = i; keys[5 * m + 1] = r.triangleIndex; keys[5 * m + 2] = r.numPoints; keys[5 * m + 3] = 0; keys[5 * m + 4] = 0;
                for (int k = 0; k < 3; ++k) normal[3 * m + k] = r.normal[k];
                for (int p = 0; p < 4; ++p) for (int k = 0; k < 3; ++k) {
                    points[24 * m + 6 * p + k] = p < r.numPoints ? r.points[p].position0[k] : 0.f;
                    points[24 * m + 6 * p + 3 + k] = p < r.numPoints ? r.points[p].position1[k] : 0.f;
                }
            }
            ++m;
        }
    }
    return m;
}

// overlapping trigger pairs after the last simulate (triggerCache after the swap at Physecs.cpp:553)
int ph_num_triggers(void* hp) { return (int)((Harness*)hp)->scene->triggerCache.size(); }
void ph_get_triggers(void* hp, int* out) {
    auto* h = (Harness*)hp;
    size_t i = 0;
    for (auto& p : h->scene->triggerCache) {
        out[4 * i + 0] = (int)p.entity0; out[4 * i + 1] = p.colliderIndex0;
        out[4 * i + 2] = (int)p.entity1; out[4 * i + 3] = p.colliderIndex1;
        ++i;
    }
}
// register the recording listeners (Scene::addOnTriggerEnterCallback / addOnTriggerExitCallback)
void ph_record_trigger_events(void* hp) {
    auto* h = (Harness*)hp;
    h->scene->addOnTriggerEnterCallback(&h->recorder);
    h->scene->addOnTriggerExitCallback(&h->recorder);
}
// events since the last call, rows (kind, e0, c0, e1, c1); returns the count and clears the log
int ph_take_trigger_events(void* hp, int* out, int cap) {
    auto* h = (Harness*)hp;
    int n = (int)h->recorder.events.size();
    for (int i = 0; i < n && i < cap; ++i) for (int k = 0; k < 5; ++k) out[5 * i + k] = h->recorder.events[i][k];
    h->recorder.events.clear();
    return n;
}
// 0 = defaultContactFilter, 1 = filterParity, 2 = filterAsymmetric
void ph_set_contact_filter(void* hp, int mode) {
    auto* h = (Harness*)hp;
    h->scene->setContactFilter(mode == 1 ? filterParity : mode == 2 ? filterAsymmetric : physecs::defaultContactFilter);
}

// structural edits through the reference's public API (registry.destroy fires Scene::onRigidBodyDelete / onDynamicDelete)
void ph_destroy_entity(void* hp, int e) { auto* h = (Harness*)hp; h->registry.destroy(h->entities[e]); }
// the application reorders the dynamic-body pool (registry.sort): the reference indexes bodies by their CURRENT position in it every
// step (Physecs.cpp:116-117), so nothing else has to follow
void ph_sort_dynamic(void* hp, int greaterFirst) {     // EnTT iterates a pool back to front: a > b keeps a creation-ordered pool's packed order, a < b reverses it
    auto* h = (Harness*)hp;
    if (greaterFirst) h->registry.sort<physecs::RigidBodyDynamicComponent>([](const entt::entity a, const entt::entity b) { return a > b; });
    else h->registry.sort<physecs::RigidBodyDynamicComponent>([](const entt::entity a, const entt::entity b) { return a < b; });
}
void ph_destroy_joint(void* hp, int j) { auto* h = (Harness*)hp; h->scene->destroyJoint(h->joints[j]); h->joints[j] = nullptr; }
int ph_set_revolute_drive(void* hp, int j, int enabled, float velocity, float maxTorque) {
    auto* h = (Harness*)hp;
    auto* r = dynamic_cast<physecs::RevoluteJoint*>(h->joints[j]);
    if (!r) return -1;
    r->setDriveEnabled(enabled != 0); r->setDriveVelocity(velocity); r->setDriveMaxTorque(maxTorque);
    return 0;
}
void ph_add_collider(void* hp, int e, const float* lpos, const float* lquat, int type, const float* params, int mesh, const float* material, int flags, int data) {
    auto* h = (Harness*)hp;
    physecs::Collider col{};
    col.position = v3(lpos);
    col.orientation = q4(lquat);
    col.geometry = makeGeometry(h, type, params, mesh);
    col.material = { material[0], material[1], material[2] };
    col.isTrigger = flags & 1;
    col.enableSimulation = (flags >> 1) & 1;
    col.data = data;
    h->scene->addCollider(h->entities[e], col);
}
void ph_clear_colliders(void* hp, int e) { auto* h = (Harness*)hp; h->scene->clearColliders(h->entities[e]); }

// Scene::raycastClosest (filtered overload) with the filter `entity % mod != skip` (mod 0: accept all)
int ph_raycast(void* hp, const float* orig, const float* dir, float maxDist, int mod, int skip, float* hitPos3) {
    auto* h = (Harness*)hp;
    glm::vec3 hit(0);
    auto e = h->scene->raycastClosest(v3(orig), v3(dir), maxDist, [=](entt::entity x) { return mod == 0 || (int)((unsigned)x % (unsigned)mod) != skip; }, &hit);
    for (int k = 0; k < 3; ++k) hitPos3[k] = hit[k];
    return e == entt::null ? -1 : (int)e;
}
// Scene::overlap: rows (entity, colIndex); returns the count
int ph_overlap(void* hp, const float* pos, const float* quat, int type, const float* params, int mesh, int filter, int cap, int* out2) {
    auto* h = (Harness*)hp;
    auto hits = h->scene->overlap(v3(pos), q4(quat), makeGeometry(h, type, params, mesh), filter);
    for (size_t i = 0; i < hits.size() && (int)i < cap; ++i) { out2[2 * i] = (int)hits[i].entity; out2[2 * i + 1] = hits[i].colIndex; }
    return (int)hits.size();
}

// Scene::overlapWithMinTranslationalDistance: rows (entity, colIndex), (normal xyz, mtd); returns the count
int ph_overlap_mtd(void* hp, const float* pos, const float* quat, int type, const float* params, int mesh, int cap, int* out2, float* out4) {
    auto* h = (Harness*)hp;
    auto hits = h->scene->overlapWithMinTranslationalDistance(v3(pos), q4(quat), makeGeometry(h, type, params, mesh));
    for (size_t i = 0; i < hits.size() && (int)i < cap; ++i) {
        out2[2 * i] = (int)hits[i].entity; out2[2 * i + 1] = hits[i].colIndex;
        for (int k = 0; k < 3; ++k) out4[4 * i + k] = hits[i].normal[k];
        out4[4 * i + 3] = hits[i].mtd;
    }
    return (int)hits.size();
}

// Bring broadPhaseEntries into the order the reference's own insertion sort (Physecs.cpp:121-133) leaves them in -- that sort
// shifts an entry left only past entries with a strictly larger bounds.min.x, i.e. it is a stable sort by min.x -- but in
// O(n log n).  The first simulate() of a freshly filled 100k..1M-collider scene otherwise spends minutes in that O(n^2) pass
// (75 s at 100k, 570 s at 250k); results are unchanged (same final order), only the checker gets usable at full size.
void ph_presort(void* hp) {
    auto* h = (Harness*)hp;
    auto& be = h->scene->broadPhaseEntries;
    std::stable_sort(be.begin(), be.end(), [](const auto& a, const auto& b) { return a.bounds.min.x < b.bounds.min.x; });
    for (size_t i = 0; i < be.size(); ++i) h->scene->colToBroadPhaseEntry[{ be[i].entity, be[i].colliderIndex }] = (int)i;
}

int ph_num_dynamic(void* hp) {
    auto* h = (Harness*)hp;
    return (int)h->registry.storage<physecs::RigidBodyDynamicComponent>().size();
}

} // extern "C"

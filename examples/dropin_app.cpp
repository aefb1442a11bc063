// An application written against the REFERENCE's public API only -- the idiom of its demo (demo/Demo.cpp:29-60: aggregate-initialised
// colliders, computeCOMAndInvInertiaTensor, the three components emplaced in that order), its joints, trigger listeners, contact
// filter, collider / body edits and scene queries (include/Physecs/Physecs.h:199-228).  Nothing in this file names physecs_b200: the
// same source builds against the reference's headers + library and against this repo's include/Physecs + physecs_b200/host/*.cpp +
// libphysecs_b200.so (way A).  tests/test_cpu_dropin.py compiles, links and runs it both ways.
#include <Physecs/Physecs.h>
#include <Physecs/Components.h>
#include <Physecs/MassUtil.h>
#include <Physecs/Joints/RevoluteJoint.h>
#include <Physecs/Joints/SphericalJoint.h>
#include <Physecs/Joints/FixedJoint.h>
#include <Physecs/ConvexMesh.h>
#include <Physecs/TriangleMesh.h>
#include <Transform.h>
#ifdef DROPIN_WITH_CHARACTER_CONTROLLER
#include <CharacterController.h>       // the reference's own file, unchanged
#endif

#include <cstdio>
#include <vector>

namespace {

struct Events : physecs::OnTriggerEnterListener, physecs::OnTriggerExitListener {
    int entered = 0, left = 0;
    void onTriggerEnter(entt::entity, int, entt::entity, int) override { ++entered; }
    void onTriggerExit(entt::entity, int, entt::entity, int) override { ++left; }
};

physecs::ContactType sensorsOnly(bool isTrigger0, int data0, bool isTrigger1, int data1) {
    if ((isTrigger0 || isTrigger1) && data0 != 7 && data1 != 7) return physecs::TRIGGER;
    return physecs::COLLISION;
}

entt::entity spawn(entt::registry& registry, glm::vec3 position, physecs::Geometry geometry, bool dynamic, bool trigger = false, int data = 0) {
    auto e = registry.create();
    registry.emplace<TransformComponent>(e, position, glm::quat(1, 0, 0, 0), glm::vec3(1, 1, 1));
    physecs::Collider collider = { glm::vec3(0), glm::quat(1, 0, 0, 0), geometry, { 0.4, 0.2 }, trigger, true, data };
    auto col = registry.emplace<physecs::RigidBodyCollisionComponent>(e, std::vector{ collider });
    if (dynamic) {
        glm::vec3 com;
        glm::mat3 invInertiaTensor;
        physecs::computeCOMAndInvInertiaTensor(col, 1.f, com, invInertiaTensor);
        registry.emplace<physecs::RigidBodyDynamicComponent>(e, false, glm::vec3(0), glm::vec3(0), 1.f, com, invInertiaTensor);
    }
    return e;
}

// a unit cube as a convex mesh: vertices + faces (indices, outward normal, centroid), as the reference's ConvexMesh wants them
physecs::ConvexMesh makeCubeHull() {
    std::vector<glm::vec3> v;
    for (int i = 0; i < 8; ++i) v.emplace_back((i & 1) ? 0.5f : -0.5f, (i & 2) ? 0.5f : -0.5f, (i & 4) ? 0.5f : -0.5f);
    const int quads[6][4] = { { 0, 4, 6, 2 }, { 1, 3, 7, 5 }, { 0, 1, 5, 4 }, { 2, 6, 7, 3 }, { 0, 2, 3, 1 }, { 4, 5, 7, 6 } };
    const glm::vec3 normals[6] = { { -1, 0, 0 }, { 1, 0, 0 }, { 0, -1, 0 }, { 0, 1, 0 }, { 0, 0, -1 }, { 0, 0, 1 } };
    std::vector<physecs::ConvexMeshFace> faces;
    for (int f = 0; f < 6; ++f) {
        physecs::ConvexMeshFace face;
        glm::vec3 c(0);
        for (int k = 0; k < 4; ++k) { face.indices.push_back(quads[f][k]); c += v[quads[f][k]]; }
        face.normal = normals[f];
        face.centroid = c / 4.f;
        faces.push_back(face);
    }
    return physecs::ConvexMesh(std::move(v), std::move(faces));
}

// a bumpy 12 x 12 patch of triangles
physecs::TriangleMesh makePatch(glm::vec3 origin) {
    const int n = 12;
    std::vector<glm::vec3> v;
    std::vector<unsigned int> idx;
    for (int z = 0; z <= n; ++z)
        for (int x = 0; x <= n; ++x) v.push_back(origin + glm::vec3(x * 0.5f, 0.1f * ((x * 7 + z * 3) % 5), z * 0.5f));
    for (int z = 0; z < n; ++z)
        for (int x = 0; x < n; ++x) {
            unsigned a = z * (n + 1) + x, b = a + 1, c = a + n + 1, d = c + 1;
            idx.insert(idx.end(), { a, c, b, b, c, d });
        }
    return physecs::TriangleMesh(v, idx);
}

} // namespace

int main() {
    entt::registry registry;
    physecs::Scene scene(registry, 2);
    scene.setNumSubSteps(4);
    scene.setNumIterations(2);
    scene.setGravity(9.81f);

    physecs::Geometry ground = { physecs::BOX };   ground.box = { glm::vec3(50, 1, 50) };
    physecs::Geometry box = { physecs::BOX };      box.box = { glm::vec3(0.5f, 0.5f, 0.5f) };
    physecs::Geometry ball = { physecs::SPHERE };  ball.sphere = { 0.4f };
    physecs::Geometry pill = { physecs::CAPSULE }; pill.capsule = { 0.5f, 0.25f };

    auto floor = spawn(registry, glm::vec3(0, -1, 0), ground, false);
    std::vector<entt::entity> bodies;
    for (int i = 0; i < 24; ++i)
        bodies.push_back(spawn(registry, glm::vec3((i % 6) * 1.5f - 4.f, 0.6f + (i / 6) * 1.2f, (i % 3) * 0.1f), i % 3 == 0 ? box : i % 3 == 1 ? ball : pill, true));
    auto sensor = spawn(registry, glm::vec3(0, 0.5f, 0), box, false, true, 1);

    // user-owned meshes (Colliders.h:24-31: raw pointers that must outlive the Scene): hulls dropped onto a triangle-mesh patch
    physecs::ConvexMesh cube = makeCubeHull();
    physecs::TriangleMesh patch = makePatch(glm::vec3(30, 0, 30));
    physecs::Geometry hull = { physecs::CONVEX_MESH };       hull.convex = { &cube, glm::vec3(0.8f, 0.6f, 0.8f) };
    physecs::Geometry terrain = { physecs::TRIANGLE_MESH };  terrain.triangleMesh = { &patch };
    spawn(registry, glm::vec3(0, 0, 0), terrain, false);
    std::vector<entt::entity> onPatch;
    for (int i = 0; i < 6; ++i) onPatch.push_back(spawn(registry, glm::vec3(31.f + i * 0.9f, 1.2f, 32.f + (i % 2)), i % 2 ? hull : ball, true));
    // the triangle mesh as the application sees it: post-build triangle order and BVH (TriangleMesh.h: public members)
    unsigned long long meshPrint = 1469598103934665603ull;
    for (auto& t : patch.triangles) for (unsigned k : t.indices) { meshPrint ^= k; meshPrint *= 1099511628211ull; }
    for (auto& nd : patch.bvh) { meshPrint ^= (unsigned)nd.triCount * 977u + (unsigned)nd.index; meshPrint *= 1099511628211ull; }
    std::printf("mesh: %zu triangles %zu nodes order %016llx, %zu hit by a query box\n", patch.triangles.size(), patch.bvh.size(), meshPrint,
                patch.overlapBvh({ glm::vec3(31, -1, 31), glm::vec3(32.2f, 1, 32.2f) }).size());

    Events events;
    scene.addOnTriggerEnterCallback(&events);
    scene.addOnTriggerExitCallback(&events);
    scene.setContactFilter(sensorsOnly);

    // a three-link chain hanging from the floor entity, one driven hinge
    auto* hinge = scene.createJoint<physecs::RevoluteJoint>(floor, glm::vec3(8, 4, 0), glm::quat(1, 0, 0, 0), bodies[0], glm::vec3(0, 0.5f, 0), glm::quat(1, 0, 0, 0));
    hinge->setDriveEnabled(true); hinge->setDriveVelocity(1.f); hinge->setDriveMaxTorque(5.f);
    auto* socket = scene.createJoint<physecs::SphericalJoint>(bodies[0], glm::vec3(0, -0.5f, 0), glm::quat(1, 0, 0, 0), bodies[1], glm::vec3(0, 0.4f, 0), glm::quat(1, 0, 0, 0));
    auto* weld = scene.createJoint<physecs::FixedJoint>(bodies[1], glm::vec3(0, -0.4f, 0), glm::quat(1, 0, 0, 0), bodies[2], glm::vec3(0, 0.75f, 0), glm::quat(1, 0, 0, 0));
    scene.setCanCollide(bodies[3], bodies[4], false);

    for (int step = 0; step < 90; ++step) {
        if (step == 20) bodies.push_back(spawn(registry, glm::vec3(0, 6, 0), ball, true));                 // a body appears
        if (step == 30) { registry.destroy(bodies[5]); bodies[5] = entt::null; }                            // ... one disappears
        if (step == 40) scene.setIsKinematic(bodies[6], true);
        if (step == 45) {                                                                                    // a second collider on a live body
            physecs::Collider extra = { glm::vec3(0.5f, 0, 0), glm::quat(1, 0, 0, 0), ball, { 0.4, 0.2 }, false, true, 0 };
            scene.addCollider(bodies[7], extra);
        }
        if (step == 50) {                                                                                    // a static body moved by the application
            registry.patch<TransformComponent>(sensor, [](TransformComponent& t) { t.position.x += 0.5f; });
        }
        if (step == 60) { scene.destroyJoint(weld); weld = nullptr; }
        if (step == 70) hinge->setDriveVelocity(-1.f);
        scene.simulate(1.f / 60.f);
    }
    (void)socket;

    glm::vec3 hitPos(0);
    entt::entity below = scene.raycastClosest(glm::vec3(20, 5, 20), glm::vec3(0, -1, 0), 100.f, &hitPos);
    entt::entity filtered = scene.raycastClosest(glm::vec3(20, 5, 20), glm::vec3(0, -1, 0), 100.f, [floor](entt::entity e) { return e != floor; });
    auto touching = scene.overlap(glm::vec3(0, 0.5f, 0), glm::quat(1, 0, 0, 0), box, 0);
    auto pushed = scene.overlapWithMinTranslationalDistance(glm::vec3(20, -0.2f, 20), glm::quat(1, 0, 0, 0), ball);
#ifdef DROPIN_WITH_CHARACTER_CONTROLLER
    auto walker = registry.create();
    registry.emplace<TransformComponent>(walker, glm::vec3(20, 0.05f, 20), glm::quat(1, 0, 0, 0), glm::vec3(1, 1, 1));
    physecs::CharacterController controller(scene, walker, 1.2f, 0.3f);
    auto flags = controller.move(glm::vec3(0.1f, -0.2f, 0), 1.f / 60.f);
    std::printf("character: down %d side %d up %d\n", (int)flags.isDown(), (int)flags.isSide(), (int)flags.isUp());
#endif
    scene.removeOnTriggerEnterCallback(&events);
    scene.removeOnTriggerExitCallback(&events);

    float patchLowest = 1e9f;
    for (auto e : onPatch) patchLowest = std::min(patchLowest, registry.get<TransformComponent>(e).position.y);
    std::printf("on the patch: lowest y %.3f\n", patchLowest);
    float lowest = 1e9f;
    for (auto e : bodies) if (e != entt::null) lowest = std::min(lowest, registry.get<TransformComponent>(e).position.y);
    std::printf("bodies %zu lowest y %.3f ray hit %d (y %.3f) filtered ray hit %d overlaps %zu mtd rows %zu trigger enter %d exit %d contacts %zu\n",
                bodies.size(), lowest, below == entt::null ? -1 : (int)entt::to_integral(below), hitPos.y, filtered == entt::null ? -1 : (int)entt::to_integral(filtered),
                touching.size(), pushed.size(), events.entered, events.left, scene.getContactPoints().size());
    return 0;
}

// physecs_b200 host layer -- physecs::Scene over an entt::registry, source-compatible with the reference
// (include/Physecs/Physecs.h:40-46 listeners / ContactType, :199-228 Scene's public section).
//
// The Scene keeps the reference's contract with the application:
//   * it borrows the registry, connects the same five EnTT signals (src/Physecs.cpp:92-98) and never scans for bodies
//     the signals did not announce;
//   * simulate(dt) reads TransformComponent / RigidBodyDynamicComponent of every body from the registry, advances the
//     world by dt with the scene's substep / iteration / gravity knobs, and writes poses and velocities back;
//   * registry.patch<TransformComponent>(e) announces a moved static / kinematic body (bounds follow, +0.01 margin).
// What it does NOT do is compute: broadphase, narrowphase, constraint build and the TGS substep loop all run on the GPU
// behind the C ABI of include/physecs_b200.h.  The host side gathers registry state into pinned SoA staging buffers
// (multi-threaded when numThreads > 0), calls pb_set_state / pb_step / pb_get_state and scatters the result.
// There is no CPU fallback: constructing a Scene without a CUDA device makes the first simulate() throw.
#pragma once
#include <functional>
#include "../BVH.h"
#include <memory>
#include <vector>
#include <entt.hpp>
#include <glm/glm.hpp>
#include "b200_types.hpp"
#include "b200_joints.hpp"

struct pb_ctx;

namespace physecs {

class OnTriggerEnterListener {
public:
    virtual void onTriggerEnter(entt::entity, int, entt::entity, int) = 0;
    virtual ~OnTriggerEnterListener() = default;
};

class OnTriggerExitListener {
public:
    virtual void onTriggerExit(entt::entity, int, entt::entity, int) = 0;
    virtual ~OnTriggerExitListener() = default;
};

struct OverlapHit {
    entt::entity entity;
    int colIndex;
};

struct OverlapMtdHit {
    entt::entity entity;
    int colIndex;
    glm::vec3 normal;   // collider -> query shape
    float mtd;          // deepest penetration along the normal (>= 0)
};

enum ContactType { COLLISION, TRIGGER };
PHYSECS_API ContactType defaultContactFilter(bool isTrigger0, int data0, bool isTrigger1, int data1);

class PHYSECS_API Scene {
    struct Impl;
    entt::registry& registry;
    int numSubSteps = 8;      // reference defaults, Physecs.h:156-159
    int numIterations = 2;
    float g = 9.81f;
    std::unique_ptr<Impl> impl;

    void onRigidBodyCreate(entt::registry&, entt::entity);
    void onRigidBodyDelete(entt::registry&, entt::entity);
    void onRigidBodyUpdate(entt::registry&, entt::entity);
    void onRigidBodyMove(entt::registry&, entt::entity);
    void onDynamicCreate(entt::registry&, entt::entity);
    void onDynamicDelete(entt::registry&, entt::entity);
    void addJoint(Joint* joint);
    void prepareDevice();

public:
    // numThreads: host worker threads for the registry gather / scatter (the reference's pool runs its narrowphase,
    // ThreadPool.cpp:30-46; here that stage is device work).  The README's one-argument form is kept.
    explicit Scene(entt::registry& registry, int numThreads = 0);
    ~Scene();
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    void setNumSubSteps(int numSubSteps) { this->numSubSteps = numSubSteps; }
    void setNumIterations(int numIterations) { this->numIterations = numIterations; }
    void setGravity(float gravity) { this->g = gravity; }
    // == reference Scene::simulate (src/Physecs.cpp:112-561).  Throws std::runtime_error on CUDA failure.
    void simulate(float timeStep);

    // Scene queries (reference Physecs.h:205-207).  They see the state of the last simulate() plus every change announced since
    // (structural edits, registry.patch<TransformComponent>).  Hits are exact per collider; results of overlap() are sorted by
    // (entity, collider index).  Triangle-mesh colliders are invisible to raycastClosest and overlap, as in the reference (no ray / overlap routine).
    entt::entity raycastClosest(glm::vec3 rayOrig, glm::vec3 rayDir, float maxDistance, glm::vec3* hitPos = nullptr);
    entt::entity raycastClosest(glm::vec3 rayOrig, glm::vec3 rayDir, float maxDistance, const std::function<bool(entt::entity)>& filter, glm::vec3* hitPos = nullptr);
    std::vector<OverlapHit> overlap(glm::vec3 pos, glm::quat ori, Geometry geometry, int filter);
    // one hit per contact manifold between a collider (triangle meshes included: one per touched triangle) and the query shape
    std::vector<OverlapMtdHit> overlapWithMinTranslationalDistance(glm::vec3 pos, glm::quat ori, Geometry geometry);

    template <typename T>
    T* createJoint(entt::entity entity0, glm::vec3 anchor0Pos, glm::quat anchor0Or, entt::entity entity1, glm::vec3 anchor1Pos, glm::quat anchor1Or) {
        T* joint = new T(entity0, anchor0Pos, anchor0Or, entity1, anchor1Pos, anchor1Or);
        addJoint(joint);
        return joint;
    }
    void destroyJoint(Joint* joint);
    void clearColliders(entt::entity entity);
    void addCollider(entt::entity entity, const Collider& collider);
    void setIsKinematic(entt::entity entity, bool isKinematic);
    void addOnTriggerEnterCallback(OnTriggerEnterListener* callback);
    void addOnTriggerExitCallback(OnTriggerExitListener* callback);
    void removeOnTriggerEnterCallback(OnTriggerEnterListener* callback);
    void removeOnTriggerExitCallback(OnTriggerExitListener* callback);
    void setCanCollide(entt::entity entity0, entt::entity entity1, bool canCollide);
    void setContactFilter(ContactType (*filter)(bool, int, bool, int));
    entt::registry& getRegistry() { return registry; }
    // world-space contact points (position1 of every manifold point) of the last step -- debug getter, downloads on call
    const std::vector<glm::vec3>& getContactPoints();
    // snapshot of the device tree in the reference's node format (debug getters, Physecs.h:227-228); root = getBHVRootId()
    const std::vector<BVHNode>& getBVH();
    const int getBHVRootId();

    // ---- additions of this implementation ----------------------------------------------------------------------------------
    void setDevice(int cudaDevice);          // before the first simulate(); default 0
    // Per-step host <-> device traffic.  SYNC_FULL (default) mirrors the reference exactly: every body's pose and velocity is
    // read from the registry before the step and written back after it.  SYNC_DEVICE_AUTHORITATIVE skips the pre-step read of
    // dynamic bodies the application did not announce through registry.patch<TransformComponent> / notifyBodyChanged, for
    // applications that never write simulated bodies directly (removes the gather and the H2D copy from the step).
    enum SyncMode { SYNC_FULL, SYNC_DEVICE_AUTHORITATIVE };
    void setSyncMode(SyncMode mode);
    void notifyBodyChanged(entt::entity entity);   // velocity / pose of a dynamic body was written directly (device-authoritative mode)
    // Initial sizes of the per-step device arenas (candidate pairs, manifolds).  Defaults scale with the collider count
    // (8x / 6x); a step that overflows an arena is re-run transparently with 4x larger arenas (nothing is written back
    // to the registry before the step has succeeded).
    void setArenaCapacity(int maxPairs, int maxManifolds);
    pb_ctx* nativeContext();                 // the C-ABI context (parity taps: pb_get_pairs / pb_get_manifolds / pb_get_bounds ...)
    struct StepStats { int pairs, manifolds, points, colors, triggers; float deviceMs; double gatherMs, scatterMs, totalMs, prepareMs; };   // prepareMs: bringing the device scene description up to date (structural edits, joints, filters) -- part of totalMs
    StepStats getLastStepStats() const;
};

} // namespace physecs

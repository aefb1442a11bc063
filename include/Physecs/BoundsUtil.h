// Public bounds helpers of the reference API (reference include/Physecs/BoundsUtil.h:8-18, src/BoundsUtil.cpp:6-104), host code.
// Applications use them next to the Scene (culling, spawn placement, their own spatial queries); the character controller of the
// reference builds its query boxes with them.  Header-only over GLM: the same expressions in the same order as the reference, so
// the values equal the collider bounds the device computes for the broadphase (csrc/np_bounds.cuh restates the same lines).
#pragma once
#include <limits>
#include <entt.hpp>
#include <glm/gtc/quaternion.hpp>
#include "detail/b200_types.hpp"

namespace physecs {

inline Bounds getBoundsSphere(glm::vec3 pos, float radius) { return { pos - glm::vec3(radius), pos + glm::vec3(radius) }; }      // BoundsUtil.cpp:36-38

inline Bounds getBoundsCapsule(glm::vec3 pos, glm::quat ori, float halfHeight, float radius) {                                  // :40-46
    const glm::vec3 p0 = pos + ori * glm::vec3(0, halfHeight, 0), p1 = pos + ori * glm::vec3(0, -halfHeight, 0);
    return { glm::min(p0, p1) - radius, glm::max(p0, p1) + radius };
}

inline Bounds getBoundsBox(glm::vec3 pos, glm::quat ori, glm::vec3 halfExtents) {                                               // :49-60
    const glm::mat3 u = glm::mat3_cast(ori);      // == glm::toMat3
    const glm::vec3 w = glm::mat3(glm::abs(u[0]), glm::abs(u[1]), glm::abs(u[2])) * halfExtents;
    return { pos - w, pos + w };
}

inline Bounds getBoundsConvexMesh(glm::vec3 pos, glm::quat ori, ConvexMesh* mesh, glm::vec3 scale) {                            // :62-69 (every padded vertex)
    glm::vec3 mn(std::numeric_limits<float>::max()), mx(std::numeric_limits<float>::lowest());
    for (const auto& v : mesh->vertices) { const glm::vec3 p = pos + ori * (scale * v); mn = glm::min(mn, p); mx = glm::max(mx, p); }
    return { mn, mx };
}

inline Bounds getBoundsTriangle(glm::vec3 a, glm::vec3 b, glm::vec3 c) { return { glm::min(glm::min(a, b), c), glm::max(glm::max(a, b), c) }; }   // :71-75

inline Bounds getBoundsTriangleMesh(glm::vec3 pos, glm::quat ori, TriangleMesh* mesh) {                                          // :77-85
    glm::vec3 mn(std::numeric_limits<float>::max()), mx(std::numeric_limits<float>::lowest());
    for (const auto& v : mesh->vertices) { const glm::vec3 p = pos + ori * v; mn = glm::min(mn, p); mx = glm::max(mx, p); }
    return { mn, mx };
}

inline Bounds getBounds(glm::vec3 pos, glm::quat ori, const Geometry& geom) {                                                    // :23-33
    switch (geom.type) {
        case SPHERE: return getBoundsSphere(pos, geom.sphere.radius);
        case CAPSULE: return getBoundsCapsule(pos, ori, geom.capsule.halfHeight, geom.capsule.radius);
        case BOX: return getBoundsBox(pos, ori, geom.box.halfExtents);
        case CONVEX_MESH: return getBoundsConvexMesh(pos, ori, geom.convex.mesh, geom.convex.scale);
        case TRIANGLE_MESH: return getBoundsTriangleMesh(pos, ori, geom.triangleMesh.mesh);
    }
    return Bounds{ glm::vec3(0), glm::vec3(0) };
}

inline bool intersects(const Bounds& a, const Bounds& b) {                                                                       // :87-92 (closed intervals)
    if (a.max[0] < b.min[0] || a.min[0] > b.max[0]) return false;
    if (a.max[1] < b.min[1] || a.min[1] > b.max[1]) return false;
    if (a.max[2] < b.min[2] || a.min[2] > b.max[2]) return false;
    return true;
}

inline Bounds getUnion(const Bounds& a, const Bounds& b) { return { glm::min(a.min, b.min), glm::max(a.max, b.max) }; }          // :94-104

// union over every collider of an entity at its current transform (BoundsUtil.cpp:6-21); an entity without colliders gives the zero box
inline Bounds getBounds(entt::registry& registry, entt::entity entity) {
    const auto& t = registry.get<::TransformComponent>(entity);
    const auto& col = registry.get<RigidBodyCollisionComponent>(entity);
    Bounds b{ glm::vec3(0), glm::vec3(0) };
    for (size_t i = 0; i < col.colliders.size(); ++i) {
        const auto& c = col.colliders[i];
        const Bounds cb = getBounds(t.position + t.orientation * c.position, t.orientation * c.orientation, c.geometry);
        b = i ? getUnion(b, cb) : cb;
    }
    return b;
}

} // namespace physecs

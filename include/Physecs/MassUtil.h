// Setup-time mass helpers of the reference API (include/Physecs/MassUtil.h:7-12), host code.
// They are not part of simulate(); they exist so applications that call physecs::computeCOMAndInvInertiaTensor /
// setMassProps when spawning bodies (reference demo/Demo.cpp:42-45) keep compiling and get the same numbers,
// including the reference's box convention (half extents in the full-extent formula, volume = hx*hy*hz).
#pragma once
#include <array>
#include "detail/b200_types.hpp"

namespace physecs {
PHYSECS_API glm::mat3 getInertiaSphere(float mass, float radius);
PHYSECS_API glm::mat3 getInertiaCapsule(float mass, float halfHeight, float radius);
PHYSECS_API glm::mat3 getInertiaBox(float mass, glm::vec3 halfExtents);
PHYSECS_API glm::mat3 getInertiaTetrahedron(float mass, const std::array<glm::vec3, 4>& v);
PHYSECS_API void computeCOMAndInvInertiaTensor(const RigidBodyCollisionComponent& collisionComponent, float mass, glm::vec3& com, glm::mat3& invInertiaTensor);
PHYSECS_API void setMassProps(RigidBodyDynamicComponent& dynamicComponent, const RigidBodyCollisionComponent& collisionComponent, float mass);
}

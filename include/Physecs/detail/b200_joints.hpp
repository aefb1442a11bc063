// physecs_b200 host layer -- the Joint class surface of the reference (include/Physecs/Joint.h:47-70 and
// include/Physecs/Joints/{Fixed,Revolute,Spherical,Universal,Prismatic,Gear,Servo}Joint.h), source-compatible:
// same class names, constructor signature, getters and setters.
//
// What differs from the reference: a joint here is a DESCRIPTION.  The reference lowers joints to 1-D constraint rows on
// the CPU through Joint::getSolverDesc / makeConstraints (src/Joints/*.cpp); on this path the row builders are device
// code (physecs_b200/csrc/joints.cu) selected by `kind`, and the per-type knobs travel as eight floats (the params8
// layout of pb_upload_joints).  User-defined Joint subclasses therefore have no device row builder: Scene::createJoint
// accepts only the seven built-in types.
#pragma once
#include <entt.hpp>
#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include "b200_types.hpp"

namespace physecs {

class Scene;

class PHYSECS_API Joint {
    friend class Scene;
    int color = -1;
    int kind;                 // PB_JOINT_* of include/physecs_b200.h
    bool paramsDirty = true;  // a setter ran since the parameters were last sent to the device

protected:
    entt::entity entity0;
    entt::entity entity1;
    glm::vec3 anchor0Pos;
    glm::quat anchor0Or;
    glm::vec3 anchor1Pos;
    glm::quat anchor1Or;
    float params[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };

    Joint(int kind, entt::entity e0, glm::vec3 a0p, glm::quat a0q, entt::entity e1, glm::vec3 a1p, glm::quat a1q)
        : kind(kind), entity0(e0), entity1(e1), anchor0Pos(a0p), anchor0Or(a0q), anchor1Pos(a1p), anchor1Or(a1q) {}
    void setParam(int i, float v) { if (params[i] != v) { params[i] = v; paramsDirty = true; } }

public:
    entt::entity getEntity0() const { return entity0; }
    entt::entity getEntity1() const { return entity1; }
    glm::vec3 getAnchor0Pos() const { return anchor0Pos; }
    glm::quat getAnchor0Or() const { return anchor0Or; }
    glm::vec3 getAnchor1Pos() const { return anchor1Pos; }
    glm::quat getAnchor1Or() const { return anchor1Or; }
    void setColor(int c) { color = c; }
    int getColor() const { return color; }    // 0..7 parallel colours, 8 = sequential overflow bucket
    virtual ~Joint() = default;
};

#define PHYSECS_B200_JOINT_CTOR(Name, Kind) \
    Name(entt::entity e0, glm::vec3 a0p, glm::quat a0q, entt::entity e1, glm::vec3 a1p, glm::quat a1q) : Joint(Kind, e0, a0p, a0q, e1, a1p, a1q)

// kinds: 0 fixed, 1 revolute, 2 spherical, 3 universal, 4 prismatic, 5 gear, 6 servo

// all six relative degrees of freedom locked (FixedJoint.cpp:6-34)
class PHYSECS_API FixedJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(FixedJoint, 0) {}
};

// hinge about the anchor frames' x axis, optional velocity motor with a torque limit (RevoluteJoint.cpp:8-50)
class PHYSECS_API RevoluteJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(RevoluteJoint, 1) {}
    void setDriveEnabled(bool enabled) { setParam(0, enabled ? 1.f : 0.f); }
    void setDriveVelocity(float velocity) { setParam(1, velocity); }
    void setDriveMaxTorque(float maxTorque) { setParam(2, maxTorque); }
};

// ball and socket (SphericalJoint.cpp:8-12)
class PHYSECS_API SphericalJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(SphericalJoint, 2) {}
};

// U-joint: point-to-point + the two frames' z axes kept perpendicular (UniversalJoint.cpp:8-20)
class PHYSECS_API UniversalJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(UniversalJoint, 3) { }
};

// slider along the anchor frames' x axis with limits and an optional soft position drive (PrismaticJoint.cpp:6-114;
// defaults PrismaticJoint.h:8-17)
class PHYSECS_API PrismaticJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(PrismaticJoint, 4) { params[0] = 1.f; params[1] = 0.f; params[4] = 5.f; params[5] = 1.f; }
    void setUpperLimit(float v) { setParam(0, v); }
    void setLowerLimit(float v) { setParam(1, v); }
    void setDriveEnabled(bool enabled) { setParam(2, enabled ? 1.f : 0.f); }
    void setTargetPosition(float v) { setParam(3, v); }
    void setDriveStiffness(float v) { setParam(4, v); }
    void setDriveDamping(float v) { setParam(5, v); }
};

// couples the rotation of two bodies about their anchor x axes with a ratio; the accumulated angles are device-side
// state that persists across steps (GearJoint.cpp:10-50; default ratio GearJoint.h:8)
class PHYSECS_API GearJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(GearJoint, 5) { params[0] = 1.f; }
    void setGearRatio(float ratio) { setParam(0, ratio); }
};

// hinge with a soft angular position drive (ServoJoint.cpp:9-49; defaults ServoJoint.h:8-12)
class PHYSECS_API ServoJoint final : public Joint {
public:
    PHYSECS_B200_JOINT_CTOR(ServoJoint, 6) { params[1] = 30.f; params[2] = 1.f; }
    void setTargetAngle(float angle) { setParam(0, angle); }
    void setDriveStiffness(float stiffness) { setParam(1, stiffness); }
    void setDriveDamping(float damping) { setParam(2, damping); }
};

#undef PHYSECS_B200_JOINT_CTOR

} // namespace physecs

// Setup-time mass properties (include/Physecs/MassUtil.h).  Same formulas as the reference helper
// (src/MassUtil.cpp:6-192) so bodies spawned through either library start from identical invMass / com / inertia:
//   * solid sphere 2/5 m r^2; capsule = cylinder + two hemispheres split by volume (:10-21);
//   * box: m/12 (y^2 + z^2) evaluated on HALF extents and "volume" hx*hy*hz (:23-29, :92) -- the reference's convention,
//     kept because applications tune masses against it;
//   * convex mesh: fan of tetrahedra around the mean of the PADDED vertex buffer divided by the unpadded count (:99-103);
//   * compound bodies: volume-weighted centre of mass, parallel-axis shift per collider; trigger colliders carry no mass.
#include <MassUtil.h>
#include <glm/gtc/constants.hpp>
#include <glm/gtx/quaternion.hpp>
#include <glm/gtx/norm.hpp>

namespace physecs {

static glm::mat3 diag(float a, float b, float c) {
    glm::mat3 m(0.f);
    m[0][0] = a; m[1][1] = b; m[2][2] = c;
    return m;
}

glm::mat3 getInertiaSphere(float mass, float radius) {
    float i = 2.f * mass * radius * radius / 5.f;
    return diag(i, i, i);
}

glm::mat3 getInertiaCapsule(float mass, float halfHeight, float radius) {
    const float pi = glm::pi<float>();
    float volCyl = 2 * radius * radius * pi * halfHeight;
    float volSph = 4 * glm::pow(radius, 3) * pi / 3.f;
    float vol = volCyl + volSph;
    float mCyl = volCyl * mass / vol, mSph = volSph * mass / vol;
    float side = mCyl * (halfHeight * halfHeight / 3.f + radius * radius / 4.f) +
                 mSph * (halfHeight * halfHeight + 3.f * halfHeight * radius / 4.f + 2 * radius * radius / 5.f);
    float axial = mCyl * radius * radius / 2.f + mSph * 2 * radius * radius / 5.f;
    return diag(side, axial, side);
}

glm::mat3 getInertiaBox(float mass, glm::vec3 he) {
    float m = mass / 12.f;
    float xx = he.x * he.x, yy = he.y * he.y, zz = he.z * he.z;
    return diag(m * (yy + zz), m * (xx + zz), m * (xx + yy));
}

// inertia of a tetrahedron about the origin (vertices given relative to it)
glm::mat3 getInertiaTetrahedron(float mass, const std::array<glm::vec3, 4>& v) {
    glm::vec3 sq(0);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j <= i; j++) sq += v[j] * v[i];
    float a = (sq.y + sq.z) / 10.f, b = (sq.x + sq.z) / 10.f, c = (sq.x + sq.y) / 10.f;
    float yz = 0, xz = 0, xy = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float w = i == j ? 2 : 1;
            yz += w * v[i].y * v[j].z;
            xz += w * v[i].x * v[j].z;
            xy += w * v[i].x * v[j].y;
        }
    yz /= 20.f; xz /= 20.f; xy /= 20.f;
    glm::mat3 I;
    I[0][0] = a; I[1][1] = b; I[2][2] = c;
    I[0][1] = I[1][0] = -xz;     // the reference stores b' (x.z products) in the xy slot and c' (x.y) in xz (MassUtil.cpp:62-68)
    I[2][0] = I[0][2] = -xy;
    I[1][2] = I[2][1] = -yz;
    return mass * I;
}

namespace {
float colliderVolume(const Collider& c) {
    const float pi = glm::pi<float>();
    switch (c.geometry.type) {
        case SPHERE: return 4 * pi * glm::pow(c.geometry.sphere.radius, 3) / 3.f;
        case CAPSULE: {
            float r = c.geometry.capsule.radius;
            return 4 * pi * glm::pow(r, 3) / 3.f + r * r * pi * c.geometry.capsule.halfHeight * 2.f;
        }
        case BOX: return c.geometry.box.halfExtents.x * c.geometry.box.halfExtents.y * c.geometry.box.halfExtents.z;
        default: return 0.f;
    }
}
glm::vec3 paddedCenter(const ConvexMeshGeometry& g) {
    glm::vec3 center(0);
    for (auto& v : g.mesh->vertices) center += g.scale * v;
    center /= g.mesh->vertices.size();
    return center;
}
template <class F> void forEachTet(const ConvexMeshGeometry& g, glm::vec3 center, F&& f) {
    for (auto& face : g.mesh->faces) {
        glm::vec3 v0 = g.scale * g.mesh->vertices[face.indices[0]];
        for (int i = 1; i < (int)face.indices.size() - 1; i++) {
            glm::vec3 v1 = g.scale * g.mesh->vertices[face.indices[i]];
            glm::vec3 v2 = g.scale * g.mesh->vertices[face.indices[i + 1]];
            float volume = glm::abs(glm::dot(glm::cross(v0 - center, v1 - center), v2 - center)) / 6.f;
            f(v0, v1, v2, volume);
        }
    }
}
glm::mat3 parallelAxis(float m, glm::vec3 r) { return m * (glm::mat3(glm::length2(r)) - glm::outerProduct(r, r)); }
} // namespace

void computeCOMAndInvInertiaTensor(const RigidBodyCollisionComponent& cc, float mass, glm::vec3& com, glm::mat3& invInertiaTensor) {
    com = glm::vec3(0);
    float totalVolume = 0;
    for (const Collider& c : cc.colliders) {
        if (c.isTrigger) continue;
        if (c.geometry.type == CONVEX_MESH) {
            glm::vec3 center = paddedCenter(c.geometry.convex);
            forEachTet(c.geometry.convex, center, [&](glm::vec3 v0, glm::vec3 v1, glm::vec3 v2, float volume) {
                glm::vec3 centroid = (center + v0 + v1 + v2) / 4.f;
                com += (c.position + centroid) * volume;
                totalVolume += volume;
            });
        } else if (c.geometry.type != TRIANGLE_MESH) {
            float volume = colliderVolume(c);
            com += c.position * volume;
            totalVolume += volume;
        }
    }
    com /= totalVolume;

    glm::mat3 inertia(0);
    for (const Collider& c : cc.colliders) {
        if (c.isTrigger) continue;
        switch (c.geometry.type) {
            case SPHERE: {
                float m = mass * colliderVolume(c) / totalVolume;
                inertia += getInertiaSphere(m, c.geometry.sphere.radius) + parallelAxis(m, com - c.position);
            } break;
            case CAPSULE: {
                float m = mass * colliderVolume(c) / totalVolume;
                glm::mat3 rot = glm::toMat3(c.orientation);
                inertia += rot * getInertiaCapsule(m, c.geometry.capsule.halfHeight, c.geometry.capsule.radius) * glm::transpose(rot) + parallelAxis(m, com - c.position);
            } break;
            case BOX: {
                float m = mass * colliderVolume(c) / totalVolume;
                glm::mat3 rot = glm::toMat3(c.orientation);
                inertia += rot * getInertiaBox(m, c.geometry.box.halfExtents) * glm::transpose(rot) + parallelAxis(m, com - c.position);
            } break;
            case CONVEX_MESH: {
                glm::vec3 center = paddedCenter(c.geometry.convex);
                glm::vec3 d = c.position - com;
                forEachTet(c.geometry.convex, center, [&](glm::vec3 v0, glm::vec3 v1, glm::vec3 v2, float volume) {
                    float m = mass * volume / totalVolume;
                    inertia += getInertiaTetrahedron(m, { center + d, v0 + d, v1 + d, v2 + d });
                });
            } break;
            default: break;
        }
    }
    invInertiaTensor = glm::inverse(inertia);
}

void setMassProps(RigidBodyDynamicComponent& d, const RigidBodyCollisionComponent& cc, float mass) {
    d.invMass = 1.f / mass;
    computeCOMAndInvInertiaTensor(cc, mass, d.com, d.invInertiaTensor);
}

} // namespace physecs

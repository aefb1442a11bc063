// physecs_b200 host layer -- public value types of the Physecs API (source-compatible with the reference).
//
// Applications written against thfProjects/Physecs keep compiling: the names, namespaces, member names and aggregate
// layouts below are the ones the reference declares in
//   include/Physecs/Colliders.h:9-58      GeometryType, *Geometry, Geometry (tagged union), Material, Collider
//   include/Physecs/Components.h:6-17     RigidBodyCollisionComponent, RigidBodyDynamicComponent
//   include/Physecs/Bounds.h              Bounds
//   include/Physecs/ConvexMesh.h:7-34     ConvexMeshFace, ConvexMeshVertices<N>, ConvexMesh
//   include/Physecs/TriangleMesh.h:9-41   Triangle, TriangleMeshBVHNode, TriangleMesh
//   src/Transform.h:6-10                  TransformComponent (global; EnTT keys storage by type name, so name + layout matter)
// Only the data model lives on the host.  Everything Scene::simulate computes runs on the GPU behind include/physecs_b200.h.
#pragma once
#include <vector>
#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>

#ifndef PHYSECS_API
#define PHYSECS_API __attribute__((visibility("default")))
#endif

// The application may already define the identical struct (the reference's demo does, demo/Components.h:3-7).
#ifndef PHYSECS_TRANSFORM_COMPONENT_DEFINED
#define PHYSECS_TRANSFORM_COMPONENT_DEFINED
struct TransformComponent {
    glm::vec3 position;
    glm::quat orientation;
    glm::vec3 scale;
};
#endif

namespace physecs {

struct Bounds {
    glm::vec3 min, max;
    glm::vec3 getCenter() const { return (min + max) * 0.5f; }
    glm::vec3 getHalfExtents() const { return (max - min) * 0.5f; }
    void addMargin(const glm::vec3& m) { min -= m; max += m; }
    void expand(glm::vec3 e) {      // grows the box on the side the expansion points to (reference src/Bounds.cpp:12-21)
        if (e.x < 0) min.x += e.x; else max.x += e.x;
        if (e.y < 0) min.y += e.y; else max.y += e.y;
        if (e.z < 0) min.z += e.z; else max.z += e.z;
    }
    float area() const { const glm::vec3 d = max - min; return 2 * (d.x * d.y + d.y * d.z + d.z * d.x); }
};

// ---- meshes (user-owned, must outlive the Scene: Colliders.h:24-31 holds raw pointers) ---------------------------------
struct ConvexMeshFace {
    std::vector<int> indices;   // CCW loop seen from outside
    glm::vec3 normal;
    glm::vec3 centroid;
};

// vertex array padded to a multiple of N by repeating the last vertex (reference src/ConvexMesh.cpp:4-10); size() is the
// unpadded count.  The device support function scans the padded array four lanes at a time like the reference (quirk Q13).
template <int N>
class ConvexMeshVertices {
    std::vector<glm::vec3> buffer;
    int count = 0;

public:
    ConvexMeshVertices(const std::vector<glm::vec3>&& vertices) : buffer(vertices), count((int)vertices.size()) {
        if (count) buffer.resize((size_t)(count + N - 1) / N * N, buffer.back());
    }
    const glm::vec3& operator[](int i) const { return buffer[i]; }
    int size() const { return count; }
    std::vector<glm::vec3>::const_iterator begin() const { return buffer.begin(); }
    std::vector<glm::vec3>::const_iterator end() const { return buffer.end(); }
};

struct ConvexMesh {
    ConvexMeshVertices<4> vertices;
    std::vector<ConvexMeshFace> faces;
    ConvexMesh(const std::vector<glm::vec3>&& v, const std::vector<ConvexMeshFace>&& f) : vertices(std::move(v)), faces(f) {}
};

struct Triangle {
    unsigned int indices[3];
    Bounds bounds;
    glm::vec3 normal;
    glm::vec3 centroid;
};

struct TriangleMeshBVHNode {
    Bounds bounds;
    int triCount = 0;   // > 0: leaf over triangles [index, index + triCount); 0: children are nodes index, index + 1
    int index = 0;
};

// Static triangle mesh.  The constructor runs the registration-time binned-SAH build (pb_build_trimesh; identical tree and
// triangle order to the reference's TriangleMesh ctor, src/TriangleMesh.cpp:99-164) so `triangles` / `bvh` read like the
// reference's members and contact-cache triangle indices agree.
class PHYSECS_API TriangleMesh {
    std::vector<unsigned int> sourceIndices;
    std::vector<int> overlapScratch;

public:
    std::vector<glm::vec3> vertices;
    std::vector<Triangle> triangles;        // post-build order
    std::vector<TriangleMeshBVHNode> bvh;
    const int rootId = 0;

    TriangleMesh(const std::vector<glm::vec3>& vertices, const std::vector<unsigned int>& indices);
    const std::vector<unsigned int>& getSourceIndices() const { return sourceIndices; }
    // triangles whose bounds overlap `bounds` (host-side helper, reference TriangleMesh::overlapBvh :166-192)
    const std::vector<int>& overlapBvh(const Bounds& bounds);
};

// ---- colliders ------------------------------------------------------------------------------------------------------------
enum GeometryType { SPHERE, CAPSULE, BOX, CONVEX_MESH, TRIANGLE_MESH };

struct SphereGeometry { float radius; };
struct CapsuleGeometry { float halfHeight; float radius; };          // axis = local +Y
struct BoxGeometry { glm::vec3 halfExtents; };
struct ConvexMeshGeometry { ConvexMesh* mesh; glm::vec3 scale; };
struct TriangleMeshGeometry { TriangleMesh* mesh; };

struct Geometry {
    GeometryType type;
    union {
        SphereGeometry sphere;
        CapsuleGeometry capsule;
        BoxGeometry box;
        ConvexMeshGeometry convex;
        TriangleMeshGeometry triangleMesh;
    };
};

struct Material {
    float friction;
    float restitution;   // read as spring frequency when damping != 0 (soft contact, reference Physecs.cpp:246-262)
    float damping;
};

struct Collider {
    glm::vec3 position;      // local to the body
    glm::quat orientation;
    Geometry geometry;
    Material material;
    bool isTrigger;
    bool enableSimulation;
    int data;                // user tag handed to the contact filter
};

// ---- components -----------------------------------------------------------------------------------------------------------
struct RigidBodyCollisionComponent {
    std::vector<Collider> colliders;
};

struct RigidBodyDynamicComponent {
    bool isKinematic;
    glm::vec3 velocity;
    glm::vec3 angularVelocity;
    float invMass;
    glm::vec3 com;
    glm::mat3 invInertiaTensor;
};

} // namespace physecs

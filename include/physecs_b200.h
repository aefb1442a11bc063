/* physecs_b200 -- C ABI of the B200 (sm_100a) implementation of Physecs' per-step pipeline.
 *
 * This is the drop-in boundary: the host-side physecs::Scene mirror (physecs_b200/host/) binds to
 * exactly these entry points, and they are what a maintainer of the reference would call from
 * physecs::Scene::simulate (reference src/Physecs.cpp:112-561) instead of the CPU stages.
 * Plain pointers and sizes only; caller-owned host buffers (pinned recommended, see pb_host_alloc);
 * device memory is owned by the context.  Every function returns 0 on success or a PB_E* code and
 * never falls back to a CPU path: without a CUDA device pb_ctx_create fails.
 *
 * Index spaces
 *   body row   : rows [0, n_dynamic) are the packed RigidBodyDynamicComponent storage order
 *                (the reference's b0/b1 index space, src/Physecs.cpp:116-117, :271-272);
 *                rows [n_dynamic, n_dynamic+n_static) are entities with colliders but no dynamic component.
 *   collider   : one row per (entity, colliderIndex) -- the reference's BroadPhaseEntry (Physecs.h:80-88).
 *   quaternions: x,y,z,w in memory (glm::quat layout, reference src/Transform.h:6-10).
 *   mat3       : column-major 9 floats (glm::mat3).
 */
#ifndef PHYSECS_B200_H
#define PHYSECS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

enum {
    PB_OK = 0,
    PB_ECUDA = 1,        /* CUDA runtime error (message in pb_last_error) */
    PB_ECAPACITY = 2,    /* a per-step ARENA overflowed (candidate pairs / manifolds / trigger pairs): nothing persistent was touched, grow
                            (pb_grow_arenas) and run the step again.  Per-pair work (EPA polytope, clip polygons, triangles met by one
                            shape) is unbounded like the reference's std::vectors: pairs that outgrow the per-thread buffers are redone
                            by spill kernels on global-memory scratch and never fail a step (pb_counts.cause tells when that happened) */
    PB_EINVAL = 3,       /* bad argument */
    PB_EUNSUPPORTED = 4  /* a shape pair the reference itself has no routine for (mesh vs mesh) reached the narrowphase */
};

/* geometry types == physecs::GeometryType (reference include/Physecs/Colliders.h:9) */
enum { PB_SPHERE = 0, PB_CAPSULE = 1, PB_BOX = 2, PB_CONVEX_MESH = 3, PB_TRIANGLE_MESH = 4 };
/* joint types (reference include/Physecs/Joints/<Type>Joint.h) */
enum { PB_JOINT_FIXED = 0, PB_JOINT_REVOLUTE = 1, PB_JOINT_SPHERICAL = 2, PB_JOINT_UNIVERSAL = 3,
       PB_JOINT_PRISMATIC = 4, PB_JOINT_GEAR = 5, PB_JOINT_SERVO = 6 };
/* collider flag bits */
enum { PB_COL_TRIGGER = 1, PB_COL_ENABLE_SIM = 2 };

typedef struct pb_caps {
    int max_bodies;      /* dynamic + static rows */
    int max_colliders;
    int max_pairs;       /* candidate pairs per step (potentialContacts) */
    int max_manifolds;   /* contact manifolds per step */
    int max_joints;
    int reserved[3];
} pb_caps;

/* pb_counts.cause bits: which limit the last step (or query) ran into.  PAIRS / MANIFOLDS / TRIGGERS come with PB_ECAPACITY; the
 * SPILLED_* bits are informational (those pairs took the spill path, results are complete); SPILL_LIST / SPILL_SCRATCH mean even the
 * spill path's (much larger) bounds were exceeded: the step completed, the affected pairs carry the contacts that fit. */
enum {
    PB_CAUSE_PAIRS = 0x1, PB_CAUSE_MANIFOLDS = 0x2, PB_CAUSE_TRIGGERS = 0x4, PB_CAUSE_WALK_STACK = 0x8,
    PB_CAUSE_SPILLED_EPA_FACES = 0x10, PB_CAUSE_SPILLED_EPA_LOOSE = 0x20, PB_CAUSE_SPILLED_EPA_VERTS = 0x40, PB_CAUSE_SPILLED_CLIP = 0x80,
    PB_CAUSE_SPILLED_TRI_CAND = 0x100, PB_CAUSE_SPILLED_TRI_CONTACTS = 0x200, PB_CAUSE_SPILLED_MESH_STACK = 0x400,
    PB_CAUSE_SPILL_LIST = 0x1000, PB_CAUSE_SPILL_SCRATCH = 0x2000
};

/* per-step counters (pb_get_counts) */
typedef struct pb_counts {
    int n_pairs;         /* broadphase candidate pairs (== reference potentialContacts.size()) */
    int n_manifolds;     /* contact manifolds with >=1 point (== contactConstraints.size()) */
    int n_points;        /* contact points */
    int n_colors;        /* contact colours used this step */
    int n_overflow;      /* manifolds in the sequential overflow bucket (colour 63) */
    int status;          /* PB_OK or PB_ECAPACITY */
    int n_mesh_pairs;
    int n_triggers;      /* overlapping trigger pairs (== triggerCacheTemp.size()) */
    int cause;           /* PB_CAUSE_* bits of the last step */
    int n_spilled;       /* pairs the spill kernels redid in the last step (GJK / EPA bin + mesh bins) */
} pb_counts;

/* device times of the last pb_step in milliseconds (CUDA events on the context's stream) */
typedef struct pb_timings {
    float broadphase, narrowphase, contact_build, solve, total;
    float solve_kernel;  /* the persistent substep kernel alone (k_substeps), CUDA events around its launch */
    float reserved[2];
} pb_timings;

/* ---- context ------------------------------------------------------------------------------------ */
int  pb_ctx_create(int device, const pb_caps* caps, pb_ctx** out);
void pb_ctx_destroy(pb_ctx* ctx);
/* enlarge the per-step arenas of a live context (after pb_step returned PB_ECAPACITY): scene, bounds and the contact
 * cache of the previous step are kept, so the failed step can be run again */
int  pb_grow_arenas(pb_ctx* ctx, int max_pairs, int max_manifolds);
const char* pb_last_error(pb_ctx* ctx);
int  pb_host_alloc(void** ptr, unsigned long long bytes);   /* pinned host staging */
void pb_host_free(void* ptr);
void* pb_stream(pb_ctx* ctx);                                /* cudaStream_t the context launches on */

/* ---- scene description (replaces the reference's registry walks: Physecs.cpp:25-98 signal hooks) - */
/* entity[n_dynamic+n_static]: entt::entity integer of each row (pair ordering + same-entity test,
 * Physecs.cpp:145,:158).  kinematic/vel/angvel/inv_mass/com/inv_inertia: n_dynamic rows
 * (RigidBodyDynamicComponent, reference include/Physecs/Components.h:10-17). */
int pb_upload_bodies(pb_ctx* ctx, int n_dynamic, int n_static, const int* entity, const float* pos3,
                     const float* quat4, const int* kinematic, const float* vel3, const float* angvel3,
                     const float* inv_mass, const float* com3, const float* inv_inertia9);
/* One row per collider (reference Collider, include/Physecs/Colliders.h:50-58). params4: sphere {r},
 * capsule {halfHeight, r}, box {hx,hy,hz}, convex {sx,sy,sz}; mesh: handle from pb_register_*.
 * material3 = {friction, restitution, damping}. Creation-time bounds carry no margin (Physecs.cpp:32). */
int pb_upload_colliders(pb_ctx* ctx, int n, const int* body_row, const int* collider_index,
                        const float* local_pos3, const float* local_quat4, const int* type,
                        const float* params4, const int* mesh, const float* material3, const int* flags,
                        const int* data);
/* Convex mesh (reference ConvexMesh, include/Physecs/ConvexMesh.h:29-34): faces as index loops. */
int pb_register_convex(pb_ctx* ctx, const float* verts3, int n_verts, const int* face_offsets,
                       const int* face_indices, int n_faces, const float* face_normals3,
                       const float* face_centroids3, int* handle);
/* Static triangle mesh.  Builds, on the host at registration time, the same binned-SAH BVH the reference
 * TriangleMesh constructor builds (src/TriangleMesh.cpp:99-164) so post-build triangle indices (part of the
 * contact-cache key) agree; tri_order_out (optional, n_tris ints) receives original index of each built triangle. */
int pb_register_trimesh(pb_ctx* ctx, const float* verts3, int n_verts, const unsigned* indices, int n_indices,
                        int* handle, int* tri_order_out);
/* The host-side BVH build alone (no device needed): outputs sized n_tris*3 (tri_idx), n_tris (tri_orig),
 * 2*n_tris*6 (node_bounds), 2*n_tris*2 (node_count_index); *n_nodes receives the node count. */
int pb_build_trimesh(const float* verts3, int n_verts, const unsigned* indices, int n_indices, unsigned* tri_idx,
                     int* tri_orig, float* node_bounds6, int* node_count_index2, int* n_nodes);
/* joints: params per type -- revolute {driveEnabled, driveVelocity, driveMaxTorque}; prismatic {upper, lower,
 * driveEnabled, targetPosition, stiffness, damping}; gear {ratio}; servo {targetAngle, stiffness, damping}.
 * color[n]: colour assigned by the host-side greedy colouring (reference Physecs.cpp:690-710); 0..7 are solved one
 * thread per joint with the SIMD-path semantics (Constraint1DW.cpp), 8 = the overflow bucket, solved sequentially
 * with the scalar-path semantics (Constraint1D.cpp, quirk Q9). */
int pb_upload_joints(pb_ctx* ctx, int n, const int* type, const int* body_row0, const int* body_row1,
                     const float* anchor0_pos3, const float* anchor0_quat4, const float* anchor1_pos3,
                     const float* anchor1_quat4, const float* params8, const int* color);
/* new parameters of the uploaded joints after a setter call on a live joint (reference Joints/<Type>Joint.cpp setters); same
 * layout and joint order as pb_upload_joints; accumulated impulses and gear angle state are kept */
int pb_update_joint_params(pb_ctx* ctx, int n, const float* params8);
/* after a re-upload: old_index[j] = position joint j had in the previous pb_upload_joints call (-1 = new joint);
 * surviving joints keep their persistent state (GearJoint angle tracking, GearJoint.cpp:32-44) */
int pb_keep_joint_state(pb_ctx* ctx, int n, const int* old_index);
/* entity pairs that must not collide (reference nonCollidingPairs, Physecs.cpp:209,:694,:788): rows (e0<e1) */
int pb_set_noncolliding_pairs(pb_ctx* ctx, int n, const int* entity_pairs2);
/* after pb_upload_colliders: old_to_new[i] = new index of the collider that was row i before the upload (-1 = removed).
 * Re-keys the contact cache of the last step (reference contactCache, Physecs.cpp:237, :291-300, :553) so restitution
 * targets of persisting contacts survive adding / removing bodies, as they do in the reference */
int pb_keep_contact_cache(pb_ctx* ctx, int n_old, const int* old_to_new);
/* Contact filter (reference Scene::setContactFilter, Physecs.cpp:200, :795): the user's function
 * ContactType f(bool isTrigger0, int data0, bool isTrigger1, int data1) depends only on the (isTrigger, data) class of
 * each collider, so the host tabulates it: collider_class[n_colliders] in [0, n_classes), lut[c0 * n_classes + c1] = 1
 * when f answers TRIGGER for (class c0 = side 0 = lower entity, class c1).  n_classes = 0 restores defaultContactFilter
 * (Physecs.cpp:20-23).  Must be called again after pb_upload_colliders. */
int pb_set_contact_filter(pb_ctx* ctx, int n_colliders, const int* collider_class, int n_classes, const unsigned char* lut);
/* isKinematic of every dynamic row changed (Scene::setIsKinematic, Physecs.cpp:753-770): solver flags and the
 * BroadPhaseEntry::isDynamic bit of the owning colliders follow */
int pb_set_kinematic(pb_ctx* ctx, int n_dynamic, const int* kinematic);
/* mass properties of the dynamic rows changed in the registry (the reference reads them live every step) */
int pb_set_mass(pb_ctx* ctx, int n_dynamic, const float* inv_mass, const float* com3, const float* inv_inertia9);
/* overwrite persistent collider bounds (BroadPhaseEntry::bounds): lets the host carry bounds across a re-upload so the
 * creation-without-margin / margin-after-update history of surviving colliders is kept (quirk Q6) */
int pb_set_bounds(pb_ctx* ctx, int n, const int* colliders, const float* bounds6);
/* the same carry-over on the device: pb_keep_bounds_begin BEFORE pb_upload_colliders keeps a copy of the current bounds,
 * pb_keep_bounds AFTER it gives collider old_to_new[i] the bounds collider i had (-1 = none: removed, or a collider the
 * reference would have created anew, Physecs.cpp:20-33, :738-751); n_old = the collider count at pb_keep_bounds_begin */
int pb_keep_bounds_begin(pb_ctx* ctx);
int pb_keep_bounds(pb_ctx* ctx, int n_old, const int* old_to_new);

/* ---- per-step state exchange ---------------------------------------------------------------------- */
/* push registry state of the dynamic rows (what the reference reads through registry.get each step) */
int pb_set_state(pb_ctx* ctx, int n_dynamic, const float* pos3, const float* quat4, const float* vel3,
                 const float* angvel3);
/* the same for rows [first, first + count) of the dynamic storage; the pointers address the first row of the range.  Lets a host
 * that gathers its registry in chunks start each chunk's upload as soon as it is gathered. */
int pb_set_state_rows(pb_ctx* ctx, int first, int count, const float* pos3, const float* quat4, const float* vel3,
                      const float* angvel3);
/* transforms of the static rows [n_dynamic, n_dynamic + n_static): the reference reads every Transform live in the
 * narrowphase (Physecs.cpp:191-198) while bounds only follow registry.patch (pb_move_rows) */
int pb_set_static_poses(pb_ctx* ctx, int n_static, const float* pos3, const float* quat4);
/* a static/kinematic row was moved through registry.patch<TransformComponent> (Physecs.cpp:51-54,:79-90):
 * new transform + bounds refresh with the 0.01 margin */
int pb_move_rows(pb_ctx* ctx, int n, const int* rows, const float* pos3, const float* quat4);
/* recompute bounds (+0.01 margin) of every collider of a non-kinematic dynamic body, as the end of
 * simulate does (Physecs.cpp:556-559); used after pb_set_state when the caller wants bounds to follow */
int pb_refresh_bounds(pb_ctx* ctx);
/* == physecs::Scene::simulate(timeStep) with the scene's knobs (Physecs.cpp:100-110, :112-561).
 * pb_step ENQUEUES the step on the context's stream and returns without waiting for the device.  The arena checks run on the device:
 * a step whose pair / manifold arena overflowed skips its solve, integration and bounds refresh, i.e. leaves the scene exactly as
 * it was.  The outcome is reported by the next call that synchronises with the step -- pb_collect_step, pb_sync, pb_get_state,
 * pb_get_counts, the taps, or the next pb_step (which then enqueues nothing) -- as PB_ECAPACITY: grow the arenas (pb_grow_arenas)
 * and call pb_step again.  A caller that wants the status of a step before doing anything else calls pb_collect_step. */
int pb_step(pb_ctx* ctx, float dt, int substeps, int iterations, float gravity);
/* Optional head of a step: counters reset + broadphase, which need nothing from the host -- the reference's sweep, too, runs on the
 * bounds the previous simulate left behind (Physecs.cpp:119-173, bounds refreshed at :556-559 / by registry.patch), not on this
 * step's transforms.  Call it first, then gather and upload the new state (pb_set_state / pb_set_state_rows) while the broadphase
 * runs, then pb_step, which continues behind it.  Calling pb_step alone is equivalent. */
int pb_step_begin(pb_ctx* ctx);
/* Optional middle of a step: world poses + narrowphase (behind pb_step_begin's broadphase; calls it if it was not).  They read the
 * new poses but no velocity, so a caller uploads the poses, calls this, uploads the velocities while the narrowphase runs, and then
 * calls pb_step, which continues with the contact build.  Poses uploaded after this call are not seen by the narrowphase already
 * enqueued; a scene edit in between (colliders, kinematic flags, moved rows, bounds) makes pb_step start the step over. */
int pb_step_narrowphase(pb_ctx* ctx);
/* waits for the narrowphase of the last pb_step (not for the whole step) and returns its status; PB_OK when nothing is pending */
int pb_collect_step(pb_ctx* ctx);
int pb_get_state(pb_ctx* ctx, float* pos3, float* quat4, float* vel3, float* angvel3);
/* pb_get_state in up to 32 chunks of rows: _begin enqueues the copies (asynchronous; reports a failed step like pb_get_state),
 * _wait blocks until chunk `chunk` has arrived and names its row range -- the host can scatter a chunk into its registry while the
 * next one is still on the bus.  chunk >= the number of chunks: *count = 0.
 * The state is snapshotted on the device when _begin is reached in stream order and copied out on a stream of its own: calls made
 * after _begin (pb_set_state, pb_step of the NEXT step) do not wait for the transfer and do not disturb it, so a caller that can use
 * step k's result one step late -- _begin(k), enqueue step k + 1, _wait(k) -- hides the read-back behind the device's work.
 * One read-back at a time: a second _begin waits (on the device) for the previous one's copies.  Host buffers: pinned, and left
 * alone until _wait returned for every chunk. */
int pb_get_state_begin(pb_ctx* ctx, float* pos3, float* quat4, float* vel3, float* angvel3, int n_chunks);
int pb_get_state_wait(pb_ctx* ctx, int chunk, int* first, int* count);
/* _wait_poses returns as soon as pos3 / quat4 of the chunk are in (pb_get_state_wait: all four arrays).  pb_set_readback_order(1):
 * later read-backs copy the poses of EVERY chunk out before any velocity (default 0: chunk by chunk, poses then velocities).  A loop
 * that sends the result back up for the next step uploads poses first -- the narrowphase waits for them -- and the velocities behind
 * them, which only the contact build reads. */
int pb_get_state_wait_poses(pb_ctx* ctx, int chunk, int* first, int* count);
int pb_set_readback_order(pb_ctx* ctx, int poses_first);
int pb_sync(pb_ctx* ctx);

/* ---- parity / debug taps ----------------------------------------------------------------------------- */
int pb_get_counts(pb_ctx* ctx, pb_counts* out);
int pb_get_timings(pb_ctx* ctx, pb_timings* out);
/* candidate pairs of the last step per narrowphase bin: sphere-sphere, sphere-capsule, capsule-capsule, sphere-box, capsule-box, box-box,
 * GJK/EPA (any pair with a convex mesh), mesh-sphere, mesh-capsule, mesh-box, mesh-convex, trigger (reference dispatch Collision.cpp:895-1016) */
int pb_get_bin_counts(pb_ctx* ctx, int* out12);
/* candidate pairs of the last step as rows (entity0, colIdx0, entity1, colIdx1), entity0 < entity1 */
int pb_get_pairs(pb_ctx* ctx, int* out4, int cap, int* n);
/* overlapping TRIGGER pairs of the last step (reference triggerCacheTemp, Physecs.cpp:200-207) as rows
 * (entity0, colIdx0, entity1, colIdx1), entity0 < entity1; the enter / exit diff and listener calls
 * (Physecs.cpp:538-552) are host work */
int pb_get_triggers(pb_ctx* ctx, int* out4, int cap, int* n);
/* collider bounds rows (min xyz, max xyz) in collider order */
int pb_get_bounds(pb_ctx* ctx, float* out6);
/* manifolds of the last step in SOLVE ORDER (colour-major): keys rows (entity0, colIdx0, entity1, colIdx1, tri),
 * num_points, normal3, points rows [4][2][3] world space at narrowphase time, color */
int pb_get_manifolds(pb_ctx* ctx, int cap, int* keys5, int* num_points, float* normal3, float* points24,
                     int* color, int* n);
/* the device key / payload sort that orders the broadphase leaves (stands where the reference keeps its entries ordered by
 * insertion sort, Physecs.cpp:121-133): sorts n host (key, payload) pairs by the low `bits` key bits (rounded up to whole 8-bit passes), stable, and returns them */
int pb_debug_sort_pairs(pb_ctx* ctx, int n, int bits, const unsigned int* keys_in, const int* vals_in, unsigned int* keys_out, int* vals_out);

/* ---- scene queries (reference Scene::raycastClosest / overlap, Physecs.cpp:571-650) -------------------------- */
/* Every collider hit by each ray within max_dist (leaf bounds test, then the geometry at its current pose; triangle meshes
 * have no ray routine in the reference).  Rows (ray, entity, colIdx, t), unordered; *n_hits may exceed cap (retry larger).
 * The caller picks the closest hit that passes its filter (the reference's filter is a host std::function). */
int pb_query_raycast(pb_ctx* ctx, int n_rays, const float* orig3, const float* dir3, float max_dist, int cap, int* out_ray,
                     int* out_entity, int* out_col_idx, float* out_t, int* n_hits);
/* Colliders overlapping a query shape (physecs::overlap, query shape first); filter != 0 keeps colliders whose data & filter
 * is non-zero.  mesh: convex handle for a convex query shape. */
int pb_query_overlap(pb_ctx* ctx, const float* pos3, const float* quat4, int type, const float* params4, int mesh, int filter,
                     int cap, int* out_entity, int* out_col_idx, int* n_hits);

/* Scene::overlapWithMinTranslationalDistance (Physecs.cpp:652-688): physecs::collision(collider, query shape) for every collider
 * whose bounds meet the query bounds; one row per manifold with points (triangle-mesh colliders give one per touched triangle):
 * entity, colIdx, normal (collider -> query shape), mtd = deepest penetration along the normal (>= 0).  Rows are unordered.
 * When *n_hits > cap nothing was written: call again with cap >= *n_hits. */
int pb_query_overlap_mtd(pb_ctx* ctx, const float* pos3, const float* quat4, int type, const float* params4, int mesh, int cap,
                         int* out_entity, int* out_col_idx, float* out_normal3, float* out_mtd, int* n_hits);

/* The device tree the queries (and the broadphase) walk, for debug drawing (reference Scene::getBVH / getBHVRootId,
 * Physecs.cpp:807-813).  Internal nodes 0 .. *n_internal-1, node 0 is the root; per node the boxes of its two children
 * (min xyz, max xyz, left then right) and two links: >= 0 an internal node, < 0 the collider -1 - link (pb_collider_ids names it).
 * Fewer than two colliders: *n_internal = 0.  cap < *n_internal: nothing written. */
int pb_get_tree(pb_ctx* ctx, int cap, float* child_boxes12, int* child_links2, int* n_internal);
int pb_collider_ids(pb_ctx* ctx, int n, const int* cols, int* out_entity, int* out_col_idx);

/* Simulation islands.  Connected components of the body / constraint graph that are small enough are solved inside one CTA each
 * (no device-wide barrier per colour): what makes batches of independent little scenes (config 5) fast.  Results do not depend on
 * the mode.  0 = off, 1 = on, 2 = auto (default: on while at least half of the constraints sit in small islands).
 * pb_get_island_stats: {on in the last step, constraints in small islands, constraints in all islands} of the last step that looked. */
int pb_set_islands(pb_ctx* ctx, int mode);
/* Which broadphase the last enqueued step took: out3 = {1 = all pairs through shared-memory tiles / 0 = tree, (group of 32
 * consecutive colliders, tile of 128) pairs whose boxes met in the last step that looked (-1: none yet), tiles}.  Up to 8192 colliders the step always tests all pairs; up to 131 072 it does when
 * colliders that were created together lie together (a batch of scenes side by side: every tile of 128 consecutive colliders meets a
 * handful of tiles), judged from the previous step's tile statistics; otherwise the LBVH.  Same pair set either way. */
int pb_get_broadphase_info(pb_ctx* ctx, int* out3);
int pb_get_island_stats(pb_ctx* ctx, int* out3);

/* Deterministic mode (also env PB_DETERMINISTIC=1 at pb_ctx_create).  The reference is reproducible with numThreads = 0
 * (src/ThreadPool.cpp:30-46: everything on the caller); the device's default colouring claims colours with atomics, so the solve order
 * and with it the trajectory differ from run to run (every run is a valid colour-batched order, which is what the parity gates
 * compare through).  With this switch the colours come from fixed priorities (a hash of each manifold's collider pair + triangle)
 * and the sequential bucket is ordered by the same hash: two runs of the same scene then produce bit-identical states.  Costs a few
 * colouring rounds per step; meant for debugging and bisecting. */
int pb_set_deterministic(pb_ctx* ctx, int on);

/* optional in-kernel phase profile of the persistent substep kernel (bench.py roofline), accumulated since pb_set_profile(ctx, 1):
 * kinds 0 = integrate velocities, 1 = joint NGS phases, 2 = contact colour phases of the device-wide sweep, 3 = joint solve phases,
 * 4 = integrate positions, 5 = the per-CTA island sweeps (all their contact and joint colours; islands on);
 * slot 6 = the k_substep_solve launches of the LAST step timed by CUDA events on the context's stream (ms summed, number of launches). */
int pb_set_profile(pb_ctx* ctx, int on);
int pb_get_profile(pb_ctx* ctx, double* ms8, long long* count8);
/* the contact-pass share of it per solver colour: accumulated ms and number of phases for colours 0..63 */
int pb_get_profile_colors(pb_ctx* ctx, double* ms64, long long* count64);
/* kernels launched by this context since creation */
unsigned long long pb_get_launches(pb_ctx* ctx);
/* cudaProfilerStart/Stop, so `ncu --profile-from-start off` captures only the timed region of bench.py */
void pb_profiler_range(int start);

/* ---- batches of independent scenes over several devices (BASELINE.json configs[4]) ------------------------
 * The reference has one Scene driven by one thread (src/Physecs.cpp:112); an application simulating many independent worlds runs
 * many Scenes.  A pb_batch shards such a set by scene: shard k = one context on devices[k] holding a contiguous block of scenes
 * concatenated into one arena (upload it through pb_batch_ctx(batch, k) with the ordinary pb_upload_* calls: scenes that do not
 * touch produce no cross-scene pairs, and simulation islands keep them apart in the solver), driven by its own host thread on its
 * own stream.  No collective and no peer access on the data path.  All calls are asynchronous except sync / get_state. */
typedef struct pb_batch pb_batch;
/* contiguous block [begin, end) of scene indices owned by `shard`; block sizes differ by at most one */
int  pb_batch_shard_range(int n_scenes, int n_shards, int shard, int* begin, int* end);
int  pb_batch_create(int n_shards, const int* devices, const pb_caps* caps_per_shard, pb_batch** out);
void pb_batch_destroy(pb_batch* batch);
int  pb_batch_shards(pb_batch* batch);
pb_ctx* pb_batch_ctx(pb_batch* batch, int shard);
/* every shard's thread enqueues n_steps x pb_step; returns at once */
int  pb_batch_step(pb_batch* batch, int n_steps, float dt, int substeps, int iterations, float gravity);
/* waits for every shard; first non-OK status of any call since the last sync (pb_batch_last_error names shard and cause) */
int  pb_batch_sync(pb_batch* batch);
/* per-shard host arrays (array of n_shards pointers each, any of them null = not transferred); n_dynamic[k] rows in shard k */
int  pb_batch_set_state(pb_batch* batch, const float* const* pos3, const float* const* quat4, const float* const* vel3,
                        const float* const* angvel3, const int* n_dynamic);
int  pb_batch_get_state(pb_batch* batch, float* const* pos3, float* const* quat4, float* const* vel3, float* const* angvel3);
const char* pb_batch_last_error(pb_batch* batch);

#ifdef __cplusplus
}
#endif
#endif

// Sanitizer run of the host layer's structural-edit path over the recording double (test infrastructure, see pb_recorder.cpp):
// a scene large enough for the worker threads to take part (parallelFor goes parallel from 4096 items), then destroy / spawn / collider
// edits / a reordered pool, a simulate() after each.  Built and run by tests/abi_recorder/sanitize.sh with -fsanitize=thread and
// -fsanitize=address,undefined.
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {
void* psh_create(int numThreads, int device);
void psh_destroy(void* h);
const char* psh_last_error(void* h);
int psh_add_entities(void* hp, int n, const float* pos, const float* quat, const int* flags, const float* vel, const float* angvel, const float* invMass,
                     const float* com, const float* invI, const int* colOffsets, const float* colLPos, const float* colLQuat, const int* colType,
                     const float* colParams, const int* colMesh, const float* colMaterial, const int* colFlags, const int* colData);
int psh_destroy_entity(void* hp, int e);
int psh_add_collider(void* hp, int e, const float* lpos, const float* lquat, int type, const float* params, int mesh, const float* material, int flags, int data);
int psh_clear_colliders(void* hp, int e);
void psh_sort_dynamic(void* hp, int greaterFirst);
void psh_set_sync_mode(void* hp, int deviceAuthoritative);
double psh_simulate(void* hp, float dt);
void psh_get_state(void* hp, float* pos, float* quat, float* vel, float* angvel);
int psh_num_entities(void* hp);
}

static int add(void* h, int n, int dynamicFlag, float x0) {
    std::vector<float> pos(3 * n), quat(4 * n, 0.f), vel(3 * n, 0.f), ang(3 * n, 0.f), invMass(n, 1.f), com(3 * n, 0.f), invI(9 * n, 0.f);
    std::vector<int> flags(n, 1 | dynamicFlag), off(n + 1), type(n, 0), mesh(n, -1), cflags(n, 2), data(n, 0);
    std::vector<float> lpos(3 * n, 0.f), lquat(4 * n, 0.f), prm(4 * n, 0.f), mat(3 * n, 0.f);
    for (int i = 0; i < n; ++i) {
        pos[3 * i] = x0 + i; pos[3 * i + 1] = 1.f; pos[3 * i + 2] = 0.f; quat[4 * i + 3] = 1.f; lquat[4 * i + 3] = 1.f; prm[4 * i] = 0.3f;
        invI[9 * i] = invI[9 * i + 4] = invI[9 * i + 8] = 1.f; off[i] = i;
    }
    off[n] = n;
    return psh_add_entities(h, n, pos.data(), quat.data(), flags.data(), vel.data(), ang.data(), invMass.data(), com.data(), invI.data(), off.data(), lpos.data(),
                            lquat.data(), type.data(), prm.data(), mesh.data(), mat.data(), cflags.data(), data.data());
}

static void step(void* h, const char* what) {
    if (psh_simulate(h, 1.f / 60.f) < 0) { std::fprintf(stderr, "simulate failed after %s: %s\n", what, psh_last_error(h)); std::exit(1); }
}

int main(int argc, char** argv) {
    void* h = psh_create(6, 0);
    const int N = argc > 1 ? std::atoi(argv[1]) : 30000;      // from 65 536 bodies the gather / upload / read-back runs in 8 chunks, poses before velocities
    add(h, 50, 0, -1000.f);             // statics
    int first = add(h, N, 2, 0.f);      // dynamics
    step(h, "build");
    step(h, "steady");
    for (int e = first + 10; e < first + N; e += 997) psh_destroy_entity(h, e);
    step(h, "destroy");
    add(h, 500, 2, 50000.f);
    step(h, "spawn");
    const float lp[3] = { 0.4f, 0, 0 }, lq[4] = { 0, 0, 0, 1 }, prm[4] = { 0.2f, 0, 0, 0 }, mat[3] = { 0.4f, 0.3f, 0 };
    for (int e = first; e < first + N; e += 1009) if (e % 997 != (first + 10) % 997) psh_add_collider(h, e, lp, lq, 0, prm, -1, mat, 2, 0);
    psh_clear_colliders(h, first + 1); psh_add_collider(h, first + 1, lp, lq, 0, prm, -1, mat, 2, 0);
    step(h, "collider edits");
    psh_sort_dynamic(h, 0);
    step(h, "sort");
    psh_set_sync_mode(h, 1);
    step(h, "device-authoritative");
    add(h, 3, 2, 90000.f);
    step(h, "spawn in device-authoritative mode");
    const int n = psh_num_entities(h);
    std::vector<float> p(3 * n), q(4 * n), v(3 * n), w(3 * n);
    psh_get_state(h, p.data(), q.data(), v.data(), w.data());
    // a body that lived through all 8 steps moved by 8, whatever row it sat in
    const int e = first + 5;
    if (p[3 * e] < 5.f + 7.99f || p[3 * e] > 5.f + 8.01f) { std::fprintf(stderr, "entity %d at x = %f, expected 13\n", e, p[3 * e]); return 1; }
    psh_destroy(h);
    std::puts("sanitize driver ok");
    return 0;
}

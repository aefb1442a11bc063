// Sanitizer run of the batch driver (physecs_b200/csrc/batch.cpp: one host thread + task queue per shard) over the recording double:
// shards stepped concurrently, state pushed / fetched while other shards still work, a failing shard, destroy with work in flight.
#include "../../include/physecs_b200.h"
#include <cstdio>
#include <vector>

extern "C" void pbr_fail_next_steps(pb_ctx* c, int n, int needPairs, int needManifolds);

static void fill(pb_ctx* ctx, int n, float x0) {
    std::vector<int> ent(n), kin(n, 0), row(n), idx(n, 0), type(n, 0), mesh(n, -1), flags(n, 2), data(n, 0);
    std::vector<float> pos(3 * n, 0.f), quat(4 * n, 0.f), vel(3 * n, 0.f), ang(3 * n, 0.f), invMass(n, 1.f), com(3 * n, 0.f), invI(9 * n, 0.f), lp(3 * n, 0.f), lq(4 * n, 0.f), prm(4 * n, 0.3f), mat(3 * n, 0.f);
    for (int i = 0; i < n; ++i) { ent[i] = i; row[i] = i; pos[3 * i] = x0 + i; quat[4 * i + 3] = lq[4 * i + 3] = 1.f; }
    pb_upload_bodies(ctx, n, 0, ent.data(), pos.data(), quat.data(), kin.data(), vel.data(), ang.data(), invMass.data(), com.data(), invI.data());
    pb_upload_colliders(ctx, n, row.data(), idx.data(), lp.data(), lq.data(), type.data(), prm.data(), mesh.data(), mat.data(), flags.data(), data.data());
}

int main() {
    const int K = 6, N = 2000;
    std::vector<int> dev(K, 0);
    std::vector<pb_caps> caps(K);
    for (auto& c : caps) { c = pb_caps{}; c.max_bodies = N; c.max_colliders = N; c.max_pairs = 8 * N; c.max_manifolds = 8 * N; c.max_joints = 16; }
    pb_batch* b = nullptr;
    if (pb_batch_create(K, dev.data(), caps.data(), &b) != PB_OK) return 1;
    for (int k = 0; k < K; ++k) fill(pb_batch_ctx(b, k), N - k, 1000.f * k);
    std::vector<std::vector<float>> P(K), Q(K), V(K), W(K);
    std::vector<float*> p(K), q(K), v(K), w(K);
    std::vector<int> nd(K);
    for (int k = 0; k < K; ++k) { nd[k] = N - k; P[k].resize(3 * nd[k]); Q[k].resize(4 * nd[k]); V[k].resize(3 * nd[k]); W[k].resize(3 * nd[k]); p[k] = P[k].data(); q[k] = Q[k].data(); v[k] = V[k].data(); w[k] = W[k].data(); }
    for (int round = 0; round < 20; ++round) {
        if (pb_batch_step(b, 25, 1.f / 60.f, 4, 2, 9.81f) != PB_OK) return 2;
        if (pb_batch_get_state(b, p.data(), q.data(), v.data(), w.data()) != PB_OK) return 3;
        if (pb_batch_set_state(b, p.data(), q.data(), v.data(), w.data(), nd.data()) != PB_OK) return 4;
    }
    if (pb_batch_sync(b) != PB_OK) return 5;
    for (int k = 0; k < K; ++k) if (P[k][0] != 1000.f * k + 500.f) { std::fprintf(stderr, "shard %d: x = %f\n", k, P[k][0]); return 6; }
    // one shard fails: the status surfaces at the sync, names the shard, and the batch goes on
    pbr_fail_next_steps(pb_batch_ctx(b, 3), 1, -1, -1);
    pb_batch_step(b, 1, 1.f / 60.f, 4, 2, 9.81f);
    if (pb_batch_get_state(b, p.data(), q.data(), v.data(), w.data()) != PB_ECAPACITY) return 7;
    std::printf("failing shard reported as: %s\n", pb_batch_last_error(b));
    if (pb_batch_step(b, 3, 1.f / 60.f, 4, 2, 9.81f) != PB_OK || pb_batch_sync(b) != PB_OK) return 8;
    pb_batch_step(b, 200, 1.f / 60.f, 4, 2, 9.81f);      // destroyed with work in flight: the destructor drains
    pb_batch_destroy(b);
    std::puts("batch sanitize driver ok");
    return 0;
}

// Batches of independent scenes sharded over devices (SURVEY.md §8b "pb_batch_create / pb_batch_step", §8e; BASELINE.json configs[4]).
//
// Scenes of a batch never interact, so a batch shards by scene: shard k holds a contiguous block of scenes concatenated into ONE
// context on device k (one SoA arena; the broadphase finds no cross-scene pairs, simulation islands keep the scenes apart in the
// solver) and is driven by its own host thread on its own stream.  No collective, no peer access: the only cross-shard operation is
// the host-side join in pb_batch_sync.  The reference has no counterpart (one Scene, one thread: src/Physecs.cpp:112); an
// application that ran N reference Scenes in N threads maps to one pb_batch.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/physecs_b200.h"

namespace {
struct Shard {
    int device = 0;
    pb_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::function<int()>> queue;      // tasks, FIFO
    size_t done = 0, posted = 0;
    int status = PB_OK;                           // first non-OK task result since the last pb_batch_sync
    std::string err;
    bool quit = false;
};
}

struct pb_batch {
    std::vector<Shard*> shards;
    std::string err;
};

static void shardLoop(Shard* s) {
    cudaSetDevice(s->device);
    for (;;) {
        std::function<int()> task;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->quit || s->done < s->posted; });
            if (s->done >= s->posted) { if (s->quit) return; continue; }
            task = s->queue[s->done];
        }
        int rc = task();
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (rc != PB_OK && s->status == PB_OK) { s->status = rc; s->err = s->ctx ? pb_last_error(s->ctx) : "no context"; }
            ++s->done;
            if (s->done == s->posted) { s->queue.clear(); s->done = s->posted = 0; }
        }
        s->cv.notify_all();
    }
}

static void post(Shard* s, std::function<int()> f) {
    { std::lock_guard<std::mutex> lk(s->mu); s->queue.push_back(std::move(f)); ++s->posted; }
    s->cv.notify_all();
}
static void drain(Shard* s) {
    std::unique_lock<std::mutex> lk(s->mu);
    s->cv.wait(lk, [&] { return s->done == s->posted; });
}
// first failure over the shards since the last call (clears them)
static int harvest(pb_batch* b) {
    int rc = PB_OK;
    for (size_t k = 0; k < b->shards.size(); ++k) {
        Shard* s = b->shards[k];
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->status != PB_OK && rc == PB_OK) { rc = s->status; b->err = "shard " + std::to_string(k) + " (device " + std::to_string(s->device) + "): " + s->err; }
        s->status = PB_OK;
    }
    return rc;
}

extern "C" {

int pb_batch_shard_range(int n_scenes, int n_shards, int shard, int* begin, int* end) {
    if (n_shards < 1 || shard < 0 || shard >= n_shards || n_scenes < 0) return PB_EINVAL;
    const int base = n_scenes / n_shards, extra = n_scenes % n_shards;
    const int b = shard * base + (shard < extra ? shard : extra);
    if (begin) *begin = b;
    if (end) *end = b + base + (shard < extra ? 1 : 0);
    return PB_OK;
}

int pb_batch_create(int n_shards, const int* devices, const pb_caps* caps, pb_batch** out) {
    if (n_shards < 1 || !devices || !caps || !out) return PB_EINVAL;
    pb_batch* b = new pb_batch();
    for (int k = 0; k < n_shards; ++k) {
        Shard* s = new Shard();
        s->device = devices[k];
        int rc = pb_ctx_create(devices[k], &caps[k], &s->ctx);       // fails without a device: there is no CPU path
        if (rc != PB_OK) {
            delete s;
            for (Shard* t : b->shards) { { std::lock_guard<std::mutex> lk(t->mu); t->quit = true; } t->cv.notify_all(); t->th.join(); pb_ctx_destroy(t->ctx); delete t; }
            delete b;
            return rc;
        }
        s->th = std::thread(shardLoop, s);
        b->shards.push_back(s);
    }
    *out = b;
    return PB_OK;
}

void pb_batch_destroy(pb_batch* b) {
    if (!b) return;
    for (Shard* s : b->shards) {
        drain(s);
        { std::lock_guard<std::mutex> lk(s->mu); s->quit = true; }
        s->cv.notify_all();
        s->th.join();
        pb_ctx_destroy(s->ctx);
        delete s;
    }
    delete b;
}

int pb_batch_shards(pb_batch* b) { return b ? (int)b->shards.size() : 0; }
pb_ctx* pb_batch_ctx(pb_batch* b, int shard) { return (b && shard >= 0 && shard < (int)b->shards.size()) ? b->shards[shard]->ctx : nullptr; }
const char* pb_batch_last_error(pb_batch* b) { return b ? b->err.c_str() : "null batch"; }

int pb_batch_step(pb_batch* b, int n_steps, float dt, int substeps, int iterations, float gravity) {
    if (!b || n_steps < 0) return PB_EINVAL;
    for (Shard* s : b->shards) {
        pb_ctx* ctx = s->ctx;
        post(s, [=]() {
            for (int i = 0; i < n_steps; ++i) { int rc = pb_step(ctx, dt, substeps, iterations, gravity); if (rc) return rc; }
            return (int)PB_OK;
        });
    }
    return PB_OK;      // the shards' threads enqueue concurrently; failures surface in pb_batch_sync / pb_batch_get_state
}

int pb_batch_sync(pb_batch* b) {
    if (!b) return PB_EINVAL;
    for (Shard* s : b->shards) { pb_ctx* ctx = s->ctx; post(s, [=]() { return pb_sync(ctx); }); }
    for (Shard* s : b->shards) drain(s);
    return harvest(b);
}

int pb_batch_set_state(pb_batch* b, const float* const* pos3, const float* const* quat4, const float* const* vel3, const float* const* angvel3, const int* n_dynamic) {
    if (!b || !n_dynamic) return PB_EINVAL;
    for (size_t k = 0; k < b->shards.size(); ++k) {
        Shard* s = b->shards[k];
        pb_ctx* ctx = s->ctx;
        const float* p = pos3 ? pos3[k] : nullptr; const float* q = quat4 ? quat4[k] : nullptr;
        const float* v = vel3 ? vel3[k] : nullptr; const float* w = angvel3 ? angvel3[k] : nullptr;
        const int n = n_dynamic[k];
        post(s, [=]() { return pb_set_state(ctx, n, p, q, v, w); });
    }
    return PB_OK;
}

int pb_batch_get_state(pb_batch* b, float* const* pos3, float* const* quat4, float* const* vel3, float* const* angvel3) {
    if (!b) return PB_EINVAL;
    for (size_t k = 0; k < b->shards.size(); ++k) {
        Shard* s = b->shards[k];
        pb_ctx* ctx = s->ctx;
        float* p = pos3 ? pos3[k] : nullptr; float* q = quat4 ? quat4[k] : nullptr; float* v = vel3 ? vel3[k] : nullptr; float* w = angvel3 ? angvel3[k] : nullptr;
        post(s, [=]() { return pb_get_state(ctx, p, q, v, w); });
    }
    for (Shard* s : b->shards) drain(s);
    return harvest(b);
}

} // extern "C"

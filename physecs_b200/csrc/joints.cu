// Joints: host-side store / upload and the once-per-step kernel; the per-substep device code is in joints.cuh.
#include "joints.cuh"
#include <algorithm>
#include <vector>

// once per step (Physecs.cpp:322-354): reset accumulated impulses, prismatic limit selection (PrismaticJoint.cpp:116-155)
__global__ void k_joint_begin(JointDev J, int2* __restrict__ bodies, const int* __restrict__ kinematic, const float4* __restrict__ pos, const float4* __restrict__ quat) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J.n) return;
    {   // b0 / b1 of the step (Physecs.cpp:335-336): kinematic bodies and statics are index -1
        int2 rr = J.rows[j];
        bodies[j] = make_int2(jSolverIndex(rr.x, J.nDyn, kinematic), jSolverIndex(rr.y, J.nDyn, kinematic));
    }
    for (int r = 0; r < MAXR; ++r) J.lambda[r * J.n + j] = 0.f;
    if (J.type[j] == PB_JOINT_PRISMATIC) {
        int2 rr = J.rows[j];
        Q4 q0 = mkq(quat[rr.x]), q1 = mkq(quat[rr.y]);
        V3 p0 = mk3(pos[rr.x]) + rotate(q0, mk3(J.a0p[j]));
        V3 p1 = mk3(pos[rr.y]) + rotate(q1, mk3(J.a1p[j]));
        V3 d = p1 - p0;
        V3 u00 = rotate(qmul(q0, mkq(J.a0q[j])), mk3(1.f, 0.f, 0.f));
        float dx = dot(d, u00);
        float4 prm0 = J.prm[2 * j];
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dx > prm0.x) st.x = 1.f;
        else if (dx < prm0.y) st.y = 1.f;
        J.state[2 * j] = st;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct JointStore {
    int n = 0;
    int* type = nullptr; int2* rows = nullptr; int2* bodies = nullptr; float4* a0p = nullptr; float4* a0q = nullptr; float4* a1p = nullptr; float4* a1q = nullptr;
    float4* prm = nullptr; float4* state = nullptr;
    float4* linC = nullptr; float4* a0T = nullptr; float4* a1K = nullptr; float4* a0tMin = nullptr; float4* a1tMax = nullptr;
    float2* soft = nullptr; float* lambda = nullptr;
    std::vector<int> order;   // device slot -> caller's joint index
    // persistent state of the previous upload (gear angles), kept so surviving joints can carry it over
    std::vector<float4> prevState; std::vector<int> prevOrder;
};
static JointStore* store(pb_ctx* ctx) { return (JointStore*)ctx->jointStore; }

void pb_joints_free(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    if (!s) return;
    cudaFree(s->type); cudaFree(s->rows); cudaFree(s->bodies); cudaFree(s->a0p); cudaFree(s->a0q); cudaFree(s->a1p); cudaFree(s->a1q); cudaFree(s->prm); cudaFree(s->state);
    cudaFree(s->linC); cudaFree(s->a0T); cudaFree(s->a1K); cudaFree(s->a0tMin); cudaFree(s->a1tMax); cudaFree(s->soft); cudaFree(s->lambda);
    delete s;
    ctx->jointStore = nullptr;
    ctx->nJoints = 0;
}

static JointDev devView(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    JointDev J;
    J.n = s->n; J.nDyn = ctx->nDyn; J.type = s->type; J.rows = s->rows; J.bodies = s->bodies; J.a0p = s->a0p; J.a0q = s->a0q; J.a1p = s->a1p; J.a1q = s->a1q;
    J.prm = s->prm; J.state = s->state; J.linC = s->linC; J.a0T = s->a0T; J.a1K = s->a1K; J.a0tMin = s->a0tMin; J.a1tMax = s->a1tMax;
    J.soft = s->soft; J.lambda = s->lambda;
    return J;
}

int pb_joints_upload(pb_ctx* ctx, int n, const int* type, const int* row0, const int* row1, const float* a0p, const float* a0q,
                     const float* a1p, const float* a1q, const float* params8, const int* color) {
    std::vector<float4> prevState; std::vector<int> prevOrder;
    if (JointStore* old = store(ctx)) {
        cudaStreamSynchronize(ctx->stream);
        prevState.resize(2 * (size_t)old->n);
        cudaMemcpy(prevState.data(), old->state, sizeof(float4) * 2 * old->n, cudaMemcpyDeviceToHost);
        prevOrder = old->order;
    }
    pb_joints_free(ctx);
    if (n == 0) return PB_OK;
    for (int j = 0; j < n; ++j) {
        if (color[j] < 0 || color[j] > 8) return pb_fail(ctx, PB_EINVAL, "joint colour out of range");
        if (type[j] < 0 || type[j] > PB_JOINT_SERVO) return pb_fail(ctx, PB_EINVAL, "joint type");
    }
    JointStore* s = new JointStore();
    ctx->jointStore = s;
    s->n = n;
    s->prevState.swap(prevState); s->prevOrder.swap(prevOrder);
    // colour-major order, stable inside a colour (== the reference's per-colour joint vectors)
    s->order.resize(n);
    for (int j = 0; j < n; ++j) s->order[j] = j;
    std::stable_sort(s->order.begin(), s->order.end(), [&](int a, int b) { return color[a] < color[b]; });
    for (int c = 0; c <= PB_JOINT_COLORS; ++c) ctx->jointColorStart[c] = 0;
    for (int j = 0; j < n; ++j) ctx->jointColorStart[color[j] + 1]++;
    for (int c = 0; c < PB_JOINT_COLORS; ++c) ctx->jointColorStart[c + 1] += ctx->jointColorStart[c];
    std::vector<int> t(n); std::vector<int2> rr(n); std::vector<float4> p0(n), q0(n), p1(n), q1(n), prm(2 * (size_t)n), st(2 * (size_t)n, make_float4(0, 0, 0, 0));
    for (int k = 0; k < n; ++k) {
        int j = s->order[k];
        t[k] = type[j]; rr[k] = make_int2(row0[j], row1[j]);
        p0[k] = make_float4(a0p[3 * j], a0p[3 * j + 1], a0p[3 * j + 2], 0.f);
        q0[k] = make_float4(a0q[4 * j], a0q[4 * j + 1], a0q[4 * j + 2], a0q[4 * j + 3]);
        p1[k] = make_float4(a1p[3 * j], a1p[3 * j + 1], a1p[3 * j + 2], 0.f);
        q1[k] = make_float4(a1q[4 * j], a1q[4 * j + 1], a1q[4 * j + 2], a1q[4 * j + 3]);
        const float* P = params8 + 8 * (size_t)j;
        prm[2 * k] = make_float4(P[0], P[1], P[2], P[3]);
        prm[2 * k + 1] = make_float4(P[4], P[5], P[6], P[7]);
    }
    int rc = 0;
    size_t R = (size_t)MAXR * n;
#define A(p, cnt) if (!rc) rc = pb_alloc(ctx, &s->p, (cnt))
    A(type, n); A(rows, n); A(bodies, n); A(a0p, n); A(a0q, n); A(a1p, n); A(a1q, n); A(prm, 2 * (size_t)n); A(state, 2 * (size_t)n);
    A(linC, R); A(a0T, R); A(a1K, R); A(a0tMin, R); A(a1tMax, R); A(soft, R); A(lambda, R);
#undef A
    if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpy(s->type, t.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->rows, rr.data(), sizeof(int2) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a0p, p0.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a0q, q0.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a1p, p1.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->a1q, q1.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->prm, prm.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemcpy(s->state, st.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    PB_CUDA(ctx, cudaMemset(s->lambda, 0, sizeof(float) * R));
    ctx->nJoints = n;
    return PB_OK;
}

// setters called on live joints (RevoluteJoint::setDriveVelocity, ServoJoint::setTargetAngle, ...): new params in the
// caller's joint order; persistent joint state (gear angles, accumulated impulses) is kept
int pb_joints_update_params(pb_ctx* ctx, int n, const float* params8) {
    JointStore* s = store(ctx);
    if (!s || n != s->n) return pb_fail(ctx, PB_EINVAL, "pb_update_joint_params: joint count mismatch");
    std::vector<float4> prm(2 * (size_t)n);
    for (int k = 0; k < n; ++k) {
        const float* P = params8 + 8 * (size_t)s->order[k];
        prm[2 * k] = make_float4(P[0], P[1], P[2], P[3]);
        prm[2 * k + 1] = make_float4(P[4], P[5], P[6], P[7]);
    }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(s->prm, prm.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    return PB_OK;
}

// oldIndex[j] = index the caller's joint j had in the PREVIOUS pb_upload_joints call, or -1 for a new joint
int pb_joints_keep_state(pb_ctx* ctx, int n, const int* oldIndex) {
    JointStore* s = store(ctx);
    if (!s || n != s->n) return n == 0 ? PB_OK : pb_fail(ctx, PB_EINVAL, "pb_keep_joint_state: joint count mismatch");
    if (s->prevOrder.empty()) return PB_OK;
    std::vector<int> oldSlot(s->prevOrder.size(), -1);
    for (size_t k = 0; k < s->prevOrder.size(); ++k) oldSlot[s->prevOrder[k]] = (int)k;
    std::vector<float4> st(2 * (size_t)n);
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(st.data(), s->state, sizeof(float4) * 2 * n, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; ++k) {
        int oj = oldIndex[s->order[k]];
        if (oj < 0 || oj >= (int)oldSlot.size() || oldSlot[oj] < 0) continue;
        st[2 * k] = s->prevState[2 * (size_t)oldSlot[oj]];
        st[2 * k + 1] = s->prevState[2 * (size_t)oldSlot[oj] + 1];
    }
    PB_CUDA(ctx, cudaMemcpy(s->state, st.data(), sizeof(float4) * 2 * n, cudaMemcpyHostToDevice));
    return PB_OK;
}

int pb_joint_begin_step(pb_ctx* ctx) {
    if (!ctx->nJoints) return PB_OK;
    JointDev J = devView(ctx);
    ++ctx->launches, k_joint_begin<<<pb_grid(J.n, 128), 128, 0, ctx->stream>>>(J, store(ctx)->bodies, ctx->kinematic, ctx->pos, ctx->quat);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// ---- per-group joint lists (islands on) -------------------------------------------------------------------------------------------
// jointOrder lists the joints by (group, colour), stable; jointStart[g * 8 + c] (local groups g < G, colours 0..7) and
// jointStart[G * 8 + c] (global group, colours 0..8) are the run starts.  A local group's CTA walks its own runs; the device-wide
// sweep walks the global runs.  Overflow-bucket joints are always global (islands.cu), so that run equals the static range.
struct JointColorStarts { int s[PB_JOINT_COLORS + 1]; };
__global__ void k_joint_keys(int nJ, const int2* __restrict__ bodies, const int* __restrict__ bodyGroup, int G, JointColorStarts cs,
                             unsigned int* __restrict__ key, int* __restrict__ hist) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nJ) return;
    int2 bb = bodies[j];
    int b = bb.x >= 0 ? bb.x : bb.y;
    int g = b >= 0 ? bodyGroup[b] : G;
    int c = 0;
    while (c < PB_JOINT_COLORS - 1 && j >= cs.s[c + 1]) ++c;
    if (c == 8) g = G;
    unsigned int k = (unsigned int)(g * 8 + c);
    key[j] = k;
    atomicAdd(&hist[k], 1);
}

// placed by a counting sort: the run table (scanned key histogram) gives every (group, colour) run its first slot, a joint takes run
// start + arrival rank.  The order inside a run is immaterial: joints of one colour share no body (Physecs.cpp:690-710), and the
// sequential overflow bucket is not listed here (the solver walks its static range in creation order).
__global__ void k_joint_scatter(int nJ, const unsigned int* __restrict__ key, const int* __restrict__ start, int* __restrict__ fill, int* __restrict__ order) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nJ) return;
    const unsigned int k = key[j];
    order[start[k] + atomicAdd(&fill[k], 1)] = j;
}

int pb_joint_lists_alloc(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    if (!s || !ctx->nJoints) return PB_OK;
    const int n = s->n, G = ctx->islandGroups, nKeys = G * 8 + 9;
    int rc;
    if (ctx->jointListCap < n || !ctx->jointStart) {
        if ((rc = pb_alloc(ctx, &ctx->jointKey, (size_t)n)) || (rc = pb_alloc(ctx, &ctx->jointSortTmp[0], (size_t)nKeys + 1)) || (rc = pb_alloc(ctx, &ctx->jointSortTmp[1], (size_t)n)) ||
            (rc = pb_alloc(ctx, &ctx->jointStart, (size_t)nKeys + 1))) return rc;
        ctx->jointListCap = n;
    }
    return PB_OK;
}

// (jointStart and the fill counters jointSortTmp[0] arrive zeroed: contacts.cu k_build_clear)
int pb_joint_lists(pb_ctx* ctx) {
    JointStore* s = store(ctx);
    if (!s || !ctx->nJoints) return PB_OK;
    const int n = s->n, G = ctx->islandGroups, nKeys = G * 8 + 9;
    int rc;
    if ((rc = pb_joint_lists_alloc(ctx))) return rc;
    JointColorStarts cs;
    for (int c = 0; c <= PB_JOINT_COLORS; ++c) cs.s[c] = ctx->jointColorStart[c];
    ++ctx->launches, k_joint_keys<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, s->bodies, ctx->bodyGroup, G, cs, (unsigned int*)ctx->jointKey, ctx->jointStart);
    if ((rc = pb_exclusive_scan(ctx, ctx->jointStart, ctx->jointStart, nKeys + 1, (int*)ctx->radixHist))) return rc;
    ++ctx->launches, k_joint_scatter<<<pb_grid(n, 256), 256, 0, ctx->stream>>>(n, (const unsigned int*)ctx->jointKey, ctx->jointStart, ctx->jointSortTmp[0], ctx->jointSortTmp[1]);
    ctx->jointOrder = ctx->jointSortTmp[1];
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// device view for the persistent substep kernel (solver.cu); false when the scene has no joints
bool pb_joint_view(pb_ctx* ctx, JointDev* out) {
    if (!ctx->nJoints || !store(ctx)) return false;
    *out = devView(ctx);
    return true;
}

// TGS "soft step" substep loop on the device.
//
// Per substep (reference src/Physecs.cpp:364-531):
//   k_integrate_v     gravity + implicit gyroscopic update, massTemp (world inverse inertia), reset pseudo velocities  :443-468
//   k_contact_prep    world arms, separation, r x n, friction direction, effective masses                            :368-426 + ContactConstraints.cpp:4-31
//   [joint kernels]   joints.cu
//   k_contact_solve   one launch per colour and iteration (normal rows then friction rows of each manifold)           ContactConstraints.cpp:33-124
//   k_integrate_x     positions / orientations (+ pseudo velocities, COM re-anchoring)                                :494-511
//   k_contact_solve   relaxation pass: hard contacts only, no bias                                                    :516-519
// Velocity triple buffering replaces the reference's component <-> velocityTemp copies:
//   vel      = component velocity at substep start (what contact prep reads for the friction direction, :406-412)
//   velPre   = post-gravity/gyro component velocity (what friction rows read all substep long, quirk Q3, ContactConstraints.cpp:92-102)
//   velLive  = velocityTemp, iterated by the solver; becomes `vel` of the next substep by pointer swap (:523-530).
// The friction increment relVel_t / kT is therefore constant within a substep and is precomputed in k_contact_prep.
#include "pb_ctx.h"
#include "pb_math.cuh"

__device__ __forceinline__ M3 loadM3(const float4* __restrict__ p, int i) {
    M3 r; r.c[0] = mk3(p[3 * i]); r.c[1] = mk3(p[3 * i + 1]); r.c[2] = mk3(p[3 * i + 2]); return r;
}
__device__ __forceinline__ void storeM3(float4* __restrict__ p, int i, const M3& a) {
    p[3 * i] = f4(a.c[0]); p[3 * i + 1] = f4(a.c[1]); p[3 * i + 2] = f4(a.c[2]);
}
// MathUtil.h:10-21
__device__ __forceinline__ V3 solve33(const M3& A, V3 b) {
    V3 c12 = cross(A.c[1], A.c[2]);
    float det = dot(A.c[0], c12);
    if (det == 0.f) return mk3(0.f);
    float inv = 1.f / det;
    return mk3(inv * dot(b, c12), inv * dot(A.c[0], cross(b, A.c[2])), inv * dot(A.c[0], cross(A.c[1], b)));
}

__global__ void __launch_bounds__(128) k_integrate_v(int nDyn, float h, float g, const int* __restrict__ kinematic,
    const float4* __restrict__ quat, const float4* __restrict__ vel, const float4* __restrict__ angvel, const float4* __restrict__ invIL,
    float4* __restrict__ velPre, float4* __restrict__ angvelPre, float4* __restrict__ velLive, float4* __restrict__ angvelLive,
    float4* __restrict__ invIW, float4* __restrict__ pseudoLin, float4* __restrict__ pseudoAng) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDyn) return;
    if (kinematic[i]) return;
    M3 rot = mat3_cast(mkq(quat[i]));
    M3 invRot = transpose(rot);
    V3 v = mk3(vel[i]) + h * mk3(0.f, -g, 0.f);
    V3 wl = mul(invRot, mk3(angvel[i]));
    M3 invI = loadM3(invIL, i);
    M3 I = inverse(invI);
    V3 Iw = mul(I, wl);
    V3 f = h * cross(wl, Iw);
    M3 J = I + h * (mul(matrixCross3(wl), I) - matrixCross3(Iw));
    wl = wl - solve33(J, f);
    V3 w = mul(rot, wl);
    velPre[i] = f4(v); angvelPre[i] = f4(w);
    velLive[i] = f4(v); angvelLive[i] = f4(w);
    pseudoLin[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    pseudoAng[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    storeM3(invIW, i, mul(mul(rot, invI), invRot));
}

__global__ void __launch_bounds__(128) k_contact_prep(int nM, const int2* __restrict__ cBodies, const int2* __restrict__ cRowsT,
    const float4* __restrict__ cNormal, const int* __restrict__ cPointOfs, const int* __restrict__ cNp,
    const float4* __restrict__ pR0T, const float4* __restrict__ pR1,
    const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ comInvMass,
    const float4* __restrict__ vel, const float4* __restrict__ angvel, const float4* __restrict__ velPre, const float4* __restrict__ angvelPre,
    const float4* __restrict__ invIW,
    float4* __restrict__ rowA, float4* __restrict__ rowB, float4* __restrict__ rowC, float4* __restrict__ rowD,
    float4* __restrict__ rowE, float4* __restrict__ rowF, float4* __restrict__ rowG, float2* __restrict__ rowL) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nM) return;
    int2 bb = cBodies[s];
    int2 rr = cRowsT[s];
    V3 n = mk3(cNormal[s]);
    Q4 q0 = mkq(quat[rr.x]), q1 = mkq(quat[rr.y]);
    V3 com0 = mk3(0.f), v0 = mk3(0.f), w0 = mk3(0.f), vp0 = mk3(0.f), wp0 = mk3(0.f);
    V3 com1 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f), vp1 = mk3(0.f), wp1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    M3 I0, I1;
    I0.c[0] = I0.c[1] = I0.c[2] = mk3(0.f); I1 = I0;
    if (bb.x >= 0) {
        float4 c = comInvMass[bb.x];
        com0 = mk3(pos[rr.x]) + rotate(q0, mk3(c)); im0 = c.w;
        v0 = mk3(vel[bb.x]); w0 = mk3(angvel[bb.x]); vp0 = mk3(velPre[bb.x]); wp0 = mk3(angvelPre[bb.x]);
        I0 = loadM3(invIW, bb.x);
    }
    if (bb.y >= 0) {
        float4 c = comInvMass[bb.y];
        com1 = mk3(pos[rr.y]) + rotate(q1, mk3(c)); im1 = c.w;
        v1 = mk3(vel[bb.y]); w1 = mk3(angvel[bb.y]); vp1 = mk3(velPre[bb.y]); wp1 = mk3(angvelPre[bb.y]);
        I1 = loadM3(invIW, bb.y);
    }
    int po = cPointOfs[s], np = cNp[s];
    for (int k = 0; k < np; ++k) {
        float4 a = pR0T[po + k];
        V3 r0 = rotate(q0, mk3(a));
        V3 r1 = rotate(q1, mk3(pR1[po + k]));
        V3 cp0 = com0 + r0, cp1 = com1 + r1;
        float cn = dot(cp1 - cp0, n);
        V3 r0xn = cross(r0, n), r1xn = cross(r1, n);
        V3 rel = v1 + cross(w1, r1) - v0 - cross(w0, r0);
        float relN = dot(rel, n);
        V3 t = rel - relN * n;
        float tl = length(t);
        if (tl) t = t / tl;
        V3 r0xt = cross(r0, t), r1xt = cross(r1, t);
        V3 r0xnt = mul(I0, r0xn), r1xnt = mul(I1, r1xn), r0xtt = mul(I0, r0xt), r1xtt = mul(I1, r1xt);
        float kN = dot(n, n) * (im0 + im1) + dot(r0xn, r0xnt) + dot(r1xn, r1xnt);
        float kT = dot(t, t) * (im0 + im1) + dot(r0xt, r0xtt) + dot(r1xt, r1xtt);
        // friction rows read the stale component velocity (quirk Q3): constant for the whole substep
        float lamT0 = 0.f;
        if (kT != 0.f) {
            float relT = dot(-t, vp0) + dot(-r0xt, wp0) + dot(t, vp1) + dot(r1xt, wp1);
            lamT0 = relT / kT;
        }
        rowA[po + k] = f4(r0xn, cn);
        rowB[po + k] = f4(r1xn, kN);
        rowC[po + k] = f4(r0xnt, a.w);
        rowD[po + k] = f4(r1xnt, lamT0);
        rowE[po + k] = f4(t, kT != 0.f ? 1.f : 0.f);
        rowF[po + k] = f4(r0xtt, 0.f);
        rowG[po + k] = f4(r1xtt, 0.f);
        rowL[po + k] = make_float2(0.f, 0.f);
    }
}

// One colour: manifolds [start, start+count).  useBias=0 && skipSoft=1 is the relaxation pass.
__global__ void __launch_bounds__(128) k_contact_solve(int start, int count, int useBias, int skipSoft, float h,
    const int2* __restrict__ cBodies, const float4* __restrict__ cNormal, const float4* __restrict__ cSoft,
    const int* __restrict__ cPointOfs, const int* __restrict__ cNp, const float4* __restrict__ comInvMass,
    float4* __restrict__ velLive, float4* __restrict__ angvelLive,
    const float4* __restrict__ rowA, const float4* __restrict__ rowB, const float4* __restrict__ rowC, const float4* __restrict__ rowD,
    const float4* __restrict__ rowE, const float4* __restrict__ rowF, const float4* __restrict__ rowG, float2* __restrict__ rowL) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int s = start + i;
    float4 soft = cSoft[s];
    if (skipSoft && soft.x != 0.f) return;
    int2 bb = cBodies[s];
    float4 nf = cNormal[s];
    V3 n = mk3(nf);
    float friction = nf.w;
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (bb.x >= 0) { v0 = mk3(velLive[bb.x]); w0 = mk3(angvelLive[bb.x]); im0 = comInvMass[bb.x].w; }
    if (bb.y >= 0) { v1 = mk3(velLive[bb.y]); w1 = mk3(angvelLive[bb.y]); im1 = comInvMass[bb.y].w; }
    int po = cPointOfs[s], np = cNp[s];
    float lamN[4], lamT[4];
    for (int k = 0; k < np; ++k) {
        float4 A = rowA[po + k], B = rowB[po + k], C = rowC[po + k], D = rowD[po + k];
        float2 L = rowL[po + k];
        lamN[k] = L.x; lamT[k] = L.y;
        float kN = B.w;
        if (kN == 0.f) continue;
        V3 r0xn = mk3(A), r1xn = mk3(B);
        float rv = dot(-n, v0) + dot(-r0xn, w0) + dot(n, v1) + dot(r1xn, w1);
        float effMass = 1.f / kN;
        float lambda;
        if (soft.x != 0.f) {
            float af = 2.f * 3.14159265358979323846f * soft.y;
            float stiffness = af * af * effMass;
            float damping = 2.f * af * soft.z * effMass;
            float gamma = 1.f / (damping + h * stiffness);
            float beta = h * stiffness / (damping + h * stiffness);
            lambda = (rv + beta * A.w / h) / (kN + gamma / h);
        } else {
            lambda = (rv - C.w + (useBias ? 0.1f * A.w / h : 0.f)) * effMass;
        }
        float prev = lamN[k];
        float tot = fminf(prev + lambda, 0.f);
        lamN[k] = tot;
        lambda = tot - prev;
        v0 += lambda * im0 * n; w0 += lambda * mk3(C);
        v1 -= lambda * im1 * n; w1 -= lambda * mk3(D);
    }
    for (int k = 0; k < np; ++k) {
        float4 E = rowE[po + k];
        if (E.w != 0.f) {
            float4 D = rowD[po + k];
            V3 t = mk3(E);
            float limit = friction * lamN[k];
            float prev = lamT[k];
            float tot = gclamp(prev + D.w, limit, -limit);
            lamT[k] = tot;
            float lambda = tot - prev;
            v0 += lambda * im0 * t; w0 += lambda * mk3(rowF[po + k]);
            v1 -= lambda * im1 * t; w1 -= lambda * mk3(rowG[po + k]);
        }
        rowL[po + k] = make_float2(lamN[k], lamT[k]);
    }
    if (bb.x >= 0) { velLive[bb.x] = f4(v0); angvelLive[bb.x] = f4(w0); }
    if (bb.y >= 0) { velLive[bb.y] = f4(v1); angvelLive[bb.y] = f4(w1); }
}

// Sequential overflow bucket (colour 63): one thread walks the manifolds in order.
__global__ void k_contact_solve_seq(int start, int count, int useBias, int skipSoft, float h,
    const int2* __restrict__ cBodies, const float4* __restrict__ cNormal, const float4* __restrict__ cSoft,
    const int* __restrict__ cPointOfs, const int* __restrict__ cNp, const float4* __restrict__ comInvMass,
    float4* __restrict__ velLive, float4* __restrict__ angvelLive,
    const float4* __restrict__ rowA, const float4* __restrict__ rowB, const float4* __restrict__ rowC, const float4* __restrict__ rowD,
    const float4* __restrict__ rowE, const float4* __restrict__ rowF, const float4* __restrict__ rowG, float2* __restrict__ rowL);

__global__ void __launch_bounds__(128) k_integrate_x(int nDyn, float h, const int* __restrict__ kinematic, float4* __restrict__ pos, float4* __restrict__ quat,
    const float4* __restrict__ comInvMass, const float4* __restrict__ velLive, const float4* __restrict__ angvelLive,
    const float4* __restrict__ pseudoLin, const float4* __restrict__ pseudoAng) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDyn) return;
    if (kinematic[i]) return;
    float4 pl = pseudoLin[i];
    int cnt = __float_as_int(pl.w);
    float scale = cnt ? 1.f / (float)cnt : 1.f;
    V3 p = mk3(pos[i]);
    Q4 q = mkq(quat[i]);
    V3 com = mk3(comInvMass[i]);
    p = p + (h * mk3(velLive[i]) + scale * mk3(pl));
    V3 prevCom = rotate(q, com);
    V3 hw = 0.5f * (h * mk3(angvelLive[i]) + scale * mk3(pseudoAng[i]));
    Q4 dq; dq.w = 0.f; dq.x = hw.x; dq.y = hw.y; dq.z = hw.z;
    Q4 add = qmul(dq, q);
    q.x += add.x; q.y += add.y; q.z += add.z; q.w += add.w;
    q = qnormalize(q);
    p = p + (prevCom - rotate(q, com));
    pos[i] = f4(p); quat[i] = f4(q);
}

int pb_joint_prep(pb_ctx* ctx, float h);
int pb_joint_solve(pb_ctx* ctx, float h, int warmStart);

int pb_solve(pb_ctx* ctx, float dt, int substeps, int iterations, float gravity) {
    const int nDyn = ctx->nDyn;
    if (nDyn == 0) return PB_OK;
    const int nM = ctx->lastCounts.n_manifolds;
    const int* colorStart = ctx->hCounters + CNT_COLORSTART;
    const int cur = ctx->curBuf;
    float h = dt / (float)substeps;
    int ncol = ctx->lastCounts.n_colors;
    for (int sub = 0; sub < substeps; ++sub) {
        ++ctx->launches, k_integrate_v<<<pb_grid(nDyn, 128), 128, 0, ctx->stream>>>(nDyn, h, gravity, ctx->kinematic, ctx->quat, ctx->vel, ctx->angvel, ctx->invIL,
            ctx->velPre, ctx->angvelPre, ctx->velLive, ctx->angvelLive, ctx->invIW, ctx->pseudoLin, ctx->pseudoAng);
        if (nM > 0) pb_prof_begin(ctx, PROF_CONTACT_PREP);
        if (nM > 0)
            ++ctx->launches, k_contact_prep<<<pb_grid(nM, 128), 128, 0, ctx->stream>>>(nM, ctx->cBodies, ctx->cRowsT, ctx->cNormal, ctx->cPointOfsBuf[cur], ctx->cNpBuf[cur],
                ctx->pR0T[cur], ctx->pR1, ctx->pos, ctx->quat, ctx->comInvMass, ctx->vel, ctx->angvel, ctx->velPre, ctx->angvelPre, ctx->invIW,
                ctx->rowA, ctx->rowB, ctx->rowC, ctx->rowD, ctx->rowE, ctx->rowF, ctx->rowG, ctx->rowL);
        if (nM > 0) pb_prof_end(ctx);
        if (ctx->nJoints) { int rc = pb_joint_prep(ctx, h); if (rc) return rc; }
        for (int it = 0; it <= iterations; ++it) {
            const bool relax = it == iterations;
            if (relax)
                ++ctx->launches, k_integrate_x<<<pb_grid(nDyn, 128), 128, 0, ctx->stream>>>(nDyn, h, ctx->kinematic, ctx->pos, ctx->quat, ctx->comInvMass,
                    ctx->velLive, ctx->angvelLive, ctx->pseudoLin, ctx->pseudoAng);
            if (nM > 0) pb_prof_begin(ctx, PROF_SOLVE_PASS);
            for (int c = 0; c < ncol; ++c) {
                int start = colorStart[c], count = colorStart[c + 1] - start;
                if (count <= 0) continue;
                if (c == PB_OVERFLOW_COLOR)
                    ++ctx->launches, k_contact_solve_seq<<<1, 1, 0, ctx->stream>>>(start, count, relax ? 0 : 1, relax ? 1 : 0, h, ctx->cBodies, ctx->cNormal, ctx->cSoft,
                        ctx->cPointOfsBuf[cur], ctx->cNpBuf[cur], ctx->comInvMass, ctx->velLive, ctx->angvelLive,
                        ctx->rowA, ctx->rowB, ctx->rowC, ctx->rowD, ctx->rowE, ctx->rowF, ctx->rowG, ctx->rowL);
                else
                    ++ctx->launches, k_contact_solve<<<pb_grid(count, 128), 128, 0, ctx->stream>>>(start, count, relax ? 0 : 1, relax ? 1 : 0, h, ctx->cBodies, ctx->cNormal, ctx->cSoft,
                        ctx->cPointOfsBuf[cur], ctx->cNpBuf[cur], ctx->comInvMass, ctx->velLive, ctx->angvelLive,
                        ctx->rowA, ctx->rowB, ctx->rowC, ctx->rowD, ctx->rowE, ctx->rowF, ctx->rowG, ctx->rowL);
            }
            if (nM > 0) pb_prof_end(ctx);
            if (!relax && ctx->nJoints) { int rc = pb_joint_solve(ctx, h, it == 0); if (rc) return rc; }
        }
        // write-back (Physecs.cpp:523-530): velocityTemp becomes the component velocity
        std::swap(ctx->vel, ctx->velLive);
        std::swap(ctx->angvel, ctx->angvelLive);
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

__global__ void k_contact_solve_seq(int start, int count, int useBias, int skipSoft, float h,
    const int2* __restrict__ cBodies, const float4* __restrict__ cNormal, const float4* __restrict__ cSoft,
    const int* __restrict__ cPointOfs, const int* __restrict__ cNp, const float4* __restrict__ comInvMass,
    float4* __restrict__ velLive, float4* __restrict__ angvelLive,
    const float4* __restrict__ rowA, const float4* __restrict__ rowB, const float4* __restrict__ rowC, const float4* __restrict__ rowD,
    const float4* __restrict__ rowE, const float4* __restrict__ rowF, const float4* __restrict__ rowG, float2* __restrict__ rowL) {
    for (int i = 0; i < count; ++i) {
        int s = start + i;
        float4 soft = cSoft[s];
        if (skipSoft && soft.x != 0.f) continue;
        int2 bb = cBodies[s];
        float4 nf = cNormal[s];
        V3 n = mk3(nf);
        float friction = nf.w;
        V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
        float im0 = 0.f, im1 = 0.f;
        if (bb.x >= 0) { v0 = mk3(velLive[bb.x]); w0 = mk3(angvelLive[bb.x]); im0 = comInvMass[bb.x].w; }
        if (bb.y >= 0) { v1 = mk3(velLive[bb.y]); w1 = mk3(angvelLive[bb.y]); im1 = comInvMass[bb.y].w; }
        int po = cPointOfs[s], np = cNp[s];
        for (int k = 0; k < np; ++k) {
            float4 A = rowA[po + k], B = rowB[po + k], C = rowC[po + k], D = rowD[po + k];
            float2 L = rowL[po + k];
            float kN = B.w;
            if (kN == 0.f) continue;
            float rv = dot(-n, v0) + dot(-mk3(A), w0) + dot(n, v1) + dot(mk3(B), w1);
            float effMass = 1.f / kN;
            float lambda;
            if (soft.x != 0.f) {
                float af = 2.f * 3.14159265358979323846f * soft.y;
                float stiffness = af * af * effMass;
                float damping = 2.f * af * soft.z * effMass;
                float gamma = 1.f / (damping + h * stiffness);
                float beta = h * stiffness / (damping + h * stiffness);
                lambda = (rv + beta * A.w / h) / (kN + gamma / h);
            } else lambda = (rv - C.w + (useBias ? 0.1f * A.w / h : 0.f)) * effMass;
            float tot = fminf(L.x + lambda, 0.f);
            lambda = tot - L.x;
            rowL[po + k] = make_float2(tot, L.y);
            v0 += lambda * im0 * n; w0 += lambda * mk3(C);
            v1 -= lambda * im1 * n; w1 -= lambda * mk3(D);
        }
        for (int k = 0; k < np; ++k) {
            float4 E = rowE[po + k];
            if (E.w == 0.f) continue;
            float4 D = rowD[po + k];
            float2 L = rowL[po + k];
            float limit = friction * L.x;
            float tot = gclamp(L.y + D.w, limit, -limit);
            float lambda = tot - L.y;
            rowL[po + k] = make_float2(L.x, tot);
            V3 t = mk3(E);
            v0 += lambda * im0 * t; w0 += lambda * mk3(rowF[po + k]);
            v1 -= lambda * im1 * t; w1 -= lambda * mk3(rowG[po + k]);
        }
        if (bb.x >= 0) { velLive[bb.x] = f4(v0); angvelLive[bb.x] = f4(w0); }
        if (bb.y >= 0) { velLive[bb.y] = f4(v1); angvelLive[bb.y] = f4(w1); }
    }
}

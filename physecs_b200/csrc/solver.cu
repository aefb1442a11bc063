// TGS "soft step" substep loop on the device -- per substep: two streaming kernels + ONE persistent cooperative kernel for
// everything that is ordered by constraint colour.
//
// Reference (src/Physecs.cpp:364-531), per substep:
//   integrate velocities   gravity + implicit gyroscopic update, massTemp (world inverse inertia), reset pseudo velocities  :443-468
//   contact prep           world arms, separation, r x n, friction direction, effective masses        :368-426 + ContactConstraints.cpp:4-31
//   joint prep             row fill + preSolve / NGS pass per joint colour                              joints.cuh
//   iterations             contact colours (normal rows then friction rows of each manifold), then joint colours   ContactConstraints.cpp:33-124
//   integrate positions    positions / orientations (+ pseudo velocities, COM re-anchoring)            :494-511
//   relaxation             one more contact pass: hard contacts only, no bias                           :516-519
//
// Why a persistent kernel for the coloured part: a 1 M-body step has ~9 contact colours whose sizes fall off geometrically (700 k, 600 k, 400 k, 150 k,
// 30 k, 4 k, 500, 100 manifolds).  Launched separately, every colour pays ~9 us of launch + dependent-load latency however
// small it is, 3 passes x 4 substeps per step; small scenes (ragdoll batches) are launch-bound outright.  In k_substep_solve
// the grid is sized to be co-resident (cudaLaunchCooperativeKernel), each colour is a grid-stride loop and colours are
// separated by a device-wide barrier (one atomic per CTA, ~1-2 us).  The arithmetic is unchanged, so parity is unchanged:
// inside a colour no two manifolds share a dynamic body and colours still run in order.
// The two big streaming phases (integrate velocities, contact prep) stay ordinary launches: contact prep needs ~125
// registers, and folding it into the persistent kernel would cap the solve colours at 16 warps/SM (measured: 2.8 TB/s
// instead of 4.1 TB/s on the large colours).  The persistent kernel is held to 64 registers (32 warps/SM).
//
// Velocity triple buffering replaces the reference's component <-> velocityTemp copies:
//   vel      = component velocity at substep start (what contact prep reads for the friction direction, :406-412)
//   velPre   = post-gravity/gyro component velocity (what friction rows read all substep long, quirk Q3, ContactConstraints.cpp:92-102)
//   velLive  = velocityTemp, iterated by the solver; becomes `vel` of the next substep by pointer swap (:523-530).
// Each buffer interleaves {v.xyz, invMass} and {w.xyz, 0} per body (32 bytes = one DRAM sector), so a solver gather of a
// body's velocity state is one sector instead of three (v, w and the invMass word of comInvMass).
// The friction increment relVel_t / kT is therefore constant within a substep and is precomputed in the prep phase.
//
// Memory: every array a step mutates is read with __ldcg (L2) -- other SMs write it between barriers and L1 is not
// coherent; arrays that are constant for the whole launch keep the read-only path.
#include "pb_ctx.h"
#include "pb_math.cuh"
#include "joints.cuh"
#include <cstdlib>
#include <type_traits>

bool pb_joint_view(pb_ctx* ctx, JointDev* out);

struct SubstepParams {
    int nDyn, substeps, iterations;
    float h, g;
    const int* counters;             // device counters block: CNT_MANIFOLDS, CNT_COLORSTART..
    // bodies
    const int* kinematic; const float4* comInvMass; const float4* invIL;
    float4* pos; float4* quat;
    float4* velA; float4* angvelA;   // substep-start velocity on entry (buffers A / B swap every substep)
    float4* velB; float4* angvelB;
    float4* bodyRec;                 // 8 float4 (128 B) per dynamic body, written by integrate-v, read by the prep kernels (layout below)
    float4* pseudoLin; float4* pseudoAng;
    // contact constraints
    const int4* cHead; const int2* cBodies; const int2* cRowsT; const float4* cNormal; const float4* cSoft; const int* cPointOfs; const int* cNp;
    const float4* pR0T; const float4* pR1;
    float4* rowA; float4* rowB; float4* rowC; float4* rowD; float4* rowE; float4* rowF; float4* rowG; float2* rowL;
    int rowExtra;                    // rows of a manifold's FIRST point live at its solve slot s; points k >= 1 at rowExtra + (firstPoint - s) + k - 1
    // joints
    int hasJoints; JointDev J; int jointColorStart[PB_JOINT_COLORS + 1];
    // solve-order run table (contacts.cu): keyStart[g * PB_KEY_COLORS + 2 c (+1)] = first single- (multi-) point slot of colour c in group g.
    // Groups 0..G-1 are islands small enough for one CTA (islands.cu, only when islandsOn), group G is the device-wide sweep.
    const int* keyStart; int G; int islandsOn;
    const int* jointOrder; const int* jointStart;     // per-group joint runs (islandsOn): jointStart[g * 8 + c]
    const int* bodyOrder; const int* bodyStart;       // per-group body lists (islands.cu; nullptr unless the whole-step kernel may run group by group)
    // grid barrier + optional phase timing (ns per phase kind, accumulated by CTA 0)
    unsigned int* barrier;
    unsigned long long* profNs;
    int clusterBarrier;              // the launch is one thread-block cluster: GridBarrier uses barrier.cluster
};

template <bool LOCAL_SYNC, bool L1> struct SweepMode { static constexpr bool local = LOCAL_SYNC, l1 = L1; };

enum { PH_INTEGRATE_V = 0, PH_PREP, PH_CONTACT_PASS, PH_JOINT_SOLVE, PH_INTEGRATE_X, PH_LOCAL, PH_KINDS };
#define PROF_LOCAL (2 * PH_KINDS + 2 * PB_MAX_COLORS)   // CTA 0's own island sweep, split by colour-phase kind: ns[3] then count[3] (NGS, contact, joint)
#define PROF_WORDS (PROF_LOCAL + 6)                     // ns + count per phase kind, then ns + count per contact colour, then the local split

__device__ __forceinline__ M3 loadM3ro(const float4* __restrict__ p, int i) {
    M3 r; r.c[0] = mk3(p[3 * i]); r.c[1] = mk3(p[3 * i + 1]); r.c[2] = mk3(p[3 * i + 2]); return r;
}
// MathUtil.h:10-21
__device__ __forceinline__ V3 solve33(const M3& A, V3 b) {
    V3 c12 = cross(A.c[1], A.c[2]);
    float det = dot(A.c[0], c12);
    if (det == 0.f) return mk3(0.f);
    float inv = 1.f / det;
    return mk3(inv * dot(b, c12), inv * dot(A.c[0], cross(b, A.c[2])), inv * dot(A.c[0], cross(A.c[1], b)));
}

// Per-substep body record (128 B, one aligned line): everything contact prep and joint fill need from a dynamic body, so a
// gather costs two 64-byte DRAM accesses instead of seven scattered sectors (ncu: prep traffic was 1.9x algorithmic).
//   r0 q.xyzw | r1 comWorld.xyz, invMass | r2 v.xyz (substep start), I.c2.z | r3 w.xyz (substep start) |
//   r4 vPre.xyz | r5 wPre.xyz | r6 I.c0.xyz, I.c1.x | r7 I.c1.y, I.c1.z, I.c2.x, I.c2.y        (I = world inverse inertia)
struct BodyRec { Q4 q; V3 com; float im; V3 v, w, vp, wp; M3 I; };
// CL = "CTA-local": the caller is the whole-step kernel carrying ONE group of islands through the step (k_step_solve_small, group by
// group): every mutable word it touches is written by its own CTA only, for the whole kernel, so plain (L1-cached) loads and stores
// are coherent -- a dependent access then costs an L1 hit instead of an L2 round trip.  Otherwise L2 (.cg), as everywhere else.
template <bool CL, class T> __device__ __forceinline__ T ldq(const T* p) { if (CL) return *p; return __ldcg(p); }
template <bool CL, class T> __device__ __forceinline__ void stq(T* p, T v) { if (CL) *p = v; else __stcg(p, v); }
template <bool CL = false>
__device__ __forceinline__ BodyRec loadBodyRec(const float4* rec, int b) {
    const float4* r = rec + 8 * (size_t)b;
    float4 r0 = ldq<CL>(r), r1 = ldq<CL>(r + 1), r2 = ldq<CL>(r + 2), r3 = ldq<CL>(r + 3), r4 = ldq<CL>(r + 4), r5 = ldq<CL>(r + 5), r6 = ldq<CL>(r + 6), r7 = ldq<CL>(r + 7);
    BodyRec B;
    B.q = mkq(r0); B.com = mk3(r1); B.im = r1.w; B.v = mk3(r2); B.w = mk3(r3); B.vp = mk3(r4); B.wp = mk3(r5);
    B.I.c[0] = mk3(r6); B.I.c[1] = mk3(r6.w, r7.x, r7.y); B.I.c[2] = mk3(r7.z, r7.w, r2.w);
    return B;
}


// ---- L2 eviction hints -------------------------------------------------------------------------------------------------------------
// A contact pass streams ~150 B of rows per manifold (hundreds of MB per pass, no reuse before the next pass: L2 cannot hold
// them) and gathers / scatters 32-byte velocity records that ARE reused (every body sits in several manifolds, all colours, all
// passes; 32 MB at 1 M bodies).  Rows are therefore read and written evict-first and velocity records evict-last, so the row
// stream does not flush the velocity working set out of L2.  All of it stays .cg (L2 only): other SMs write these arrays
// between grid barriers and L1 is not coherent.
struct L2Hints { unsigned long long first, last; };
__device__ __forceinline__ L2Hints makeL2Hints() {
    L2Hints h;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(h.first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(h.last));
    return h;
}
__device__ __forceinline__ float4 ldHint(const float4* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int4 ldHint(const int4* p, unsigned long long pol) {
    int4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ldHint(const float2* p, unsigned long long pol) {
    float2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stHint(float4* p, float4 v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void stHint(float2* p, float2 v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

// the same accessors for a per-CTA island sweep (L1 = true): plain cached loads / stores, see ldm() in joints.cuh
template <bool L1, class T> __device__ __forceinline__ T ldSolve(const T* p, unsigned long long pol) { if (L1) return *p; return ldHint(p, pol); }
template <bool L1, class T> __device__ __forceinline__ void stSolve(T* p, T v, unsigned long long pol) { if (L1) *p = v; else stHint(p, v, pol); }

// Row addressing: the first point's rows sit at the manifold's own slot, so the solver can load them together with the
// header instead of after it (one dependent memory round trip less for the single-point manifolds that dominate big scenes).
// firstPoint - s is the number of extra points of all earlier manifolds (every manifold has >= 1 point).
__device__ __forceinline__ int rowIndex(const SubstepParams& P, int s, int po, int k) { return k == 0 ? s : P.rowExtra + (po - s) + (k - 1); }

// ---- phases (one unit of work each) -------------------------------------------------------------------------------------------------
template <bool CL = false>
__device__ __forceinline__ void integrateV(const SubstepParams& P, int i, const float4* vel, const float4* angvel, float4* velLive, float4* angvelLive) {
    if (P.kinematic[i]) {    // not integrated (Physecs.cpp:446); the component value just follows the buffer swap
        velLive[2 * i] = ldq<CL>(&vel[2 * i]); angvelLive[2 * i] = ldq<CL>(&angvel[2 * i]);
        return;
    }
    M3 rot = mat3_cast(mkq(ldq<CL>(&P.quat[i])));
    M3 invRot = transpose(rot);
    float4 vin = ldq<CL>(&vel[2 * i]);          // .w carries invMass (the solver's gathers get it for free)
    V3 v = mk3(vin) + P.h * mk3(0.f, -P.g, 0.f);
    V3 wl = mul(invRot, mk3(ldq<CL>(&angvel[2 * i])));
    M3 invI = loadM3ro(P.invIL, i);
    M3 I = inverse(invI);
    V3 Iw = mul(I, wl);
    V3 f = P.h * cross(wl, Iw);
    M3 J = I + P.h * (mul(matrixCross3(wl), I) - matrixCross3(Iw));
    wl = wl - solve33(J, f);
    V3 w = mul(rot, wl);
    velLive[2 * i] = f4(v, vin.w); angvelLive[2 * i] = f4(w);
    if (P.hasJoints) {       // only the joints' NGS pass accumulates into these; without joints integrate-x takes them as zero
        P.pseudoLin[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        P.pseudoAng[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    M3 IW = mul(mul(rot, invI), invRot);
    float4 c = P.comInvMass[i];
    Q4 q = mkq(ldq<CL>(&P.quat[i]));
    V3 comW = mk3(ldq<CL>(&P.pos[i])) + rotate(q, mk3(c));      // same expression contact prep used to evaluate per manifold
    float4* r = P.bodyRec + 8 * (size_t)i;
    r[0] = f4(q); r[1] = f4(comW, c.w);
    r[2] = f4(mk3(vin), IW.c[2].z); r[3] = f4(mk3(ldq<CL>(&angvel[2 * i])));
    r[4] = f4(v); r[5] = f4(w);
    r[6] = f4(IW.c[0], IW.c[1].x); r[7] = make_float4(IW.c[1].y, IW.c[1].z, IW.c[2].x, IW.c[2].y);
}

template <bool CL = false>
__device__ __forceinline__ void contactPrep(const SubstepParams& P, int s, const float4* vel, const float4* angvel) {
    int4 hd = P.cHead[s];
    int2 bb = make_int2(hd.x, hd.y);
    int2 rr = P.cRowsT[s];
    V3 n = mk3(P.cNormal[s]);
    Q4 q0, q1;
    V3 com0 = mk3(0.f), v0 = mk3(0.f), w0 = mk3(0.f), vp0 = mk3(0.f), wp0 = mk3(0.f);
    V3 com1 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f), vp1 = mk3(0.f), wp1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    M3 I0, I1;
    I0.c[0] = I0.c[1] = I0.c[2] = mk3(0.f); I1 = I0;
    // (Measured and dropped: the static side's orientation kept per manifold by the contact build.  The transform rows load in the
    // first wave next to the header, so the quaternion is no deeper in the dependency chain than the body records: no gain, 16 B more.)
    if (bb.x >= 0) { BodyRec B = loadBodyRec<CL>(P.bodyRec, bb.x); q0 = B.q; com0 = B.com; im0 = B.im; v0 = B.v; w0 = B.w; vp0 = B.vp; wp0 = B.wp; I0 = B.I; }
    else q0 = mkq(ldq<CL>(&P.quat[rr.x]));       // static / kinematic side: only its orientation matters (quirk Q25)
    if (bb.y >= 0) { BodyRec B = loadBodyRec<CL>(P.bodyRec, bb.y); q1 = B.q; com1 = B.com; im1 = B.im; v1 = B.v; w1 = B.w; vp1 = B.vp; wp1 = B.wp; I1 = B.I; }
    else q1 = mkq(ldq<CL>(&P.quat[rr.y]));
    int po = hd.z, np = hd.w & 0xff;
    for (int k = 0; k < np; ++k) {
        float4 a = P.pR0T[po + k];
        V3 r0 = rotate(q0, mk3(a));
        V3 r1 = rotate(q1, mk3(P.pR1[po + k]));
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
        const int ri = rowIndex(P, s, po, k);
        stq<CL>(&P.rowA[ri], f4(r0xn, cn));
        stq<CL>(&P.rowB[ri], f4(r1xn, kN));
        stq<CL>(&P.rowC[ri], f4(r0xnt, a.w));
        stq<CL>(&P.rowD[ri], f4(r1xnt, lamT0));
        stq<CL>(&P.rowE[ri], f4(t, kT != 0.f ? 1.f : 0.f));
        stq<CL>(&P.rowF[ri], f4(r0xtt, 0.f));
        stq<CL>(&P.rowG[ri], f4(r1xtt, 0.f));
        stq<CL>(&P.rowL[ri], make_float2(0.f, 0.f));
    }
}

// one manifold of the current colour.  useBias=0 && skipSoft=1 is the relaxation pass.
// Normal row then friction row of one point (ContactConstraints.cpp:33-124); shared by the single-point fast path and the loop.
__device__ __forceinline__ void normalRow(float4 A, float4 B, float4 C, float4 D, float4 soft, int useBias, float h, V3 n, float im0, float im1,
                                          V3& v0, V3& w0, V3& v1, V3& w1, float& lamN) {
    float kN = B.w;
    if (kN == 0.f) return;
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
    float prev = lamN;
    float tot = fminf(prev + lambda, 0.f);
    lamN = tot;
    lambda = tot - prev;
    v0 += lambda * im0 * n; w0 += lambda * mk3(C);
    v1 -= lambda * im1 * n; w1 -= lambda * mk3(D);
}
__device__ __forceinline__ void frictionRow(float4 D, float4 E, float4 F, float4 G, float friction, float lamN, float im0, float im1,
                                            V3& v0, V3& w0, V3& v1, V3& w1, float& lamT) {
    V3 t = mk3(E);
    float limit = friction * lamN;
    float prev = lamT;
    float tot = gclamp(prev + D.w, limit, -limit);
    lamT = tot;
    float lambda = tot - prev;
    v0 += lambda * im0 * t; w0 += lambda * mk3(F);
    v1 -= lambda * im1 * t; w1 -= lambda * mk3(G);
}

template <bool L1>
__device__ __forceinline__ void contactSolve(const SubstepParams& P, int s, int useBias, int skipSoft, float4* velLive, float4* angvelLive, const L2Hints& H) {
    // first wave: header + every row of the first point (their address is the slot itself, no dependence on the header)
    int4 hd = ldSolve<L1>(&P.cHead[s], H.first);
    float4 nf = ldSolve<L1>(&P.cNormal[s], H.first);
    float4 A = ldSolve<L1>(&P.rowA[s], H.first), B = ldSolve<L1>(&P.rowB[s], H.first), C = ldSolve<L1>(&P.rowC[s], H.first), D = ldSolve<L1>(&P.rowD[s], H.first);
    float4 E = ldSolve<L1>(&P.rowE[s], H.first), F = ldSolve<L1>(&P.rowF[s], H.first), G = ldSolve<L1>(&P.rowG[s], H.first);
    float2 L = ldSolve<L1>(&P.rowL[s], H.first);
    const bool isSoft = (hd.w & 0x100) != 0;
    if (skipSoft && isSoft) return;
    float4 soft = isSoft ? P.cSoft[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    V3 n = mk3(nf);
    float friction = nf.w;
    const float h = P.h;
    const int b0 = hd.x, b1 = hd.y, po = hd.z, np = hd.w & 0xff;
    // second wave: the body velocities (one 32-byte sector per body, invMass rides in v.w)
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (b0 >= 0) { float4 t_ = ldSolve<L1>(&velLive[2 * b0], H.last); v0 = mk3(t_); im0 = t_.w; w0 = mk3(ldSolve<L1>(&angvelLive[2 * b0], H.last)); }
    if (b1 >= 0) { float4 t_ = ldSolve<L1>(&velLive[2 * b1], H.last); v1 = mk3(t_); im1 = t_.w; w1 = mk3(ldSolve<L1>(&angvelLive[2 * b1], H.last)); }
    if (np == 1) {
        // the common case (a body resting on a mesh triangle, a sphere pair): two memory round trips in total
        float lamN = L.x, lamT = L.y;
        normalRow(A, B, C, D, soft, useBias, h, n, im0, im1, v0, w0, v1, w1, lamN);
        if (E.w != 0.f) frictionRow(D, E, F, G, friction, lamN, im0, im1, v0, w0, v1, w1, lamT);
        stSolve<L1>(&P.rowL[s], make_float2(lamN, lamT), H.first);
    } else {
        float lamN[4], lamT[4];
        lamN[0] = L.x; lamT[0] = L.y;
        normalRow(A, B, C, D, soft, useBias, h, n, im0, im1, v0, w0, v1, w1, lamN[0]);
        for (int k = 1; k < np; ++k) {
            const int ri = rowIndex(P, s, po, k);
            float4 Ak = ldSolve<L1>(&P.rowA[ri], H.first), Bk = ldSolve<L1>(&P.rowB[ri], H.first), Ck = ldSolve<L1>(&P.rowC[ri], H.first), Dk = ldSolve<L1>(&P.rowD[ri], H.first);
            float2 Lk = ldSolve<L1>(&P.rowL[ri], H.first);
            lamN[k] = Lk.x; lamT[k] = Lk.y;
            normalRow(Ak, Bk, Ck, Dk, soft, useBias, h, n, im0, im1, v0, w0, v1, w1, lamN[k]);
        }
        if (E.w != 0.f) frictionRow(D, E, F, G, friction, lamN[0], im0, im1, v0, w0, v1, w1, lamT[0]);
        stSolve<L1>(&P.rowL[s], make_float2(lamN[0], lamT[0]), H.first);
        for (int k = 1; k < np; ++k) {
            const int ri = rowIndex(P, s, po, k);
            float4 Ek = ldSolve<L1>(&P.rowE[ri], H.first);
            if (Ek.w != 0.f)
                frictionRow(ldSolve<L1>(&P.rowD[ri], H.first), Ek, ldSolve<L1>(&P.rowF[ri], H.first), ldSolve<L1>(&P.rowG[ri], H.first), friction, lamN[k], im0, im1, v0, w0, v1, w1, lamT[k]);
            stSolve<L1>(&P.rowL[ri], make_float2(lamN[k], lamT[k]), H.first);
        }
    }
    if (b0 >= 0) { stSolve<L1>(&velLive[2 * b0], f4(v0, im0), H.last); stSolve<L1>(&angvelLive[2 * b0], f4(w0), H.last); }
    if (b1 >= 0) { stSolve<L1>(&velLive[2 * b1], f4(v1, im1), H.last); stSolve<L1>(&angvelLive[2 * b1], f4(w1), H.last); }
}

// FOUR lanes per manifold: lane k holds point k, so all rows of the manifold are loaded in one wave; the points are then
// applied in order (normal rows, then friction rows) by handing the running body velocities from lane to lane with shuffles.
// Same arithmetic as contactSolve; used when manifolds have several points on average (box stacks, ragdolls), where the
// one-thread version serialises ~3 memory round trips per point.
template <bool L1>
__device__ __forceinline__ void contactSolveQuad(const SubstepParams& P, int s, int lane4, unsigned gmask, int useBias, int skipSoft,
                                                 float4* velLive, float4* angvelLive, const L2Hints& H) {
    int4 hd = ldSolve<L1>(&P.cHead[s], H.first);
    const bool isSoft = (hd.w & 0x100) != 0;
    if (skipSoft && isSoft) return;          // uniform across the four lanes
    float4 nf = ldSolve<L1>(&P.cNormal[s], H.first);
    float4 soft = isSoft ? P.cSoft[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    V3 n = mk3(nf);
    float friction = nf.w;
    const float h = P.h;
    const int b0 = hd.x, b1 = hd.y, po = hd.z, np = hd.w & 0xff;
    const bool mine = lane4 < np;
    float4 A = make_float4(0, 0, 0, 0), B = A, C = A, D = A, E = A, F = A, G = A;
    float2 L = make_float2(0.f, 0.f);
    const int ri = rowIndex(P, s, po, lane4);
    if (mine) {
        A = ldSolve<L1>(&P.rowA[ri], H.first); B = ldSolve<L1>(&P.rowB[ri], H.first); C = ldSolve<L1>(&P.rowC[ri], H.first); D = ldSolve<L1>(&P.rowD[ri], H.first);
        E = ldSolve<L1>(&P.rowE[ri], H.first); F = ldSolve<L1>(&P.rowF[ri], H.first); G = ldSolve<L1>(&P.rowG[ri], H.first);
        L = ldSolve<L1>(&P.rowL[ri], H.first);
    }
    V3 v0 = mk3(0.f), w0 = mk3(0.f), v1 = mk3(0.f), w1 = mk3(0.f);
    float im0 = 0.f, im1 = 0.f;
    if (b0 >= 0) { float4 t_ = ldSolve<L1>(&velLive[2 * b0], H.last); v0 = mk3(t_); im0 = t_.w; w0 = mk3(ldSolve<L1>(&angvelLive[2 * b0], H.last)); }
    if (b1 >= 0) { float4 t_ = ldSolve<L1>(&velLive[2 * b1], H.last); v1 = mk3(t_); im1 = t_.w; w1 = mk3(ldSolve<L1>(&angvelLive[2 * b1], H.last)); }
    float lamN = L.x, lamT = L.y;
#define PB_PASS_ON(r) \
    v0.x = __shfl_sync(gmask, v0.x, r, 4); v0.y = __shfl_sync(gmask, v0.y, r, 4); v0.z = __shfl_sync(gmask, v0.z, r, 4); \
    w0.x = __shfl_sync(gmask, w0.x, r, 4); w0.y = __shfl_sync(gmask, w0.y, r, 4); w0.z = __shfl_sync(gmask, w0.z, r, 4); \
    v1.x = __shfl_sync(gmask, v1.x, r, 4); v1.y = __shfl_sync(gmask, v1.y, r, 4); v1.z = __shfl_sync(gmask, v1.z, r, 4); \
    w1.x = __shfl_sync(gmask, w1.x, r, 4); w1.y = __shfl_sync(gmask, w1.y, r, 4); w1.z = __shfl_sync(gmask, w1.z, r, 4);
    for (int k = 0; k < np; ++k) {
        if (lane4 == k) normalRow(A, B, C, D, soft, useBias, h, n, im0, im1, v0, w0, v1, w1, lamN);
        PB_PASS_ON(k)
    }
    for (int k = 0; k < np; ++k) {
        if (lane4 == k && E.w != 0.f) frictionRow(D, E, F, G, friction, lamN, im0, im1, v0, w0, v1, w1, lamT);
        PB_PASS_ON(k)
    }
#undef PB_PASS_ON
    if (mine) stSolve<L1>(&P.rowL[ri], make_float2(lamN, lamT), H.first);
    if (lane4 == 0) {
        if (b0 >= 0) { stSolve<L1>(&velLive[2 * b0], f4(v0, im0), H.last); stSolve<L1>(&angvelLive[2 * b0], f4(w0), H.last); }
        if (b1 >= 0) { stSolve<L1>(&velLive[2 * b1], f4(v1, im1), H.last); stSolve<L1>(&angvelLive[2 * b1], f4(w1), H.last); }
    }
}

template <bool CL = false>
__device__ __forceinline__ void integrateX(const SubstepParams& P, int i, const float4* velLive, const float4* angvelLive) {
    if (P.kinematic[i]) return;
    // no joints: nobody wrote the pseudo velocities, the same zeros go through the same arithmetic without the 64 B per body of traffic
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    float4 pl = P.hasJoints ? ldq<CL>(&P.pseudoLin[i]) : zero4;
    int cnt = __float_as_int(pl.w);
    float scale = cnt ? 1.f / (float)cnt : 1.f;
    V3 p = mk3(ldq<CL>(&P.pos[i]));
    Q4 q = mkq(ldq<CL>(&P.quat[i]));
    V3 com = mk3(P.comInvMass[i]);
    p = p + (P.h * mk3(ldq<CL>(&velLive[2 * i])) + scale * mk3(pl));
    V3 prevCom = rotate(q, com);
    V3 hw = 0.5f * (P.h * mk3(ldq<CL>(&angvelLive[2 * i])) + scale * mk3(P.hasJoints ? ldq<CL>(&P.pseudoAng[i]) : zero4));
    Q4 dq; dq.w = 0.f; dq.x = hw.x; dq.y = hw.y; dq.z = hw.z;
    Q4 add = qmul(dq, q);
    q.x += add.x; q.y += add.y; q.z += add.z; q.w += add.w;
    q = qnormalize(q);
    p = p + (prevCom - rotate(q, com));
    P.pos[i] = f4(p); P.quat[i] = f4(q);
}

// ---- device-wide barrier ------------------------------------------------------------------------------------------------------------
// Monotonic arrival counter (zeroed before the launch): barrier k completes when it reaches k * gridDim.x.  The release /
// acquire pair on the counter plus the CTA barriers around it order all earlier global writes of every CTA before all later
// reads of every CTA (the same construction cooperative_groups::grid_group::sync uses).
struct GridBarrier {
    unsigned int* counter;
    unsigned int target;
    unsigned long long* profNs;
    unsigned long long tPrev;
    int cluster;        // the grid is ONE thread-block cluster (k_step_solve_small on small scenes): the hardware cluster barrier replaces the counter
    __device__ __forceinline__ void sync(int kind, int color = -1) {
        if (cluster) {
            // barrier.cluster with release / acquire semantics orders the global writes of every CTA of the cluster before the reads
            // that follow it (PTX ISA, "barrier.cluster"): the same guarantee as the counter below, in hardware, for <= 16 CTAs
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            if (profNs && blockIdx.x == 0 && threadIdx.x == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                atomicAdd(&profNs[kind], t - tPrev);
                atomicAdd(&profNs[PH_KINDS + kind], 1ull);
                if (color >= 0 && color < PB_MAX_COLORS) { atomicAdd(&profNs[2 * PH_KINDS + color], t - tPrev); atomicAdd(&profNs[2 * PH_KINDS + PB_MAX_COLORS + color], 1ull); }
                tPrev = t;
            }
            return;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            // release: the CTA barrier above ordered every thread's writes before this increment (cumulativity);
            // acquire: the polling load orders every later read of the CTA after the last arrival
            // (atom, not red: the returned count tells the last arriver that it is last, so it leaves without a polling round trip --
            // measured 1.29 us instead of 1.67 us per barrier on 444 CTAs, tools/micro/barrier_bench.cu)
            unsigned int seen;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(seen) : "l"(counter) : "memory");
            ++seen;
            while (seen < target) { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); }
            if (profNs && blockIdx.x == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                atomicAdd(&profNs[kind], t - tPrev);
                atomicAdd(&profNs[PH_KINDS + kind], 1ull);
                if (color >= 0 && color < PB_MAX_COLORS) { atomicAdd(&profNs[2 * PH_KINDS + color], t - tPrev); atomicAdd(&profNs[2 * PH_KINDS + PB_MAX_COLORS + color], 1ull); }
                tPrev = t;
            }
        }
        __syncthreads();
    }
};

// The narrowphase overflowed an arena (or met an unsupported shape pair): the step's solve is skipped as a whole -- poses, velocities,
// joint state and bounds stay what they were, so the host can grow the arenas and run the step again (capi.cu collectStep).
__device__ __forceinline__ bool stepSkipped(const int* __restrict__ counters) { return (counters[CNT_STATUS] & (PB_ECAPACITY | 0x100)) != 0; }

__global__ void __launch_bounds__(256) k_integrate_v(const __grid_constant__ SubstepParams P) {
    if (stepSkipped(P.counters)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.nDyn) integrateV(P, i, P.velA, P.angvelA, P.velB, P.angvelB);
}

// grid-stride: the grid follows the host's guess of the manifold count, the count itself stays on the device
__global__ void __launch_bounds__(128) k_contact_prep(const __grid_constant__ SubstepParams P) {
    if (stepSkipped(P.counters)) return;
    const int n = P.counters[CNT_MANIFOLDS];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) contactPrep(P, s, P.velA, P.angvelA);
}

// Joint routines are called, not inlined: their register appetite (row builders, 3x3 products) then spills inside the
// callee only, and the contact colour loops -- the bandwidth-critical part -- keep a spill-free 64-register budget.
template <bool L1>
__device__ __noinline__ void jointNgsCall(const SubstepParams& P, int j) {
    jointNgsOne<L1>(P.J, j, P.kinematic, P.comInvMass, P.pseudoLin, P.pseudoAng);
}
template <bool L1>
__device__ __noinline__ void jointSolveCall(const SubstepParams& P, int j, int lane8, unsigned gmask, int warmStart, float4* velLive, float4* angvelLive) {
    jointSolveOct<L1>(P.J, j, lane8, gmask, P.h, warmStart, P.kinematic, P.comInvMass, velLive, angvelLive);
}
__device__ __noinline__ void jointNgsSeqCall(const SubstepParams& P, int start, int count) {
    jointNgsSeq(P.J, start, count, P.kinematic, P.comInvMass, P.pseudoLin, P.pseudoAng);
}
__device__ __noinline__ void jointSolveSeqCall(const SubstepParams& P, int start, int count, int warmStart, float4* velLive, float4* angvelLive) {
    jointSolveSeq(P.J, start, count, P.h, warmStart, P.kinematic, P.comInvMass, velLive, angvelLive);
}
__device__ __noinline__ void contactSolveSeqCall(const SubstepParams& P, int start, int count, int useBias, int skipSoft, float4* velLive, float4* angvelLive) {
    L2Hints H = makeL2Hints();
    for (int i = 0; i < count; ++i) contactSolve<false>(P, start + i, useBias, skipSoft, velLive, angvelLive, H);
}

// joint row fill for every joint of the scene (makeConstraints + effective masses): independent of the joint colours
__global__ void __launch_bounds__(128) k_joint_fill(const __grid_constant__ SubstepParams P) {
    if (stepSkipped(P.counters)) return;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P.J.n) jointPrepOne(P.J, j, 0, P.kinematic, P.pos, P.quat, P.comInvMass, P.bodyRec, P.pseudoLin, P.pseudoAng);
}

// Everything of one substep that is ordered by colour: joint NGS pass (colours share bodies through the pseudo velocities),
// the solver iterations (contact colours, then joint colours), position integration and the relaxation pass.
//
// Two sweeps run the same colour sequence.  The DEVICE-WIDE sweep (group G of the run table) spreads every colour over the whole
// grid and separates colours with the grid barrier.  The LOCAL sweeps (islandsOn) give each group of small islands to one CTA,
// which walks its colours with __syncthreads only: a ragdoll batch or a field of separate little piles then pays two grid
// barriers per substep instead of one per colour and pass.  Islands share no dynamic body, so the interleaving is immaterial.
// L1LOCAL: the per-CTA sweeps may use L1-cached accesses -- true when everything they read was written before this kernel started or
// by their own CTA (k_substep_solve); false in the fused small-scene kernel, where other SMs rewrite the rows every substep.
// ONLYGROUP: the caller (k_step_solve_small, every constraint of the scene in small islands) walks its groups one at a time through the
// whole substep loop: this call then sweeps group `onlyGroup` alone, integrates that group's bodies (bodyOrder list) and never touches
// the grid barrier -- a batch of little scenes steps without a single device-wide synchronisation.
template <bool L1LOCAL, bool ONLYGROUP = false>
__device__ __forceinline__ void substepColoured(const SubstepParams& P, GridBarrier& bar, float4* velLive, float4* angvelLive, int* sRuns, int* sJoint, int onlyGroup = -1) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    const int ncol = P.counters[CNT_NCOLORS];
    const int lane = threadIdx.x & 31;
    const L2Hints H = makeL2Hints();
    const int G = P.G;
    unsigned long long tLocal = 0;
    auto stampLocal = [&](int kind) {       // profiling only: CTA 0 times its own local colour phases
        if (P.profNs && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (kind >= 0) { atomicAdd(&P.profNs[PROF_LOCAL + kind], t - tLocal); atomicAdd(&P.profNs[PROF_LOCAL + 3 + kind], 1ull); }
            tLocal = t;
        }
    };
    // sweep modes: (CTA-local barriers?, L1-cached accesses?)
    const SweepMode<true, L1LOCAL> LOCAL;
    const SweepMode<false, false> GLOBAL;

    // One contact pass over the colour runs `runs` (129 ints).  id / nthr: this thread's index and the thread count of the sweep.
    // Inside a colour the single-point manifolds come first (one thread each: two memory round trips), then the multi-point ones.
    // A colour that fits one round of the sweep lasts as long as its longest dependent chain, so its multi-point manifolds take
    // four lanes each (lane k = point k, all rows in one wave); over several rounds throughput matters and every manifold gets one
    // thread (the multi-point ones sit together at the end of the colour, so their longer path diverges in few warps).
    // one colour of a pass: `id` / `nthr` = this thread's index and the thread count of whoever sweeps it
    auto contactColour = [&](auto mode, const int* runs, int c, int id, int nthr, int useBias, int skipSoft) {
        constexpr bool l1 = decltype(mode)::l1;
        const int start = runs[2 * c], mid = runs[2 * c + 1], count = runs[2 * c + 2] - start;
        const int singles = mid - start, multis = count - singles;
        if (singles + 4 * multis <= nthr) {
            if (id < singles) contactSolve<l1>(P, start + id, useBias, skipSoft, velLive, angvelLive, H);
            int g = (nthr >> 2) - 1 - (id >> 2);      // from the far end: the low threads hold the singles
            if (g < multis) contactSolveQuad<l1>(P, mid + g, lane & 3, 0xFu << (lane & 28), useBias, skipSoft, velLive, angvelLive, H);
        } else {
            for (int i = id; i < count; i += nthr) contactSolve<l1>(P, start + i, useBias, skipSoft, velLive, angvelLive, H);
        }
    };
    // (Measured and dropped: the small trailing colours of a pile -- a 100 k-body pile has 18 colours, the last ten with a few hundred
    // manifolds in all -- swept by CTA 0 alone behind one grid barrier: 2.34 vs 2.21 ms/step at 100 k bodies; one CTA's L2 round
    // trips per colour are no shorter than the grid's.)
    auto contactPass = [&](auto mode, const int* runs, int id, int nthr, int useBias, int skipSoft) {
        constexpr bool local = decltype(mode)::local;
        for (int c = 0; c < ncol; ++c) {
            const int start = runs[2 * c], count = runs[2 * c + 2] - start;
            if (count <= 0) continue;
            if (c == PB_OVERFLOW_COLOR) {       // sequential bucket: manifolds may share bodies
                if (id == 0) contactSolveSeqCall(P, start, count, useBias, skipSoft, velLive, angvelLive);
            } else contactColour(mode, runs, c, id, nthr, useBias, skipSoft);
            if (local) { __syncthreads(); stampLocal(1); } else bar.sync(PH_CONTACT_PASS, c);
        }
    };
    // joint colour runs of a sweep: jr[c] .. jr[c + 1] index P.jointOrder (islandsOn) or are the joint slots themselves
    auto jointNgsPass = [&](auto mode, const int* jr, int id, int nthr) {
        constexpr bool local = decltype(mode)::local, l1 = decltype(mode)::l1;
        for (int c = 0; c < 8; ++c) {
            const int start = jr[c], count = jr[c + 1] - start;
            if (count <= 0) continue;
            for (int i = id; i < count; i += nthr) jointNgsCall<l1>(P, P.jointOrder ? P.jointOrder[start + i] : start + i);
            if (local) { __syncthreads(); stampLocal(0); } else bar.sync(PH_PREP);
        }
    };
    auto jointSolvePass = [&](auto mode, const int* jr, int id, int nthr, int warmStart) {
        constexpr bool local = decltype(mode)::local, l1 = decltype(mode)::l1;
        for (int c = 0; c < 8; ++c) {
            const int start = jr[c], count = jr[c + 1] - start;
            if (count <= 0) continue;
            for (int i = id >> 3; i < count; i += nthr >> 3)
                jointSolveCall<l1>(P, P.jointOrder ? P.jointOrder[start + i] : start + i, lane & 7, 0xFFu << (lane & 24), warmStart, velLive, angvelLive);
            if (local) { __syncthreads(); stampLocal(2); } else bar.sync(PH_JOINT_SOLVE);
        }
    };
    // a local group's tables into shared memory (the colour loops then cost no global round trip per colour)
    auto loadLocal = [&](int g) {
        __syncthreads();
        for (int i = threadIdx.x; i <= PB_KEY_COLORS; i += blockDim.x) sRuns[i] = P.keyStart[g * PB_KEY_COLORS + i];
        if (P.hasJoints && threadIdx.x <= 8) sJoint[threadIdx.x] = P.jointStart[g * 8 + threadIdx.x];
        __syncthreads();
    };

    if (ONLYGROUP) {
        const int g = onlyGroup;
        loadLocal(g);
        const bool empty = sRuns[PB_KEY_COLORS] == sRuns[0] && (!P.hasJoints || sJoint[8] == sJoint[0]);
        if (!empty) {
            if (P.hasJoints) jointNgsPass(LOCAL, sJoint, threadIdx.x, blockDim.x);
            for (int it = 0; it < P.iterations; ++it) {
                contactPass(LOCAL, sRuns, threadIdx.x, blockDim.x, 1, 0);
                if (P.hasJoints) jointSolvePass(LOCAL, sJoint, threadIdx.x, blockDim.x, it == 0);
            }
        }
        __syncthreads();
        for (int k = P.bodyStart[g] + threadIdx.x; k < P.bodyStart[g + 1]; k += blockDim.x) integrateX<true>(P, P.bodyOrder[k], velLive, angvelLive);
        __syncthreads();
        if (!empty) contactPass(LOCAL, sRuns, threadIdx.x, blockDim.x, 0, 1);
        return;
    }
    // ---- phase A: NGS pass of the joints, then the iterations --------------------------------------------------------------------
    if (P.islandsOn) {
        for (int g = blockIdx.x; g < G; g += gridDim.x) {
            loadLocal(g);
            stampLocal(-1);
            if (sRuns[PB_KEY_COLORS] == sRuns[0] && (!P.hasJoints || sJoint[8] == sJoint[0])) continue;     // empty group
            if (P.hasJoints) jointNgsPass(LOCAL, sJoint, threadIdx.x, blockDim.x);
            for (int it = 0; it < P.iterations; ++it) {
                contactPass(LOCAL, sRuns, threadIdx.x, blockDim.x, 1, 0);
                if (P.hasJoints) jointSolvePass(LOCAL, sJoint, threadIdx.x, blockDim.x, it == 0);
            }
        }
    }
    const int* gRuns = P.keyStart + G * PB_KEY_COLORS;
    const int* gJoint = P.islandsOn ? P.jointStart + G * 8 : P.jointColorStart;
    if (P.hasJoints) {
        // rows were filled by k_joint_fill; the NGS pass accumulates into per-body pseudo velocities, colour by colour
        jointNgsPass(GLOBAL, gJoint, tid, nth);
        int start = P.jointColorStart[8], count = P.jointColorStart[9] - start;
        if (count > 0) {                    // overflow bucket: the sequential NGS pass (always part of the device-wide sweep)
            if (tid == 0) jointNgsSeqCall(P, start, count);
            bar.sync(PH_PREP);
        }
    }
    for (int it = 0; it < P.iterations; ++it) {
        contactPass(GLOBAL, gRuns, tid, nth, 1, 0);
        if (P.hasJoints) {
            jointSolvePass(GLOBAL, gJoint, tid, nth, it == 0);
            int start = P.jointColorStart[8], count = P.jointColorStart[9] - start;
            if (count > 0) {
                if (tid == 0) jointSolveSeqCall(P, start, count, it == 0, velLive, angvelLive);
                bar.sync(PH_JOINT_SOLVE);
            }
        }
    }
    // the local sweeps end without a grid barrier: bodies are integrated by whichever thread owns their index
    if (P.islandsOn) bar.sync(PH_LOCAL);

    for (int i = tid; i < P.nDyn; i += nth) integrateX(P, i, velLive, angvelLive);
    bar.sync(PH_INTEGRATE_X);

    // ---- relaxation ----------------------------------------------------------------------------------------------------------------
    if (P.islandsOn) {
        for (int g = blockIdx.x; g < G; g += gridDim.x) {
            loadLocal(g);
            stampLocal(-1);
            contactPass(LOCAL, sRuns, threadIdx.x, blockDim.x, 0, 1);
        }
    }
    contactPass(GLOBAL, gRuns, tid, nth, 0, 1);
}


__global__ void __launch_bounds__(256, 3) k_substep_solve(const __grid_constant__ SubstepParams P) {
    __shared__ int sRuns[PB_KEY_COLORS + 1];
    __shared__ int sJoint[PB_JOINT_COLORS + 1];
    if (stepSkipped(P.counters)) return;       // uniform over the grid: nobody reaches a barrier
    GridBarrier bar;
    bar.counter = P.barrier; bar.target = 0; bar.profNs = P.profNs; bar.tPrev = 0; bar.cluster = 0;
    if (P.profNs && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(bar.tPrev));
    substepColoured<true>(P, bar, P.velB, P.angvelB, sRuns, sJoint);
    if (P.islandsOn && P.profNs) bar.sync(PH_LOCAL);     // profiling only: closes the local relaxation sweeps of every CTA
}

// Small scenes (a box pyramid, a few hundred ragdolls): the WHOLE substep loop of a step in one launch -- velocity integration,
// contact prep and joint row fill become grid-stride phases in front of the coloured part, and the velocity buffers swap inside the
// kernel.  A step then costs one launch instead of four or five per substep, which is most of what such a step costs; the register
// appetite of the prep phases (114) does not matter at this size.  Same device functions, same arithmetic, same order.
__device__ __forceinline__ void stepSolveSmall(const SubstepParams& P, int* sRuns, int* sJoint) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    if (stepSkipped(P.counters)) return;
    GridBarrier bar;
    bar.counter = P.barrier; bar.target = 0; bar.profNs = P.profNs; bar.tPrev = 0; bar.cluster = P.clusterBarrier;
    if (P.profNs && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(bar.tPrev));
    const int nManifolds = P.counters[CNT_MANIFOLDS];
    float4* vel = P.velA; float4* angvel = P.angvelA; float4* velLive = P.velB; float4* angvelLive = P.angvelB;
    // Every constraint sits in a small island (nothing in the device-wide group G, no overflow-bucket joints): groups are independent
    // for the WHOLE step, so each CTA takes its groups through all substeps on its own -- integration, prep, sweeps, relaxation --
    // with CTA barriers only.  (Decided from device data, uniformly over the grid; otherwise the grid-barrier form below runs.)
    if (P.islandsOn && P.bodyOrder) {
        const int* gRuns = P.keyStart + P.G * PB_KEY_COLORS;
        const bool globalEmpty = gRuns[PB_KEY_COLORS] == gRuns[0] && (!P.hasJoints || P.jointStart[P.G * 8 + 9] == P.jointStart[P.G * 8]);
        if (globalEmpty) {
            for (int g = blockIdx.x; g < P.G; g += gridDim.x) {
                const int b0 = P.bodyStart[g], b1 = P.bodyStart[g + 1];
                if (b0 == b1) continue;
                const int m0 = P.keyStart[g * PB_KEY_COLORS], m1 = P.keyStart[(g + 1) * PB_KEY_COLORS];
                const int j0 = P.hasJoints ? P.jointStart[g * 8] : 0, j1 = P.hasJoints ? P.jointStart[g * 8 + 8] : 0;
                float4* v = P.velA; float4* w = P.angvelA; float4* vL = P.velB; float4* wL = P.angvelB;
                for (int sub = 0; sub < P.substeps; ++sub) {
                    for (int k = b0 + threadIdx.x; k < b1; k += blockDim.x) integrateV<true>(P, P.bodyOrder[k], v, w, vL, wL);
                    __syncthreads();
                    for (int s = m0 + threadIdx.x; s < m1; s += blockDim.x) contactPrep<true>(P, s, v, w);
                    for (int k = j0 + threadIdx.x; k < j1; k += blockDim.x)
                        jointPrepOne<true>(P.J, P.jointOrder[k], 0, P.kinematic, P.pos, P.quat, P.comInvMass, P.bodyRec, P.pseudoLin, P.pseudoAng);
                    __syncthreads();
                    substepColoured<true, true>(P, bar, vL, wL, sRuns, sJoint, g);
                    __syncthreads();
                    float4* t = v; v = vL; vL = t;
                    t = w; w = wL; wL = t;
                }
            }
            return;
        }
    }
    for (int sub = 0; sub < P.substeps; ++sub) {
        for (int i = tid; i < P.nDyn; i += nth) integrateV(P, i, vel, angvel, velLive, angvelLive);
        bar.sync(PH_INTEGRATE_V);
        for (int s = tid; s < nManifolds; s += nth) contactPrep(P, s, vel, angvel);
        if (P.hasJoints) for (int j = tid; j < P.J.n; j += nth) jointPrepOne(P.J, j, 0, P.kinematic, P.pos, P.quat, P.comInvMass, P.bodyRec, P.pseudoLin, P.pseudoAng);
        bar.sync(PH_PREP);
        substepColoured<false>(P, bar, velLive, angvelLive, sRuns, sJoint);
        bar.sync(PH_LOCAL);       // the relaxation pass of every sweep is in before the next substep reads the velocities
        // write-back (Physecs.cpp:523-530): velocityTemp becomes the component velocity of the next substep
        float4* t = vel; vel = velLive; velLive = t;
        t = angvel; angvel = angvelLive; angvelLive = t;
    }
}

__global__ void __launch_bounds__(256, 3) k_step_solve_small(const __grid_constant__ SubstepParams P) {
    __shared__ int sRuns[PB_KEY_COLORS + 1];
    __shared__ int sJoint[PB_JOINT_COLORS + 1];
    stepSolveSmall(P, sRuns, sJoint);
}
// The same kernel for CTAs of 128 threads, three per SM: 168 registers per thread instead of 80.  At 80 the prep and joint routines
// spill ~2 KB per thread, and a batch of few little scenes -- one ragdoll per group: a step IS the dependent instruction chain of one
// CTA carrying one ragdoll through four substeps -- pays for every spill on that chain.  Used when the groups hold little work.
__global__ void __launch_bounds__(128, 3) k_step_solve_small_w(const __grid_constant__ SubstepParams P) {
    __shared__ int sRuns[PB_KEY_COLORS + 1];
    __shared__ int sJoint[PB_JOINT_COLORS + 1];
    stepSolveSmall(P, sRuns, sJoint);
}

int pb_solve(pb_ctx* ctx, float dt, int substeps, int iterations, float gravity) {
    const int nDyn = ctx->nDyn;
    if (nDyn == 0) return PB_OK;
    // manifold count as far as the host knows it: the previous step's, with headroom (no step collected yet: the arena capacity).
    // Shapes grids and picks between equivalent kernels, nothing else -- the kernels read the count from the device.
    const int workBound = ctx->rawHint < 0 ? ctx->caps.max_manifolds : std::min(ctx->caps.max_manifolds, ctx->rawHint + ctx->rawHint / 4 + 1024);
    if (!ctx->solveGrid) {
        int perSM = 0;
        PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_substep_solve, 256, 0));
        if (perSM < 1) return pb_fail(ctx, PB_ECUDA, "k_substep_solve does not fit on an SM");
        int perSM2 = 0;
        PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM2, k_step_solve_small, 256, 0));
        if (perSM2 < perSM) perSM = perSM2;       // both persistent kernels use the same co-resident grid
        if (perSM < 1) return pb_fail(ctx, PB_ECUDA, "k_step_solve_small does not fit on an SM");
        ctx->solveGrid = perSM * ctx->numSMs;
        // largest cluster of 256-thread CTAs of the whole-step kernel the device schedules (16 needs the non-portable opt-in, 8 is portable)
        if (ctx->clusterSize < 0) {
            ctx->clusterSize = 0;
            for (int c : { 16, 8 }) {
                if (c > 8 && cudaFuncSetAttribute(k_step_solve_small, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(c); cfg.blockDim = dim3(256);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, k_step_solve_small, &cfg) == cudaSuccess && n >= 1) { ctx->clusterSize = c; break; }
                cudaGetLastError();
            }
        }
        int rc = pb_alloc(ctx, &ctx->solveBarrier, 64); if (rc) return rc;
        rc = pb_alloc(ctx, &ctx->solveProfNs, PROF_WORDS); if (rc) return rc;
        PB_CUDA(ctx, cudaMemsetAsync(ctx->solveProfNs, 0, sizeof(unsigned long long) * PROF_WORDS, ctx->stream));
    }
    const int cur = ctx->curBuf;
    SubstepParams P{};
    P.nDyn = nDyn; P.substeps = substeps; P.iterations = iterations; P.h = dt / (float)substeps; P.g = gravity;
    P.counters = ctx->counters;
    P.kinematic = ctx->kinematic; P.comInvMass = ctx->comInvMass; P.invIL = ctx->invIL;
    P.pos = ctx->pos; P.quat = ctx->quat;
    P.bodyRec = ctx->bodyRec; P.pseudoLin = ctx->pseudoLin; P.pseudoAng = ctx->pseudoAng;
    P.cHead = ctx->cHead; P.cBodies = ctx->cBodies; P.cRowsT = ctx->cRowsT; P.cNormal = ctx->cNormal; P.cSoft = ctx->cSoft;
    P.cPointOfs = ctx->cPointOfsBuf[cur]; P.cNp = ctx->cNpBuf[cur]; P.pR0T = ctx->pR0T[cur]; P.pR1 = ctx->pR1;
    P.rowExtra = ctx->caps.max_manifolds;
    P.rowA = ctx->rowA; P.rowB = ctx->rowB; P.rowC = ctx->rowC; P.rowD = ctx->rowD; P.rowE = ctx->rowE; P.rowF = ctx->rowF; P.rowG = ctx->rowG; P.rowL = ctx->rowL;
    P.hasJoints = pb_joint_view(ctx, &P.J) ? 1 : 0;
    for (int c = 0; c <= PB_JOINT_COLORS; ++c) P.jointColorStart[c] = ctx->jointColorStart[c];
    P.keyStart = ctx->keyStart; P.G = ctx->islandGroups; P.islandsOn = ctx->islandsOn ? 1 : 0;
    P.jointOrder = (ctx->islandsOn && P.hasJoints) ? ctx->jointOrder : nullptr; P.jointStart = ctx->jointStart;
    P.bodyOrder = (ctx->islandsOn && ctx->bodyListsBuilt) ? ctx->bodyOrder : nullptr; P.bodyStart = ctx->bodyStart;
    P.barrier = ctx->solveBarrier;
    P.profNs = ctx->profile ? ctx->solveProfNs : nullptr;
    // persistent grid: co-resident by construction; small scenes use fewer CTAs so the barrier stays cheap
    long long work = std::max<long long>(std::max<long long>(nDyn, 4ll * workBound), 8ll * ctx->nJoints);
    int grid = (int)std::min<long long>(ctx->solveGrid, (work + 255) / 256);
    if (grid < 1) grid = 1;
    // islands on: one CTA per group, so that no CTA walks several groups' colour sequences back to back (a sweep costs its ~25 colour
    // phases of latency however few constraints the group holds); the two grid barriers per substep are cheap next to that
    if (ctx->islandsOn) grid = std::min(ctx->solveGrid, ctx->islandGroups);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    ctx->evSubCount = 0;
    // small scenes: the whole substep loop in one launch (k_step_solve_small); PB_FUSED=0 / 1 overrides the size rule
    const int fusedEnv = ctx->fusedMode;      // env PB_FUSED at context creation
    // (measured: equal or slightly ahead up to ~1 k bodies -- 64 ragdolls 0.78 vs 0.80 ms/step, 1 k-box pyramid 1.50 vs 1.52 -- and behind
    // from ~5 k bodies on, where the 3-CTA/SM register cap makes the prep phases spill; such steps are bound by the latency of their
    // ~30 colour phases per substep, not by launches)
    // ... except when every constraint sits in a small island (as far as the last collected step knows; the kernel decides from device
    // data and falls back to its grid-barrier form otherwise): then the groups go through the whole step independently, no grid barrier
    // at all, and the one-launch form wins up to batches of tens of thousands of bodies (4096 ragdoll scenes)
    const bool allLocal = ctx->islandsOn && ctx->bodyListsBuilt && (ctx->lastIslandTotal == 0 || ctx->lastIslandLocal == ctx->lastIslandTotal);
    const bool fused = fusedEnv >= 0 ? fusedEnv != 0 : ((nDyn <= 2048 && workBound <= 8192 && ctx->nJoints <= 4096) || (allLocal && nDyn <= ctx->fusedLocalMax));
    if (fused) {
        P.velA = ctx->vel; P.angvelA = ctx->angvel; P.velB = ctx->velLive; P.angvelB = ctx->angvelLive;
        int fgrid = ctx->islandsOn ? std::min(ctx->solveGrid, ctx->islandGroups) : grid;
        PB_CUDA(ctx, cudaMemsetAsync(ctx->solveBarrier, 0, sizeof(unsigned int), ctx->stream));
        void* args[] = { &P };
        ++ctx->launches;
        // few constraints per group (a batch of few little scenes): the 128-thread form, whose threads keep their state in registers
        const long long perGroup = ((long long)workBound + ctx->nJoints) / std::max(1, ctx->islandGroups);
        const bool narrow = allLocal && ctx->fusedNarrowMax > 0 && perGroup <= ctx->fusedNarrowMax && ctx->rawHint >= 0;
        // A small scene whose constraints form one pile (islands off: a box pyramid) sweeps its colours device-wide.  When a handful of
        // CTAs hold it, they are launched as ONE thread-block cluster: co-scheduled on a GPC by construction (no cooperative launch), and
        // the ~240 colour phases of a step meet at the hardware cluster barrier instead of an atomic counter in L2.
        const bool asCluster = !ctx->islandsOn && ctx->clusterSize > 0 && fgrid <= 4 * ctx->clusterSize;
        if (asCluster) {
            P.clusterBarrier = 1;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(ctx->clusterSize); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = ctx->clusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            PB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_step_solve_small, P));
        }
        else if (narrow) PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_step_solve_small_w, dim3(fgrid), dim3(128), args, 0, ctx->stream));
        else PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_step_solve_small, dim3(fgrid), dim3(256), args, 0, ctx->stream));
        if (substeps & 1) { std::swap(ctx->vel, ctx->velLive); std::swap(ctx->angvel, ctx->angvelLive); ++ctx->undoVelSwaps; }
    } else
    for (int sub = 0; sub < substeps; ++sub) {
        P.velA = ctx->vel; P.angvelA = ctx->angvel; P.velB = ctx->velLive; P.angvelB = ctx->angvelLive;
        ++ctx->launches, k_integrate_v<<<pb_grid(nDyn, 256), 256, 0, ctx->stream>>>(P);
        ++ctx->launches, k_contact_prep<<<pb_grid(workBound, 128), 128, 0, ctx->stream>>>(P);
        if (P.hasJoints) ++ctx->launches, k_joint_fill<<<pb_grid(ctx->nJoints, 128), 128, 0, ctx->stream>>>(P);
        PB_CUDA(ctx, cudaMemsetAsync(ctx->solveBarrier, 0, sizeof(unsigned int), ctx->stream));
        void* args[] = { &P };
        ++ctx->launches;
        // profiling (pb_set_profile): CUDA events around every k_substep_solve launch of the step (bench.py: roofline of the dominant kernel)
        const bool timeIt = ctx->profile && sub < 8;
        if (timeIt) {
            if (!ctx->evSub[2 * sub]) { cudaEventCreate(&ctx->evSub[2 * sub]); cudaEventCreate(&ctx->evSub[2 * sub + 1]); }
            cudaEventRecord(ctx->evSub[2 * sub], ctx->stream);
        }
        PB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_substep_solve, dim3(grid), dim3(256), args, 0, ctx->stream));
        if (timeIt) { cudaEventRecord(ctx->evSub[2 * sub + 1], ctx->stream); ctx->evSubCount = sub + 1; }
        // write-back (Physecs.cpp:523-530): velocityTemp becomes the component velocity of the next substep
        std::swap(ctx->vel, ctx->velLive);
        std::swap(ctx->angvel, ctx->angvelLive);
        ++ctx->undoVelSwaps;
    }
    cudaEventRecord(ctx->ev[6], ctx->stream);
    PB_CUDA(ctx, cudaGetLastError());
    return PB_OK;
}

// accumulated per-phase device time of the persistent kernel since profiling was switched on: ns[PH_KINDS], count[PH_KINDS]
int pb_solve_profile(pb_ctx* ctx, unsigned long long* out, bool reset) {
    if (!ctx->solveProfNs) { for (int i = 0; i < 2 * PH_KINDS; ++i) out[i] = 0; return PB_OK; }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(out, ctx->solveProfNs, sizeof(unsigned long long) * 2 * PH_KINDS, cudaMemcpyDeviceToHost));
    if (reset) PB_CUDA(ctx, cudaMemset(ctx->solveProfNs, 0, sizeof(unsigned long long) * PROF_WORDS));
    return PB_OK;
}

// per contact colour: accumulated ns and phase count (out[0..63] ns, out[64..127] counts); then CTA 0's local split (ns[3], count[3])
int pb_solve_profile_colors(pb_ctx* ctx, unsigned long long* out134) {
    for (int i = 0; i < 2 * PB_MAX_COLORS + 6; ++i) out134[i] = 0;
    if (!ctx->solveProfNs) return PB_OK;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaMemcpy(out134, ctx->solveProfNs + 2 * PH_KINDS, sizeof(unsigned long long) * (2 * PB_MAX_COLORS + 6), cudaMemcpyDeviceToHost));
    return PB_OK;
}

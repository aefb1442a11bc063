// Does a thread that re-reads a global word it (or a CTA-mate) just stored hit L1?  Dependent load -> add -> store -> __syncthreads
// chains on global memory (plain ld/st, ld.cg/st.cg) and on shared memory; cycles per iteration.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1_rw_bench l1_rw_bench.cu && ./l1_rw_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void chain(float4* g, long long* out, int iters) {
    __shared__ float4 s[256];
    float4* p = MODE == 2 ? s : g + blockIdx.x * 256;
    const int me = threadIdx.x, other = (threadIdx.x + 1) & 127;      // read a neighbour's word: written by a CTA-mate last round
    if (MODE == 2) s[me] = make_float4(1, 2, 3, 4);
    __syncthreads();
    long long t0 = clock64();
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        float4 v;
        if (MODE == 1) v = __ldcg(&p[other]); else v = p[other];
        acc += v.x;
        v.x = acc * 0.5f + 1.f;
        __syncthreads();
        if (MODE == 1) __stcg(&p[me], v); else p[me] = v;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = (long long)acc; }
}
int main() {
    float4* g; long long* out; cudaMalloc(&g, 4096 * 256 * sizeof(float4)); cudaMemset(g, 0, 4096 * 256 * sizeof(float4)); cudaMallocManaged(&out, 8192 * sizeof(long long));
    const int iters = 2000;
    const char* names[3] = { "global, plain ld / st (L1 path)", "global, ld.cg / st.cg (L2)", "shared memory" };
    for (int blocks : { 1, 444 }) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) chain<0><<<blocks, 128>>>(g, out, iters);
                if (mode == 1) chain<1><<<blocks, 128>>>(g, out, iters);
                if (mode == 2) chain<2><<<blocks, 128>>>(g, out, iters);
                cudaDeviceSynchronize();
            }
            printf("%3d CTAs  %-34s %8.1f cycles / round\n", blocks, names[mode], (double)out[0] / iters);
        }
    }
    return 0;
}

// Microbenchmark: latency of the device-wide barrier used by k_substep_solve, in a few variants, plus the cost of a minimal
// "phase" (dependent load -> dependent load -> store -> barrier).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int VARIANT>
__device__ __forceinline__ void gridBarrier(unsigned int* counter, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        if (VARIANT == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
            unsigned int seen;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
        } else if (VARIANT == 1) {
            __threadfence();
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
            unsigned int seen;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
            __threadfence();
        } else {
            // atom returns the old value: the last arriver knows it is last without polling once more
            unsigned int old;
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
            if (old + 1 < target) {
                unsigned int seen;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
            } else __threadfence();
        }
    }
    __syncthreads();
}

template <int VARIANT>
__global__ void k_barriers(unsigned int* counter, int iters, unsigned long long* out) {
    unsigned int target = 0;
    unsigned long long t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) gridBarrier<VARIANT>(counter, target);
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}

__global__ void k_cg_barriers(int iters, unsigned long long* out) {
    cg::grid_group g = cg::this_grid();
    unsigned long long t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}

// a minimal phase: n work items; item i loads idx[i] (streamed), then gathers val[idx] (dependent), adds, stores back; barrier.
__global__ void k_phases(unsigned int* counter, int iters, int n, const int* __restrict__ idx, float4* val, const float4* __restrict__ rows, unsigned long long* out) {
    unsigned int target = 0;
    unsigned long long t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        for (int i = tid; i < n; i += nth) {
            int b = __ldcg(&idx[i]);
            float4 r = __ldcg(&rows[(size_t)it % 4 * n + i]);
            float4 v = __ldcg(&val[b]);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            __stcg(&val[b], v);
        }
        gridBarrier<0>(counter, target);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    unsigned int* counter; unsigned long long* out;
    cudaMalloc(&counter, 256); cudaMalloc(&out, 64);
    const int iters = 2000;
    auto run = [&](const char* name, void* fn, int grid, int block, void** args) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(counter, 0, 4);
            cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(block), args, 0, 0);
            if (e != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(e)); return; }
            cudaDeviceSynchronize();
        }
        unsigned long long ns; cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost);
        printf("%-44s grid %4d x %4d : %7.3f us per barrier/phase\n", name, grid, block, ns / 1e3 / iters);
    };
    int it = iters;
    for (int cfg = 0; cfg < 4; ++cfg) {
        int block = cfg == 0 ? 256 : cfg == 1 ? 256 : cfg == 2 ? 768 : 1024;
        int per = cfg == 0 ? 3 : 1;
        int grid = sms * per;
        void* a0[] = { &counter, &it, &out };
        run("red.release + ld.acquire poll", (void*)k_barriers<0>, grid, block, a0);
        run("fence + relaxed red/poll + fence", (void*)k_barriers<1>, grid, block, a0);
        run("atom.release (last arriver skips poll)", (void*)k_barriers<2>, grid, block, a0);
        void* a1[] = { &it, &out };
        run("cooperative_groups grid.sync", (void*)k_cg_barriers, grid, block, a1);
    }
    // phases with work
    int nmax = 1 << 20;
    int* idx; float4* val; float4* rows;
    cudaMalloc(&idx, sizeof(int) * nmax); cudaMalloc(&val, sizeof(float4) * nmax); cudaMalloc(&rows, sizeof(float4) * 4 * (size_t)nmax);
    int* h = new int[nmax];
    unsigned int s = 12345;
    for (int i = 0; i < nmax; ++i) h[i] = i;
    for (int i = nmax - 1; i > 0; --i) { s = s * 1664525u + 1013904223u; int j = s % (i + 1); int t = h[i]; h[i] = h[j]; h[j] = t; }   // a permutation: no two items share a target
    cudaMemcpy(idx, h, sizeof(int) * nmax, cudaMemcpyHostToDevice);
    cudaMemset(val, 0, sizeof(float4) * nmax); cudaMemset(rows, 0, sizeof(float4) * 4 * (size_t)nmax);
    for (int n : { 1, 32, 1024, 16384, 65536, 113664, 262144, 1048576 }) {
        int grid = sms * 3, block = 256;
        void* a2[] = { &counter, &it, &n, &idx, &val, &rows, &out };
        char name[64]; snprintf(name, 64, "phase: %d items (load->gather->store)", n);
        run(name, (void*)k_phases, grid, block, a2);
    }
    return 0;
}

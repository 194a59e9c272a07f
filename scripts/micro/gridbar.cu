// micro-benchmark: cost of one grid-wide barrier, cooperative groups vs a hand-written
// sense-reversing barrier (one red.release per block, thread 0 spins on ld.acquire)
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

__global__ void k_cg(int iters, unsigned long long *sink)
{
    cg::grid_group grid = cg::this_grid();
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) { acc += i; grid.sync(); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

struct Bar { unsigned int count; unsigned int gen; };

__device__ __forceinline__ void bar_sync(Bar *b, unsigned int nblocks, unsigned int &my_gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int target = my_gen + 1;
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&b->count) : "memory");
        if (prev == nblocks - 1) {
            b->count = 0;                       // ordered before the release below
            asm volatile("st.release.gpu.u32 [%0], %1;" :: "l"(&b->gen), "r"(target) : "memory");
        } else {
            unsigned int g;
            do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(&b->gen) : "memory"); } while (g != target);
        }
    }
    my_gen++;
    __syncthreads();
}

__global__ void k_own(int iters, Bar *b, unsigned long long *sink)
{
    unsigned int gen = 0;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) { acc += i; bar_sync(b, gridDim.x, gen); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

int main(int argc, char **argv)
{
    const int iters = 20000;
    unsigned long long *sink; Bar *bar;
    cudaMalloc(&sink, 8); cudaMalloc(&bar, sizeof(Bar)); 
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int threads : {128, 448, 512}) for (int blocks : {148, 296}) {
        float ms;
        int it = iters;
        void *args1[] = {&it, &sink};
        cudaLaunchCooperativeKernel((void *)k_cg, dim3(blocks), dim3(threads), args1, 0, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void *)k_cg, dim3(blocks), dim3(threads), args1, 0, 0);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("cg   blocks %3d threads %3d: %.3f us per barrier (%s)\n", blocks, threads, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        cudaMemset(bar, 0, sizeof(Bar));
        void *args2[] = {&it, &bar, &sink};
        cudaLaunchCooperativeKernel((void *)k_own, dim3(blocks), dim3(threads), args2, 0, 0);
        cudaDeviceSynchronize();
        cudaMemset(bar, 0, sizeof(Bar));
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void *)k_own, dim3(blocks), dim3(threads), args2, 0, 0);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("own  blocks %3d threads %3d: %.3f us per barrier (%s)\n", blocks, threads, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

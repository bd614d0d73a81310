// Micro-benchmark (diagnostic): hand-off latency through mbarriers between two warps of one CTA, with try_wait (default
// and with a suspend-time hint) and test_wait polling, and with bystander warps that block on a barrier which never completes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mbar_pingpong tools/ubench/mbar_pingpong.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
template <int MODE> __device__ __forceinline__ void wait(uint64_t *b, uint32_t par, uint32_t hint) {
    uint32_t ok;
    do {
        if (MODE == 0)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
        else if (MODE == 1)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par), "r"(hint) : "memory");
    } while (!ok);
}

// warp 0 <-> warp 1 ping-pong; warps 2.. are bystanders: BY = 0 none, 1 blocked in try_wait, 2 spinning on test_wait
template <int MODE, int BY> __global__ void k_pingpong(int iters, uint32_t hint, long long *out, int done_wait_reps) {
    __shared__ uint64_t bars[4];
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0 && lane == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            mbar_arrive(&bars[0]);
            wait<MODE>(&bars[1], (uint32_t)i & 1u, hint);
        }
        const long long t1 = clock64();
        out[0] = (t1 - t0) / iters;  // cycles per round trip (two hand-offs)
        // cost of waiting on a barrier whose phase completed long ago
        const long long t2 = clock64();
        for (int i = 0; i < done_wait_reps; ++i) wait<MODE>(&bars[1], (uint32_t)(iters - 1) & 1u, hint);
        out[1] = (clock64() - t2) / done_wait_reps;
        stop = 1;
        mbar_arrive(&bars[2]);
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            wait<MODE>(&bars[0], (uint32_t)i & 1u, hint);
            mbar_arrive(&bars[1]);
        }
    } else if (warp >= 2 && lane == 0 && BY) {
        if (BY == 1) wait<0>(&bars[2], 0, 0);      // sleeps in try_wait until the very end
        else while (!stop) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bars[3])), "r"(0u) : "memory"); }
    }
}

template <int MODE, int BY> void run(const char *name, int warps, uint32_t hint) {
    long long *out, h[2];
    cudaMalloc(&out, 16);
    k_pingpong<MODE, BY><<<1, warps * 32>>>(2000, hint, out, 200);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-44s warps %2d hint %6u : round trip %5lld clk (%4lld per hand-off), wait on completed barrier %4lld clk  [%s]\n", name, warps, hint, h[0], h[0] / 2, h[1],
           cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    run<0, 0>("try_wait, no bystanders", 2, 0);
    run<1, 0>("test_wait poll, no bystanders", 2, 0);
    run<2, 0>("try_wait + hint, no bystanders", 2, 20);
    run<2, 0>("try_wait + hint, no bystanders", 2, 1000);
    run<0, 1>("try_wait, 14 bystanders asleep in try_wait", 16, 0);
    run<1, 1>("test_wait poll, 14 bystanders asleep", 16, 0);
    run<0, 2>("try_wait, 14 bystanders polling test_wait", 16, 0);
    run<1, 2>("test_wait poll, 14 bystanders polling", 16, 0);
    run<0, 2>("try_wait, 6 bystanders polling test_wait", 8, 0);
    run<0, 1>("try_wait, 18 bystanders asleep in try_wait", 20, 0);
    return 0;
}

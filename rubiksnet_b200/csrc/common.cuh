// Shared helpers for librubiks_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <utility>

#include "../../include/rubiks_b200.h"

namespace rb {

// ---- error reporting -------------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local int g_last_impl;
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_forced_impl;
extern std::atomic<int> g_pdl;  // programmatic dependent launch on (default) / off: rb_set_dependent_launch()

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// call right after every <<<>>>: counts the launch and surfaces launch-configuration errors
// (the reference never checks: SURVEY.md 2.2)
inline int launched(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return RB_OK;
}

inline size_t dtype_size(int dt) {
    switch (dt) {
        case RB_F32: return 4;
        case RB_F64: return 8;
        case RB_F16: return 2;
        case RB_BF16: return 2;
    }
    return 0;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();
int launch_fence(cudaStream_t s);  // abi.cu

// ---- programmatic dependent launch ---------------------------------------------------------------
// A training step is ~1 400 launches of 10-60 us each, so the launch gap, the drain of the previous grid and every
// kernel's own prologue (barrier init, TMEM allocation, the weight block of the 1x1 convs) are a measurable share of it.
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and follows ONE rule:
//     prologue that touches no global memory written by an earlier launch  ->  pdl_wait()  ->  pdl_trigger()  ->  body
// pdl_wait() (griddepcontrol.wait) returns once the previous kernel in the stream has completed and its writes are
// visible; triggering only AFTER the wait bounds the overlap to one kernel: when a grid passes its wait, everything
// older than its predecessor is complete, so a prologue may read data produced two or more launches earlier (the packed
// weights of a step: rb_pw_weight_pack* launch a fence kernel behind the pack) and never anything newer.  No global
// writes happen before the wait, so there are no write-after-read hazards against the predecessor either.  Kernels of
// other libraries (cuDNN conv1, the optimizer) are launched without the attribute and serialize as usual.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_wait();
    pdl_trigger();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- element conversion ----------------------------------------------------------------------
template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename A, typename T> __device__ __forceinline__ A ld(const T *p) { return (A)(*p); }
template <> __device__ __forceinline__ float ld<float, __half>(const __half *p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld<float, __nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(*p);
}

template <typename T, typename A> __device__ __forceinline__ T cvt(A v) { return (T)v; }
template <> __device__ __forceinline__ __half cvt<__half, float>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16, float>(float v) {
    return __float2bfloat16_rn(v);
}

// shift parameters may be stored in a different dtype than the activations (fp32 master copy with
// bf16 activations); dt is block-uniform so the switch does not diverge
template <typename A> __device__ __forceinline__ A ld_param(const void *p, int dt, int i) {
    switch (dt) {
        case RB_F32: return (A)((const float *)p)[i];
        case RB_F64: return (A)((const double *)p)[i];
        case RB_F16: return (A)__half2float(((const __half *)p)[i]);
        default: return (A)__bfloat162float(((const __nv_bfloat16 *)p)[i]);
    }
}
template <typename A> __device__ __forceinline__ void st_param(void *p, int dt, int i, A v) {
    switch (dt) {
        case RB_F32: ((float *)p)[i] = (float)v; break;
        case RB_F64: ((double *)p)[i] = (double)v; break;
        case RB_F16: ((__half *)p)[i] = __float2half_rn((float)v); break;
        default: ((__nv_bfloat16 *)p)[i] = __float2bfloat16_rn((float)v); break;
    }
}

// floorf() on the shift even in double: cuda_src/rubiks3d_kernels.cu:65-69
template <typename A> __device__ __forceinline__ int floor3d(A s) { return (int)floorf((float)s); }
// cuda_src/rubiks2d_kernels.cu:69-73
template <typename A> __device__ __forceinline__ int floor_fast(A x) {
    int ix = (int)x;
    return ix - (x < (A)ix);
}
// cuda_src/rubiks2d_kernels.cu:76-82
template <typename A> __device__ __forceinline__ int round_fast(A x) {
    return (x < (A)0) ? (int)(x - (A)0.5f) : (int)(x + (A)0.5f);
}

template <typename A> __device__ __forceinline__ A warp_sum(A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Geom3 {
    int N, T, C, H, W;
    int To, Ho, Wo;
    int sT, sH, sW, pT, pH, pW;
};
struct Geom2 {
    int N, C, H, W;
    int Ho, Wo;
    int sH, sW, pH, pW;
};

#define RB_DISPATCH_DTYPE(dt, ...)                                   \
    switch (dt) {                                                    \
        case RB_F32: { using T = float; __VA_ARGS__; } break;          \
        case RB_F64: { using T = double; __VA_ARGS__; } break;         \
        case RB_F16: { using T = __half; __VA_ARGS__; } break;         \
        case RB_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
        default: return fail(RB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dt); \
    }

}  // namespace rb

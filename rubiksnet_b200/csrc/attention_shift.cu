// AttentionShift temporal mix (rubiksnet/attention_shift.py:32-39).  The reference materialises a
// [C*H*W,1,3] weight with repeat_interleave and runs a depth-wise F.conv1d with up to 903 168 groups
// on a transposed view followed by .contiguous(): several passes over HBM.  Here every (n,c,p) column
// of T frames is streamed once through registers: x is read once and out written once.
#include "common.cuh"

namespace rb {

static constexpr int kAThreads = 128;

// x/out: [N, T, C, HW]; taps fp32 [C,3]
template <typename T>
__global__ void __launch_bounds__(kAThreads)
k_attn_fwd(const T *__restrict__ x, const float *__restrict__ taps, T *__restrict__ out, int Tn, int C,
           int HW) {
    pdl_sync();
    const int p = blockIdx.x * kAThreads + threadIdx.x;
    const int c = blockIdx.y, n = blockIdx.z;
    if (p >= HW) return;
    const float a0 = taps[c * 3 + 0], a1 = taps[c * 3 + 1], a2 = taps[c * 3 + 2];
    const int64_t fs = (int64_t)C * HW;
    const T *xp = x + ((int64_t)n * Tn * C + c) * HW + p;
    T *op = out + ((int64_t)n * Tn * C + c) * HW + p;
    float prev = 0.f, cur = ld<float, T>(xp);
    for (int t = 0; t < Tn; ++t) {
        const float nxt = (t + 1 < Tn) ? ld<float, T>(xp + (t + 1) * fs) : 0.f;
        op[t * fs] = cvt<T, float>(a0 * prev + a1 * cur + a2 * nxt);
        prev = cur;
        cur = nxt;
    }
}

// gx[t] = a0 g[t+1] + a1 g[t] + a2 g[t-1];  gtaps[c,k] = sum g[t] x[t+k-1]
template <typename T>
__global__ void __launch_bounds__(kAThreads)
k_attn_bwd(const T *__restrict__ x, const float *__restrict__ taps, const T *__restrict__ og,
           T *__restrict__ gx, float *__restrict__ partial, int Tn, int C, int HW, int want_gx,
           int want_gt) {
    pdl_sync();
    const int p = blockIdx.x * kAThreads + threadIdx.x;
    const int c = blockIdx.y, n = blockIdx.z;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (p < HW) {
        const float a0 = taps[c * 3 + 0], a1 = taps[c * 3 + 1], a2 = taps[c * 3 + 2];
        const int64_t fs = (int64_t)C * HW;
        const int64_t base = ((int64_t)n * Tn * C + c) * HW + p;
        const T *xp = x + base;
        const T *gp = og + base;
        float gprev = 0.f, gcur = ld<float, T>(gp);
        float xprev = 0.f, xcur = want_gt ? ld<float, T>(xp) : 0.f;
        for (int t = 0; t < Tn; ++t) {
            const bool more = t + 1 < Tn;
            const float gnxt = more ? ld<float, T>(gp + (t + 1) * fs) : 0.f;
            if (want_gx) gx[base + t * fs] = cvt<T, float>(a0 * gnxt + a1 * gcur + a2 * gprev);
            if (want_gt) {
                const float xnxt = more ? ld<float, T>(xp + (t + 1) * fs) : 0.f;
                s0 += gcur * xprev;
                s1 += gcur * xcur;
                s2 += gcur * xnxt;
                xprev = xcur;
                xcur = xnxt;
            }
            gprev = gcur;
            gcur = gnxt;
        }
    }
    if (!want_gt) return;
    __shared__ float red[3][kAThreads / 32];
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = s0;
        red[1][warp] = s1;
        red[2][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAThreads / 32; ++w) s += red[threadIdx.x][w];
        const int parts = gridDim.x * gridDim.z;
        const int part = blockIdx.z * gridDim.x + blockIdx.x;
        partial[((int64_t)c * parts + part) * 3 + threadIdx.x] = s;
    }
}

__global__ void k_attn_finalize(const float *__restrict__ partial, int parts, float *gtaps, int C) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0, s2 = 0;
    const float *p = partial + (int64_t)c * parts * 3;
    for (int i = lane; i < parts; i += 32) {
        s0 += p[i * 3 + 0];
        s1 += p[i * 3 + 1];
        s2 += p[i * 3 + 2];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        gtaps[c * 3 + 0] = (float)s0;
        gtaps[c * 3 + 1] = (float)s1;
        gtaps[c * 3 + 2] = (float)s2;
    }
}

}  // namespace rb

using namespace rb;

// NB: the frame count is called Tn below because RB_DISPATCH_DTYPE binds the element type to `T`.
extern "C" int rb_attention_shift_forward(const void *x, const float *taps, void *out, int dtype,
                                          int N, int Tn, int C, int HW, void *stream) {
    if (N < 0 || Tn < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0) return fail(RB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if ((int64_t)N * Tn * C * HW == 0) return RB_OK;
    if (!x || !taps || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (C > 65535 || N > 65535) return fail(RB_ERR_UNSUPPORTED, "attention shift: C or N > 65535");
    dim3 grid(cdiv(HW, kAThreads), C, N);
    cudaStream_t s = (cudaStream_t)stream;
    RB_DISPATCH_DTYPE(dtype, (launch_kernel(k_attn_fwd<T>, dim3(grid), dim3(kAThreads), 0, s, (const T *)x, taps, (T *)out,
                                                                      Tn, C, HW)));
    return launched("k_attn_fwd");
}

extern "C" size_t rb_attention_shift_backward_workspace_bytes(int N, int Tn, int C, int HW) {
    (void)Tn;
    if (N <= 0 || C <= 0 || HW <= 0) return 0;
    return (size_t)C * N * cdiv(HW, kAThreads) * 3 * sizeof(float);
}

extern "C" int rb_attention_shift_backward(const void *x, const float *taps, const void *out_grad,
                                           void *x_grad, float *taps_grad, int dtype, int N, int Tn,
                                           int C, int HW, void *workspace, size_t workspace_bytes,
                                           void *stream) {
    if (N < 0 || Tn < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0) return fail(RB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    if ((int64_t)N * Tn * C * HW == 0) {
        if (taps_grad && C > 0) cudaMemsetAsync(taps_grad, 0, (size_t)C * 3 * sizeof(float), s);
        return RB_OK;
    }
    if (!x || !taps || !out_grad) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (!x_grad && !taps_grad) return RB_OK;
    if (C > 65535 || N > 65535) return fail(RB_ERR_UNSUPPORTED, "attention shift: C or N > 65535");
    const size_t need = rb_attention_shift_backward_workspace_bytes(N, Tn, C, HW);
    if (taps_grad && (!workspace || workspace_bytes < need))
        return fail(RB_ERR_WORKSPACE, "attention shift backward needs %zu workspace bytes, got %zu",
                    need, workspace_bytes);
    dim3 grid(cdiv(HW, kAThreads), C, N);
    RB_DISPATCH_DTYPE(dtype, (launch_kernel(k_attn_bwd<T>, dim3(grid), dim3(kAThreads), 0, s, 
                                 (const T *)x, taps, (const T *)out_grad, (T *)x_grad,
                                 (float *)workspace, Tn, C, HW, x_grad != nullptr,
                                 taps_grad != nullptr)));
    int rc = launched("k_attn_bwd");
    if (rc || !taps_grad) return rc;
    const int warps = 4;
    launch_kernel(k_attn_finalize, dim3(cdiv(C, warps)), dim3(warps * 32), 0, s, (const float *)workspace,
                                                          (int)(grid.x * grid.z), taps_grad, C);
    return launched("k_attn_finalize");
}

// AttentionShift temporal mix (rubiksnet/attention_shift.py:32-39).  The reference materialises a
// [C*H*W,1,3] weight with repeat_interleave and runs a depth-wise F.conv1d with up to 903 168 groups
// on a transposed view followed by .contiguous(): several passes over HBM.  Here every (n,c,p) column
// of T frames is streamed once through registers: x is read once and out written once.
//
// AFF variants: the input is relu(x * scale[c] + bias[c]) rounded to the element type -- bn1 -> relu of the attention-quantized
// block (rubiksnet/backbone.py:123-125 with models.py:100-104) folded into the load, so that the normalised tensor is never
// written: the value is computed exactly as k_bn_apply would store it (one FMA, max, one rounding), hence bit-identical.
#include "common.cuh"

namespace rb {

static constexpr int kAThreads = 128;

template <typename T, bool AFF> __device__ __forceinline__ float a_in(const T *p, float sc, float bi) {
    const float f = ld<float, T>(p);
    if (!AFF) return f;
    const T r = cvt<T, float>(fmaxf(f * sc + bi, 0.f));
    return ld<float, T>(&r);
}

template <typename T, int V> struct alignas(sizeof(T) * V) APack { T v[V]; };

// Thread layout shared by both kernels: a thread owns V consecutive pixels of one channel plane (one 2V- / 4V-byte vector
// per frame) and streams the T frames of its clip through registers.  Planes with fewer than 128 vectors (14x14, 7x7) share a
// CTA: CG = 128 / (vectors per plane) channels per block.  grid = (blocks per plane, channel groups, clips).
struct ASlot {
    int c, p, cl;   // channel, first pixel, channel index inside the CTA
    bool active;
};
template <int V> __device__ __forceinline__ ASlot a_slot(int C, int HW, int vpc, int CG) {
    ASlot s;
    if (CG > 1) {
        s.cl = threadIdx.x / vpc;
        s.p = (threadIdx.x - s.cl * vpc) * V;
        s.c = blockIdx.y * CG + s.cl;
        s.active = s.cl < CG && s.c < C;
    } else {
        s.cl = 0;
        s.p = (blockIdx.x * kAThreads + threadIdx.x) * V;
        s.c = blockIdx.y;
        s.active = s.p < HW;
    }
    return s;
}

// x/out: [N, T, C, HW]; taps fp32 [C,3]
template <typename T, int V, bool AFF>
__global__ void __launch_bounds__(kAThreads)
k_attn_fwd(const T *__restrict__ x, const float *__restrict__ in_sb, const float *__restrict__ taps, T *__restrict__ out, int Tn, int C,
           int HW, int vpc, int CG) {
    pdl_sync();
    const ASlot sl = a_slot<V>(C, HW, vpc, CG);
    if (!sl.active) return;
    const int c = sl.c, n = blockIdx.z;
    const float a0 = taps[c * 3 + 0], a1 = taps[c * 3 + 1], a2 = taps[c * 3 + 2];
    const float sc = AFF ? in_sb[2 * c] : 1.f, bi = AFF ? in_sb[2 * c + 1] : 0.f;
    const int64_t fs = (int64_t)C * HW;
    const int64_t base = ((int64_t)n * Tn * C + c) * HW + sl.p;
    using P = APack<T, V>;
    float prev[V], cur[V];
    {
        const P v0 = *reinterpret_cast<const P *>(x + base);
#pragma unroll
        for (int i = 0; i < V; ++i) { prev[i] = 0.f; cur[i] = a_in<T, AFF>(&v0.v[i], sc, bi); }
    }
    for (int t = 0; t < Tn; ++t) {
        float nxt[V];
        if (t + 1 < Tn) {
            const P vn = *reinterpret_cast<const P *>(x + base + (t + 1) * fs);
#pragma unroll
            for (int i = 0; i < V; ++i) nxt[i] = a_in<T, AFF>(&vn.v[i], sc, bi);
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) nxt[i] = 0.f;
        }
        P o;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            o.v[i] = cvt<T, float>(a0 * prev[i] + a1 * cur[i] + a2 * nxt[i]);
            prev[i] = cur[i];
            cur[i] = nxt[i];
        }
        *reinterpret_cast<P *>(out + base + t * fs) = o;
    }
}

// gx[t] = a0 g[t+1] + a1 g[t] + a2 g[t-1];  gtaps[c,k] = sum g[t] x[t+k-1]
template <typename T, int V, bool AFF>
__global__ void __launch_bounds__(kAThreads)
k_attn_bwd(const T *__restrict__ x, const float *__restrict__ in_sb, const float *__restrict__ taps, const T *__restrict__ og,
           T *__restrict__ gx, float *__restrict__ partial, int Tn, int C, int HW, int vpc, int CG, int want_gx, int want_gt) {
    pdl_sync();
    const ASlot sl = a_slot<V>(C, HW, vpc, CG);
    const int n = blockIdx.z;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    using P = APack<T, V>;
    if (sl.active) {
        const int c = sl.c;
        const float a0 = taps[c * 3 + 0], a1 = taps[c * 3 + 1], a2 = taps[c * 3 + 2];
        const float sc = AFF ? in_sb[2 * c] : 1.f, bi = AFF ? in_sb[2 * c + 1] : 0.f;
        const int64_t fs = (int64_t)C * HW;
        const int64_t base = ((int64_t)n * Tn * C + c) * HW + sl.p;
        float gprev[V], gcur[V], xprev[V], xcur[V];
        {
            const P g0 = *reinterpret_cast<const P *>(og + base);
#pragma unroll
            for (int i = 0; i < V; ++i) { gprev[i] = 0.f; gcur[i] = ld<float, T>(&g0.v[i]); xprev[i] = 0.f; xcur[i] = 0.f; }
            if (want_gt) {
                const P x0 = *reinterpret_cast<const P *>(x + base);
#pragma unroll
                for (int i = 0; i < V; ++i) xcur[i] = a_in<T, AFF>(&x0.v[i], sc, bi);
            }
        }
        for (int t = 0; t < Tn; ++t) {
            const bool more = t + 1 < Tn;
            float gnxt[V], xnxt[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { gnxt[i] = 0.f; xnxt[i] = 0.f; }
            if (more) {
                const P gn = *reinterpret_cast<const P *>(og + base + (t + 1) * fs);
#pragma unroll
                for (int i = 0; i < V; ++i) gnxt[i] = ld<float, T>(&gn.v[i]);
                if (want_gt) {
                    const P xn = *reinterpret_cast<const P *>(x + base + (t + 1) * fs);
#pragma unroll
                    for (int i = 0; i < V; ++i) xnxt[i] = a_in<T, AFF>(&xn.v[i], sc, bi);
                }
            }
            if (want_gx) {
                P o;
#pragma unroll
                for (int i = 0; i < V; ++i) o.v[i] = cvt<T, float>(a0 * gnxt[i] + a1 * gcur[i] + a2 * gprev[i]);
                *reinterpret_cast<P *>(gx + base + t * fs) = o;
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                if (want_gt) {
                    s0 += gcur[i] * xprev[i];
                    s1 += gcur[i] * xcur[i];
                    s2 += gcur[i] * xnxt[i];
                    xprev[i] = xcur[i];
                    xcur[i] = xnxt[i];
                }
                gprev[i] = gcur[i];
                gcur[i] = gnxt[i];
            }
        }
    }
    if (!want_gt) return;
    // deterministic per-channel sums: every thread parks its three sums, thread (channel, axis) adds its channel's slots in order
    __shared__ float red[kAThreads][3];
    red[threadIdx.x][0] = s0;
    red[threadIdx.x][1] = s1;
    red[threadIdx.x][2] = s2;
    __syncthreads();
    if (threadIdx.x < CG * 3) {
        const int l = threadIdx.x / 3, ax = threadIdx.x - l * 3;
        const int c = CG > 1 ? blockIdx.y * CG + l : blockIdx.y;
        if (c < C) {
            const int lo = CG > 1 ? l * vpc : 0, hi = CG > 1 ? lo + vpc : kAThreads;
            float s = 0.f;
            for (int i = lo; i < hi; ++i) s += red[i][ax];
            const int parts = gridDim.x * gridDim.z;
            const int part = blockIdx.z * gridDim.x + blockIdx.x;
            partial[((int64_t)c * parts + part) * 3 + ax] = s;
        }
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

namespace {

struct APlan {
    int V, vpc, CG;
    dim3 grid;
};

// widest vector (elements) that divides the plane and that every base pointer is aligned for
APlan a_plan(int dtype, int N, int C, int HW, const void *p0, const void *p1, const void *p2) {
    const size_t es = dtype_size(dtype);
    const uintptr_t al = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2);
    APlan p;
    p.V = (int)(16 / es);
    if (p.V > 8) p.V = 8;
    while (p.V > 1 && (HW % p.V != 0 || (al & (uintptr_t)(p.V * es - 1)) != 0)) p.V >>= 1;
    p.vpc = HW / p.V;
    p.CG = p.vpc >= kAThreads ? 1 : kAThreads / p.vpc;
    p.grid = p.CG > 1 ? dim3(1, (unsigned)cdiv(C, p.CG), (unsigned)N) : dim3((unsigned)cdiv(p.vpc, kAThreads), (unsigned)C, (unsigned)N);
    return p;
}

}  // namespace

#define RB_ATTN_V(V_, ...)                 \
    switch (V_) {                          \
        case 8: { constexpr int VV = 8; __VA_ARGS__; } break; \
        case 4: { constexpr int VV = 4; __VA_ARGS__; } break; \
        case 2: { constexpr int VV = 2; __VA_ARGS__; } break; \
        default: { constexpr int VV = 1; __VA_ARGS__; } break; \
    }

// NB: the frame count is called Tn below because RB_DISPATCH_DTYPE binds the element type to `T`.
static int attn_forward(const void *x, const float *in_sb, const float *taps, void *out, int dtype, int N, int Tn, int C, int HW,
                        void *stream) {
    if (N < 0 || Tn < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0) return fail(RB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if ((int64_t)N * Tn * C * HW == 0) return RB_OK;
    if (!x || !taps || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (C > 65535 || N > 65535) return fail(RB_ERR_UNSUPPORTED, "attention shift: C or N > 65535");
    const APlan p = a_plan(dtype, N, C, HW, x, out, nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    if (in_sb) {
        RB_DISPATCH_DTYPE(dtype, RB_ATTN_V(p.V, (launch_kernel(k_attn_fwd<T, (VV * sizeof(T) <= 16 ? VV : 1), true>, p.grid, dim3(kAThreads), 0,
                                                               s, (const T *)x, in_sb, taps, (T *)out, Tn, C, HW, p.vpc, p.CG))));
    } else {
        RB_DISPATCH_DTYPE(dtype, RB_ATTN_V(p.V, (launch_kernel(k_attn_fwd<T, (VV * sizeof(T) <= 16 ? VV : 1), false>, p.grid, dim3(kAThreads), 0,
                                                               s, (const T *)x, in_sb, taps, (T *)out, Tn, C, HW, p.vpc, p.CG))));
    }
    return launched("k_attn_fwd");
}

extern "C" int rb_attention_shift_forward(const void *x, const float *taps, void *out, int dtype,
                                          int N, int Tn, int C, int HW, void *stream) {
    return attn_forward(x, nullptr, taps, out, dtype, N, Tn, C, HW, stream);
}

extern "C" int rb_bn_attention_shift_forward(const void *x, const float *in_scale_bias, const float *taps, void *out, int dtype,
                                             int N, int Tn, int C, int HW, void *stream) {
    if (!in_scale_bias) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return attn_forward(x, in_scale_bias, taps, out, dtype, N, Tn, C, HW, stream);
}

extern "C" size_t rb_attention_shift_backward_workspace_bytes(int N, int Tn, int C, int HW) {
    (void)Tn;
    if (N <= 0 || C <= 0 || HW <= 0) return 0;
    return (size_t)C * N * cdiv(HW, kAThreads) * 3 * sizeof(float);  // upper bound over every vector width
}

static int attn_backward(const void *x, const float *in_sb, const float *taps, const void *out_grad, void *x_grad, float *taps_grad,
                         int dtype, int N, int Tn, int C, int HW, void *workspace, size_t workspace_bytes, void *stream) {
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
    const APlan p = a_plan(dtype, N, C, HW, x, out_grad, x_grad);
    if (in_sb) {
        RB_DISPATCH_DTYPE(dtype, RB_ATTN_V(p.V, (launch_kernel(k_attn_bwd<T, (VV * sizeof(T) <= 16 ? VV : 1), true>, p.grid, dim3(kAThreads), 0,
                                                               s, (const T *)x, in_sb, taps, (const T *)out_grad, (T *)x_grad,
                                                               (float *)workspace, Tn, C, HW, p.vpc, p.CG, x_grad != nullptr,
                                                               taps_grad != nullptr))));
    } else {
        RB_DISPATCH_DTYPE(dtype, RB_ATTN_V(p.V, (launch_kernel(k_attn_bwd<T, (VV * sizeof(T) <= 16 ? VV : 1), false>, p.grid, dim3(kAThreads), 0,
                                                               s, (const T *)x, in_sb, taps, (const T *)out_grad, (T *)x_grad,
                                                               (float *)workspace, Tn, C, HW, p.vpc, p.CG, x_grad != nullptr,
                                                               taps_grad != nullptr))));
    }
    int rc = launched("k_attn_bwd");
    if (rc || !taps_grad) return rc;
    const int warps = 4;
    launch_kernel(k_attn_finalize, dim3(cdiv(C, warps)), dim3(warps * 32), 0, s, (const float *)workspace,
                  (int)(p.grid.x * p.grid.z), taps_grad, C);
    return launched("k_attn_finalize");
}

extern "C" int rb_attention_shift_backward(const void *x, const float *taps, const void *out_grad,
                                           void *x_grad, float *taps_grad, int dtype, int N, int Tn,
                                           int C, int HW, void *workspace, size_t workspace_bytes,
                                           void *stream) {
    return attn_backward(x, nullptr, taps, out_grad, x_grad, taps_grad, dtype, N, Tn, C, HW, workspace, workspace_bytes, stream);
}

extern "C" int rb_bn_attention_shift_backward(const void *x, const float *in_scale_bias, const float *taps, const void *out_grad,
                                              void *x_grad, float *taps_grad, int dtype, int N, int Tn, int C, int HW,
                                              void *workspace, size_t workspace_bytes, void *stream) {
    if (!in_scale_bias) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return attn_backward(x, in_scale_bias, taps, out_grad, x_grad, taps_grad, dtype, N, Tn, C, HW, workspace, workspace_bytes, stream);
}

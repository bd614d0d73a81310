// Generic 2D learnable-shift kernels ([N,C,H,W], shift [2,C]); restates
// /root/reference/cuda_src/rubiks2d_kernels.cu (:94-145 forward, :147-266 shift gradient with the
// 0.5 * central-difference rule for integer shifts, :269-379 input gradient, :381-397 normalise)
// with one (n,c) plane per block row and a deterministic two-stage shift-gradient reduction.
#include "common.cuh"

namespace rb {

static constexpr int kThreads = 256;
static constexpr int kItems = 4;

template <typename T, typename A>
__device__ __forceinline__ A tap2(const T *p, int h, int w, int H, int W) {
    if (h < 0 || w < 0 || h >= H || w >= W) return (A)0;
    return ld<A, T>(p + (int64_t)h * W + w);
}
// rubiks2d_kernels.cu:60-66
template <typename A> __device__ __forceinline__ A interp2d(A p00, A p01, A p10, A p11, A rh, A rw) {
    return p00 * (1 - rh) * (1 - rw) + p01 * (1 - rh) * rw + p10 * rh * (1 - rw) + p11 * rh * rw;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift2d_fwd(const T *__restrict__ x, const void *__restrict__ shift, int sdt, T *__restrict__ out,
              Geom2 g, int bpp, int quantize) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int plane = blockIdx.x / bpp, chunk = blockIdx.x % bpp;
    const int c = plane % g.C;
    const A offh = ld_param<A>(shift, sdt, c), offw = ld_param<A>(shift, sdt, g.C + c);
    const int fh = floor_fast(offh), fw = floor_fast(offw);
    const A rh = offh - fh, rw = offw - fw;
    const int HWo = g.Ho * g.Wo;
    const T *xp = x + (int64_t)plane * g.H * g.W;
    T *op = out + (int64_t)plane * HWo;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int p = (chunk * kItems + k) * kThreads + threadIdx.x;
        if (p >= HWo) break;
        const int ho = p / g.Wo, wo = p - ho * g.Wo;
        const int bh = ho * g.sH - g.pH, bw = wo * g.sW - g.pW;
        A v;
        if (quantize) {  // :116-121; out-of-bounds taps stay 0 in the reference's pre-zeroed output
            v = tap2<T, A>(xp, round_fast<A>(bh + offh), round_fast<A>(bw + offw), g.H, g.W);
        } else {
            const int h0 = bh + fh, w0 = bw + fw;
            v = interp2d<A>(tap2<T, A>(xp, h0, w0, g.H, g.W), tap2<T, A>(xp, h0, w0 + 1, g.H, g.W),
                            tap2<T, A>(xp, h0 + 1, w0, g.H, g.W),
                            tap2<T, A>(xp, h0 + 1, w0 + 1, g.H, g.W), rh, rw);
        }
        op[p] = cvt<T, A>(v);
    }
}

template <typename T, typename A>
__device__ __forceinline__ A tap2_adj(const T *gp, int h, int w, const Geom2 &g) {
    if (h % g.sH != 0 || w % g.sW != 0) return (A)0;
    h /= g.sH;
    w /= g.sW;
    if (h < 0 || w < 0 || h >= g.Ho || w >= g.Wo) return (A)0;
    return ld<A, T>(gp + (int64_t)h * g.Wo + w);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift2d_bwd_input(const void *__restrict__ shift, int sdt, const T *__restrict__ og,
                    T *__restrict__ gin, Geom2 g, int bpp, int quantize) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int plane = blockIdx.x / bpp, chunk = blockIdx.x % bpp;
    const int c = plane % g.C;
    const A sh = -ld_param<A>(shift, sdt, c), sw = -ld_param<A>(shift, sdt, g.C + c);
    const int fh = floor_fast(sh), fw = floor_fast(sw);
    const A rh = sh - fh, rw = sw - fw;
    const bool zero_shift = (sw == 0 && sh == 0);  // :322
    const int HW = g.H * g.W;
    const T *gp = og + (int64_t)plane * g.Ho * g.Wo;
    T *ip = gin + (int64_t)plane * HW;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int p = (chunk * kItems + k) * kThreads + threadIdx.x;
        if (p >= HW) break;
        const int h = p / g.W, w = p - h * g.W;
        const int bh = h + g.pH, bw = w + g.pW;
        A v;
        if (quantize) {  // :294-309
            v = tap2_adj<T, A>(gp, round_fast<A>(bh + sh), round_fast<A>(bw + sw), g);
        } else if (zero_shift) {
            v = tap2_adj<T, A>(gp, bh, bw, g);
        } else {
            v = interp2d<A>(tap2_adj<T, A>(gp, bh + fh, bw + fw, g),
                            tap2_adj<T, A>(gp, bh + fh, bw + fw + 1, g),
                            tap2_adj<T, A>(gp, bh + fh + 1, bw + fw, g),
                            tap2_adj<T, A>(gp, bh + fh + 1, bw + fw + 1, g), rh, rw);
        }
        ip[p] = cvt<T, A>(v);
    }
}

// grid = (chunks, C); block (chunk, c) owns images {chunk, chunk+chunks, ...}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift2d_bwd_shift(const T *__restrict__ x, const void *__restrict__ shift, int sdt,
                    const T *__restrict__ og, double *__restrict__ partial, Geom2 g, int chunks) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int chunk = blockIdx.x, c = blockIdx.y;
    const A offh = ld_param<A>(shift, sdt, c), offw = ld_param<A>(shift, sdt, g.C + c);
    const int fh = floor_fast(offh), fw = floor_fast(offw);
    A rh = offh - fh, rw = offw - fw;
    const A tol = (A)1e-7f;  // :189
    bool ih = false, iw = false;
    if (tol > rh && rh > -tol) { ih = true; rh = 0; }
    if (tol > rw && rw > -tol) { iw = true; rw = 0; }
    const int HWo = g.Ho * g.Wo;
    A accH = 0, accW = 0;
    for (int n = chunk; n < g.N; n += chunks) {
        const T *xp = x + ((int64_t)n * g.C + c) * g.H * g.W;
        const T *gp = og + ((int64_t)n * g.C + c) * HWo;
        for (int p = threadIdx.x; p < HWo; p += kThreads) {
            const int ho = p / g.Wo, wo = p - ho * g.Wo;
            const int h0 = ho * g.sH - g.pH + fh, w0 = wo * g.sW - g.pW + fw;
            const A p00 = tap2<T, A>(xp, h0, w0, g.H, g.W), p01 = tap2<T, A>(xp, h0, w0 + 1, g.H, g.W);
            const A p10 = tap2<T, A>(xp, h0 + 1, w0, g.H, g.W),
                    p11 = tap2<T, A>(xp, h0 + 1, w0 + 1, g.H, g.W);
            A gH = (1 - rw) * (p10 - p00) + rw * (p11 - p01);
            A gW = (1 - rh) * (p01 - p00) + rh * (p11 - p10);
            if (ih)  // P[a][b] = x(h0+a-1, w0+b-1); :238-244
                gH = (A)0.5f * ((1 - rw) * (p10 - tap2<T, A>(xp, h0 - 1, w0, g.H, g.W)) +
                                rw * (p11 - tap2<T, A>(xp, h0 - 1, w0 + 1, g.H, g.W)));
            if (iw)  // :246-252
                gW = (A)0.5f * ((1 - rh) * (p01 - tap2<T, A>(xp, h0, w0 - 1, g.H, g.W)) +
                                rh * (p11 - tap2<T, A>(xp, h0 + 1, w0 - 1, g.H, g.W)));
            const A up = ld<A, T>(gp + p);
            accH += gH * up;
            accW += gW * up;
        }
    }
    __shared__ double red[2][kThreads / 32];
    accH = warp_sum(accH);
    accW = warp_sum(accW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = (double)accH;
        red[1][warp] = (double)accW;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[threadIdx.x][w];
        partial[((int64_t)c * chunks + chunk) * 2 + threadIdx.x] = s;
    }
}

template <typename A>
__global__ void k_shift2d_finalize(const double *__restrict__ partial, int parts, void *shift_grad,
                                   int sdt, int C, int normalize) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0;
    const double *p = partial + (int64_t)c * parts * 2;
    for (int i = lane; i < parts; i += 32) {
        s0 += p[i * 2 + 0];
        s1 += p[i * 2 + 1];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane != 0) return;
    A gh = (A)s0, gw = (A)s1;
    if (normalize) {  // rubiks2d_kernels.cu:386-396
        const A mag = sqrt(gh * gh + gw * gw);
        if (mag > 0) {
            gh = gh / mag;
            gw = gw / mag;
        }
    }
    st_param<A>(shift_grad, sdt, c, gh);
    st_param<A>(shift_grad, sdt, C + c, gw);
}

int shift2d_bwd_chunks(const Geom2 &g) {
    int want = cdiv(2048, g.C);
    return g.N < want ? g.N : want;
}

int shift2d_forward_generic(const void *x, const void *shift, void *out, int dt, int sdt,
                            const Geom2 &g, int quantize, cudaStream_t s) {
    const int bpp = cdiv(g.Ho * g.Wo, kThreads * kItems);
    const int64_t blocks = (int64_t)g.N * g.C * bpp;
    if (blocks == 0) return RB_OK;
    if (blocks > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "shift2d forward: tensor too large");
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift2d_fwd<T>, dim3((unsigned)blocks), dim3(kThreads), 0, s, 
                              (const T *)x, shift, sdt, (T *)out, g, bpp, quantize)));
    return launched("k_shift2d_fwd");
}

int shift2d_bwd_input_generic(const void *shift, const void *og, void *gin, int dt, int sdt,
                              const Geom2 &g, int quantize, cudaStream_t s) {
    const int bpp = cdiv(g.H * g.W, kThreads * kItems);
    const int64_t blocks = (int64_t)g.N * g.C * bpp;
    if (blocks == 0) return RB_OK;
    if (blocks > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "shift2d backward: tensor too large");
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift2d_bwd_input<T>, dim3((unsigned)blocks), dim3(kThreads), 0, s, 
                              shift, sdt, (const T *)og, (T *)gin, g, bpp, quantize)));
    return launched("k_shift2d_bwd_input");
}

int shift2d_bwd_shift_generic(const void *x, const void *shift, const void *og, void *shift_grad,
                              int dt, int sdt, const Geom2 &g, int normalize, double *partial,
                              cudaStream_t s) {
    const int chunks = shift2d_bwd_chunks(g);
    if (g.C > 65535) return fail(RB_ERR_UNSUPPORTED, "shift2d backward: C > 65535");
    dim3 grid(chunks, g.C);
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift2d_bwd_shift<T>, dim3(grid), dim3(kThreads), 0, s, 
                              (const T *)x, shift, sdt, (const T *)og, partial, g, chunks)));
    int rc = launched("k_shift2d_bwd_shift");
    if (rc) return rc;
    const int warps = 4;
    if (dt == RB_F64)
        launch_kernel(k_shift2d_finalize<double>, dim3(cdiv(g.C, warps)), dim3(warps * 32), 0, s, partial, chunks, shift_grad,
                                                                          sdt, g.C, normalize);
    else
        launch_kernel(k_shift2d_finalize<float>, dim3(cdiv(g.C, warps)), dim3(warps * 32), 0, s, partial, chunks, shift_grad,
                                                                         sdt, g.C, normalize);
    return launched("k_shift2d_finalize");
}

}  // namespace rb

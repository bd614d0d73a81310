// Generic (any stride / padding / quantize / dtype) 3D learnable-shift kernels.
//
// These are the always-correct gather kernels the dispatcher falls back to when the tiled sm_100a
// kernels (shift3d_tiled.cu) do not cover a geometry.  They restate the arithmetic of
// /root/reference/cuda_src/rubiks3d_kernels.cu (:15-205 forward, :218-452 shift gradient,
// :455-929 input gradient, :932-960 normalisation) with a different execution shape:
//   * one (n,t,c) plane per block row, so floor / remainder / weights are block-uniform and the
//     per-element 4 div + 4 mod of the reference disappear;
//   * the shift gradient is reduced in registers -> warp shuffles -> per-block partials -> one
//     finalize kernel (deterministic), instead of 3 global atomics per element into a [3C,Ho,Wo]
//     buffer followed by a cuBLAS GEMV (rubiks.cpp:295-299,344-345).
#include "common.cuh"

namespace rb {

static constexpr int kThreads = 256;
static constexpr int kItems = 4;

template <typename T, typename A>
__device__ __forceinline__ A tap(const T *frame0, int t, int h, int w, int Tn, int H, int W,
                                 int64_t frame_stride) {
    if (t < 0 || h < 0 || w < 0 || t >= Tn || h >= H || w >= W) return (A)0;
    return ld<A, T>(frame0 + t * frame_stride + (int64_t)h * W + w);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift3d_fwd_generic(const T *__restrict__ x, const void *__restrict__ shift, int sdt,
                      T *__restrict__ out, Geom3 g, int bpp, int quantize) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int plane = blockIdx.x / bpp, chunk = blockIdx.x % bpp;
    const int c = plane % g.C, nt = plane / g.C, to = nt % g.To, n = nt / g.To;
    const A st = ld_param<A>(shift, sdt, c), sh = ld_param<A>(shift, sdt, g.C + c),
            sw = ld_param<A>(shift, sdt, 2 * g.C + c);
    const int ft = floor3d(st), fh = floor3d(sh), fw = floor3d(sw);
    const A rt = st - ft, rh = sh - fh, rw = sw - fw;
    const int HW = g.H * g.W, HWo = g.Ho * g.Wo;
    const int64_t fs = (int64_t)g.C * HW;
    const T *x0 = x + ((int64_t)n * g.T * g.C + c) * HW;
    T *op = out + (int64_t)plane * HWo;
    const int bt = to * g.sT - g.pT;
    int kt = 0, kh = 0, kw = 0;
    if (quantize) {  // rubiks3d_kernels.cu:76-79
        kt = (rt < (A)0.5f) ? ft : ft + 1;
        kh = (rh < (A)0.5f) ? fh : fh + 1;
        kw = (rw < (A)0.5f) ? fw : fw + 1;
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int p = (chunk * kItems + k) * kThreads + threadIdx.x;
        if (p >= HWo) break;
        const int ho = p / g.Wo, wo = p - ho * g.Wo;
        const int bh = ho * g.sH - g.pH, bw = wo * g.sW - g.pW;
        A v;
        if (quantize) {
            v = tap<T, A>(x0, bt + kt, bh + kh, bw + kw, g.T, g.H, g.W, fs);
        } else {
            const int t0 = bt + ft, h0 = bh + fh, w0 = bw + fw;
            const A q111 = tap<T, A>(x0, t0, h0, w0, g.T, g.H, g.W, fs);
            const A q112 = tap<T, A>(x0, t0, h0, w0 + 1, g.T, g.H, g.W, fs);
            const A q121 = tap<T, A>(x0, t0, h0 + 1, w0, g.T, g.H, g.W, fs);
            const A q122 = tap<T, A>(x0, t0, h0 + 1, w0 + 1, g.T, g.H, g.W, fs);
            const A q211 = tap<T, A>(x0, t0 + 1, h0, w0, g.T, g.H, g.W, fs);
            const A q212 = tap<T, A>(x0, t0 + 1, h0, w0 + 1, g.T, g.H, g.W, fs);
            const A q221 = tap<T, A>(x0, t0 + 1, h0 + 1, w0, g.T, g.H, g.W, fs);
            const A q222 = tap<T, A>(x0, t0 + 1, h0 + 1, w0 + 1, g.T, g.H, g.W, fs);
            v = (1 - rt) * ((1 - rh) * (q111 * (1 - rw) + q112 * rw) +
                            rh * (q121 * (1 - rw) + q122 * rw)) +
                rt * ((1 - rh) * (q211 * (1 - rw) + q212 * rw) +
                      rh * (q221 * (1 - rw) + q222 * rw));
        }
        op[p] = cvt<T, A>(v);
    }
}

// adjoint tap: valid iff every numerator is divisible by its stride (C '%', truncating) and the
// quotient is inside the output (rubiks3d_kernels.cu:586-594)
template <typename T, typename A>
__device__ __forceinline__ A tap_adj(const T *g0, int t, int h, int w, const Geom3 &g,
                                     int64_t frame_stride) {
    if (t % g.sT != 0 || h % g.sH != 0 || w % g.sW != 0) return (A)0;
    t /= g.sT;
    h /= g.sH;
    w /= g.sW;
    if (t < 0 || h < 0 || w < 0 || t >= g.To || h >= g.Ho || w >= g.Wo) return (A)0;
    return ld<A, T>(g0 + t * frame_stride + (int64_t)h * g.Wo + w);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift3d_bwd_input_generic(const void *__restrict__ shift, int sdt, const T *__restrict__ og,
                            T *__restrict__ gin, Geom3 g, int bpp, int quantize) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int plane = blockIdx.x / bpp, chunk = blockIdx.x % bpp;
    const int c = plane % g.C, nt = plane / g.C, t = nt % g.T, n = nt / g.T;
    const A st = -ld_param<A>(shift, sdt, c), sh = -ld_param<A>(shift, sdt, g.C + c),
            sw = -ld_param<A>(shift, sdt, 2 * g.C + c);
    const int ft = floor3d(st), fh = floor3d(sh), fw = floor3d(sw);
    const A rt = st - ft, rh = sh - fh, rw = sw - fw;
    const int HW = g.H * g.W, HWo = g.Ho * g.Wo;
    const int64_t fs = (int64_t)g.C * HWo;
    const T *g0 = og + ((int64_t)n * g.To * g.C + c) * HWo;
    T *ip = gin + (int64_t)plane * HW;
    const int bt = t + g.pT;
    const bool zero_shift = (st == 0 && sh == 0 && sw == 0);  // :561
    int kt = 0, kh = 0, kw = 0;
    if (quantize) {  // :534-536, on the NEGATED shift
        kt = (rt < (A)0.5f) ? ft : ft + 1;
        kh = (rh < (A)0.5f) ? fh : fh + 1;
        kw = (rw < (A)0.5f) ? fw : fw + 1;
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int p = (chunk * kItems + k) * kThreads + threadIdx.x;
        if (p >= HW) break;
        const int h = p / g.W, w = p - h * g.W;
        const int bh = h + g.pH, bw = w + g.pW;
        A v;
        if (quantize) {
            v = tap_adj<T, A>(g0, bt + kt, bh + kh, bw + kw, g, fs);
        } else if (zero_shift) {
            v = tap_adj<T, A>(g0, bt, bh, bw, g, fs);
        } else {
            const int t0 = bt + ft, h0 = bh + fh, w0 = bw + fw;
            const A q111 = tap_adj<T, A>(g0, t0, h0, w0, g, fs);
            const A q112 = tap_adj<T, A>(g0, t0, h0, w0 + 1, g, fs);
            const A q121 = tap_adj<T, A>(g0, t0, h0 + 1, w0, g, fs);
            const A q122 = tap_adj<T, A>(g0, t0, h0 + 1, w0 + 1, g, fs);
            const A q211 = tap_adj<T, A>(g0, t0 + 1, h0, w0, g, fs);
            const A q212 = tap_adj<T, A>(g0, t0 + 1, h0, w0 + 1, g, fs);
            const A q221 = tap_adj<T, A>(g0, t0 + 1, h0 + 1, w0, g, fs);
            const A q222 = tap_adj<T, A>(g0, t0 + 1, h0 + 1, w0 + 1, g, fs);
            v = (1 - rt) * ((1 - rh) * (q111 * (1 - rw) + q112 * rw) +
                            rh * (q121 * (1 - rw) + q122 * rw)) +
                rt * ((1 - rh) * (q211 * (1 - rw) + q212 * rw) +
                      rh * (q221 * (1 - rw) + q222 * rw));
        }
        ip[p] = cvt<T, A>(v);
    }
}

// rubiks3d_kernels.cu:208-215
template <typename A>
__device__ __forceinline__ A interp2(A p11, A p12, A p21, A p22, A d1, A d2) {
    return p11 * (1 - d1) * (1 - d2) + p12 * (1 - d1) * d2 + p21 * d1 * (1 - d2) + p22 * d1 * d2;
}

// Per-pixel shift gradient (rubiks3d_kernels.cu:283-446) reduced per block.
// grid = (chunks, C); block (chunk, c) owns the (n,to) planes {chunk, chunk+chunks, ...} of channel c
// and writes partial[(c * chunks + chunk) * 3 + axis] (double).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_shift3d_bwd_shift_generic(const T *__restrict__ x, const void *__restrict__ shift, int sdt,
                            const T *__restrict__ og, double *__restrict__ partial, Geom3 g,
                            int chunks) {
    pdl_sync();
    using A = typename Acc<T>::type;
    const int chunk = blockIdx.x, c = blockIdx.y;
    const A st = ld_param<A>(shift, sdt, c), sh = ld_param<A>(shift, sdt, g.C + c),
            sw = ld_param<A>(shift, sdt, 2 * g.C + c);
    const int ft = floor3d(st), fh = floor3d(sh), fw = floor3d(sw);
    const A rt = st - ft, rh = sh - fh, rw = sw - fw;
    // exact-integer rule: the SMALL coordinate moves to floor-1 on every axis whose remainder is 0
    const int at = (rt == 0) ? -1 : 0, ah = (rh == 0) ? -1 : 0, aw = (rw == 0) ? -1 : 0;
    const int HW = g.H * g.W, HWo = g.Ho * g.Wo;
    const int64_t fs = (int64_t)g.C * HW;
    A accT = 0, accH = 0, accW = 0;
    const int planes = g.N * g.To;
    for (int pl = chunk; pl < planes; pl += chunks) {
        const int n = pl / g.To, to = pl - n * g.To;
        const T *x0 = x + ((int64_t)n * g.T * g.C + c) * HW;
        const T *gp = og + ((int64_t)pl * g.C + c) * HWo;
        const int bt = to * g.sT - g.pT;
        const int tl = bt + ft + at, th = bt + ft + 1;
        for (int p = threadIdx.x; p < HWo; p += kThreads) {
            const int ho = p / g.Wo, wo = p - ho * g.Wo;
            const int bh = ho * g.sH - g.pH, bw = wo * g.sW - g.pW;
            const int hl = bh + fh + ah, hh = bh + fh + 1;
            const int wl = bw + fw + aw, wh = bw + fw + 1;
            const A q111 = tap<T, A>(x0, tl, hl, wl, g.T, g.H, g.W, fs);
            const A q112 = tap<T, A>(x0, tl, hl, wh, g.T, g.H, g.W, fs);
            const A q121 = tap<T, A>(x0, tl, hh, wl, g.T, g.H, g.W, fs);
            const A q122 = tap<T, A>(x0, tl, hh, wh, g.T, g.H, g.W, fs);
            const A q211 = tap<T, A>(x0, th, hl, wl, g.T, g.H, g.W, fs);
            const A q212 = tap<T, A>(x0, th, hl, wh, g.T, g.H, g.W, fs);
            const A q221 = tap<T, A>(x0, th, hh, wl, g.T, g.H, g.W, fs);
            const A q222 = tap<T, A>(x0, th, hh, wh, g.T, g.H, g.W, fs);
            const A gT = -interp2(q111, q112, q121, q122, rh, rw) + interp2(q211, q212, q221, q222, rh, rw);
            const A gH = -interp2(q111, q112, q211, q212, rt, rw) + interp2(q121, q122, q221, q222, rt, rw);
            const A gW = -interp2(q111, q121, q211, q221, rt, rh) + interp2(q112, q122, q212, q222, rt, rh);
            const A up = ld<A, T>(gp + p);
            accT += gT * up;
            accH += gH * up;
            accW += gW * up;
        }
    }
    __shared__ double red[3][kThreads / 32];
    accT = warp_sum(accT);
    accH = warp_sum(accH);
    accW = warp_sum(accW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = (double)accT;
        red[1][warp] = (double)accH;
        red[2][warp] = (double)accW;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[threadIdx.x][w];
        partial[((int64_t)c * chunks + chunk) * 3 + threadIdx.x] = s;
    }
}

// Sums the per-block partials of a channel (fixed order => deterministic), optionally normalises
// (rubiks3d_kernels.cu:941-958: nothing is rewritten when the norm is 0) and OVERWRITES shift_grad
// (addmv_ beta = 0, rubiks.cpp:344-345).  One warp per channel.  AccT is the reference's T.
template <typename A>
__global__ void k_shift3d_finalize(const double *__restrict__ partial, int parts, void *shift_grad,
                                   int sdt, int C, int normalize, A factor) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0, s2 = 0;
    const double *p = partial + (int64_t)c * parts * 3;
    for (int i = lane; i < parts; i += 32) {
        s0 += p[i * 3 + 0];
        s1 += p[i * 3 + 1];
        s2 += p[i * 3 + 2];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane != 0) return;
    A gt = (A)s0, gh = (A)s1, gw = (A)s2;
    if (normalize) {
        A ct, ch, cw;
        if (factor < 0) {
            ct = gt;
            ch = 0;
            cw = 0;
        } else {
            ct = gt * factor;
            ch = gh;
            cw = gw;
        }
        const A mag = sqrt(ct * ct + ch * ch + cw * cw);
        if (mag > 0) {
            gt = ct / mag;
            gh = ch / mag;
            gw = cw / mag;
        }
    }
    st_param<A>(shift_grad, sdt, c, gt);
    st_param<A>(shift_grad, sdt, C + c, gh);
    st_param<A>(shift_grad, sdt, 2 * C + c, gw);
}

// ---------------------------------------------------------------------------------------------

int generic_bwd_chunks(const Geom3 &g) {
    int want = cdiv(2048, g.C);
    int planes = g.N * g.To;
    return planes < want ? planes : want;
}

int shift3d_forward_generic(const void *x, const void *shift, void *out, int dt, int sdt,
                            const Geom3 &g, int quantize, cudaStream_t s) {
    const int HWo = g.Ho * g.Wo;
    const int bpp = cdiv(HWo, kThreads * kItems);
    const int64_t blocks = (int64_t)g.N * g.To * g.C * bpp;
    if (blocks == 0) return RB_OK;
    if (blocks > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "shift3d forward: tensor too large");
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift3d_fwd_generic<T>, dim3((unsigned)blocks), dim3(kThreads), 0, s, 
                              (const T *)x, shift, sdt, (T *)out, g, bpp, quantize)));
    return launched("k_shift3d_fwd_generic");
}

int shift3d_bwd_input_generic(const void *shift, const void *og, void *gin, int dt, int sdt,
                              const Geom3 &g, int quantize, cudaStream_t s) {
    const int HW = g.H * g.W;
    const int bpp = cdiv(HW, kThreads * kItems);
    const int64_t blocks = (int64_t)g.N * g.T * g.C * bpp;
    if (blocks == 0) return RB_OK;
    if (blocks > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "shift3d backward: tensor too large");
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift3d_bwd_input_generic<T>, dim3((unsigned)blocks), dim3(kThreads), 0, s, 
                              shift, sdt, (const T *)og, (T *)gin, g, bpp, quantize)));
    return launched("k_shift3d_bwd_input_generic");
}

int shift3d_finalize(const double *partial, int parts, void *shift_grad, int dt, int sdt, int C,
                     int normalize, double factor, cudaStream_t s) {
    const int warps = 4;
    if (dt == RB_F64)
        launch_kernel(k_shift3d_finalize<double>, dim3(cdiv(C, warps)), dim3(warps * 32), 0, s, partial, parts, shift_grad,
                                                                         sdt, C, normalize, factor);
    else
        launch_kernel(k_shift3d_finalize<float>, dim3(cdiv(C, warps)), dim3(warps * 32), 0, s, 
            partial, parts, shift_grad, sdt, C, normalize, (float)factor);
    return launched("k_shift3d_finalize");
}

int shift3d_bwd_shift_generic(const void *x, const void *shift, const void *og, void *shift_grad,
                              int dt, int sdt, const Geom3 &g, int normalize, double factor,
                              double *partial, cudaStream_t s) {
    const int chunks = generic_bwd_chunks(g);
    if (g.C > 65535) return fail(RB_ERR_UNSUPPORTED, "shift3d backward: C > 65535");
    dim3 grid(chunks, g.C);
    RB_DISPATCH_DTYPE(dt, (launch_kernel(k_shift3d_bwd_shift_generic<T>, dim3(grid), dim3(kThreads), 0, s, 
                              (const T *)x, shift, sdt, (const T *)og, partial, g, chunks)));
    int rc = launched("k_shift3d_bwd_shift_generic");
    if (rc) return rc;
    return shift3d_finalize(partial, chunks, shift_grad, dt, sdt, g.C, normalize, factor, s);
}

}  // namespace rb

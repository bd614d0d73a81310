// Fused BatchNorm2d (+ReLU) for the RubiksShiftBlock (rubiksnet/backbone.py:50-53,123-135: bn1->relu,
// bn2->relu, bn_last->relu) on NCHW activations [NI, C, HW] in bf16 / fp16 / fp32 with fp32 parameters.
//
// The reference leaves these to ATen/cuDNN; in bf16 that path costs 90 of 135 ms of a RubiksNet-Large
// training step on B200 (profiles/r01a_*): three generic ATen kernels per BN.  These are pure streaming
// ops, so each one here is a single vectorised (128-bit) pass bound by HBM:
//   forward  (training): stats pass (1 read)            -> apply pass  y = relu(x*scale+bias)   (1R + 1W)
//   backward           : reduce pass (reads dy, x)      -> apply pass  dx = c1*dyr + c2*x + c3 [+ residual]
// The ReLU mask is recomputed from x (no mask tensor), the per-channel sums are reduced
// registers -> warp shuffles -> smem -> double partials -> finalize kernel (deterministic, no atomics).
#include "common.cuh"

namespace rb {

static constexpr int kBT = 256;

// the streaming passes use 128-bit vector accesses on every activation tensor
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct FastDiv {  // exact unsigned division by a runtime constant for n < 2^31 (Granlund-Montgomery)
    uint32_t d, mul, shr;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t l = 0;
    while ((1u << l) < d) ++l;  // ceil(log2 d)
    uint64_t m = ((uint64_t(1) << (31 + l)) + d - 1) / d;  // fits in 32 bits for n < 2^31
    f.mul = (uint32_t)m;
    f.shr = l - 1;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) {
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

template <typename T> struct Vec {};
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__half> { static constexpr int N = 8; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct Vec<double> { static constexpr int N = 2; };

template <typename T, int N> struct alignas(sizeof(T) * N) Pack { T v[N]; };

template <typename T> __device__ __forceinline__ float tof(T v) { return (float)v; }
template <> __device__ __forceinline__ float tof<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float tof<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- per-channel reductions ---------------------------------------------------------------------
// grid (SPLIT, C).  MODE 0: sum x, sum x^2.  MODE 1: sum dyr, sum dyr * xhat  (dyr = relu-masked dy).
// Block (sp, c) owns images sp, sp+SPLIT, ... of channel c and walks the flattened (image, vector) space so
// that all threads stay busy even when a plane holds fewer vectors than the block has threads (14x14, 7x7).
// V = widest vector (elements) dividing HW, so every plane start is V-aligned.
template <typename T, int MODE, int V>
__global__ void __launch_bounds__(kBT)
k_bn_reduce(const T *__restrict__ x, const T *__restrict__ dy, const float *__restrict__ mean_invstd,
            const float *__restrict__ scale_bias, double *__restrict__ partial, int NI, int C, int HW, FastDiv vpp,
            int relu) {
    pdl_sync();
    const int sp = blockIdx.x, c = blockIdx.y, splits = gridDim.x;
    float s0 = 0.f, s1 = 0.f;
    float mean = 0.f, invstd = 1.f, sc = 1.f, bi = 0.f;
    if (MODE == 1) {
        mean = mean_invstd[2 * c];
        invstd = mean_invstd[2 * c + 1];
        sc = scale_bias[2 * c];
        bi = scale_bias[2 * c + 1];
    }
    const int n_count = (NI - sp + splits - 1) / splits;
    const uint32_t total = (uint32_t)n_count * vpp.d;
    const int64_t img_stride = (int64_t)splits * C * HW;
    const int64_t base0 = ((int64_t)sp * C + c) * HW;
    for (uint32_t idx = threadIdx.x; idx < total; idx += kBT) {
        const uint32_t nl = fdiv(idx, vpp);
        const uint32_t i = idx - nl * vpp.d;
        const int64_t off = base0 + (int64_t)nl * img_stride + (int64_t)i * V;
        const Pack<T, V> xv = *reinterpret_cast<const Pack<T, V> *>(x + off);
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float f = tof(xv.v[k]);
                s0 += f;
                s1 += f * f;
            }
        } else {
            const Pack<T, V> gv = *reinterpret_cast<const Pack<T, V> *>(dy + off);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float f = tof(xv.v[k]);
                float g = tof(gv.v[k]);
                if (relu && !(f * sc + bi > 0.f)) g = 0.f;
                s0 += g;
                s1 += g * ((f - mean) * invstd);
            }
        }
    }
    __shared__ double red[2][kBT / 32];
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = (double)s0;
        red[1][warp] = (double)s1;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kBT / 32; ++w) s += red[threadIdx.x][w];
        partial[((int64_t)c * splits + sp) * 2 + threadIdx.x] = s;
    }
}

template <typename T, int MODE>
static void launch_bn_reduce(const T *x, const T *dy, const float *mean_invstd, const float *scale_bias, double *partial,
                             int NI, int C, int HW, int relu, int splits, cudaStream_t s) {
    dim3 grid(splits, C);
    constexpr int VMAX = Vec<T>::N;
    int V = VMAX;
    while (V > 1 && HW % V != 0) V >>= 1;
    const FastDiv vpp = make_fastdiv((uint32_t)(HW / V));
#define RB_BN_REDUCE(VV)                                                                                          \
    launch_kernel(k_bn_reduce<T, MODE, (VV <= VMAX ? VV : 1)>, dim3(grid), dim3(kBT), 0, s, x, dy, mean_invstd, scale_bias, partial, NI, C, HW, \
                                                                    vpp, relu)
    switch (V) {
        case 8: RB_BN_REDUCE(8); break;
        case 4: RB_BN_REDUCE(4); break;
        case 2: RB_BN_REDUCE(2); break;
        default: RB_BN_REDUCE(1); break;
    }
#undef RB_BN_REDUCE
}

// one warp per channel: batch statistics -> save (mean, invstd), (scale, bias); running stats update
__global__ void k_bn_stats_finalize(const double *__restrict__ partial, int splits, int C, double count,
                                    const float *__restrict__ gamma, const float *__restrict__ beta,
                                    float *running_mean, float *running_var, float momentum, float eps,
                                    float *__restrict__ mean_invstd, float *__restrict__ scale_bias) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s = 0, ss = 0;
    for (int i = lane; i < splits; i += 32) {
        s += partial[((int64_t)c * splits + i) * 2];
        ss += partial[((int64_t)c * splits + i) * 2 + 1];
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane != 0) return;
    const double mean = s / count;
    double var = ss / count - mean * mean;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    mean_invstd[2 * c] = (float)mean;
    mean_invstd[2 * c + 1] = invstd;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale_bias[2 * c] = g * invstd;
    scale_bias[2 * c + 1] = b - (float)mean * g * invstd;
    if (running_mean) {  // torch.nn.BatchNorm2d: unbiased variance in the running estimate
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// eval mode: (scale, bias) and (mean, invstd) from the running statistics
__global__ void k_bn_eval_coeffs(const float *__restrict__ gamma, const float *__restrict__ beta,
                                 const float *__restrict__ running_mean, const float *__restrict__ running_var,
                                 float eps, int C, float *__restrict__ mean_invstd, float *__restrict__ scale_bias) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = rsqrtf(running_var[c] + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    mean_invstd[2 * c] = running_mean[c];
    mean_invstd[2 * c + 1] = invstd;
    scale_bias[2 * c] = g * invstd;
    scale_bias[2 * c + 1] = b - running_mean[c] * g * invstd;
}

// backward finalize: dgamma, dbeta and the per-channel coefficients of dx = c1*dyr + c2*x + c3
__global__ void k_bn_bwd_finalize(const double *__restrict__ partial, int splits, int C, double count,
                                  const float *__restrict__ gamma, const float *__restrict__ mean_invstd,
                                  int training, float *__restrict__ dgamma, float *__restrict__ dbeta,
                                  float *__restrict__ coef) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s = 0, sx = 0;
    for (int i = lane; i < splits; i += 32) {
        s += partial[((int64_t)c * splits + i) * 2];
        sx += partial[((int64_t)c * splits + i) * 2 + 1];
    }
    s = warp_sum(s);
    sx = warp_sum(sx);
    if (lane != 0) return;
    if (dgamma) dgamma[c] = (float)sx;
    if (dbeta) dbeta[c] = (float)s;
    const float g = gamma ? gamma[c] : 1.f;
    const float mean = mean_invstd[2 * c], invstd = mean_invstd[2 * c + 1];
    const float k = g * invstd;
    float c1 = k, c2 = 0.f, c3 = 0.f;
    if (training) {
        const float m1 = (float)(s / count), m2 = (float)(sx / count);
        c2 = -k * m2 * invstd;
        c3 = k * m2 * invstd * mean - k * m1;
    }
    coef[4 * c] = c1;
    coef[4 * c + 1] = c2;
    coef[4 * c + 2] = c3;
}

// ---- streaming apply passes ----------------------------------------------------------------------
// MODE 0: y  = act(x*scale + bias)
// MODE 1: dx = c1*dyr + c2*x + c3 (+ residual), dyr = relu-masked dy
template <typename T, int MODE>
__global__ void __launch_bounds__(kBT)
k_bn_apply(const T *__restrict__ x, const T *__restrict__ dy, const T *__restrict__ residual,
           const float *__restrict__ scale_bias, const float *__restrict__ coef, T *__restrict__ out,
           int64_t total, int C, FastDiv hw, int relu) {
    pdl_sync();
    constexpr int V = Vec<T>::N;
    const int64_t nvec = total / V;
    const Pack<T, V> *xp = reinterpret_cast<const Pack<T, V> *>(x);
    const Pack<T, V> *gp = reinterpret_cast<const Pack<T, V> *>(dy);
    const Pack<T, V> *rp = reinterpret_cast<const Pack<T, V> *>(residual);
    Pack<T, V> *op = reinterpret_cast<Pack<T, V> *>(out);
    const int HW = (int)hw.d;
    for (int64_t v = (int64_t)blockIdx.x * kBT + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kBT) {
        const uint32_t e0 = (uint32_t)(v * V);
        const uint32_t plane = fdiv(e0, hw);
        const int within = (int)(e0 - plane * (uint32_t)HW);
        int c = (int)(plane % (uint32_t)C);
        int nextb = HW - within;  // index within this vector of the first element of the next plane
        const Pack<T, V> xv = xp[v];
        Pack<T, V> gv, rv, ov;
        if (MODE == 1) {
            gv = gp[v];
            if (residual) rv = rp[v];
        }
        float sc = __ldg(scale_bias + 2 * c), bi = __ldg(scale_bias + 2 * c + 1);
        float c1 = 0.f, c2 = 0.f, c3 = 0.f;
        if (MODE == 1) {
            c1 = __ldg(coef + 4 * c);
            c2 = __ldg(coef + 4 * c + 1);
            c3 = __ldg(coef + 4 * c + 2);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (k == nextb) {  // vector straddles a plane boundary (HW % V != 0)
                nextb += HW;
                c = (c + 1 == C) ? 0 : c + 1;
                sc = __ldg(scale_bias + 2 * c);
                bi = __ldg(scale_bias + 2 * c + 1);
                if (MODE == 1) {
                    c1 = __ldg(coef + 4 * c);
                    c2 = __ldg(coef + 4 * c + 1);
                    c3 = __ldg(coef + 4 * c + 2);
                }
            }
            const float f = tof(xv.v[k]);
            const float y = f * sc + bi;
            if (MODE == 0) {
                ov.v[k] = cvt<T, float>(relu ? fmaxf(y, 0.f) : y);
            } else {
                float g = tof(gv.v[k]);
                if (relu && !(y > 0.f)) g = 0.f;
                float d = c1 * g + c2 * f + c3;
                if (residual) d += tof(rv.v[k]);
                ov.v[k] = cvt<T, float>(d);
            }
        }
        op[v] = ov;
    }
    // tail (total % V elements), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (int)(total - nvec * V)) {
        const int64_t e = nvec * V + threadIdx.x;
        const int c = (int)((e / HW) % C);
        const float f = tof(x[e]);
        const float y = f * scale_bias[2 * c] + scale_bias[2 * c + 1];
        if (MODE == 0) {
            out[e] = cvt<T, float>(relu ? fmaxf(y, 0.f) : y);
        } else {
            float g = tof(dy[e]);
            if (relu && !(y > 0.f)) g = 0.f;
            float d = coef[4 * c] * g + coef[4 * c + 1] * f + coef[4 * c + 2];
            if (residual) d += tof(residual[e]);
            out[e] = cvt<T, float>(d);
        }
    }
}

// The same passes for 16-bit types on planes whose size is a multiple of 4 (every map of RubiksNet: 112x112 ... 14x14; not 7x7):
// the two 4-element halves of a 128-bit vector each lie inside ONE plane, so the per-element "did the plane end here" test
// of k_bn_apply disappears (two coefficient sets per vector instead), and the arithmetic runs on packed fp32 pairs (FFMA2).
// ncu on k_bn_apply<bf16, 0> at 14x14: 77 % issue utilisation -- the streaming passes are instruction-bound on L2-resident maps.
template <typename T, int MODE>
__global__ void __launch_bounds__(kBT)
k_bn_apply_h(const T *__restrict__ x, const T *__restrict__ dy, const T *__restrict__ residual,
             const float *__restrict__ scale_bias, const float *__restrict__ coef, T *__restrict__ out,
             int64_t total, int C, FastDiv hw, int relu) {
    pdl_sync();
    constexpr int V = 8;
    static_assert(sizeof(T) == 2, "16-bit element types");
    const int64_t nvec = total / V;  // HW % 4 == 0 and the tensor holds whole planes: total % 4 == 0; a 4-element tail is possible
    const Pack<T, V> *xp = reinterpret_cast<const Pack<T, V> *>(x);
    const Pack<T, V> *gp = reinterpret_cast<const Pack<T, V> *>(dy);
    const Pack<T, V> *rp = reinterpret_cast<const Pack<T, V> *>(residual);
    Pack<T, V> *op = reinterpret_cast<Pack<T, V> *>(out);
    const int HW = (int)hw.d;
    const float2 zero2 = make_float2(0.f, 0.f);
    for (int64_t v = (int64_t)blockIdx.x * kBT + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kBT) {
        const uint32_t e0 = (uint32_t)(v * V);
        const uint32_t plane = fdiv(e0, hw);
        const int within = (int)(e0 - plane * (uint32_t)HW);
        const int c0 = (int)(plane % (uint32_t)C);
        const int c1 = within + 4 < HW ? c0 : (c0 + 1 == C ? 0 : c0 + 1);  // channel of the second half
        const Pack<T, V> xv = xp[v];
        Pack<T, V> gv, rv, ov;
        if (MODE == 1) {
            gv = gp[v];
            if (residual) rv = rp[v];
        }
        const float2 sb0 = __ldg(reinterpret_cast<const float2 *>(scale_bias) + c0), sb1 = __ldg(reinterpret_cast<const float2 *>(scale_bias) + c1);
        float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0;
        if (MODE == 1) {
            k0 = __ldg(reinterpret_cast<const float4 *>(coef) + c0);
            k1 = __ldg(reinterpret_cast<const float4 *>(coef) + c1);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 sc = h ? make_float2(sb1.x, sb1.x) : make_float2(sb0.x, sb0.x);
            const float2 bi = h ? make_float2(sb1.y, sb1.y) : make_float2(sb0.y, sb0.y);
            const float4 kk = h ? k1 : k0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = h * 4 + j * 2;
                const float2 f = make_float2(tof(xv.v[k]), tof(xv.v[k + 1]));
                const float2 y = __ffma2_rn(f, sc, bi);
                if (MODE == 0) {
                    ov.v[k] = cvt<T, float>(relu ? fmaxf(y.x, 0.f) : y.x);
                    ov.v[k + 1] = cvt<T, float>(relu ? fmaxf(y.y, 0.f) : y.y);
                } else {
                    float2 g = make_float2(tof(gv.v[k]), tof(gv.v[k + 1]));
                    if (relu && !(y.x > 0.f)) g.x = 0.f;
                    if (relu && !(y.y > 0.f)) g.y = 0.f;
                    // c1 * g + c2 * f + c3
                    float2 d = __ffma2_rn(make_float2(kk.x, kk.x), g, __ffma2_rn(make_float2(kk.y, kk.y), f, make_float2(kk.z, kk.z)));
                    if (residual) d = __fadd2_rn(d, make_float2(tof(rv.v[k]), tof(rv.v[k + 1])));
                    ov.v[k] = cvt<T, float>(d.x);
                    ov.v[k + 1] = cvt<T, float>(d.y);
                }
            }
        }
        op[v] = ov;
    }
    (void)zero2;
    // tail: total % 8 is 0 or 4 elements (one half vector), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (int)(total - nvec * V)) {
        const int64_t e = nvec * V + threadIdx.x;
        const int c = (int)((e / HW) % C);
        const float f = tof(x[e]);
        const float y = f * scale_bias[2 * c] + scale_bias[2 * c + 1];
        if (MODE == 0) {
            out[e] = cvt<T, float>(relu ? fmaxf(y, 0.f) : y);
        } else {
            float g = tof(dy[e]);
            if (relu && !(y > 0.f)) g = 0.f;
            float d = coef[4 * c] * g + (coef[4 * c + 1] * f + coef[4 * c + 2]);
            if (residual) d += tof(residual[e]);
            out[e] = cvt<T, float>(d);
        }
    }
}

// 16-bit element types on planes of a multiple of 4 elements take the halves kernel
template <typename T, int MODE>
static void launch_bn_apply(int blocks, cudaStream_t s, const T *x, const T *dy, const T *residual, const float *scale_bias,
                            const float *coef, T *out, int64_t total, int C, FastDiv hw, int relu) {
    if constexpr (sizeof(T) == 2) {
        if (hw.d % 4 == 0 && total < (int64_t(1) << 32)) {  // the halves kernel indexes elements in 32 bits
            launch_kernel(k_bn_apply_h<T, MODE>, dim3(blocks), dim3(kBT), 0, s, x, dy, residual, scale_bias, coef, out, total, C, hw, relu);
            return;
        }
    }
    launch_kernel(k_bn_apply<T, MODE>, dim3(blocks), dim3(kBT), 0, s, x, dy, residual, scale_bias, coef, out, total, C, hw, relu);
}

// ---- channel-resident forward pass (small maps) -------------------------------------------------------------------------
// On the 14x14 / 7x7 maps (39 of the 51 blocks of RubiksNet-Large) a whole channel -- all NI planes of HW elements -- fits
// one CTA's shared memory: 256 x 196 x 2 B = 100 KB.  One CTA per channel then does what the streaming path needs three
// launches and a second read of the tensor for: it loads the channel ONCE (8 vectors in flight per thread), reduces the
// statistics locally (no partial slices, no finalize kernel: the channel has one owner), and applies from shared memory;
// with y == NULL it is the statistics pass alone in one launch.  Measured at 256 x 288 x 14x14 (tools/bench_bn.py):
// stats+apply 0.030 -> 0.022-0.025 ms.  Same formulas as k_bn_stats_finalize; sums are fp32 per thread over <= 100
// elements, then double.  (The backward twin -- x and dy resident, 200 KB, ONE CTA per SM -- was built and measured slower
// than the streaming passes, 0.049 vs 0.046 ms: with one CTA per SM the load and the store phases of a wave do not
// overlap; it was removed, gpurun_out/r02aj_bench_bn.log.)
static constexpr int kRT = 512;  // two CTAs per SM (100 KB each): the load phase of one overlaps the store phase of the other
static constexpr int kResSmem = 216 * 1024;
static constexpr int kResMinC = 128;  // fewer channels than this leave most SMs without a CTA: streaming path instead
static std::atomic<int> g_bn_resident{1};

template <int MODE> static bool bn_resident_ok(int dtype, int NI, int C, int HW) {
    if (!g_bn_resident.load(std::memory_order_relaxed)) return false;
    if (dtype != RB_BF16 && dtype != RB_F16) return false;
    if (C < kResMinC) return false;
    return (size_t)NI * HW * 2 <= (size_t)kResSmem;
}

__device__ __forceinline__ void block_sum2(double &a, double &b, double (*red)[kRT / 32]) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
    __syncthreads();
    a = 0; b = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kRT / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(kRT)
k_bn_fwd_resident(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta, float *running_mean,
                  float *running_var, T *__restrict__ y, float *__restrict__ mean_invstd, float *__restrict__ scale_bias, int NI,
                  int C, int HW, FastDiv vpp, float momentum, float eps, int relu) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char res_smem[];
    T *xs = reinterpret_cast<T *>(res_smem);
    __shared__ double red[2][kRT / 32];
    __shared__ float coef[2];
    const int c = blockIdx.x;
    const uint32_t total = (uint32_t)NI * vpp.d;
    const int64_t cs = (int64_t)C * HW;
    const T *xc = x + (int64_t)c * HW;
    constexpr int U = 8;  // loads in flight per thread
    float s0 = 0.f, s1 = 0.f;
    for (uint32_t i0 = threadIdx.x; i0 < total; i0 += kRT * U) {
        Pack<T, V> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = i0 + u * kRT;
            if (idx < total) {
                const uint32_t n = fdiv(idx, vpp), j = idx - n * vpp.d;
                v[u] = *reinterpret_cast<const Pack<T, V> *>(xc + n * cs + j * V);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = i0 + u * kRT;
            if (idx < total) {
                if (y != nullptr) *reinterpret_cast<Pack<T, V> *>(xs + (size_t)idx * V) = v[u];  // statistics only: nothing to keep
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const float f = tof(v[u].v[k]);
                    s0 += f;
                    s1 += f * f;
                }
            }
        }
    }
    double d0 = (double)s0, d1 = (double)s1;
    block_sum2(d0, d1, red);
    if (threadIdx.x == 0) {
        const double count = (double)NI * HW;
        const double mean = d0 / count;
        double var = d1 / count - mean * mean;
        if (var < 0) var = 0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        mean_invstd[2 * c] = (float)mean;
        mean_invstd[2 * c + 1] = invstd;
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        coef[0] = g * invstd;
        coef[1] = b - (float)mean * g * invstd;
        scale_bias[2 * c] = coef[0];
        scale_bias[2 * c + 1] = coef[1];
        if (running_mean) {
            const double unbiased = count > 1 ? var * count / (count - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
    if (y == nullptr) return;
    __syncthreads();
    const float sc = coef[0], bi = coef[1];
    T *yc = y + (int64_t)c * HW;
    for (uint32_t idx = threadIdx.x; idx < total; idx += kRT) {
        const uint32_t n = fdiv(idx, vpp), j = idx - n * vpp.d;
        const Pack<T, V> xv = *reinterpret_cast<const Pack<T, V> *>(xs + (size_t)idx * V);
        Pack<T, V> ov;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float v = tof(xv.v[k]) * sc + bi;
            ov.v[k] = cvt<T, float>(relu ? fmaxf(v, 0.f) : v);
        }
        *reinterpret_cast<Pack<T, V> *>(yc + n * cs + j * V) = ov;
    }
}

template <typename T, int V, int MODE> static int bn_resident_attr() {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_bn_fwd_resident<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResSmem);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(bn resident): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    return RB_OK;
}

// widest vector (elements of a 2-byte type) dividing the plane size: plane starts are then aligned to it (base 16-byte aligned)
static int bn_resident_vec(int HW) {
    int V = 8;
    while (V > 1 && HW % V != 0) V >>= 1;
    return V;
}

static int bn_splits(int NI, int C) {
    int want = cdiv(4 * 148, C);
    if (want < 1) want = 1;
    return want < NI ? want : NI;
}
int bn_stats_finalize(const double *partial, int splits, int C, double count, const float *gamma, const float *beta,
                      float *running_mean, float *running_var, float momentum, float eps, float *mean_invstd, float *scale_bias,
                      cudaStream_t s);
int bn_apply_forward(const void *x, const float *scale_bias, void *y, int dtype, int NI, int C, int HW, int relu, cudaStream_t s);
static size_t bn_ws_bytes(int NI, int C) {
    return ((size_t)C * bn_splits(NI, C) * 2 * sizeof(double) + (size_t)C * 4 * sizeof(float) + 255) & ~(size_t)255;
}

}  // namespace rb

using namespace rb;

// statistics reduced elsewhere (the epilogue of the producing GEMM, pw_conv.cu): partial[(c * splits + sp) * 2 + {0,1}]
int rb::bn_stats_finalize(const double *partial, int splits, int C, double count, const float *gamma, const float *beta,
                          float *running_mean, float *running_var, float momentum, float eps, float *mean_invstd,
                          float *scale_bias, cudaStream_t s) {
    launch_kernel(k_bn_stats_finalize, dim3(cdiv(C, 4)), dim3(128), 0, s, partial, splits, C, count, gamma, beta, running_mean, running_var, momentum,
                                                   eps, mean_invstd, scale_bias);
    return launched("k_bn_stats_finalize");
}

// y = act(x * scale + bias) with given per-channel coefficients (the apply pass alone)
int rb::bn_apply_forward(const void *x, const float *scale_bias, void *y, int dtype, int NI, int C, int HW, int relu,
                         cudaStream_t s) {
    if (dtype_size(dtype) == 0 || dtype == RB_F64) return fail(RB_ERR_INVALID_ARGUMENT, "bn: dtype %d not supported", dtype);
    if (C > 65535) return fail(RB_ERR_UNSUPPORTED, "bn: C > 65535");
    if (!aligned16(x) || !aligned16(y)) return fail(RB_ERR_INVALID_ARGUMENT, "bn: activation pointers must be 16-byte aligned");
    const int64_t total = (int64_t)NI * C * HW;
    const FastDiv hw = make_fastdiv((uint32_t)HW);
    RB_DISPATCH_DTYPE(dtype, {
        const int64_t nvec = total / Vec<T>::N;
        int blocks = (int)((nvec + kBT - 1) / kBT);
        const int cap = sm_count() * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        launch_bn_apply<T, 0>(blocks, s, (const T *)x, nullptr, nullptr, scale_bias, nullptr, (T *)y, total, C, hw, relu);
    });
    return launched("k_bn_apply<fwd>");
}

extern "C" size_t rb_bn_workspace_bytes(int NI, int C) {
    if (NI <= 0 || C <= 0) return 0;
    return bn_ws_bytes(NI, C);
}

// NB: `Tn`-style naming is not needed here, but RB_DISPATCH_DTYPE binds `T`, so no parameter is called T.
extern "C" int rb_bn_act_forward(const void *x, const float *gamma, const float *beta, float *running_mean,
                                 float *running_var, void *y, float *mean_invstd, float *scale_bias, int dtype,
                                 int NI, int C, int HW, int training, float momentum, float eps, int relu,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    if (NI < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0 || dtype == RB_F64) return fail(RB_ERR_INVALID_ARGUMENT, "bn: dtype %d not supported", dtype);
    const int64_t total = (int64_t)NI * C * HW;
    if (total == 0) return RB_OK;
    if (total > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "bn: tensor too large");
    if (!x || !mean_invstd || !scale_bias) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (!training && (!running_mean || !running_var)) return fail(RB_ERR_INVALID_ARGUMENT, "eval mode needs running stats");
    if (!aligned16(x) || !aligned16(y)) return fail(RB_ERR_INVALID_ARGUMENT, "bn: activation pointers must be 16-byte aligned");
    if (C > 65535) return fail(RB_ERR_UNSUPPORTED, "bn: C > 65535");
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (training && bn_resident_ok<0>(dtype, NI, C, HW)) {
        // small maps: one CTA per channel keeps the channel in shared memory -- statistics + apply in ONE launch (statistics
        // only, y == NULL: the same launch without the shared-memory copy, no partial slices and no finalize kernel)
        const int V = bn_resident_vec(HW);
        const FastDiv vpp = make_fastdiv((uint32_t)(HW / V));
        const size_t smem = y ? (size_t)NI * HW * 2 : 0;
#define RB_BN_RES_FWD(TT, VV)                                                                                                   \
    do {                                                                                                                        \
        if ((rc = bn_resident_attr<TT, VV, 0>())) return rc;                                                                    \
        launch_kernel(k_bn_fwd_resident<TT, VV>, dim3(C), dim3(kRT), smem, s, (const TT *)x, gamma, beta, running_mean, running_var, \
                      (TT *)y, mean_invstd, scale_bias, NI, C, HW, vpp, momentum, eps, relu);                                   \
    } while (0)
#define RB_BN_RES_FWD_T(TT)                                                                       \
    switch (V) {                                                                                  \
        case 8: RB_BN_RES_FWD(TT, 8); break;                                                      \
        case 4: RB_BN_RES_FWD(TT, 4); break;                                                      \
        case 2: RB_BN_RES_FWD(TT, 2); break;                                                      \
        default: RB_BN_RES_FWD(TT, 1); break;                                                     \
    }
        if (dtype == RB_BF16) { RB_BN_RES_FWD_T(__nv_bfloat16) } else { RB_BN_RES_FWD_T(__half) }
#undef RB_BN_RES_FWD_T
#undef RB_BN_RES_FWD
        return launched("k_bn_fwd_resident");
    }
    if (training) {
        if (!workspace || workspace_bytes < bn_ws_bytes(NI, C))
            return fail(RB_ERR_WORKSPACE, "bn forward needs %zu workspace bytes", bn_ws_bytes(NI, C));
        const int splits = bn_splits(NI, C);
        RB_DISPATCH_DTYPE(dtype, (launch_bn_reduce<T, 0>((const T *)x, (const T *)x, nullptr, nullptr, (double *)workspace,
                                                         NI, C, HW, 0, splits, s)));
        if ((rc = launched("k_bn_reduce<stats>"))) return rc;
        launch_kernel(k_bn_stats_finalize, dim3(cdiv(C, 4)), dim3(128), 0, s, (const double *)workspace, splits, C, (double)NI * HW, gamma, beta,
                                                       running_mean, running_var, momentum, eps, mean_invstd, scale_bias);
        if ((rc = launched("k_bn_stats_finalize"))) return rc;
    } else {
        launch_kernel(k_bn_eval_coeffs, dim3(cdiv(C, 128)), dim3(128), 0, s, gamma, beta, running_mean, running_var, eps, C, mean_invstd,
                                                      scale_bias);
        if ((rc = launched("k_bn_eval_coeffs"))) return rc;
    }
    if (!y) return RB_OK;  // statistics only: the apply pass is folded into the consumer
    const FastDiv hw = make_fastdiv((uint32_t)HW);
    RB_DISPATCH_DTYPE(dtype, {
        const int64_t nvec = total / Vec<T>::N;
        int blocks = (int)((nvec + kBT - 1) / kBT);
        const int cap = sm_count() * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        launch_bn_apply<T, 0>(blocks, s, (const T *)x, nullptr, nullptr, scale_bias, nullptr, (T *)y, total, C, hw, relu);
    });
    return launched("k_bn_apply<fwd>");
}

extern "C" int rb_bn_act_backward(const void *x, const void *dy, const void *residual, const float *gamma,
                                  const float *mean_invstd, const float *scale_bias, void *dx, float *dgamma,
                                  float *dbeta, int dtype, int NI, int C, int HW, int training, int relu,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (NI < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0 || dtype == RB_F64) return fail(RB_ERR_INVALID_ARGUMENT, "bn: dtype %d not supported", dtype);
    const int64_t total = (int64_t)NI * C * HW;
    cudaStream_t s = (cudaStream_t)stream;
    if (total == 0) {
        if (dgamma && C > 0) cudaMemsetAsync(dgamma, 0, C * sizeof(float), s);
        if (dbeta && C > 0) cudaMemsetAsync(dbeta, 0, C * sizeof(float), s);
        return RB_OK;
    }
    if (total > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "bn: tensor too large");
    if (!x || !dy || !mean_invstd || !scale_bias) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (C > 65535) return fail(RB_ERR_UNSUPPORTED, "bn: C > 65535");
    if (!aligned16(x) || !aligned16(dy) || !aligned16(residual) || !aligned16(dx))
        return fail(RB_ERR_INVALID_ARGUMENT, "bn: activation pointers must be 16-byte aligned");
    if (!workspace || workspace_bytes < bn_ws_bytes(NI, C))
        return fail(RB_ERR_WORKSPACE, "bn backward needs %zu workspace bytes", bn_ws_bytes(NI, C));
    int rc;
    const int splits = bn_splits(NI, C);
    double *partial = (double *)workspace;
    float *coef = (float *)((char *)workspace + (size_t)C * splits * 2 * sizeof(double));
    RB_DISPATCH_DTYPE(dtype, (launch_bn_reduce<T, 1>((const T *)x, (const T *)dy, mean_invstd, scale_bias, partial, NI, C,
                                                     HW, relu, splits, s)));
    if ((rc = launched("k_bn_reduce<bwd>"))) return rc;
    launch_kernel(k_bn_bwd_finalize, dim3(cdiv(C, 4)), dim3(128), 0, s, partial, splits, C, (double)NI * HW, gamma, mean_invstd, training, dgamma,
                                                 dbeta, coef);
    if ((rc = launched("k_bn_bwd_finalize"))) return rc;
    if (!dx) return RB_OK;
    const FastDiv hw = make_fastdiv((uint32_t)HW);
    RB_DISPATCH_DTYPE(dtype, {
        const int64_t nvec = total / Vec<T>::N;
        int blocks = (int)((nvec + kBT - 1) / kBT);
        const int cap = sm_count() * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        launch_bn_apply<T, 1>(blocks, s, (const T *)x, (const T *)dy, (const T *)residual, scale_bias, coef, (T *)dx, total, C, hw, relu);
    });
    return launched("k_bn_apply<bwd>");
}

// 1 (default): small maps take the channel-resident one-launch passes; 0: always the streaming passes (A/B, tests)
extern "C" void rb_bn_set_resident(int enabled) { g_bn_resident.store(enabled ? 1 : 0); }

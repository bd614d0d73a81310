// Patch matrix of the network's first layer (conv1: 3x3, stride 2, padding 1, 3 -> width channels,
// rubiksnet/backbone.py:148-149): cols[i, t, ho, wo] = x[i, ci, ho*S - 1 + kh, wo*S - 1 + kw], t = (ci*3 + kh)*3 + kw, rows
// t >= 9*Cin are zero (channel padding up to a multiple of the GEMM's K step).  With it conv1 and its weight gradient are the
// library's own tensor-core GEMMs (k_pw_conv / k_pw_tf32 on [NI, Tpad, Ho*Wo], k_pw_wgrad) instead of cuDNN's implicit GEMM
// plus two NCHW<->NHWC transposes per direction.  16-byte (bf16) / 2 x 16-byte (fp32) coalesced stores, stride-S gathers that
// neighbouring threads share through L1.
#include "common.cuh"

namespace rb {
namespace {

template <typename TO> struct Out8 {};
template <> struct Out8<__nv_bfloat16> {
    static __device__ __forceinline__ void store(__nv_bfloat16 *dst, const float (&v)[8]) {
        __nv_bfloat162 p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(p);
    }
};
template <> struct Out8<float> {
    static __device__ __forceinline__ void store(float *dst, const float (&v)[8]) {
        reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// One thread = 8 consecutive outputs of one (image, input channel, kernel row): it loads the 7*S + 3 inputs the three kernel
// columns share ONCE and writes three tap rows; units beyond the 3*Cin real ones write the zero padding rows.
template <typename TI, typename TO, int S>
__global__ void __launch_bounds__(256) k_im2col3x3(const TI *__restrict__ x, TO *__restrict__ cols, int NI, int Cin, int H, int W,
                                                   int Ho, int Wo, int T, int Tpad, int units_per_image, uint32_t total) {
    pdl_sync();
    constexpr int SEG = 7 * S + 3;
    const uint32_t HoWo = (uint32_t)(Ho * Wo), G = HoWo >> 3, gpr = (uint32_t)Wo >> 3;
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
        const uint32_t g = idx % G, r0 = idx / G;
        const uint32_t unit = r0 % (uint32_t)units_per_image, i = r0 / (uint32_t)units_per_image;
        const int t0 = (int)unit * 3;  // first tap row of this unit
        float seg[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) seg[j] = 0.f;
        const uint32_t ho = g / gpr, wo0 = (g - ho * gpr) << 3;
        if (t0 < T) {
            const int ci = (int)unit / 3, kh = (int)unit - ci * 3;
            const int hi = (int)ho * S - 1 + kh;
            if (hi >= 0 && hi < H) {
                const TI *row = x + ((int64_t)(i * Cin + ci) * H + hi) * W;
                const int w0 = (int)wo0 * S - 1;
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    const int wi = w0 + j;
                    if (wi >= 0 && wi < W) seg[j] = ld<float, TI>(row + wi);
                }
            }
        }
        TO *dst = cols + ((int64_t)i * Tpad + t0) * HoWo + (g << 3);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            if (t0 + kw >= Tpad) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = seg[j * S + kw];
            Out8<TO>::store(dst + (int64_t)kw * HoWo, v);
        }
    }
}

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" int rb_im2col3x3(const void *x, void *cols, int in_dtype, int out_dtype, int NI, int Cin, int H, int W, int stride,
                            int Tpad, void *stream) {
    if (NI < 0 || Cin <= 0 || H <= 0 || W <= 0 || stride <= 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent");
    const int T = 9 * Cin;
    if (Tpad < T) return fail(RB_ERR_INVALID_ARGUMENT, "im2col: Tpad %d < 9*Cin", Tpad);
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    if (Wo % 8 != 0) return fail(RB_ERR_UNSUPPORTED, "im2col: output width %d is not a multiple of 8", Wo);
    if (NI == 0) return RB_OK;
    if (!x || !cols) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (reinterpret_cast<uintptr_t>(cols) & 15) return fail(RB_ERR_INVALID_ARGUMENT, "im2col: cols must be 16-byte aligned");
    if (stride != 1 && stride != 2) return fail(RB_ERR_UNSUPPORTED, "im2col: stride %d (1 or 2 supported)", stride);
    const int units = cdiv(Tpad, 3);  // 3*Cin real (channel, kernel row) units + the zero rows
    const int64_t total = (int64_t)NI * units * ((Ho * Wo) >> 3);
    if ((int64_t)NI * Tpad * Ho * Wo > 0x7fffffffLL || (int64_t)NI * Cin * H * W > 0x7fffffffLL || total > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "im2col: tensor too large");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
#define RB_IM2COL(TI, TO)                                                                                                       \
    do {                                                                                                                        \
        if (stride == 2)                                                                                                        \
            launch_kernel(k_im2col3x3<TI, TO, 2>, dim3((unsigned)blocks), dim3(256), 0, s, (const TI *)x, (TO *)cols, NI, Cin, H, W, \
                          Ho, Wo, T, Tpad, units, (uint32_t)total);                                                             \
        else                                                                                                                    \
            launch_kernel(k_im2col3x3<TI, TO, 1>, dim3((unsigned)blocks), dim3(256), 0, s, (const TI *)x, (TO *)cols, NI, Cin, H, W, \
                          Ho, Wo, T, Tpad, units, (uint32_t)total);                                                             \
    } while (0)
    if (in_dtype == RB_F32 && out_dtype == RB_BF16) RB_IM2COL(float, __nv_bfloat16);
    else if (in_dtype == RB_BF16 && out_dtype == RB_BF16) RB_IM2COL(__nv_bfloat16, __nv_bfloat16);
    else if (in_dtype == RB_F32 && out_dtype == RB_F32) RB_IM2COL(float, float);
    else return fail(RB_ERR_UNSUPPORTED, "im2col: dtype pair (%d -> %d) not supported", in_dtype, out_dtype);
#undef RB_IM2COL
    return launched("k_im2col3x3");
}

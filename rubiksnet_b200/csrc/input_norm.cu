// Input pipeline of the evaluation / training loops on the GPU: decoded frames arrive as ONE uint8 array per video,
// H x W x (3 T) (the reference's `Stack` transform, rubiksnet/transforms.py:329-342), and the reference then spends most of its
// loader time on the CPU turning it into the network input: ToTorchFormatTensor (HWC -> CHW permute + .float().div(255),
// :345-363, "this transpose takes 80% of the loading time/CPU") and GroupNormalize (per-channel (x - mean) / std, :66-79).
// Here the uint8 stack is what crosses PCIe (4x fewer bytes than fp32) and one kernel does permute + scale + normalise +
// cast:    out[n, c, h, w] = (x[n, h, w, c] / 255 - mean[c % 3]) / std[c % 3],   c in [0, 3T),  out fp32 or bf16
// which viewed [N*T, 3, H, W] is the model input (rubiksnet/models.py:106).  A CTA transposes a (32 pixels x all 3T channels)
// tile through shared memory: coalesced uint8 reads along the channel-minor input, coalesced stores along W.
#include "common.cuh"

namespace rb {
namespace {

constexpr int kINPix = 32;  // pixels per CTA tile

template <typename TO>
__global__ void __launch_bounds__(256) k_frames_to_clip(const unsigned char *__restrict__ x, TO *__restrict__ out, int HW, int CT,
                                                        float m0, float m1, float m2, float i0, float i1, float i2, int div255) {
    pdl_sync();
    extern __shared__ unsigned char tile[];  // [kINPix][CT + 1]
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * kINPix;
    const int np = min(kINPix, HW - p0);
    const int pitch = CT + 1;
    const unsigned char *src = x + ((int64_t)n * HW + p0) * CT;  // np * CT contiguous bytes
    for (int i = threadIdx.x; i < np * CT; i += 256) {
        const int p = i / CT, c = i - p * CT;
        tile[p * pitch + c] = src[i];
    }
    __syncthreads();
    const float scale = div255 ? 1.f / 255.f : 1.f;
    for (int i = threadIdx.x; i < CT * kINPix; i += 256) {
        const int c = i / kINPix, p = i - c * kINPix;
        if (p >= np) continue;
        const int rgb = c % 3;
        const float mean = rgb == 0 ? m0 : (rgb == 1 ? m1 : m2), inv = rgb == 0 ? i0 : (rgb == 1 ? i1 : i2);
        const float v = ((float)tile[p * pitch + c] * scale - mean) * inv;
        out[((int64_t)n * CT + c) * HW + p0 + p] = cvt<TO, float>(v);
    }
}

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" int rb_frames_to_clip(const void *frames_u8, void *out, int out_dtype, int N, int H, int W, int channels,
                                 const float *mean3, const float *std3, int div255, void *stream) {
    if (N < 0 || H < 0 || W < 0 || channels <= 0 || channels % 3 != 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent");
    if ((int64_t)N * H * W == 0) return RB_OK;
    if (!frames_u8 || !out || !mean3 || !std3) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if ((int64_t)N * H * W * channels > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "frames_to_clip: tensor too large");
    if (N > 65535) return fail(RB_ERR_UNSUPPORTED, "frames_to_clip: N > 65535");
    for (int i = 0; i < 3; ++i)
        if (!(std3[i] > 0.f)) return fail(RB_ERR_INVALID_ARGUMENT, "frames_to_clip: std must be positive");
    const int HW = H * W;
    const size_t smem = (size_t)kINPix * (channels + 1);
    if (smem > 48 * 1024) return fail(RB_ERR_UNSUPPORTED, "frames_to_clip: %d channels per pixel", channels);
    const dim3 grid((unsigned)cdiv(HW, kINPix), (unsigned)N);
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == RB_F32)
        launch_kernel(k_frames_to_clip<float>, grid, dim3(256), smem, s, (const unsigned char *)frames_u8, (float *)out, HW, channels,
                      mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2], div255);
    else if (out_dtype == RB_BF16)
        launch_kernel(k_frames_to_clip<__nv_bfloat16>, grid, dim3(256), smem, s, (const unsigned char *)frames_u8, (__nv_bfloat16 *)out,
                      HW, channels, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2], div255);
    else
        return fail(RB_ERR_UNSUPPORTED, "frames_to_clip: output dtype %d (fp32 or bf16)", out_dtype);
    return launched("k_frames_to_clip");
}

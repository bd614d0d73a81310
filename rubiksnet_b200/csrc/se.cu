// Squeeze-and-excitation gate of RubiksShiftBlock (rubiksnet/backbone.py:56-71, tier "small"): the two passes over the
// activation tensor -- global average pool per (image, channel) plane, and the per-plane rescale -- and their gradients.
// The [NI, C] -> [NI, C/r] -> [NI, C] gate MLP in between works on a few thousand numbers and stays in the host framework.
//   rb_plane_reduce:  out[plane] = scale * sum_p a[plane, p] * (b ? b[plane, p] : 1)        (pool; gate gradient sum g*x)
//   rb_plane_scale:   out[plane, p] = a[plane, p] * s[plane] + (t ? t[plane] : 0)            (x * gate; g * gate + dpool / HW)
// Streaming, HBM-bound: one warp per plane for the reduction (16-byte loads when the plane size allows), a grid-stride
// vector pass for the rescale.
#include "common.cuh"

namespace rb {
namespace {

template <typename T> __device__ __forceinline__ float se_tof(T v) { return (float)v; }
template <> __device__ __forceinline__ float se_tof<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float se_tof<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, int V> struct alignas(sizeof(T) * V) SePack { T v[V]; };

template <typename T, int V>
__global__ void __launch_bounds__(256) k_plane_reduce(const T *__restrict__ a, const T *__restrict__ b, float *__restrict__ out,
                                                      int planes, int HW, float scale) {
    pdl_sync();
    const int plane = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (plane >= planes) return;
    const T *pa = a + (int64_t)plane * HW, *pb = b ? b + (int64_t)plane * HW : nullptr;
    float s = 0.f;
    for (int i = lane * V; i < HW; i += 32 * V) {
        const SePack<T, V> va = *reinterpret_cast<const SePack<T, V> *>(pa + i);
        if (pb) {
            const SePack<T, V> vb = *reinterpret_cast<const SePack<T, V> *>(pb + i);
#pragma unroll
            for (int k = 0; k < V; ++k) s = fmaf(se_tof(va.v[k]), se_tof(vb.v[k]), s);
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) s += se_tof(va.v[k]);
        }
    }
    s = warp_sum(s);
    if (lane == 0) out[plane] = s * scale;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) k_plane_scale(const T *__restrict__ a, const float *__restrict__ s, const float *__restrict__ t,
                                                     T *__restrict__ out, int64_t nvec, int vec_per_plane) {
    pdl_sync();
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * 256) {
        const int plane = (int)(v / vec_per_plane);
        const float sc = __ldg(s + plane), tt = t ? __ldg(t + plane) : 0.f;
        const SePack<T, V> va = reinterpret_cast<const SePack<T, V> *>(a)[v];
        SePack<T, V> vo;
#pragma unroll
        for (int k = 0; k < V; ++k) vo.v[k] = cvt<T, float>(fmaf(se_tof(va.v[k]), sc, tt));
        reinterpret_cast<SePack<T, V> *>(out)[v] = vo;
    }
}

int pick_v(int HW, size_t es, const void *p0, const void *p1, const void *p2) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2);
    int v = (int)(16 / es);
    while (v > 1 && (HW % v != 0 || (al & (uintptr_t)(v * es - 1)) != 0)) v >>= 1;
    return v;
}

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" int rb_plane_reduce(const void *a, const void *b, float *out, int dtype, int planes, int HW, float scale, void *stream) {
    if (planes < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0 || dtype == RB_F64) return fail(RB_ERR_INVALID_ARGUMENT, "plane_reduce: dtype %d not supported", dtype);
    if (planes == 0) return RB_OK;
    if (!a || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if ((int64_t)planes * HW > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "plane_reduce: tensor too large");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)cdiv(planes, 8);
    const int v = pick_v(HW, dtype_size(dtype), a, b, nullptr);
#define RB_SE_RED(VV)                                                                                                    \
    RB_DISPATCH_DTYPE(dtype, (launch_kernel(k_plane_reduce<T, (VV * sizeof(T) <= 16 ? VV : 1)>, dim3(blocks), dim3(256), 0, s, (const T *)a, (const T *)b, \
                                                                                                       out, planes, HW, scale)))
    switch (v) {
        case 8: RB_SE_RED(8); break;
        case 4: RB_SE_RED(4); break;
        case 2: RB_SE_RED(2); break;
        default: RB_SE_RED(1); break;
    }
#undef RB_SE_RED
    return launched("k_plane_reduce");
}

extern "C" int rb_plane_scale(const void *a, const float *s_, const float *t, void *out, int dtype, int planes, int HW, void *stream) {
    if (planes < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (dtype_size(dtype) == 0 || dtype == RB_F64) return fail(RB_ERR_INVALID_ARGUMENT, "plane_scale: dtype %d not supported", dtype);
    if ((int64_t)planes * HW == 0) return RB_OK;
    if (!a || !s_ || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if ((int64_t)planes * HW > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "plane_scale: tensor too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int v = pick_v(HW, dtype_size(dtype), a, out, nullptr);
    const int64_t nvec = (int64_t)planes * HW / v;
    int blocks = (int)((nvec + 255) / 256);
    const int cap = sm_count() * 16;
    if (blocks > cap) blocks = cap;
#define RB_SE_SC(VV)                                                                                                   \
    RB_DISPATCH_DTYPE(dtype, (launch_kernel(k_plane_scale<T, (VV * sizeof(T) <= 16 ? VV : 1)>, dim3(blocks), dim3(256), 0, s, (const T *)a, s_, t, (T *)out, \
                                                                                                      nvec, HW / v)))
    switch (v) {
        case 8: RB_SE_SC(8); break;
        case 4: RB_SE_SC(4); break;
        case 2: RB_SE_SC(2); break;
        default: RB_SE_SC(1); break;
    }
#undef RB_SE_SC
    return launched("k_plane_scale");
}

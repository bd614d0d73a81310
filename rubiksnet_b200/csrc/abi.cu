// extern "C" entry points of librubiks_b200.so (include/rubiks_b200.h): argument checks, output
// geometry, workspace carving and implementation dispatch -- the role of the host drivers in
// /root/reference/cuda_src/rubiks.cpp:44-67,94-155,181-253,256-379, without ATen.
#include "common.cuh"

namespace rb {

thread_local char g_err[512] = {0};
thread_local int g_last_impl = RB_IMPL_AUTO;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_forced_impl{RB_IMPL_AUTO};
std::atomic<int> g_pdl{1};

// Plain (not programmatically serialized) empty kernel: it starts only after everything before it in the stream has
// completed, and a programmatically launched successor starts only after it -- whatever was written before the fence is
// "two launches old" for every later kernel (common.cuh).
__global__ void k_launch_fence() {}
int launch_fence(cudaStream_t s) {
    k_launch_fence<<<1, 32, 0, s>>>();
    return launched("k_launch_fence");
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
        cached_dev = dev;
    }
    return cached;
}

// shift3d_generic.cu
int generic_bwd_chunks(const Geom3 &g);
int shift3d_forward_generic(const void *, const void *, void *, int, int, const Geom3 &, int, cudaStream_t);
int shift3d_bwd_input_generic(const void *, const void *, void *, int, int, const Geom3 &, int, cudaStream_t);
int shift3d_bwd_shift_generic(const void *, const void *, const void *, void *, int, int, const Geom3 &, int,
                              double, double *, cudaStream_t);
// shift3d_tiled.cu
bool shift3d_tiled_supported(int dt, const Geom3 &g, int quantize);
int shift3d_forward_tiled(const void *, const void *, void *, int, int, const Geom3 &, cudaStream_t);
size_t shift3d_backward_tiled_workspace(int dt, const Geom3 &g);
int shift3d_backward_tiled(const void *, const void *, const void *, void *, void *, int, int, const Geom3 &,
                           int, double, void *, cudaStream_t);
// shift3d_strip.cu
bool shift3d_strip_supported(int dt, const Geom3 &g, int quantize);
int shift3d_forward_strip(const void *, const void *, void *, int, int, const Geom3 &, cudaStream_t);
size_t shift3d_backward_strip_workspace(int dt, const Geom3 &g);
int shift3d_backward_strip(const void *, const void *, const void *, void *, void *, int, int, const Geom3 &, int,
                           double, void *, cudaStream_t);
// shift3d_strip.cu: the 2D shift through the strip kernels (one-frame clips)
bool shift2d_strip_supported(int dt, const Geom2 &g, int quantize);
int shift2d_forward_strip(const void *x, const void *shift, void *out, int dt, int sdt, const Geom2 &g, cudaStream_t s);
size_t shift2d_backward_strip_workspace(int dt, const Geom2 &g);
bool shift2d_tiled_supported(int dt, const Geom2 &g, int quantize);
int shift2d_forward_tiled(const void *x, const void *shift, void *out, int dt, int sdt, const Geom2 &g, cudaStream_t s);
size_t shift2d_backward_tiled_workspace(int dt, const Geom2 &g);
int shift2d_backward_tiled(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt, int sdt,
                           const Geom2 &g, int normalize, void *workspace, cudaStream_t s);
int shift2d_backward_strip(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt, int sdt,
                           const Geom2 &g, int normalize, void *workspace, cudaStream_t s);
// shift2d_generic.cu
int shift2d_bwd_chunks(const Geom2 &g);
int shift2d_forward_generic(const void *, const void *, void *, int, int, const Geom2 &, int, cudaStream_t);
int shift2d_bwd_input_generic(const void *, const void *, void *, int, int, const Geom2 &, int, cudaStream_t);
int shift2d_bwd_shift_generic(const void *, const void *, const void *, void *, int, int, const Geom2 &, int,
                              double *, cudaStream_t);

// pw_conv.cu
int pw_conv_forward(const void *x, const void *w, int w_dt, int w_trans, const void *residual, void *out, int NI, int K,
                    int N, int HW, const float *a_sb, const void *shift, int shift_dt, int T, int H, int W, cudaStream_t s,
                    double *stats = nullptr, size_t stats_bytes = 0, int *stats_splits = nullptr);
int bn_stats_finalize(const double *partial, int splits, int C, double count, const float *gamma, const float *beta,
                      float *running_mean, float *running_var, float momentum, float eps, float *mean_invstd, float *scale_bias,
                      cudaStream_t s);
int bn_apply_forward(const void *x, const float *scale_bias, void *y, int dtype, int NI, int C, int HW, int relu, cudaStream_t s);
size_t pw_conv_wgrad_workspace(int NI, int M, int N, int HW);
void pw_conv_set_tuning(int min_n_splits);
// pw_conv2.cu
size_t pw2_weight_image_bytes(int rows, int contraction);
int pw2_weight_pack(const float *w, int N, int K, int trans, void *image, cudaStream_t s);
int pw2_supported(int NI, int K, int N, int HW, int has_bn);
int pw2_weight_pack_multi(const void *items_device, int count, cudaStream_t s);
void pw2_set_tuning(int op_stages, int kc, int wait_ns);
#ifdef RB_DEBUG_TRACE
void pw2_set_debug(int flags);
#endif
bool pw3_supported(const void *x, const void *out, const void *res, int NI, int K, int N, int HW);
int pw3_forward(const void *x, const void *w, int w_dt, const void *res, void *out, int NI, int K, int N, int HW, const float *a_sb,
                cudaStream_t s, double *stats = nullptr, size_t stats_bytes = 0, int *stats_splits = nullptr);
bool pw3_stats_preferred(int NI, int K, int N, int HW);
void pw3_set_enabled(int on);
size_t wg3_workspace(int NI, int M, int N, int HW);
void wg3_set_tuning(int burst, int l2_256, int max_stages);
bool wg3_supported(const void *g, const void *x, int NI, int M, int N, int HW);
int wg3_run(const void *g, const void *x, float *dw, int NI, int M, int N, int HW, const float *x_sb, void *workspace, cudaStream_t s);
int pw_tf32_forward(const float *x, const float *w, const float *res, float *out, int NI, int K, int N, int HW,
                    const float *in_sb, const float *out_sb, int relu, int resident, cudaStream_t s);
int pw2_forward(const void *x, const void *wimg, const void *residual, void *out, int NI, int K, int N, int HW,
                const float *a_sb, cudaStream_t s);
#ifdef RB_DEBUG_TRACE
void pw_conv_set_trace(void *p);
#endif
int pw_weight_pack(const float *w, void *w_nk, void *w_kn, int N, int K, cudaStream_t s);
int pw_weight_pack_multi(const void *items_device, int count, cudaStream_t s);
int pw_conv_wgrad(const void *g, const void *x, float *dw, int NI, int M, int N, int HW, const float *x_sb,
                  const void *shift, int shift_dt, int T, int H, int W, void *workspace, cudaStream_t s);

static int make_geom3(Geom3 &g, int N, int T, int C, int H, int W, int sT, int sH, int sW, int pT, int pH,
                      int pW) {
    if (N < 0 || T < 0 || C < 0 || H < 0 || W < 0)
        return fail(RB_ERR_INVALID_ARGUMENT, "negative extent [%d,%d,%d,%d,%d]", N, T, C, H, W);
    // the reference only prints to stderr for stride <= 0 (rubiks.cpp:162-164) and then divides by it
    if (sT <= 0 || sH <= 0 || sW <= 0)
        return fail(RB_ERR_INVALID_ARGUMENT, "stride must be > 0, got (%d,%d,%d)", sT, sH, sW);
    if (pT < 0 || pH < 0 || pW < 0)
        return fail(RB_ERR_INVALID_ARGUMENT, "padding must be >= 0, got (%d,%d,%d)", pT, pH, pW);
    g.N = N; g.T = T; g.C = C; g.H = H; g.W = W;
    g.sT = sT; g.sH = sH; g.sW = sW; g.pT = pT; g.pH = pH; g.pW = pW;
    g.To = T > 0 ? rb_out_len(T, sT, pT) : 0;
    g.Ho = H > 0 ? rb_out_len(H, sH, pH) : 0;
    g.Wo = W > 0 ? rb_out_len(W, sW, pW) : 0;
    if ((int64_t)N * T * C * H * W > 0x7fffffffLL || (int64_t)N * g.To * C * g.Ho * g.Wo > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    return RB_OK;
}

static int check_dtypes(int dt, int sdt) {
    if (dtype_size(dt) == 0) return fail(RB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dt);
    if (dtype_size(sdt) == 0) return fail(RB_ERR_INVALID_ARGUMENT, "unknown shift dtype %d", sdt);
    if (dt == RB_F64 && sdt != RB_F64)
        return fail(RB_ERR_INVALID_ARGUMENT, "float64 activations need a float64 shift");
    return RB_OK;
}

// AUTO: strip kernels (stride 1) > tiled kernels (stride 1 / 2) > generic gather kernels
static int pick_impl(int dt, const Geom3 &g, int quantize, int *err) {
    const int forced = g_forced_impl.load(std::memory_order_relaxed);
    const bool strip_ok = shift3d_strip_supported(dt, g, quantize);
    const bool tiled_ok = shift3d_tiled_supported(dt, g, quantize);
    *err = RB_OK;
    if (forced == RB_IMPL_GENERIC) return RB_IMPL_GENERIC;
    if (forced == RB_IMPL_TILED) {
        if (!tiled_ok) *err = fail(RB_ERR_UNSUPPORTED, "RB_IMPL_TILED forced but geometry is not covered by the tiled kernels");
        return RB_IMPL_TILED;
    }
    if (forced == RB_IMPL_STRIP) {
        if (!strip_ok) *err = fail(RB_ERR_UNSUPPORTED, "RB_IMPL_STRIP forced but geometry is not covered by the strip kernels");
        return RB_IMPL_STRIP;
    }
    return strip_ok ? RB_IMPL_STRIP : (tiled_ok ? RB_IMPL_TILED : RB_IMPL_GENERIC);
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_abi_version(void) { return RB_ABI_VERSION; }
const char *rb_last_error(void) { return g_err; }
uint64_t rb_launch_count(void) { return g_launches.load(); }
void rb_launch_count_reset(void) { g_launches.store(0); }
void rb_set_impl(int impl) { g_forced_impl.store(impl); }
void rb_set_dependent_launch(int enabled) { g_pdl.store(enabled ? 1 : 0); }
int rb_last_impl(void) { return g_last_impl; }

int rb_out_len(int in_len, int stride, int pad) { return (in_len + 2 * pad - 1) / stride + 1; }

int rb_shift3d_forward(const void *x, const void *shift, void *out, int dtype, int shift_dtype, int N,
                       int T, int C, int H, int W, int sT, int sH, int sW, int pT, int pH, int pW,
                       int quantize, void *stream) {
    Geom3 g;
    int rc = make_geom3(g, N, T, C, H, W, sT, sH, sW, pT, pH, pW);
    if (rc) return rc;
    if ((rc = check_dtypes(dtype, shift_dtype))) return rc;
    if ((int64_t)N * g.To * C * g.Ho * g.Wo == 0) return RB_OK;
    if (!x || !shift || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int impl = pick_impl(dtype, g, quantize, &rc);
    if (rc) return rc;
    g_last_impl = impl;
    if (impl == RB_IMPL_STRIP) return shift3d_forward_strip(x, shift, out, dtype, shift_dtype, g, s);
    if (impl == RB_IMPL_TILED) return shift3d_forward_tiled(x, shift, out, dtype, shift_dtype, g, s);
    return shift3d_forward_generic(x, shift, out, dtype, shift_dtype, g, quantize, s);
}

size_t rb_shift3d_backward_workspace_bytes(int dtype, int N, int T, int C, int H, int W, int sT, int sH,
                                           int sW, int pT, int pH, int pW) {
    Geom3 g;
    if (make_geom3(g, N, T, C, H, W, sT, sH, sW, pT, pH, pW)) return 0;
    if ((int64_t)N * g.To * C * g.Ho * g.Wo == 0) return 0;
    size_t generic = (size_t)C * generic_bwd_chunks(g) * 3 * sizeof(double);
    size_t tiled = shift3d_tiled_supported(dtype, g, 0) ? shift3d_backward_tiled_workspace(dtype, g) : 0;
    size_t strip = shift3d_strip_supported(dtype, g, 0) ? shift3d_backward_strip_workspace(dtype, g) : 0;
    size_t need = generic > tiled ? generic : tiled;
    if (strip > need) need = strip;
    return (need + 255) & ~(size_t)255;
}

int rb_shift3d_backward(const void *x, const void *shift, const void *out_grad, void *x_grad,
                        void *shift_grad, int dtype, int shift_dtype, int N, int T, int C, int H, int W,
                        int sT, int sH, int sW, int pT, int pH, int pW, int normalize_grad,
                        double normalize_t_factor, int quantize, void *workspace, size_t workspace_bytes,
                        void *stream) {
    Geom3 g;
    int rc = make_geom3(g, N, T, C, H, W, sT, sH, sW, pT, pH, pW);
    if (rc) return rc;
    if ((rc = check_dtypes(dtype, shift_dtype))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!x_grad && !shift_grad) return RB_OK;
    if ((int64_t)N * g.To * C * g.Ho * g.Wo == 0) {
        // empty upstream gradient: x_grad is all zeros, shift_grad = addmv_ of nothing = 0
        if (x_grad && (int64_t)N * T * C * H * W > 0)
            cudaMemsetAsync(x_grad, 0, (size_t)N * T * C * H * W * dtype_size(dtype), s);
        if (shift_grad && C > 0) cudaMemsetAsync(shift_grad, 0, (size_t)3 * C * dtype_size(shift_dtype), s);
        return RB_OK;
    }
    if (!x || !shift || !out_grad) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    const size_t need = rb_shift3d_backward_workspace_bytes(dtype, N, T, C, H, W, sT, sH, sW, pT, pH, pW);
    if (shift_grad && (!workspace || workspace_bytes < need))
        return fail(RB_ERR_WORKSPACE, "shift3d backward needs %zu workspace bytes, got %zu", need,
                    workspace_bytes);
    const int impl = pick_impl(dtype, g, quantize, &rc);
    if (rc) return rc;
    g_last_impl = impl;
    if (impl == RB_IMPL_STRIP)
        return shift3d_backward_strip(x, shift, out_grad, x_grad, shift_grad, dtype, shift_dtype, g,
                                      normalize_grad, normalize_t_factor, workspace, s);
    if (impl == RB_IMPL_TILED)
        return shift3d_backward_tiled(x, shift, out_grad, x_grad, shift_grad, dtype, shift_dtype, g,
                                      normalize_grad, normalize_t_factor, workspace, s);
    // host order of rubiks.cpp:324-376: shift gradient (+reduce, normalise), then input gradient
    if (shift_grad) {
        rc = shift3d_bwd_shift_generic(x, shift, out_grad, shift_grad, dtype, shift_dtype, g, normalize_grad,
                                       normalize_t_factor, (double *)workspace, s);
        if (rc) return rc;
    }
    if (x_grad) rc = shift3d_bwd_input_generic(shift, out_grad, x_grad, dtype, shift_dtype, g, quantize, s);
    return rc;
}

static int make_geom2(Geom2 &g, int N, int C, int H, int W, int sH, int sW, int pH, int pW) {
    if (N < 0 || C < 0 || H < 0 || W < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if (sH <= 0 || sW <= 0) return fail(RB_ERR_INVALID_ARGUMENT, "stride must be > 0, got (%d,%d)", sH, sW);
    if (pH < 0 || pW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "padding must be >= 0, got (%d,%d)", pH, pW);
    g.N = N; g.C = C; g.H = H; g.W = W; g.sH = sH; g.sW = sW; g.pH = pH; g.pW = pW;
    g.Ho = H > 0 ? rb_out_len(H, sH, pH) : 0;
    g.Wo = W > 0 ? rb_out_len(W, sW, pW) : 0;
    if ((int64_t)N * C * H * W > 0x7fffffffLL || (int64_t)N * C * g.Ho * g.Wo > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    return RB_OK;
}

// stride-1 / pad-0 / non-quantized 2D shifts (what RubiksNet's attention-quantized variant uses outside its 4 down-sampling
// blocks) run on the TMA-staged strip kernels; rb_set_impl(RB_IMPL_GENERIC) forces the generic gather
static bool use_strip2d(int dtype, int shift_dtype, const Geom2 &g, int quantize) {
    const int forced = g_forced_impl.load(std::memory_order_relaxed);
    if (forced == RB_IMPL_GENERIC || forced == RB_IMPL_TILED) return false;
    if (shift_dtype == RB_F64) return false;
    return shift2d_strip_supported(dtype, g, quantize);
}
// ... and the stride-2 ones of the down-sampling blocks (or any geometry under rb_set_impl(RB_IMPL_TILED)) on the tiled kernel
static bool use_tiled2d(int dtype, int shift_dtype, const Geom2 &g, int quantize) {
    const int forced = g_forced_impl.load(std::memory_order_relaxed);
    if (forced == RB_IMPL_GENERIC || forced == RB_IMPL_STRIP) return false;
    if (shift_dtype == RB_F64) return false;
    return shift2d_tiled_supported(dtype, g, quantize);
}

int rb_shift2d_forward(const void *x, const void *shift, void *out, int dtype, int shift_dtype, int N,
                       int C, int H, int W, int sH, int sW, int pH, int pW, int quantize, void *stream) {
    Geom2 g;
    int rc = make_geom2(g, N, C, H, W, sH, sW, pH, pW);
    if (rc) return rc;
    if ((rc = check_dtypes(dtype, shift_dtype))) return rc;
    if ((int64_t)N * C * g.Ho * g.Wo == 0) return RB_OK;
    if (!x || !shift || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (use_strip2d(dtype, shift_dtype, g, quantize)) {
        g_last_impl = RB_IMPL_STRIP;
        return shift2d_forward_strip(x, shift, out, dtype, shift_dtype, g, (cudaStream_t)stream);
    }
    if (use_tiled2d(dtype, shift_dtype, g, quantize)) {
        g_last_impl = RB_IMPL_TILED;
        return shift2d_forward_tiled(x, shift, out, dtype, shift_dtype, g, (cudaStream_t)stream);
    }
    g_last_impl = RB_IMPL_GENERIC;
    return shift2d_forward_generic(x, shift, out, dtype, shift_dtype, g, quantize, (cudaStream_t)stream);
}

size_t rb_shift2d_backward_workspace_bytes(int dtype, int N, int C, int H, int W, int sH, int sW, int pH,
                                           int pW) {
    Geom2 g;
    if (make_geom2(g, N, C, H, W, sH, sW, pH, pW)) return 0;
    if ((int64_t)N * C * g.Ho * g.Wo == 0) return 0;
    size_t need = (size_t)C * shift2d_bwd_chunks(g) * 2 * sizeof(double);
    const size_t strip = shift2d_strip_supported(dtype, g, 0) ? shift2d_backward_strip_workspace(dtype, g) : 0;
    if (strip > need) need = strip;
    const size_t tiled = shift2d_tiled_supported(dtype, g, 0) ? shift2d_backward_tiled_workspace(dtype, g) : 0;
    if (tiled > need) need = tiled;
    return (need + 255) & ~(size_t)255;
}

int rb_shift2d_backward(const void *x, const void *shift, const void *out_grad, void *x_grad,
                        void *shift_grad, int dtype, int shift_dtype, int N, int C, int H, int W, int sH,
                        int sW, int pH, int pW, int normalize_grad, int enable_shift_grad, int quantize,
                        void *workspace, size_t workspace_bytes, void *stream) {
    Geom2 g;
    int rc = make_geom2(g, N, C, H, W, sH, sW, pH, pW);
    if (rc) return rc;
    if ((rc = check_dtypes(dtype, shift_dtype))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!enable_shift_grad) shift_grad = nullptr;  // rubiks.cpp:126
    if (!x_grad && !shift_grad) return RB_OK;
    if ((int64_t)N * C * g.Ho * g.Wo == 0) {
        if (x_grad && (int64_t)N * C * H * W > 0)
            cudaMemsetAsync(x_grad, 0, (size_t)N * C * H * W * dtype_size(dtype), s);
        if (shift_grad && C > 0) cudaMemsetAsync(shift_grad, 0, (size_t)2 * C * dtype_size(shift_dtype), s);
        return RB_OK;
    }
    if (!x || !shift || !out_grad) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    const size_t need = rb_shift2d_backward_workspace_bytes(dtype, N, C, H, W, sH, sW, pH, pW);
    if (shift_grad && (!workspace || workspace_bytes < need))
        return fail(RB_ERR_WORKSPACE, "shift2d backward needs %zu workspace bytes, got %zu", need,
                    workspace_bytes);
    if (use_strip2d(dtype, shift_dtype, g, quantize)) {  // input gradient + shift gradient in one launch (+ finalize)
        g_last_impl = RB_IMPL_STRIP;
        return shift2d_backward_strip(x, shift, out_grad, x_grad, shift_grad, dtype, shift_dtype, g, normalize_grad, workspace, s);
    }
    if (use_tiled2d(dtype, shift_dtype, g, quantize)) {
        g_last_impl = RB_IMPL_TILED;
        return shift2d_backward_tiled(x, shift, out_grad, x_grad, shift_grad, dtype, shift_dtype, g, normalize_grad, workspace, s);
    }
    g_last_impl = RB_IMPL_GENERIC;
    if (shift_grad) {
        rc = shift2d_bwd_shift_generic(x, shift, out_grad, shift_grad, dtype, shift_dtype, g, normalize_grad,
                                       (double *)workspace, s);
        if (rc) return rc;
    }
    if (x_grad) rc = shift2d_bwd_input_generic(shift, out_grad, x_grad, dtype, shift_dtype, g, quantize, s);
    return rc;
}

static int check_weight_dtype(int wdt) {
    wdt &= ~RB_W_RESIDENT;
    if (wdt != RB_F32 && wdt != RB_BF16) return fail(RB_ERR_INVALID_ARGUMENT, "weight dtype must be RB_F32 or RB_BF16, got %d", wdt);
    return RB_OK;
}

int rb_pw_conv_forward(const void *x, const void *weight, int weight_dtype, int weight_transposed, const void *residual,
                       void *out, int dtype, int NI, int K, int N, int HW, const float *in_scale_bias, void *stream) {
    if (dtype != RB_BF16) return fail(RB_ERR_UNSUPPORTED, "rb_pw_conv_forward: bf16 activations only (got dtype %d)", dtype);
    if ((weight_dtype & ~RB_W_RESIDENT) == RB_W_IMAGE) weight_dtype = RB_W_IMAGE;
    int rc = weight_dtype == RB_W_IMAGE ? RB_OK : check_weight_dtype(weight_dtype);
    if (rc) return rc;
    if (NI < 0 || K <= 0 || N <= 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d,%d,%d]", NI, K, N, HW);
    if ((int64_t)NI * K * HW > 0x7fffffffLL || (int64_t)NI * N * HW > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    if ((int64_t)NI * HW == 0) return RB_OK;
    if (!x || !weight || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (weight_dtype == RB_W_IMAGE) {
        if (weight_transposed)
            return fail(RB_ERR_INVALID_ARGUMENT, "a packed weight image carries its orientation (rb_pw_weight_image_pack)");
        return pw2_forward(x, weight, residual, out, NI, K, N, HW, in_scale_bias, (cudaStream_t)stream);
    }
    // large maps (16-byte row pitch): the tensor-map TMA schedule (csrc/pw_conv3.cu)
    if (!weight_transposed && pw3_supported(x, out, residual, NI, K, N, HW))
        return pw3_forward(x, weight, weight_dtype, residual, out, NI, K, N, HW, in_scale_bias, (cudaStream_t)stream);
    return pw_conv_forward(x, weight, weight_dtype, weight_transposed != 0, residual, out, NI, K, N, HW, in_scale_bias,
                           nullptr, 0, 0, 0, 0, (cudaStream_t)stream);
}

void rb_pw_conv_tma_set_enabled(int enabled) { pw3_set_enabled(enabled); }

int rb_pw_conv_forward_f32(const float *x, const float *weight, const float *residual, float *out, int NI, int K, int N, int HW,
                           const float *in_scale_bias, const float *out_scale_bias, int out_relu, int flags, void *stream) {
    if (NI < 0 || K <= 0 || N <= 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d,%d,%d]", NI, K, N, HW);
    if ((int64_t)NI * K * HW > 0x7fffffffLL || (int64_t)NI * N * HW > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    if ((int64_t)NI * HW == 0) return RB_OK;
    if (!x || !weight || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(residual)) & 3)
        return fail(RB_ERR_INVALID_ARGUMENT, "rb_pw_conv_forward_f32: pointers must be 4-byte aligned");
    return pw_tf32_forward(x, weight, residual, out, NI, K, N, HW, in_scale_bias, out_scale_bias, out_relu != 0,
                           (flags & RB_W_RESIDENT) != 0, (cudaStream_t)stream);
}

size_t rb_pw_weight_image_bytes(int rows, int contraction) { return pw2_weight_image_bytes(rows, contraction); }

int rb_pw_weight_image_pack(const float *weight, int N, int K, int transposed, void *image, void *stream) {
    if (N <= 0 || K <= 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d]", N, K);
    if (!weight || !image) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (reinterpret_cast<uintptr_t>(image) & 15) return fail(RB_ERR_INVALID_ARGUMENT, "weight image must be 16-byte aligned");
    return pw2_weight_pack(weight, N, K, transposed != 0, image, (cudaStream_t)stream);
}

int rb_pw_weight_pack_multi(const rb_pw_pack_item_t *items_device, int count, void *stream) {
    if (count < 0 || count > 65535) return fail(RB_ERR_INVALID_ARGUMENT, "bad item count %d", count);
    if (count == 0) return RB_OK;
    if (!items_device) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return pw_weight_pack_multi(items_device, count, (cudaStream_t)stream);
}

int rb_pw_weight_image_pack_multi(const rb_pw_pack_item_t *items_device, int count, void *stream) {
    if (count < 0 || count > 65535) return fail(RB_ERR_INVALID_ARGUMENT, "bad item count %d", count);
    if (count == 0) return RB_OK;
    if (!items_device) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return pw2_weight_pack_multi(items_device, count, (cudaStream_t)stream);
}

int rb_pw_conv_image_supported(int NI, int K, int N, int HW, int has_in_scale_bias) {
    return pw2_supported(NI, K, N, HW, has_in_scale_bias);
}

int rb_pw_conv_forward_stats(const void *x, const void *weight, int weight_dtype, int weight_transposed, const void *residual,
                             void *out, int dtype, int NI, int K, int N, int HW, const float *in_scale_bias,
                             double *stats_partial, size_t stats_bytes, int *stats_splits, void *stream) {
    if (dtype != RB_BF16) return fail(RB_ERR_UNSUPPORTED, "rb_pw_conv_forward_stats: bf16 activations only (got dtype %d)", dtype);
    int rc = check_weight_dtype(weight_dtype);
    if (rc) return rc;
    if (NI <= 0 || K <= 0 || N <= 0 || HW <= 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d,%d,%d]", NI, K, N, HW);
    if ((int64_t)NI * K * HW > 0x7fffffffLL || (int64_t)NI * N * HW > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    if (!x || !weight || !out || !stats_partial || !stats_splits) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (!weight_transposed && pw3_supported(x, out, residual, NI, K, N, HW))  // thread-local sums in the TMA schedule's epilogue
        return pw3_forward(x, weight, weight_dtype, residual, out, NI, K, N, HW, in_scale_bias, (cudaStream_t)stream, stats_partial,
                           stats_bytes, stats_splits);
    return pw_conv_forward(x, weight, weight_dtype, weight_transposed != 0, residual, out, NI, K, N, HW, in_scale_bias,
                           nullptr, 0, 0, 0, 0, (cudaStream_t)stream, stats_partial, stats_bytes, stats_splits);
}

int rb_pw_conv_stats_preferred(int NI, int K, int N, int HW) {
    if (NI <= 0 || K <= 0 || N <= 0 || HW <= 0) return 0;
    return pw3_stats_preferred(NI, K, N, HW) ? 1 : 0;
}

int rb_bn_stats_finalize(const double *partial, int splits, int C, double count, const float *gamma, const float *beta,
                         float *running_mean, float *running_var, float momentum, float eps, float *mean_invstd,
                         float *scale_bias, void *stream) {
    if (splits <= 0 || C <= 0 || !(count > 0)) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent");
    if (!partial || !mean_invstd || !scale_bias) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return bn_stats_finalize(partial, splits, C, count, gamma, beta, running_mean, running_var, momentum, eps, mean_invstd,
                             scale_bias, (cudaStream_t)stream);
}

int rb_bn_apply_forward(const void *x, const float *scale_bias, void *y, int dtype, int NI, int C, int HW, int relu, void *stream) {
    if (NI < 0 || C < 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    if ((int64_t)NI * C * HW == 0) return RB_OK;
    if ((int64_t)NI * C * HW > 0x7fffffffLL) return fail(RB_ERR_UNSUPPORTED, "bn: tensor too large");
    if (!x || !scale_bias || !y) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return bn_apply_forward(x, scale_bias, y, dtype, NI, C, HW, relu, (cudaStream_t)stream);
}

int rb_pw_weight_pack(const float *weight, void *weight_nk, void *weight_kn, int N, int K, void *stream) {
    if (N <= 0 || K <= 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d]", N, K);
    if (!weight || !weight_nk || !weight_kn) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return pw_weight_pack(weight, weight_nk, weight_kn, N, K, (cudaStream_t)stream);
}

int rb_shift3d_pw_conv_forward(const void *x, const void *shift, const void *weight, int weight_dtype, const void *residual,
                               void *out, int dtype, int shift_dtype, int N, int T, int C, int H, int W, int Cout,
                               void *stream) {
    if (dtype != RB_BF16)
        return fail(RB_ERR_UNSUPPORTED, "rb_shift3d_pw_conv_forward: bf16 activations only (got dtype %d)", dtype);
    int rc = check_dtypes(dtype, shift_dtype);
    if (rc) return rc;
    if ((rc = check_weight_dtype(weight_dtype))) return rc;
    if (N < 0 || T < 0 || C <= 0 || H < 0 || W < 0 || Cout <= 0)
        return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d,%d,%d,%d] -> %d", N, T, C, H, W, Cout);
    if ((int64_t)N * T * C * H * W > 0x7fffffffLL || (int64_t)N * T * Cout * H * W > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    if ((int64_t)N * T * H * W == 0) return RB_OK;
    if (!x || !shift || !weight || !out) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    return pw_conv_forward(x, weight, weight_dtype, 0, residual, out, N * T, C, Cout, H * W, nullptr, shift, shift_dtype, T,
                           H, W, (cudaStream_t)stream);
}

size_t rb_pw_conv_wgrad_workspace_bytes(int NI, int K, int N, int HW) {
    if (NI <= 0 || K <= 0 || N <= 0 || HW <= 0) return 0;
    const size_t first = pw_conv_wgrad_workspace(NI, N, K, HW), tma = wg3_workspace(NI, N, K, HW);
    return ((first > tma ? first : tma) + 255) & ~(size_t)255;
}

static int wgrad_common(const void *out_grad, const void *x, float *weight_grad, int dtype, int NI, int K, int N, int HW,
                        const float *in_scale_bias, const void *shift, int shift_dtype, int T, int H, int W,
                        void *workspace, size_t workspace_bytes, void *stream) {
    if (dtype != RB_BF16) return fail(RB_ERR_UNSUPPORTED, "pw_conv wgrad: bf16 activations only (got dtype %d)", dtype);
    if (NI < 0 || K <= 0 || N <= 0 || HW < 0) return fail(RB_ERR_INVALID_ARGUMENT, "bad extent [%d,%d,%d,%d]", NI, K, N, HW);
    if ((int64_t)NI * K * HW > 0x7fffffffLL || (int64_t)NI * N * HW > 0x7fffffffLL)
        return fail(RB_ERR_UNSUPPORTED, "tensors with more than 2^31-1 elements are not supported");
    if (!weight_grad) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if ((int64_t)NI * HW == 0) {
        cudaMemsetAsync(weight_grad, 0, (size_t)N * K * sizeof(float), s);
        return RB_OK;
    }
    if (!out_grad || !x) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    const size_t need = rb_pw_conv_wgrad_workspace_bytes(NI, K, N, HW);
    if (!workspace || workspace_bytes < need)
        return fail(RB_ERR_WORKSPACE, "pw_conv wgrad needs %zu workspace bytes, got %zu", need, workspace_bytes);
    // GEMM rows m = output channels (N of the conv), columns = input channels (K of the conv)
    if (!shift && wg3_supported(out_grad, x, NI, N, K, HW))  // both operands through tensor-map TMA (csrc/pw_wgrad3.cu)
        return wg3_run(out_grad, x, weight_grad, NI, N, K, HW, in_scale_bias, workspace, s);
    return pw_conv_wgrad(out_grad, x, weight_grad, NI, N, K, HW, in_scale_bias, shift, shift_dtype, T, H, W, workspace, s);
}

int rb_pw_conv_wgrad(const void *out_grad, const void *x, float *weight_grad, int dtype, int NI, int K, int N, int HW,
                     const float *in_scale_bias, void *workspace, size_t workspace_bytes, void *stream) {
    return wgrad_common(out_grad, x, weight_grad, dtype, NI, K, N, HW, in_scale_bias, nullptr, 0, 0, 0, 0, workspace,
                        workspace_bytes, stream);
}

int rb_shift3d_pw_conv_wgrad(const void *out_grad, const void *x, const void *shift, float *weight_grad, int dtype,
                             int shift_dtype, int N, int T, int C, int H, int W, int Cout, void *workspace,
                             size_t workspace_bytes, void *stream) {
    int rc = check_dtypes(dtype, shift_dtype);
    if (rc) return rc;
    if (!shift) return fail(RB_ERR_INVALID_ARGUMENT, "null pointer");
    if (N < 0 || T < 0 || H < 0 || W < 0) return fail(RB_ERR_INVALID_ARGUMENT, "negative extent");
    return wgrad_common(out_grad, x, weight_grad, dtype, N * T, C, Cout, H * W, nullptr, shift, shift_dtype, T, H, W,
                        workspace, workspace_bytes, stream);
}

void rb_pw_conv_set_tuning(int min_n_splits) { pw_conv_set_tuning(min_n_splits); }
void rb_pw_conv_wgrad_set_tuning(int burst, int l2_256, int max_stages) { wg3_set_tuning(burst, l2_256, max_stages); }
void rb_pw_conv_image_set_tuning(int operand_stages, int k_chunk, int wait_hint_ns) {
    pw2_set_tuning(operand_stages, k_chunk, wait_hint_ns);
}

#ifdef RB_DEBUG_TRACE
/* debug builds only: work-skipping switches of the image kernel (critical-path experiments, tools/trace_pw.py --dbg) */
void rb_debug_pw_flags(int flags) { pw2_set_debug(flags); }
/* debug builds only: device buffer (128 x uint64 per CTA) receiving globaltimer stamps of k_pw_conv; NULL = off */
void rb_debug_pw_trace(void *device_buffer) { pw_conv_set_trace(device_buffer); }
#endif

}  // extern "C"

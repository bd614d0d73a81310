/*
 * rubiks_b200.h -- C ABI of librubiks_b200.so: RubiksNet's learnable-shift hot path for B200 (sm_100a).
 *
 * This is the drop-in boundary.  Each entry point replaces one function of the reference's native
 * module `rubiksnet_cuda` (pybind11, /root/reference/cuda_src/rubiks.cpp:384-396); the citation next
 * to each prototype names the reference interface it stands in for.  Plain pointers and sizes only:
 * no torch / ATen types cross this boundary (INTEGRATION.md shows the reference-side binding).
 *
 * Conventions
 *  - All tensor pointers are DEVICE pointers to contiguous memory of the given dtype.
 *      3D: x [N,T,C,H,W], shift [3,C] rows (T,H,W), out / out_grad [N,To,C,Ho,Wo]
 *      2D: x [N,C,H,W],   shift [2,C] rows (H,W),   out / out_grad [N,C,Ho,Wo]
 *      output extent per axis = (in + 2*pad - 1) / stride + 1     (rubiks.cpp:14-30,161-178)
 *  - `dtype` is the activation type, `shift_dtype` the type of shift AND shift_grad (RB_F32 when
 *    activations are bf16/f16 and the parameter stays fp32; the reference requires them equal:
 *    rubiksnet/shiftlib/rubiks3d/primitive.py:65).
 *  - `stream` is a cudaStream_t (NULL = legacy default stream).  Every launch goes to that stream;
 *    nothing synchronises.  The caller selects the device (cudaSetDevice) before calling.
 *  - Every output element is written: callers need not pre-zero buffers (the reference's Python side
 *    does, rubiksnet/utils.py:25-26; elements it would have left at 0 are written as 0 here).
 *  - shift_grad is OVERWRITTEN, as in the reference (addmv_ with beta = 0, rubiks.cpp:344-345).
 *  - Return value: RB_OK (0) on success -- the reference's entry points always return 0 and the
 *    Python side asserts that (primitive.py:79,139).  Invalid arguments and CUDA launch errors
 *    return a non-zero rb_status_t; rb_last_error() gives the message (the reference throws a C++
 *    exception -> RuntimeError, which is what the Python mirror raises on non-zero).
 */
#ifndef RUBIKS_B200_H_
#define RUBIKS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 1

typedef enum { RB_F32 = 0, RB_F64 = 1, RB_F16 = 2, RB_BF16 = 3 } rb_dtype_t;

typedef enum {
    RB_OK = 0,
    RB_ERR_INVALID_ARGUMENT = 1, /* bad dtype / non-positive extent or stride / null pointer */
    RB_ERR_UNSUPPORTED = 2,      /* combination not implemented (message says which) */
    RB_ERR_CUDA = 3,             /* cudaGetLastError() after a launch was not cudaSuccess */
    RB_ERR_WORKSPACE = 4         /* workspace pointer null or smaller than *_workspace_bytes() */
} rb_status_t;

/* which 3D-shift implementation the dispatcher may use.  RB_IMPL_AUTO picks, in this order, the strip
 * kernels (spatial stride 1), the tiled kernels (spatial stride 1 or 2) -- both TMA-staged sm_100a kernels
 * for temporal stride 1 / no padding / no quantize -- and the generic gather kernels otherwise */
typedef enum { RB_IMPL_AUTO = 0, RB_IMPL_GENERIC = 1, RB_IMPL_TILED = 2, RB_IMPL_STRIP = 3 } rb_impl_t;

int rb_abi_version(void);
/* message of the last non-zero status returned on this thread ("" if none) */
const char *rb_last_error(void);
/* number of kernels this library has launched since load / the last reset (all threads) */
uint64_t rb_launch_count(void);
void rb_launch_count_reset(void);
/* force an implementation for subsequent calls from any thread (tests / benchmarks) */
void rb_set_impl(int impl);
/* implementation the last forward / backward call on this thread actually used (rb_impl_t) */
int rb_last_impl(void);

/* compute_output_shape, rubiks.cpp:161-178 (and compute_output_len :14-30) */
int rb_out_len(int in_len, int stride, int pad);

/* ---------------------------------------------------------------- 3D learnable shift ---------- */

/* replaces rubiks_shift_3d_forward<T> (rubiks.cpp:181-253) -> RubiksShift3DForward<T>
 * (rubiks3d_kernels.cu:1002-1038) -> rubiks_shift_3d_forward_cuda (:15-205).
 * strides / pads are (T,H,W). */
int rb_shift3d_forward(const void *x, const void *shift, void *out, int dtype, int shift_dtype,
                       int N, int T, int C, int H, int W, int sT, int sH, int sW, int pT, int pH,
                       int pW, int quantize, void *stream);

/* scratch bytes rb_shift3d_backward needs (replaces the torch::zeros({3C,Ho,Wo}) + torch::ones
 * temporaries of rubiks.cpp:294-299) */
size_t rb_shift3d_backward_workspace_bytes(int dtype, int N, int T, int C, int H, int W, int sT,
                                           int sH, int sW, int pT, int pH, int pW);

/* replaces rubiks_shift_3d_backward<T> (rubiks.cpp:256-379): shift-grad kernel
 * (rubiks3d_kernels.cu:218-452) + addmv_ reduction (rubiks.cpp:344-345) + optional
 * normalisation (:932-960) + input-grad kernel (:455-929).
 * x_grad and/or shift_grad may be NULL to skip that output.  normalize_t_factor < 0 selects the
 * temporal-only normalisation.  `quantize` affects x_grad only (as in the reference). */
int rb_shift3d_backward(const void *x, const void *shift, const void *out_grad, void *x_grad,
                        void *shift_grad, int dtype, int shift_dtype, int N, int T, int C, int H,
                        int W, int sT, int sH, int sW, int pT, int pH, int pW, int normalize_grad,
                        double normalize_t_factor, int quantize, void *workspace,
                        size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- 2D learnable shift ---------- */

/* replaces rubiks2d_forward (rubiks.cpp:44-67) -> rubiks2d_forward_kernel
 * (rubiks2d_kernels.cu:94-145).  strides / pads are (H,W). */
int rb_shift2d_forward(const void *x, const void *shift, void *out, int dtype, int shift_dtype,
                       int N, int C, int H, int W, int sH, int sW, int pH, int pW, int quantize,
                       void *stream);

size_t rb_shift2d_backward_workspace_bytes(int dtype, int N, int C, int H, int W, int sH, int sW,
                                           int pH, int pW);

/* replaces rubiks2d_backward (rubiks.cpp:94-155): rubiks2d_backward_shift_kernel
 * (rubiks2d_kernels.cu:147-266) + addmv_ + rubiks2d_normalize_shift_grad_kernel (:381-397) +
 * rubiks2d_backward_input_kernel (:269-379).  enable_shift_grad = 0 leaves shift_grad untouched
 * (rubiks.cpp:126). */
int rb_shift2d_backward(const void *x, const void *shift, const void *out_grad, void *x_grad,
                        void *shift_grad, int dtype, int shift_dtype, int N, int C, int H, int W,
                        int sH, int sW, int pH, int pW, int normalize_grad, int enable_shift_grad,
                        int quantize, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- AttentionShift -------------- */

/* The 3-tap temporal mix of rubiksnet/attention_shift.py:32-39 (the depth-wise F.conv1d with
 * C*H*W groups): out[n,t,c,:] = taps[c,0] x[n,t-1,c,:] + taps[c,1] x[n,t,c,:] + taps[c,2] x[n,t+1,c,:]
 * with zeros beyond the clip.  x / out are [N*T, C, HW] contiguous, taps is fp32 [C,3] (the softmax
 * of :29-30 is a [C,3] computation done by the host mirror). */
int rb_attention_shift_forward(const void *x, const float *taps, void *out, int dtype, int N, int T,
                               int C, int HW, void *stream);

size_t rb_attention_shift_backward_workspace_bytes(int N, int T, int C, int HW);

/* adjoint of the above: x_grad [N*T,C,HW] and taps_grad fp32 [C,3] (overwritten); either may be
 * NULL. */
int rb_attention_shift_backward(const void *x, const float *taps, const void *out_grad, void *x_grad,
                                float *taps_grad, int dtype, int N, int T, int C, int HW,
                                void *workspace, size_t workspace_bytes, void *stream);

/* The same pair with the block's bn1 -> relu folded into the load (rubiksnet/backbone.py:123-125 in front of the
 * AttentionShift of models.py:100-104): the mixed input is relu(x * scale[c] + bias[c]) rounded to `dtype`, with
 * in_scale_bias fp32 [C,2] as written by rb_bn_act_forward -- the normalised tensor never exists in memory.
 * x_grad is the gradient with respect to that normalised input (what rb_bn_act_backward takes as dy). */
int rb_bn_attention_shift_forward(const void *x, const float *in_scale_bias, const float *taps, void *out, int dtype,
                                  int N, int T, int C, int HW, void *stream);
int rb_bn_attention_shift_backward(const void *x, const float *in_scale_bias, const float *taps, const void *out_grad,
                                   void *x_grad, float *taps_grad, int dtype, int N, int T, int C, int HW,
                                   void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- BatchNorm (+ReLU) ------------ */

/* The bn1->relu / bn2->relu / bn_last->relu stages of RubiksShiftBlock / RubiksNetBackbone
 * (rubiksnet/backbone.py:50-53,123-135,196), which the reference delegates to nn.BatchNorm2d + nn.ReLU.
 * x / y / dy / dx / residual are [NI, C, HW] contiguous (NCHW) with 16-byte aligned base pointers (checked:
 * RB_ERR_INVALID_ARGUMENT otherwise); gamma / beta / running_mean / running_var and all statistics are FP32
 * arrays of C elements -- a caller holding 16-bit BatchNorm buffers (model.half()) must pass fp32 copies.
 *   mean_invstd [C,2] and scale_bias [C,2] are OUTPUTS of the forward pass that the backward pass
 *   (and the fused shift) consume: y = act(x * scale + bias), scale = gamma * invstd,
 *   bias = beta - mean * scale.
 * training != 0: batch statistics (biased variance), running stats updated with `momentum`
 * (unbiased variance), exactly torch.nn.BatchNorm2d.  training == 0: running statistics.
 * relu != 0 applies max(.,0) in forward and masks dy with (y > 0) in backward.
 * rb_bn_act_forward with y == NULL computes only the statistics / coefficients (the apply pass is then
 * folded into a consumer, e.g. rb_pw_conv_forward's in_scale_bias). */
size_t rb_bn_workspace_bytes(int NI, int C);

int rb_bn_act_forward(const void *x, const float *gamma, const float *beta, float *running_mean,
                      float *running_var, void *y, float *mean_invstd, float *scale_bias, int dtype,
                      int NI, int C, int HW, int training, float momentum, float eps, int relu,
                      void *workspace, size_t workspace_bytes, void *stream);

/* dx = dL/dx (+ residual if non-NULL: same shape as x, e.g. the identity-shortcut gradient),
 * dgamma / dbeta fp32 [C] (overwritten; may be NULL).  dx may be NULL (parameter gradients only). */
int rb_bn_act_backward(const void *x, const void *dy, const void *residual, const float *gamma,
                       const float *mean_invstd, const float *scale_bias, void *dx, float *dgamma,
                       float *dbeta, int dtype, int NI, int C, int HW, int training, int relu,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- squeeze-and-excitation ------- */

/* The two passes over the activation tensor of SELayer (rubiksnet/backbone.py:56-71; tier "small") and their gradients:
 *   rb_plane_reduce: out[plane] = scale * sum_p a[plane, p] * (b ? b[plane, p] : 1)   fp32 out, plane = (image, channel)
 *                    (avg pool: b = NULL, scale = 1/HW; gate gradient: a = dL/dy, b = x, scale = 1)
 *   rb_plane_scale:  out[plane, p] = a[plane, p] * s[plane] + (t ? t[plane] : 0)
 *                    (y = x * gate;  dL/dx = dL/dy * gate + dpool / HW)
 * a / b / out are [planes, HW] contiguous in fp32 / fp16 / bf16; s, t and the reduction result are fp32. */
int rb_plane_reduce(const void *a, const void *b, float *out, int dtype, int planes, int HW, float scale, void *stream);
int rb_plane_scale(const void *a, const float *s, const float *t, void *out, int dtype, int planes, int HW, void *stream);

/* ---------------------------------------------------------------- pointwise (1x1) convolutions -- */

/* Conv1x1 of RubiksShiftBlock (rubiksnet/backbone.py:45-47: nn.Conv2d(k=1, bias=False); conv2 / conv3 /
 * shortcut at :123-135) as ONE tcgen05 tensor-core GEMM launch on NCHW bf16 activations:
 *     out[i, n, p] = sum_k weight[n, k] * A[i, k, p]  (+ residual[i, n, p])
 * x [NI, K, HW], residual / out [NI, N, HW]; dtype must be RB_BF16 (fp32 accumulation in tensor memory).
 * weight is the Conv2d parameter [N, K, 1, 1] in weight_dtype RB_F32 (the fp32 master copy, rounded to bf16
 * by the kernel) or RB_BF16.  weight_transposed != 0: the buffer holds [K, N] instead -- passing a conv's own
 * weight with N and K swapped computes that conv's INPUT GRADIENT (out_grad -> x_grad) with the same kernel.
 * residual may be NULL; it may alias out.
 * in_scale_bias (fp32 [K, 2] = the scale_bias output of rb_bn_act_forward, or NULL) folds
 * A = relu(x * scale[k] + bias[k]) -- the bn1 -> relu in front of conv2 (backbone.py:123,127) -- into the
 * operand producer; NULL: A = x. */
int rb_pw_conv_forward(const void *x, const void *weight, int weight_dtype, int weight_transposed,
                       const void *residual, void *out, int dtype, int NI, int K, int N, int HW,
                       const float *in_scale_bias, void *stream);

/* rb_pw_conv_forward that also reduces the BatchNorm statistics of its own output in the GEMM epilogue (the block
 * feeds every conv output into a BatchNorm: conv2 -> bn2, conv3 + shortcut -> the next block's bn1, backbone.py:123-135),
 * so the separate statistics pass over the tensor disappears.  stats_partial receives per channel and split the fp64
 * pair (sum, sum of squares) of the stored bf16 values, laid out [N][*stats_splits][2]; *stats_splits (<= 2 x SM count) is
 * written by the call; stats_bytes >= N * 2 * 148 * 16 always suffices.  Feed both to rb_bn_stats_finalize. */
int rb_pw_conv_forward_stats(const void *x, const void *weight, int weight_dtype, int weight_transposed,
                             const void *residual, void *out, int dtype, int NI, int K, int N, int HW,
                             const float *in_scale_bias, double *stats_partial, size_t stats_bytes, int *stats_splits,
                             void *stream);
/* 1 when the statistics epilogue costs (almost) nothing for this geometry -- the tensor-map schedule, where an epilogue thread
 * owns a whole channel row -- i.e. when a caller should prefer rb_pw_conv_forward_stats over a separate statistics pass. */
int rb_pw_conv_stats_preferred(int NI, int K, int N, int HW);

/* 1x1 convolution on fp32 activations as a tcgen05 kind::tf32 GEMM (csrc/pw_conv_tf32.cu): the fp32 inference path of
 * RubiksShiftBlock (rubiksnet/backbone.py:123-135; the reference runs cuDNN TF32 convolutions between separate BatchNorm /
 * ReLU passes).  x [NI, K, HW], weight [N, K] fp32 (the Conv2d parameter), residual / out [NI, N, HW] fp32:
 *     out = [relu]( conv(A) * out_scale[n] + out_bias[n] ) + residual,   A = x  or  relu(x * in_scale[k] + in_bias[k])
 * in_scale_bias [K, 2] / out_scale_bias [N, 2] are eval-mode BatchNorm coefficients as produced by rb_bn_act_forward
 * (training = 0); each may be NULL, as may residual (which may alias out).  Operands are rounded to TF32 (10-bit
 * mantissa), accumulation and everything after it is fp32.  flags: RB_W_RESIDENT or 0. */
int rb_pw_conv_forward_f32(const float *x, const float *weight, const float *residual, float *out, int NI, int K, int N,
                           int HW, const float *in_scale_bias, const float *out_scale_bias, int out_relu, int flags,
                           void *stream);

/* Input pipeline on the GPU (rubiksnet/transforms.py:66-79,329-363: Stack -> ToTorchFormatTensor -> GroupNormalize, which the
 * reference runs on the CPU per sample): frames_u8 [N, H, W, channels] uint8 (channels = 3 * frames, RGB-minor, as `Stack`
 * concatenates them) -> out [N, channels, H, W] (= [N * frames, 3, H, W], the model input) in out_dtype RB_F32 / RB_BF16,
 *     out[n, c, h, w] = ((div255 ? x / 255 : x) - mean3[c % 3]) / std3[c % 3].
 * mean3 / std3 are HOST arrays of 3 floats (RubiksNet.input_mean / input_std). */
int rb_frames_to_clip(const void *frames_u8, void *out, int out_dtype, int N, int H, int W, int channels, const float *mean3,
                      const float *std3, int div255, void *stream);

/* Patch matrix of the 3x3 / stride-S / padding-1 first convolution (rubiksnet/backbone.py:148-149): cols [NI, Tpad, Ho, Wo],
 * cols[i, (ci*3 + kh)*3 + kw, ho, wo] = x[i, ci, ho*S - 1 + kh, wo*S - 1 + kw] (0 outside the image), rows >= 9*Cin zero.
 * conv1(x) = rb_pw_conv_forward / rb_pw_conv_forward_f32 (cols, weight viewed [Cout, 9*Cin] and zero-padded to Tpad);
 * its weight gradient = rb_pw_conv_wgrad(out_grad, cols).  x [NI, Cin, H, W] in in_dtype (RB_F32 / RB_BF16), cols in
 * out_dtype (RB_BF16, or RB_F32 from RB_F32); Wo must be a multiple of 8; cols 16-byte aligned. */
int rb_im2col3x3(const void *x, void *cols, int in_dtype, int out_dtype, int NI, int Cin, int H, int W, int stride, int Tpad,
                 void *stream);

/* BatchNorm passes on small maps (a whole channel of a 16-bit tensor fits one CTA's shared memory: NI * HW * 2 bytes <= 216 KiB
 * for the forward, twice that for the backward, and C >= 128) run as ONE channel-resident launch instead of the reduce /
 * finalize / apply sequence.  1 (default) = on, 0 = always the streaming passes; process-global, for A/B measurements. */
void rb_bn_set_resident(int enabled);

/* Batch statistics -> (mean, invstd), (scale, bias) and the running-statistics update of nn.BatchNorm2d in training
 * mode, from partial sums [C][splits][2] over `count` elements per channel (what rb_bn_act_forward does after its own
 * reduction pass).  running_mean / running_var may be NULL. */
int rb_bn_stats_finalize(const double *partial, int splits, int C, double count, const float *gamma, const float *beta,
                         float *running_mean, float *running_var, float momentum, float eps, float *mean_invstd,
                         float *scale_bias, void *stream);

/* y = relu?(x * scale[c] + bias[c]): the apply pass of rb_bn_act_forward alone, for coefficients obtained from
 * rb_bn_stats_finalize. */
int rb_bn_apply_forward(const void *x, const float *scale_bias, void *y, int dtype, int NI, int C, int HW, int relu,
                        void *stream);

/* Per-step preparation of a conv weight for the two GEMMs above: the fp32 master [N, K] (nn.Conv2d(k=1).weight,
 * backbone.py:45-47) is rounded to bf16 once, in both orientations -- weight_nk [N, K] for the forward
 * (rb_pw_conv_forward(..., weight_nk, RB_BF16, 0, ...)) and weight_kn [K, N], which is the [N'=K, K'=N] weight
 * matrix of the input gradient (rb_pw_conv_forward(out_grad, weight_kn, RB_BF16, 0, ..., K'=N, N'=K)).  Same
 * rounding as the kernels apply when handed the fp32 master; saves every CTA the conversion (and, for the input
 * gradient, a transposing scatter) of the whole weight block.  One launch. */
int rb_pw_weight_pack(const float *weight, void *weight_nk, void *weight_kn, int N, int K, void *stream);

/* as3 -> conv3 -> `out += shortcut` of RubiksShiftBlock.forward (backbone.py:129-135, with as3 the
 * _Rubiks3DWrap of rubiksnet/models.py:128-145) in ONE launch: the 3D learnable shift
 * (rubiks_shift_3d_forward_cuda, rubiks3d_kernels.cu:15-205; stride (1,1,1), padding 0, no quantize) is the
 * A-operand producer of the tensor-core GEMM, so the shifted tensor never exists in HBM.
 * x [N, T, C, H, W] bf16, shift [3, C] (shift_dtype), weight [Cout, C] (weight_dtype), residual / out
 * [N*T, Cout, H, W] bf16.  Same arithmetic as rb_shift3d_forward followed by rb_pw_conv_forward: fp32 trilinear
 * interpolation in the reference's association order, rounded to bf16 once, then the GEMM. */
int rb_shift3d_pw_conv_forward(const void *x, const void *shift, const void *weight, int weight_dtype,
                               const void *residual, void *out, int dtype, int shift_dtype, int N, int T,
                               int C, int H, int W, int Cout, void *stream);

/* Weight gradient of the same 1x1 convolution (what autograd computes for nn.Conv2d(k=1) in the reference's
 * backward pass): weight_grad[n, k] = sum_{i,p} out_grad[i, n, p] * A[i, k, p], fp32 [N, K], OVERWRITTEN.
 * A is recomputed from x by the operand producer exactly as in the forward (in_scale_bias as above).
 * One tensor-core launch that reduces disjoint pixel ranges into fp32 partial slices in `workspace`, plus a
 * fixed-order reduction over the slices (deterministic). */
size_t rb_pw_conv_wgrad_workspace_bytes(int NI, int K, int N, int HW);
int rb_pw_conv_wgrad(const void *out_grad, const void *x, float *weight_grad, int dtype, int NI, int K, int N,
                     int HW, const float *in_scale_bias, void *workspace, size_t workspace_bytes,
                     void *stream);

/* conv3 weight gradient with the 3D shift recomputed in the operand producer (the shifted tensor was never
 * stored by rb_shift3d_pw_conv_forward): weight_grad[co, c] = sum out_grad[i, co, p] * shift3d(x)[i, c, p]. */
int rb_shift3d_pw_conv_wgrad(const void *out_grad, const void *x, const void *shift, float *weight_grad,
                             int dtype, int shift_dtype, int N, int T, int C, int H, int W, int Cout,
                             void *workspace, size_t workspace_bytes, void *stream);

/* Second-generation pointwise-conv path (csrc/pw_conv2.cu: all global traffic through TMA bulk copies, output channels
 * split over grid.y, whole-image tiles on 14x14 / 7x7 maps).  The weight is handed over as a PACKED IMAGE -- the bf16
 * shared-memory layout of every output-channel slice, so that a CTA fetches its slice with one bulk copy:
 *   rb_pw_weight_image_bytes(rows, contraction)  size of the image of a [rows x contraction] matrix;
 *   rb_pw_weight_image_pack(weight fp32 [N, K], N, K, transposed, image)
 *        transposed == 0: image of W   (rows = N, contraction = K)  -> the forward GEMM of the conv;
 *        transposed == 1: image of W^T (rows = K, contraction = N)  -> its input-gradient GEMM;
 *   rb_pw_conv_forward(x, image, RB_W_IMAGE, 0, ...)  runs it: K = contraction, N = rows of the packed matrix;
 *   rb_pw_conv_image_supported(NI, K, N, HW, has_in_scale_bias): 0 = this geometry has no image path (call
 *        rb_pw_conv_forward with a plain fp32 / bf16 weight), 1 = supported, 2 = supported and measured faster than the
 *        plain path (several whole images per tile: maps of <= 112 pixels, e.g. 7x7).  Needs K % 8 == 0, N % 8 == 0 and
 *        HW <= 224, or HW % 8 == 0 with a divisor in [64, 256] that is a multiple of 8.
 * Same arithmetic as the plain path (bf16 operands, fp32 accumulation, bf16 result, `+= residual` on the rounded value). */
#define RB_W_IMAGE 16
/* OR-ed into weight_dtype (RB_F32 / RB_BF16): the weight buffer was written by rb_pw_weight_pack* (which fence
 * themselves) or at least two launches before this call on `stream`.  Every kernel of the library is launched with
 * programmatic stream serialization: it sets up barriers / tensor memory while its predecessor in the stream drains and
 * waits for that predecessor (griddepcontrol.wait) before touching its results; a resident weight block is staged
 * before that wait as well.  Packed images (RB_W_IMAGE) are always resident.  Without the flag nothing is read early. */
#define RB_W_RESIDENT 256
/* Programmatic dependent launch on (default, 1) / off (0): process-global switch for A/B measurements. */
void rb_set_dependent_launch(int enabled);
size_t rb_pw_weight_image_bytes(int rows, int contraction);
int rb_pw_weight_image_pack(const float *weight, int N, int K, int transposed, void *image, void *stream);
int rb_pw_conv_image_supported(int NI, int K, int N, int HW, int has_in_scale_bias);
/* Both images of MANY conv weights in one launch (once per training step, after the optimizer update): `items_device` is
 * an array in DEVICE memory; image_fwd / image_bwd are 16-byte aligned buffers of rb_pw_weight_image_bytes(N, K) /
 * rb_pw_weight_image_bytes(K, N) bytes. */
typedef struct rb_pw_pack_item {
    const float *weight; /* fp32 [N, K] */
    void *image_fwd;     /* image of W   (rows = N, contraction = K) */
    void *image_bwd;     /* image of W^T (rows = K, contraction = N) */
    int N, K;
} rb_pw_pack_item_t;
int rb_pw_weight_image_pack_multi(const rb_pw_pack_item_t *items_device, int count, void *stream);
/* The plain copies of rb_pw_weight_pack for many weights in one launch: image_fwd = weight_nk [N,K] bf16,
 * image_bwd = weight_kn [K,N] bf16. */
int rb_pw_weight_pack_multi(const rb_pw_pack_item_t *items_device, int count, void *stream);

/* rb_pw_conv_forward runs plain-producer convolutions (in_scale_bias == NULL, weight not transposed) on maps whose plane
 * size is a multiple of 8 pixels (16-byte row pitch: 112x112, 56x56, 28x28) with 16-byte aligned tensors on the
 * tensor-map TMA kernel (csrc/pw_conv3.cu: the TMA unit writes the UMMA operand layout and stores the output tile, no
 * register round trip).  1 (default) = on, 0 = always the first kernel; process-global, for A/B measurements and tests. */
void rb_pw_conv_tma_set_enabled(int enabled);

/* Tiling override for rb_pw_conv_forward (process-global, like rb_set_impl): lower bound on the number of
 * output-channel splits (grid.y); more splits = a smaller resident weight block and a deeper activation ring per CTA
 * at the price of reading the activations once per split (from L2).  0 = automatic (default).  Changes the schedule
 * only, never the arithmetic. */
void rb_pw_conv_set_tuning(int min_n_splits);
/* tensor-map weight gradient (csrc/pw_wgrad3.cu), measurement knobs: pixel chunks issued together (1..4), 256-byte L2
 * promotion in the tensor maps (0/1), ring depth cap (2..8; 0 = default) */
void rb_pw_conv_wgrad_set_tuning(int burst, int l2_256, int max_stages);
/* Same for the image kernel: depth of the operand ring (2..8), channels per K chunk (16 / 32) and the suspend-time hint
 * (ns) of its mbarrier waits; 0 = automatic / hardware default. */
void rb_pw_conv_image_set_tuning(int operand_stages, int k_chunk, int wait_hint_ns);

#ifdef RB_DEBUG_TRACE
/* Debug builds only (python -m rubiksnet_b200.build --trace; tools/trace_pw.py): when non-NULL, every k_pw_conv CTA
 * writes globaltimer stamps of its pipeline events (128 x uint64 per CTA) into this device buffer.  The product
 * library does not export this symbol and contains no tracing or work-skipping code. */
void rb_debug_pw_trace(void *device_buffer);
void rb_debug_pw_flags(int flags); /* image kernel: skip relayout (1) / MMA (2) / epilogue (4) / stores (8) / raw loads (16) */
#endif

#ifdef __cplusplus
}
#endif
#endif /* RUBIKS_B200_H_ */

// Pointwise (1x1) convolutions of RubiksShiftBlock as tcgen05 tensor-core GEMMs on NCHW bf16 activations, with the
// operation in front of the convolution folded into the kernel's activation-operand producer:
//
//     out[i, n, p] = sum_k  W[n, k] * A(i, k, p)   (+ residual[i, n, p])          i = image (clip*T + t), p = pixel
//
//     PROD_PLAIN   A = x                                   conv2 / shortcut / the input gradient of any 1x1 conv
//     PROD_BNRELU  A = relu(x * scale[k] + bias[k])          bn1 -> relu -> conv2          (backbone.py:123-128)
//     PROD_SHIFT3D A = RubiksShift3D(x)[i, k, p]             as3 -> conv3 -> += shortcut   (backbone.py:129-135)
//                  trilinear gather of cuda_src/rubiks3d_kernels.cu:54-203 (stride 1, pad 0) in fp32, rounded to
//                  bf16 once -- what the stand-alone shift kernel would have stored; the shifted tensor never
//                  exists in HBM.
//
// k_pw_conv (forward / input gradient): GEMM view per tile  D[channel n, pixel p] = W[n, :] . A[:, p]
//     M = output channels (TMEM lanes; the weight block [Ncta x Kpad] is the K-major MMA "A" operand and stays resident in
//         shared memory for the CTA's lifetime), N = 64 or 128 pixels (TMEM columns; the activation tile is the MN-major,
//         i.e. pixel-contiguous, MMA "B" operand), K = input channels.
//     Pixels on the column axis mean that an epilogue thread owns one output channel and 16 CONSECUTIVE pixels per
//     tcgen05.ld, so results (and the residual) move with 8/16-byte vector accesses along NCHW rows.
//   persistent CTAs (one per SM), 17 warps with fixed roles:
//     warp 0       allocates TMEM; one elected thread issues tcgen05.mma (128 x Npx x 16 per instruction)
//     warps 1-8    epilogue: tcgen05.ld the fp32 accumulator, add the residual, store bf16
//     warps 9-16   producers: build the activation tile in shared memory in the UMMA canonical no-swizzle layout
//                  (generic-proxy stores -> fence.proxy.async -> mbarrier), ring of up to 6 stages of 16 KiB
//   accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// k_pw_wgrad (weight gradient):  dW[m, n] = sum_{i,p} G[i, m, p] * A(i, n, p), both operands K-major (pixels are the
//     reduction axis), 128-byte-swizzled canonical layout; split over pixel ranges, fp32 partial slices, fixed-order
//     reduction (deterministic).
#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int kEpiWarp0 = 1;
constexpr int kNumEpiWarps = 8;
constexpr int kProdWarp0 = kEpiWarp0 + kNumEpiWarps;  // 9
constexpr int kNumProdWarps = 8;                             // gather (3D shift) producers, weight-gradient kernel
constexpr int kThreads = (kProdWarp0 + kNumProdWarps) * 32;  // 544
constexpr int kRowProdWarps = 7;                             // plain / BN+ReLU producers of k_pw_conv: 16 warps = 512 threads,
constexpr int kRowThreads = (kProdWarp0 + kRowProdWarps) * 32;  // so that a thread may use up to 128 registers
constexpr int kStageBytes = 16384;  // largest activation stage: Npx pixels x (8192 / Npx) channels
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHdrBytes = 256;
constexpr int kStgBytes = kNumEpiWarps * 2048;  // epilogue staging: 32 channels x 64 B per warp

enum { PROD_PLAIN = 0, PROD_BNRELU = 1, PROD_SHIFT3D = 2 };

__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint2 ldg8(const void *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ uint32_t ldg4(const void *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ uint32_t ldg2(const void *p) { return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p)); }

// 8 consecutive pixels of one channel row -> 4 packed words; VEC = alignment granule in elements (HW % VEC == 0,
// p % VEC == 0); elements at or beyond `nvalid` read as zero
template <int VEC> __device__ __forceinline__ void load_unit(const __nv_bfloat16 *src, int nvalid, uint32_t (&r)[4]) {
    if (VEC == 8) {
        if (nvalid >= 8) {
            const uint4 v = ldg16(src);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
            r[0] = r[1] = r[2] = r[3] = 0u;
        }
    } else if (VEC == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (nvalid >= 4 * h + 4) {
                const uint2 v = ldg8(src + 4 * h);
                r[2 * h] = v.x; r[2 * h + 1] = v.y;
            } else {
                r[2 * h] = r[2 * h + 1] = 0u;
            }
        }
    } else if (VEC == 2) {
#pragma unroll
        for (int h = 0; h < 4; ++h) r[h] = (nvalid >= 2 * h + 2) ? ldg4(src + 2 * h) : 0u;
    } else {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t lo = (nvalid > 2 * h) ? ldg2(src + 2 * h) : 0u;
            const uint32_t hi = (nvalid > 2 * h + 1) ? ldg2(src + 2 * h + 1) : 0u;
            r[h] = lo | (hi << 16);
        }
    }
}

// A "unit" is 8 consecutive positions of the flattened (image, pixel) axis of one channel row of an NCHW tensor
// [NI, C, HW].  It starts at pixel p of some image (p % VEC == 0, HW % VEC == 0) and may run into the next image, whose
// same channel row lies (C-1)*HW elements further than the plain continuation.  piece_offsets gives, for each VEC-wide
// piece, its element offset relative to the unit's first element.
template <int VEC> __device__ __forceinline__ void piece_offsets(int p, int HW, int wrap, int (&o)[8 / VEC]) {
#pragma unroll
    for (int s_ = 0; s_ < 8 / VEC; ++s_) {
        int po = p + s_ * VEC, w = 0;
        while (po >= HW) { po -= HW; w += wrap; }
        o[s_] = s_ * VEC + w;
    }
}

// loads the unit at element offset `off` of `base`; pieces at or beyond `nleft` positions read as zero
template <int VEC>
__device__ __forceinline__ void load_unit_flat(const __nv_bfloat16 *base, int off, const int (&o)[8 / VEC], int nleft, uint32_t (&r)[4]) {
    if (VEC == 8) {
        if (nleft >= 8) {
            const uint4 v = ldg16(base + off);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
            r[0] = r[1] = r[2] = r[3] = 0u;
        }
    } else if (VEC == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (nleft >= 4 * h + 4) {
                const uint2 v = ldg8(base + off + o[h]);
                r[2 * h] = v.x; r[2 * h + 1] = v.y;
            } else {
                r[2 * h] = r[2 * h + 1] = 0u;
            }
        }
    } else if (VEC == 2) {
#pragma unroll
        for (int h = 0; h < 4; ++h) r[h] = (nleft >= 2 * h + 2) ? ldg4(base + off + o[h]) : 0u;
    } else {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t lo = (nleft > 2 * h) ? ldg2(base + off + o[2 * h]) : 0u;
            const uint32_t hi = (nleft > 2 * h + 1) ? ldg2(base + off + o[2 * h + 1]) : 0u;
            r[h] = lo | (hi << 16);
        }
    }
}

template <int VEC>
__device__ __forceinline__ void store_unit_flat(__nv_bfloat16 *base, int off, const int (&o)[8 / VEC], int nleft, const uint32_t (&v)[4]) {
    if (VEC == 8) {
        if (nleft >= 8) *reinterpret_cast<uint4 *>(base + off) = make_uint4(v[0], v[1], v[2], v[3]);
    } else if (VEC == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (nleft >= 4 * h + 4) *reinterpret_cast<uint2 *>(base + off + o[h]) = make_uint2(v[2 * h], v[2 * h + 1]);
    } else if (VEC == 2) {
#pragma unroll
        for (int h = 0; h < 4; ++h)
            if (nleft >= 2 * h + 2) *reinterpret_cast<uint32_t *>(base + off + o[h]) = v[h];
    } else {
        unsigned short *us = reinterpret_cast<unsigned short *>(base);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (nleft > 2 * h) us[off + o[2 * h]] = (unsigned short)(v[h] & 0xffffu);
            if (nleft > 2 * h + 1) us[off + o[2 * h + 1]] = (unsigned short)(v[h] >> 16);
        }
    }
}

// relu(x*sc + bi) on the first `nvalid` elements; the rest stay zero (padding must not become relu(bias))
__device__ __forceinline__ void bn_relu_unit(uint32_t (&r)[4], float sc, float bi, int nvalid) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float lo = (nvalid > 2 * h) ? fmaxf(fmaf(bf16_lo(r[h]), sc, bi), 0.f) : 0.f;
        const float hi = (nvalid > 2 * h + 1) ? fmaxf(fmaf(bf16_hi(r[h]), sc, bi), 0.f) : 0.f;
        r[h] = pack_bf16x2(lo, hi);
    }
}

// ---- 3D shift gather ---------------------------------------------------------------------------------------------
struct ShiftSrc {
    const __nv_bfloat16 *x;  // [clips, T, K, H, W]
    const void *shift;       // [3, K] rows (T, H, W)
    int shift_dt, T, H, W, HW, K;
    int last_word;           // index of the 32-bit word holding the tensor's last element (< 2^30 elements pairs)
};

// per (channel, frame) constants: floors and (1-r, r) weights of cuda_src/rubiks3d_kernels.cu:65-74, folded with the frame
struct ShiftCh {
    int fH, fW;
    int ebase[2];   // element index of (clip, t+fT+a, k, row 0, col 0); only used when tok[a]
    bool tok[2];    // source frame t+fT+a exists
    float wab[4];   // wT[a] * wH[b]
    float wW0, wW1;
};

__device__ __forceinline__ ShiftCh shift_channel(const ShiftSrc &s, int k, int clip, int t) {
    const float sT = ld_param<float>(s.shift, s.shift_dt, k), sH = ld_param<float>(s.shift, s.shift_dt, s.K + k),
                sW = ld_param<float>(s.shift, s.shift_dt, 2 * s.K + k);
    ShiftCh c;
    const int fT = floor3d(sT);
    c.fH = floor3d(sH); c.fW = floor3d(sW);
    const float rT = sT - fT, rH = sH - c.fH;
    c.wW1 = sW - c.fW; c.wW0 = 1.f - c.wW1;
    c.wab[0] = (1.f - rT) * (1.f - rH); c.wab[1] = (1.f - rT) * rH;
    c.wab[2] = rT * (1.f - rH); c.wab[3] = rT * rH;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int ts = t + fT + a;
        c.tok[a] = ts >= 0 && ts < s.T;
        c.ebase[a] = ((clip * s.T + ts) * s.K + k) * s.HW;
    }
    return c;
}

// RubiksShift3D forward (stride 1, pad 0; cuda_src/rubiks3d_kernels.cu:54-74,97-203) for the 8 output pixels
// (row h, columns c0..c0+7) of one (channel, frame):
//     out[c] = sum_{a,b,d in {0,1}} wT[a] wH[b] wW[d] * X[t+fT+a, k, h+fH+b, c+fW+d],   X = 0 outside the tensor.
// shift3d_load fetches the 9 source columns needed from each of the 4 (frame, row) pairs as 5 aligned 32-bit words;
// shift3d_compute realigns them with a funnel shift, combines the (frame, row) pairs first,
// V[j] = sum_ab wT[a] wH[b] X_ab[j], then the two column taps, out[i] = V[i] wW0 + V[i+1] wW1  (fp32 throughout; same
// taps and weights as the reference, summed in a different order).  Columns outside [0, W) are zeroed after the
// combination, rows / frames outside are never loaded.  The split lets callers keep the loads of the next run in
// flight while the current one is being combined.
struct RunLoad {
    uint32_t wd[4][5];
    uint32_t shr[4];  // 0 or 16 per (frame, row): parity of the first source element
};

__device__ __forceinline__ void shift3d_load(const ShiftSrc &s, const ShiftCh &ch, int h, int c0, RunLoad &L) {
    const int ws = c0 + ch.fW;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(s.x);
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        const int hs = h + ch.fH + (ab & 1);
        const bool rowok = ch.tok[ab >> 1] && (unsigned)hs < (unsigned)s.H;
        const int e0 = ch.ebase[ab >> 1] + hs * s.W + ws;  // may be slightly negative on the tensor's first row
        const int wi0 = e0 >> 1;                            // floor
        L.shr[ab] = (uint32_t)(e0 & 1) * 16u;
        if (wi0 >= 0 && wi0 + 4 <= s.last_word) {
            const uint32_t *wp = words + wi0;
#pragma unroll
            for (int i = 0; i < 5; ++i) L.wd[ab][i] = rowok ? __ldg(wp + i) : 0u;
        } else {
            // first / last row of the whole tensor: the words that fall outside hold only columns outside [0, W)
#pragma unroll
            for (int i = 0; i < 5; ++i) L.wd[ab][i] = rowok ? __ldg(words + min(max(wi0 + i, 0), s.last_word)) : 0u;
        }
    }
}

__device__ __forceinline__ void shift3d_compute(const ShiftSrc &s, const ShiftCh &ch, int c0, const RunLoad &L, float (&o)[8]) {
    const int ws = c0 + ch.fW;
    float V[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) V[j] = 0.f;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        const float w = ch.wab[ab];
        const uint32_t sh = L.shr[ab];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t nw = __funnelshift_r(L.wd[ab][i], L.wd[ab][i + 1], sh);
            V[2 * i] = fmaf(w, bf16_lo(nw), V[2 * i]);
            V[2 * i + 1] = fmaf(w, bf16_hi(nw), V[2 * i + 1]);
        }
        V[8] = fmaf(w, bf16_lo(L.wd[ab][4] >> sh), V[8]);
    }
#pragma unroll
    for (int j = 0; j < 9; ++j)
        if ((unsigned)(ws + j) >= (unsigned)s.W) V[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = V[i] * ch.wW0 + V[i + 1] * ch.wW1;
}

// run g of an image (rows of `rpr` runs of 8 columns) -> (row, first column)
__device__ __forceinline__ void run_coords(int g, int rpr, uint32_t magic, int &h, int &c0) {
    h = (int)(((uint32_t)g * magic) >> 16);
    c0 = (g - h * rpr) * 8;
}

// =====================================================================================================================
struct PwArgs {
    const __nv_bfloat16 *x;    // [NI, K, HW]
    const void *w;             // [N, K] (w_trans: [K, N]), bf16 or fp32 (w_dt)
    int w_dt, w_trans;
    int w_resident;            // RB_W_RESIDENT: the weight buffer may be read before griddepcontrol.wait
    const __nv_bfloat16 *res;  // [NI, N, HW] or null
    __nv_bfloat16 *out;        // [NI, N, HW]
    const float *a_sb;         // PROD_BNRELU: per input channel (scale, bias) pairs [K, 2]
    const void *shift;         // PROD_SHIFT3D: [3, K]
    int shift_dt;
    int T, H, W;               // PROD_SHIFT3D: frames per clip, map size (HW = H*W)
    int NI, K, N, HW;
    int Kpad, Ncta, Mt, Npx, kstage, acc_stages, stages, tmem_cols, stage_bytes, reserved0;
    int NP, total_tiles, k_stages;  // NP = NI*HW positions on the flattened (image, pixel) axis, tiled by Npx
    uint32_t off_w, off_a, off_sb, off_stg, w_lbo, a_lbo;
    int b_rows;                // activation operand staged as 128-byte-swizzled pixel rows (cp.async path) instead of 8x16-byte core matrices
    uint32_t rpr, rpr_magic;   // runs of 8 columns per image row; (r * rpr_magic) >> 16 == r / rpr for r < 4096
    double *stats;             // non-null: per-channel (sum, sum of squares) of `out` -> stats[(c * stats_splits + split) * 2 + {0,1}]
    int stats_splits;          // = 2 * gridDim.x (every CTA column and pixel half is one split)
    uint32_t hw_mul, hw_shr;   // exact division by HW for positions < 2^31: (umulhi(P, hw_mul) >> hw_shr), HW == 1: mul 0
#ifdef RB_DEBUG_TRACE
    unsigned long long *trace; // debug builds only: per-CTA event timestamps (globaltimer ns), 128 slots per CTA; null = off
#endif
};

#ifdef RB_DEBUG_TRACE
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif
// P / HW without the ~30-instruction runtime division (Granlund-Montgomery, exact for P < 2^31)
__device__ __forceinline__ int div_hw(const PwArgs &a, int P) {
    return a.hw_mul == 0u ? P : (int)(__umulhi((uint32_t)P, a.hw_mul) >> a.hw_shr);
}
// Pipeline time stamps exist only in debug builds (python -m rubiksnet_b200.build --trace -> -DRB_DEBUG_TRACE, used by
// tools/trace_pw.py); the product library contains no tracing code and no work-skipping switches.
#ifdef RB_DEBUG_TRACE
#define PW_TRACE_IF(cond, slot) do { if (a.trace) { if (cond) a.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 128 + (slot)] = gtime(); } } while (0)
#define PW_TRACE(slot) do { if (a.trace) a.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 128 + (slot)] = gtime(); } while (0)
#else
#define PW_TRACE_IF(cond, slot) do { } while (0)
#define PW_TRACE(slot) do { } while (0)
#endif

struct Hdr {
    uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(Hdr) <= kHdrBytes, "header");

// Resident weight block of a CTA: W[n0 + n, k] (n < nrows) -> shared memory at
//     (k/8)*w_lbo + (n/8)*128 + (n%8)*16 + (k%8)*2
// (K-major 8x16-byte core matrices; w_lbo is an odd multiple of 16 bytes so that consecutive k-groups fall into different
// bank groups).  Rows n >= nrows and columns k >= K are zero.  Threads follow the contiguous axis of the weight buffer
// ([N,K]: 8 consecutive k = one 16-byte shared store; transposed [K,N]: 8 consecutive n = eight 2-byte stores), keep 8
// units in flight each, and every CTA starts at a different offset so that the grid does not walk the same L2 lines in
// lockstep.
__device__ __forceinline__ void stage_weights(const PwArgs &a, unsigned char *smem_w, int n0, int nrows, int tid, int nthreads) {
    const int kgroups = a.Kpad >> 3, rows8 = (a.Ncta + 7) & ~7;
    const __nv_bfloat16 *wb = reinterpret_cast<const __nv_bfloat16 *>(a.w);
    const float *wf = reinterpret_cast<const float *>(a.w);
    const bool f32 = a.w_dt != RB_BF16;
    constexpr int UB = 8;
    if (!a.w_trans && !f32 && (a.K & 7) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0) {
        // packed bf16 [N, K] weights (rb_pw_weight_pack): every unit is one 16-byte cp.async straight into the operand
        // layout; all of a thread's copies are in flight together, so the whole block costs about one memory latency
        const int total = rows8 * kgroups;
        const int rot = (int)(((int64_t)blockIdx.x * total) / gridDim.x);
        const uint32_t sw = smem_u32(smem_w);
        for (int ul = tid; ul < total; ul += nthreads) {
            const int u = ul + rot < total ? ul + rot : ul + rot - total;
            const int n = u / kgroups, kg = u - n * kgroups;
            const bool ok = n < nrows && kg * 8 < a.K;
            cp_async16(sw + (uint32_t)kg * a.w_lbo + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u,
                       ok ? wb + (int64_t)(n0 + n) * a.K + kg * 8 : wb, ok ? 16u : 0u);
        }
        cp_async_commit();
        cp_async_wait<0>();
        return;
    }
    if (!a.w_trans) {
        const bool vec_ok = (a.K & 7) == 0;
        const int total = rows8 * kgroups;
        const int rot = (int)(((int64_t)blockIdx.x * total) / gridDim.x);
        for (int u0 = tid; u0 < total; u0 += UB * nthreads) {
            uint4 lo[UB], hi[UB];  // fp32: two float4 (as bits); bf16: lo only
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int ul = u0 + b * nthreads;
                const int u = ul + rot < total ? ul + rot : ul + rot - total;
                const int n = u / kgroups, kg = u - n * kgroups;
                lo[b] = hi[b] = make_uint4(0u, 0u, 0u, 0u);
                if (ul < total && n < nrows && kg * 8 < a.K) {
                    const int64_t e = (int64_t)(n0 + n) * a.K + kg * 8;
                    if (vec_ok) {
                        if (f32) { lo[b] = ldg16(wf + e); hi[b] = ldg16(wf + e + 4); }
                        else lo[b] = ldg16(wb + e);
                    } else {
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            f[i] = kg * 8 + i < a.K ? (f32 ? __ldg(wf + e + i) : __bfloat162float(wb[e + i])) : 0.f;
                        lo[b] = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
                        hi[b] = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int ul = u0 + b * nthreads;
                if (ul >= total) continue;
                const int u = ul + rot < total ? ul + rot : ul + rot - total;
                const int n = u / kgroups, kg = u - n * kgroups;
                uint4 o = lo[b];
                if (f32 || !vec_ok)
                    o = make_uint4(pack_bf16x2(__uint_as_float(lo[b].x), __uint_as_float(lo[b].y)),
                                   pack_bf16x2(__uint_as_float(lo[b].z), __uint_as_float(lo[b].w)),
                                   pack_bf16x2(__uint_as_float(hi[b].x), __uint_as_float(hi[b].y)),
                                   pack_bf16x2(__uint_as_float(hi[b].z), __uint_as_float(hi[b].w)));
                *reinterpret_cast<uint4 *>(smem_w + (size_t)kg * a.w_lbo + (size_t)(n >> 3) * 128 + (n & 7) * 16) = o;
            }
        }
    } else {
        // buffer is [K, N]: unit = (k, group of 8 consecutive rows n); consecutive threads take consecutive k
        const bool vec_ok = (a.N & 7) == 0;  // then n0 and the row groups are 8-aligned as well (Ncta % 8 == 0)
        const int ngroups = rows8 >> 3, total = ngroups * a.Kpad;
        const int rot = (int)(((int64_t)blockIdx.x * total) / gridDim.x);
        for (int u0 = tid; u0 < total; u0 += UB * nthreads) {
            uint4 lo[UB], hi[UB];
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int ul = u0 + b * nthreads;
                const int u = ul + rot < total ? ul + rot : ul + rot - total;
                const int ng = u / a.Kpad, k = u - ng * a.Kpad;
                lo[b] = hi[b] = make_uint4(0u, 0u, 0u, 0u);
                if (ul < total && k < a.K && ng * 8 < nrows) {
                    const int64_t e = (int64_t)k * a.N + n0 + ng * 8;
                    if (vec_ok) {  // rows beyond nrows inside the group belong to the next CTA or do not exist: masked below
                        if (f32) { lo[b] = ldg16(wf + e); hi[b] = ldg16(wf + e + 4); }
                        else lo[b] = ldg16(wb + e);
                    } else {
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            f[i] = ng * 8 + i < nrows ? (f32 ? __ldg(wf + e + i) : __bfloat162float(wb[e + i])) : 0.f;
                        lo[b] = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
                        hi[b] = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int ul = u0 + b * nthreads;
                if (ul >= total) continue;
                const int u = ul + rot < total ? ul + rot : ul + rot - total;
                const int ng = u / a.Kpad, k = u - ng * a.Kpad;
                uint32_t w[4];  // 8 bf16 values: rows ng*8 .. ng*8+7 at column k
                if (f32 || !vec_ok) {
                    w[0] = pack_bf16x2(__uint_as_float(lo[b].x), __uint_as_float(lo[b].y));
                    w[1] = pack_bf16x2(__uint_as_float(lo[b].z), __uint_as_float(lo[b].w));
                    w[2] = pack_bf16x2(__uint_as_float(hi[b].x), __uint_as_float(hi[b].y));
                    w[3] = pack_bf16x2(__uint_as_float(hi[b].z), __uint_as_float(hi[b].w));
                } else {
                    w[0] = lo[b].x; w[1] = lo[b].y; w[2] = lo[b].z; w[3] = lo[b].w;
                }
                unsigned short *d = reinterpret_cast<unsigned short *>(smem_w + (size_t)(k >> 3) * a.w_lbo + (size_t)ng * 128 + (k & 7) * 2);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t v = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                    d[i * 8] = (ng * 8 + i < nrows) ? (unsigned short)v : (unsigned short)0;
                }
            }
        }
    }
}

// MMA issue loop of k_pw_conv, run by one thread.  Per stage: wait for the producers, issue ksteps x MT instructions
// (128 channels x Npx pixels x 16 input channels each), commit to the stage's `empty` barrier.
template <int MT>
__device__ __forceinline__ void mma_loop(const PwArgs &a, Hdr *hdr, unsigned char *smem_w, unsigned char *smem_a,
                                         uint32_t tmem_base, int tile0, int tstride) {
    const uint32_t idesc = instr_desc_bf16(128, a.Npx, /*weights K-major*/ 0, /*activations MN-major*/ 1);
    const uint64_t adesc0 = smem_desc(smem_u32(smem_w), a.w_lbo, 128, LAYOUT_NONE);
    // activations, MN-major.  Row layout: per 8-channel group 8 rows of 64 pixels (128 bytes, 16-byte chunks XOR-swizzled
    // by the row), groups 1 KiB apart (SBO), the second 64-pixel half of a 128-pixel tile (kstage/8) KiB further (LBO)
    const uint64_t bdesc0 = a.b_rows ? smem_desc(smem_u32(smem_a), (uint32_t)(a.kstage >> 3) * 1024u, 1024, LAYOUT_SW128)
                                     : smem_desc(smem_u32(smem_a), a.a_lbo, 128, LAYOUT_NONE);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
    // descriptors advance by adding (bytes >> 4) to the start-address field (all addresses stay below 256 KiB)
    const uint32_t a_kstep = (2 * a.w_lbo) >> 4, b_kstep = a.b_rows ? (2048u >> 4) : (2 * a.a_lbo) >> 4;
    const int ksteps_per_stage = a.kstage >> 4;
    const int acc_cols = a.Mt * a.Npx;
    const uint32_t npx = (uint32_t)a.Npx, stage16 = (uint32_t)(a.stage_bytes >> 4);
    int slot = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
        const int as = it % a.acc_stages;
        const uint32_t aph = (uint32_t)(it / a.acc_stages) & 1u;
        const uint32_t tacc = tmem_base + as * acc_cols;
        mbar_wait(&hdr->tmem_empty[as], aph ^ 1u);
        tc_fence_after();
        PW_TRACE_IF(it < 7, 8 + it * 8);
        uint32_t a_lo = a_lo0;  // weights: first k-group of the current stage
        uint32_t acc = 0u;
        int kleft = a.Kpad >> 4;
        for (int st = 0; st < a.k_stages; ++st) {
            mbar_wait(&hdr->full[slot], phase);
            tc_fence_after();
            PW_TRACE_IF(it == 2 && st < 8, 64 + st * 8);
            PW_TRACE_IF(it < 7 && st == 0, 9 + it * 8);
            PW_TRACE_IF(it < 7 && st == a.k_stages - 1, 10 + it * 8);
            const int ksteps = min(ksteps_per_stage, kleft);
            kleft -= ksteps;
            uint32_t b_lo = b_lo0 + (uint32_t)slot * stage16;
            for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    mma_bf16_lohi(tacc + mt * npx, a_lo + mt * (2048 >> 4), a_hi, b_lo, b_hi, idesc, acc);
                acc = 1u;
                a_lo += a_kstep;
                b_lo += b_kstep;
            }
            mma_commit(&hdr->empty[slot]);
            PW_TRACE_IF(it == 2 && st < 8, 65 + st * 8);
            if (++slot == a.stages) { slot = 0; phase ^= 1u; }
        }
        mma_commit(&hdr->tmem_full[as]);
        PW_TRACE_IF(it < 7, 4 + it * 8);
    }
}

template <int PROD, int VEC>
__global__ void __launch_bounds__(PROD == PROD_SHIFT3D ? kThreads : kRowThreads, 1) k_pw_conv(const PwArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Hdr *hdr = reinterpret_cast<Hdr *>(smem);
    unsigned char *smem_w = smem + a.off_w;
    unsigned char *smem_a = smem + a.off_a;
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * a.Ncta;
    const int nrows = min(a.Ncta, a.N - n0);  // output channels owned by this CTA

    // ---- one-time setup: barriers, TMEM, resident weight block --------------------------------------------------
    if (tid == 0) {
        PW_TRACE(0);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&hdr->full[i], a.b_rows ? 1 : kNumProdWarps);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hdr->tmem_full[i], 1);
            mbar_init(&hdr->tmem_empty[i], kNumEpiWarps * 32);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, (uint32_t)a.tmem_cols);
    }
    // Everything up to here touches no global memory.  A weight buffer flagged RB_W_RESIDENT was written at least two
    // launches ago (common.cuh: programmatic dependent launch), so the weight block is staged while the previous kernel of
    // the stream is still draining; activations, residual and BN coefficients come from that kernel and wait for it.
    if (a.w_resident && warp < kProdWarp0) stage_weights(a, smem_w, n0, nrows, tid, kProdWarp0 * 32);
    pdl_sync();
    if (PROD == PROD_BNRELU)
        for (int k = tid; k < a.Kpad; k += (int)blockDim.x) {
            smem_sb[k] = k < a.K ? a.a_sb[2 * k] : 0.f;
            smem_sb[a.Kpad + k] = k < a.K ? a.a_sb[2 * k + 1] : 0.f;
        }
    // barriers, TMEM address and BN coefficients are visible to everybody after this sync; the producer warps then start
    // loading activations at once, while the MMA + epilogue warps (the only readers of the weight block, through the
    // tensor core) stage the weights and meet again on a named barrier of their own
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    if (warp < kProdWarp0) {
        if (!a.w_resident) stage_weights(a, smem_w, n0, nrows, tid, kProdWarp0 * 32);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kProdWarp0 * 32) : "memory");
    }
    if (tid == 0) PW_TRACE(2);

    if (tid == kProdWarp0 * 32) PW_TRACE(1);

    const int tile0 = blockIdx.x, tstride = gridDim.x;
    const int acc_cols = a.Mt * a.Npx;

    if (warp == 0) {
        // ===================================== MMA issuer ========================================================
        // ONE elected thread runs the whole loop (barrier waits included): no divergence inside it, descriptors stepped
        // with 32-bit adds, the M tiles of a K step unrolled at compile time.
        if (elect_one()) {
            switch (a.Mt) {
                case 1: mma_loop<1>(a, hdr, smem_w, smem_a, tmem_base, tile0, tstride); break;
                case 2: mma_loop<2>(a, hdr, smem_w, smem_a, tmem_base, tile0, tstride); break;
                case 3: mma_loop<3>(a, hdr, smem_w, smem_a, tmem_base, tile0, tstride); break;
                default: mma_loop<4>(a, hdr, smem_w, smem_a, tmem_base, tile0, tstride); break;
            }
        }
        __syncwarp();
    } else if (warp < kProdWarp0) {
        // ===================================== epilogue ==========================================================
        // warp -> TMEM lane quarter (hardware: warp id % 4) and one half of the tile's pixel columns
        const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
        const int cbeg = half * (a.Npx >> 1), cend = cbeg + (a.Npx >> 1);
        unsigned char *stg = smem + a.off_stg + (warp - kEpiWarp0) * 2048;
        // BatchNorm statistics of the output (a.stats): every lane owns the same 4 channel rows of each M tile for the
        // whole kernel, so the sums stay in registers until the end
        float st_s[4][4], st_q[4][4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int i = 0; i < 4; ++i) st_s[m][i] = st_q[m][i] = 0.f;
        int it = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            const int as = it % a.acc_stages;
            const uint32_t aph = (uint32_t)(it / a.acc_stages) & 1u;
            const int P0 = tile * a.Npx;  // first position of the tile on the flattened (image, pixel) axis
            mbar_wait(&hdr->tmem_full[as], aph);
            tc_fence_after();
            PW_TRACE_IF(tid == kEpiWarp0 * 32 && it < 7, 5 + it * 8);
            // rounds of (32 channels = this warp's TMEM lanes of one M tile) x (32 pixels): TMEM -> registers (thread =
            // channel) -> bf16 -> per-warp staging tile in shared memory (32 channels x 64 B, 16-byte chunks
            // XOR-swizzled) -> read back with 4 lanes per channel row, so every global access instruction touches 8 rows
            // x 64 contiguous bytes instead of 32 rows x 16 bytes.  The TMEM load of round r+1 is in flight while round r
            // goes through shared memory, and the accumulator stage is released as soon as the last load has landed.
            const int ncr = (cend - cbeg) >> 5;                                  // column rounds per M tile
            const int mts = nrows > q * 32 ? (nrows - q * 32 + 127) >> 7 : 0;    // M tiles holding channels of this warp
            int nr = mts * ncr;
            // One loop body for every round (the three roles of the CTA share the instruction caches: a second inlined
            // copy of this body, tried for a software-pipelined TMEM load, cost more than the overlap gained).
            for (int ridx = 0; ridx < nr; ++ridx) {
                const int mt = ncr == 2 ? ridx >> 1 : (ncr == 1 ? ridx : ridx / ncr), c0 = cbeg + (ridx - mt * ncr) * 32;
                const int r0 = mt * 128 + q * 32;  // first CTA-local channel of this warp's 32 TMEM lanes
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols + mt * a.Npx + c0;
                uint32_t v[2][16];
                __syncwarp();
                tmem_ld16(taddr, v[0]);
                tmem_ld16(taddr + 16, v[1]);
                // this lane's four (channel row, 8 positions) pieces of the round: addresses first, and the residual
                // loads go out while the accumulator is being fetched
                int off[4], nv[4], po[8 / VEC];
                uint32_t rv[4][4];
                {
                    const int P = P0 + c0 + (lane & 3) * 8;
                    const int img = div_hw(a, P), pp = P - img * a.HW;
                    piece_offsets<VEC>(pp, a.HW, (a.N - 1) * a.HW, po);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = (lane >> 2) + 8 * i;
                        nv[i] = (r0 + row < nrows) ? a.NP - P : 0;  // <= 0 beyond the tensor: nothing moves
                        off[i] = ((img * a.N + n0 + r0 + row) * a.HW) + pp;
                        if (a.res != nullptr) load_unit_flat<VEC>(a.res, off[i], po, nv[i], rv[i]);
                    }
                }
                tmem_ld_wait();
                if (ridx == nr - 1) {  // every TMEM load of this tile has landed: hand the accumulator stage back
                    tc_fence_before();
                    mbar_arrive(&hdr->tmem_empty[as]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t *vv = &v[c >> 1][(c & 1) * 8];
                    *reinterpret_cast<uint4 *>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                        make_uint4(pack_bf16x2(__uint_as_float(vv[0]), __uint_as_float(vv[1])),
                                   pack_bf16x2(__uint_as_float(vv[2]), __uint_as_float(vv[3])),
                                   pack_bf16x2(__uint_as_float(vv[4]), __uint_as_float(vv[5])),
                                   pack_bf16x2(__uint_as_float(vv[6]), __uint_as_float(vv[7])));
                }
                __syncwarp();
                float rs1[4] = {0.f, 0.f, 0.f, 0.f}, rs2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = (lane >> 2) + 8 * i, ch = lane & 3;
                    const uint4 sv = *reinterpret_cast<const uint4 *>(stg + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
                    uint32_t ov[4] = {sv.x, sv.y, sv.z, sv.w};
                    if (a.res != nullptr) {
                        // `out += shortcut` on the bf16 conv result (the rounding order of conv3 followed by the add)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            ov[e] = pack_bf16x2(bf16_lo(ov[e]) + bf16_lo(rv[i][e]), bf16_hi(ov[e]) + bf16_hi(rv[i][e]));
                    }
                    store_unit_flat<VEC>(a.out, off[i], po, nv[i], ov);
                    if (a.stats != nullptr) {  // of the stored (bf16-rounded) values, as BatchNorm would read them back
                        float s1 = 0.f, s2 = 0.f;
                        if (nv[i] >= 8) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = bf16_lo(ov[e]), hi = bf16_hi(ov[e]);
                                s1 += lo + hi;
                                s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = 2 * e < nv[i] ? bf16_lo(ov[e]) : 0.f, hi = 2 * e + 1 < nv[i] ? bf16_hi(ov[e]) : 0.f;
                                s1 += lo + hi;
                                s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                            }
                        }
                        rs1[i] = s1;
                        rs2[i] = s2;
                    }
                }
                if (a.stats != nullptr) {  // warp-uniform choice of the M tile's accumulators
                    if (mt == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { st_s[0][i] += rs1[i]; st_q[0][i] += rs2[i]; }
                    } else if (mt == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { st_s[1][i] += rs1[i]; st_q[1][i] += rs2[i]; }
                    } else if (mt == 2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { st_s[2][i] += rs1[i]; st_q[2][i] += rs2[i]; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { st_s[3][i] += rs1[i]; st_q[3][i] += rs2[i]; }
                    }
                }
            }
            if (nr == 0) {
                tc_fence_before();
                mbar_arrive(&hdr->tmem_empty[as]);
            }
            PW_TRACE_IF(tid == kEpiWarp0 * 32 && it < 7, 6 + it * 8);
        }
        if (a.stats != nullptr) {
            const int sp = blockIdx.x * 2 + half;
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s1 = st_s[m][i], s2 = st_q[m][i];  // the 4 lanes of a row hold its 4 groups of 8 pixels
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    const int chl = m * 128 + q * 32 + (lane >> 2) + 8 * i;
                    if ((lane & 3) == 0 && m < a.Mt && chl < nrows) {
                        double *o = a.stats + ((int64_t)(n0 + chl) * a.stats_splits + sp) * 2;
                        o[0] = (double)s1;
                        o[1] = (double)s2;
                    }
                }
        }
    } else {
        // ===================================== activation producers ==============================================
        const int pw = warp - kProdWarp0;
        const int kq = lane & 7, mgq = lane >> 3;  // gather producer: channel / group-of-runs lanes
        const ShiftSrc ssrc{a.x, a.shift, a.shift_dt, a.T, a.H, a.W, a.HW, a.K,
                            (int)(((int64_t)a.NI * a.K * a.HW - 1) >> 1)};

        int tile = tile0, st = 0, slot = 0;
        uint32_t phase = 0;
        if constexpr (PROD != PROD_SHIFT3D) {
            // One warp per stage: producer warp w owns stages w, w+8, ... and moves a whole 8 KiB stage through its
            // registers (64 per lane): all loads of the stage are issued back to back, then -- once the MMA warp has
            // released the slot -- stored into the operand layout.  A warp's loads share one scoreboard, so a warp can
            // only wait for ALL of its outstanding loads; giving every warp exactly one stage in flight makes that the
            // right wait, and the eight warps keep eight stages (64 KiB per SM) in flight regardless of the depth of the
            // shared-memory ring.  BN+ReLU is applied to the registers.
            //
            // A stage holds kstage channel rows per 64-pixel half: rows of 128 bytes, 16-byte chunks XOR-swizzled by
            // the row (MN-major SWIZZLE_128B operand).  A piece is PB = 2*VEC bytes of one row; the lanes of a warp
            // cover whole rows (RPI rows per instruction), so global reads are contiguous 128-byte segments and the
            // shared-memory stores are conflict free.
            // VEC == 1 (odd plane sizes, 7x7): pieces of two pixels, fetched with two 2-byte loads that may lie in
            // different images
            constexpr int EPP = VEC == 1 ? 2 : VEC;  // elements per piece
            // (their halves are merged only after every load has been issued -- a dependent OR between the loads would
            // serialise them -- so VEC == 1 stages are 4 KiB to keep both halves of all pieces in registers)
            constexpr int STAGE = VEC == 1 ? 4096 : 8192;
            constexpr int PB = 2 * EPP, PPR = 128 / PB, RPI = 32 / PPR, NPIECE = STAGE / (32 * PB), WPP = PB / 4;
            const int pc = lane % PPR, rl = lane / PPR;  // piece column, row inside one instruction
            const uint32_t smem_a_u32 = smem_u32(smem_a);
            const uint32_t chunk = (uint32_t)(pc * PB) >> 4, sub = (uint32_t)(pc * PB) & 15u;
            const uint32_t sb_u32 = smem_u32(smem_sb);
            // stage cursor of this warp
            // Only min(7, stages) warps take part: a warp waiting for the slot of stage m has stored stage m - W, so
            // m < c + stages + W (c = stage the MMA warp is consuming); W <= stages keeps every waiter less than two
            // ring rounds ahead of the consumer, which is what the one-bit mbarrier phase parity can tell apart.
            const int W = a.stages < kRowProdWarps ? a.stages : kRowProdWarps;
            int n = pw;  // global stage index
            st = pw % a.k_stages;
            tile = pw < W ? tile0 + (pw / a.k_stages) * tstride : a.total_tiles;
            while (tile < a.total_tiles) {
                slot = n % a.stages;
                phase = (uint32_t)(n / a.stages) & 1u;
                // pixel column of this lane in each 64-pixel half of the tile
                int offb[2], offb1[2];  // offb1 / pvalid1: second pixel of the piece (VEC == 1 only)
                bool pvalid[2], pvalid1[2];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int P = tile * a.Npx + hf * 64 + pc * EPP;
                    pvalid[hf] = hf * 64 < a.Npx && P < a.NP;
                    const int img = div_hw(a, P), pp = P - img * a.HW;
                    offb[hf] = img * a.K * a.HW + pp;
                    pvalid1[hf] = hf * 64 < a.Npx && P + 1 < a.NP;
                    offb1[hf] = pp + 1 < a.HW ? offb[hf] + 1 : (img + 1) * a.K * a.HW;
                }
                const int kbase = st * a.kstage;
                uint32_t r[NPIECE][WPP];
                uint32_t rhi[VEC == 1 ? NPIECE : 1];  // VEC == 1: second pixel of every piece
#pragma unroll
                for (int i = 0; i < NPIECE; ++i) {
                    const int R = i * RPI + rl;
                    const int hf = R >= a.kstage ? 1 : 0;
                    const int k = kbase + R - hf * a.kstage;
                    const bool ok = pvalid[hf] && k < a.K;
                    // uniform 64-bit base + 32-bit byte offset: one register per address (tensors stay below 4 GiB)
                    const __nv_bfloat16 *src = reinterpret_cast<const __nv_bfloat16 *>(
                        reinterpret_cast<const char *>(a.x) + (uint32_t)(offb[hf] + k * a.HW) * 2u);
                    if (VEC == 8) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (ok) v = ldg16(src);
                        r[i][0] = v.x; r[i][WPP > 1 ? 1 : 0] = v.y; r[i][WPP > 2 ? 2 : 0] = v.z; r[i][WPP > 3 ? 3 : 0] = v.w;
                    } else if (VEC == 4) {
                        uint2 v = make_uint2(0u, 0u);
                        if (ok) v = ldg8(src);
                        r[i][0] = v.x; r[i][WPP > 1 ? 1 : 0] = v.y;
                    } else if (VEC == 2) {
                        r[i][0] = ok ? ldg4(src) : 0u;
                    } else {
                        r[i][0] = ok ? ldg2(src) : 0u;
                        rhi[VEC == 1 ? i : 0] = (pvalid1[hf] && k < a.K)
                            ? ldg2(reinterpret_cast<const char *>(a.x) + (uint32_t)(offb1[hf] + k * a.HW) * 2u) : 0u;
                    }
                }
                if (VEC == 1) {
#pragma unroll
                    for (int i = 0; i < NPIECE; ++i) r[i][0] |= rhi[VEC == 1 ? i : 0] << 16;
                }
                if (PROD == PROD_BNRELU) {
                    // rows of a half are valid up to a per-lane limit (k < K), so validity is a compare against the
                    // unrolled index; the coefficient loads are volatile so that they are not hoisted into 2 x NPIECE
                    // live registers
                    const int half_i = a.kstage / RPI;  // pieces per half
                    const int lim0 = pvalid[0] ? (a.K - kbase - rl + RPI - 1) / RPI : 0;
                    const int lim1 = pvalid[1] ? half_i + (a.K - kbase - rl + RPI - 1) / RPI : 0;
                    uint32_t sbp = sb_u32 + (uint32_t)(kbase + rl) * 4u;
#pragma unroll
                    for (int i = 0; i < NPIECE; ++i) {
                        const bool second = i >= half_i;
                        const bool valid = second ? i < lim1 : i < lim0;
                        if (valid) {
                            const uint32_t ap = sbp + (uint32_t)((second ? i - half_i : i) * RPI) * 4u;
                            float sc, bi;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc) : "r"(ap));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bi) : "r"(ap + (uint32_t)a.Kpad * 4u));
#pragma unroll
                            for (int e = 0; e < WPP; ++e)
                                r[i][e] = pack_bf16x2(fmaxf(fmaf(bf16_lo(r[i][e]), sc, bi), 0.f),
                                                      (VEC > 1 || pvalid1[second ? 1 : 0]) ? fmaxf(fmaf(bf16_hi(r[i][e]), sc, bi), 0.f) : 0.f);
                        }
                    }
                }
                mbar_wait(&hdr->empty[slot], phase ^ 1u);
                const uint32_t sp = smem_a_u32 + (uint32_t)(slot * a.stage_bytes);
#pragma unroll
                for (int i = 0; i < NPIECE; ++i) {
                    const int R = i * RPI + rl;
                    const uint32_t d = sp + (uint32_t)((R >> 3) * 1024 + (R & 7) * 128) + ((chunk ^ (uint32_t)(R & 7)) << 4) + sub;
                    if (VEC == 8)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(r[i][0]), "r"(r[i][WPP > 1 ? 1 : 0]),
                                     "r"(r[i][WPP > 2 ? 2 : 0]), "r"(r[i][WPP > 3 ? 3 : 0]) : "memory");
                    else if (VEC == 4)
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d), "r"(r[i][0]), "r"(r[i][WPP > 1 ? 1 : 0]) : "memory");
                    else
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(d), "r"(r[i][0]) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr->full[slot]);
                PW_TRACE_IF(lane == 0 && st == a.k_stages - 1 && (tile - tile0) / tstride < 7, 7 + ((tile - tile0) / tstride) * 8);
                n += W;
                st += W;
                while (st >= a.k_stages) { st -= a.k_stages; tile += tstride; }
            }
        } else {
            // 3D-shift gather: work items = (channel, row run of <= 8 pixels); results go to shared memory with 2-byte
            // stores because a run need not be aligned to the 8-pixel groups of the operand layout
            const int rpr = (int)a.rpr;  // runs of 8 columns per image row
            for (; tile < a.total_tiles; tile += tstride) {
                const int Plo = tile * a.Npx, Phi = min(Plo + a.Npx, a.NP);  // flattened (image, pixel) range of the tile
                const int img_lo = Plo / a.HW, img_hi = (Phi - 1) / a.HW;
                for (st = 0; st < a.k_stages; ++st) {
                    mbar_wait(&hdr->empty[slot], phase ^ 1u);
                    unsigned char *sp = smem_a + (size_t)slot * a.stage_bytes;
                    for (int kl = pw * 8 + kq; kl < a.kstage; kl += 64) {
                        const int k = st * a.kstage + kl;
                        if (k >= a.Kpad) break;
                        unsigned char *srow = sp + (size_t)(kl >> 3) * a.a_lbo + (kl & 7) * 16;
                        for (int img = img_lo; img <= img_hi; ++img) {
                            // the part of this image inside the tile: pixels [p0, pend), tile-relative base mb
                            const int ib = img * a.HW, p0 = max(Plo - ib, 0), pend = min(Phi - ib, a.HW), mb = ib - Plo;
                            const int clip = img / a.T, t = img - clip * a.T;
                            const int h0 = p0 / a.W, h1 = (pend - 1) / a.W;
                            const int g0 = h0 * rpr + ((p0 - h0 * a.W) >> 3), g1 = h1 * rpr + ((pend - 1 - h1 * a.W) >> 3);
                            // results of a run (pixels pr0 .. pr0+7 of the image, clipped to [p0, pend) and to the row)
                            auto store_run = [&](int h, int c0, const float (&o)[8]) {
                                const int pr0 = h * a.W + c0, m0 = mb + pr0;
                                const int lo = max(p0 - pr0, 0), hi = min(min(8, a.W - c0), pend - pr0);
                                if (((m0 & 7) | lo) == 0 && hi == 8) {  // whole, group-aligned run: one 16-byte store
                                    *reinterpret_cast<uint4 *>(srow + (m0 >> 3) * 128) =
                                        make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                                   pack_bf16x2(o[6], o[7]));
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if ((unsigned)(i - lo) < (unsigned)(hi - lo)) {
                                            const int m = m0 + i;
                                            *reinterpret_cast<unsigned short *>(srow + (m >> 3) * 128 + (m & 7) * 2) =
                                                __bfloat16_as_ushort(__float2bfloat16_rn(o[i]));
                                        }
                                }
                            };
                            if (k >= a.K) {  // channel padding of the last K step: zeros
                                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                for (int g = g0 + mgq; g <= g1; g += 4) {
                                    int h, c0;
                                    run_coords(g, rpr, a.rpr_magic, h, c0);
                                    store_run(h, c0, z);
                                }
                                continue;
                            }
                            const ShiftCh ch = shift_channel(ssrc, k, clip, t);
                            // two runs in flight: the loads of run g+4 are issued before run g is combined
                            RunLoad LA, LB;
                            int g = g0 + mgq, hA = 0, cA = 0, hB = 0, cB = 0;
                            if (g <= g1) {
                                run_coords(g, rpr, a.rpr_magic, hA, cA);
                                shift3d_load(ssrc, ch, hA, cA, LA);
                            }
                            while (g <= g1) {
                                float o[8];
                                const bool hasB = g + 4 <= g1;
                                if (hasB) {
                                    run_coords(g + 4, rpr, a.rpr_magic, hB, cB);
                                    shift3d_load(ssrc, ch, hB, cB, LB);
                                }
                                shift3d_compute(ssrc, ch, cA, LA, o);
                                store_run(hA, cA, o);
                                if (!hasB) break;
                                if (g + 8 <= g1) {
                                    run_coords(g + 8, rpr, a.rpr_magic, hA, cA);
                                    shift3d_load(ssrc, ch, hA, cA, LA);
                                }
                                shift3d_compute(ssrc, ch, cB, LB, o);
                                store_run(hB, cB, o);
                                g += 8;
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&hdr->full[slot]);
                    if (++slot == a.stages) { slot = 0; phase ^= 1u; }
                }
            }
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) PW_TRACE(3);
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// tiling override (rb_pw_conv_set_tuning): lower bound on the number of output-channel splits; 0 = automatic
int g_min_n_splits = 0;

// splits the output channels over grid.y so that the resident weight block fits shared memory, picks the pixel-tile width
// so that two accumulator stages fit the 512 TMEM columns
bool plan(PwArgs &a, int prod, int vec, dim3 *grid, size_t *smem_bytes) {
    a.b_rows = prod != PROD_SHIFT3D ? 1 : 0;
    a.Kpad = round_up(a.K, 16);
    const int sb_bytes = prod == PROD_BNRELU ? round_up(2 * a.Kpad * 4, 128) : 0;
    int gy = 0;
    for (int cand = g_min_n_splits > 0 ? g_min_n_splits : 1; cand <= 16 && !gy; ++cand) {
        const int nc = round_up(cdiv(a.N, cand), 8);
        const int mt = cdiv(nc, 128);
        if (mt > 4) continue;
        const int w_lbo = nc * 16 + 16;
        const int w_bytes = round_up(kHdrBytes + sb_bytes + kStgBytes + (a.Kpad >> 3) * w_lbo, 1024) - (kHdrBytes + sb_bytes + kStgBytes);  // stages start 1 KiB aligned
        const int room = kSmemLimit - kHdrBytes - sb_bytes - kStgBytes - w_bytes;
        if (room / kStageBytes < 2) continue;
        // little room next to a large weight block: halve the stages so that more of them are in flight
        const int stage_bytes = a.b_rows ? (vec == 1 ? 4096 : 8192) : kStageBytes;  // row path: one stage = one warp's registers
        const int st = room / stage_bytes;
        // the last M tile reads (garbage, ignored) rows up to mt*128 of the last k-group: keep that inside the allocation
        const int reach = ((a.Kpad >> 3) - 1) * w_lbo + mt * 2048;
        const int stages = st < kMaxStages ? st : kMaxStages;
        if (reach > w_bytes + stages * stage_bytes) continue;
        gy = cand;
        a.Ncta = nc; a.Mt = mt; a.w_lbo = (uint32_t)w_lbo; a.stages = stages; a.stage_bytes = stage_bytes;
        a.off_sb = kHdrBytes;
        a.off_stg = kHdrBytes + sb_bytes;
        a.off_w = a.off_stg + kStgBytes;
        a.off_a = a.off_w + w_bytes;
    }
    if (!gy) return false;
    a.Npx = (2 * a.Mt * 128 <= 512) ? 128 : 64;
    a.acc_stages = 2;
    a.kstage = a.stage_bytes / 2 / a.Npx;
    a.a_lbo = (uint32_t)a.Npx * 16;
    a.k_stages = cdiv(a.Kpad, a.kstage);
    int cols = 32;
    while (cols < a.acc_stages * a.Mt * a.Npx) cols <<= 1;
    a.tmem_cols = cols;
    a.NP = a.NI * a.HW;
    a.total_tiles = cdiv(a.NP, a.Npx);
    if (a.HW <= 1) {
        a.hw_mul = 0u; a.hw_shr = 0u;
    } else {
        uint32_t l = 0;
        while ((1u << l) < (uint32_t)a.HW) ++l;  // ceil(log2 HW)
        a.hw_mul = (uint32_t)(((uint64_t(1) << (31 + l)) + (uint32_t)a.HW - 1) / (uint32_t)a.HW);
        a.hw_shr = l - 1;
    }
    a.rpr = (uint32_t)(a.W > 0 ? (a.W + 7) / 8 : 1);
    a.rpr_magic = 65536u / a.rpr + 1u;
    for (uint32_t r = 0; r < 4096; ++r)
        if (((r * a.rpr_magic) >> 16) != r / a.rpr) return false;
    *smem_bytes = (size_t)a.off_a + (size_t)a.stages * a.stage_bytes;
    const int gy_real = cdiv(a.N, a.Ncta);
    int ctas_x = sm_count() / gy_real;
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > a.total_tiles) ctas_x = a.total_tiles;
    *grid = dim3((unsigned)ctas_x, (unsigned)gy_real, 1);
    return true;
}

template <int PROD, int VEC> int launch(const PwArgs &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_pw_conv<PROD, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_pw_conv): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    launch_kernel(k_pw_conv<PROD, VEC>, dim3(grid), dim3(PROD == PROD_SHIFT3D ? kThreads : kRowThreads), smem_bytes, s, a);
    return launched("k_pw_conv");
}

// widest alignment granule (elements) shared by the tensors' base addresses and the plane size
int pick_vec(const PwArgs &a) {
    uintptr_t align = reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.res);
    const bool base16 = (align & 15) == 0;
    if (base16 && a.HW % 8 == 0) return 8;
    if (base16 && a.HW % 4 == 0) return 4;
    if (base16 && a.HW % 2 == 0) return 2;
    return 1;
}

template <int PROD> int launch_vec(const PwArgs &a, int vec, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    if (vec == 8) return launch<PROD, 8>(a, grid, smem_bytes, s);
    if (vec == 4) return launch<PROD, 4>(a, grid, smem_bytes, s);
    if (vec == 2) return launch<PROD, 2>(a, grid, smem_bytes, s);
    return launch<PROD, 1>(a, grid, smem_bytes, s);
}

// =====================================================================================================================
// Weight gradient of the 1x1 convolution:  dW[m, n] = sum_{i, p} G[i, m, p] * A(i, n, p)      (A as in the forward)
// Both operands are K-major here (the reduction runs over pixels, which are contiguous in NCHW), staged in the
// 128-byte-swizzled canonical layout: a row of 64 pixels = 128 bytes, 16-byte chunk c of row r stored at chunk
// c ^ (r % 8), so that 8 lanes reading one 128-byte global row segment write 8 different bank groups.
// grid = (pixel splits, N blocks, M blocks); every CTA reduces its pixel range into TMEM and writes one fp32
// partial [M, N] slice; k_wg_reduce sums the slices in a fixed order (deterministic, unlike atomics).
constexpr int kWgMaxStages = 4;
constexpr int kWgChunk = 64;             // pixels per stage
constexpr int kWgTileBytes = 128 * 128;  // one 128-row operand tile
constexpr int kWgProdWarps = 14;  // producer warps (the first 8 also drain the accumulators at the end); with the MMA warp
constexpr int kWgProdThreads = kWgProdWarps * 32;  // 480 threads, so that a thread may hold 16 units (64 registers) of a stage
constexpr int kWgThreads = kWgProdThreads + 32;
constexpr int kWgGroups = 2;      // plain / BN+ReLU producers: two groups of 7 warps fill alternate stages
constexpr int kWgUB = 16;         // units a thread keeps in flight

struct WgArgs {
    const __nv_bfloat16 *g;  // [NI, M, HW]
    const __nv_bfloat16 *x;  // [NI, N, HW]
    float *partial;          // [splits, M, N]
    const float *x_sb;       // PROD_BNRELU: (scale, bias) pairs [N, 2]
    const void *shift;
    int shift_dt, T, H, W;
    int NI, M, N, HW;
    int Mb, Mt, Nc, n_sub, sub_n, stages, tmem_cols;
    int cpi, total_chunks, chunks_per_split;
    uint32_t off_sb, off_stage, stage_bytes;
    uint32_t rpr, rpr_magic;
};

struct WgHdr {
    uint64_t full[kWgMaxStages], empty[kWgMaxStages], tmem_full;
    uint32_t tmem_base;
};
static_assert(sizeof(WgHdr) <= kHdrBytes, "header");

// byte offset of (row r, 16-byte chunk c) inside a 128-byte-swizzled K-major operand block (8-row atoms of 1 KiB)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

template <int PROD, int VEC>
__global__ void __launch_bounds__(kWgThreads, 1) k_pw_wgrad(const WgArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    WgHdr *hdr = reinterpret_cast<WgHdr *>(smem);
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);
    unsigned char *stage0 = smem + a.off_stage;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.z * a.Mb, mrows = min(a.Mb, a.M - m0);
    const int n0 = blockIdx.y * a.Nc, nrows = min(a.Nc, a.N - n0);
    const int c_begin = blockIdx.x * a.chunks_per_split;
    const int c_end = min(a.total_chunks, c_begin + a.chunks_per_split);

    if (tid == 0) {
        for (int i = 0; i < kWgMaxStages; ++i) {
            mbar_init(&hdr->full[i], PROD == PROD_SHIFT3D ? kWgProdWarps : kWgProdWarps / kWgGroups);
            mbar_init(&hdr->empty[i], 1);
        }
        mbar_init(&hdr->tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, (uint32_t)a.tmem_cols);
    }
    pdl_sync();  // barriers and tensor memory are set up while the previous kernel drains
    if (PROD == PROD_BNRELU)
        for (int k = tid; k < a.N; k += kWgThreads) {
            smem_sb[k] = a.x_sb[2 * k];
            smem_sb[a.N + k] = a.x_sb[2 * k + 1];
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp == 0) {
        // ONE elected thread runs the whole issue loop (see mma_loop of k_pw_conv): descriptors as 32-bit halves
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(128, a.sub_n, 0, 0);
            const uint64_t desc0 = smem_desc(smem_u32(stage0), 16, 1024, LAYOUT_SW128);
            const uint32_t d_hi = (uint32_t)(desc0 >> 32), d_lo0 = (uint32_t)desc0;
            const uint32_t b_off = (uint32_t)((a.Mt * kWgTileBytes) >> 4), sub_step = (uint32_t)((a.sub_n * 128) >> 4);
            int slot = 0;
            uint32_t phase = 0, acc = 0u;
            int pc = c_begin % a.cpi;
            for (int q = c_begin; q < c_end; ++q) {
                const int kvalid = min(kWgChunk, a.HW - pc * kWgChunk);
                const int ksteps = (kvalid + 15) >> 4;
                if (++pc == a.cpi) pc = 0;
                mbar_wait(&hdr->full[slot], phase);
                tc_fence_after();
                const uint32_t a_lo_s = d_lo0 + (uint32_t)((slot * a.stage_bytes) >> 4);
                for (int ks = 0; ks < ksteps; ++ks) {
                    // 32 bytes per K step inside the 128-byte swizzle atom
                    uint32_t a_lo = a_lo_s + (uint32_t)(ks * 2);
                    uint32_t tcol = tmem_base;
                    for (int mt = 0; mt < a.Mt; ++mt, a_lo += kWgTileBytes >> 4) {
                        uint32_t b_lo = a_lo_s + b_off + (uint32_t)(ks * 2);
                        for (int j = 0; j < a.n_sub; ++j, b_lo += sub_step, tcol += a.sub_n)
                            mma_bf16_lohi(tcol, a_lo, d_hi, b_lo, d_hi, idesc, acc);
                    }
                    acc = 1u;
                }
                mma_commit(&hdr->empty[slot]);
                if (++slot == a.stages) { slot = 0; phase ^= 1u; }
            }
            mma_commit(&hdr->tmem_full);
        }
        __syncwarp();
    } else {
        // all 16 non-MMA warps build the operand stages; warps 1-8 drain the accumulators afterwards
        const int pt = tid - 32, pw = pt >> 5;
        // operand rows loaded verbatim: G always, x unless it goes through the shift gather
        const int unit_rows = (PROD == PROD_SHIFT3D) ? mrows : mrows + nrows;
        const int total_units = unit_rows * 8;
        const ShiftSrc ssrc{a.x, a.shift, a.shift_dt, a.T, a.H, a.W, a.HW, a.N, (int)(((int64_t)a.NI * a.N * a.HW - 1) >> 1)};
        const int rpr = (int)a.rpr;
        // Plain / BN+ReLU operands: the warps form kWgGroups groups that fill alternate stages, each thread with all its
        // units of the stage (<= kWgUB) in flight at once -- two stages are being loaded while a third is consumed.  (A
        // warp's loads share one scoreboard: more than one batch per thread and stage would serialise on it.)  The shift
        // gather keeps every warp on every stage.
        constexpr int G = PROD == PROD_SHIFT3D ? 1 : kWgGroups;
        constexpr int GT = kWgProdThreads / G;          // threads per group
        const int grp = pt / GT, gt = pt - grp * GT;    // group, thread inside the group
        for (int q = c_begin + grp; q < c_end; q += G) {
            const int j = q - c_begin;                  // stage index of this CTA
            const int slot = j % a.stages;
            const uint32_t phase = (uint32_t)(j / a.stages) & 1u;
            const int img = q / a.cpi, pc = q - img * a.cpi;
            const int p0 = pc * kWgChunk;
            const int kvalid = min(kWgChunk, a.HW - p0);
            const int nchunks16 = ((kvalid + 15) >> 4) * 2;  // 16-byte chunks the MMAs of this stage will read
            unsigned char *abase = stage0 + (size_t)slot * a.stage_bytes, *bbase = abase + (size_t)a.Mt * kWgTileBytes;
            bool waited = false;
            for (int u0 = gt; u0 < total_units; u0 += kWgUB * GT) {
                uint32_t r[kWgUB][4];
#pragma unroll
                for (int b = 0; b < kWgUB; ++b) {
                    const int u = u0 + b * GT;
                    const int rowi = u >> 3, c = u & 7;
                    r[b][0] = r[b][1] = r[b][2] = r[b][3] = 0u;
                    if (u < total_units && c < nchunks16) {
                        const int p = p0 + c * 8;
                        if (rowi < mrows) load_unit<VEC>(a.g + ((int64_t)img * a.M + m0 + rowi) * a.HW + p, a.HW - p, r[b]);
                        else load_unit<VEC>(a.x + ((int64_t)img * a.N + n0 + rowi - mrows) * a.HW + p, a.HW - p, r[b]);
                    }
                }
                if (!waited) {  // the loads above do not need the slot: wait for it only now
                    mbar_wait(&hdr->empty[slot], phase ^ 1u);
                    waited = true;
                }
#pragma unroll
                for (int b = 0; b < kWgUB; ++b) {
                    const int u = u0 + b * GT;
                    const int rowi = u >> 3, c = u & 7;
                    if (u < total_units && c < nchunks16) {
                        if (PROD == PROD_BNRELU && rowi >= mrows) {  // applied after all loads of the batch were issued
                            const int k = n0 + rowi - mrows;
                            bn_relu_unit(r[b], smem_sb[k], smem_sb[a.N + k], a.HW - (p0 + c * 8));
                        }
                        unsigned char *d = (rowi < mrows) ? abase + (rowi >> 7) * kWgTileBytes + sw128_off(rowi & 127, c)
                                                          : bbase + sw128_off(rowi - mrows, c);
                        *reinterpret_cast<uint4 *>(d) = make_uint4(r[b][0], r[b][1], r[b][2], r[b][3]);
                    }
                }
            }
            if (!waited) mbar_wait(&hdr->empty[slot], phase ^ 1u);
            if (PROD == PROD_SHIFT3D) {
                // shifted activations: (channel, row run) items, 2-byte stores into the swizzled K-major block
                const int pend = p0 + kvalid, npad = nchunks16 * 8;
                const int clip = img / a.T, t = img - clip * a.T;
                const int h0 = p0 / a.W, h1 = (pend - 1) / a.W;
                const int g0 = h0 * rpr + ((p0 - h0 * a.W) >> 3), g1 = h1 * rpr + ((pend - 1 - h1 * a.W) >> 3);
                const int kq = lane & 7, mgq = lane >> 3;
                for (int rb = pw * 8 + kq; rb < nrows; rb += 8 * kWgProdWarps) {
                    const int k = n0 + rb;
                    const ShiftCh ch = shift_channel(ssrc, k, clip, t);
                    unsigned char *brow = bbase + (rb >> 3) * 1024 + (rb & 7) * 128;
                    const int sw = rb & 7;
                    for (int m = kvalid + mgq; m < npad; m += 4)  // zero the reduction padding
                        *reinterpret_cast<unsigned short *>(brow + (((m >> 3) ^ sw) << 4) + (m & 7) * 2) = 0;
                    auto store_run = [&](int h, int c0, const float (&o)[8]) {
                        const int m0 = h * a.W + c0 - p0;
                        const int lo = max(-m0, 0), hi = min(min(8, a.W - c0), kvalid - m0);
                        if (((m0 & 7) | lo) == 0 && hi == 8) {
                            *reinterpret_cast<uint4 *>(brow + (((m0 >> 3) ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                           pack_bf16x2(o[6], o[7]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if ((unsigned)(i - lo) < (unsigned)(hi - lo)) {
                                    const int m = m0 + i;
                                    *reinterpret_cast<unsigned short *>(brow + (((m >> 3) ^ sw) << 4) + (m & 7) * 2) =
                                        __bfloat16_as_ushort(__float2bfloat16_rn(o[i]));
                                }
                        }
                    };
                    RunLoad LA, LB;
                    int g = g0 + mgq, hA = 0, cA = 0, hB = 0, cB = 0;
                    if (g <= g1) {
                        run_coords(g, rpr, a.rpr_magic, hA, cA);
                        shift3d_load(ssrc, ch, hA, cA, LA);
                    }
                    while (g <= g1) {
                        float o[8];
                        const bool hasB = g + 4 <= g1;
                        if (hasB) {
                            run_coords(g + 4, rpr, a.rpr_magic, hB, cB);
                            shift3d_load(ssrc, ch, hB, cB, LB);
                        }
                        shift3d_compute(ssrc, ch, cA, LA, o);
                        store_run(hA, cA, o);
                        if (!hasB) break;
                        if (g + 8 <= g1) {
                            run_coords(g + 8, rpr, a.rpr_magic, hA, cA);
                            shift3d_load(ssrc, ch, hA, cA, LA);
                        }
                        shift3d_compute(ssrc, ch, cB, LB, o);
                        store_run(hB, cB, o);
                        g += 8;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->full[slot]);
        }
        if (warp < kProdWarp0) {
            // epilogue: warp -> TMEM lane quarter + every other 16-column chunk; fp32 partial slice of this pixel split
            const int q4 = warp & 3, half = (warp - kEpiWarp0) >> 2, row = q4 * 32 + lane;
            mbar_wait(&hdr->tmem_full, 0);
            tc_fence_after();
            float *dst = a.partial + (int64_t)blockIdx.x * a.M * a.N;
            const int nchunks = (nrows + 15) >> 4;
            for (int mt = 0; mt < a.Mt; ++mt) {
                const int ml = mt * 128 + row;
                const bool valid = ml < mrows;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + mt * a.Nc;
                for (int ci = half; ci < nchunks; ci += 2) {
                    const int c0 = ci * 16;
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
                        // slice layout [N, M]: for a fixed column the 32 lanes (= 32 consecutive rows m) write 128
                        // contiguous bytes
                        float *o = dst + (int64_t)(n0 + c0) * a.M + m0 + ml;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < nrows) o[(int64_t)j * a.M] = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

bool wg_plan(WgArgs &a, dim3 *grid, size_t *smem_bytes) {
    const int sb_bytes = round_up(2 * a.N * 4, 128);
    a.off_sb = kHdrBytes;
    a.off_stage = (uint32_t)round_up(kHdrBytes + sb_bytes, 1024);
    int best = 1 << 30, bmb = 0, bnb = 0;
    for (int mb = 1; mb <= 8; ++mb)
        for (int nb = 1; nb <= 8; ++nb) {
            const int Mb = round_up(cdiv(a.M, mb), 8), Mt = cdiv(Mb, 128);
            int Nc = round_up(cdiv(a.N, nb), 16);
            const int ns = cdiv(Nc, 256);
            Nc = round_up(Nc, 16 * ns);
            if (Mt > 4 || Mt * Nc > 512) continue;
            const int stage = Mt * kWgTileBytes + round_up(Nc * 128, 1024);
            if ((kSmemLimit - (int)a.off_stage) / stage < 2) continue;
            const int cost = (nb + mb) * 16 + mb * nb;  // operand re-reads first, CTA count second
            if (cost < best) { best = cost; bmb = mb; bnb = nb; }
        }
    if (!bmb) return false;
    a.Mb = round_up(cdiv(a.M, bmb), 8);
    a.Mt = cdiv(a.Mb, 128);
    a.Nc = round_up(cdiv(a.N, bnb), 16);
    a.n_sub = cdiv(a.Nc, 256);
    a.Nc = round_up(a.Nc, 16 * a.n_sub);
    a.sub_n = a.Nc / a.n_sub;
    a.stage_bytes = (uint32_t)(a.Mt * kWgTileBytes + round_up(a.Nc * 128, 1024));
    a.stages = (kSmemLimit - (int)a.off_stage) / (int)a.stage_bytes;
    if (a.stages > kWgMaxStages) a.stages = kWgMaxStages;
    int cols = 32;
    while (cols < a.Mt * a.Nc) cols <<= 1;
    a.tmem_cols = cols;
    a.cpi = cdiv(a.HW, kWgChunk);
    a.total_chunks = a.NI * a.cpi;
    a.rpr = (uint32_t)(a.W > 0 ? (a.W + 7) / 8 : 1);
    a.rpr_magic = 65536u / a.rpr + 1u;
    const int gy = cdiv(a.N, a.Nc), gz = cdiv(a.M, a.Mb);
    int want = sm_count() / (gy * gz);
    if (want < 1) want = 1;
    if (want > a.total_chunks) want = a.total_chunks;
    a.chunks_per_split = cdiv(a.total_chunks, want);
    const int splits = cdiv(a.total_chunks, a.chunks_per_split);
    *grid = dim3((unsigned)splits, (unsigned)gy, (unsigned)gz);
    *smem_bytes = (size_t)a.off_stage + (size_t)a.stages * a.stage_bytes;
    return true;
}

template <int PROD, int VEC> int wg_launch(const WgArgs &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_pw_wgrad<PROD, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_pw_wgrad): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    launch_kernel(k_pw_wgrad<PROD, VEC>, dim3(grid), dim3(kWgThreads), smem_bytes, s, a);
    return launched("k_pw_wgrad");
}

template <int PROD> int wg_launch_vec(const WgArgs &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    const bool base16 = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.g)) & 15) == 0;
    if (base16 && a.HW % 8 == 0) return wg_launch<PROD, 8>(a, grid, smem_bytes, s);
    if (base16 && a.HW % 4 == 0) return wg_launch<PROD, 4>(a, grid, smem_bytes, s);
    if (base16 && a.HW % 2 == 0) return wg_launch<PROD, 2>(a, grid, smem_bytes, s);
    return wg_launch<PROD, 1>(a, grid, smem_bytes, s);
}

}  // namespace

// fp32 [N, K] conv weight -> bf16 copies in both orientations, one launch: w_nk [N, K] (forward) and w_kn [K, N] (the
// weight matrix of the input-gradient GEMM).  32x32 tiles through shared memory: reads and both writes are coalesced.
__global__ void k_pw_weight_pack(const float *__restrict__ w, __nv_bfloat16 *__restrict__ w_nk, __nv_bfloat16 *__restrict__ w_kn,
                                 int N, int K) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        const float v = (n < N && k < K) ? w[(int64_t)n * K + k] : 0.f;
        tile[r][tx] = v;
        if (n < N && k < K) w_nk[(int64_t)n * K + k] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        if (n < N && k < K) w_kn[(int64_t)k * N + n] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// the same for MANY weights in one launch (grid.y = weight, blocks stride over its 32x32 tiles); items as rb_pw_pack_item_t
struct PackItem {
    const float *w;
    __nv_bfloat16 *w_nk, *w_kn;
    int N, K;
};
__global__ void k_pw_weight_pack_multi(const PackItem *__restrict__ items) {
    pdl_sync();
    __shared__ float tile[32][33];
    const PackItem it = items[blockIdx.y];
    const int tk = (it.K + 31) / 32, tn = (it.N + 31) / 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int t = blockIdx.x; t < tk * tn; t += gridDim.x) {
        const int k0 = (t % tk) * 32, n0 = (t / tk) * 32;
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int n = n0 + r, k = k0 + tx;
            const float v = (n < it.N && k < it.K) ? it.w[(int64_t)n * it.K + k] : 0.f;
            tile[r][tx] = v;
            if (n < it.N && k < it.K) it.w_nk[(int64_t)n * it.K + k] = __float2bfloat16_rn(v);
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int k = k0 + r, n = n0 + tx;
            if (n < it.N && k < it.K) it.w_kn[(int64_t)k * it.N + n] = __float2bfloat16_rn(tile[tx][r]);
        }
    }
}

int pw_weight_pack_multi(const void *items_device, int count, cudaStream_t s) {
    static_assert(sizeof(PackItem) == 32, "rb_pw_pack_item_t layout");
    launch_kernel(k_pw_weight_pack_multi, dim3(dim3(24, (unsigned)count)), dim3(256), 0, s, (const PackItem *)items_device);
    if (int rc = launched("k_pw_weight_pack_multi")) return rc;
    return launch_fence(s);  // the packed operands are RB_W_RESIDENT for every later launch
}

int pw_weight_pack(const float *w, void *w_nk, void *w_kn, int N, int K, cudaStream_t s) {
    dim3 grid((unsigned)cdiv(K, 32), (unsigned)cdiv(N, 32));
    launch_kernel(k_pw_weight_pack, dim3(grid), dim3(256), 0, s, w, (__nv_bfloat16 *)w_nk, (__nv_bfloat16 *)w_kn, N, K);
    if (int rc = launched("k_pw_weight_pack")) return rc;
    return launch_fence(s);  // the packed operands are RB_W_RESIDENT for every later launch
}

#ifdef RB_DEBUG_TRACE
static unsigned long long *g_pw_trace = nullptr;
void pw_conv_set_trace(void *p) { g_pw_trace = (unsigned long long *)p; }
unsigned long long *pw_conv_get_trace() { return g_pw_trace; }
#endif

void pw_conv_set_tuning(int min_n_splits) { g_min_n_splits = min_n_splits < 0 ? 0 : (min_n_splits > 16 ? 16 : min_n_splits); }

int pw_conv_forward(const void *x, const void *w, int w_dt, int w_trans, const void *residual, void *out, int NI, int K,
                    int N, int HW, const float *a_sb, const void *shift, int shift_dt, int T, int H, int W, cudaStream_t s,
                    double *stats, size_t stats_bytes, int *stats_splits) {
    PwArgs a{};
    a.x = (const __nv_bfloat16 *)x; a.w = w; a.w_dt = w_dt & ~RB_W_RESIDENT; a.w_resident = (w_dt & RB_W_RESIDENT) != 0; a.w_trans = w_trans; a.res = (const __nv_bfloat16 *)residual;
    a.out = (__nv_bfloat16 *)out; a.a_sb = a_sb; a.shift = shift; a.shift_dt = shift_dt;
    a.T = T; a.H = H; a.W = W; a.NI = NI; a.K = K; a.N = N; a.HW = HW;
#ifdef RB_DEBUG_TRACE
    a.trace = g_pw_trace;
#endif
    const int prod = shift ? PROD_SHIFT3D : (a_sb ? PROD_BNRELU : PROD_PLAIN);
    dim3 grid;
    size_t smem_bytes = 0;
    const int vec = pick_vec(a);
    if (!plan(a, prod, vec, &grid, &smem_bytes))
        return fail(RB_ERR_UNSUPPORTED, "pw_conv: no tiling for K=%d N=%d (weight block does not fit shared memory)", K, N);
    if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return fail(RB_ERR_INVALID_ARGUMENT, "pw_conv: x must be 4-byte aligned");
    if (stats != nullptr) {
        a.stats = stats;
        a.stats_splits = 2 * (int)grid.x;
        if (stats_bytes < (size_t)N * a.stats_splits * 2 * sizeof(double))
            return fail(RB_ERR_WORKSPACE, "pw_conv: statistics need %zu bytes", (size_t)N * a.stats_splits * 2 * sizeof(double));
        if (stats_splits) *stats_splits = a.stats_splits;
    }
    if (prod == PROD_SHIFT3D) return launch_vec<PROD_SHIFT3D>(a, vec, grid, smem_bytes, s);
    if (prod == PROD_BNRELU) return launch_vec<PROD_BNRELU>(a, vec, grid, smem_bytes, s);
    return launch_vec<PROD_PLAIN>(a, vec, grid, smem_bytes, s);
}

// scratch bytes needed by pw_conv_wgrad: fp32 [splits, M, N]
int wg_reduce(const float *partial, float *dw, int splits, int M, int N, cudaStream_t s);  // pw_wgrad3.cu

size_t pw_conv_wgrad_workspace(int NI, int M, int N, int HW) {
    WgArgs a{};
    a.NI = NI; a.M = M; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem_bytes = 0;
    if (!wg_plan(a, &grid, &smem_bytes)) return 0;
    return (size_t)grid.x * M * N * sizeof(float);
}

int pw_conv_wgrad(const void *g, const void *x, float *dw, int NI, int M, int N, int HW, const float *x_sb,
                  const void *shift, int shift_dt, int T, int H, int W, void *workspace, cudaStream_t s) {
    WgArgs a{};
    a.g = (const __nv_bfloat16 *)g; a.x = (const __nv_bfloat16 *)x; a.partial = (float *)workspace;
    a.x_sb = x_sb; a.shift = shift; a.shift_dt = shift_dt; a.T = T; a.H = H; a.W = W;
    a.NI = NI; a.M = M; a.N = N; a.HW = HW;
    const int prod = shift ? PROD_SHIFT3D : (x_sb ? PROD_BNRELU : PROD_PLAIN);
    dim3 grid;
    size_t smem_bytes = 0;
    if (!wg_plan(a, &grid, &smem_bytes)) return fail(RB_ERR_UNSUPPORTED, "pw_conv_wgrad: no tiling for M=%d N=%d", M, N);
    if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return fail(RB_ERR_INVALID_ARGUMENT, "pw_conv_wgrad: x must be 4-byte aligned");
    int rc;
    if (prod == PROD_SHIFT3D) rc = wg_launch_vec<PROD_SHIFT3D>(a, grid, smem_bytes, s);
    else if (prod == PROD_BNRELU) rc = wg_launch_vec<PROD_BNRELU>(a, grid, smem_bytes, s);
    else rc = wg_launch_vec<PROD_PLAIN>(a, grid, smem_bytes, s);
    if (rc) return rc;
    return wg_reduce(a.partial, dw, (int)grid.x, M, N, s);
}

}  // namespace rb

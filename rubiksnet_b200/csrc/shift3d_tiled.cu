// Tiled sm_100a kernels for the 3D learnable shift (forward and fused backward).
//
// Replaces rubiks_shift_3d_forward_cuda / rubiks_shift_3d_backward_cuda /
// rubiks_shift_3d_backward_input{,_s1p0}_cuda / normalize_shift_grad_3d_cuda
// (/root/reference/cuda_src/rubiks3d_kernels.cu:15-205,218-452,455-929,932-960) for the geometries
// RubiksNet uses: temporal stride 1, no padding, spatial stride (1,1) or (2,2), no quantize.
//
// Design (B200-first, HBM-bound op):
//  * A CTA owns one tile = (clip n, CG consecutive channels, a band of TH destination rows) for ALL T
//    frames.  For each source frame the rows that tile needs are ONE contiguous range of HBM, so a
//    single elected thread per frame issues one 1-D TMA bulk copy (cp.async.bulk global->shared,
//    16-byte aligned body; the <16 B unaligned head/tail is patched with plain loads) that completes
//    on a per-frame mbarrier.  No LSU instruction is spent on moving input data.
//  * Each warp works on ONE channel, so floor(shift), the remainders and the 8 trilinear weights are
//    warp-uniform registers (the reference recomputes them plus 4 div + 4 mod per element).
//  * Frames stream through registers: the bilinear (H,W) interpolation B[t] of a source frame is
//    computed once and used by the two destination frames that touch it,
//        dst[t] = w0 * B[t + fT] + w1 * B[t + fT + 1]
//    which is exactly the association order of the reference (:193-203, :709-719).  Every source
//    element is read from HBM once and every destination element written once.
//  * Backward is ONE kernel in the adjoint ("dual") form: staged tensor = out_grad, and for every
//    INPUT position i it produces
//        x_grad[i]   = sum_taps wT wH wW og[tap]                       (reference :455-929)
//        dL/dshift  += x[i] * sum_taps (sign on one axis, weights on the others) og[tap]
//    The second line is the reference's per-output-pixel sum (:283-446) regrouped by input index:
//    identical terms, so out_grad is read once, x once, x_grad written once (the reference reads
//    out_grad twice, x eight-fold through L1/L2 and does 3 global atomics per element).
//    Channels with an exactly-integer shift component take a warp-uniform slow path that applies the
//    reference's "small := floor-1" rule (:290-298,359-431) tap by tap.
//  * dL/dshift: registers -> warp shuffles -> CTA smem -> one double per (tile, channel, axis) ->
//    finalize kernel (deterministic, also normalises).  No atomics, no memsets, no GEMV.
#include "common.cuh"

namespace rb {

int shift3d_finalize(const double *partial, int parts, void *shift_grad, int dt, int sdt, int C, int normalize,
                     double factor, cudaStream_t s);

static constexpr int kNW = 8;          // warps per CTA
static constexpr int kMaxFrames = 16;  // frames resident in shared memory at once
static constexpr int kHdrBytes = 512;  // mbarriers + reduction scratch
static constexpr int kZeroBytes = 32;  // zeroed region at the head of every stage (target of invalid taps)

enum { MODE_FWD = 0, MODE_BWD = 1 };

struct TileCfg {
    int CG, WPC, TH, row_tiles, K, groups;
    int stage_bytes;  // shared bytes per frame stage (multiple of 16)
    int smem_bytes;
};

struct TiledArgs {
    const void *src;    // staged tensor: x (FWD) / out_grad (BWD)
    void *dst;          // out (FWD) / x_grad (BWD); may be null in BWD
    const void *xin;    // BWD: x, read once per input position; null if no shift grad wanted
    const void *shift;
    double *partial;    // BWD: [C][parts][3]
    int sdt;
    int N, Tn, C;
    int Hs, Ws, Hd, Wd; // source / destination plane extents
    int S;              // spatial stride (1 or 2)
    int mode2d;         // 2D shift (cuda_src/rubiks2d_kernels.cu): shift is [2, C] (H, W); the Tn frames of a CTA are independent images
    TileCfg cfg;
};

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
template <typename T> __device__ __forceinline__ float lds(const T *p) { return ld<float, T>(p); }
template <typename T> __device__ __forceinline__ float ldg_stream(const T *p) { return ld<float, T>(p); }
template <> __device__ __forceinline__ float ldg_stream<float>(const float *p) { return __ldcs(p); }

__device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b;
    return q - ((a % b != 0) && ((a < 0) != (b < 0)));
}
__device__ __forceinline__ int ceil_div(int a, int b) { return -floor_div(-a, b); }

// Per-axis coefficients of the integer-shift slow path for a tap at numerator offset
// delta in {-1,0,1} relative to i + floor(-s):
//   coef_a   weight of x_grad                                   (reference :709-719)
//   coef_b   interpolation weight used by the OTHER axes' gradients
//   coef_sg  sign of this axis' finite difference
// non-integer axis: a = b = (0, 1-r', r'),  sg = (0, +1, -1)
// integer axis    : a = (0,1,0), b = (0,0,1), sg = (+1,0,-1)   ("small := floor-1", :290-298)
__device__ __forceinline__ float coef_a(int d, float r) { return d == 0 ? 1.f - r : (d == 1 ? r : 0.f); }
__device__ __forceinline__ float coef_b(int d, float r, bool integer) {
    return integer ? (d == 1 ? 1.f : 0.f) : coef_a(d, r);
}
__device__ __forceinline__ float coef_sg(int d, bool integer) {
    if (d == 1) return -1.f;
    return integer ? (d == -1 ? 1.f : 0.f) : (d == 0 ? 1.f : 0.f);
}

// S2 (backward, spatial stride 2): of the four (dy, dx) taps of an input position exactly one has even numerators, the
// other three read the zero slot -- that path keeps ONE offset per position and multiplies by the selected weights, which
// is the general expression with the zero terms dropped (bit-identical: x*w + 0*w' == x*w).
template <typename T, int MODE, int K, bool S2 = false>
__global__ void __launch_bounds__(kNW * 32) k_shift3d_tiled(const TiledArgs a) {
    pdl_sync();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);           // [kMaxFrames]
    double *red = reinterpret_cast<double *>(smem_raw + 128);          // [kNW][3]
    unsigned char *stages = smem_raw + kHdrBytes;
    constexpr int ES = (int)sizeof(T);
    constexpr int ZOFF = kZeroBytes / ES;  // element index of source element 0 relative to the stage pointer

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileCfg &cf = a.cfg;
    const int CG = cf.CG, WPC = cf.WPC;
    int b = blockIdx.x;
    const int rtile = b % cf.row_tiles; b /= cf.row_tiles;
    const int grp = b % cf.groups;
    const int n = b / cf.groups;
    const int c0 = grp * CG;
    const int cl = warp / WPC, wsub = warp - cl * WPC;
    const int c = c0 + cl;
    const bool c_ok = c < a.C;
    const int cc = c_ok ? c : a.C - 1;
    const int hd0 = rtile * cf.TH;
    const int th = min(cf.TH, a.Hd - hd0);
    const int S = a.S, Hs = a.Hs, Ws = a.Ws, Hd = a.Hd, Wd = a.Wd, Tn = a.Tn;
    const int HWs = Hs * Ws, HWd = Hd * Wd;

    // ---- warp-uniform channel parameters --------------------------------------------------------
    float sT, sH, sW;
    if (a.mode2d) {
        sT = 0.f;
        sH = ld_param<float>(a.shift, a.sdt, cc);
        sW = ld_param<float>(a.shift, a.sdt, a.C + cc);
        if (MODE == MODE_BWD) {
            // the 2D shift gradient treats |r| < 1e-7 as an exact integer shift (rubiks2d_kernels.cu:189): snap the shift
            const float fh = floorf(sH), fw = floorf(sW);
            if (fabsf(sH - fh) < 1e-7f) sH = fh;
            if (fabsf(sW - fw) < 1e-7f) sW = fw;
        }
    } else {
        sT = ld_param<float>(a.shift, a.sdt, cc);
        sH = ld_param<float>(a.shift, a.sdt, a.C + cc);
        sW = ld_param<float>(a.shift, a.sdt, 2 * a.C + cc);
    }
    if (MODE == MODE_BWD) { sT = -sT; sH = -sH; sW = -sW; }  // adjoint gathers at the negated shift (:505-507)
    const int fT = floor3d(sT), fH = floor3d(sH), fW = floor3d(sW);
    const float rT = sT - fT, rH = sH - fH, rW = sW - fW;
    // 2D: no temporal component (frame t reads frame t with weight 1; must not trigger the integer rule)
    const bool intT = !a.mode2d && (rT == 0.f), intH = (rH == 0.f), intW = (rW == 0.f);
    const bool slow = (MODE == MODE_BWD) && (intT || intH || intW);

    // ---- rows of the source plane this CTA stages (CTA-uniform) ---------------------------------
    int rlo = 0, rhi = Hs - 1;
    if (CG == 1) {
        if (MODE == MODE_FWD) {
            rlo = hd0 * S + fH;
            rhi = (hd0 + th - 1) * S + fH + 1;
        } else {
            rlo = ceil_div(hd0 + fH - (intH ? 1 : 0), S);
            rhi = floor_div(hd0 + th - 1 + fH + 1, S);
        }
        rlo = max(rlo, 0);
        rhi = min(rhi, Hs - 1);
    }
    const int cg_eff = min(CG, a.C - c0);
    const int count = (rhi >= rlo) ? (cg_eff - 1) * HWs + (rhi - rlo + 1) * Ws : 0;  // elements per frame
    const bool any_data = count > 0;

    // ---- issue the per-frame bulk copies ---------------------------------------------------------
    if (tid == 0) {
        for (int t = 0; t < Tn; ++t) mbar_init(&bars[t], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const T *src = reinterpret_cast<const T *>(a.src);
    auto frame_src = [&](int t) -> const T * {
        return src + ((int64_t)(n * Tn + t) * a.C + c0) * HWs + (int64_t)rlo * Ws;
    };
    if (any_data) {
        if (tid < Tn) {
            const uintptr_t sb = reinterpret_cast<uintptr_t>(frame_src(tid));
            const uintptr_t eb = sb + (uintptr_t)count * ES;
            const uintptr_t ab = (sb + 15) & ~(uintptr_t)15, ae = eb & ~(uintptr_t)15;
            unsigned char *stage = stages + (size_t)tid * cf.stage_bytes;
            if (ae > ab) {
                const uint32_t bytes = (uint32_t)(ae - ab);
                mbar_arrive_expect_tx(&bars[tid], bytes);
                bulk_g2s(stage + kZeroBytes + (ab - (sb & ~(uintptr_t)15)), reinterpret_cast<const void *>(ab), bytes,
                         &bars[tid]);
            } else {
                mbar_arrive_expect_tx(&bars[tid], 0);
            }
        }
        // unaligned head / tail elements (< 16 bytes each) with plain loads: 16 slots per frame
        {
            const int t = tid >> 4, slot = tid & 15;
            if (t < Tn) {
                const T *fs = frame_src(t);
                const uintptr_t sb = reinterpret_cast<uintptr_t>(fs);
                const uintptr_t eb = sb + (uintptr_t)count * ES;
                const uintptr_t ab = (sb + 15) & ~(uintptr_t)15, ae = eb & ~(uintptr_t)15;
                const int mis = (int)(sb & 15);
                T *sdata = reinterpret_cast<T *>(stages + (size_t)t * cf.stage_bytes + kZeroBytes + mis);
                int head, tail_start;
                if (ae > ab) {
                    head = (int)((ab - sb) / ES);
                    tail_start = (int)((ae - sb) / ES);
                } else {
                    head = count;
                    tail_start = count;
                }
                if (slot < 8) {
                    if (slot < head) sdata[slot] = fs[slot];
                } else {
                    const int j = tail_start + (slot - 8);
                    if (j < count) sdata[j] = fs[j];
                }
            }
        }
        // zero region at the head of every stage
        if (tid < Tn * (kZeroBytes / 4))
            reinterpret_cast<uint32_t *>(stages + (size_t)(tid / (kZeroBytes / 4)) * cf.stage_bytes)[tid % (kZeroBytes / 4)] = 0u;
    }

    // ---- per-item tap offsets (element index relative to the frame's stage pointer; 0 = zero) ----
    const int P = th * Wd;  // destination positions per channel in this tile
    // position k of a thread: lanes interleaved (adjacent lanes -> adjacent elements).  (K CONSECUTIVE positions per thread,
    // one K-element vector access to x / x_grad per frame, was measured on the stride-2 backward and is slower in bf16:
    // 0.822 -> 0.911 ms at 72 ch x 112x112, profiles/r02z_bench_shift.log vs gpurun_out/wg3_shift.log.)
    const int pstride = WPC * 32;
    const int p0 = wsub * 32 + lane;
    int off[S2 ? 1 : K][4];
    int osel[S2 ? K : 1];
    float wWs[S2 ? K : 1], wHs[S2 ? K : 1], sgH[S2 ? K : 1], sgW[S2 ? K : 1];
    const float wH0 = 1.f - rH, wH1 = rH, wW0 = 1.f - rW, wW1 = rW, wT0 = 1.f - rT, wT1 = rT;
    const int chan_base = ZOFF + cl * HWs - rlo * Ws;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = p0 + k * pstride;
        const bool ok = c_ok && p < P && any_data;
        const int hl = p / Wd, wd = p - hl * Wd, hd = hd0 + hl;
        if (S2) {
            const int dy = (hd + fH) & 1, dx = (wd + fW) & 1;  // the tap whose numerators are even
            const int hs = (hd + fH + dy) >> 1, ws = (wd + fW + dx) >> 1;
            const bool v = ok && hs >= rlo && hs <= rhi && ws >= 0 && ws < Ws;
            osel[k] = v ? chan_base + hs * Ws + ws : 0;
            wHs[k] = dy ? wH1 : wH0;
            wWs[k] = dx ? wW1 : wW0;
            sgH[k] = dy ? -1.f : 1.f;
            sgW[k] = dx ? -1.f : 1.f;
            continue;
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                int hs, ws;
                bool v = ok;
                if (MODE == MODE_FWD) {
                    hs = hd * S + fH + dy;
                    ws = wd * S + fW + dx;
                } else {
                    const int nh = hd + fH + dy, nw = wd + fW + dx;
                    v = v && ((nh & (S - 1)) == 0) && ((nw & (S - 1)) == 0);
                    hs = nh >> (S - 1);
                    ws = nw >> (S - 1);
                }
                v = v && hs >= rlo && hs <= rhi && ws >= 0 && ws < Ws;
                off[S2 ? 0 : k][dy * 2 + dx] = v ? chan_base + hs * Ws + ws : 0;
            }
    }
    __syncthreads();  // head/tail/zero stores visible to every warp

    T *dst = reinterpret_cast<T *>(a.dst);
    const T *xin = reinterpret_cast<const T *>(a.xin);
    const bool want_dst = dst != nullptr;
    const bool want_grad = (MODE == MODE_BWD) && xin != nullptr;
    const int64_t dst_chan = ((int64_t)n * Tn * a.C + c) * HWd + (int64_t)hd0 * Wd;  // frame 0
    const int64_t dst_fs = (int64_t)a.C * HWd;
    auto stage_ptr = [&](int t) -> const T * {
        const int mis = (int)(reinterpret_cast<uintptr_t>(frame_src(t)) & 15);
        return reinterpret_cast<const T *>(stages + (size_t)t * cf.stage_bytes + mis);
    };

    float accT = 0.f, accH = 0.f, accW = 0.f;
    if (!slow && S2) {
        // ================= streaming path, stride-2 backward =================
        // ncu on the generic loop below (72 ch x 112x112, profiles/r02z_ncu_s2_bwd.txt): 84 warp instructions per position and
        // frame at 67 % issue utilisation, and 27 % of the stall samples on the first use of the x value loaded at the top of
        // the same step.  Here the per-position predicates and the x / x_grad pointers are hoisted out of the frame loop
        // (one pointer bump per frame), and x of frame t is fetched one step ahead of its use.
        bool okk[K];
#pragma unroll
        for (int k = 0; k < K; ++k) okk[k] = c_ok && p0 + k * pstride < P;
        const T *xnext = want_grad ? xin + dst_chan + p0 : nullptr;  // x of the frame fetched next
        T *dcur = want_dst ? dst + dst_chan + p0 : nullptr;          // x_grad of the frame completed next
        T xpre[K];
        float pB[K], pDH[K], pDW[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { pB[k] = 0.f; pDH[k] = 0.f; pDW[k] = 0.f; xpre[k] = cvt<T, float>(0.f); }
        for (int step = 0; step <= Tn; ++step) {
            const int ts = step + fT;
            const bool have = any_data && ts >= 0 && ts < Tn;
            const bool emit = step >= 1;  // destination frame step - 1 is completed by this step
            float xv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) xv[k] = ld<float, T>(&xpre[k]);  // fetched during the previous step
            if (want_grad && step < Tn) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (okk[k]) xpre[k] = xnext[k * pstride];
                xnext += dst_fs;
            }
            const T *sp = nullptr;
            if (have) {
                mbar_wait(&bars[ts], 0);
                sp = stage_ptr(ts);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float cB = 0.f, cDH = 0.f, cDW = 0.f;
                if (have) {
                    const float q = lds<T>(sp + osel[k]);
                    const float rw = q * wWs[k];
                    cB = wHs[k] * rw;
                    cDH = sgH[k] * rw;
                    cDW = wHs[k] * (sgW[k] * q);
                }
                if (emit) {
                    const float v = a.mode2d ? pB[k] : wT0 * pB[k] + wT1 * cB;
                    if (want_dst && okk[k]) dcur[k * pstride] = cvt<T, float>(v);
                    accT += xv[k] * (pB[k] - cB);
                    accH += xv[k] * (a.mode2d ? pDH[k] : wT0 * pDH[k] + wT1 * cDH);
                    accW += xv[k] * (a.mode2d ? pDW[k] : wT0 * pDW[k] + wT1 * cDW);
                }
                pB[k] = cB;
                pDH[k] = cDH;
                pDW[k] = cDW;
            }
            if (emit && want_dst) dcur += dst_fs;
        }
    } else if (!slow) {
        // ================= streaming path =================
        float pB[K], pDH[K], pDW[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { pB[k] = 0.f; pDH[k] = 0.f; pDW[k] = 0.f; }
        for (int step = 0; step <= Tn; ++step) {
            const int ts = step + fT;  // source frame entering the window
            const int td = step - 1;   // destination frame completed by it
            const bool have = any_data && ts >= 0 && ts < Tn;
            float xv[K];
            if (MODE == MODE_BWD) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int p = p0 + k * pstride;
                    xv[k] = (want_grad && td >= 0 && c_ok && p < P) ? ldg_stream<T>(xin + dst_chan + td * dst_fs + p) : 0.f;
                }
            }
            const T *sp = nullptr;
            if (have) {
                mbar_wait(&bars[ts], 0);
                sp = stage_ptr(ts);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float cB = 0.f, cDH = 0.f, cDW = 0.f;
                if (have) {
                    const float q00 = lds<T>(sp + off[S2 ? 0 : k][0]), q01 = lds<T>(sp + off[S2 ? 0 : k][1]);
                    const float q10 = lds<T>(sp + off[S2 ? 0 : k][2]), q11 = lds<T>(sp + off[S2 ? 0 : k][3]);
                    const float r0 = q00 * wW0 + q01 * wW1, r1 = q10 * wW0 + q11 * wW1;
                    cB = wH0 * r0 + wH1 * r1;
                    if (MODE == MODE_BWD) {
                        cDH = r0 - r1;
                        cDW = wH0 * (q00 - q01) + wH1 * (q10 - q11);
                    }
                }
                if (td >= 0) {
                    const int p = p0 + k * pstride;
                    const float v = a.mode2d ? pB[k] : wT0 * pB[k] + wT1 * cB;
                    if (want_dst && c_ok && p < P) dst[dst_chan + td * dst_fs + p] = cvt<T, float>(v);
                    if (MODE == MODE_BWD) {
                        accT += xv[k] * (pB[k] - cB);
                        accH += xv[k] * (a.mode2d ? pDH[k] : wT0 * pDH[k] + wT1 * cDH);
                        accW += xv[k] * (a.mode2d ? pDW[k] : wT0 * pDW[k] + wT1 * cDW);
                    }
                }
                pB[k] = cB;
                if (MODE == MODE_BWD) { pDH[k] = cDH; pDW[k] = cDW; }
            }
        }
    } else {
        // ================= integer-shift slow path (BWD only, warp-uniform) =================
        if (any_data)
            for (int t = 0; t < Tn; ++t) mbar_wait(&bars[t], 0);
        const bool zero_shift = (sT == 0.f && sH == 0.f && sW == 0.f);  // :561-576
        auto tap = [&](int ts, int nh, int nw) -> float {
            bool v = any_data && ts >= 0 && ts < Tn && ((nh & (S - 1)) == 0) && ((nw & (S - 1)) == 0);
            const int hs = nh >> (S - 1), ws = nw >> (S - 1);
            v = v && hs >= rlo && hs <= rhi && ws >= 0 && ws < Ws;
            return v ? lds<T>(stage_ptr(ts) + chan_base + hs * Ws + ws) : 0.f;
        };
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const int p = p0 + k * pstride;
            if (!(c_ok && p < P)) continue;
            const int hl = p / Wd, wd = p - hl * Wd, hd = hd0 + hl;
#pragma unroll 1
            for (int td = 0; td < Tn; ++td) {
                const int t0 = td + fT, h0 = hd + fH, w0 = wd + fW;
                if (want_dst) {
                    float v;
                    if (zero_shift) {
                        v = tap(t0, h0, w0);
                    } else {
                        const float q111 = tap(t0, h0, w0), q112 = tap(t0, h0, w0 + 1);
                        const float q121 = tap(t0, h0 + 1, w0), q122 = tap(t0, h0 + 1, w0 + 1);
                        const float q211 = tap(t0 + 1, h0, w0), q212 = tap(t0 + 1, h0, w0 + 1);
                        const float q221 = tap(t0 + 1, h0 + 1, w0), q222 = tap(t0 + 1, h0 + 1, w0 + 1);
                        v = wH0 * (q111 * wW0 + q112 * wW1) + wH1 * (q121 * wW0 + q122 * wW1);
                        if (!a.mode2d) v = wT0 * v + wT1 * (wH0 * (q211 * wW0 + q212 * wW1) + wH1 * (q221 * wW0 + q222 * wW1));
                    }
                    dst[dst_chan + td * dst_fs + p] = cvt<T, float>(v);
                }
                if (want_grad && a.mode2d) {
                    // 2D rule (rubiks2d_kernels.cu:189-253) in adjoint form: regular axis = difference of the two taps;
                    // exact-integer axis = 0.5 * central difference (taps -1 and +1); the other axis always interpolates
                    // with (1 - r, r)
                    float gH = 0.f, gW = 0.f;
#pragma unroll 1
                    for (int dy = -1; dy <= 1; ++dy) {
                        const float bh = coef_a(dy, rH);
                        const float sh = intH ? (dy == -1 ? 0.5f : (dy == 1 ? -0.5f : 0.f)) : (dy == 0 ? 1.f : (dy == 1 ? -1.f : 0.f));
#pragma unroll 1
                        for (int dx = -1; dx <= 1; ++dx) {
                            const float bw = coef_a(dx, rW);
                            const float sw = intW ? (dx == -1 ? 0.5f : (dx == 1 ? -0.5f : 0.f)) : (dx == 0 ? 1.f : (dx == 1 ? -1.f : 0.f));
                            const float cH = sh * bw, cW = bh * sw;
                            if (cH == 0.f && cW == 0.f) continue;
                            const float qq = tap(t0, h0 + dy, w0 + dx);
                            gH += cH * qq;
                            gW += cW * qq;
                        }
                    }
                    const float xv = ld<float, T>(xin + dst_chan + td * dst_fs + p);
                    accH += xv * gH;
                    accW += xv * gW;
                } else if (want_grad) {
                    float gT = 0.f, gH = 0.f, gW = 0.f;
#pragma unroll 1
                    for (int dt = -1; dt <= 1; ++dt) {
                        const float bt = coef_b(dt, rT, intT), st = coef_sg(dt, intT);
#pragma unroll 1
                        for (int dy = -1; dy <= 1; ++dy) {
                            const float bh = coef_b(dy, rH, intH), sh = coef_sg(dy, intH);
#pragma unroll 1
                            for (int dx = -1; dx <= 1; ++dx) {
                                const float bw = coef_b(dx, rW, intW), sw = coef_sg(dx, intW);
                                const float cT = st * bh * bw, cH = bt * sh * bw, cW = bt * bh * sw;
                                if (cT == 0.f && cH == 0.f && cW == 0.f) continue;
                                const float qq = tap(t0 + dt, h0 + dy, w0 + dx);
                                gT += cT * qq;
                                gH += cH * qq;
                                gW += cW * qq;
                            }
                        }
                    }
                    const float xv = ld<float, T>(xin + dst_chan + td * dst_fs + p);
                    accT += xv * gT;
                    accH += xv * gH;
                    accW += xv * gW;
                }
            }
        }
    }

    if (MODE == MODE_BWD && a.partial != nullptr) {
        accT = warp_sum(accT);
        accH = warp_sum(accH);
        accW = warp_sum(accW);
        if (lane == 0) {
            red[warp * 3 + 0] = (double)accT;
            red[warp * 3 + 1] = (double)accH;
            red[warp * 3 + 2] = (double)accW;
        }
        __syncthreads();
        if (tid < CG * 3) {
            const int l = tid / 3, ax = tid - l * 3;
            if (c0 + l < a.C) {
                double s = 0;
                for (int w = 0; w < WPC; ++w) s += red[(l * WPC + w) * 3 + ax];
                const int parts = a.N * cf.row_tiles;
                a.partial[((int64_t)(c0 + l) * parts + (n * cf.row_tiles + rtile)) * 3 + ax] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side

static constexpr int kSmemTarget = 72 * 1024;   // ~3 CTAs / SM
static constexpr int kSmemLimit = 227 * 1024;

static int stage_bytes_for(int count_elems, int es) {
    int b = kZeroBytes + 16 + count_elems * es;
    return (b + 15) & ~15;
}

// Picks the tile shape.  mode: MODE_FWD stages x (extent Hs x Ws = input plane), MODE_BWD stages out_grad.
static bool pick_cfg(int mode, int es, int Tn, int C, int Hs, int Ws, int Hd, int Wd, int S, int kmax,
                     TileCfg *out) {
    if (Tn > kMaxFrames || Tn <= 0) return false;
    const int lanes = kNW * 32;
    const int Pplane = Hd * Wd;
    TileCfg best;
    bool found = false;
    // regime A: whole planes, CG channels per CTA
    for (int CG = kNW; CG >= 1 && !found; CG >>= 1) {
        if (CG > 1 && CG > C) continue;
        const int WPC = kNW / CG;
        const int k = cdiv(Pplane, WPC * 32);
        if (k > kmax) continue;
        const int sb = stage_bytes_for(CG * Hs * Ws, es);
        const int smem = kHdrBytes + Tn * sb;
        if (smem > kSmemTarget) continue;
        best.CG = CG; best.WPC = WPC; best.TH = Hd; best.row_tiles = 1; best.K = k;
        best.stage_bytes = sb; best.smem_bytes = smem;
        found = true;
    }
    if (!found) {
        // regime B: one channel, band of TH destination rows
        int th_k = (lanes * kmax) / Wd;
        if (th_k < 1) return false;
        if (th_k > Hd) th_k = Hd;
        for (int pass = 0; pass < 2 && !found; ++pass) {
            const int budget = pass == 0 ? kSmemTarget : kSmemLimit;
            for (int TH = th_k; TH >= 1; --TH) {
                int rows = (mode == MODE_FWD) ? (TH - 1) * S + 2 : (TH + 2) / S + 2;
                if (rows > Hs) rows = Hs;
                const int sb = stage_bytes_for(rows * Ws, es);
                const int smem = kHdrBytes + Tn * sb;
                if (smem > budget) continue;
                const int row_tiles = cdiv(Hd, TH);
                const int THb = cdiv(Hd, row_tiles);  // balance the bands
                best.CG = 1; best.WPC = kNW; best.TH = THb; best.row_tiles = cdiv(Hd, THb);
                best.K = cdiv(THb * Wd, lanes);
                best.stage_bytes = sb; best.smem_bytes = smem;
                found = true;
                break;
            }
        }
    }
    if (!found) return false;
    int K = 1;
    while (K < best.K) K <<= 1;
    best.K = K;
    best.groups = cdiv(C, best.CG);
    *out = best;
    return true;
}

static constexpr int kKmaxFwd = 8;
static constexpr int kKmaxBwd = 4;

bool shift3d_tiled_supported(int dt, const Geom3 &g, int quantize) {
    if (quantize) return false;
    if (dt != RB_F32 && dt != RB_F16 && dt != RB_BF16) return false;
    if (g.sT != 1 || g.pT != 0 || g.pH != 0 || g.pW != 0) return false;
    if (g.sH != g.sW || (g.sH != 1 && g.sH != 2)) return false;
    if (g.Ho <= 0 || g.Wo <= 0) return false;
    const int es = (int)dtype_size(dt);
    TileCfg c;
    if (!pick_cfg(MODE_FWD, es, g.T, g.C, g.H, g.W, g.Ho, g.Wo, g.sH, kKmaxFwd, &c)) return false;
    if (!pick_cfg(MODE_BWD, es, g.T, g.C, g.Ho, g.Wo, g.H, g.W, g.sH, kKmaxBwd, &c)) return false;
    if ((int64_t)g.N * c.groups * c.row_tiles > 0x7fffffffLL) return false;
    return true;
}

template <typename T, int MODE, int K, bool S2> static int launch_ks(const TiledArgs &a, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_shift3d_tiled<T, MODE, K, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemLimit);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    const unsigned blocks = (unsigned)((int64_t)a.N * a.cfg.groups * a.cfg.row_tiles);
    launch_kernel(k_shift3d_tiled<T, MODE, K, S2>, dim3(blocks), dim3(kNW * 32), a.cfg.smem_bytes, s, a);
    return launched(MODE == MODE_FWD ? "k_shift3d_tiled<fwd>" : "k_shift3d_tiled<bwd>");
}
template <typename T, int MODE, int K> static int launch_k(const TiledArgs &a, cudaStream_t s) {
    if (MODE == MODE_BWD && a.S == 2) return launch_ks<T, MODE, K, (MODE == MODE_BWD)>(a, s);
    return launch_ks<T, MODE, K, false>(a, s);
}

template <typename T, int MODE> static int launch_mode(const TiledArgs &a, cudaStream_t s) {
    switch (a.cfg.K) {
        case 1: return launch_k<T, MODE, 1>(a, s);
        case 2: return launch_k<T, MODE, 2>(a, s);
        case 4: return launch_k<T, MODE, 4>(a, s);
        case 8: return launch_k<T, MODE, 8>(a, s);
    }
    return fail(RB_ERR_UNSUPPORTED, "tiled kernel: K=%d", a.cfg.K);
}

template <int MODE> static int launch_dtype(int dt, const TiledArgs &a, cudaStream_t s) {
    switch (dt) {
        case RB_F32: return launch_mode<float, MODE>(a, s);
        case RB_F16: return launch_mode<__half, MODE>(a, s);
        case RB_BF16: return launch_mode<__nv_bfloat16, MODE>(a, s);
    }
    return fail(RB_ERR_UNSUPPORTED, "tiled kernel: dtype %d", dt);
}

int shift3d_forward_tiled(const void *x, const void *shift, void *out, int dt, int sdt, const Geom3 &g,
                          cudaStream_t s) {
    TiledArgs a{};
    a.src = x; a.dst = out; a.xin = nullptr; a.shift = shift; a.partial = nullptr; a.sdt = sdt;
    a.N = g.N; a.Tn = g.T; a.C = g.C; a.Hs = g.H; a.Ws = g.W; a.Hd = g.Ho; a.Wd = g.Wo; a.S = g.sH;
    if (!pick_cfg(MODE_FWD, (int)dtype_size(dt), g.T, g.C, g.H, g.W, g.Ho, g.Wo, g.sH, kKmaxFwd, &a.cfg))
        return fail(RB_ERR_UNSUPPORTED, "tiled forward: no tile configuration");
    return launch_dtype<MODE_FWD>(dt, a, s);
}

size_t shift3d_backward_tiled_workspace(int dt, const Geom3 &g) {
    TileCfg c;
    if (!pick_cfg(MODE_BWD, (int)dtype_size(dt), g.T, g.C, g.Ho, g.Wo, g.H, g.W, g.sH, kKmaxBwd, &c)) return 0;
    return (size_t)g.C * g.N * c.row_tiles * 3 * sizeof(double);
}

int shift3d_backward_tiled(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt,
                           int sdt, const Geom3 &g, int normalize, double factor, void *workspace,
                           cudaStream_t s) {
    TiledArgs a{};
    a.src = og; a.dst = gin; a.xin = gshift ? x : nullptr; a.shift = shift;
    a.partial = gshift ? (double *)workspace : nullptr; a.sdt = sdt;
    a.N = g.N; a.Tn = g.T; a.C = g.C; a.Hs = g.Ho; a.Ws = g.Wo; a.Hd = g.H; a.Wd = g.W; a.S = g.sH;
    if (!pick_cfg(MODE_BWD, (int)dtype_size(dt), g.T, g.C, g.Ho, g.Wo, g.H, g.W, g.sH, kKmaxBwd, &a.cfg))
        return fail(RB_ERR_UNSUPPORTED, "tiled backward: no tile configuration");
    int rc = launch_dtype<MODE_BWD>(dt, a, s);
    if (rc || !gshift) return rc;
    return shift3d_finalize((const double *)workspace, g.N * a.cfg.row_tiles, gshift, dt, sdt, g.C, normalize,
                            factor, s);
}

// ---- 2D shift (stride 1 or 2) on the same kernel: groups of Tn images per CTA, no temporal taps ---------------------
int shift2d_strip_finalize(const double *partial, int parts, void *gshift, int sdt, int C, int normalize, cudaStream_t s);

static int tiled2d_frames(int dt, const Geom2 &g, TileCfg *cf, TileCfg *cb) {
    const int es = (int)dtype_size(dt);
    for (int tn = 8; tn >= 1; tn >>= 1)
        if (g.N % tn == 0 && pick_cfg(MODE_FWD, es, tn, g.C, g.H, g.W, g.Ho, g.Wo, g.sH, kKmaxFwd, cf) &&
            pick_cfg(MODE_BWD, es, tn, g.C, g.Ho, g.Wo, g.H, g.W, g.sH, kKmaxBwd, cb))
            return tn;
    return 0;
}

bool shift2d_tiled_supported(int dt, const Geom2 &g, int quantize) {
    if (quantize) return false;
    if (dt != RB_F32 && dt != RB_F16 && dt != RB_BF16) return false;
    if (g.pH != 0 || g.pW != 0 || g.sH != g.sW || (g.sH != 1 && g.sH != 2)) return false;
    if (g.Ho <= 0 || g.Wo <= 0) return false;
    TileCfg cf, cb;
    const int tn = tiled2d_frames(dt, g, &cf, &cb);
    if (!tn) return false;
    return (int64_t)(g.N / tn) * cf.groups * cf.row_tiles <= 0x7fffffffLL && (int64_t)(g.N / tn) * cb.groups * cb.row_tiles <= 0x7fffffffLL;
}

int shift2d_forward_tiled(const void *x, const void *shift, void *out, int dt, int sdt, const Geom2 &g, cudaStream_t s) {
    TiledArgs a{};
    TileCfg cb;
    a.src = x; a.dst = out; a.xin = nullptr; a.shift = shift; a.partial = nullptr; a.sdt = sdt;
    a.C = g.C; a.Hs = g.H; a.Ws = g.W; a.Hd = g.Ho; a.Wd = g.Wo; a.S = g.sH; a.mode2d = 1;
    a.Tn = tiled2d_frames(dt, g, &a.cfg, &cb);
    if (!a.Tn) return fail(RB_ERR_UNSUPPORTED, "2D tiled forward: no tile configuration");
    a.N = g.N / a.Tn;
    return launch_dtype<MODE_FWD>(dt, a, s);
}

size_t shift2d_backward_tiled_workspace(int dt, const Geom2 &g) {
    TileCfg cf, cb;
    const int tn = tiled2d_frames(dt, g, &cf, &cb);
    if (!tn) return 0;
    return (size_t)g.C * (g.N / tn) * cb.row_tiles * 3 * sizeof(double);
}

int shift2d_backward_tiled(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt, int sdt,
                           const Geom2 &g, int normalize, void *workspace, cudaStream_t s) {
    TiledArgs a{};
    TileCfg cf;
    a.src = og; a.dst = gin; a.xin = gshift ? x : nullptr; a.shift = shift;
    a.partial = gshift ? (double *)workspace : nullptr; a.sdt = sdt;
    a.C = g.C; a.Hs = g.Ho; a.Ws = g.Wo; a.Hd = g.H; a.Wd = g.W; a.S = g.sH; a.mode2d = 1;
    a.Tn = tiled2d_frames(dt, g, &cf, &a.cfg);
    if (!a.Tn) return fail(RB_ERR_UNSUPPORTED, "2D tiled backward: no tile configuration");
    a.N = g.N / a.Tn;
    int rc = launch_dtype<MODE_BWD>(dt, a, s);
    if (rc || !gshift) return rc;
    return shift2d_strip_finalize((const double *)workspace, a.N * a.cfg.row_tiles, gshift, sdt, g.C, normalize, s);
}

}  // namespace rb

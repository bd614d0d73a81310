// "Strip" kernels for the stride-1 3D learnable shift (47 of the 51 shift layers of RubiksNet-Large):
// second-generation sm_100a kernels, instruction-lean so that bf16 can approach the HBM roofline.
//
// Same staging as shift3d_tiled.cu (one 1-D TMA bulk copy per source frame into shared memory, completion on
// per-frame mbarriers; one CTA = clip n x CG channels x a band of TH rows, all T frames), different compute:
//   * a thread owns a vertical strip of R output rows x CW columns (CW = 1 for fp32, 2 for 16-bit types, so a
//     warp always writes 128 contiguous bytes per row) and walks the T frames with its R*CW partial results in
//     registers;
//   * the W-interpolated row  L[j] = q[j][c]*w0 + q[j][c+1]*w1  is computed once per source row and shared by
//     the two output rows that use it, so per output element the kernel issues ~2.25/CW shared loads and ~6 FP
//     ops instead of 4 loads + 8 ops (+ address arithmetic) per element;
//   * no per-tap predicates: out-of-range source ROWS are redirected to a zeroed row at the head of every
//     stage, out-of-range COLUMNS get weight 0 (their address is clamped onto initialised shared memory), both
//     precomputed once per thread.  [Consequence: a non-finite input next to the zero padding can turn
//     0*inf into NaN where the reference yields a finite value; finite inputs are unaffected.]
//   * backward (adjoint form, see shift3d_tiled.cu) regroups the shift-gradient sums by source frame u:
//        dT += B[u]  * (x[u-f] - x[u-f-1]),   dH += DH[u] * xm,   dW += DW[u] * xm,  xm = w0 x[u-f] + w1 x[u-f-1]
//     which needs only B (already needed for x_grad) and the previous x in registers.
// Reference semantics restated: /root/reference/cuda_src/rubiks3d_kernels.cu:15-205 (forward), :218-452
// (shift gradient incl. the exact-integer rule, taken by a per-thread slow path), :726-929 (input gradient,
// stride-1/pad-0 variant), :932-960 (normalisation, in k_shift3d_finalize).
#include "common.cuh"

namespace rb {

int shift3d_finalize(const double *partial, int parts, void *shift_grad, int dt, int sdt, int C, int normalize,
                     double factor, cudaStream_t s);

static constexpr int kSNT = 128;        // threads per CTA
static constexpr int kSMaxFrames = 16;  // frames resident in shared memory
static constexpr int kSHdr = 2048;      // [0,128) mbarriers  [128,192) per-frame misalignment  [512,2048) reduction
static constexpr int kSZero = 32;       // zeroed bytes at the head of every stage
static constexpr int kSSlack = 16;      // zeroed bytes after the staged data

enum { SMODE_FWD = 0, SMODE_BWD = 1 };

struct StripCfg {
    int CG, spc, ncg, nstr, TH, row_tiles, groups, R;
    int stage_bytes, smem_bytes;
};
struct StripArgs {
    const void *src;  // x (FWD) / out_grad (BWD)
    void *dst;        // out (FWD) / x_grad (BWD, may be null)
    const void *xin;  // BWD: x (null when the shift gradient is not wanted)
    const void *shift;
    double *partial;  // BWD: [C][N*row_tiles][3]
    int sdt;
    int N, Tn, C, H, W;
    int mode2d;       // 2D shift (cuda_src/rubiks2d_kernels.cu): shift is [2, C] (H, W); the Tn "frames" of a CTA are independent images
    StripCfg cfg;
};

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void s_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void s_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     s_u32(dst)),
                 "l"(src), "r"(bytes), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void s_wait(uint64_t *bar) {
    uint32_t ok;
    const uint32_t addr = s_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr)
            : "memory");
    } while (!ok);
}

template <typename T> struct StripTraits { static constexpr int CW = 2; };
template <> struct StripTraits<float> { static constexpr int CW = 1; };

template <typename T> __device__ __forceinline__ float s_tof(T v) { return (float)v; }
template <> __device__ __forceinline__ float s_tof<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float s_tof<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// store CW consecutive outputs (vectorised to one 32-bit store for 16-bit types when VEC)
template <typename T, bool VEC> __device__ __forceinline__ void s_store(T *p, const float *v, int ncols) {
    constexpr int CW = StripTraits<T>::CW;
    if (CW == 1) {
        p[0] = cvt<T, float>(v[0]);
    } else if (VEC) {
        T pair[2] = {cvt<T, float>(v[0]), cvt<T, float>(v[1])};
        *reinterpret_cast<uint32_t *>(p) = *reinterpret_cast<const uint32_t *>(pair);
    } else {
        p[0] = cvt<T, float>(v[0]);
        if (ncols > 1) p[1] = cvt<T, float>(v[1]);
    }
}
template <typename T, bool VEC> __device__ __forceinline__ void s_load(const T *p, float *v, int ncols) {
    constexpr int CW = StripTraits<T>::CW;
    if (CW == 1) {
        v[0] = s_tof(p[0]);
    } else if (VEC) {
        const uint32_t raw = __ldg(reinterpret_cast<const uint32_t *>(p));
        T pair[2];
        *reinterpret_cast<uint32_t *>(pair) = raw;
        v[0] = s_tof(pair[0]);
        v[1] = s_tof(pair[1]);
    } else {
        v[0] = s_tof(p[0]);
        v[1] = ncols > 1 ? s_tof(p[1]) : 0.f;
    }
}

// the same CW = 2 elements as one raw 32-bit word (element 0 in the low half), converted where they are used: the packed
// backward loop keeps x of three frames in flight (previous / current / prefetched) at one register per row each
template <typename T, bool VEC> __device__ __forceinline__ uint32_t s_load_raw(const T *p, int ncols) {
    if (sizeof(T) != 2) return 0u;  // fp32 takes the scalar loop (CW == 1)
    if (VEC) return __ldg(reinterpret_cast<const uint32_t *>(p));
    const uint32_t lo = *reinterpret_cast<const unsigned short *>(p);
    const uint32_t hi = ncols > 1 ? *reinterpret_cast<const unsigned short *>(p + 1) : 0u;
    return lo | (hi << 16);
}
template <typename T> __device__ __forceinline__ float2 s_raw_to_f2(uint32_t w);
template <> __device__ __forceinline__ float2 s_raw_to_f2<__nv_bfloat16>(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 s_raw_to_f2<__half>(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
template <> __device__ __forceinline__ float2 s_raw_to_f2<float>(uint32_t) { return make_float2(0.f, 0.f); }  // never used (CW == 1)

__device__ __forceinline__ float sc_a(int d, float r) { return d == 0 ? 1.f - r : (d == 1 ? r : 0.f); }
__device__ __forceinline__ float sc_b(int d, float r, bool integer) { return integer ? (d == 1 ? 1.f : 0.f) : sc_a(d, r); }
__device__ __forceinline__ float sc_sg(int d, bool integer) {
    if (d == 1) return -1.f;
    return integer ? (d == -1 ? 1.f : 0.f) : (d == 0 ? 1.f : 0.f);
}

template <typename T, int MODE, int R, bool VEC>
// (forcing 5 CTAs/SM on the backward -- 96 registers, 28 bytes of spills instead of 125 registers -- was measured SLOWER on every
// geometry, 0.631 -> 0.761 ms at 112x112 and 0.052 -> 0.058 ms at 14x14, gpurun_out/r02ae: the kernel is issue-bound, not
// residency-bound)
__global__ void __launch_bounds__(kSNT, MODE == SMODE_BWD ? 4 : 7) k_shift3d_strip(const StripArgs a) {
    pdl_sync();
    constexpr int ES = (int)sizeof(T);
    constexpr int CW = StripTraits<T>::CW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    int *mis_tab = reinterpret_cast<int *>(smem_raw + 128);
    float *red = reinterpret_cast<float *>(smem_raw + 512);
    unsigned char *stages = smem_raw + kSHdr;

    const int tid = threadIdx.x;
    const StripCfg &cf = a.cfg;
    const int CG = cf.CG, Tn = a.Tn, H = a.H, W = a.W, HW = H * W;
    int b = blockIdx.x;
    const int rtile = b % cf.row_tiles; b /= cf.row_tiles;
    const int grp = b % cf.groups;
    const int n = b / cf.groups;
    const int c0 = grp * CG;
    const int cg_eff = min(CG, a.C - c0);
    const int hd0 = rtile * cf.TH;
    const int th = min(cf.TH, H - hd0);

    // ---- slot -> (channel, strip, column group) ----------------------------------------------------
    const int cl_raw = tid / cf.spc;
    const int rem = tid - cl_raw * cf.spc;
    const int strip = rem / cf.ncg;
    const int cgi = rem - strip * cf.ncg;
    const int hl0 = strip * R;
    const int nrows = min(R, th - hl0);
    const int wd0 = cgi * CW;
    const bool active = cl_raw < cg_eff && nrows > 0;
    const int cl = min(cl_raw, cg_eff - 1);
    const int c = c0 + cl;
    const int ncols = min(CW, W - wd0);

    float sT, sH, sW;
    if (a.mode2d) {
        sT = 0.f;
        sH = ld_param<float>(a.shift, a.sdt, c);
        sW = ld_param<float>(a.shift, a.sdt, a.C + c);
        if (MODE == SMODE_BWD) {
            // the 2D shift gradient treats |r| < 1e-7 as an exact integer shift (rubiks2d_kernels.cu:189): snap the shift
            // itself, which moves the input gradient by at most 1e-7 of a tap
            const float fh = floorf(sH), fw = floorf(sW);
            if (fabsf(sH - fh) < 1e-7f) sH = fh;
            if (fabsf(sW - fw) < 1e-7f) sW = fw;
        }
    } else {
        sT = ld_param<float>(a.shift, a.sdt, c);
        sH = ld_param<float>(a.shift, a.sdt, a.C + c);
        sW = ld_param<float>(a.shift, a.sdt, 2 * a.C + c);
    }
    if (MODE == SMODE_BWD) { sT = -sT; sH = -sH; sW = -sW; }
    const int fT = floor3d(sT), fH = floor3d(sH), fW = floor3d(sW);
    const float rT = sT - fT, rH = sH - fH, rW = sW - fW;
    // 2D: there is no temporal component (sT = 0 selects frame t with weight 1 and must not trigger the integer rule)
    const bool intT = !a.mode2d && rT == 0.f, intH = rH == 0.f, intW = rW == 0.f;
    const bool slow = (MODE == SMODE_BWD) && (intT || intH || intW);

    // ---- staged source rows (CTA-uniform: CG == 1 implies every thread reads channel c0's shift) ----
    int rlo = 0, rhi = H - 1;
    if (CG == 1) {
        rlo = max(hd0 + fH - ((MODE == SMODE_BWD && intH) ? 1 : 0), 0);
        rhi = min(hd0 + th - 1 + fH + 1, H - 1);
    }
    const int count = (rhi >= rlo) ? (cg_eff - 1) * HW + (rhi - rlo + 1) * W : 0;
    const bool any_data = count > 0;

    const T *src = reinterpret_cast<const T *>(a.src);
    auto frame_src = [&](int t) -> const T * {
        return src + ((int64_t)(n * Tn + t) * a.C + c0) * HW + (int64_t)rlo * W;
    };
    if (tid == 0) {
        for (int t = 0; t < Tn; ++t) s_mbar_init(&bars[t], 1);
        s_fence_init();
    }
    __syncthreads();
    {
        const int job = tid >> 4, t = tid & 15;  // job 0: bulk copy, 1: zero head + head elements, 2: tail + slack
        if (t < Tn && job < 3) {
            const T *fs = frame_src(t);
            const uintptr_t sb = reinterpret_cast<uintptr_t>(fs);
            const uintptr_t eb = sb + (uintptr_t)count * ES;
            const uintptr_t ab = (sb + 15) & ~(uintptr_t)15, ae = eb & ~(uintptr_t)15;
            const int mis = any_data ? (int)(sb & 15) : 0;
            unsigned char *stage = stages + (size_t)t * cf.stage_bytes;
            T *sdata = reinterpret_cast<T *>(stage + kSZero + mis);
            const bool body = any_data && ae > ab;
            const int head = !any_data ? 0 : (body ? (int)((ab - sb) / ES) : count);
            const int tail_start = body ? (int)((ae - sb) / ES) : count;
            if (job == 0) {
                mis_tab[t] = mis;
                if (any_data) {
                    const uint32_t bytes = body ? (uint32_t)(ae - ab) : 0u;
                    s_expect_tx(&bars[t], bytes);
                    if (body) s_bulk_g2s(stage + kSZero + (ab - (sb & ~(uintptr_t)15)), reinterpret_cast<const void *>(ab),
                                         bytes, &bars[t]);
                }
            } else if (job == 1) {
                T *z = reinterpret_cast<T *>(stage);
                for (int i = 0; i < (kSZero + mis) / ES; ++i) z[i] = cvt<T, float>(0.f);
                for (int i = 0; i < head; ++i) sdata[i] = fs[i];
            } else {
                for (int i = tail_start; i < count; ++i) sdata[i] = fs[i];
                for (int i = 0; i < kSSlack / ES; ++i) sdata[count + i] = cvt<T, float>(0.f);
            }
        }
    }

    // ---- per-thread tap geometry -----------------------------------------------------------------------
    const float wH0 = 1.f - rH, wH1 = rH, wW0 = 1.f - rW, wW1 = rW, wT0 = 1.f - rT, wT1 = rT;
    const int cb = min(max(wd0 + fW, -CW), W);
    float vf[CW + 1];
#pragma unroll
    for (int i = 0; i <= CW; ++i) {
        const int col = wd0 + fW + i;
        vf[i] = (active && col >= 0 && col < W) ? 1.f : 0.f;
    }
    float wa[CW][2];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        wa[i][0] = wW0 * vf[i];
        wa[i][1] = wW1 * vf[i + 1];
    }
    int rowoff[R + 1];  // byte offset of (source row j, column cb) relative to the frame pointer; 0 = zero row
#pragma unroll
    for (int j = 0; j <= R; ++j) {
        const int hs = hd0 + hl0 + j + fH;
        const bool v = active && any_data && j <= nrows && hs >= rlo && hs <= rhi;
        rowoff[j] = v ? kSZero + (cl * HW + (hs - rlo) * W + cb) * ES : 0;
    }
    __syncthreads();  // zero rows, head/tail elements and mis_tab visible to everyone

    T *dst = reinterpret_cast<T *>(a.dst);
    const T *xin = reinterpret_cast<const T *>(a.xin);
    const bool want_dst = dst != nullptr;
    const bool want_grad = (MODE == SMODE_BWD) && xin != nullptr;
    const int64_t dst_fs = (int64_t)a.C * HW;
    const int64_t dbase = ((int64_t)n * Tn * a.C + c) * HW + (int64_t)(hd0 + hl0) * W + wd0;
    float accT = 0.f, accH = 0.f, accW = 0.f;

    if (any_data && (CG > 1 || slow))
        for (int t = 0; t < Tn; ++t) s_wait(&bars[t]);

    if (!slow && CW == 2) {
        // 16-bit types: a thread's two columns go through every formula together as PACKED fp32 pairs (FFMA2 / FMUL2 on
        // sm_100).  ncu on the scalar loop below (profiles/r02z_ncu_shift_kernels.txt): 53-56 warp instructions per element in
        // the backward pass at IPC 1.9-2.2 -- the three-register FFMA / FMUL issue every other cycle per scheduler, so
        // the fma pipe, not memory, bounds the kernel; the packed forms do two lanes per issue slot.  Same formulas, same
        // association (mul then fma); the shift-gradient sums are kept per column and added at the end.
        const float2 wa0 = make_float2(wa[0][0], wa[CW - 1][0]), wa1 = make_float2(wa[0][1], wa[CW - 1][1]);
        const float2 vf0 = make_float2(vf[0], vf[1]), nvf1 = make_float2(-vf[1], -vf[CW]);
        const float2 wH0v = make_float2(wH0, wH0), wH1v = make_float2(wH1, wH1);
        const float2 wT0v = make_float2(wT0, wT0), wT1v = make_float2(wT1, wT1);
        const float2 neg1 = make_float2(-1.f, -1.f), zero2 = make_float2(0.f, 0.f);
        float2 pB[R];
        // x of frame step - 1 / step / step + 1 as raw words: the frame needed at step s + 1 is fetched during step s, so its
        // latency hides behind a whole step (ncu: 15-30 % of the stall samples sat on the first use of a same-step load)
        uint32_t xwp[R], xwc[R], xwn[R];
        const bool fetch = (MODE == SMODE_BWD) && want_grad && active;
        const T *xrow = fetch ? xin + dbase : nullptr;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            pB[k] = zero2;
            xwp[k] = 0u;
            xwc[k] = (fetch && k < nrows && Tn > 0) ? s_load_raw<T, VEC>(xrow + k * W, ncols) : 0u;
            xwn[k] = 0u;
        }
        float2 aT = zero2, aH = zero2, aW = zero2;
        const bool mode2d = a.mode2d != 0;
        const uint32_t rmask = active ? ((1u << nrows) - 1u) : 0u;  // row slots of this thread that exist
        T *dprev = want_dst ? dst + dbase : nullptr;                 // destination frame step - 1, row 0 of this thread
        for (int step = 0; step <= Tn; ++step) {
            const int ts = step + fT;
            const bool have = any_data && ts >= 0 && ts < Tn;
            if (MODE == SMODE_BWD) {
                xrow += dst_fs;  // frame step + 1
#pragma unroll
                for (int k = 0; k < R; ++k) xwn[k] = (fetch && step + 1 < Tn && k < nrows) ? s_load_raw<T, VEC>(xrow + k * W, ncols) : 0u;
            }
            const bool emit = step >= 1 && want_dst;
            auto consume = [&](int k, const float2 Bk, const float2 DHk, const float2 DWk) {
                // the value is computed for every row slot, only the store is predicated (a branch around the whole body cost a
                // BSSY / BRA / BSYNC triple and a three-term 64-bit address per row)
                const float2 v2 = mode2d ? pB[k] : __ffma2_rn(wT1v, Bk, __fmul2_rn(wT0v, pB[k]));  // 2D: frames are independent images
                if (emit && ((rmask >> k) & 1u)) {
                    float v[CW];
                    v[0] = v2.x;
                    v[CW - 1] = v2.y;
                    s_store<T, VEC>(dprev + k * W, v, ncols);
                }
                if (MODE == SMODE_BWD) {
                    const float2 xn = s_raw_to_f2<T>(xwc[k]), xq = s_raw_to_f2<T>(xwp[k]);
                    const float2 xm = mode2d ? xn : __ffma2_rn(wT1v, xq, __fmul2_rn(wT0v, xn));
                    const float2 xd = __ffma2_rn(xq, neg1, xn);
                    aT = __ffma2_rn(Bk, xd, aT);
                    aH = __ffma2_rn(DHk, xm, aH);
                    aW = __ffma2_rn(DWk, xm, aW);
                }
                pB[k] = Bk;
            };
            if (have) {
                if (CG == 1) s_wait(&bars[ts]);
                const unsigned char *fp = stages + (size_t)ts * cf.stage_bytes + mis_tab[ts];
                float2 Lp = zero2, Ep = zero2;
#pragma unroll
                for (int j = 0; j <= R; ++j) {
                    const T *rp = reinterpret_cast<const T *>(fp + rowoff[j]);
                    const float q0 = s_tof(rp[0]), q1 = s_tof(rp[1]), q2 = s_tof(rp[CW]);
                    const float2 qa = make_float2(q0, q1), qb = make_float2(q1, q2);
                    const float2 L = __ffma2_rn(qb, wa1, __fmul2_rn(qa, wa0));
                    const float2 E = (MODE == SMODE_BWD) ? __ffma2_rn(qb, nvf1, __fmul2_rn(qa, vf0)) : zero2;
                    if (j >= 1) {
                        const float2 Bk = __ffma2_rn(wH1v, L, __fmul2_rn(wH0v, Lp));
                        const float2 DHk = __ffma2_rn(L, neg1, Lp);
                        const float2 DWk = __ffma2_rn(wH1v, E, __fmul2_rn(wH0v, Ep));
                        consume(j - 1, Bk, DHk, DWk);
                    }
                    Lp = L;
                    Ep = E;
                }
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) consume(k, zero2, zero2, zero2);
            }
            if (MODE == SMODE_BWD) {
#pragma unroll
                for (int k = 0; k < R; ++k) { xwp[k] = xwc[k]; xwc[k] = xwn[k]; }
            }
            if (emit) dprev += dst_fs;
        }
        accT = aT.x + aT.y;
        accH = aH.x + aH.y;
        accW = aW.x + aW.y;
    } else if (!slow) {
        float pB[R][CW], xp[R][CW];
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int i = 0; i < CW; ++i) { pB[k][i] = 0.f; xp[k][i] = 0.f; }
        const bool mode2d = a.mode2d != 0;
        const uint32_t rmask = active ? ((1u << nrows) - 1u) : 0u;  // row slots of this thread that exist
        T *dprev = want_dst ? dst + dbase : nullptr;                 // destination frame step - 1, row 0 of this thread
        for (int step = 0; step <= Tn; ++step) {
            const int ts = step + fT;
            const bool have = any_data && ts >= 0 && ts < Tn;
            float xn[R][CW];
            if (MODE == SMODE_BWD) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
#pragma unroll
                    for (int i = 0; i < CW; ++i) xn[k][i] = 0.f;
                    if (want_grad && step < Tn && active && k < nrows)
                        s_load<T, VEC>(xin + dbase + step * dst_fs + k * W, xn[k], ncols);
                }
            }
            const bool emit = step >= 1 && want_dst;
            // consume(k, B, DH, DW): finish output row k of destination frame step-1 and accumulate the shift gradient (the value
            // is computed for every row slot, only the store is predicated -- as in the packed loop above)
            auto consume = [&](int k, const float *Bk, const float *DHk, const float *DWk) {
                float v[CW];
#pragma unroll
                for (int i = 0; i < CW; ++i) v[i] = mode2d ? pB[k][i] : wT0 * pB[k][i] + wT1 * Bk[i];  // 2D: frames are independent images
                if (emit && ((rmask >> k) & 1u)) s_store<T, VEC>(dprev + k * W, v, ncols);
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                    if (MODE == SMODE_BWD) {
                        const float xm = mode2d ? xn[k][i] : wT0 * xn[k][i] + wT1 * xp[k][i];
                        const float xd = xn[k][i] - xp[k][i];
                        accT += Bk[i] * xd;
                        accH += DHk[i] * xm;
                        accW += DWk[i] * xm;
                        xp[k][i] = xn[k][i];
                    }
                    pB[k][i] = Bk[i];
                }
            };
            if (have) {
                if (CG == 1) s_wait(&bars[ts]);
                const unsigned char *fp = stages + (size_t)ts * cf.stage_bytes + mis_tab[ts];
                float Lp[CW], Ep[CW];
#pragma unroll
                for (int j = 0; j <= R; ++j) {
                    const T *rp = reinterpret_cast<const T *>(fp + rowoff[j]);
                    float q[CW + 1];
#pragma unroll
                    for (int i = 0; i <= CW; ++i) q[i] = s_tof(rp[i]);
                    float L[CW], E[CW];
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        L[i] = q[i] * wa[i][0] + q[i + 1] * wa[i][1];
                        E[i] = (MODE == SMODE_BWD) ? q[i] * vf[i] - q[i + 1] * vf[i + 1] : 0.f;
                    }
                    if (j >= 1) {
                        float Bk[CW], DHk[CW], DWk[CW];
#pragma unroll
                        for (int i = 0; i < CW; ++i) {
                            Bk[i] = wH0 * Lp[i] + wH1 * L[i];
                            DHk[i] = Lp[i] - L[i];
                            DWk[i] = wH0 * Ep[i] + wH1 * E[i];
                        }
                        consume(j - 1, Bk, DHk, DWk);
                    }
#pragma unroll
                    for (int i = 0; i < CW; ++i) { Lp[i] = L[i]; Ep[i] = E[i]; }
                }
            } else {
                float zero[CW];
#pragma unroll
                for (int i = 0; i < CW; ++i) zero[i] = 0.f;
#pragma unroll
                for (int k = 0; k < R; ++k) consume(k, zero, zero, zero);
            }
            if (emit) dprev += dst_fs;
        }
    } else if (active) {
        // ---- exact-integer shift component: tap-by-tap evaluation of the reference's rule ----------------
        const bool zero_shift = (sT == 0.f && sH == 0.f && sW == 0.f);  // :561-576
        auto tap = [&](int ts, int hs, int ws) -> float {
            const bool v = any_data && ts >= 0 && ts < Tn && hs >= rlo && hs <= rhi && ws >= 0 && ws < W;
            if (!v) return 0.f;
            const unsigned char *fp = stages + (size_t)ts * cf.stage_bytes + mis_tab[ts] + kSZero;
            return s_tof(reinterpret_cast<const T *>(fp)[cl * HW + (hs - rlo) * W + ws]);
        };
#pragma unroll 1
        for (int k = 0; k < nrows; ++k) {
#pragma unroll 1
            for (int i = 0; i < ncols; ++i) {
                const int hd = hd0 + hl0 + k, wd = wd0 + i;
#pragma unroll 1
                for (int td = 0; td < Tn; ++td) {
                    const int t0 = td + fT, h0 = hd + fH, w0 = wd + fW;
                    const int64_t di = dbase + td * dst_fs + k * W + i;
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
                        dst[di] = cvt<T, float>(v);
                    }
                    if (want_grad && a.mode2d) {
                        // 2D rule (rubiks2d_kernels.cu:189-253) in adjoint form: regular axis = difference of the two
                        // taps; exact-integer axis = 0.5 * central difference (taps -1 and +1); the other axis always
                        // interpolates with (1 - r, r) -- an integer shift there does not move the taps (unlike 3D)
                        float gH = 0.f, gW = 0.f;
#pragma unroll 1
                        for (int dy = -1; dy <= 1; ++dy) {
                            const float bh = sc_a(dy, rH);
                            const float sh = intH ? (dy == -1 ? 0.5f : (dy == 1 ? -0.5f : 0.f)) : (dy == 0 ? 1.f : (dy == 1 ? -1.f : 0.f));
#pragma unroll 1
                            for (int dx = -1; dx <= 1; ++dx) {
                                const float bw = sc_a(dx, rW);
                                const float sw = intW ? (dx == -1 ? 0.5f : (dx == 1 ? -0.5f : 0.f)) : (dx == 0 ? 1.f : (dx == 1 ? -1.f : 0.f));
                                const float cH = sh * bw, cW2 = bh * sw;
                                if (cH == 0.f && cW2 == 0.f) continue;
                                const float qq = tap(t0, h0 + dy, w0 + dx);
                                gH += cH * qq;
                                gW += cW2 * qq;
                            }
                        }
                        const float xv = s_tof(xin[di]);
                        accH += xv * gH;
                        accW += xv * gW;
                    } else if (want_grad) {
                        float gT = 0.f, gH = 0.f, gW = 0.f;
#pragma unroll 1
                        for (int dt = -1; dt <= 1; ++dt) {
                            const float bt = sc_b(dt, rT, intT), st = sc_sg(dt, intT);
#pragma unroll 1
                            for (int dy = -1; dy <= 1; ++dy) {
                                const float bh = sc_b(dy, rH, intH), sh = sc_sg(dy, intH);
#pragma unroll 1
                                for (int dx = -1; dx <= 1; ++dx) {
                                    const float bw = sc_b(dx, rW, intW), sw = sc_sg(dx, intW);
                                    const float cT = st * bh * bw, cH = bt * sh * bw, cW2 = bt * bh * sw;
                                    if (cT == 0.f && cH == 0.f && cW2 == 0.f) continue;
                                    const float qq = tap(t0 + dt, h0 + dy, w0 + dx);
                                    gT += cT * qq;
                                    gH += cH * qq;
                                    gW += cW2 * qq;
                                }
                            }
                        }
                        const float xv = s_tof(xin[di]);
                        accT += xv * gT;
                        accH += xv * gH;
                        accW += xv * gW;
                    }
                }
            }
        }
    }

    if (MODE == SMODE_BWD && a.partial != nullptr) {
        red[tid * 3 + 0] = active ? accT : 0.f;
        red[tid * 3 + 1] = active ? accH : 0.f;
        red[tid * 3 + 2] = active ? accW : 0.f;
        __syncthreads();
        if (tid < cg_eff * 3) {
            const int l = tid / 3, ax = tid - l * 3;
            double s = 0;
            for (int i = l * cf.spc; i < (l + 1) * cf.spc; ++i) s += (double)red[i * 3 + ax];
            const int parts = a.N * cf.row_tiles;
            a.partial[((int64_t)(c0 + l) * parts + (n * cf.row_tiles + rtile)) * 3 + ax] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
static constexpr int kSSmemTarget = 40 * 1024;
static constexpr int kSSmemLimit = 227 * 1024;

static int s_stage_bytes(int count_elems, int es) {
    int bytes = kSZero + 16 + count_elems * es + kSSlack;
    return (bytes + 15) & ~15;
}

static bool pick_strip_cfg(int es, int Tn, int C, int H, int W, StripCfg *out) {
    if (Tn <= 0 || Tn > kSMaxFrames || H <= 0 || W <= 0) return false;
    const int CW = es == 4 ? 1 : 2;
    StripCfg c;
    c.R = (H % 8 == 0) ? 8 : ((H % 7 == 0) ? 7 : 8);
    c.ncg = cdiv(W, CW);
    if (c.ncg > kSNT) return false;
    const int nstr_plane = cdiv(H, c.R);
    const int spc_full = nstr_plane * c.ncg;
    if (spc_full <= kSNT) {  // whole planes, several channels per CTA
        int CG = kSNT / spc_full;
        if (CG > C) CG = C;
        while (CG > 1 && kSHdr + Tn * s_stage_bytes(CG * H * W, es) > kSSmemTarget) --CG;
        c.CG = CG; c.spc = spc_full; c.nstr = nstr_plane; c.TH = H; c.row_tiles = 1;
        c.stage_bytes = s_stage_bytes(CG * H * W, es);
    } else {  // one channel, band of rows
        int nstr = kSNT / c.ncg;
        if (nstr < 1) return false;
        int TH = nstr * c.R;
        int row_tiles = cdiv(H, TH);
        nstr = cdiv(cdiv(H, row_tiles), c.R);  // balance the bands
        TH = nstr * c.R;
        c.CG = 1; c.spc = nstr * c.ncg; c.nstr = nstr; c.TH = TH; c.row_tiles = cdiv(H, TH);
        int rows = TH + 3;
        if (rows > H) rows = H;
        c.stage_bytes = s_stage_bytes(rows * W, es);
    }
    c.smem_bytes = kSHdr + Tn * c.stage_bytes;
    if (c.smem_bytes > kSSmemLimit) return false;
    c.groups = cdiv(C, c.CG);
    *out = c;
    return true;
}

bool shift3d_strip_supported(int dt, const Geom3 &g, int quantize) {
    if (quantize) return false;
    if (dt != RB_F32 && dt != RB_F16 && dt != RB_BF16) return false;
    if (g.sT != 1 || g.sH != 1 || g.sW != 1 || g.pT != 0 || g.pH != 0 || g.pW != 0) return false;
    StripCfg c;
    if (!pick_strip_cfg((int)dtype_size(dt), g.T, g.C, g.H, g.W, &c)) return false;
    return (int64_t)g.N * c.groups * c.row_tiles <= 0x7fffffffLL;
}

template <typename T, int MODE, int R, bool VEC> static int strip_launch(const StripArgs &a, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_shift3d_strip<T, MODE, R, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSSmemLimit);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    const unsigned blocks = (unsigned)((int64_t)a.N * a.cfg.groups * a.cfg.row_tiles);
    launch_kernel(k_shift3d_strip<T, MODE, R, VEC>, dim3(blocks), dim3(kSNT), a.cfg.smem_bytes, s, a);
    return launched(MODE == SMODE_FWD ? "k_shift3d_strip<fwd>" : "k_shift3d_strip<bwd>");
}

template <typename T, int MODE> static int strip_dispatch(const StripArgs &a, cudaStream_t s) {
    const bool vec = StripTraits<T>::CW == 1 || (a.W % 2 == 0);
    if (a.cfg.R == 8) return vec ? strip_launch<T, MODE, 8, true>(a, s) : strip_launch<T, MODE, 8, false>(a, s);
    return vec ? strip_launch<T, MODE, 7, true>(a, s) : strip_launch<T, MODE, 7, false>(a, s);
}

template <int MODE> static int strip_dtype(int dt, const StripArgs &a, cudaStream_t s) {
    switch (dt) {
        case RB_F32: return strip_dispatch<float, MODE>(a, s);
        case RB_F16: return strip_dispatch<__half, MODE>(a, s);
        case RB_BF16: return strip_dispatch<__nv_bfloat16, MODE>(a, s);
    }
    return fail(RB_ERR_UNSUPPORTED, "strip kernel: dtype %d", dt);
}

int shift3d_forward_strip(const void *x, const void *shift, void *out, int dt, int sdt, const Geom3 &g, cudaStream_t s) {
    StripArgs a{};
    a.src = x; a.dst = out; a.xin = nullptr; a.shift = shift; a.partial = nullptr; a.sdt = sdt;
    a.N = g.N; a.Tn = g.T; a.C = g.C; a.H = g.H; a.W = g.W;
    if (!pick_strip_cfg((int)dtype_size(dt), g.T, g.C, g.H, g.W, &a.cfg))
        return fail(RB_ERR_UNSUPPORTED, "strip forward: no configuration");
    return strip_dtype<SMODE_FWD>(dt, a, s);
}

size_t shift3d_backward_strip_workspace(int dt, const Geom3 &g) {
    StripCfg c;
    if (!pick_strip_cfg((int)dtype_size(dt), g.T, g.C, g.H, g.W, &c)) return 0;
    return (size_t)g.C * g.N * c.row_tiles * 3 * sizeof(double);
}

int shift3d_backward_strip(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt, int sdt,
                           const Geom3 &g, int normalize, double factor, void *workspace, cudaStream_t s) {
    StripArgs a{};
    a.src = og; a.dst = gin; a.xin = gshift ? x : nullptr; a.shift = shift;
    a.partial = gshift ? (double *)workspace : nullptr; a.sdt = sdt;
    a.N = g.N; a.Tn = g.T; a.C = g.C; a.H = g.H; a.W = g.W;
    if (!pick_strip_cfg((int)dtype_size(dt), g.T, g.C, g.H, g.W, &a.cfg))
        return fail(RB_ERR_UNSUPPORTED, "strip backward: no configuration");
    int rc = strip_dtype<SMODE_BWD>(dt, a, s);
    if (rc || !gshift) return rc;
    return shift3d_finalize((const double *)workspace, g.N * a.cfg.row_tiles, gshift, dt, sdt, g.C, normalize, factor, s);
}

// ---- 2D shift on [N, C, H, W] through the same kernels: every image is a one-frame clip, the shift has no T row ---------
__global__ void k_shift2d_strip_finalize(const double *__restrict__ partial, int parts, void *shift_grad, int sdt, int C,
                                         int normalize) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0;
    const double *p = partial + (int64_t)c * parts * 3;
    for (int i = lane; i < parts; i += 32) {
        s0 += p[i * 3 + 1];
        s1 += p[i * 3 + 2];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane != 0) return;
    float gh = (float)s0, gw = (float)s1;
    if (normalize) {  // rubiks2d_kernels.cu:386-396
        const float mag = sqrtf(gh * gh + gw * gw);
        if (mag > 0) {
            gh = gh / mag;
            gw = gw / mag;
        }
    }
    st_param<float>(shift_grad, sdt, c, gh);
    st_param<float>(shift_grad, sdt, C + c, gw);
}

// 2D images are processed in groups of Tn "frames" per CTA (independent images: no temporal taps), which gives a CTA the
// same amount of work per staging round trip as a 3D clip; Tn = the largest of 8, 4, 2, 1 that divides the batch
static int strip2d_frames(int dt, const Geom2 &g, StripCfg *c) {
    for (int tn = 8; tn >= 1; tn >>= 1)
        if (g.N % tn == 0 && pick_strip_cfg((int)dtype_size(dt), tn, g.C, g.H, g.W, c)) return tn;
    return 0;
}

int shift2d_strip_finalize(const double *partial, int parts, void *gshift, int sdt, int C, int normalize, cudaStream_t s);

bool shift2d_strip_supported(int dt, const Geom2 &g, int quantize) {
    if (quantize) return false;
    if (dt != RB_F32 && dt != RB_F16 && dt != RB_BF16) return false;
    if (g.sH != 1 || g.sW != 1 || g.pH != 0 || g.pW != 0) return false;
    StripCfg c;
    const int tn = strip2d_frames(dt, g, &c);
    if (!tn) return false;
    return (int64_t)(g.N / tn) * c.groups * c.row_tiles <= 0x7fffffffLL;
}

int shift2d_forward_strip(const void *x, const void *shift, void *out, int dt, int sdt, const Geom2 &g, cudaStream_t s) {
    StripArgs a{};
    a.src = x; a.dst = out; a.xin = nullptr; a.shift = shift; a.partial = nullptr; a.sdt = sdt;
    a.C = g.C; a.H = g.H; a.W = g.W; a.mode2d = 1;
    a.Tn = strip2d_frames(dt, g, &a.cfg);
    if (!a.Tn) return fail(RB_ERR_UNSUPPORTED, "2D strip forward: no configuration");
    a.N = g.N / a.Tn;
    return strip_dtype<SMODE_FWD>(dt, a, s);
}

size_t shift2d_backward_strip_workspace(int dt, const Geom2 &g) {
    StripCfg c;
    const int tn = strip2d_frames(dt, g, &c);
    if (!tn) return 0;
    return (size_t)g.C * (g.N / tn) * c.row_tiles * 3 * sizeof(double);
}

int shift2d_backward_strip(const void *x, const void *shift, const void *og, void *gin, void *gshift, int dt, int sdt,
                           const Geom2 &g, int normalize, void *workspace, cudaStream_t s) {
    StripArgs a{};
    a.src = og; a.dst = gin; a.xin = gshift ? x : nullptr; a.shift = shift;
    a.partial = gshift ? (double *)workspace : nullptr; a.sdt = sdt;
    a.C = g.C; a.H = g.H; a.W = g.W; a.mode2d = 1;
    a.Tn = strip2d_frames(dt, g, &a.cfg);
    if (!a.Tn) return fail(RB_ERR_UNSUPPORTED, "2D strip backward: no configuration");
    a.N = g.N / a.Tn;
    int rc = strip_dtype<SMODE_BWD>(dt, a, s);
    if (rc || !gshift) return rc;
    return shift2d_strip_finalize((const double *)workspace, a.N * a.cfg.row_tiles, gshift, sdt, g.C, normalize, s);
}

// partial [C][parts][3] (temporal slot unused) -> shift_grad [2, C] incl. the 2D normalisation; also used by the tiled kernel
int shift2d_strip_finalize(const double *partial, int parts, void *gshift, int sdt, int C, int normalize, cudaStream_t s) {
    const int warps = 4;
    launch_kernel(k_shift2d_strip_finalize, dim3(cdiv(C, warps)), dim3(warps * 32), 0, s, partial, parts, gshift, sdt, C, normalize);
    return launched("k_shift2d_strip_finalize");
}

}  // namespace rb

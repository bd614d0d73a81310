// Pointwise (1x1) convolution of RubiksShiftBlock as a tcgen05 tensor-core GEMM on NCHW bf16 activations, with the
// operation in front of the convolution folded into the kernel's A-operand producer:
//
//     out[i, n, p] = sum_k  W[n, k] * A(i, k, p)   (+ residual[i, n, p])          i = image (clip*T + t), p = pixel
//
//     PROD_PLAIN   A = x                                   conv2 / shortcut / the dgrad of any 1x1 conv
//     PROD_BNRELU  A = relu(x * scale[k] + bias[k])          bn1 -> relu -> conv2          (backbone.py:123-128)
//     PROD_SHIFT3D A = RubiksShift3D(x)[i, k, p]             as3 -> conv3 -> += shortcut   (backbone.py:129-135)
//                  trilinear gather of cuda_src/rubiks3d_kernels.cu:54-203 (stride 1, pad 0), rounded to bf16 once,
//                  i.e. exactly what the stand-alone shift kernel would have stored: the shifted tensor never
//                  exists in HBM.
//
// One launch, persistent CTAs (one per SM), 13 warps with fixed roles:
//     warp 0       allocates TMEM; one elected thread issues tcgen05.mma (M=128 pixels x N<=256 channels x K=16)
//     warps 1-4    epilogue: tcgen05.ld the fp32 accumulator (thread = pixel row), add the residual, store bf16
//     warps 5-12   producers: build the A tile [128 pixels x 64 channels] in shared memory in the UMMA canonical
//                  MN-major (pixel-contiguous) no-swizzle layout, so global reads stay coalesced along pixels;
//                  generic-proxy stores -> fence.proxy.async -> mbarrier, ring of up to 6 stages
// The weight block B [Ncta x Kpad] stays resident in shared memory (K-major canonical layout) for the CTA's
// lifetime.  GEMM view per tile: D[128 px, Ncta] = A[128 px, K] * B[Ncta, K]^T, accumulators in TMEM
// (double-buffered when 2*Ncta <= 512 columns so the epilogue of tile i overlaps the MMAs of tile i+1).
#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int kProdWarp0 = 5;
constexpr int kNumProdWarps = 8;
constexpr int kThreads = (kProdWarp0 + kNumProdWarps) * 32;  // 416
constexpr int kTileM = 128;
constexpr int kStageK = 64;
constexpr int kStageBytes = kStageK * kTileM * 2;  // 16 KiB
constexpr int kMaxStages = 6;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHdrBytes = 256;

enum { PROD_PLAIN = 0, PROD_BNRELU = 1, PROD_SHIFT3D = 2 };

// what the shift producer reads (shared by the forward and the weight-gradient kernels)
struct ShiftSrc {
    const __nv_bfloat16 *x;  // [clips, T, K, H, W]
    const void *shift;       // [3, K]
    int shift_dt, T, H, W, HW, K;
};

struct PwArgs {
    const __nv_bfloat16 *x;    // [NI, K, HW]
    const void *w;             // [N, K] (w_trans: [K, N]), bf16 or fp32 (w_dt)
    int w_dt, w_trans;
    const __nv_bfloat16 *res;  // [NI, N, HW] or null
    __nv_bfloat16 *out;        // [NI, N, HW]
    const float *a_sb;              // PROD_BNRELU: per input channel (scale, bias) pairs [K, 2]
    const void *shift;              // PROD_SHIFT3D: [3, K]
    int shift_dt;
    int T, H, W;                    // PROD_SHIFT3D: frames per clip, map size (HW = H*W)
    int NI, K, N, HW;
    int Kpad, Ncta, n_sub, sub_n, acc_stages, stages, tmem_cols;
    int tiles_per_img, total_tiles, k_stages;
    uint32_t off_b, off_a, off_sb;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // physical strides: 8-channel group / 8-row group
    uint32_t ad_lbo, ad_sbo, bd_lbo, bd_sbo;  // the same, as written into the UMMA descriptors
};

struct Hdr {
    uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(Hdr) <= kHdrBytes, "header");

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

__device__ __forceinline__ void bn_relu_unit(uint32_t (&r)[4], float sc, float bi) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float lo = fmaxf(fmaf(bf16_lo(r[h]), sc, bi), 0.f);
        const float hi = fmaxf(fmaf(bf16_hi(r[h]), sc, bi), 0.f);
        r[h] = pack_bf16x2(lo, hi);
    }
}

// RubiksShift3D forward for 8 consecutive output pixels p..p+7 of (image img = clip*T + t, channel k), stride 1 /
// pad 0: out = (1-rT)((1-rH)(q111(1-rW)+q112 rW) + rH(q121(1-rW)+q122 rW)) + rT(...), zero outside the tensor
// (cuda_src/rubiks3d_kernels.cu:54-74,97-203; same association order).  Walks the pixels left to right and reuses
// the right-hand taps of pixel e as the left-hand taps of pixel e+1 inside an image row.
__device__ __forceinline__ void shift3d_unit(const ShiftSrc &a, int img, int k, int p, uint32_t (&r)[4]) {
    const int T = a.T, H = a.H, W = a.W, HW = a.HW;
    const float sT = ld_param<float>(a.shift, a.shift_dt, k), sH = ld_param<float>(a.shift, a.shift_dt, a.K + k),
                sW = ld_param<float>(a.shift, a.shift_dt, 2 * a.K + k);
    const int fT = floor3d(sT), fH = floor3d(sH), fW = floor3d(sW);
    const float rT = sT - fT, rH = sH - fH, rW = sW - fW;
    const float wT0 = 1.f - rT, wH0 = 1.f - rH, wW0 = 1.f - rW;
    const int clip = img / T, t = img - clip * T;
    int h = p / W, w = p - h * W;
    // tap rows c = 2*a + b  (a: frame ts = t+fT+a, b: row hs = h+fH+b)
    const __nv_bfloat16 *rowp[4];
    bool rowok[4];
    auto set_rows = [&]() {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int ts = t + fT + (c >> 1), hs = h + fH + (c & 1);
            rowok[c] = ts >= 0 && ts < T && hs >= 0 && hs < H;
            rowp[c] = a.x + ((int64_t)((clip * T + ts) * a.K + k) * HW + (int64_t)hs * W);
        }
    };
    set_rows();
    float right[4] = {0.f, 0.f, 0.f, 0.f};
    bool cont = false;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float left[4];
        const int ws = w + fW;
        if (p + e < HW) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                left[c] = cont ? right[c] : ((rowok[c] && ws >= 0 && ws < W) ? bf16_lo(ldg2(rowp[c] + ws)) : 0.f);
                right[c] = (rowok[c] && ws + 1 >= 0 && ws + 1 < W) ? bf16_lo(ldg2(rowp[c] + ws + 1)) : 0.f;
            }
            const float l0 = left[0] * wW0 + right[0] * rW, l1 = left[1] * wW0 + right[1] * rW;
            const float l2 = left[2] * wW0 + right[2] * rW, l3 = left[3] * wW0 + right[3] * rW;
            o[e] = wT0 * (wH0 * l0 + rH * l1) + rT * (wH0 * l2 + rH * l3);
            cont = true;
            if (++w == W) {
                w = 0;
                ++h;
                cont = false;
                set_rows();
            }
        } else {
            o[e] = 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = pack_bf16x2(o[2 * i], o[2 * i + 1]);
}

template <int PROD, int VEC>
__global__ void __launch_bounds__(kThreads, 1) k_pw_conv(const PwArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Hdr *hdr = reinterpret_cast<Hdr *>(smem);
    unsigned char *smem_b = smem + a.off_b;
    unsigned char *smem_a = smem + a.off_a;
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * a.Ncta;

    // ---- one-time setup: barriers, TMEM, resident weight block --------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&hdr->full[i], kNumProdWarps);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hdr->tmem_full[i], 1);
            mbar_init(&hdr->tmem_empty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, (uint32_t)a.tmem_cols);
    }
    {
        // B[n, k] -> (k/8)*b_lbo + (n/8)*b_sbo + (n%8)*16 + (k%8)*2 ; rows n0+n >= N and columns k >= K are zero
        const int kgroups = a.Kpad >> 3;
        const bool fast = a.w_dt == RB_BF16 && !a.w_trans && (a.K & 7) == 0;
        const __nv_bfloat16 *wb = reinterpret_cast<const __nv_bfloat16 *>(a.w);
        const float *wf = reinterpret_cast<const float *>(a.w);
        for (int u = tid; u < a.Ncta * kgroups; u += kThreads) {
            const int kg = u / a.Ncta, n = u - kg * a.Ncta;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (n0 + n < a.N) {
                if (fast) {
                    if (kg * 8 < a.K) v = ldg16(wb + (int64_t)(n0 + n) * a.K + kg * 8);
                } else {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = kg * 8 + e;
                        const int64_t idx = a.w_trans ? (int64_t)k * a.N + (n0 + n) : (int64_t)(n0 + n) * a.K + k;
                        f[e] = k < a.K ? (a.w_dt == RB_BF16 ? __bfloat162float(wb[idx]) : __ldg(wf + idx)) : 0.f;
                    }
                    v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                   pack_bf16x2(f[6], f[7]));
                }
            }
            *reinterpret_cast<uint4 *>(smem_b + (size_t)kg * a.b_lbo + (size_t)(n >> 3) * a.b_sbo + (n & 7) * 16) = v;
        }
        if (PROD == PROD_BNRELU)
            for (int k = tid; k < a.Kpad; k += kThreads) {
                smem_sb[k] = k < a.K ? a.a_sb[2 * k] : 0.f;
                smem_sb[a.Kpad + k] = k < a.K ? a.a_sb[2 * k + 1] : 0.f;
            }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    const int tile0 = blockIdx.x, tstride = gridDim.x;
    const ShiftSrc ssrc{a.x, a.shift, a.shift_dt, a.T, a.H, a.W, a.HW, a.K};

    if (warp == 0) {
        // ===================================== MMA issuer ========================================================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(kTileM, a.sub_n, /*A MN-major*/ 1, /*B K-major*/ 0);
            const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
            int slot = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
                const int as = it % a.acc_stages;
                const uint32_t aph = (uint32_t)(it / a.acc_stages) & 1u;
                mbar_wait(&hdr->tmem_empty[as], aph ^ 1u);
                tc_fence_after();
                for (int st = 0; st < a.k_stages; ++st) {
                    mbar_wait(&hdr->full[slot], phase);
                    tc_fence_after();
                    const int ksteps = min(kStageK / 16, (a.Kpad - st * kStageK) >> 4);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t adesc =
                            smem_desc(a_base + slot * kStageBytes + ks * 2 * a.a_lbo, a.ad_lbo, a.ad_sbo, LAYOUT_NONE);
                        const int kg = (st * kStageK >> 3) + ks * 2;
                        for (int j = 0; j < a.n_sub; ++j) {
                            const uint64_t bdesc = smem_desc(b_base + kg * a.b_lbo + j * (a.sub_n >> 3) * a.b_sbo, a.bd_lbo,
                                                             a.bd_sbo, LAYOUT_NONE);
                            mma_bf16(tmem_base + as * a.Ncta + j * a.sub_n, adesc, bdesc, idesc, (st | ks) ? 1u : 0u);
                        }
                    }
                    mma_commit(&hdr->empty[slot]);
                    if (++slot == a.stages) { slot = 0; phase ^= 1u; }
                }
                mma_commit(&hdr->tmem_full[as]);
            }
        }
        __syncwarp();
    } else if (warp < kProdWarp0) {
        // ===================================== epilogue ==========================================================
        const int q = warp & 3, row = q * 32 + lane;
        int it = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            const int as = it % a.acc_stages;
            const uint32_t aph = (uint32_t)(it / a.acc_stages) & 1u;
            const int img = tile / a.tiles_per_img, p = (tile - img * a.tiles_per_img) * kTileM + row;
            const bool valid = p < a.HW;
            mbar_wait(&hdr->tmem_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * a.Ncta;
            const int64_t obase = ((int64_t)img * a.N + n0) * a.HW + p;
            for (int c0 = 0; c0 < a.Ncta && n0 + c0 < a.N; c0 += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(taddr + c0, v);
                float rr[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    rr[j] = (a.res != nullptr && valid && n0 + c0 + j < a.N)
                                ? __bfloat162float(a.res[obase + (int64_t)(c0 + j) * a.HW])
                                : 0.f;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (valid && n0 + c0 + j < a.N)
                        a.out[obase + (int64_t)(c0 + j) * a.HW] = __float2bfloat16_rn(__uint_as_float(v[j]) + rr[j]);
            }
            tc_fence_before();
            mbar_arrive(&hdr->tmem_empty[as]);
        }
    } else {
        // ===================================== A producers =======================================================
        const int pw = warp - kProdWarp0;
        const int mg = (pw & 3) * 4 + (lane >> 3);  // 8-pixel group inside the tile
        const int kk0 = (pw >> 2) * 8 + (lane & 7);   // channel inside a 16-channel slab (+ j*16)
        const uint32_t soff0 = (uint32_t)(pw >> 2) * a.a_lbo + (uint32_t)mg * a.a_sbo + (uint32_t)(lane & 7) * 16u;

        auto load_stage = [&](int tile, int st, uint32_t (&r)[4][4]) {
            const int img = tile / a.tiles_per_img, p = (tile - img * a.tiles_per_img) * kTileM + mg * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = st * kStageK + j * 16 + kk0;
                if (k < a.K && p < a.HW) {
                    if (PROD == PROD_SHIFT3D) {
                        shift3d_unit(ssrc, img, k, p, r[j]);
                    } else {
                        load_unit<VEC>(a.x + ((int64_t)img * a.K + k) * a.HW + p, a.HW - p, r[j]);
                        if (PROD == PROD_BNRELU) bn_relu_unit(r[j], smem_sb[k], smem_sb[a.Kpad + k]);
                    }
                } else {
                    r[j][0] = r[j][1] = r[j][2] = r[j][3] = 0u;
                }
            }
        };

        int tile = tile0, st = 0, slot = 0;
        uint32_t phase = 0;
        uint32_t cur[4][4];
        if (tile < a.total_tiles) load_stage(tile, st, cur);
        while (tile < a.total_tiles) {
            int ntile = tile, nst = st + 1;
            if (nst == a.k_stages) { nst = 0; ntile += tstride; }
            uint32_t nxt[4][4];
            if (ntile < a.total_tiles) load_stage(ntile, nst, nxt);
            mbar_wait(&hdr->empty[slot], phase ^ 1u);
            unsigned char *sp = smem_a + (size_t)slot * kStageBytes + soff0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4 *>(sp + (size_t)j * 2 * a.a_lbo) = make_uint4(cur[j][0], cur[j][1], cur[j][2], cur[j][3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->full[slot]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) cur[j][i] = nxt[j][i];
            tile = ntile;
            st = nst;
            if (++slot == a.stages) { slot = 0; phase ^= 1u; }
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// splits N over grid.y so that the resident weight block fits shared memory and the accumulators fit TMEM
bool plan(PwArgs &a, int prod, dim3 *grid, size_t *smem_bytes) {
    a.Kpad = round_up(a.K, 16);
    a.k_stages = cdiv(a.Kpad, kStageK);
    const int sb_bytes = prod == PROD_BNRELU ? round_up(2 * a.Kpad * 4, 128) : 0;
    int gy = 0, Ncta = 0, n_sub = 0, stages = 0;
    for (int cand = 1; cand <= 16; ++cand) {
        int nc = round_up(cdiv(a.N, cand), 16);
        const int ns = cdiv(nc, 256);
        nc = round_up(nc, 16 * ns);
        if (nc > 512) continue;
        const int b_bytes = nc * a.Kpad * 2;
        const int st = (kSmemLimit - kHdrBytes - sb_bytes - b_bytes) / kStageBytes;
        if (st < 2) continue;
        gy = cand; Ncta = nc; n_sub = ns; stages = st < kMaxStages ? st : kMaxStages;
        break;
    }
    if (!gy) return false;
    a.Ncta = Ncta; a.n_sub = n_sub; a.sub_n = Ncta / n_sub; a.stages = stages;
    a.acc_stages = (2 * Ncta <= 512) ? 2 : 1;
    int cols = 32;
    while (cols < a.acc_stages * Ncta) cols <<= 1;
    a.tmem_cols = cols;
    a.tiles_per_img = cdiv(a.HW, kTileM);
    a.total_tiles = a.NI * a.tiles_per_img;
    a.off_sb = kHdrBytes;
    a.off_b = kHdrBytes + sb_bytes;
    a.off_a = a.off_b + round_up(Ncta * a.Kpad * 2, 128);
    // canonical no-swizzle layouts: 8 x 16-byte core matrices, contiguous along the non-strided direction
    a.a_sbo = 128; a.a_lbo = (kTileM / 8) * 128;
    a.b_sbo = 128; a.b_lbo = (uint32_t)Ncta * 16;
    a.ad_lbo = a.a_lbo; a.ad_sbo = a.a_sbo; a.bd_lbo = a.b_lbo; a.bd_sbo = a.b_sbo;
    *smem_bytes = (size_t)a.off_a + (size_t)stages * kStageBytes;
    int ctas_x = sm_count() / gy;
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > a.total_tiles) ctas_x = a.total_tiles;
    *grid = dim3((unsigned)ctas_x, (unsigned)gy, 1);
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
    k_pw_conv<PROD, VEC><<<grid, kThreads, smem_bytes, s>>>(a);
    return launched("k_pw_conv");
}

template <int PROD> int launch_vec(const PwArgs &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    if constexpr (PROD == PROD_SHIFT3D) return launch<PROD, 1>(a, grid, smem_bytes, s);
    else {
    const bool base16 = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    if (base16 && a.HW % 8 == 0) return launch<PROD, 8>(a, grid, smem_bytes, s);
    if (base16 && a.HW % 4 == 0) return launch<PROD, 4>(a, grid, smem_bytes, s);
    if (base16 && a.HW % 2 == 0) return launch<PROD, 2>(a, grid, smem_bytes, s);
    return launch<PROD, 1>(a, grid, smem_bytes, s);
    }
}


// =====================================================================================================================
// Weight gradient of the 1x1 convolution:  dW[m, n] = sum_{i, p} G[i, m, p] * A(i, n, p)      (A as in the forward)
// Both operands are K-major here (the reduction runs over pixels, which are contiguous in NCHW), staged in the
// 128-byte-swizzled canonical layout: a row of 64 pixels = 128 bytes, 16-byte chunk c of row r stored at chunk
// c ^ (r % 8), so that 8 lanes reading one 128-byte global row segment write 8 different bank groups.
// grid = (pixel splits, N blocks, M blocks); every CTA reduces its pixel range into TMEM and writes one fp32
// partial [M, N] slice; k_wg_reduce sums the slices in a fixed order (deterministic, unlike atomics).
constexpr int kWgMaxStages = 4;
constexpr int kWgChunk = 64;          // pixels per stage
constexpr int kWgTileBytes = 128 * 128;  // one 128-row operand tile

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
};

struct WgHdr {
    uint64_t full[kWgMaxStages], empty[kWgMaxStages], tmem_full;
    uint32_t tmem_base;
};
static_assert(sizeof(WgHdr) <= kHdrBytes, "header");

template <int PROD, int VEC>
__global__ void __launch_bounds__(kThreads, 1) k_pw_wgrad(const WgArgs a) {
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
            mbar_init(&hdr->full[i], kNumProdWarps);
            mbar_init(&hdr->empty[i], 1);
        }
        mbar_init(&hdr->tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, (uint32_t)a.tmem_cols);
    }
    if (PROD == PROD_BNRELU)
        for (int k = tid; k < a.N; k += kThreads) {
            smem_sb[k] = a.x_sb[2 * k];
            smem_sb[a.N + k] = a.x_sb[2 * k + 1];
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(128, a.sub_n, 0, 0);
            const uint32_t sbase = smem_u32(stage0);
            int slot = 0;
            uint32_t phase = 0, acc = 0;
            for (int q = c_begin; q < c_end; ++q) {
                const int pc = q % a.cpi;
                const int kvalid = min(kWgChunk, a.HW - pc * kWgChunk);
                const int ksteps = (kvalid + 15) >> 4;
                mbar_wait(&hdr->full[slot], phase);
                tc_fence_after();
                const uint32_t abase = sbase + slot * a.stage_bytes, bbase = abase + a.Mt * kWgTileBytes;
                for (int ks = 0; ks < ksteps; ++ks) {
                    for (int mt = 0; mt < a.Mt; ++mt) {
                        const uint64_t adesc = smem_desc(abase + mt * kWgTileBytes + ks * 32, 16, 1024, LAYOUT_SW128);
                        for (int j = 0; j < a.n_sub; ++j) {
                            const uint64_t bdesc = smem_desc(bbase + j * a.sub_n * 128 + ks * 32, 16, 1024, LAYOUT_SW128);
                            mma_bf16(tmem_base + mt * a.Nc + j * a.sub_n, adesc, bdesc, idesc, acc);
                        }
                    }
                    acc = 1u;
                }
                mma_commit(&hdr->empty[slot]);
                if (++slot == a.stages) { slot = 0; phase ^= 1u; }
            }
            mma_commit(&hdr->tmem_full);
        }
        __syncwarp();
    } else if (warp < kProdWarp0) {
        const int q4 = warp & 3, row = q4 * 32 + lane;
        mbar_wait(&hdr->tmem_full, 0);
        tc_fence_after();
        float *dst = a.partial + (int64_t)blockIdx.x * a.M * a.N;
        for (int mt = 0; mt < a.Mt; ++mt) {
            const int ml = mt * 128 + row;
            const bool valid = ml < mrows;
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + mt * a.Nc;
            for (int c0 = 0; c0 < nrows; c0 += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                    float *o = dst + (int64_t)(m0 + ml) * a.N + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < nrows) o[j] = __uint_as_float(v[j]);
                }
            }
        }
    } else {
        const int pt = tid - kProdWarp0 * 32;
        const int total_units = (mrows + nrows) * 8;
        const ShiftSrc ssrc{a.x, a.shift, a.shift_dt, a.T, a.H, a.W, a.HW, a.N};
        int slot = 0;
        uint32_t phase = 0;
        for (int q = c_begin; q < c_end; ++q) {
            const int img = q / a.cpi, pc = q - img * a.cpi;
            const int p0 = pc * kWgChunk;
            const int kvalid = min(kWgChunk, a.HW - p0);
            const int nchunks16 = ((kvalid + 15) >> 4) * 2;  // 16-byte chunks the MMAs of this stage will read
            mbar_wait(&hdr->empty[slot], phase ^ 1u);
            unsigned char *abase = stage0 + (size_t)slot * a.stage_bytes, *bbase = abase + (size_t)a.Mt * kWgTileBytes;
            for (int u0 = pt; u0 < total_units; u0 += 8 * kNumProdWarps * 32) {
                uint32_t r[8][4];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int u = u0 + b * kNumProdWarps * 32;
                    const int rowi = u >> 3, c = u & 7;
                    r[b][0] = r[b][1] = r[b][2] = r[b][3] = 0u;
                    if (u < total_units && c < nchunks16) {
                        const int p = p0 + c * 8;
                        if (rowi < mrows) {
                            load_unit<VEC>(a.g + ((int64_t)img * a.M + m0 + rowi) * a.HW + p, a.HW - p, r[b]);
                        } else {
                            const int k = n0 + rowi - mrows;
                            if (PROD == PROD_SHIFT3D) {
                                if (p < a.HW) shift3d_unit(ssrc, img, k, p, r[b]);
                            } else {
                                load_unit<VEC>(a.x + ((int64_t)img * a.N + k) * a.HW + p, a.HW - p, r[b]);
                                if (PROD == PROD_BNRELU) {
                                    // padding pixels must stay zero: relu(0*s + b) may not be
                                    const int nv = a.HW - p;
                                    bn_relu_unit(r[b], smem_sb[k], smem_sb[a.N + k]);
                                    if (nv < 8) {
#pragma unroll
                                        for (int hh = 0; hh < 4; ++hh) {
                                            if (nv <= 2 * hh) r[b][hh] = 0u;
                                            else if (nv == 2 * hh + 1) r[b][hh] &= 0xffffu;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int u = u0 + b * kNumProdWarps * 32;
                    const int rowi = u >> 3, c = u & 7;
                    if (u < total_units && c < nchunks16) {
                        unsigned char *d;
                        if (rowi < mrows) {
                            d = abase + (rowi >> 7) * kWgTileBytes + ((rowi & 127) >> 3) * 1024 + (rowi & 7) * 128 +
                                ((c ^ (rowi & 7)) << 4);
                        } else {
                            const int rb = rowi - mrows;
                            d = bbase + (rb >> 3) * 1024 + (rb & 7) * 128 + ((c ^ (rb & 7)) << 4);
                        }
                        *reinterpret_cast<uint4 *>(d) = make_uint4(r[b][0], r[b][1], r[b][2], r[b][3]);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->full[slot]);
            if (++slot == a.stages) { slot = 0; phase ^= 1u; }
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

__global__ void k_wg_reduce(const float *__restrict__ partial, float *__restrict__ out, int splits, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * count + i];
    out[i] = s;
}

bool wg_plan(WgArgs &a, int prod, dim3 *grid, size_t *smem_bytes) {
    const int sb_bytes = prod == PROD_BNRELU ? round_up(2 * a.N * 4, 128) : 0;
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
    int want = sm_count() / (bmb * bnb);
    if (want < 1) want = 1;
    if (want > a.total_chunks) want = a.total_chunks;
    a.chunks_per_split = cdiv(a.total_chunks, want);
    const int splits = cdiv(a.total_chunks, a.chunks_per_split);
    // the last M / N block may be empty after rounding Mb / Nc up: shrink the grid to the blocks that own rows
    *grid = dim3((unsigned)splits, (unsigned)cdiv(a.N, a.Nc), (unsigned)cdiv(a.M, a.Mb));
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
    k_pw_wgrad<PROD, VEC><<<grid, kThreads, smem_bytes, s>>>(a);
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

int pw_conv_forward(const void *x, const void *w, int w_dt, int w_trans, const void *residual, void *out, int NI, int K,
                    int N, int HW, const float *a_sb, const void *shift, int shift_dt, int T, int H, int W, cudaStream_t s) {
    PwArgs a{};
    a.x = (const __nv_bfloat16 *)x; a.w = w; a.w_dt = w_dt; a.w_trans = w_trans; a.res = (const __nv_bfloat16 *)residual;
    a.out = (__nv_bfloat16 *)out; a.a_sb = a_sb; a.shift = shift; a.shift_dt = shift_dt;
    a.T = T; a.H = H; a.W = W; a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    const int prod = shift ? PROD_SHIFT3D : (a_sb ? PROD_BNRELU : PROD_PLAIN);
    dim3 grid;
    size_t smem_bytes = 0;
    if (!plan(a, prod, &grid, &smem_bytes))
        return fail(RB_ERR_UNSUPPORTED, "pw_conv: no tiling for K=%d N=%d (weight block does not fit shared memory)", K, N);
    if (const char *dbg = getenv("RB_PW_SWAP")) {  // debug: swap leading/stride offsets of A (bit 0) / B (bit 1)
        const int m = atoi(dbg);
        if (m & 1) { uint32_t t = a.ad_lbo; a.ad_lbo = a.ad_sbo; a.ad_sbo = t; }
        if (m & 2) { uint32_t t = a.bd_lbo; a.bd_lbo = a.bd_sbo; a.bd_sbo = t; }
    }
    if (prod == PROD_SHIFT3D) return launch_vec<PROD_SHIFT3D>(a, grid, smem_bytes, s);
    if (prod == PROD_BNRELU) return launch_vec<PROD_BNRELU>(a, grid, smem_bytes, s);
    return launch_vec<PROD_PLAIN>(a, grid, smem_bytes, s);
}

}  // namespace rb

namespace rb {

// scratch floats needed by pw_conv_wgrad: [splits, M, N]
size_t pw_conv_wgrad_workspace(int NI, int M, int N, int HW) {
    WgArgs a{};
    a.NI = NI; a.M = M; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem_bytes = 0;
    if (!wg_plan(a, PROD_BNRELU, &grid, &smem_bytes)) return 0;
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
    if (!wg_plan(a, PROD_BNRELU, &grid, &smem_bytes))
        return fail(RB_ERR_UNSUPPORTED, "pw_conv_wgrad: no tiling for M=%d N=%d", M, N);
    int rc;
    if (prod == PROD_SHIFT3D) rc = wg_launch_vec<PROD_SHIFT3D>(a, grid, smem_bytes, s);
    else if (prod == PROD_BNRELU) rc = wg_launch_vec<PROD_BNRELU>(a, grid, smem_bytes, s);
    else rc = wg_launch_vec<PROD_PLAIN>(a, grid, smem_bytes, s);
    if (rc) return rc;
    const int64_t count = (int64_t)M * N;
    k_wg_reduce<<<(unsigned)cdiv64(count, 256), 256, 0, s>>>(a.partial, dw, (int)grid.x, count);
    return launched("k_wg_reduce");
}

}  // namespace rb

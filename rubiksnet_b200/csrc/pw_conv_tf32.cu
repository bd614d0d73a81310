// Pointwise (1x1) convolution of RubiksShiftBlock on fp32 NCHW activations as a tcgen05 kind::tf32 GEMM -- the fp32
// inference path (BASELINE C2: RubiksNet-Tiny, eval mode), where the reference runs cuDNN's TF32 convolutions
// (torch.backends.cudnn.allow_tf32 defaults to True) between separate BatchNorm and ReLU passes.
//
//     out[i, n, p] = post( sum_k W[n, k] * A(i, k, p) ),      A = x  or  relu(x * in_scale[k] + in_bias[k])
//     post(v)      = [relu]( v * out_scale[n] + out_bias[n] + pre[i, n, p] ) + residual[i, n, p]
//
// so that an eval-mode block is three launches: conv2 with bn1+relu folded into the operand producer and bn2+relu folded
// into the epilogue (rubiksnet/backbone.py:123-128), the shift kernel, conv3 with the shortcut added in the epilogue
// (backbone.py:133-135).  Operands are fp32 words read by the tensor core as TF32 (10-bit mantissa; weights are rounded
// to nearest once, activations are truncated by the hardware like in cuDNN's kernels), accumulation is fp32 in tensor
// memory, everything after the accumulator is fp32.
//
// Structure (same plan as the bf16 kernel k_pw_conv, csrc/pw_conv.cu): the GEMM is turned around -- M = output channels
// (TMEM lanes; the [128 x Kpad] weight block stays resident in shared memory as the K-major operand), N = 128 positions
// of the flattened (image, pixel) axis (TMEM columns; the activation tile is the MN-major, pixel-contiguous operand in the
// 128-byte-swizzled layout with 32-byte granules that 32-bit MN-major operands require: rows of 32 pixels), K = input
// channels in stages of 16.  Persistent CTAs, 16 warps:
//     warp 0      one elected thread issues tcgen05.mma.kind::tf32 (128 x 128 x 8 per instruction)
//     warps 1-8   epilogue: tcgen05.ld -> scale/bias -> per-warp swizzled staging tile -> coalesced 128-byte row segments
//     warps 9-15  producers, one warp per 8 KiB stage: 16 x LDG.128 per lane in flight, BN+ReLU in registers, swizzled STS
// Output channels beyond 128 go to grid.y; a contraction too long for a resident weight block is split over launches
// that chain through `pre` (the partial sums take the place of the bias).
#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int kTfEpiWarp0 = 1, kTfNumEpi = 8;
constexpr int kTfProdWarp0 = kTfEpiWarp0 + kTfNumEpi, kTfNumProd = 7;
constexpr int kTfThreads = (kTfProdWarp0 + kTfNumProd) * 32;  // 512: a thread may use 128 registers
constexpr int kTfMaxStages = 12;
constexpr int kTfStageCh = 16;                                // channels per stage
constexpr int kTfNpx = 128;                                   // positions per tile
constexpr int kTfStageBytes = kTfStageCh * kTfNpx * 4;        // 8 KiB
constexpr int kTfHdr = 256;
constexpr int kTfStgWarp = 32 * 128;                          // per epilogue warp: 32 channels x 32 positions fp32
constexpr int kTfStgBytes = kTfNumEpi * kTfStgWarp;
constexpr int kTfSmem = 227 * 1024;
constexpr int kTfRows = 128;                                  // output channels per CTA (= MMA M)
constexpr int kTfWLbo = kTfRows * 16 + 16;                    // bytes between K-adjacent core matrices of the weight block
constexpr int kTfMinStages = 4;

struct TfArgs {
    const float *x;       // [NI, Ktot, HW], already offset to the first contracted channel
    const float *w;       // [N, Ktot], already offset to the first contracted column
    const float *res;     // [NI, N, HW] or null
    const float *pre;     // [NI, N, HW] or null (partial sums of an earlier K slice)
    float *out;           // [NI, N, HW]
    const float *in_sb;   // [K, 2] (scale, bias) of the producer's relu(x*s+b), or null
    const float *out_sb;  // [N, 2] (scale, bias) applied to the accumulator, or null
    int out_bias, relu, w_resident;
    int NI, K, N, HW;
    int x_img_stride;     // elements between images of x (Ktot * HW)
    int w_row_stride;     // elements between rows of w (Ktot)
    int Kpad, k_stages, stages, NP, total_tiles;
    uint32_t off_sb, off_stg, off_w, off_a;
    uint32_t hw_mul, hw_shr;
    uint32_t kp_mul, kp_shr;  // exact division by Kpad for indices < 2^31
    uint32_t kq_mul, kq_shr;  // ... by Kpad / 4
    int w_vec4;               // weight rows may be read with 16-byte loads
};

struct TfHdr {
    uint64_t full[kTfMaxStages], empty[kTfMaxStages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(TfHdr) <= kTfHdr, "header");

__device__ __forceinline__ int tf_div_hw(const TfArgs &a, int P) {
    return a.hw_mul == 0u ? P : (int)(__umulhi((uint32_t)P, a.hw_mul) >> a.hw_shr);
}

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
}

// kind::tf32 instruction descriptor: D fp32, A/B tf32, A (weights) K-major, B (activations) MN-major, 128 x 128
__device__ __forceinline__ uint32_t tf_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(kTfNpx >> 3) << 17) | ((uint32_t)(kTfRows >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Resident weight block: W[n0 + n, k] (n < nrows, k < K) -> (k/4)*kTfWLbo + (n/8)*128 + (n%8)*16 + (k%4)*4, i.e. K-major
// 8-row x 16-byte core matrices; rows >= nrows and columns >= K are zero.  The 128 x Kpad elements are walked as one flat
// index (consecutive threads = consecutive k of a row: coalesced 4-byte reads, and kTfWLbo / 4 = 516 = 4 (mod 32) keeps the
// scattered 4-byte stores of a warp in different banks); a thread keeps 16 loads in flight, so the whole block costs a few
// memory latencies instead of one per element.
__device__ __forceinline__ void tf_stage_weights(const TfArgs &a, unsigned char *smem_w, int n0, int nrows, int t, int nthreads) {
    if (a.w_vec4) {
        // rows are 16-byte aligned and K % 4 == 0: one 16-byte load = one core-matrix row (4 consecutive k of a weight row)
        constexpr int UB = 8;
        const int kq = a.Kpad >> 2, total = kTfRows * kq;
        for (int u0 = t; u0 < total; u0 += UB * nthreads) {
            float4 v[UB];
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int u = u0 + b * nthreads;
                const int n = (int)(__umulhi((uint32_t)u, a.kq_mul) >> a.kq_shr), q = u - n * kq;
                v[b] = (u < total && n < nrows && 4 * q < a.K)
                           ? __ldg(reinterpret_cast<const float4 *>(a.w + (int64_t)(n0 + n) * a.w_row_stride) + q)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int u = u0 + b * nthreads;
                if (u >= total) continue;
                const int n = (int)(__umulhi((uint32_t)u, a.kq_mul) >> a.kq_shr), q = u - n * kq;
                *reinterpret_cast<uint4 *>(smem_w + q * kTfWLbo + (n >> 3) * 128 + (n & 7) * 16) =
                    make_uint4(to_tf32(v[b].x), to_tf32(v[b].y), to_tf32(v[b].z), to_tf32(v[b].w));
            }
        }
        return;
    }
    constexpr int UB = 16;
    const int total = kTfRows * a.Kpad;
    for (int e0 = t; e0 < total; e0 += UB * nthreads) {
        float v[UB];
#pragma unroll
        for (int b = 0; b < UB; ++b) {
            const int e = e0 + b * nthreads;
            const int n = (int)(__umulhi((uint32_t)e, a.kp_mul) >> a.kp_shr), k = e - n * a.Kpad;
            v[b] = (e < total && n < nrows && k < a.K) ? __ldg(a.w + (int64_t)(n0 + n) * a.w_row_stride + k) : 0.f;
        }
#pragma unroll
        for (int b = 0; b < UB; ++b) {
            const int e = e0 + b * nthreads;
            if (e >= total) continue;
            const int n = (int)(__umulhi((uint32_t)e, a.kp_mul) >> a.kp_shr), k = e - n * a.Kpad;
            *reinterpret_cast<uint32_t *>(smem_w + (k >> 2) * kTfWLbo + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4) = to_tf32(v[b]);
        }
    }
}

template <int VEC, bool BN>
__global__ void __launch_bounds__(kTfThreads, 1) k_pw_tf32(const TfArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    TfHdr *hdr = reinterpret_cast<TfHdr *>(smem);
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);
    unsigned char *smem_w = smem + a.off_w;
    unsigned char *smem_a = smem + a.off_a;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * kTfRows;
    const int nrows = min(kTfRows, a.N - n0);

    if (tid == 0) {
        for (int i = 0; i < kTfMaxStages; ++i) {
            mbar_init(&hdr->full[i], 1);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hdr->tmem_full[i], 1);
            mbar_init(&hdr->tmem_empty[i], kTfNumEpi * 32);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, 256u);
    }
    // nothing above touches global memory; a resident weight block (RB_W_RESIDENT: parameters of an eval-mode network) is
    // staged while the previous kernel of the stream drains (common.cuh: programmatic dependent launch)
    if (a.w_resident) {  // every warp: nobody has anything else to do yet
        tf_stage_weights(a, smem_w, n0, nrows, tid, kTfThreads);
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core after the CTA sync below
    }
    pdl_sync();
    if (BN)
        for (int k = tid; k < a.Kpad; k += kTfThreads) {
            smem_sb[k] = k < a.K ? a.in_sb[2 * k] : 0.f;
            smem_sb[a.Kpad + k] = k < a.K ? a.in_sb[2 * k + 1] : 0.f;
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    if (!a.w_resident && warp < kTfProdWarp0) {
        // the producer warps already load activations while the MMA + epilogue warps stage the weight block
        tf_stage_weights(a, smem_w, n0, nrows, tid, kTfProdWarp0 * 32);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kTfProdWarp0 * 32) : "memory");
    }

    const int tile0 = blockIdx.x, tstride = gridDim.x;

    if (warp == 0) {
        // ===================================== MMA issuer: one thread ============================================
        if (elect_one()) {
            const uint32_t idesc = tf_idesc();
            const uint64_t adesc0 = smem_desc(smem_u32(smem_w), kTfWLbo, 128, LAYOUT_NONE);
            // activations, MN-major.  32-bit MN-major operands have exactly one shared-memory layout (cute: Layout_MN_SW128_32B_Atom,
            // "for mn-major tf32 operands, SW128_32B is the only available smem layout"): atoms of 4 channel rows x 32 positions
            // (128 B per row), the 32-byte chunks of a row XOR-ed with (row % 4); atoms of consecutive channel groups 512 B
            // apart (SBO), the next 32 positions (kTfStageCh / 4) * 512 B further (LBO)
            const uint64_t bdesc0 = smem_desc(smem_u32(smem_a), (kTfStageCh >> 2) * 512u, 512, LAYOUT_SW128_BASE32B);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
            const uint32_t a_kstep = (2u * kTfWLbo) >> 4;  // one K = 8 step: two core-matrix columns
            int slot = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
                const int as = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                const uint32_t tacc = tmem_base + (uint32_t)as * kTfNpx;
                mbar_wait(&hdr->tmem_empty[as], aph ^ 1u);
                tc_fence_after();
                uint32_t a_lo = a_lo0, acc = 0u;
                int kleft = a.Kpad >> 3;
                for (int st = 0; st < a.k_stages; ++st) {
                    mbar_wait(&hdr->full[slot], phase);
                    tc_fence_after();
                    const int ksteps = min(kTfStageCh >> 3, kleft);
                    kleft -= ksteps;
                    uint32_t b_lo = b_lo0 + (uint32_t)slot * (kTfStageBytes >> 4);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        mma_tf32_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                        acc = 1u;
                        a_lo += a_kstep;
                        b_lo += 1024u >> 4;
                    }
                    mma_commit(&hdr->empty[slot]);
                    if (++slot == a.stages) { slot = 0; phase ^= 1u; }
                }
                mma_commit(&hdr->tmem_full[as]);
            }
        }
        __syncwarp();
    } else if (warp < kTfProdWarp0) {
        // ===================================== epilogue ==========================================================
        // warp -> TMEM lane quarter (hardware: warp id % 4) and one half of the tile's 128 columns
        const int q = warp & 3, half = (warp - kTfEpiWarp0) >> 2;
        unsigned char *stg = smem + a.off_stg + (warp - kTfEpiWarp0) * kTfStgWarp;
        const int myrow = q * 32 + lane;  // CTA-local output channel of this thread's TMEM lane
        float os = 1.f, ob = 0.f;
        if (a.out_sb != nullptr && myrow < nrows) {
            os = a.out_sb[2 * (n0 + myrow)];
            ob = a.out_bias ? a.out_sb[2 * (n0 + myrow) + 1] : 0.f;
        }
        const int chunk = lane & 7, rsub = lane >> 3;  // read-back geometry: 8 lanes per 128-byte row, 4 rows per access
        int it = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            const int as = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            const int P0 = tile * kTfNpx;
            mbar_wait(&hdr->tmem_full[as], aph);
            tc_fence_after();
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                const int c0 = half * 64 + r * 32;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kTfNpx + c0);
                uint32_t v[2][16];
                __syncwarp();  // the previous round's read-back of the staging tile is over
                tmem_ld16(taddr, v[0]);
                tmem_ld16(taddr + 16, v[1]);
                // this lane's positions of the read-back phase: 4 consecutive positions from P
                const int P = P0 + c0 + chunk * 4;
                int off[VEC == 4 ? 1 : 4];
                bool ok[VEC == 4 ? 1 : 4];
#pragma unroll
                for (int e = 0; e < (VEC == 4 ? 1 : 4); ++e) {
                    const int Pe = P + e;
                    const int img = tf_div_hw(a, Pe), pp = Pe - img * a.HW;
                    ok[e] = Pe < a.NP;
                    off[e] = (img * a.N + n0) * a.HW + pp;
                }
                tmem_ld_wait();
                if (r == 1) {  // every TMEM load of this tile has landed: hand the accumulator stage back
                    tc_fence_before();
                    mbar_arrive(&hdr->tmem_empty[as]);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t *vv = &v[c >> 2][(c & 3) * 4];
                    float4 o;
                    o.x = fmaf(__uint_as_float(vv[0]), os, ob);
                    o.y = fmaf(__uint_as_float(vv[1]), os, ob);
                    o.z = fmaf(__uint_as_float(vv[2]), os, ob);
                    o.w = fmaf(__uint_as_float(vv[3]), os, ob);
                    *reinterpret_cast<float4 *>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = o;
                }
                __syncwarp();
                // read-back: lane = (row rsub + 4 i, 16-byte chunk).  `pre` / `res` of all 8 rows are fetched first (all loads
                // in flight together; `pre` may alias `out`, so the compiler would not move them above the stores itself)
                float4 pv[8], rv[8];
                bool live[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int grow = q * 32 + rsub + 4 * i;
                    const int o = off[0] + grow * a.HW;
                    live[i] = grow < nrows && (VEC == 4 ? ok[0] : true);
                    pv[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!live[i]) continue;
                    if (VEC == 4) {
                        if (a.pre != nullptr) pv[i] = *reinterpret_cast<const float4 *>(a.pre + o);
                        if (a.res != nullptr) rv[i] = __ldg(reinterpret_cast<const float4 *>(a.res + o));
                    } else {
                        const int roff = grow * a.HW;
                        float *pp = reinterpret_cast<float *>(&pv[i]), *rp = reinterpret_cast<float *>(&rv[i]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!ok[VEC == 4 ? 0 : e]) continue;
                            const int oe = off[VEC == 4 ? 0 : e] + roff;
                            if (a.pre != nullptr) pp[e] = a.pre[oe];
                            if (a.res != nullptr) rp[e] = a.res[oe];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rrow = rsub + 4 * i;
                    const int grow = q * 32 + rrow;
                    float4 t = *reinterpret_cast<const float4 *>(stg + rrow * 128 + ((chunk ^ (rrow & 7)) << 4));
                    if (!live[i]) continue;
                    t.x += pv[i].x; t.y += pv[i].y; t.z += pv[i].z; t.w += pv[i].w;
                    if (a.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
                    t.x += rv[i].x; t.y += rv[i].y; t.z += rv[i].z; t.w += rv[i].w;
                    const int roff = grow * a.HW;
                    if (VEC == 4) {
                        *reinterpret_cast<float4 *>(a.out + off[0] + roff) = t;
                    } else {
                        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (ok[VEC == 4 ? 0 : e]) a.out[off[VEC == 4 ? 0 : e] + roff] = tv[e];
                    }
                }
            }
        }
    } else {
        // ===================================== producers: one warp per stage =====================================
        const int W = min(kTfNumProd, a.stages);
        const int w = warp - kTfProdWarp0;
        if (w < W) {
            int slot = 0, next = w, cnt = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride) {
                const int P = tile * kTfNpx + lane * 4;
                int off[VEC == 4 ? 1 : 4];
                bool ok[VEC == 4 ? 1 : 4];
#pragma unroll
                for (int e = 0; e < (VEC == 4 ? 1 : 4); ++e) {
                    const int Pe = P + e;
                    const int img = tf_div_hw(a, Pe), pp = Pe - img * a.HW;
                    ok[e] = Pe < a.NP;
                    off[e] = img * a.x_img_stride + pp;
                }
                for (int st = 0; st < a.k_stages; ++st, ++cnt) {
                    if (cnt == next) {
                        next += W;
                        const int cbase = st * kTfStageCh;
                        float4 v[kTfStageCh];
#pragma unroll
                        for (int j = 0; j < kTfStageCh; ++j) {
                            const int c = cbase + j;
                            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (c < a.K) {
                                const float *src = a.x + c * a.HW;
                                if (VEC == 4) {
                                    if (ok[0]) v[j] = __ldg(reinterpret_cast<const float4 *>(src + off[0]));
                                } else {
                                    if (ok[0]) v[j].x = __ldg(src + off[0]);
                                    if (ok[VEC == 4 ? 0 : 1]) v[j].y = __ldg(src + off[VEC == 4 ? 0 : 1]);
                                    if (ok[VEC == 4 ? 0 : 2]) v[j].z = __ldg(src + off[VEC == 4 ? 0 : 2]);
                                    if (ok[VEC == 4 ? 0 : 3]) v[j].w = __ldg(src + off[VEC == 4 ? 0 : 3]);
                                }
                            }
                        }
                        mbar_wait(&hdr->empty[slot], phase ^ 1u);
                        unsigned char *dst = smem_a + slot * kTfStageBytes + (lane >> 3) * ((kTfStageCh >> 3) * 1024);
#pragma unroll
                        for (int j = 0; j < kTfStageCh; ++j) {
                            float4 o = v[j];
                            if (BN) {
                                const int c = cbase + j;  // < Kpad: coefficients of padded channels are zero
                                const float s = smem_sb[c], b = smem_sb[a.Kpad + c];
                                o.x = fmaxf(fmaf(o.x, s, b), 0.f);
                                o.y = fmaxf(fmaf(o.y, s, b), 0.f);
                                o.z = fmaxf(fmaf(o.z, s, b), 0.f);
                                o.w = fmaxf(fmaf(o.w, s, b), 0.f);
                            }
                            // row j of the stage: atom j / 4, row j % 4; this lane's 16 bytes are half (lane & 1) of 32-byte
                            // chunk (lane & 7) / 2 of the row
                            *reinterpret_cast<float4 *>(dst + (j >> 2) * 512 + (j & 3) * 128 +
                                                        (((((lane & 7) >> 1) ^ (j & 3)) << 5) | ((lane & 1) << 4))) = o;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&hdr->full[slot]);
                    }
                    if (++slot == a.stages) { slot = 0; phase ^= 1u; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256u);
    }
}

int tf_round_up(int v, int m) { return (v + m - 1) / m * m; }

// shared-memory plan for a contraction of K channels; false = the weight block leaves fewer than kTfMinStages stages
bool tf_plan(TfArgs &a) {
    a.Kpad = tf_round_up(a.K, 8);
    a.k_stages = cdiv(a.Kpad, kTfStageCh);
    const int sb_bytes = a.in_sb ? tf_round_up(2 * a.Kpad * 4, 128) : 0;
    a.off_sb = kTfHdr;
    a.off_stg = a.off_sb + sb_bytes;
    a.off_w = a.off_stg + kTfStgBytes;
    a.off_a = (uint32_t)tf_round_up((int)a.off_w + (a.Kpad >> 2) * kTfWLbo, 1024);
    const int room = kTfSmem - (int)a.off_a;
    if (room < kTfMinStages * kTfStageBytes) return false;
    a.stages = room / kTfStageBytes;
    if (a.stages > kTfMaxStages) a.stages = kTfMaxStages;
    return true;
}

template <int VEC, bool BN> int tf_launch(const TfArgs &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_pw_tf32<VEC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTfSmem);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_pw_tf32): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    launch_kernel(k_pw_tf32<VEC, BN>, grid, dim3(kTfThreads), smem_bytes, s, a);
    return launched("k_pw_tf32");
}

}  // namespace

int pw_tf32_forward(const float *x, const float *w, const float *res, float *out, int NI, int K, int N, int HW,
                    const float *in_sb, const float *out_sb, int relu, int resident, cudaStream_t s) {
    // longest contraction one launch can keep resident: a multiple of the stage depth
    int kmax = 0;
    for (int k = kTfStageCh; k <= 4096; k += kTfStageCh) {
        TfArgs t{};
        t.K = k; t.in_sb = in_sb;
        if (!tf_plan(t)) break;
        kmax = k;
    }
    if (kmax == 0) return fail(RB_ERR_UNSUPPORTED, "pw_tf32: no tiling");
    const int parts = cdiv(K, kmax);
    const int kpart = parts == 1 ? K : tf_round_up(cdiv(K, parts), kTfStageCh);
    const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res);
    const int vec = ((align & 15) == 0 && HW % 4 == 0) ? 4 : 1;
    for (int p = 0, k0 = 0; k0 < K; ++p, k0 += kpart) {
        const bool last = k0 + kpart >= K;
        TfArgs a{};
        a.x = x + (int64_t)k0 * HW; a.w = w + k0; a.out = out;
        a.in_sb = in_sb ? in_sb + 2 * k0 : nullptr;
        a.out_sb = out_sb;
        a.pre = p > 0 ? out : nullptr;
        a.res = last ? res : nullptr;
        a.out_bias = last ? 1 : 0;
        a.relu = last ? relu : 0;
        a.w_resident = resident;
        a.NI = NI; a.K = (last ? K - k0 : kpart); a.N = N; a.HW = HW;
        a.x_img_stride = K * HW; a.w_row_stride = K;
        if (!tf_plan(a)) return fail(RB_ERR_UNSUPPORTED, "pw_tf32: no tiling for K=%d", a.K);
        a.NP = NI * HW;
        a.total_tiles = cdiv(a.NP, kTfNpx);
        if (HW <= 1) {
            a.hw_mul = 0u; a.hw_shr = 0u;
        } else {
            uint32_t l = 0;
            while ((1u << l) < (uint32_t)HW) ++l;
            a.hw_mul = (uint32_t)(((uint64_t(1) << (31 + l)) + (uint32_t)HW - 1) / (uint32_t)HW);
            a.hw_shr = l - 1;
        }
        {
            uint32_t l = 0;
            while ((1u << l) < (uint32_t)a.Kpad) ++l;  // Kpad >= 8
            a.kp_mul = (uint32_t)(((uint64_t(1) << (31 + l)) + (uint32_t)a.Kpad - 1) / (uint32_t)a.Kpad);
            a.kp_shr = l - 1;
            const uint32_t kq = (uint32_t)a.Kpad >> 2;  // >= 2
            l = 0;
            while ((1u << l) < kq) ++l;
            a.kq_mul = (uint32_t)(((uint64_t(1) << (31 + l)) + kq - 1) / kq);
            a.kq_shr = l - 1;
            a.w_vec4 = (a.K % 4 == 0 && K % 4 == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0) ? 1 : 0;
        }
        const int gy = cdiv(N, kTfRows);
        int gx = sm_count() / gy;
        if (gx < 1) gx = 1;
        if (gx > a.total_tiles) gx = a.total_tiles;
        const dim3 grid((unsigned)gx, (unsigned)gy, 1);
        const size_t smem_bytes = (size_t)a.off_a + (size_t)a.stages * kTfStageBytes;
        int rc;
        if (in_sb) rc = vec == 4 ? tf_launch<4, true>(a, grid, smem_bytes, s) : tf_launch<1, true>(a, grid, smem_bytes, s);
        else rc = vec == 4 ? tf_launch<4, false>(a, grid, smem_bytes, s) : tf_launch<1, false>(a, grid, smem_bytes, s);
        if (rc) return rc;
    }
    return RB_OK;
}

}  // namespace rb

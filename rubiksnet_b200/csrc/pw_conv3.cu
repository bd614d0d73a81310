// Pointwise (1x1) convolution for the LARGE maps (112x112, 56x56, 28x28: row pitch of a channel plane = a multiple of 16
// bytes), third schedule: the tensor-map (tiled) mode of the TMA unit writes the activation tile straight into the
// MN-major SWIZZLE_128B UMMA operand layout and reads the finished tile back out of a swizzled staging block -- no register
// round trip on either side, which is what bounds k_pw_conv on these maps (csrc/pw_conv.cu: LDG -> registers -> STS
// producers, TMEM -> registers -> staging -> STG epilogue; 0.28 ms against cuBLAS's 0.17 ms at 72 ch x 112x112).
//
//     out[i, n, p] = sum_k W[n, k] * A(i, k, p)  (+ residual[i, n, p])     A = x  or  relu(x * scale[k] + bias[k])
//
// (conv3 + shortcut, the shortcut conv, conv1 on the patch matrix, every input gradient, and conv2 behind bn1 -> relu: eight
// extra warps apply the affine + ReLU in place on a landed stage -- 16-byte shared-memory vectors, a channel is a row -- and
// hand it to the MMA warp through `ready`.)
//
//   tile       = 128 consecutive pixels of ONE image x all K channels: two boxes of (64 pixels x K rows) = the two 64-pixel
//                column blocks of the operand (LBO apart); pixels beyond the plane are zero-filled by the TMA unit on the
//                way in and clipped on the way out.
//   warp 0     TMA loads: activation boxes into a ring of stages, residual boxes into the staging buffers
//   warp 1     one elected thread issues tcgen05.mma (128 x 128 x 16), accumulators double-buffered in tensor memory
//   warps 2-9  epilogue: tcgen05.ld -> bf16 (+ residual from the staging buffer) -> staging buffer (same swizzled box
//              layout) -> one thread issues the two TMA stores of the tile
//   warps 10-17 (BN+ReLU producer only) relu(x*s+b) on the stage between `full` and `ready`
//   weights    resident [128 x Kpad] K-major block (<= 128 output channels per CTA, more go to grid.y), staged before the
//              dependency wait when RB_W_RESIDENT
// Arithmetic is the same as k_pw_conv: bf16 operands, fp32 accumulation in TMEM, result rounded to bf16, `+= shortcut` on
// the rounded value.
#include <cuda.h>

#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int kP3Warps = 10;
constexpr int kP3Threads = kP3Warps * 32;
constexpr int kP3EpiWarp0 = 2, kP3NumEpi = 8;
constexpr int kP3BnWarp0 = 10, kP3NumBn = 8;  // launched only with the BN+ReLU producer
constexpr int kP3ThreadsBn = (kP3BnWarp0 + kP3NumBn) * 32;
constexpr int kP3MaxStages = 6;
constexpr int kP3Smem = 227 * 1024;
constexpr int kP3Hdr = 4096;  // barriers [0, 512) + (scale, bias) of up to 256 input channels [512, 2560)
constexpr int kP3SbOff = 512;
constexpr int kP3Rows = 128;  // output channels per CTA = MMA M
constexpr int kP3Npx = 128;

struct P3Args {
    const void *w;  // bf16 [N, K] (rb_pw_weight_pack) or fp32 [N, K]
    const float *a_sb;  // BN+ReLU producer: (scale, bias) pairs [K, 2]; NULL = plain
    double *stats;      // BatchNorm statistics of the output: [N][stats_splits][2] (sum, sum of squares); NULL = none
    int stats_splits;   // 2 * gridDim.x: one slot per CTA and 64-pixel column block
    int w_f32, w_resident, has_res;
    int NI, K, N, HW;
    int Kpad, Ncta, stages, tiles_per_image, total_tiles;
    uint32_t w_lbo, off_w, off_a, off_stg, stage_bytes, stg_bytes;
};

struct P3Hdr {
    uint64_t full[kP3MaxStages], empty[kP3MaxStages], ready[kP3MaxStages], tmem_full[2], tmem_empty[2], res_full[2], stg_free[2];
    uint32_t tmem_base;
};
static_assert(sizeof(P3Hdr) <= kP3SbOff, "header");
__device__ __forceinline__ uint32_t p3_bn_relu2(uint32_t w, float sc, float bi) {
    return pack_bf16x2(fmaxf(fmaf(bf16_lo(w), sc, bi), 0.f), fmaxf(fmaf(bf16_hi(w), sc, bi), 0.f));
}

__device__ __forceinline__ void p3_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// box (64 pixels x rows x 1 image) at (pixel, channel, image): global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void p3_load_box(uint32_t dst, const CUtensorMap *map, int px, int ch, int img, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(map), "r"(px), "r"(ch), "r"(img), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void p3_store_box(const CUtensorMap *map, uint32_t src, int px, int ch, int img) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(px),
                 "r"(ch), "r"(img)
                 : "memory");
}
__device__ __forceinline__ void p3_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void p3_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void p3_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 p3_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void p3_sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// bf16(acc pair) + bf16 residual pair, rounded once more (the module graph's `out += shortcut` on bf16 tensors)
__device__ __forceinline__ uint32_t p3_add_bf16x2(uint32_t a, uint32_t b) {
    return pack_bf16x2(bf16_lo(a) + bf16_lo(b), bf16_hi(a) + bf16_hi(b));
}

// weight block W[n0 + n, k] -> (k/8)*w_lbo + (n/8)*128 + (n%8)*16 + (k%8)*2; rows >= nrows and columns >= K are zero
__device__ __forceinline__ void p3_stage_weights(const P3Args &a, unsigned char *smem_w, int n0, int nrows, int tid) {
    const int kg = a.Kpad >> 3, total = kP3Rows * kg, nthreads = (int)blockDim.x;
    const __nv_bfloat16 *wb = reinterpret_cast<const __nv_bfloat16 *>(a.w);
    const float *wf = reinterpret_cast<const float *>(a.w);
    for (int u = tid; u < total; u += nthreads) {
        const int n = u / kg, g = u - n * kg;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (n < nrows && g * 8 < a.K) {
            const int64_t e = (int64_t)(n0 + n) * a.K + g * 8;  // K % 8 == 0: 16-byte aligned in bf16
            if (!a.w_f32) {
                o = __ldg(reinterpret_cast<const uint4 *>(wb + e));
            } else {
                const float4 lo = __ldg(reinterpret_cast<const float4 *>(wf + e)), hi = __ldg(reinterpret_cast<const float4 *>(wf + e + 4));
                o = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
            }
        }
        *reinterpret_cast<uint4 *>(smem_w + (size_t)g * a.w_lbo + (n >> 3) * 128 + (n & 7) * 16) = o;
    }
}

// STATS: the epilogue also reduces the BatchNorm statistics of the output (a separate instantiation: the extra registers and
// arithmetic slow the plain epilogue down, 0.176 -> 0.217 ms at 72 ch x 112x112, even when they are skipped at run time)
template <bool STATS>
__global__ void __launch_bounds__(kP3ThreadsBn, 1)
k_pw3(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
      const P3Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    P3Hdr *hdr = reinterpret_cast<P3Hdr *>(smem);
    unsigned char *smem_w = smem + a.off_w;
    const uint32_t s_a = smem_u32(smem + a.off_a), s_stg = smem_u32(smem + a.off_stg);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * a.Ncta;
    const int nrows = min(a.Ncta, a.N - n0);

    if (tid == 0) {
        for (int i = 0; i < kP3MaxStages; ++i) {
            mbar_init(&hdr->full[i], 1);
            mbar_init(&hdr->empty[i], 1);
            mbar_init(&hdr->ready[i], kP3NumBn);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hdr->tmem_full[i], 1);
            mbar_init(&hdr->tmem_empty[i], kP3NumEpi * 32);
            mbar_init(&hdr->res_full[i], 1);
            mbar_init(&hdr->stg_free[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, 256u);
    }
    // channel rows K .. Kpad-1 of every stage are never written by the TMA (the box has K rows): zero them once, their
    // weight columns are zero as well (0 * garbage could be NaN)
    if (a.Kpad > a.K) {
        const int pad_rows = a.Kpad - a.K;  // a whole 8-row group (K % 8 == 0, Kpad % 16 == 0): 1 KiB per column block
        const int per_stage = 2 * pad_rows * 8;  // 16-byte chunks
        for (int u = tid; u < a.stages * per_stage; u += (int)blockDim.x) {
            const int st = u / per_stage, r = u - st * per_stage;
            const int blk = r / (pad_rows * 8), c = r - blk * (pad_rows * 8);
            *reinterpret_cast<uint4 *>(smem + a.off_a + (size_t)st * a.stage_bytes + (size_t)blk * (a.Kpad * 128) + (size_t)a.K * 128 + c * 16) =
                make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (a.w_resident) p3_stage_weights(a, smem_w, n0, nrows, tid);
    pdl_sync();
    if (!a.w_resident) p3_stage_weights(a, smem_w, n0, nrows, tid);
    float *smem_sb = reinterpret_cast<float *>(smem + kP3SbOff);
    if (a.a_sb != nullptr)
        for (int k = tid; k < 2 * a.K; k += (int)blockDim.x) smem_sb[k] = a.a_sb[k];  // (scale, bias) interleaved as given
    fence_proxy_async_smem();  // weights + zero rows (generic proxy) -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    const int tile0 = blockIdx.x, tstride = gridDim.x;
    const uint32_t box_bytes_x = (uint32_t)a.K * 128u, box_bytes_o = (uint32_t)nrows * 128u;

    if (warp == 0) {
        // ================================ TMA loads ================================================================
        if (lane == 0) {
            int slot = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
                const int img = tile / a.tiles_per_image, px0 = (tile - img * a.tiles_per_image) * kP3Npx;
                const int b = it & 1;
                const uint32_t bph = (uint32_t)(it >> 1) & 1u;
                if (a.has_res) {
                    // the staging buffer is free once the store of its previous tile has read it
                    mbar_wait(&hdr->stg_free[b], bph ^ 1u);
                    p3_expect_tx(&hdr->res_full[b], 2u * box_bytes_o);
                    p3_load_box(s_stg + (uint32_t)b * a.stg_bytes, &map_res, px0, n0, img, &hdr->res_full[b]);
                    p3_load_box(s_stg + (uint32_t)b * a.stg_bytes + (uint32_t)a.Ncta * 128u, &map_res, px0 + 64, n0, img, &hdr->res_full[b]);
                }
                mbar_wait(&hdr->empty[slot], phase ^ 1u);
                p3_expect_tx(&hdr->full[slot], 2u * box_bytes_x);
                const uint32_t dst = s_a + (uint32_t)slot * a.stage_bytes;
                p3_load_box(dst, &map_x, px0, 0, img, &hdr->full[slot]);
                p3_load_box(dst + (uint32_t)a.Kpad * 128u, &map_x, px0 + 64, 0, img, &hdr->full[slot]);
                if (++slot == a.stages) { slot = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer: one thread ===================================================
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(kP3Rows, kP3Npx, /*weights K-major*/ 0, /*activations MN-major*/ 1);
            const uint64_t adesc0 = smem_desc(smem_u32(smem_w), a.w_lbo, 128, LAYOUT_NONE);
            // activations: 8-channel groups of 8 rows x 64 pixels (128 B, XOR-swizzled by the TMA), groups 1 KiB apart (SBO),
            // the second 64-pixel block Kpad * 128 B further (LBO)
            const uint64_t bdesc0 = smem_desc(s_a, (uint32_t)a.Kpad * 128u, 1024, LAYOUT_SW128);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
            const uint32_t a_kstep = (2 * a.w_lbo) >> 4;
            const int ksteps = a.Kpad >> 4;
            int slot = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
                const int as = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&hdr->tmem_empty[as], aph ^ 1u);
                mbar_wait(a.a_sb != nullptr ? &hdr->ready[slot] : &hdr->full[slot], phase);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)as * kP3Npx;
                uint32_t a_lo = a_lo0, b_lo = b_lo0 + (uint32_t)slot * (a.stage_bytes >> 4);
                for (int ks = 0; ks < ksteps; ++ks) {
                    mma_bf16_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, ks ? 1u : 0u);
                    a_lo += a_kstep;
                    b_lo += 2048u >> 4;
                }
                mma_commit(&hdr->empty[slot]);
                mma_commit(&hdr->tmem_full[as]);
                if (++slot == a.stages) { slot = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp >= kP3BnWarp0) {
        // ================================ BN + ReLU on the landed stage ==============================================
        const int bt = tid - kP3BnWarp0 * 32;
        const int units = a.K * 8;  // 16-byte chunks per 64-pixel column block; chunk u lies in channel row u / 8
        int slot = 0;
        uint32_t phase = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride) {
            mbar_wait(&hdr->full[slot], phase);
            const uint32_t base = s_a + (uint32_t)slot * a.stage_bytes;
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                const uint32_t bb = base + (uint32_t)blk * ((uint32_t)a.Kpad * 128u);
                for (int u = bt; u < units; u += kP3NumBn * 32) {
                    const float2 sb = reinterpret_cast<const float2 *>(smem_sb)[u >> 3];
                    uint4 v = p3_lds128(bb + (uint32_t)u * 16u);
                    v.x = p3_bn_relu2(v.x, sb.x, sb.y);
                    v.y = p3_bn_relu2(v.y, sb.x, sb.y);
                    v.z = p3_bn_relu2(v.z, sb.x, sb.y);
                    v.w = p3_bn_relu2(v.w, sb.x, sb.y);
                    p3_sts128(bb + (uint32_t)u * 16u, v);
                }
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->ready[slot]);
            if (++slot == a.stages) { slot = 0; phase ^= 1u; }
        }
    } else {
        // ================================ epilogue =================================================================
        // warp -> TMEM lane quarter (hardware: warp id % 4) and one of the two 64-pixel column blocks
        const int q = warp & 3, blk = (warp - kP3EpiWarp0) >> 2;
        const int row = q * 32 + lane;  // CTA-local output channel of this thread's TMEM lane
        const bool row_ok = row < nrows;
        // BatchNorm statistics of the output (a.stats): this thread owns output channel `row` of every tile's column block
        // `blk` for the whole kernel, so the two sums never leave its registers until the end
        float st1 = 0.f, st2 = 0.f;
        int it = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            const int as = it & 1, b = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            const int img = tile / a.tiles_per_image, px0 = (tile - img * a.tiles_per_image) * kP3Npx;
            const uint32_t stg = s_stg + (uint32_t)b * a.stg_bytes + (uint32_t)blk * ((uint32_t)a.Ncta * 128u) + (uint32_t)row * 128u;
            mbar_wait(&hdr->tmem_full[as], aph);
            tc_fence_after();
            if (a.has_res) mbar_wait(&hdr->res_full[b], aph);
            else mbar_wait(&hdr->stg_free[b], aph ^ 1u);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kP3Npx + blk * 64);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {  // two rounds of 32 columns
                uint32_t v[2][16];
                tmem_ld16(taddr + h * 32, v[0]);
                tmem_ld16(taddr + h * 32 + 16, v[1]);
                tmem_ld_wait();
                if (h == 1) {  // every TMEM load of this tile has landed: hand the accumulator stage back
                    tc_fence_before();
                    mbar_arrive(&hdr->tmem_empty[as]);
                }
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t *vv = &v[c >> 1][(c & 1) * 8];
                        uint4 o = make_uint4(pack_bf16x2(__uint_as_float(vv[0]), __uint_as_float(vv[1])),
                                             pack_bf16x2(__uint_as_float(vv[2]), __uint_as_float(vv[3])),
                                             pack_bf16x2(__uint_as_float(vv[4]), __uint_as_float(vv[5])),
                                             pack_bf16x2(__uint_as_float(vv[6]), __uint_as_float(vv[7])));
                        const uint32_t addr = stg + ((uint32_t)((h * 4 + c) ^ (row & 7)) << 4);
                        if (a.has_res) {
                            const uint4 r = p3_lds128(addr);
                            o = make_uint4(p3_add_bf16x2(o.x, r.x), p3_add_bf16x2(o.y, r.y), p3_add_bf16x2(o.z, r.z), p3_add_bf16x2(o.w, r.w));
                        }
                        p3_sts128(addr, o);
                        // of the stored (bf16-rounded) values, as a statistics pass over the tensor would read them; chunks
                        // beyond the plane (the half-empty last tile of an image) are not part of it
                        if (STATS && px0 + blk * 64 + (h * 4 + c) * 8 < a.HW) {
                            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = bf16_lo(ow[e]), hi = bf16_hi(ow[e]);
                                st1 += lo + hi;
                                st2 = fmaf(lo, lo, fmaf(hi, hi, st2));
                            }
                        }
                    }
                }
            }
            fence_proxy_async_smem();  // staging writes (generic proxy) -> visible to the TMA store
            asm volatile("bar.sync 1, %0;" ::"n"(kP3NumEpi * 32) : "memory");
            if (warp == kP3EpiWarp0 && lane == 0) {
                const uint32_t src = s_stg + (uint32_t)b * a.stg_bytes;
                p3_store_box(&map_out, src, px0, n0, img);
                if (px0 + 64 < a.HW) p3_store_box(&map_out, src + (uint32_t)a.Ncta * 128u, px0 + 64, n0, img);
                p3_commit();
                p3_wait_read0();  // the stores have read the staging buffer: it may be refilled
                mbar_arrive(&hdr->stg_free[b]);
            }
        }
        if (warp == kP3EpiWarp0 && lane == 0) p3_wait_all();  // global writes complete before the grid retires
        if (STATS && row_ok) {
            double *o = a.stats + ((int64_t)(n0 + row) * a.stats_splits + (blockIdx.x * 2 + blk)) * 2;
            o[0] = (double)st1;
            o[1] = (double)st2;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256u);
    }
}

typedef CUresult (*P3EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

P3EncodeFn p3_encoder() {
    static P3EncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<P3EncodeFn>(p);
    }();
    return fn;
}

// [NI, C, HW] bf16 tensor, boxes of (64 pixels, rows channels, 1 image), 128-byte swizzle
bool p3_make_map(CUtensorMap *map, const void *base, int NI, int C, int HW, int rows) {
    P3EncodeFn enc = p3_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)NI};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)rows, 1u}, estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int p3_round_up(int v, int m) { return (v + m - 1) / m * m; }

bool p3_plan(P3Args &a, dim3 *grid, size_t *smem_bytes) {
    if (a.HW % 8 != 0 || a.K % 8 != 0 || a.K > 256 || a.K < 16 || a.HW < 128) return false;
    a.Kpad = p3_round_up(a.K, 16);
    const int gy = cdiv(a.N, kP3Rows);
    a.Ncta = p3_round_up(cdiv(a.N, gy), 8);
    if (a.N % 8 != 0 || a.Ncta * gy != a.N) return false;  // every slice a whole number of 8-row groups, no ragged last slice
    a.w_lbo = (uint32_t)(kP3Rows * 16 + 16);
    a.off_w = kP3Hdr;
    a.off_a = (uint32_t)p3_round_up((int)a.off_w + (a.Kpad >> 3) * (int)a.w_lbo, 1024);
    a.stage_bytes = (uint32_t)a.Kpad * 256u;  // two 64-pixel column blocks of Kpad rows x 128 B
    a.stg_bytes = (uint32_t)p3_round_up(a.Ncta * 256, 1024);
    const int room = kP3Smem - (int)a.off_a - 2 * (int)a.stg_bytes;
    a.stages = room / (int)a.stage_bytes;
    if (a.stages > kP3MaxStages) a.stages = kP3MaxStages;
    if (a.stages < 2) return false;
    a.off_stg = a.off_a + (uint32_t)a.stages * a.stage_bytes;
    a.tiles_per_image = cdiv(a.HW, kP3Npx);
    a.total_tiles = a.NI * a.tiles_per_image;
    int gx = sm_count() / gy;
    if (gx < 1) gx = 1;
    if (gx > a.total_tiles) gx = a.total_tiles;
    *grid = dim3((unsigned)gx, (unsigned)gy, 1);
    *smem_bytes = (size_t)a.off_stg + 2 * (size_t)a.stg_bytes;
    return true;
}

std::atomic<int> g_p3_enabled{1};

}  // namespace

void pw3_set_enabled(int on) { g_p3_enabled.store(on ? 1 : 0); }
bool pw3_enabled() { return g_p3_enabled.load(std::memory_order_relaxed) != 0; }

// 1 when the geometry (and the pointers' alignment) can run on k_pw3
bool pw3_supported(const void *x, const void *out, const void *res, int NI, int K, int N, int HW) {
    if (!g_p3_enabled.load(std::memory_order_relaxed)) return false;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 15) return false;
    P3Args a{};
    a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem;
    return p3_plan(a, &grid, &smem) && p3_encoder() != nullptr;
}

// BatchNorm statistics in the epilogue are worth it on this schedule (an epilogue thread owns a channel row: thread-local sums)
bool pw3_stats_preferred(int NI, int K, int N, int HW) {
    if (!g_p3_enabled.load(std::memory_order_relaxed) || p3_encoder() == nullptr) return false;
    P3Args a{};
    a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem;
    return p3_plan(a, &grid, &smem);
}

int pw3_forward(const void *x, const void *w, int w_dt, const void *res, void *out, int NI, int K, int N, int HW, const float *a_sb,
                cudaStream_t s, double *stats, size_t stats_bytes, int *stats_splits) {
    P3Args a{};
    a.a_sb = a_sb;
    a.w = w; a.w_f32 = (w_dt & ~RB_W_RESIDENT) == RB_F32; a.w_resident = (w_dt & RB_W_RESIDENT) != 0; a.has_res = res != nullptr;
    a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem_bytes = 0;
    if (!p3_plan(a, &grid, &smem_bytes)) return fail(RB_ERR_UNSUPPORTED, "pw3: geometry not supported");
    if (reinterpret_cast<uintptr_t>(w) & 15) return fail(RB_ERR_INVALID_ARGUMENT, "pw3: weight must be 16-byte aligned");
    if (stats != nullptr) {
        a.stats = stats;
        a.stats_splits = 2 * (int)grid.x;
        if (stats_bytes < (size_t)N * a.stats_splits * 2 * sizeof(double))
            return fail(RB_ERR_WORKSPACE, "pw3: statistics need %zu bytes", (size_t)N * a.stats_splits * 2 * sizeof(double));
        if (stats_splits) *stats_splits = a.stats_splits;
    }
    CUtensorMap mx, mo, mr;
    if (!p3_make_map(&mx, x, NI, K, HW, K) || !p3_make_map(&mo, out, NI, N, HW, a.Ncta) ||
        !p3_make_map(&mr, res ? res : out, NI, N, HW, a.Ncta))
        return fail(RB_ERR_CUDA, "pw3: cuTensorMapEncodeTiled failed");
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_pw3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pw3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP3Smem);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_pw3): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    if (stats != nullptr) launch_kernel(k_pw3<true>, grid, dim3(a_sb ? kP3ThreadsBn : kP3Threads), smem_bytes, s, mx, mo, mr, a);
    else launch_kernel(k_pw3<false>, grid, dim3(a_sb ? kP3ThreadsBn : kP3Threads), smem_bytes, s, mx, mo, mr, a);
    return launched("k_pw3");
}

}  // namespace rb

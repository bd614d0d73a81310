// Weight gradient of the 1x1 convolutions, second schedule: both operands arrive through the tensor-map (tiled) mode of
// the TMA unit, straight into the K-major SWIZZLE_128B UMMA layout -- no register round trip (k_pw_wgrad, csrc/pw_conv.cu,
// moves every operand byte LDG -> registers -> STS with 14 producer warps and runs at 0.19 of the HBM peak in a step).
//
//     dW[m, n] = sum_{i, p} G[i, m, p] * A(i, n, p)        A = x  or  relu(x * scale[n] + bias[n])   (bn1 -> relu -> conv2)
//
// The reduction runs over pixels, which are contiguous in NCHW: a box of (64 pixels x rows channels x 1 image) written
// with the 128-byte swizzle IS the canonical K-major operand block (row = channel = 128 bytes of K, 8-row atoms of 1 KiB).
// Pixels beyond the plane are zero-filled by the TMA unit, i.e. they add nothing to the sums.
//
// Only maps whose row pitch is a multiple of 16 bytes qualify (HW % 8 == 0: 112x112, 56x56, 28x28).  On 14x14 (392-byte
// rows) and 7x7 (98-byte rows) every other channel row starts 8 / 2 bytes off a 16-byte boundary; splitting the channels
// into residue classes with one tensor map each satisfies the encoder (base and strides 16-byte multiples) but the box
// START must be 16-byte aligned as well -- a pixel coordinate 4 elements into the row faults with "illegal instruction"
// (measured, gpurun_out/probe_wg3.log) -- so those maps stay on k_pw_wgrad.
//
//   warp 0      TMA loads into a ring of stages (G rows | A rows), one elected thread
//   warp 1      one elected thread issues tcgen05.mma (128 x sub_n x 16), accumulators [Mt x Nc] fp32 in tensor memory
//   warps 2-9   BN+ReLU producer only: relu(x*s+b) in place on the A rows of a landed stage (16-byte shared-memory
//               vectors), then `ready`; after the last stage: TMEM -> fp32 partial slice of this pixel split
//   grid        (pixel splits, N blocks, M blocks); k_wg_reduce sums the slices in a fixed order (deterministic)
#include <cuda.h>

#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int kW3Warps = 10;
constexpr int kW3Threads = kW3Warps * 32;
constexpr int kW3EpiWarp0 = 2, kW3NumEpi = 8;
constexpr int kW3MaxStages = 8;
constexpr int kW3Smem = 227 * 1024;
constexpr int kW3Hdr = 1024;
constexpr int kW3MaxBurst = 4;
constexpr int kW3Chunk = 64;             // pixels per stage = one swizzle atom of K
constexpr int kW3TileBytes = 128 * 128;  // one 128-row operand tile

struct W3Maps {
    CUtensorMap g, x;
};

struct W3Args {
    float *partial;     // [splits, N, M]
    const float *x_sb;  // BN+ReLU producer: (scale, bias) pairs [N, 2]
    int NI, M, N, HW;
    int Mb, Mt, Ncr, Nc, n_sub, sub_n, stages, tmem_cols;
    int boxM, boxN, nboxM, nboxN;
    int cpi, total_chunks, chunks_per_split, burst;
    uint32_t off_sb, off_stage, stage_bytes, b_off;
};

struct W3Hdr {
    uint64_t full[kW3MaxStages], empty[kW3MaxStages], ready[kW3MaxStages], tmem_full;
    uint32_t tmem_base;
};
static_assert(sizeof(W3Hdr) <= kW3Hdr, "header");

__device__ __forceinline__ void w3_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w3_load_box(uint32_t dst, const CUtensorMap *map, int px, int ch, int img, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(map), "r"(px), "r"(ch), "r"(img), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 w3_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void w3_sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t w3_bn_relu2(uint32_t w, float sc, float bi) {
    return pack_bf16x2(fmaxf(fmaf(bf16_lo(w), sc, bi), 0.f), fmaxf(fmaf(bf16_hi(w), sc, bi), 0.f));
}
template <bool BN>
__global__ void __launch_bounds__(kW3Threads, 1) k_wg3(const __grid_constant__ W3Maps maps, const W3Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    W3Hdr *hdr = reinterpret_cast<W3Hdr *>(smem);
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);
    const uint32_t s_stage0 = smem_u32(smem + a.off_stage);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.z * a.Mb, n0 = blockIdx.y * a.Ncr;
    const int c_begin = blockIdx.x * a.chunks_per_split;
    const int c_end = min(a.total_chunks, c_begin + a.chunks_per_split);

    if (tid == 0) {
        for (int i = 0; i < kW3MaxStages; ++i) {
            mbar_init(&hdr->full[i], 1);
            mbar_init(&hdr->empty[i], 1);
            mbar_init(&hdr->ready[i], kW3NumEpi);
        }
        mbar_init(&hdr->tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, (uint32_t)a.tmem_cols);
    }
    pdl_sync();  // barriers and tensor memory are set up while the previous kernel drains
    if (BN)
        for (int k = tid; k < a.Ncr; k += kW3Threads) {
            smem_sb[k] = a.x_sb[2 * (n0 + k)];
            smem_sb[a.Ncr + k] = a.x_sb[2 * (n0 + k) + 1];
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp == 0) {
        // ================================ TMA loads ================================================================
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(a.Mb + a.Ncr) * 128u;
            int slot = 0;
            uint32_t phase = 0;
            int img = c_begin / a.cpi, pc = c_begin - img * a.cpi;
            // `burst` consecutive chunks are issued together, operand by operand: G(q), G(q+1), .., A(q), A(q+1), .. so that
            // the requests for adjacent 128-byte pieces of a channel row reach the memory system back to back
            for (int q = c_begin; q < c_end;) {
                const int nb = min(a.burst, c_end - q);
                uint32_t dst[kW3MaxBurst];
                int p0[kW3MaxBurst], im[kW3MaxBurst];
                uint64_t *bar[kW3MaxBurst];
#pragma unroll
                for (int u = 0; u < kW3MaxBurst; ++u)
                    if (u < nb) {
                        mbar_wait(&hdr->empty[slot], phase ^ 1u);
                        w3_expect_tx(&hdr->full[slot], tx);
                        dst[u] = s_stage0 + (uint32_t)slot * a.stage_bytes;
                        bar[u] = &hdr->full[slot];
                        p0[u] = pc * kW3Chunk;
                        im[u] = img;
                        if (++pc == a.cpi) { pc = 0; ++img; }
                        if (++slot == a.stages) { slot = 0; phase ^= 1u; }
                    }
                for (int j = 0; j < a.nboxM; ++j)
#pragma unroll
                    for (int u = 0; u < kW3MaxBurst; ++u)
                        if (u < nb) w3_load_box(dst[u] + (uint32_t)(j * a.boxM) * 128u, &maps.g, p0[u], m0 + j * a.boxM, im[u], bar[u]);
                for (int j = 0; j < a.nboxN; ++j)
#pragma unroll
                    for (int u = 0; u < kW3MaxBurst; ++u)
                        if (u < nb) w3_load_box(dst[u] + a.b_off + (uint32_t)(j * a.boxN) * 128u, &maps.x, p0[u], n0 + j * a.boxN, im[u], bar[u]);
                q += nb;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer: one thread ===================================================
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(128, a.sub_n, 0, 0);
            const uint64_t desc0 = smem_desc(s_stage0, 16, 1024, LAYOUT_SW128);
            const uint32_t d_hi = (uint32_t)(desc0 >> 32), d_lo0 = (uint32_t)desc0;
            const uint32_t b_off = a.b_off >> 4, sub_step = (uint32_t)((a.sub_n * 128) >> 4);
            int slot = 0;
            uint32_t phase = 0, acc = 0u;
            int pc = c_begin % a.cpi;
            for (int q = c_begin; q < c_end; ++q) {
                const int kvalid = min(kW3Chunk, a.HW - pc * kW3Chunk);
                const int ksteps = (kvalid + 15) >> 4;
                if (++pc == a.cpi) pc = 0;
                mbar_wait(BN ? &hdr->ready[slot] : &hdr->full[slot], phase);
                tc_fence_after();
                const uint32_t lo_s = d_lo0 + (uint32_t)((slot * a.stage_bytes) >> 4);
                for (int ks = 0; ks < ksteps; ++ks) {
                    uint32_t a_lo = lo_s + (uint32_t)(ks * 2);  // 32 bytes per K step inside the swizzle atom
                    uint32_t tcol = tmem_base;
                    for (int mt = 0; mt < a.Mt; ++mt, a_lo += kW3TileBytes >> 4) {
                        uint32_t b_lo = lo_s + b_off + (uint32_t)(ks * 2);
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
        const int et = tid - kW3EpiWarp0 * 32;
        if (BN) {
            // ============================ relu(x*s+b) in place on the landed A rows =================================
            const int units = a.Ncr * 8;  // 16-byte chunks; chunk u lives in row u / 8 whatever the swizzle
            int slot = 0;
            uint32_t phase = 0;
            int pc = c_begin % a.cpi;
            for (int q = c_begin; q < c_end; ++q) {
                const int kvalid = min(kW3Chunk, a.HW - pc * kW3Chunk);
                const int nch = ((kvalid + 15) >> 4) * 2;  // logical chunks the MMAs of this stage read
                if (++pc == a.cpi) pc = 0;
                mbar_wait(&hdr->full[slot], phase);
                const uint32_t base = s_stage0 + (uint32_t)slot * a.stage_bytes + a.b_off;
                for (int u = et; u < units; u += kW3NumEpi * 32) {
                    const int row = u >> 3;
                    if (((u ^ row) & 7) < nch) {
                        const float sc = smem_sb[row], bi = smem_sb[a.Ncr + row];
                        uint4 v = w3_lds128(base + (uint32_t)u * 16u);
                        v.x = w3_bn_relu2(v.x, sc, bi);
                        v.y = w3_bn_relu2(v.y, sc, bi);
                        v.z = w3_bn_relu2(v.z, sc, bi);
                        v.w = w3_bn_relu2(v.w, sc, bi);
                        w3_sts128(base + (uint32_t)u * 16u, v);
                    }
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr->ready[slot]);
                if (++slot == a.stages) { slot = 0; phase ^= 1u; }
            }
        }
        // ================================ epilogue: fp32 partial slice of this pixel split ==========================
        const int q4 = warp & 3, half = (warp - kW3EpiWarp0) >> 2, row = q4 * 32 + lane;
        mbar_wait(&hdr->tmem_full, 0);
        tc_fence_after();
        float *dst = a.partial + (int64_t)blockIdx.x * a.M * a.N;
        const int nchunks = (a.Ncr + 15) >> 4;
        for (int mt = 0; mt < a.Mt; ++mt) {
            const int ml = mt * 128 + row;
            const bool valid = ml < a.Mb;
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(mt * a.Nc);
            for (int ci = half; ci < nchunks; ci += 2) {
                const int c0 = ci * 16;
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                    // slice layout [N, M]: for a fixed column the 32 lanes write 128 contiguous bytes
                    float *o = dst + (int64_t)(n0 + c0) * a.M + m0 + ml;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < a.Ncr) o[(int64_t)j * a.M] = __uint_as_float(v[j]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

// partial: [splits][N][M] (M contiguous); out: [M][N].  One CTA = 32 consecutive outputs x 8 warps, warp g sums the slices
// g, g+8, .. (coalesced 128-byte reads, 8 x more loads in flight than one thread per output walking all the slices: that
// version took 10 us on 74 slices of 288 x 288), then the 8 group sums are added in a fixed order: deterministic.
constexpr int kRedGroups = 8;
__global__ void __launch_bounds__(kRedGroups * 32) k_wg_reduce(const float *__restrict__ partial, float *__restrict__ out, int splits, int M, int N) {
    pdl_sync();
    __shared__ float red[kRedGroups][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int64_t count = (int64_t)M * N;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < count) {
        const float *p = partial + i;
        int k = grp;
        for (; k + 3 * kRedGroups < splits; k += 4 * kRedGroups) {
            const float v0 = __ldg(p + (int64_t)k * count), v1 = __ldg(p + (int64_t)(k + kRedGroups) * count);
            const float v2 = __ldg(p + (int64_t)(k + 2 * kRedGroups) * count), v3 = __ldg(p + (int64_t)(k + 3 * kRedGroups) * count);
            s += v0;
            s += v1;
            s += v2;
            s += v3;
        }
        for (; k < splits; k += kRedGroups) s += __ldg(p + (int64_t)k * count);
    }
    red[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && i < count) {
        float t = red[0][lane];
#pragma unroll
        for (int g = 1; g < kRedGroups; ++g) t += red[g][lane];
        const int n = (int)(i / M), m = (int)(i - (int64_t)n * M);
        out[(int64_t)m * N + n] = t;
    }
}

typedef CUresult (*W3EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

W3EncodeFn w3_encoder() {
    static W3EncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<W3EncodeFn>(p);
    }();
    return fn;
}

// [NI, C, HW] bf16 tensor, boxes of (64 pixels, rows channels, 1 image), 128-byte swizzle
// measured at 72 ch x 112x112 / 56x56 (gpurun_out/wg3_bench.log): burst 1 -> 2 -> 4 = 0.316 -> 0.230 -> 0.233 ms / 0.110 -> 0.109 -> 0.091 ms;
// 256-byte L2 promotion and a shallower ring change nothing
std::atomic<int> g_w3_l2_256{0}, g_w3_burst{4}, g_w3_max_stages{kW3MaxStages};

bool w3_make_map(CUtensorMap *map, const void *base, int NI, int C, int HW, int rows) {
    W3EncodeFn enc = w3_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)NI};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
    cuuint32_t box[3] = {(cuuint32_t)kW3Chunk, (cuuint32_t)rows, 1u}, estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, g_w3_l2_256.load() ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int w3_round_up(int v, int m) { return (v + m - 1) / m * m; }
bool w3_plan(W3Args &a, dim3 *grid, size_t *smem_bytes) {
    if (a.HW <= 0 || a.NI <= 0) return false;
    if (a.HW % 8 != 0 || a.M % 8 != 0 || a.N % 8 != 0) return false;  // 16-byte row pitch; whole 8-row swizzle atoms
    a.off_sb = kW3Hdr;
    int best = 1 << 30, bmb = 0, bnb = 0;
    for (int mb = 1; mb <= 8; ++mb)
        for (int nb = 1; nb <= 8; ++nb) {
            if (a.M % mb != 0 || a.N % nb != 0) continue;
            const int Mb = a.M / mb, Ncr = a.N / nb;
            if (Mb % 8 != 0 || Ncr % 8 != 0) continue;
            const int Mt = cdiv(Mb, 128);
            int Nc = w3_round_up(Ncr, 16);
            const int ns = cdiv(Nc, 256);
            Nc = w3_round_up(Nc, 16 * ns);
            if (Mt > 4 || Mt * Nc > 512) continue;
            if (Mb % (cdiv(Mb, 256)) != 0 || (Mb / cdiv(Mb, 256)) % 8 != 0) continue;  // boxes of <= 256 rows, whole atoms
            if (Ncr % (cdiv(Ncr, 256)) != 0 || (Ncr / cdiv(Ncr, 256)) % 8 != 0) continue;
            const int off_stage = w3_round_up(kW3Hdr + 2 * Ncr * 4, 1024);
            const int stage = Mt * kW3TileBytes + w3_round_up(Nc * 128, 1024);
            if ((kW3Smem - off_stage) / stage < 2) continue;
            const int cost = (nb + mb) * 16 + mb * nb;  // operand re-reads first, CTA count second
            if (cost < best) { best = cost; bmb = mb; bnb = nb; }
        }
    if (!bmb) return false;
    a.Mb = a.M / bmb;
    a.Mt = cdiv(a.Mb, 128);
    a.Ncr = a.N / bnb;
    a.Nc = w3_round_up(a.Ncr, 16);
    a.n_sub = cdiv(a.Nc, 256);
    a.Nc = w3_round_up(a.Nc, 16 * a.n_sub);
    a.sub_n = a.Nc / a.n_sub;
    a.nboxM = cdiv(a.Mb, 256);
    a.nboxN = cdiv(a.Ncr, 256);
    a.boxM = a.Mb / a.nboxM;
    a.boxN = a.Ncr / a.nboxN;
    a.off_stage = (uint32_t)w3_round_up(kW3Hdr + 2 * a.Ncr * 4, 1024);
    a.b_off = (uint32_t)(a.Mt * kW3TileBytes);
    a.stage_bytes = a.b_off + (uint32_t)w3_round_up(a.Nc * 128, 1024);
    a.stages = (kW3Smem - (int)a.off_stage) / (int)a.stage_bytes;
    if (a.stages > g_w3_max_stages.load()) a.stages = g_w3_max_stages.load();
    a.burst = g_w3_burst.load();
    if (a.burst > a.stages / 2) a.burst = a.stages / 2;  // the consumer must be able to run while a burst is being issued
    if (a.burst < 1) a.burst = 1;
    int cols = 32;
    while (cols < a.Mt * a.Nc) cols <<= 1;
    a.tmem_cols = cols;
    a.cpi = cdiv(a.HW, kW3Chunk);
    a.total_chunks = a.NI * a.cpi;
    int want = sm_count() / (bmb * bnb);
    if (want < 1) want = 1;
    if (want > a.total_chunks) want = a.total_chunks;
    a.chunks_per_split = cdiv(a.total_chunks, want);
    const int splits = cdiv(a.total_chunks, a.chunks_per_split);
    *grid = dim3((unsigned)splits, (unsigned)bnb, (unsigned)bmb);
    *smem_bytes = (size_t)a.off_stage + (size_t)a.stages * a.stage_bytes;
    return true;
}

}  // namespace

// burst: chunks issued together (1..4); l2_256: 256-byte L2 promotion in the tensor maps; max_stages: ring depth cap (2..8)
void wg3_set_tuning(int burst, int l2_256, int max_stages) {
    g_w3_burst.store(burst < 1 ? 1 : (burst > kW3MaxBurst ? kW3MaxBurst : burst));
    g_w3_l2_256.store(l2_256 ? 1 : 0);
    g_w3_max_stages.store(max_stages < 2 ? kW3MaxStages : (max_stages > kW3MaxStages ? kW3MaxStages : max_stages));
}

// fixed-order sum of the fp32 partial slices of either weight-gradient kernel
int wg_reduce(const float *partial, float *dw, int splits, int M, int N, cudaStream_t s) {
    const int64_t count = (int64_t)M * N;
    launch_kernel(k_wg_reduce, dim3((unsigned)cdiv64(count, 32)), dim3(kRedGroups * 32), 0, s, partial, dw, splits, M, N);
    return launched("k_wg_reduce");
}

bool pw3_enabled();  // pw_conv3.cu: rb_pw_conv_tma_set_enabled switches both tensor-map schedules

// fp32 [splits, M, N] scratch of the tensor-map schedule; 0 when the geometry cannot run on it
size_t wg3_workspace(int NI, int M, int N, int HW) {
    W3Args a{};
    a.NI = NI; a.M = M; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem;
    if (!w3_plan(a, &grid, &smem) || w3_encoder() == nullptr) return 0;
    return (size_t)grid.x * M * N * sizeof(float);
}

bool wg3_supported(const void *g, const void *x, int NI, int M, int N, int HW) {
    if (!pw3_enabled()) return false;
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x)) & 15) return false;
    return wg3_workspace(NI, M, N, HW) != 0;
}

int wg3_run(const void *g, const void *x, float *dw, int NI, int M, int N, int HW, const float *x_sb, void *workspace, cudaStream_t s) {
    W3Args a{};
    a.partial = (float *)workspace; a.x_sb = x_sb;
    a.NI = NI; a.M = M; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem_bytes = 0;
    if (!w3_plan(a, &grid, &smem_bytes)) return fail(RB_ERR_UNSUPPORTED, "wg3: geometry not supported");
    W3Maps maps;
    if (!w3_make_map(&maps.g, g, NI, M, HW, a.boxM) || !w3_make_map(&maps.x, x, NI, N, HW, a.boxN))
        return fail(RB_ERR_CUDA, "wg3: cuTensorMapEncodeTiled failed");
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_wg3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kW3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wg3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kW3Smem);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_wg3): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    if (x_sb) launch_kernel(k_wg3<true>, grid, dim3(kW3Threads), smem_bytes, s, maps, a);
    else launch_kernel(k_wg3<false>, grid, dim3(kW3Threads), smem_bytes, s, maps, a);
    if (int rc = launched("k_wg3")) return rc;
    return wg_reduce((const float *)a.partial, dw, (int)grid.x, M, N, s);
}

}  // namespace rb

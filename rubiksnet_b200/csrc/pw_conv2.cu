// Pointwise (1x1) convolution of RubiksShiftBlock, second generation: every byte of global traffic moves through the
// TMA unit (cp.async.bulk, SASS UBLKCP), warps only touch shared memory and tensor memory.
//
//     out[i, n, p] = sum_k W[n, k] * A(i, k, p)  (+ residual[i, n, p]),   A = x  or  relu(x * scale[k] + bias[k])
//
// (conv2 / conv3 / shortcut of rubiksnet/backbone.py:123-135 and, on the transposed weight image, their input gradients.)
//
// Why a second kernel (measured on B200, profiles/r01g_*): the first-generation k_pw_conv keeps the WHOLE [N x K] weight
// block resident (166 KiB at 288 channels), which leaves a 44 KiB activation ring -- one tile deep -- and moves activations
// with per-thread LDG/STG, so its producers and its epilogue (15 GB/s per SM), not HBM, set the pace.  Here:
//   * output channels are split over grid.y so that a CTA owns <= 128 of them: ONE M tile, a weight slice of <= 112 KiB that
//     arrives as a single bulk copy of a pre-packed shared-memory image (rb_pw2_weight_pack), and room for deep rings;
//   * a tile is S segments of L consecutive pixels (S whole images when a map has <= 224 pixels -- 14x14: one image, 7x7: four
//     -- else L = a 16-byte-multiple divisor of the map), N_mma = round_up(S*L, 16) <= 256 accumulator columns, double-buffered
//     in tensor memory;
//   * raw ring: the producer warp fetches the [channels x L] rows of a K chunk exactly as they lie in NCHW (whole-image
//     chunks are ONE contiguous bulk copy) -- alignment rules of cp.async.bulk (16 bytes) are met by construction;
//   * relayout warps turn a raw stage into the MN-major SWIZZLE_128B UMMA operand (and apply bn1+relu on the way): shared ->
//     shared with 29-cycle loads instead of 600-cycle global loads, so 4 warps with a handful of registers keep up;
//   * the epilogue writes bf16 rows into a staging block laid out exactly like the output rows in global memory, the
//     residual block is bulk-loaded into the same staging buffer beforehand and added in place, and the tile leaves with
//     bulk stores (one per image for whole-image tiles).
// Arithmetic is the same as k_pw_conv: bf16 operands, fp32 accumulation in TMEM, result rounded to bf16, `+= shortcut` on
// the rounded value.
#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

// warp roles: 0 MMA issuer (+ TMEM allocation), 1 TMA loads, 2 TMA stores, 3 idle (keeps the epilogue warps aligned with the
// TMEM lane quarters: warp % 4), 4-11 relayout, 12-19 epilogue
constexpr int kP2Warps = 20;
constexpr int kP2Threads = kP2Warps * 32;  // 640
constexpr int kP2MmaWarp = 0, kP2TmaWarp = 1, kP2StoreWarp = 2;
constexpr int kP2RelWarp0 = 4, kP2NumRel = 8, kP2RelGroups = 2, kP2RelPerGroup = kP2NumRel / kP2RelGroups;
constexpr int kP2EpiWarp0 = 12, kP2NumEpi = 8;
constexpr int kP2MaxRing = 8;
constexpr int kP2Smem = 227 * 1024;
constexpr int kP2Hdr = 512;
constexpr int kP2MaxRows = 128;

struct P2Args {
    const __nv_bfloat16 *x;      // [NI, K, HW]
    const unsigned char *wimg;   // packed weight image: gy slices of w_bytes
    const __nv_bfloat16 *res;    // [NI, N, HW] or null
    __nv_bfloat16 *out;          // [NI, N, HW]
    const float *a_sb;           // bn+relu producer: (scale, bias) pairs [K, 2] or null
    int NI, K, N, HW;
    int Kpad, Ncta, gy;
    int L, S, tpi, caseA;        // segment length, segments per tile, tiles per image (case B), whole-image segments?
    int Nmma, atoms, kc, G, k_stages, raw_stages, op_stages, stg_bufs, total_tiles;
    int V;                       // elements per relayout piece (8 / 4 / 2 / 1)
    uint32_t wait_ns;            // suspend-time hint of the mbarrier waits (0 = hardware default)
    uint32_t ppr, ppr_mul, ppr_shr;  // pieces per row and its exact-division constants
    uint32_t w_lbo, w_bytes, off_sb, off_w, off_raw, off_op, off_stg;
    uint32_t raw_stage_bytes, op_stage_bytes, stg_buf_bytes, stg_pitch, stg_seg_stride;
#ifdef RB_DEBUG_TRACE
    unsigned long long *trace;  // debug builds only: 128 globaltimer stamps per CTA (tools/trace_pw.py --v2)
    int dbg;                    // debug builds only: skip work to find the critical path (1 relayout, 2 MMA, 4 epilogue, 8 stores, 16 raw loads)
#endif
};

#ifdef RB_DEBUG_TRACE
__device__ __forceinline__ unsigned long long p2_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define P2_TRACE(cond, slot) do { if (a.trace && (cond)) a.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 128 + (slot)] = p2_gtime(); } while (0)
#define P2_DBG(bit) ((a.dbg & (bit)) != 0)
#else
#define P2_TRACE(cond, slot) do { } while (0)
#define P2_DBG(bit) false
#endif

struct P2Hdr {
    uint64_t raw_full[kP2MaxRing], raw_empty[kP2MaxRing], op_full[kP2MaxRing], op_empty[kP2MaxRing];
    uint64_t tmem_full[2], tmem_empty[2], res_full[2], stg_ready[2], stg_empty[2], w_full;
    uint32_t tmem_base;
};
static_assert(sizeof(P2Hdr) <= kP2Hdr, "header");

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Whole-warp wait with ONE polling lane: mbarrier.try_wait goes through the shared-memory pipeline like an atomic, and 32
// lanes polling the same word are 32 serialised transactions -- with a dozen waiting warps that traffic alone slowed the
// relayout warps' LDS/STS and the TMA unit's writes by an order of magnitude (profiles/r02e_trace_l3.log).  Lane 0 polls,
// __syncwarp() releases the others and orders their subsequent reads after lane 0's acquire.
// mbarrier.try_wait with an explicit suspend-time hint (ns).  Measured (profiles/r02i_dbg_l3.log): pure polling
// (test_wait) floods the shared-memory pipeline -- every fence and LDS/STS of the working warps gets several times slower --
// while try_wait without a hint lets a blocked thread sleep for a hardware-chosen quantum, which in a ring of actors that
// wait on each other costs ~0.7 us per pipeline stage with no work at all.
__device__ __forceinline__ void p2_spin(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 0) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    if (hint_ns == 0) {
        mbar_wait(bar, parity);
        return;
    }
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int lane, uint32_t hint_ns = 0) {
    if (lane == 0) p2_spin(bar, parity, hint_ns);
    __syncwarp();
}
// global -> shared bulk copy (TMA, 1-D); completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy; completion through the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (uint32_t)v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t bnrelu2(uint32_t w, float sc, float bi) {
    return pack_bf16x2(fmaxf(fmaf(bf16_lo(w), sc, bi), 0.f), fmaxf(fmaf(bf16_hi(w), sc, bi), 0.f));
}

// tile -> first image, pixel offset inside the image, number of valid segments
__device__ __forceinline__ void p2_tile(const P2Args &a, int tile, int &img0, int &p0, int &nseg) {
    if (a.caseA) {
        img0 = tile * a.S;
        p0 = 0;
        nseg = min(a.S, a.NI - img0);
    } else {
        img0 = tile / a.tpi;
        p0 = (tile - img0 * a.tpi) * a.L;
        nseg = 1;
    }
}

// Operand layout: tile column c of channel kl = 8 g + r lives at
//     (atom(c) * G + g) * 1024 + r * 128 + ((chunk(c) ^ r) << 4) + sub(c),   atom = c / 64, chunk = (c % 64) / 8, sub = 2 (c % 8)
__device__ __forceinline__ uint32_t p2_col_part(uint32_t c, uint32_t G) { return (((c >> 6) * G) << 10) + ((c & 7u) << 1); }
__device__ __forceinline__ uint32_t p2_col_chunk(uint32_t c) { return (c & 63u) >> 3; }

// Relayout of one raw stage into the UMMA operand.  A work unit is (segment, 8-channel group, 32 pieces along the row); the
// relayout warps take units round-robin.  Inside a unit a lane owns ONE piece column and walks the 8 rows of the group with
// compile-time row indices: 8 independent shared-memory loads, then 8 stores whose addresses differ by constants
// (r * 128 and an XOR of the chunk index with r).  History (profiles/r02c_trace_l3.log): one piece per thread and iteration
// with run-time row arithmetic made the relayout warps -- single warps issuing dependent ALU chains at ~0.17 instructions
// per cycle -- the slowest stage of the pipeline (1.2 us per 32-channel stage against 0.1 us of MMA).
//   MODE 16 / 8: pieces of 16 / 8 bytes (L % 8 == 0 / L % 4 == 0): source and destination pieces are equally aligned.
//   MODE 4: any L (7x7 maps: rows of 49 elements start on odd elements): 4-byte destination words, fetched as two aligned
//           words + funnel shift whatever the source alignment; row ends are written as 2-byte halves so that the
//           neighbouring segment's columns in the same operand row are never touched.
// K % 8 == 0 (p2_plan), so an 8-channel group is either all real channels or all padding (zeros).
// units of a FULL stage (S segments x G groups x passes) owned by one relayout warp, decoded once per kernel: byte 0 =
// segment, byte 1 = group, byte 2 = pass.  Partial stages / tiles skip the units they do not have.
constexpr int kP2MaxUnits = 8;
struct P2Units {
    int n;
    uint32_t u[kP2MaxUnits];
};
__device__ __forceinline__ int p2_passes(const P2Args &a, int mode) {
    if (mode == 16 || mode == 8) return (int)(((uint32_t)a.L / (uint32_t)(mode / 2) + 31u) >> 5);
    return (int)((((uint32_t)a.L >> 1) + 2u + 31u) >> 5);
}
__device__ __forceinline__ P2Units p2_make_units(const P2Args &a, int mode, int rw) {
    P2Units pu;
    pu.n = 0;
    const int passes = p2_passes(a, mode), total = a.S * a.G * passes;
    for (int u = rw; u < total && pu.n < kP2MaxUnits; u += kP2RelPerGroup) {
        const int pass = u % passes, t = u / passes, g = t % a.G, sg = t / a.G;
        pu.u[pu.n++] = (uint32_t)sg | ((uint32_t)g << 8) | ((uint32_t)pass << 16);
    }
    return pu;
}

template <int MODE, bool BN>
__device__ __forceinline__ void p2_relayout_stage(const P2Args &a, uint32_t raw, uint32_t op, const float *sb, int kbase,
                                                  int rows_real, int rows_pad, int nseg, const P2Units &pu, int lane) {
    const uint32_t L = (uint32_t)a.L, Lb = L * 2u, G = (uint32_t)a.G;
    const int groups = rows_pad >> 3;
    if (MODE == 16 || MODE == 8) {
        constexpr uint32_t EPP = MODE / 2;  // elements per piece
        const uint32_t ppr = L / EPP;
#pragma unroll
        for (int iu = 0; iu < kP2MaxUnits; ++iu) {
            if (iu >= pu.n) break;
            const int sg = (int)(pu.u[iu] & 0xffu), g = (int)((pu.u[iu] >> 8) & 0xffu), pass = (int)(pu.u[iu] >> 16);
            if (sg >= nseg || g >= groups) continue;
            const uint32_t pc = (uint32_t)(lane + 32 * pass);
            if (pc >= ppr) continue;
            const uint32_t c = (uint32_t)sg * L + pc * EPP;
            const uint32_t chunk = p2_col_chunk(c);
            const uint32_t dst = op + p2_col_part(c, G) + ((uint32_t)g << 10);
            const uint32_t src = raw + ((uint32_t)sg * (uint32_t)a.kc + (uint32_t)g * 8u) * Lb + pc * MODE;
            const bool real = g * 8 < rows_real;
            uint4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = make_uint4(0u, 0u, 0u, 0u);
                if (real) {
                    if (MODE == 16) v[r] = lds128(src + (uint32_t)r * Lb);
                    else { const uint2 t2 = lds64(src + (uint32_t)r * Lb); v[r].x = t2.x; v[r].y = t2.y; }
                }
            }
            if (BN && real) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float sc = sb[kbase + g * 8 + r], bi = sb[a.Kpad + kbase + g * 8 + r];
                    v[r].x = bnrelu2(v[r].x, sc, bi); v[r].y = bnrelu2(v[r].y, sc, bi);
                    if (MODE == 16) { v[r].z = bnrelu2(v[r].z, sc, bi); v[r].w = bnrelu2(v[r].w, sc, bi); }
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t d = dst + ((uint32_t)r << 7) + ((chunk ^ (uint32_t)r) << 4);
                if (MODE == 16) sts128(d, v[r]); else sts64(d, make_uint2(v[r].x, v[r].y));
            }
        }
    } else {
#pragma unroll
        for (int iu = 0; iu < kP2MaxUnits; ++iu) {
            if (iu >= pu.n) break;
            const int sg = (int)(pu.u[iu] & 0xffu), g = (int)((pu.u[iu] >> 8) & 0xffu), pass = (int)(pu.u[iu] >> 16);
            if (sg >= nseg || g >= groups) continue;
            const uint32_t c0 = (uint32_t)sg * L, u0 = c0 >> 1;
            const uint32_t nwords = ((c0 + L - 1u) >> 1) - u0 + 1u;
            const uint32_t w = (uint32_t)(lane + 32 * pass);
            if (w >= nwords) continue;
            const uint32_t c = 2u * (u0 + w);
            const int p_lo = (int)c - (int)c0;  // -1 for the first word of a segment that starts on an odd column
            const bool ok_lo = p_lo >= 0, ok_hi = p_lo + 1 < (int)L;
            const uint32_t chunk = p2_col_chunk(c);
            const uint32_t dst = op + p2_col_part(c, G) + ((uint32_t)g << 10);
            // byte address of source column p_lo in row r = 0 of the group (row r: + r * Lb); may point 2 bytes in front of
            // the stage (p_lo = -1 in its first row): still inside this CTA's shared memory, and the value is not stored
            const uint32_t ad0 = raw + (uint32_t)((int)(((uint32_t)sg * (uint32_t)a.kc + (uint32_t)g * 8u) * L) + p_lo) * 2u;
            const bool real = g * 8 < rows_real;
            uint32_t v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = 0u;
                if (real) {
                    const uint32_t ad = ad0 + (uint32_t)r * Lb;
                    const uint32_t w0 = lds32(ad & ~3u), w1 = lds32((ad & ~3u) + 4u);
                    v[r] = __funnelshift_r(w0, w1, (ad & 2u) << 3);
                }
            }
            if (BN && real) {
#pragma unroll
                for (int r = 0; r < 8; ++r) v[r] = bnrelu2(v[r], sb[kbase + g * 8 + r], sb[a.Kpad + kbase + g * 8 + r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t d = dst + ((uint32_t)r << 7) + ((chunk ^ (uint32_t)r) << 4);
                if (ok_lo && ok_hi) sts32(d, v[r]);
                else if (ok_lo) sts16(d, v[r] & 0xffffu);
                else if (ok_hi) sts16(d + 2u, v[r] >> 16);
            }
        }
    }
}

// 2 x 16 accumulator columns of one output row -> bf16 -> (+ residual) -> staging row, in 8-byte pieces.  All residual
// pieces are fetched before the first store (the shared-memory helpers are volatile asm: program order is issue order).
__device__ __forceinline__ void p2_epi_round_fast(const uint32_t (&v0)[16], const uint32_t (&v1)[16], bool two, uint32_t rowaddr,
                                                  int c0, int ncols, bool has_res) {
    uint2 rr[8];
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + 4 * j;
            rr[j] = make_uint2(0u, 0u);
            if (c < ncols && (j < 4 || two)) rr[j] = lds64(rowaddr + (uint32_t)c * 2u);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + 4 * j;
        if (c < ncols && (j < 4 || two)) {
            const uint32_t *v = j < 4 ? &v0[4 * j] : &v1[4 * (j - 4)];
            uint2 o = make_uint2(pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])),
                                 pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3])));
            if (has_res) {
                o.x = pack_bf16x2(bf16_lo(o.x) + bf16_lo(rr[j].x), bf16_hi(o.x) + bf16_hi(rr[j].x));
                o.y = pack_bf16x2(bf16_lo(o.y) + bf16_lo(rr[j].y), bf16_hi(o.y) + bf16_hi(rr[j].y));
            }
            sts64(rowaddr + (uint32_t)c * 2u, o);
        }
    }
}

template <int MODE, bool BN> __global__ void __launch_bounds__(kP2Threads, 1) k_pw2(const P2Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    P2Hdr *hdr = reinterpret_cast<P2Hdr *>(smem);
    float *smem_sb = reinterpret_cast<float *>(smem + a.off_sb);
    const uint32_t s_w = smem_u32(smem + a.off_w), s_raw = smem_u32(smem + a.off_raw), s_op = smem_u32(smem + a.off_op),
                   s_stg = smem_u32(smem + a.off_stg);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * a.Ncta;
    const int nrows = min(a.Ncta, a.N - n0);
    const int tile0 = blockIdx.x, tstride = gridDim.x;
    const bool has_res = a.res != nullptr;
    P2_TRACE(tid == 0, 0);

    if (tid == 0) {
        for (int i = 0; i < kP2MaxRing; ++i) {
            mbar_init(&hdr->raw_full[i], 1);
            mbar_init(&hdr->raw_empty[i], kP2RelPerGroup);
            mbar_init(&hdr->op_full[i], kP2RelPerGroup);
            mbar_init(&hdr->op_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hdr->tmem_full[i], 1);
            mbar_init(&hdr->tmem_empty[i], kP2NumEpi);
            mbar_init(&hdr->res_full[i], 1);
            mbar_init(&hdr->stg_ready[i], kP2NumEpi);
            mbar_init(&hdr->stg_empty[i], 1);
        }
        mbar_init(&hdr->w_full, 1);
        mbar_fence_init();
        // the weight image was packed (and fenced) at least two launches ago: fetch it while the previous kernel of the
        // stream drains (common.cuh: programmatic dependent launch); everything else waits for that kernel below
        mbar_expect_tx(&hdr->w_full, a.w_bytes);
        bulk_g2s(s_w, a.wimg + (size_t)blockIdx.y * a.w_bytes, a.w_bytes, &hdr->w_full);
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&hdr->tmem_base, 512u);
    }
    pdl_sync();
    if (BN)
        for (int k = tid; k < a.Kpad; k += kP2Threads) {
            smem_sb[k] = k < a.K ? a.a_sb[2 * k] : 0.f;
            smem_sb[a.Kpad + k] = k < a.K ? a.a_sb[2 * k + 1] : 0.f;
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    const uint32_t Lb = (uint32_t)a.L * 2u;

    if (warp == kP2MmaWarp) {
        // ================================ MMA issuer: one thread ===================================================
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(128, a.Nmma, /*weights: K-major*/ 0, /*activations: MN-major*/ 1);
            const uint64_t adesc0 = smem_desc(s_w, a.w_lbo, 128, LAYOUT_NONE);
            // activations: per 8-channel group 8 rows of 64 pixels (128 B, chunks XOR-swizzled by the row), groups 1 KiB apart
            // (SBO), the next 64 pixels G KiB further (LBO)
            const uint64_t bdesc0 = smem_desc(s_op, (uint32_t)a.G * 1024u, 1024, LAYOUT_SW128);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
            const uint32_t a_kstep = (2 * a.w_lbo) >> 4, op16 = a.op_stage_bytes >> 4;
            P2_TRACE(true, 1);
            p2_spin(&hdr->w_full, 0, a.wait_ns);
            tc_fence_after();
            P2_TRACE(true, 2);
            // ring positions advance incrementally everywhere in this kernel: a run-time `%` is ~40 dependent instructions,
            // and the control code of a stage -- not its work -- is what sets the pace of these single-warp roles
            int it = 0, o = 0;
            uint32_t oph = 0;
            const int ksteps_full = a.kc >> 4, ksteps_last = (a.Kpad - (a.k_stages - 1) * a.kc) >> 4;
            for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
                const int as = it & 1;
                p2_spin(&hdr->tmem_empty[as], ((uint32_t)(it >> 1) & 1u) ^ 1u, a.wait_ns);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)as * 256u;
                uint32_t a_lo = a_lo0, acc = 0u;
                P2_TRACE(it < 4, 8 + it * 12 + 0);
                for (int st = 0; st < a.k_stages; ++st) {
                    p2_spin(&hdr->op_full[o], oph, a.wait_ns);
                    tc_fence_after();
                    P2_TRACE(it < 4 && st == 0, 8 + it * 12 + 1);
                    P2_TRACE(it < 4 && st == a.k_stages - 1, 8 + it * 12 + 2);
                    const int ksteps = st == a.k_stages - 1 ? ksteps_last : ksteps_full;
                    uint32_t b_lo = b_lo0 + (uint32_t)o * op16;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        if (!P2_DBG(2)) mma_bf16_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                        acc = 1u;
                        a_lo += a_kstep;
                        b_lo += 2048u >> 4;
                    }
                    mma_commit(&hdr->op_empty[o]);
                    if (++o == a.op_stages) { o = 0; oph ^= 1u; }
                }
                mma_commit(&hdr->tmem_full[as]);
            }
        }
        __syncwarp();
    } else if (warp == kP2TmaWarp) {
        // ================================ TMA load warp: weights, raw ring, residual blocks =========================
        int it = 0, r = 0, buf = 0;
        uint32_t rph = 0, bph = 0;  // raw-ring phase, staging-buffer phase
        const uint32_t rows_last = (uint32_t)(a.K - (a.k_stages - 1) * a.kc);
        const size_t stage_elems = (size_t)a.kc * a.HW;  // source advance per K chunk
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            int img0, p0, nseg;
            p2_tile(a, tile, img0, p0, nseg);
            bool res_pending = has_res;
            // staging buffer `buf` is free once the store warp has seen the store of its previous use finish reading
            const uint32_t empty_par = bph ^ 1u;
            auto issue_residual = [&]() {
                const uint32_t dst0 = s_stg + (uint32_t)buf * a.stg_buf_bytes;
                if (a.caseA) {
                    const uint32_t bytes = (uint32_t)nrows * Lb;
                    if (lane == 0) mbar_expect_tx(&hdr->res_full[buf], bytes * (uint32_t)nseg);
                    __syncwarp();
                    if (lane < nseg)
                        bulk_g2s(dst0 + (uint32_t)lane * a.stg_seg_stride,
                                 a.res + ((size_t)(img0 + lane) * a.N + n0) * a.HW, bytes, &hdr->res_full[buf]);
                } else {
                    if (lane == 0) mbar_expect_tx(&hdr->res_full[buf], (uint32_t)nrows * Lb);
                    __syncwarp();
                    for (int m = lane; m < nrows; m += 32)
                        bulk_g2s(dst0 + (uint32_t)m * a.stg_pitch, a.res + ((size_t)img0 * a.N + n0 + m) * a.HW + p0, Lb,
                                 &hdr->res_full[buf]);
                }
                res_pending = false;
            };
            // this lane's source of stage 0: case A lane = segment (image img0 + lane), case B lane = channel row
            const __nv_bfloat16 *src = a.caseA ? a.x + (size_t)(img0 + (lane < nseg ? lane : 0)) * a.K * a.HW
                                               : a.x + ((size_t)img0 * a.K + lane) * a.HW + p0;
            const uint32_t dlane = a.caseA ? (uint32_t)lane * (uint32_t)a.kc * Lb : (uint32_t)lane * Lb;
            for (int st = 0; st < a.k_stages; ++st) {
                P2_TRACE(lane == 0 && it < 4 && st == 0, 8 + it * 12 + 6);
                P2_TRACE(lane == 0 && it < 4 && st == a.k_stages - 1, 8 + it * 12 + 7);
                if (res_pending) {
                    int ok = 0;
                    if (lane == 0) ok = (int)mbar_test(&hdr->stg_empty[buf], empty_par);
                    if (__shfl_sync(0xffffffffu, ok, 0) != 0) issue_residual();
                }
                P2_TRACE(lane == 0 && it == 2 && st < 9, 56 + 2 * st);
                mbar_wait_warp(&hdr->raw_empty[r], rph ^ 1u, lane, a.wait_ns);
                P2_TRACE(lane == 0 && it == 2 && st < 9, 57 + 2 * st);
                const uint32_t rows = st == a.k_stages - 1 ? rows_last : (uint32_t)a.kc;
                const uint32_t dst0 = s_raw + (uint32_t)r * a.raw_stage_bytes;
                if (P2_DBG(16)) {
                    if (lane == 0) mbar_arrive(&hdr->raw_full[r]);
                } else if (a.caseA) {
                    const uint32_t bytes = rows * Lb;
                    if (lane == 0) mbar_expect_tx(&hdr->raw_full[r], bytes * (uint32_t)nseg);
                    __syncwarp();
                    if (lane < nseg) bulk_g2s(dst0 + dlane, src, bytes, &hdr->raw_full[r]);
                } else {
                    if (lane == 0) mbar_expect_tx(&hdr->raw_full[r], rows * Lb);
                    __syncwarp();
                    for (uint32_t row = (uint32_t)lane; row < rows; row += 32)
                        bulk_g2s(dst0 + row * Lb, src + (size_t)(row - (uint32_t)lane) * a.HW, Lb, &hdr->raw_full[r]);
                }
                __syncwarp();
                src += stage_elems;
                if (++r == a.raw_stages) { r = 0; rph ^= 1u; }
            }
            if (res_pending) {
                mbar_wait_warp(&hdr->stg_empty[buf], empty_par, lane, a.wait_ns);
                issue_residual();
            }
            if (++buf == a.stg_bufs) { buf = 0; bph ^= 1u; }
        }
    } else if (warp == kP2StoreWarp) {
        // ================================ TMA store warp ===========================================================
        // waits until the 8 epilogue warps have filled a staging buffer, sends it to global memory with bulk copies, and hands
        // the buffer back (stg_empty) once the copies have read it: to the load warp (next residual block) or, without a
        // residual, straight to the epilogue warps
        int it = 0, buf = 0;
        uint32_t bph = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            int img0, p0, nseg;
            p2_tile(a, tile, img0, p0, nseg);
            const uint32_t stg = s_stg + (uint32_t)buf * a.stg_buf_bytes;
            mbar_wait_warp(&hdr->stg_ready[buf], bph, lane, a.wait_ns);
            P2_TRACE(lane == 0 && it < 4, 8 + it * 12 + 4);
            if (P2_DBG(8)) {
            } else if (a.caseA) {
                if (lane < nseg)
                    bulk_s2g(a.out + ((size_t)(img0 + lane) * a.N + n0) * a.HW, stg + (uint32_t)lane * a.stg_seg_stride,
                             (uint32_t)nrows * Lb);
            } else {
                for (int mm = lane; mm < nrows; mm += 32)
                    bulk_s2g(a.out + ((size_t)img0 * a.N + n0 + mm) * a.HW + p0, stg + (uint32_t)mm * a.stg_pitch, Lb);
            }
            bulk_commit();
            bulk_wait_read<0>();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->stg_empty[buf]);
            P2_TRACE(lane == 0 && it < 4, 8 + it * 12 + 5);
            if (++buf == a.stg_bufs) { buf = 0; bph ^= 1u; }
        }
        bulk_wait_all();  // global writes complete before the CTA exits
    } else if (warp >= kP2RelWarp0 && warp < kP2EpiWarp0) {
        // ================================ relayout warps: raw stage -> UMMA operand =================================
        // The 8 relayout warps form kP2RelGroups groups that take alternate stages: the fixed cost of a stage for a warp (two
        // barrier waits, fence, arrivals: ~0.3 us of serial latency) is then paid every other stage period.
        const int rw = warp - kP2RelWarp0, grp = rw % kP2RelGroups, wg = rw / kP2RelGroups;
        const P2Units pu = p2_make_units(a, MODE, wg);
        int tile = tile0, st = grp, r = grp, o = grp;
        uint32_t rph = 0, oph = 0;
        while (st >= a.k_stages) { st -= a.k_stages; tile += tstride; }
        while (r >= a.raw_stages) { r -= a.raw_stages; rph ^= 1u; }
        while (o >= a.op_stages) { o -= a.op_stages; oph ^= 1u; }
        const int rows_real_last = a.K - (a.k_stages - 1) * a.kc, rows_pad_last = a.Kpad - (a.k_stages - 1) * a.kc;
        while (tile < a.total_tiles) {
            int img0, p0, nseg;
            p2_tile(a, tile, img0, p0, nseg);
            const bool last = st == a.k_stages - 1;
            const int rows_real = last ? rows_real_last : a.kc, rows_pad = last ? rows_pad_last : a.kc;
            if (P2_DBG(128) && !P2_DBG(512)) mbar_wait_warp(&hdr->op_empty[o], oph ^ 1u, lane, a.wait_ns);
            P2_TRACE(rw == 0 && lane == 0 && tile == tile0 + 2 * tstride && st < 12 && P2_DBG(128), 80 + 2 * st);
            if (!P2_DBG(256)) mbar_wait_warp(&hdr->raw_full[r], rph, lane, a.wait_ns);
            P2_TRACE(rw == 0 && lane == 0 && (tile - tile0) / tstride < 4 && st == 0, 8 + ((tile - tile0) / tstride) * 12 + 8);
            P2_TRACE(rw == 0 && lane == 0 && tile == tile0 + 2 * tstride && st < 12, 104 + 2 * st);
            if (!P2_DBG(128) && !P2_DBG(512)) mbar_wait_warp(&hdr->op_empty[o], oph ^ 1u, lane, a.wait_ns);
            P2_TRACE(rw == 0 && lane == 0 && tile == tile0 + 2 * tstride && st < 12 && !P2_DBG(128), 80 + 2 * st);
            if (!P2_DBG(1))
                p2_relayout_stage<MODE, BN>(a, s_raw + (uint32_t)r * a.raw_stage_bytes, s_op + (uint32_t)o * a.op_stage_bytes, smem_sb,
                                            st * a.kc, rows_real, rows_pad, nseg, pu, lane);
            if (!P2_DBG(32)) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&hdr->op_full[o]);
                mbar_arrive(&hdr->raw_empty[r]);
            }
            P2_TRACE(rw == 0 && lane == 0 && tile == tile0 + 2 * tstride && st < 12, 105 + 2 * st);
            P2_TRACE(rw == 0 && lane == 0 && (tile - tile0) / tstride < 4 && last, 8 + ((tile - tile0) / tstride) * 12 + 9);
            st += kP2RelGroups;
            while (st >= a.k_stages) { st -= a.k_stages; tile += tstride; }
            r += kP2RelGroups;
            while (r >= a.raw_stages) { r -= a.raw_stages; rph ^= 1u; }
            o += kP2RelGroups;
            while (o >= a.op_stages) { o -= a.op_stages; oph ^= 1u; }
        }
    } else if (warp >= kP2EpiWarp0) {
        // ================================ epilogue warps ==========================================================
        const int e = warp - kP2EpiWarp0, q = warp & 3, half = e >> 2;
        const int m = q * 32 + lane;  // output channel row (TMEM lane) of this thread inside the CTA's slice
        const bool rowok = m < nrows;
        const bool fast = (a.S == 1) && (a.L % 4 == 0);
        int it = 0, buf = 0;
        uint32_t bph = 0;
        for (int tile = tile0; tile < a.total_tiles; tile += tstride, ++it) {
            int img0, p0, nseg;
            p2_tile(a, tile, img0, p0, nseg);
            const int as = it & 1;
            const uint32_t stg = s_stg + (uint32_t)buf * a.stg_buf_bytes;
            const int ncols = nseg * a.L;
            const int nch = (ncols + 15) >> 4, nch0 = (nch + 1) >> 1;
            const int ch_lo = half ? nch0 : 0, ch_hi = half ? nch : nch0;
            mbar_wait_warp(&hdr->tmem_full[as], (uint32_t)(it >> 1) & 1u, lane, a.wait_ns);
            tc_fence_after();
            P2_TRACE(e == 0 && lane == 0 && it < 4, 8 + it * 12 + 3);
            // the staging buffer holds the residual block (which also means the previous store has released it), or is free
            if (has_res) mbar_wait_warp(&hdr->res_full[buf], bph, lane, a.wait_ns);
            else mbar_wait_warp(&hdr->stg_empty[buf], bph ^ 1u, lane, a.wait_ns);
            P2_TRACE(e == 0 && lane == 0 && it < 4, 8 + it * 12 + 10);
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
            const uint32_t rowaddr = stg + (uint32_t)m * a.stg_pitch;
            // two 16-column chunks per round: both TMEM loads are in flight before the first conversion
            for (int ch = ch_lo; ch < (P2_DBG(4) ? ch_lo : ch_hi); ch += 2) {
                const int c0 = ch << 4;
                const bool two = ch + 1 < ch_hi;
                uint32_t v0[16], v1[16];
                __syncwarp();
                tmem_ld16(tbase + (uint32_t)c0, v0);
                if (two) tmem_ld16(tbase + (uint32_t)c0 + 16u, v1);
                tmem_ld_wait();
                if (ch + 2 >= ch_hi) {  // last TMEM read of this warp for the tile: hand the accumulator stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&hdr->tmem_empty[as]);
                }
                if (!rowok) continue;
                if (fast) {
                    p2_epi_round_fast(v0, v1, two, rowaddr, c0, ncols, has_res);
                } else {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (hh == 1 && !two) break;
                        const int cb = c0 + 16 * hh;
                        int sg = cb / a.L, p = cb - sg * a.L;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (cb + j < ncols) {
                                const uint32_t ad = stg + (uint32_t)sg * a.stg_seg_stride + (uint32_t)m * a.stg_pitch + (uint32_t)p * 2u;
                                float f = __bfloat162float(__float2bfloat16_rn(__uint_as_float(hh ? v1[j] : v0[j])));
                                if (has_res) f += __uint_as_float(lds16(ad) << 16);
                                sts16(ad, (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f)));
                            }
                            if (++p == a.L) { p = 0; ++sg; }
                        }
                    }
                }
            }
            if (ch_lo >= ch_hi || P2_DBG(4)) {  // no column chunk for this warp (tiny tile): still release the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr->tmem_empty[as]);
            }
            // staging rows of this warp complete -> visible to the async proxy -> tell the store warp
            if (!P2_DBG(64)) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->stg_ready[buf]);
            if (++buf == a.stg_bufs) { buf = 0; bph ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    P2_TRACE(tid == 0, 3);
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512u);
    }
}

__host__ __device__ inline int p2_round_up(int v, int m) { return (v + m - 1) / m * m; }

// output-channel slicing shared by the weight packer and the kernel: <= 128 rows per slice (one M tile), slices of <= 112 KiB
__host__ __device__ inline void p2_slices(int rows, int contraction, int *gy, int *ncta, int *kpad, uint32_t *w_lbo, uint32_t *w_bytes) {
    const int Kpad = p2_round_up(contraction, 16);
    int g = (rows + kP2MaxRows - 1) / kP2MaxRows;
    for (;; ++g) {
        const int nc = p2_round_up((rows + g - 1) / g, 8);
        const size_t bytes = (size_t)(Kpad >> 3) * (nc * 16 + 16);
        if (nc <= kP2MaxRows && (bytes <= 112 * 1024 || nc <= 8)) {
            *gy = g; *ncta = nc; *kpad = Kpad;
            *w_lbo = (uint32_t)(nc * 16 + 16);
            *w_bytes = (uint32_t)bytes;
            return;
        }
    }
}

// schedule overrides (rb_pw_conv_set_tuning2): operand-ring depth and K chunk of the image kernel; 0 = automatic
int g_p2_op_stages = 0, g_p2_kc = 0, g_p2_wait_ns = 0;

bool p2_plan(P2Args &a, dim3 *grid, size_t *smem_bytes) {
    if (a.NI <= 0 || a.K <= 0 || a.N <= 0 || a.HW <= 0) return false;
    if (a.K % 8 != 0 || a.N % 8 != 0) return false;
    a.wait_ns = (uint32_t)g_p2_wait_ns;
    if ((int64_t)a.NI * a.HW * (a.K > a.N ? a.K : a.N) >= (int64_t(1) << 31)) return false;
    p2_slices(a.N, a.K, &a.gy, &a.Ncta, &a.Kpad, &a.w_lbo, &a.w_bytes);
    if (a.gy > 148) return false;
    if (a.HW <= 224) {  // whole images per tile
        a.caseA = 1;
        a.L = a.HW;
        a.S = 224 / a.HW;
        if (a.S > 32) a.S = 32;
        a.tpi = 1;
        a.total_tiles = cdiv(a.NI, a.S);
        a.V = a.HW % 8 == 0 ? 8 : (a.HW % 4 == 0 ? 4 : (a.HW % 2 == 0 ? 2 : 1));
        a.stg_pitch = (uint32_t)a.L * 2u;
        a.stg_seg_stride = (uint32_t)a.Ncta * a.stg_pitch;
    } else {  // L pixels of one image: the largest divisor of HW that is a multiple of 8 and <= 256
        if (a.HW % 8 != 0) return false;
        int L = 0;
        for (int c = 256; c >= 64; c -= 8)
            if (a.HW % c == 0) { L = c; break; }
        if (!L) return false;
        a.caseA = 0;
        a.L = L;
        a.S = 1;
        a.tpi = a.HW / L;
        a.total_tiles = a.NI * a.tpi;
        a.V = 8;
        a.stg_pitch = (uint32_t)L * 2u + 16u;
        a.stg_seg_stride = 0;
    }
    a.Nmma = p2_round_up(a.S * a.L, 16);
    if (a.Nmma > 256 || a.Nmma < 16) return false;
    a.atoms = cdiv(a.Nmma, 64);
    a.ppr = (uint32_t)(a.L / a.V);
    if (a.ppr <= 1) { a.ppr_mul = 0; a.ppr_shr = 0; }
    else {
        uint32_t l = 0;
        while ((1u << l) < a.ppr) ++l;
        a.ppr_mul = (uint32_t)(((uint64_t(1) << (31 + l)) + a.ppr - 1) / a.ppr);
        a.ppr_shr = l - 1;
    }
    const uint32_t sb_bytes = a.a_sb ? (uint32_t)p2_round_up(2 * a.Kpad * 4, 128) : 0u;
    a.off_sb = kP2Hdr;
    a.off_w = a.off_sb + sb_bytes;
    const uint32_t w_alloc = (uint32_t)p2_round_up((int)a.w_bytes + 2048, 1024);  // the M = 128 tile reads up to 2 KiB per k-group
    uint32_t fixed = (uint32_t)p2_round_up((int)(a.off_w + w_alloc), 1024);
    const uint32_t stg1 = (uint32_t)p2_round_up(a.caseA ? a.S * (int)a.stg_seg_stride : a.Ncta * (int)a.stg_pitch, 128);
    // Rings: the operand ring only decouples the relayout warps from the MMA issuer (2 stages); the RAW ring is what keeps
    // global memory busy -- bytes in flight per SM = raw stages x stage bytes -- so it gets all the remaining room.  Pick the
    // K chunk (32 or 16 channels) and the number of staging buffers that maximise the bytes in flight.
    int best_bytes = -1;
    const int nop = g_p2_op_stages >= 2 && g_p2_op_stages <= kP2MaxRing ? g_p2_op_stages : 3;
    for (int kc = 64; kc >= 16; kc -= 16) {
        if (kc == 48 || (g_p2_kc ? kc != g_p2_kc : kc == 64)) continue;  // 64-channel chunks only on request (tuning)
        const uint32_t raw_b = (uint32_t)p2_round_up(a.S * kc * a.L * 2, 128);
        const uint32_t op_b = (uint32_t)a.atoms * (uint32_t)(kc >> 3) * 1024u;
        for (int bufs = 2; bufs >= 1; --bufs) {
            const int64_t room = (int64_t)kP2Smem - fixed - (int64_t)bufs * stg1 - nop * (int64_t)op_b;
            if (room < (int64_t)2 * raw_b) continue;
            int stages = (int)(room / raw_b);
            if (stages > kP2MaxRing) stages = kP2MaxRing;
            int inflight = stages * (int)raw_b;
            if (bufs == 2) inflight += inflight / 16;  // mild preference for the second staging buffer at equal depth
            if (inflight <= best_bytes) continue;
            best_bytes = inflight;
            a.kc = kc; a.G = kc >> 3;
            a.k_stages = cdiv(a.Kpad, kc);
            a.raw_stages = stages;
            a.op_stages = nop;
            a.stg_bufs = bufs;
            a.raw_stage_bytes = raw_b; a.op_stage_bytes = op_b; a.stg_buf_bytes = stg1;
            a.off_op = fixed;                                    // 1 KiB aligned (SWIZZLE_128B atoms)
            a.off_raw = a.off_op + (uint32_t)nop * op_b;
            a.off_stg = a.off_raw + (uint32_t)stages * raw_b;
            *smem_bytes = (size_t)a.off_stg + (size_t)bufs * stg1;
        }
    }
    if (best_bytes < 0) return false;
    {   // every relayout warp owns at most kP2MaxUnits (segment, group, 32-piece pass) units of a stage
        const int mode = a.V == 8 ? 16 : (a.V == 4 ? 8 : 4);
        const int passes = (mode == 4) ? ((a.L >> 1) + 2 + 31) >> 5 : (a.L / (mode / 2) + 31) >> 5;
        if (a.S * a.G * passes > kP2RelPerGroup * kP2MaxUnits) return false;
    }
    int gx = sm_count() / a.gy;
    if (gx < 1) gx = 1;
    if (gx > a.total_tiles) gx = a.total_tiles;
    *grid = dim3((unsigned)gx, (unsigned)a.gy, 1);
    return true;
}

template <int MODE, bool BN> int p2_launch(const P2Args &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_pw2<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem);
        if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(k_pw2): %s", cudaGetErrorString(e));
        configured_dev = dev;
    }
    launch_kernel(k_pw2<MODE, BN>, dim3(grid), dim3(kP2Threads), smem_bytes, s, a);
    return launched("k_pw2");
}

template <bool BN> int p2_launch_v(const P2Args &a, dim3 grid, size_t smem_bytes, cudaStream_t s) {
    if (a.V == 8) return p2_launch<16, BN>(a, grid, smem_bytes, s);
    if (a.V == 4) return p2_launch<8, BN>(a, grid, smem_bytes, s);
    return p2_launch<4, BN>(a, grid, smem_bytes, s);
}

// 16-byte unit u of the image of a [rows x contraction] matrix (trans: W^T of the [contraction x rows]... see pw2_weight_pack)
__device__ __forceinline__ void p2_pack_unit(const float *__restrict__ w, unsigned char *__restrict__ img, int64_t u, int rows,
                                             int contraction, int trans, int ncta, int kgroups, uint32_t w_lbo, uint32_t w_bytes) {
    const int n = (int)(u % (ncta + 1));
    const int kg = (int)((u / (ncta + 1)) % kgroups);
    const int sl = (int)(u / ((int64_t)(ncta + 1) * kgroups));
    const int row = sl * ncta + n;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kg * 8 + j;
        f[j] = (n < ncta && row < rows && k < contraction) ? (trans ? w[(int64_t)k * rows + row] : w[(int64_t)row * contraction + k]) : 0.f;
    }
    *reinterpret_cast<uint4 *>(img + (size_t)sl * w_bytes + (size_t)kg * w_lbo + (size_t)n * 16) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// all conv weights of a model in ONE launch (grid.y = weight): both orientations of every item
struct P2PackItem {
    const float *w;
    unsigned char *img_fwd, *img_bwd;
    int N, K;
};
__global__ void k_pw2_pack_multi(const P2PackItem *__restrict__ items) {
    pdl_sync();
    const P2PackItem it = items[blockIdx.y];
    for (int trans = 0; trans < 2; ++trans) {
        const int rows = trans ? it.K : it.N, contraction = trans ? it.N : it.K;
        int gy, ncta, kpad;
        uint32_t lbo, wb;
        p2_slices(rows, contraction, &gy, &ncta, &kpad, &lbo, &wb);
        const int64_t total = (int64_t)gy * (kpad >> 3) * (ncta + 1);
        for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x)
            p2_pack_unit(it.w, trans ? it.img_bwd : it.img_fwd, u, rows, contraction, trans, ncta, kpad >> 3, lbo, wb);
    }
}

// one thread per 16-byte unit (slice, k-group, row): 8 consecutive k of weight row n0 + n (zero beyond the matrix)
__global__ void k_pw2_pack(const float *__restrict__ w, unsigned char *__restrict__ img, int rows, int contraction, int trans,
                           int gy, int ncta, int kgroups, uint32_t w_lbo, uint32_t w_bytes) {
    pdl_sync();
    const int64_t total = (int64_t)gy * kgroups * (ncta + 1);
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    p2_pack_unit(w, img, u, rows, contraction, trans, ncta, kgroups, w_lbo, w_bytes);
}

}  // namespace

#ifdef RB_DEBUG_TRACE
unsigned long long *pw_conv_get_trace();
static int g_p2_dbg = 0;
void pw2_set_debug(int flags) { g_p2_dbg = flags; }
#endif

void pw2_set_tuning(int op_stages, int kc, int wait_ns) {
    g_p2_op_stages = op_stages;
    g_p2_kc = (kc == 16 || kc == 32 || kc == 64) ? kc : 0;
    g_p2_wait_ns = wait_ns < 0 ? 0 : wait_ns;
}

// bytes of the packed image of a [rows x contraction] weight matrix (all slices)
size_t pw2_weight_image_bytes(int rows, int contraction) {
    if (rows <= 0 || contraction <= 0) return 0;
    int gy, ncta, kpad;
    uint32_t lbo, wb;
    p2_slices(rows, contraction, &gy, &ncta, &kpad, &lbo, &wb);
    return (size_t)gy * wb;
}

// w: fp32 conv weight [N, K].  trans == 0: image of W (rows = N, contraction = K: the forward GEMM); trans == 1: image of
// W^T (rows = K, contraction = N: the input-gradient GEMM).
int pw2_weight_pack(const float *w, int N, int K, int trans, void *image, cudaStream_t s) {
    const int rows = trans ? K : N, contraction = trans ? N : K;
    int gy, ncta, kpad;
    uint32_t lbo, wb;
    p2_slices(rows, contraction, &gy, &ncta, &kpad, &lbo, &wb);
    const int64_t total = (int64_t)gy * (kpad >> 3) * (ncta + 1);
    // `w` is indexed [N, K] in both cases: W^T[row = k, col = n] = w[n * K + k] = w[col * rows + row]
    launch_kernel(k_pw2_pack, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, s, w, (unsigned char *)image, rows, contraction, trans, gy, ncta, kpad >> 3,
                                                          lbo, wb);
    if (int rc = launched("k_pw2_pack")) return rc;
    return launch_fence(s);  // the packed operands are RB_W_RESIDENT for every later launch
}

// items: DEVICE array of `count` {const float *weight [N,K]; void *image_fwd; void *image_bwd; int N; int K} (rb_pw_pack_item_t)
int pw2_weight_pack_multi(const void *items_device, int count, cudaStream_t s) {
    static_assert(sizeof(P2PackItem) == 32, "rb_pw_pack_item_t layout");
    launch_kernel(k_pw2_pack_multi, dim3(dim3(48, (unsigned)count)), dim3(256), 0, s, (const P2PackItem *)items_device);
    if (int rc = launched("k_pw2_pack_multi")) return rc;
    return launch_fence(s);  // the packed operands are RB_W_RESIDENT for every later launch
}

// 0: no image path; 1: supported; 2: supported and measured faster than the first-generation kernel for this geometry.
// Measured on B200 at 32 clips (profiles/r02l_bench_pw.log, us per launch, first generation -> image kernel):
//   7x7 maps (4 images per tile): plain 161 -> 116, +residual 165 -> 116, bn+relu producer 269 -> 132   => level 2
//   14x14 maps (1 image per tile): plain 41 -> 43, +residual 57 -> 51, bn+relu 51 -> 57                  => level 1 (a wash)
//   28x28 .. 112x112 (row pieces, one small bulk copy per channel row: ~300 ns of issue each)  1.5-4x slower => level 1
int pw2_supported(int NI, int K, int N, int HW, int has_bn) {
    P2Args a{};
    a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    a.a_sb = has_bn ? reinterpret_cast<const float *>(uintptr_t(16)) : nullptr;
    dim3 grid;
    size_t smem = 0;
    if (!p2_plan(a, &grid, &smem)) return 0;
    return (a.caseA && a.S > 1) ? 2 : 1;
}

int pw2_forward(const void *x, const void *wimg, const void *residual, void *out, int NI, int K, int N, int HW,
                const float *a_sb, cudaStream_t s) {
    P2Args a{};
    a.x = (const __nv_bfloat16 *)x; a.wimg = (const unsigned char *)wimg; a.res = (const __nv_bfloat16 *)residual;
    a.out = (__nv_bfloat16 *)out; a.a_sb = a_sb;
    a.NI = NI; a.K = K; a.N = N; a.HW = HW;
    dim3 grid;
    size_t smem = 0;
    if (!p2_plan(a, &grid, &smem))
        return fail(RB_ERR_UNSUPPORTED, "pw_conv (image weights): geometry NI=%d K=%d N=%d HW=%d not supported", NI, K, N, HW);
    const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual) |
                         reinterpret_cast<uintptr_t>(wimg);
    if (al & 15) return fail(RB_ERR_INVALID_ARGUMENT, "pw_conv (image weights): pointers must be 16-byte aligned");
#ifdef RB_DEBUG_TRACE
    a.trace = pw_conv_get_trace();
    a.dbg = g_p2_dbg;
#endif
    return a_sb ? p2_launch_v<true>(a, grid, smem, s) : p2_launch_v<false>(a, grid, smem, s);
}

}  // namespace rb

// Micro-benchmark (diagnostic, not product code): how many bytes per second can ONE SM pull from global memory into
// shared memory with (a) 1-D bulk copies (cp.async.bulk, SASS UBLKCP), (b) 2-D tiled TMA (cp.async.bulk.tensor, SASS
// UTMALDG, SWIZZLE_128B boxes of 64 x R bf16), (c) plain 128-bit loads?  Ring of D slots per CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_bw tools/ubench/tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    } while (!ok);
}

// (a) 1-D bulk copies of `bytes` each, D in flight, `iters` copies per CTA, each CTA streams its own contiguous region
__global__ void k_bulk(const unsigned char *src, size_t region, int bytes, int D, int iters, int lanes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    unsigned char *ring = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < D; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *base = src + (size_t)blockIdx.x * region;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = bytes / lanes;  // each stage is split over `lanes` copies issued by different lanes
        for (int n = 0; n < iters + D; ++n) {
            const int slot = n % D;
            if (n >= D) mbar_wait(&bars[slot], (uint32_t)((n / D) - 1) & 1u);
            if (n < iters) {
                if (lane == 0) mbar_expect(&bars[slot], (uint32_t)bytes);
                __syncwarp();
                if (lane < lanes) {
                    const size_t off = ((size_t)n * bytes) % region;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     s32(ring + (size_t)slot * bytes + (size_t)lane * per)),
                                 "l"(base + off + (size_t)lane * per), "r"((uint32_t)per), "r"(s32(&bars[slot]))
                                 : "memory");
                }
                __syncwarp();
            }
        }
    }
}

// (b) tiled TMA: boxes of 64 bf16 x rows (128 B x rows, SWIZZLE_128B) from a 2-D tensor [H rows][W bf16]
__global__ void k_tensor(const __grid_constant__ CUtensorMap map, int rows, int D, int iters, int boxes_per_row, int row_blocks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    unsigned char *ring = smem + 1024;
    const int bytes = rows * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < D; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int n = 0; n < iters + D; ++n) {
            const int slot = n % D;
            if (n >= D) mbar_wait(&bars[slot], (uint32_t)((n / D) - 1) & 1u);
            if (n < iters) {
                mbar_expect(&bars[slot], (uint32_t)bytes);
                const long long t = (long long)blockIdx.x * iters + n;  // global box index
                const int c0 = (int)(t % boxes_per_row) * 64, c1 = (int)((t / boxes_per_row) % row_blocks) * rows;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                 s32(ring + (size_t)slot * bytes)),
                             "l"(&map), "r"(c0), "r"(c1), "r"(s32(&bars[slot]))
                             : "memory");
            }
        }
    }
}

// (c) plain loads: every thread keeps U 16-byte loads in flight
template <int U> __global__ void k_ldg(const uint4 *src, size_t n16_per_cta, int iters, uint4 *sink) {
    const uint4 *base = src + (size_t)blockIdx.x * n16_per_cta;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(base + (((size_t)it * U + u) * blockDim.x + threadIdx.x) % n16_per_cta);
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    if (acc.x == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t total = (size_t)2 << 30;  // 2 GiB source: far beyond the 126 MB L2
    unsigned char *src;
    uint4 *sink;
    CK(cudaMalloc(&src, total));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(src, 1, total));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("SMs %d\n", sms);
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    for (int small = 0; small < 2; ++small) {
        // small = 1: every CTA streams inside a 32 MB window (L2-resident after the first pass)
        const size_t span = small ? ((size_t)32 << 20) : total;
        for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
            const int grid = sms * ctas_per_sm;
            const size_t region = (span / grid) & ~(size_t)1023;
            const int bytes_list[] = {512, 2048, 12544, 16384, 32768};
            for (int bytes : bytes_list) {
                for (int D = 2; D <= 16; D *= 2) {
                    for (int lanes = 1; lanes <= 8; lanes *= 8) {
                        const size_t smem = 1024 + (size_t)D * bytes;
                        if (smem * ctas_per_sm > 200 * 1024 || (bytes / lanes) % 16) continue;
                        const int iters = (int)(((size_t)24 << 20) / bytes / ctas_per_sm);  // 24 MB per SM
                        k_bulk<<<grid, 64, smem>>>(src, region, bytes, D, 64, lanes);
                        CK(cudaEventRecord(e0));
                        k_bulk<<<grid, 64, smem>>>(src, region, bytes, D, iters, lanes);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        const double gb = (double)grid * iters * bytes / 1e9;
                        printf("bulk1d %s ctas/SM %d bytes %6d D %2d lanes %d : %7.1f GB/s total, %6.2f GB/s per SM\n", small ? "L2 " : "HBM",
                               ctas_per_sm, bytes, D, lanes, gb / (ms / 1e3), gb / (ms / 1e3) / sms);
                    }
                }
            }
        }
    }

    // tiled TMA
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 0; }
    for (int small = 0; small < 2; ++small) {
        const uint64_t W = 8192;  // bf16 per row (16 KB rows)
        const uint64_t H = (small ? ((uint64_t)32 << 20) : (uint64_t)total) / (W * 2);
        for (int rows : {32, 64, 128}) {
            CUtensorMap map;
            cuuint64_t dims[2] = {W, H}, strides[1] = {W * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)rows}, estr[2] = {1, 1};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm *= 2) {
                for (int D = 2; D <= 16; D *= 2) {
                    const int bytes = rows * 128;
                    const size_t smem = 1024 + (size_t)D * bytes;
                    if (smem * ctas_per_sm > 200 * 1024) continue;
                    const int grid = sms * ctas_per_sm;
                    const int iters = (int)(((size_t)24 << 20) / bytes / ctas_per_sm);
                    k_tensor<<<grid, 64, smem>>>(map, rows, D, 64, (int)(W / 64), (int)(H / rows));
                    CK(cudaEventRecord(e0));
                    k_tensor<<<grid, 64, smem>>>(map, rows, D, iters, (int)(W / 64), (int)(H / rows));
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double gb = (double)grid * iters * bytes / 1e9;
                    printf("tiled2d %s ctas/SM %d box 64x%-3d (%5d B) D %2d : %7.1f GB/s total, %6.2f GB/s per SM\n", small ? "L2 " : "HBM",
                           ctas_per_sm, rows, bytes, D, gb / (ms / 1e3), gb / (ms / 1e3) / sms);
                }
            }
        }
    }

    // plain loads
    for (int small = 0; small < 2; ++small) {
        const size_t span = small ? ((size_t)32 << 20) : total;
        for (int ctas_per_sm = 1; ctas_per_sm <= 8; ctas_per_sm *= 2) {
            const int grid = sms * ctas_per_sm, threads = 256;
            const size_t n16 = span / 16 / grid;
            const int iters = (int)(((size_t)24 << 20) / 16 / 8 / threads / ctas_per_sm);
            k_ldg<8><<<grid, threads>>>((const uint4 *)src, n16, 8, sink);
            CK(cudaEventRecord(e0));
            k_ldg<8><<<grid, threads>>>((const uint4 *)src, n16, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double gb = (double)grid * iters * 8 * threads * 16 / 1e9;
            printf("ldg128x8 %s ctas/SM %d (256 thr) : %7.1f GB/s total, %6.2f GB/s per SM\n", small ? "L2 " : "HBM", ctas_per_sm, gb / (ms / 1e3),
                   gb / (ms / 1e3) / sms);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}

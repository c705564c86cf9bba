// Microbenchmark for the round-2 GEMM plan: how many bytes per clock can the SMs pull from L2 through TMA when
// the CTAs of a cluster need the SAME 16 KB operand tile,
//   private   : every CTA streams its own tile (no sharing; the reference point for the L2->SM cap),
//   unicast   : every CTA of the cluster loads the same tile itself (what two CTA pairs working on the same weight rows do today),
//   multicast : CTA r loads 1/csz of the tile and multicasts it to all CTAs of the cluster (cp.async.bulk.tensor ... .multicast::cluster),
// for cluster sizes 1, 2, 4, 8.  The ring protocol is the one the persistent GEMM would use: per stage a `full` mbarrier (1 arrival + 16 KB
// of transaction bytes, possibly from several producers) and an `empty` mbarrier that counts one arrival from the consumer of EVERY CTA of
// the cluster, because a multicast overwrites that stage in all of them.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_mcast_bench tools/tma_mcast_bench.cu && /tmp/tma_mcast_bench
//
// Prints one line per (mode, cluster size): delivered GB/s (bytes landing in shared memory), requested L2 bytes, and a checksum that must
// agree between unicast and multicast.  Not part of the library; no GPU was available when it was written (compiles, never run).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (x);                                                                            \
        if (e_ != cudaSuccess) {                                                                         \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
            exit(1);                                                                                     \
        }                                                                                                \
    } while (0)

constexpr int TILE_ROWS = 128, TILE_COLS = 32;                 // 128 rows x 128 bytes = one k-block of a weight tile
constexpr int TILE_BYTES = TILE_ROWS * TILE_COLS * 4;          // 16 KB
constexpr int STAGES = 6;
constexpr int THREADS = 64;                                    // warp 0: producer (lane 0), warp 1: consumer

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_to_cta(uint32_t local_saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}

enum Mode { PRIVATE = 0, UNICAST = 1, MULTICAST = 2 };

// map_full: box = 32 x 128 (whole tile); map_slice: box = 32 x (128 / csz) (the slice one CTA multicasts)
__global__ void __launch_bounds__(THREADS, 1)
stream_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_slice, int mode, int csz, int row_tiles,
              int k_blocks, int iters, float* __restrict__ checksum) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * TILE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t cluster = cluster_id_x();
    const bool shared_stage = mode == MULTICAST;              // other CTAs write into my stages -> they need my consumer's release
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], shared_stage ? csz : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();                                       // every CTA's barriers exist before anyone multicasts into them
    // the tile this CTA wants: shared by the cluster (unicast / multicast) or its own (private)
    const int tile = (mode == PRIVATE ? (int)(cluster * csz + rank) : (int)cluster) % row_tiles;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (i / STAGES) & 1;
            if (i >= STAGES) mbar_wait(&empty_bar[s], ph ^ 1);
            const int kb = i % k_blocks;
            mbar_expect_tx(&full_bar[s], TILE_BYTES);
            uint8_t* dst = smem + s * TILE_BYTES;
            if (mode == MULTICAST) {
                const int rows = TILE_ROWS / csz;
                tma_load_2d_mcast(&map_slice, &full_bar[s], dst + rank * rows * TILE_COLS * 4, kb * TILE_COLS, tile * TILE_ROWS + rank * rows,
                                  (uint16_t)((1u << csz) - 1));
            } else {
                tma_load_2d(&map_full, &full_bar[s], dst, kb * TILE_COLS, tile * TILE_ROWS);
            }
        }
    } else if (warp == 1) {
        float acc = 0.f;
        for (int i = 0; i < iters; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (i / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            acc += reinterpret_cast<const float*>(smem + s * TILE_BYTES)[lane * 33 % (TILE_ROWS * TILE_COLS)];   // touch the data
            __syncwarp();
            if (shared_stage) {
                if (lane < csz) mbar_arrive_cluster(mapa_to_cta(smem_u32(&empty_bar[s]), lane));
            } else if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) checksum[blockIdx.x] = acc;
    }
    cluster_sync_all();                                       // nobody exits while a peer may still multicast into it / arrive on its barriers
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn fn, float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 4};
    cuuint32_t box[2] = {TILE_COLS, box_rows}, elem[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    CHECK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    const int row_tiles = 148, k_blocks = 32;                  // 148 tiles x 128 rows x 1024 floats = 77.6 MB: resident in the 126 MB L2
    const uint64_t rows = (uint64_t)row_tiles * TILE_ROWS, cols = (uint64_t)k_blocks * TILE_COLS;
    float* A;
    CHECK(cudaMalloc(&A, rows * cols * 4));
    {
        float* h = (float*)malloc(rows * cols * 4);
        for (uint64_t i = 0; i < rows * cols; ++i) h[i] = (float)((i * 2654435761u) % 1024) / 1024.f;
        CHECK(cudaMemcpy(A, h, rows * cols * 4, cudaMemcpyHostToDevice));
        free(h);
    }
    float* checksum;
    CHECK(cudaMalloc(&checksum, 4 * 256));
    const int smem = STAGES * TILE_BYTES + 2 * STAGES * 8 + 1024;
    CHECK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CHECK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int iters = 32 * 64;                                 // 2048 tiles of 16 KB per CTA = 32 MB per CTA
    printf("device %s, %d SMs, clock attr %d kHz; %d stages of %d B, %d iterations per CTA\n", prop.name, prop.multiProcessorCount, clock_khz, STAGES,
           TILE_BYTES, iters);
    printf("%-10s %4s %5s %10s %12s %12s %12s  %s\n", "mode", "csz", "ctas", "ms", "smem GB/s", "smem B/clk", "L2 req GB/s", "checksum");
    for (int csz : {1, 2, 4, 8}) {
        for (int mode : {PRIVATE, UNICAST, MULTICAST}) {
            if (csz == 1 && mode != PRIVATE) continue;
            CUtensorMap full = make_map(fn, A, rows, cols, TILE_ROWS), slice = make_map(fn, A, rows, cols, TILE_ROWS / csz);
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(THREADS);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cfg.gridDim = dim3(csz);
            int max_clusters = 0;
            CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, stream_kernel, &cfg));
            const int clusters = max_clusters < prop.multiProcessorCount / csz ? max_clusters : prop.multiProcessorCount / csz;
            const int ctas = clusters * csz;
            cfg.gridDim = dim3(ctas);
            CHECK(cudaMemset(checksum, 0, 4 * 256));
            cudaEvent_t e0, e1;
            CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {                // rep 0 warms L2
                CHECK(cudaEventRecord(e0));
                CHECK(cudaLaunchKernelEx(&cfg, stream_kernel, full, slice, mode, csz, row_tiles, k_blocks, iters, checksum));
                CHECK(cudaEventRecord(e1));
                CHECK(cudaDeviceSynchronize());
                CHECK(cudaEventElapsedTime(&ms, e0, e1));
            }
            float h[256];
            CHECK(cudaMemcpy(h, checksum, 4 * 256, cudaMemcpyDeviceToHost));
            double cs = 0;
            for (int i = 0; i < ctas; ++i) cs += h[i];
            const double delivered = (double)ctas * iters * TILE_BYTES;
            const double l2_req = mode == MULTICAST ? delivered / csz : delivered;
            printf("%-10s %4d %5d %10.3f %12.1f %12.1f %12.1f  %.3f\n", mode == PRIVATE ? "private" : (mode == UNICAST ? "unicast" : "multicast"), csz, ctas,
                   ms, delivered / ms / 1e6, delivered / (ms * 1e-3) / (clock_khz * 1e3), l2_req / ms / 1e6, cs);
            CHECK(cudaEventDestroy(e0)); CHECK(cudaEventDestroy(e1));
        }
    }
    return 0;
}

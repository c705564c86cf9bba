// Shared helpers for the sm_100a kernels of siu3r_b200 (device code + C-ABI plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SIU3R_OK 0
#define SIU3R_ERR_INVALID -1      // bad argument (shape / alignment / null pointer)
#define SIU3R_ERR_CAPACITY -2     // caller-provided workspace / capacity too small
#define SIU3R_ERR_CUDA -3         // CUDA runtime error (message on stderr)
#define SIU3R_ERR_UNSUPPORTED -4  // configuration not implemented by this kernel

#define SIU3R_CUDA_CHECK(expr)                                                                       \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            fprintf(stderr, "[siu3r_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),       \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                                     \
            return SIU3R_ERR_CUDA;                                                                   \
        }                                                                                            \
    } while (0)

#define SIU3R_LAUNCH_CHECK() SIU3R_CUDA_CHECK(cudaGetLastError())

#define SIU3R_REQUIRE(cond)                                                                          \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            fprintf(stderr, "[siu3r_b200] invalid argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            return SIU3R_ERR_INVALID;                                                                \
        }                                                                                            \
    } while (0)

static inline __host__ __device__ int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline __host__ __device__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline __host__ __device__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute (the opt-in to > 48 KB of dynamic shared memory) is PER DEVICE: once-per-process flags are kept per device index.
static inline bool siu3r_first_use_on_device(bool (&seen)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 63;
    if (seen[dev]) return false;
    seen[dev] = true;
    return true;
}

// Programmatic dependent launch (PDL): a kernel launched through siu3r_launch_pdl may start while the previous kernel of its stream is still
// running -- its prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps that kernel's tail -- and calls pdl_wait() before its first
// global-memory access; pdl_wait() returns once the previous kernel has COMPLETED and its writes are visible, so stream-order semantics are
// unchanged.  pdl_launch_dependents() lets the next kernel's CTAs be scheduled as this kernel's CTAs retire (single-wave / persistent kernels
// call it right after pdl_wait()).  SIU3R_PDL=0 disables the launch attribute (the device instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern "C" int siu3r_pdl_enabled(void);
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t siu3r_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = siu3r_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Every launch of one of OUR kernels bumps this counter (bench.py reports it as gpu_launches).
extern "C" void siu3r_note_launch(int n);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

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

// geodiffuser_b200/csrc/common.cuh -- shared helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define GD_OK 0
#define GD_ERR_INVALID 1     // bad argument (shape / pointer / enum)
#define GD_ERR_CUDA 2        // CUDA runtime error (see gd_last_error)
#define GD_ERR_UNSUPPORTED 3 // shape outside what the kernels were built for

namespace gd {

extern thread_local char g_last_error[256];
int set_error(int code, const char* fmt, ...);

#define GD_CHECK_ARG(cond)                                                                     \
    do {                                                                                       \
        if (!(cond)) return gd::set_error(GD_ERR_INVALID, "%s:%d: `%s`", __FILE__, __LINE__, #cond); \
    } while (0)

#define GD_CHECK_LAUNCH()                                                                      \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess)                                                                \
            return gd::set_error(GD_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

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

// block-wide sum of `v` (all threads get the result); blockDim.x <= 1024, multiple of 32
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 32 floats */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = (l < nw) ? sh[l] : 0.f;
    r = warp_sum(r);
    return r;
}

}  // namespace gd

// Shared helpers for libavtex.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avtex.h"

void avtex_set_error(const char *fmt, ...);

#define AVTEX_CUDA(call)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            avtex_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

#define AVTEX_REQUIRE(cond, ...)              \
    do {                                      \
        if (!(cond)) {                        \
            avtex_set_error(__VA_ARGS__);     \
            return 2;                         \
        }                                     \
    } while (0)

#define AVTEX_ENTER(device) AVTEX_CUDA(cudaSetDevice(device))
#define AVTEX_LAUNCH_CHECK() AVTEX_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide reductions through one shared scratch array of 32 slots (blockDim.x <= 1024,
// multiple of 32).  All threads get the result.  Trailing __syncthreads makes `scratch` reusable.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T identity, Op op, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T r = (lane < nw) ? scratch[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    __syncthreads();
    return r;
}

struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
template <typename T> struct OpAdd { __device__ T operator()(T a, T b) const { return a + b; } };

// Streaming 128-bit load that does not pollute L1 (data touched once per kernel).
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

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

// Coherent, L1-CACHED 128-bit / 32-bit loads (ld.global.ca).  For vectors that other CTAs (or peer GPUs) rewrite
// between grid barriers: a gpu-scope fence / grid.sync invalidates L1, so lines cached after the barrier are
// current, and every later reader on the SM hits L1 instead of pulling the same bytes through L2 again
// (__ldcg made the future-cost sweeps L2-bandwidth bound: two m vectors per D3 element = 3x the D3 bytes
// through the ~12 TB/s L2 fabric).  Never ld.global.nc here: the non-coherent path may keep stale lines.
__device__ __forceinline__ float4 ld_ca_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ld_ca_f(const float *p) {
    float r;
    asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}

// x ** p for x >= 0, p > 0 (the future-cost power D2 ** 0.7, classic/q_learning.py:34).
// powf() costs ~150 instructions per element and made the filter kernel compute-bound.  Here
// x = 2^e * m (m in [1,2)):  x^p = 2^(p*e) * 2^(p*log2 m).  p is split as p_hi + p_lo with p_hi on
// 12 bits, so p_hi*e is exact; its integer part goes straight into the exponent field and only
// small-magnitude terms (|.| < 2) reach ex2, keeping the result within ~3 ulp of powf (the
// reference's CPU powf and CUDA's differ by that much already).  Zero, subnormals, negatives, inf and
// nan take the out-of-line powf.  lg2.approx / ex2.approx are what log2f / exp2f evaluate for arguments
// in these ranges (m in [1,2), t in (-0.6, 1.6)); calling them directly only drops the range fix-ups.
static __device__ __noinline__ float pow_slow(float x, float p) { return x == 0.f ? 0.f : powf(x, p); }

__device__ __forceinline__ float pow_pos(float x, float p) {
    const unsigned ix = __float_as_uint(x);
    const unsigned e9 = ix >> 23;                  // sign + exponent
    if (e9 - 1u >= 254u) return pow_slow(x, p);    // zero / subnormal / inf / nan / negative
    const float e = (float)((int)e9 - 127);
    const float m = __uint_as_float((ix & 0x007fffffu) | 0x3f800000u);
    const float p_hi = __uint_as_float(__float_as_uint(p) & 0xfffff000u);
    const float p_lo = p - p_hi;
    const float a = p_hi * e;                      // exact: 12 x 8 bits
    const float ai = rintf(a);
    if (fabsf(ai) > 100.f) return pow_slow(x, p);  // result exponent near the fp32 limits
    float t = a - ai;                              // exact, |t| <= 0.5
    t = fmaf(p_lo, e, t);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(m));
    t = fmaf(p, lg, t);                            // in (-0.6, 1.6)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));   // in (0.65, 3.1)
    return __int_as_float(__float_as_int(r) + ((int)ai << 23));
}

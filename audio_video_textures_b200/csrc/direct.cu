// K1 (fallback) — pairwise L2 by direct difference, the reference's own formula
// (classic/computeD1.py:88: torch.norm(feats_A - feats_B, dim=2)), fp32 SIMT.
// Used for features that are not integer-valued bytes, for byte frames outside the Gram path's
// d^2 < 2^32 domain, and as the on-device cross-check of the tensor-core path.
// FP32-pipe bound: 2 instructions (sub, fma) per (pair, feature).
// Accumulation is two-level: each 32-feature stage is summed in fp32 and the stage sums are added
// in fp64.  A single running fp32 sum drifts by ~5e-5 at K = 12288 (measured), more than the
// reference's vectorised torch.norm; two-level keeps it below 1e-6, and for byte inputs every
// stage sum is an integer < 2^24, so the result is the exact d^2 (bit-identical to the Gram path).
#include "common.cuh"

namespace {

constexpr int DT = 64;       // output tile edge
constexpr int DK = 32;       // features per smem stage
constexpr int DTHREADS = 256;

template <typename T>
__global__ void __launch_bounds__(DTHREADS)
pairdist_direct_kernel(const T *__restrict__ x, int64_t n, int64_t k, int64_t ld, int64_t row0,
                       int64_t rows, float *__restrict__ D, int64_t ldd, double *sum,
                       unsigned long long *nnz) {
    __shared__ float As[DK][DT + 1];
    __shared__ float Bs[DK][DT + 1];
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t ra0 = row0 + int64_t(blockIdx.y) * DT;     // global row of A tile
    const int64_t rb0 = int64_t(blockIdx.x) * DT;            // global row of B tile (= output column)
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int64_t k0 = 0; k0 < k; k0 += DK) {
        for (int idx = threadIdx.x; idx < DT * DK; idx += DTHREADS) {
            const int r = idx / DK, c = idx % DK;
            const int64_t kc = k0 + c;
            const int64_t ga = ra0 + r, gb = rb0 + r;
            As[c][r] = (ga < row0 + rows && kc < k) ? float(x[ga * ld + kc]) : 0.f;
            Bs[c][r] = (gb < n && kc < k) ? float(x[gb * ld + kc]) : 0.f;
        }
        __syncthreads();
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll 8
        for (int kk = 0; kk < DK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = a[i] - b[j];
                    part[i][j] = fmaf(d, d, part[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j];
        __syncthreads();
    }
    double s = 0.0;
    unsigned long long z = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = ra0 + ty * 4 + i;
        if (gr >= row0 + rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gc = rb0 + tx * 4 + j;
            if (gc >= n) continue;
            const float d = __fsqrt_rn(__double2float_rn(acc[i][j]));   // same rounding sequence as gram.cu
            D[(gr - row0) * ldd + gc] = d;
            s += d;
            z += (d != 0.f);
        }
    }
    if (sum != nullptr) {
        s = block_reduce(s, 0.0, OpAdd<double>(), sred);
        z = block_reduce(z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, s); atomicAdd(nnz, z); }
    }
}

template <typename T>
int launch_direct(const T *x, int64_t n, int64_t k, int64_t ld, int64_t row0, int64_t rows, float *D,
                  int64_t ldd, double *sum, unsigned long long *nnz, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(n > 0 && k > 0 && ld >= k && rows > 0 && row0 >= 0 && row0 + rows <= n && ldd >= n,
                  "pairdist_direct: bad shape n=%lld k=%lld row0=%lld rows=%lld", (long long)n,
                  (long long)k, (long long)row0, (long long)rows);
    AVTEX_REQUIRE((sum == nullptr) == (nnz == nullptr), "pairdist_direct: sum and nnz go together");
    dim3 grid((unsigned)((n + DT - 1) / DT), (unsigned)((rows + DT - 1) / DT));
    AVTEX_REQUIRE(grid.y <= 65535, "pairdist_direct: too many row tiles (%u)", grid.y);
    pairdist_direct_kernel<T><<<grid, DTHREADS, 0, as_stream(stream)>>>(x, n, k, ld, row0, rows, D, ldd, sum, nnz);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int avtex_pairdist_direct_f32(const float *x, int64_t n, int64_t k, int64_t ld,
                                         int64_t row0, int64_t rows, float *D, int64_t ldd,
                                         double *sum, unsigned long long *nnz, int device,
                                         void *stream) {
    return launch_direct<float>(x, n, k, ld, row0, rows, D, ldd, sum, nnz, device, stream);
}

extern "C" int avtex_pairdist_direct_u8(const uint8_t *x, int64_t n, int64_t k, int64_t ld,
                                        int64_t row0, int64_t rows, float *D, int64_t ldd,
                                        double *sum, unsigned long long *nnz, int device,
                                        void *stream) {
    return launch_direct<uint8_t>(x, n, k, ld, row0, rows, D, ldd, sum, nnz, device, stream);
}

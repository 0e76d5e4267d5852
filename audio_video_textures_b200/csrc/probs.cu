// K5 — sigma statistics, transition probabilities, threshold, survivor lists.
//   tail(D, f) of classic/computeD1.py:240-245 == computeD2.py:44-50 == q_learning.py:53-59,
//   per-row threshold of q_learning.py:61-64, and the `nonzero()` the walk calls per step
//   (video_textures.py:76-78).  All HBM-bound row kernels: one CTA per row, 128-bit loads,
//   warp-shuffle + shared-memory block reductions.
#include <float.h>

#include <atomic>

#include "common.cuh"

namespace {

constexpr int PT = 256;

__global__ void __launch_bounds__(PT)
sum_nnz_kernel(const float *__restrict__ D, int64_t rows, int64_t cols, int64_t ld, double *sum,
               unsigned long long *nnz) {
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    double s = 0.0;
    unsigned long long z = 0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const float *row = D + r * ld;
        const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
        const int64_t cv = vec ? (cols & ~int64_t(3)) : 0;
        for (int64_t k = int64_t(threadIdx.x) * 4; k < cv; k += PT * 4) {
            const float4 v = ld_stream_f4(row + k);
            s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
            z += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
        }
        for (int64_t k = cv + threadIdx.x; k < cols; k += PT) {
            const float v = row[k];
            s += (double)v;
            z += (v != 0.f);
        }
    }
    s = block_reduce(s, 0.0, OpAdd<double>(), sred);
    z = block_reduce(z, 0ull, OpAdd<unsigned long long>(), nred);
    if (threadIdx.x == 0) { atomicAdd(sum, s); atomicAdd(nnz, z); }
}

// exp((-d)/sigma): true division, as `torch.exp(-D / sigma)`.
__device__ __forceinline__ float trans_e(float d, float sigma) { return expf(__fdiv_rn(-d, sigma)); }

// CACHE: E = exp(-D/sigma) of the whole row is kept in shared memory between the two passes, so the
// row is read from HBM once and exp is evaluated once (rows up to PROBS_SMEM_COLS floats); otherwise
// pass 2 re-reads the row and recomputes exp (deterministic -> identical values).
constexpr int PROBS_SMEM_COLS = 48 * 1024;

template <bool CACHE, int NT>
__global__ void __launch_bounds__(NT)
transition_probs_kernel(const float *__restrict__ D, int64_t ld, int64_t rows_in, int64_t cols,
                        float sigma, int shift, float *__restrict__ P, int64_t ldp, float th,
                        float *__restrict__ Pn, int64_t ldn, int *__restrict__ counts) {
    extern __shared__ float ebuf[];
    __shared__ double dred[32];
    __shared__ float fred[32];
    __shared__ int ired[32];
    const int64_t i = blockIdx.x;
    int64_t r = i + shift;
    if (r > rows_in - 1) r = rows_in - 1;
    const float *src = D + r * ld;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int64_t cv = vec ? (cols & ~int64_t(3)) : 0;
    // pass 1: row sum (fp64) and row max of E
    double s = 0.0;
    float mx = 0.f;                                   // E > 0 always
    auto one4 = [&](const float4 v, int64_t k) {
        const float e0 = trans_e(v.x, sigma), e1 = trans_e(v.y, sigma), e2 = trans_e(v.z, sigma),
                    e3 = trans_e(v.w, sigma);
        if (CACHE) *reinterpret_cast<float4 *>(ebuf + k) = make_float4(e0, e1, e2, e3);
        s += (double)e0 + (double)e1 + (double)e2 + (double)e3;
        mx = fmaxf(fmaxf(mx, fmaxf(e0, e1)), fmaxf(e2, e3));
    };
    int64_t k = int64_t(threadIdx.x) * 4;
    for (; k + NT * 4 < cv; k += 2 * NT * 4) {        // two 128-bit loads in flight
        const float4 v0 = CACHE ? ld_stream_f4(src + k) : *reinterpret_cast<const float4 *>(src + k);
        const float4 v1 = CACHE ? ld_stream_f4(src + k + NT * 4) : *reinterpret_cast<const float4 *>(src + k + NT * 4);
        one4(v0, k);
        one4(v1, k + NT * 4);
    }
    for (; k < cv; k += NT * 4) one4(CACHE ? ld_stream_f4(src + k) : *reinterpret_cast<const float4 *>(src + k), k);
    for (int64_t t = cv + threadIdx.x; t < cols; t += NT) {
        const float e = trans_e(src[t], sigma);
        if (CACHE) ebuf[t] = e;
        s += (double)e;
        mx = fmaxf(mx, e);
    }
    s = block_reduce(s, 0.0, OpAdd<double>(), dred);             // (also orders the ebuf writes)
    mx = block_reduce(mx, 0.f, OpMax(), fred);
    const float S = (float)s;
    const float pmax = __fdiv_rn(mx, S);                         // division is monotone: max P = fl(max E / S)
    const float cut = __fsub_rn(pmax, __fmul_rn(th, pmax));      // q_learning.py:63
    float *dp = (P != nullptr) ? P + i * ldp : nullptr;
    float *dn = (Pn != nullptr) ? Pn + i * ldn : nullptr;
    int cnt = 0;
    const bool vec_out = vec && (dp == nullptr || (reinterpret_cast<uintptr_t>(dp) & 15) == 0) &&
                         (dn == nullptr || (reinterpret_cast<uintptr_t>(dn) & 15) == 0);
    const int64_t cw = vec_out ? cv : 0;
    for (int64_t q = int64_t(threadIdx.x) * 4; q < cw; q += NT * 4) {
        float4 e;
        if (CACHE) e = *reinterpret_cast<const float4 *>(ebuf + q);
        else {
            const float4 v = *reinterpret_cast<const float4 *>(src + q);
            e = make_float4(trans_e(v.x, sigma), trans_e(v.y, sigma), trans_e(v.z, sigma), trans_e(v.w, sigma));
        }
        float4 pv = make_float4(__fdiv_rn(e.x, S), __fdiv_rn(e.y, S), __fdiv_rn(e.z, S), __fdiv_rn(e.w, S));
        if (dp != nullptr) *reinterpret_cast<float4 *>(dp + q) = pv;
        if (dn != nullptr) {
            pv.x = (pv.x < cut) ? 0.f : pv.x; pv.y = (pv.y < cut) ? 0.f : pv.y;
            pv.z = (pv.z < cut) ? 0.f : pv.z; pv.w = (pv.w < cut) ? 0.f : pv.w;
            *reinterpret_cast<float4 *>(dn + q) = pv;
        }
        cnt += (pv.x != 0.f) + (pv.y != 0.f) + (pv.z != 0.f) + (pv.w != 0.f);
    }
    for (int64_t q = cw + threadIdx.x; q < cols; q += NT) {
        const float pv = __fdiv_rn(CACHE ? ebuf[q] : trans_e(src[q], sigma), S);
        if (dp != nullptr) dp[q] = pv;
        float out = pv;
        if (dn != nullptr) {
            out = (pv < cut) ? 0.f : pv;
            dn[q] = out;
        }
        cnt += (out != 0.f);
    }
    if (counts != nullptr) {
        cnt = block_reduce(cnt, 0, OpAdd<int>(), ired);
        if (threadIdx.x == 0) counts[i] = cnt;
    }
}

__global__ void __launch_bounds__(PT)
row_nnz_kernel(const float *__restrict__ P, int64_t ld, int64_t cols, int *__restrict__ counts) {
    __shared__ int ired[32];
    const float *row = P + int64_t(blockIdx.x) * ld;
    int cnt = 0;
    for (int64_t k = threadIdx.x; k < cols; k += PT) cnt += (row[k] != 0.f);
    cnt = block_reduce(cnt, 0, OpAdd<int>(), ired);
    if (threadIdx.x == 0) counts[blockIdx.x] = cnt;
}

// Ordered (ascending column) compaction of the non-zeros of each row.
__global__ void __launch_bounds__(PT)
csr_fill_kernel(const float *__restrict__ P, int64_t ld, int64_t cols, const int64_t *__restrict__ rowptr,
                int *__restrict__ colidx) {
    __shared__ int wcount[PT / 32];
    __shared__ int base_s;
    const float *row = P + int64_t(blockIdx.x) * ld;
    int *dst = colidx + rowptr[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < cols; c0 += PT) {
        const int64_t c = c0 + threadIdx.x;
        const bool nz = (c < cols) && (row[c] != 0.f);
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) wcount[wid] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; ++w) off += wcount[w];
        if (nz) dst[off + __popc(bal & ((1u << lane) - 1u))] = (int)c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < PT / 32; ++w) t += wcount[w];
            base_s += t;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int avtex_sum_nnz(const float *D, int64_t rows, int64_t cols, int64_t ld, double *sum,
                             unsigned long long *nnz, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && sum != nullptr && nnz != nullptr,
                  "sum_nnz: bad arguments rows=%lld cols=%lld", (long long)rows, (long long)cols);
    const unsigned grid = (unsigned)(rows < 148 * 8 ? rows : 148 * 8);
    sum_nnz_kernel<<<grid, PT, 0, as_stream(stream)>>>(D, rows, cols, ld, sum, nnz);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_transition_probs(const float *D, int64_t ld, int64_t rows_in, int64_t cols,
                                      float sigma, int shift, int64_t rows_out, float *P, int64_t ldp,
                                      float th, float *P_new, int64_t ldn, int *counts, int device,
                                      void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows_in >= 1 && cols >= 1 && ld >= cols && rows_out >= 1 && shift >= 0,
                  "transition_probs: bad shape rows_in=%lld cols=%lld rows_out=%lld", (long long)rows_in,
                  (long long)cols, (long long)rows_out);
    AVTEX_REQUIRE(P == nullptr || ldp >= cols, "transition_probs: ldp too small");
    AVTEX_REQUIRE(P_new == nullptr || (ldn >= cols && th >= 0.f), "transition_probs: P_new needs ldn >= cols and th >= 0");
    AVTEX_REQUIRE(sigma > 0.f, "transition_probs: sigma must be positive (got %g)", (double)sigma);
    if (cols <= PROBS_SMEM_COLS) {
        const size_t smem = (size_t)cols * sizeof(float);
        static std::atomic<bool> attr_set[64];       // the attribute calls are idempotent; the flag only skips repeats
        if (device < 0 || device >= 64 || !attr_set[device].load(std::memory_order_acquire)) {
            AVTEX_CUDA(cudaFuncSetAttribute(transition_probs_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            PROBS_SMEM_COLS * (int)sizeof(float)));
            AVTEX_CUDA(cudaFuncSetAttribute(transition_probs_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            PROBS_SMEM_COLS * (int)sizeof(float)));
            if (device >= 0 && device < 64) attr_set[device].store(true, std::memory_order_release);
        }
        if (cols >= 8192)       // long rows: few CTAs fit per SM (shared memory), so make each one wide
            transition_probs_kernel<true, 1024><<<(unsigned)rows_out, 1024, smem, as_stream(stream)>>>(
                D, ld, rows_in, cols, sigma, shift, P, ldp, th, P_new, ldn, counts);
        else
            transition_probs_kernel<true, 256><<<(unsigned)rows_out, 256, smem, as_stream(stream)>>>(
                D, ld, rows_in, cols, sigma, shift, P, ldp, th, P_new, ldn, counts);
    } else {
        transition_probs_kernel<false, 1024><<<(unsigned)rows_out, 1024, 0, as_stream(stream)>>>(
            D, ld, rows_in, cols, sigma, shift, P, ldp, th, P_new, ldn, counts);
    }
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_row_nnz(const float *P, int64_t ld, int64_t rows, int64_t cols, int *counts,
                             int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && cols < (int64_t(1) << 31), "row_nnz: bad shape");
    row_nnz_kernel<<<(unsigned)rows, PT, 0, as_stream(stream)>>>(P, ld, cols, counts);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_csr_fill(const float *P, int64_t ld, int64_t rows, int64_t cols,
                              const int64_t *rowptr, int *colidx, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && cols < (int64_t(1) << 31), "csr_fill: bad shape");
    csr_fill_kernel<<<(unsigned)rows, PT, 0, as_stream(stream)>>>(P, ld, cols, rowptr, colidx);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

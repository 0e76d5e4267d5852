// K2 — diagonal temporal filter (Classic+/++) fused with the future-cost power:
//     D2[a,b] = sum_k w[k] * D1[a*s + k, b*s + k]          (classic/computeD2.py:34-42)
//     D3      = D2 ** p                                     (classic/q_learning.py:34)
// HBM-bound: reads 4 N^2 B (every D1 row is touched), writes 4 M^2 (+4 M^2) B.
//
// The reference runs a dense fs x fs cuDNN convolution whose kernel is 97.5 % zeros.  Here one thread
// walks ONE input diagonal and keeps R (16 at stride 1, 8 otherwise) consecutive outputs of that diagonal in registers, so it
// issues (R-1)*s + fs loads for R outputs instead of R*fs, and the taps are compile-time indexed
// kernel parameters (constant-bank FFMA operands: no shared-memory or LDS traffic at all).  Within
// a warp consecutive threads own consecutive diagonals, so every load is a coalesced row segment.
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int FT = 128;      // threads per CTA = diagonals per CTA

struct Taps64 { float w[64]; };
struct TapsBig { float w[960]; };

// Out-of-line so the 8/16 calls per thread do not multiply the (already fully unrolled) code size:
// the first version stalled on instruction fetch (ncu: stalled_no_instruction was the top reason).
__device__ __noinline__ float pow_pos_call(float x, float p) { return pow_pos(x, p); }

template <int FS, int S, int FR, bool PACKED>
__global__ void __launch_bounds__(FT)
diag_filter_kernel(const float *__restrict__ D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const Taps64 taps,
                   int64_t a0, int64_t rows_out, int64_t m, float *__restrict__ D2, int64_t ld2,
                   float *__restrict__ D3, int64_t ld3, float p, double *sum, unsigned long long *nnz) {
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    constexpr int T = (FR - 1) * S + FS;
    const int64_t n_in = (m - 1) * S + FS;                                  // valid input rows / cols
    const int64_t a_base = a0 + int64_t(blockIdx.y) * FR;                   // first output row of the band
    const int64_t b_blk = int64_t(blockIdx.x) * FT - (FR - 1);              // output col of thread 0's output 0
    const int64_t b0 = b_blk + threadIdx.x;
    const int64_t grow = a_base * S;                                        // global input row at t = 0
    const int64_t gcol = b0 * S;
    const float *src = D1 + (grow - in_row0) * ld1 + gcol;
    const int64_t step = ld1 + 1;
    float acc[FR];
#pragma unroll
    for (int i = 0; i < FR; ++i) acc[i] = 0.f;
    // CTA-uniform: every load of every thread is in range -> no per-load predicates (all CTAs except
    // those on the matrix border)
    const bool interior = (b_blk * S >= 0) && ((b_blk + FT - 1) * S + T - 1 < n_in) &&
                          (grow + T - 1 < in_row0 + in_rows) && (grow + T - 1 < n_in);
    if (interior && PACKED) {
        // Packed fp32x2 FMA (Blackwell FFMA2): outputs 2p and 2p+1 share one instruction.  At step t they
        // need taps q = t - 2p and q - 1, so the pair (w[q], w[q-1]) is kept as one 64-bit register value
        // (zero outside [0, FS)).  Same operation order per output as the scalar path -> identical bits,
        // half the FMA issue slots (this kernel is issue-bound at stride 1: 40 FMAs + pow per output).
        float2 wp[FS + 1];
#pragma unroll
        for (int q = 0; q <= FS; ++q)
            wp[q] = make_float2(q < FS ? taps.w[q] : 0.f, q >= 1 ? taps.w[q - 1] : 0.f);
        float2 acc2[FR / 2];
#pragma unroll
        for (int p2 = 0; p2 < FR / 2; ++p2) acc2[p2] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float x = __ldg(src);
            src += step;
            const float2 x2 = make_float2(x, x);
#pragma unroll
            for (int p2 = 0; p2 < FR / 2; ++p2) {
                const int q = t - 2 * p2;
                if (q >= 0 && q <= FS) acc2[p2] = __ffma2_rn(wp[q], x2, acc2[p2]);
            }
        }
#pragma unroll
        for (int p2 = 0; p2 < FR / 2; ++p2) { acc[2 * p2] = acc2[p2].x; acc[2 * p2 + 1] = acc2[p2].y; }
    } else if (interior) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float x = __ldg(src);
            src += step;
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const int kk = t - i * S;
                if (kk >= 0 && kk < FS) acc[i] = fmaf(taps.w[kk], x, acc[i]);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const bool ok = (gcol + t >= 0) && (gcol + t < n_in) && (grow + t < in_row0 + in_rows);
            const float x = ok ? __ldg(src) : 0.f;
            src += step;
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const int kk = t - i * S;
                if (kk >= 0 && kk < FS) acc[i] = fmaf(taps.w[kk], x, acc[i]);
            }
        }
    }
    double s = 0.0;
    int z = 0;
    float *o2 = D2 + (a_base - a0) * ld2 + b0;
    float *o3 = (D3 != nullptr) ? D3 + (a_base - a0) * ld3 + b0 : nullptr;
#pragma unroll
    for (int i = 0; i < FR; ++i) {
        const int64_t a = a_base + i, b = b0 + i;
        if (a < a0 + rows_out && b >= 0 && b < m) {
            o2[i * (ld2 + 1)] = acc[i];
            if (o3 != nullptr) o3[i * (ld3 + 1)] = pow_pos_call(acc[i], p);
            s += (double)acc[i];
            z += (acc[i] != 0.f);
        }
    }
    if (sum != nullptr) {
        const double sd = block_reduce(s, 0.0, OpAdd<double>(), sred);
        const unsigned long long zd = block_reduce((unsigned long long)z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, sd); atomicAdd(nnz, zd); }
    }
}

// ------------------------------------------------------------------ stride 1: sliding window from shared memory
// At stride 1 every input D1[r, c] feeds FS outputs of its diagonal, so the kernel above (16 outputs per
// thread, 55 loads, taps re-read from the constant bank) was ISSUE-bound at 0.34 of the HBM roofline.
// Here a thread walks its diagonal for a whole band of rows with a sliding window: FS/2 + 1 accumulator
// PAIRS rotate through compile-time register names (the loop body covers one full rotation, FS + 2 inputs),
// each input is read once from shared memory and contributes through FS/2 + 1 packed fp32x2 FMAs (FFMA2:
// two outputs per instruction); the tap pairs (w[q], w[q-1]) live in registers for the whole band.  Inputs
// are staged by cp.async in chunks of CH rows x (128 + CH) columns (6-stage ring, 128-bit, zero-filled
// outside the matrix), so the loads are decoupled from the FMA stream.  Per output: ~20.5 FFMA2 + 1 LDS
// instead of 40 FFMA + 3.4 LDG + constant-bank traffic.  The per-output operation order (k ascending
// fmaf chain from 0) is the same as diag_filter_kernel's, so the two kernels are bit-identical.
template <int FS>
struct S1Cfg {
    static constexpr int PER = FS + 2;                               // inputs per unrolled body
    static constexpr int NP = PER / 2;                               // accumulator pairs in rotation
    static constexpr int NCH = 3;                                    // smem chunks per body
    static constexpr int CH = PER / NCH;                             // rows per chunk
    static constexpr int PITCH = ((FT + CH + 3 + 3) / 4) * 4;        // floats per staged row
    static constexpr int VEC_PER_ROW = PITCH / 4;
    static constexpr int CHUNK_VECS = CH * VEC_PER_ROW;
    static constexpr int STAGES = 6;
    static constexpr int IN_BYTES = STAGES * CH * PITCH * 4;
    static constexpr int SMEM_BYTES = IN_BYTES + CH * FT * 4;        // + per-thread output staging [CH][FT]
    static_assert(FS % 2 == 0 && PER % NCH == 0, "sliding-window filter needs FS even and (FS + 2) % 3 == 0");
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, int bytes) {
    // copies `bytes` (0..16) from src and zero-fills the rest of the 16 destination bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int FS, bool STATS, bool POW>
__global__ void __launch_bounds__(FT, 3)
diag_filter_s1_kernel(const float *__restrict__ D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const Taps64 taps,
                      int64_t a0, int64_t rows_out, int64_t m, float *__restrict__ D2, int64_t ld2,
                      float *__restrict__ D3, int64_t ld3, float p, double *sum, unsigned long long *nnz, int nb) {
    using C = S1Cfg<FS>;
#ifndef AVTEX_S1_UR_PAIRS
#define AVTEX_S1_UR_PAIRS 0
#endif
    constexpr int UR_PAIRS = AVTEX_S1_UR_PAIRS < FS + 1 ? AVTEX_S1_UR_PAIRS : FS + 1;
    constexpr int KV = (C::CHUNK_VECS + FT - 1) / FT;                 // 128-bit copies per thread per chunk
    extern __shared__ __align__(16) float s1_smem[];
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    __shared__ float2 wp_s[FS + 1];
    const int64_t n_in = m - 1 + FS;                                  // valid input rows / cols (stride 1)
    const int band = C::PER * nb - FS;                                // output rows per band
    const int64_t a_band = a0 + int64_t(blockIdx.y) * band;
    const int64_t c_blk = int64_t(blockIdx.x) * FT - (band - 1);      // output column of thread 0 at i = 0
    const int64_t c_thread = c_blk + threadIdx.x;
    const int64_t row_lim = (in_row0 + in_rows < n_in) ? in_row0 + in_rows : n_in;
    const int total_chunks = nb * C::NCH;
    const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(s1_smem));

    // ---- staging: chunk j = input rows [a_band + j*CH, +CH) x columns [cb, cb + PITCH), cb = (c_blk + j*CH) & ~3.
    // The (row, column) a thread copies inside a chunk never changes, only the chunk origin does.
    uint32_t v_soff[KV];
    int64_t v_goff[KV];
#pragma unroll
    for (int k = 0; k < KV; ++k) {
        const int v = threadIdx.x + k * FT;
        const int rr = v / C::VEC_PER_ROW, cc = (v - rr * C::VEC_PER_ROW) * 4;
        v_soff[k] = uint32_t(rr * C::PITCH + cc) * 4u;
        v_goff[k] = int64_t(rr) * ld1 + cc;
    }
    auto issue_chunk = [&](int j) {
        if (j < total_chunks) {
            const int64_t r0 = a_band + int64_t(j) * C::CH;
            const int64_t cb = (c_blk + int64_t(j) * C::CH) & ~int64_t(3);
            const uint32_t dst0 = smem_base + uint32_t(j % C::STAGES) * (C::CH * C::PITCH * 4);
            const float *base = D1 + (r0 - in_row0) * ld1 + cb;
            if (r0 + C::CH <= row_lim && cb >= 0 && cb + C::PITCH <= n_in) {     // CTA-uniform: the whole chunk is inside
#pragma unroll
                for (int k = 0; k < KV; ++k)
                    if (k < KV - 1 || threadIdx.x + k * FT < C::CHUNK_VECS)
                        cp_async16_zfill(dst0 + v_soff[k], base + v_goff[k], 16);
            } else {
#pragma unroll
                for (int k = 0; k < KV; ++k)
                    if (k < KV - 1 || threadIdx.x + k * FT < C::CHUNK_VECS) {
                        // bytes beyond the copied prefix are ZERO-filled: rows past the block, columns < 0 and
                        // columns >= n_in (row padding may hold NaN bit patterns, and 0 * NaN would poison a valid
                        // output through the zero half of an edge tap pair)
                        const int v = threadIdx.x + k * FT;
                        const int rr = v / C::VEC_PER_ROW, cc = (v - rr * C::VEC_PER_ROW) * 4;
                        const int64_t r = r0 + rr, c = cb + cc;
                        int64_t cnt = (r < row_lim && c >= 0) ? n_in - c : 0;
                        cnt = cnt < 0 ? 0 : (cnt > 4 ? 4 : cnt);
                        cp_async16_zfill(dst0 + v_soff[k], cnt > 0 ? base + v_goff[k] : D1, int(cnt) * 4);
                    }
            }
        }
        cp_async_commit();                                             // empty groups keep the wait count uniform
    };
#pragma unroll
    for (int j = 0; j < C::STAGES - 1; ++j) issue_chunk(j);

    // ---- tap pairs (w[q], w[q-1]), q = 0..FS, held in REGISTERS for the whole band.  They are read back from
    // shared memory on purpose: values ptxas can trace to the constant bank get rematerialised into uniform
    // registers at every use (63 URs for 82 values), which cost more issue slots than the packing saved.
    if (threadIdx.x <= FS) {
        const int q = threadIdx.x;
        wp_s[q] = make_float2(q < FS ? taps.w[q] : 0.f, q >= 1 ? taps.w[q - 1] : 0.f);
    }
    __syncthreads();
    float2 wp[FS + 1];
#pragma unroll
    for (int q = 0; q <= FS; ++q) {
        if (q < UR_PAIRS) {            // constant-bank values: ptxas keeps these pairs in UNIFORM registers (no vector registers)
            wp[q] = make_float2(q < FS ? taps.w[q] : 0.f, q >= 1 ? taps.w[q - 1] : 0.f);
        } else {
            wp[q].x = *reinterpret_cast<volatile float *>(&wp_s[q].x);
            wp[q].y = *reinterpret_cast<volatile float *>(&wp_s[q].y);
        }
    }
    float2 acc[C::NP];
#pragma unroll
    for (int sidx = 0; sidx < C::NP; ++sidx) acc[sidx] = make_float2(0.f, 0.f);

    // ---- outputs: local index i <-> (a_band + i, c_thread + i); valid i in [i_lo, i_lo + i_span).  The pair
    // completed at input step t holds outputs i = t - FS and t - FS + 1.  Completed values are parked in a
    // per-thread column of shared memory (compile-time offsets) and written out by a small ROLLED loop after each
    // chunk: predicate, D2 store, pow, D3 store and the sigma statistics exist once in the code and their
    // addresses advance by ld + 1 per output (the fully unrolled form re-derived every address with 64-bit
    // multiplies and re-read the parameters: ~45 instructions per pair, measured).
    const int64_t lim_i = ((a0 + rows_out < a_band + band) ? a0 + rows_out : a_band + band) - a_band;
    const int64_t lo64 = c_thread < 0 ? -c_thread : 0, hi64 = (m - c_thread < lim_i) ? m - c_thread : lim_i;
    const int i_lo = int(lo64);
    const unsigned i_span = hi64 > lo64 ? unsigned(hi64 - lo64) : 0u;
    const int64_t step2 = ld2 + 1, step3 = ld3 + 1;
    float *o2 = D2 + (a_band - a0) * ld2 + c_thread - int64_t(FS) * step2;     // i = -FS (never dereferenced while invalid)
    float *o3 = POW ? D3 + (a_band - a0) * ld3 + c_thread - int64_t(FS) * step3 : nullptr;
    float *out_s = s1_smem + C::IN_BYTES / 4 + threadIdx.x;                     // out_s[r * FT]: my value of chunk row r
    int i0 = -FS;
    double s = 0.0;
    int z = 0;

#pragma unroll 1
    for (int body = 0; body < nb; ++body) {
        float fs = 0.f;
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) {
            const int j = body * C::NCH + ch;
            cp_async_wait<C::STAGES - 2>();                            // chunk j has landed (this thread's part)
            __syncthreads();                                           // ... everyone's part; chunk j-1 is fully consumed
            issue_chunk(j + C::STAGES - 1);                            // refill the stage chunk j-1 occupied
            const int off = int((c_blk + int64_t(j) * C::CH) & 3);
            const float *rowp = s1_smem + (j % C::STAGES) * (C::CH * C::PITCH) + threadIdx.x + off;
#pragma unroll
            for (int r = 0; r < C::CH; ++r) {
                const int tb = ch * C::CH + r;                         // compile-time step inside the body
                const float x = rowp[r * C::PITCH + r];
                const float2 x2 = make_float2(x, x);
#pragma unroll
                for (int sidx = 0; sidx < C::NP; ++sidx) {
                    const int q = (tb - 2 * sidx + 2 * C::PER) % C::PER;          // tap-pair index for slot sidx at this step
                    if (q <= FS) acc[sidx] = __ffma2_rn(wp[q], x2, acc[sidx]);
                    if (q == FS) {                                     // slot complete: rows r and r + 1 of this chunk's outputs
                        static_assert(C::CH % 2 == 0, "pairs must not straddle chunks");
                        out_s[r * FT] = acc[sidx].x;
                        out_s[(r + 1) * FT] = acc[sidx].y;
                        acc[sidx] = make_float2(0.f, 0.f);
                    }
                }
            }
#pragma unroll 2
            for (int r = 0; r < C::CH; ++r) {
                const float v = out_s[r * FT];
                if (unsigned(i0 - i_lo) < i_span) {
                    *o2 = v;
                    if (POW) *o3 = pow_pos(v, p);
                    if (STATS) { fs += v; z += (v != 0.f); }
                }
                o2 += step2;
                if (POW) o3 += step3;
                ++i0;
            }
        }
        s += (double)fs;
    }
    cp_async_wait<0>();
    if (STATS) {
        const double sd = block_reduce(s, 0.0, OpAdd<double>(), sred);
        const unsigned long long zd = block_reduce((unsigned long long)z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, sd); atomicAdd(nnz, zd); }
    }
}

template <int FS>
int launch_s1(const float *D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const float *h_w, int64_t a0,
              int64_t rows_out, int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p, double *sum,
              unsigned long long *nnz, int device, cudaStream_t st) {
    using C = S1Cfg<FS>;
    Taps64 taps;
    for (int i = 0; i < 64; ++i) taps.w[i] = (i < FS) ? h_w[i] : 0.f;
    static std::atomic<bool> attr_set[64];
    if (device < 0 || device >= 64 || !attr_set[device].load(std::memory_order_acquire)) {
        AVTEX_CUDA(cudaFuncSetAttribute(diag_filter_s1_kernel<FS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        AVTEX_CUDA(cudaFuncSetAttribute(diag_filter_s1_kernel<FS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        AVTEX_CUDA(cudaFuncSetAttribute(diag_filter_s1_kernel<FS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        AVTEX_CUDA(cudaFuncSetAttribute(diag_filter_s1_kernel<FS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        if (device >= 0 && device < 64) attr_set[device].store(true, std::memory_order_release);
    }
    // bodies per thread: long bands amortise the FS-1 warm-up rows (band / (band + FS)), short ones keep enough
    // CTAs in flight on small matrices
    int64_t nb = (rows_out / 6 + FS + C::PER - 1) / C::PER;
    if (nb < 2) nb = 2;
    if (nb > 12) nb = 12;
    const int64_t band = int64_t(C::PER) * nb - FS;
    dim3 grid((unsigned)((m + band - 1 + FT - 1) / FT), (unsigned)((rows_out + band - 1) / band));
#define AVTEX_S1(ST_, PW_)                                                                                          \
    diag_filter_s1_kernel<FS, ST_, PW_><<<grid, FT, C::SMEM_BYTES, st>>>(D1, ld1, in_row0, in_rows, taps, a0, rows_out, m, \
                                                                         D2, ld2, D3, ld3, p, sum, nnz, (int)nb)
    if (sum != nullptr) { if (D3 != nullptr) AVTEX_S1(true, true); else AVTEX_S1(true, false); }
    else { if (D3 != nullptr) AVTEX_S1(false, true); else AVTEX_S1(false, false); }
#undef AVTEX_S1
    return 0;
}

// Any (fs, stride): one output per thread, taps read from the parameter bank with a runtime index.
__global__ void __launch_bounds__(FT)
diag_filter_generic_kernel(const float *__restrict__ D1, int64_t ld1, int64_t in_row0, const TapsBig taps,
                           int fs, int stride, int64_t a0, int64_t rows_out, int64_t m,
                           float *__restrict__ D2, int64_t ld2, float *__restrict__ D3, int64_t ld3,
                           float p, double *sum, unsigned long long *nnz) {
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    const int64_t a = a0 + blockIdx.y;
    const int64_t b = int64_t(blockIdx.x) * FT + threadIdx.x;
    double s = 0.0;
    unsigned long long z = 0;
    if (b < m) {
        const float *src = D1 + (a * stride - in_row0) * ld1 + b * stride;
        float acc = 0.f;
        for (int k = 0; k < fs; ++k) acc = fmaf(taps.w[k], __ldg(src + int64_t(k) * (ld1 + 1)), acc);
        D2[(a - a0) * ld2 + b] = acc;
        if (D3 != nullptr) D3[(a - a0) * ld3 + b] = pow_pos(acc, p);
        s = acc;
        z = (acc != 0.f);
    }
    if (sum != nullptr) {
        s = block_reduce(s, 0.0, OpAdd<double>(), sred);
        z = block_reduce(z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, s); atomicAdd(nnz, z); }
    }
}

// outputs per thread along its diagonal: (FR-1)*S+FS loads serve FR outputs
template <int S> constexpr int filter_r() { return S == 1 ? 16 : 8; }

template <int FS, int S>
void launch_fast(const float *D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const float *h_w, int64_t a0,
                 int64_t rows_out, int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                 double *sum, unsigned long long *nnz, cudaStream_t st) {
    Taps64 taps;
    for (int i = 0; i < 64; ++i) taps.w[i] = (i < FS) ? h_w[i] : 0.f;
    constexpr int FR = filter_r<S>();
    dim3 grid((unsigned)((m + FR - 1 + FT - 1) / FT), (unsigned)((rows_out + FR - 1) / FR));
    static const bool packed_off = []() { const char *e = getenv("AVTEX_FILTER_FFMA2"); return e != nullptr && e[0] == '0'; }();
    if (S == 1 && !packed_off)
        diag_filter_kernel<FS, S, FR, S == 1><<<grid, FT, 0, st>>>(D1, ld1, in_row0, in_rows, taps, a0, rows_out, m, D2, ld2, D3,
                                                           ld3, p, sum, nnz);
    else
    diag_filter_kernel<FS, S, FR, false><<<grid, FT, 0, st>>>(D1, ld1, in_row0, in_rows, taps, a0, rows_out, m, D2, ld2, D3,
                                                   ld3, p, sum, nnz);
}

// out = D ** p elementwise (classic/q_learning.py:34 when D2 comes from the caller, not from the fused filter).
__global__ void __launch_bounds__(256)
pow_matrix_kernel(const float *__restrict__ D, int64_t ld, int64_t rows, int64_t cols, float p, float *__restrict__ out,
                  int64_t ld_out) {
    const bool vec = (ld % 4 == 0) && (ld_out % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(D) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const float *src = D + r * ld;
        float *dst = out + r * ld_out;
        const int64_t cv = vec ? (cols & ~int64_t(3)) : 0;
        for (int64_t c = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 4; c < cv; c += int64_t(gridDim.x) * 1024) {
            float4 v = ld_stream_f4(src + c);
            v.x = pow_pos(v.x, p); v.y = pow_pos(v.y, p); v.z = pow_pos(v.z, p); v.w = pow_pos(v.w, p);
            *reinterpret_cast<float4 *>(dst + c) = v;
        }
        for (int64_t c = cv + int64_t(blockIdx.x) * 256 + threadIdx.x; c < cols; c += int64_t(gridDim.x) * 256)
            dst[c] = pow_pos(src[c], p);
    }
}

}  // namespace

extern "C" int avtex_pow_matrix(const float *D, int64_t ld, int64_t rows, int64_t cols, float p, float *out,
                                int64_t ld_out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && ld_out >= cols && p > 0.f, "pow_matrix: bad arguments");
    dim3 grid((unsigned)((cols + 1023) / 1024), (unsigned)(rows < 65535 ? rows : 65535));
    pow_matrix_kernel<<<grid, 256, 0, as_stream(stream)>>>(D, ld, rows, cols, p, out, ld_out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_diag_filter_pow(const float *D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const float *h_w,
                                     int fs, int stride, int64_t a0, int64_t rows_out, int64_t m,
                                     float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                                     double *sum, unsigned long long *nnz, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(fs >= 1 && fs <= 960 && stride >= 1, "diag_filter: fs=%d stride=%d unsupported", fs, stride);
    AVTEX_REQUIRE(m >= 1 && rows_out >= 1 && a0 >= 0 && a0 + rows_out <= m && ld2 >= m,
                  "diag_filter: bad output shape a0=%lld rows=%lld m=%lld", (long long)a0,
                  (long long)rows_out, (long long)m);
    AVTEX_REQUIRE(ld1 >= (m - 1) * stride + fs, "diag_filter: ld1=%lld too small", (long long)ld1);
    AVTEX_REQUIRE(in_row0 >= 0 && in_row0 <= a0 * stride &&
                      in_row0 + in_rows >= (a0 + rows_out - 1) * stride + fs,
                  "diag_filter: D1 rows [%lld, %lld) do not cover the rows needed", (long long)in_row0,
                  (long long)(in_row0 + in_rows));
    AVTEX_REQUIRE((sum == nullptr) == (nnz == nullptr), "diag_filter: sum and nnz go together");
    AVTEX_REQUIRE(D3 == nullptr || ld3 >= m, "diag_filter: ld3 too small");
    cudaStream_t st = as_stream(stream);
    {
        static const bool s1_off = []() { const char *e = getenv("AVTEX_FILTER_S1"); return e != nullptr && e[0] == '0'; }();
        const bool aligned = (ld1 % 4 == 0) && ((reinterpret_cast<uintptr_t>(D1) & 15) == 0);
        if (stride == 1 && aligned && !s1_off && (fs == 40 || fs == 16)) {
            int rc = fs == 40 ? launch_s1<40>(D1, ld1, in_row0, in_rows, h_w, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz, device, st)
                              : launch_s1<16>(D1, ld1, in_row0, in_rows, h_w, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz, device, st);
            if (rc) return rc;
            AVTEX_LAUNCH_CHECK();
            return 0;
        }
    }
    const int key = fs * 100 + stride;
#define AVTEX_FAST(FS_, S_)                                                                            \
    case FS_ * 100 + S_:                                                                               \
        AVTEX_REQUIRE((rows_out + filter_r<S_>() - 1) / filter_r<S_>() <= 65535, "diag_filter: too many row bands"); \
        launch_fast<FS_, S_>(D1, ld1, in_row0, in_rows, h_w, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz, st); \
        break;
    switch (key) {
        AVTEX_FAST(40, 1)
        AVTEX_FAST(40, 4)
        AVTEX_FAST(16, 1)
        AVTEX_FAST(16, 4)
        AVTEX_FAST(8, 1)
        default: {
            TapsBig taps;
            for (int i = 0; i < 960; ++i) taps.w[i] = (i < fs) ? h_w[i] : 0.f;
            for (int64_t r = 0; r < rows_out; r += 65535) {          // gridDim.y <= 65535: row bands of that many rows
                const int64_t nr = rows_out - r < 65535 ? rows_out - r : 65535;
                dim3 grid((unsigned)((m + FT - 1) / FT), (unsigned)nr);
                diag_filter_generic_kernel<<<grid, FT, 0, st>>>(D1, ld1, in_row0, taps, fs, stride, a0 + r, nr, m,
                                                                D2 + r * ld2, ld2, D3 != nullptr ? D3 + r * ld3 : nullptr,
                                                                ld3, p, sum, nnz);
            }
        }
    }
#undef AVTEX_FAST
    AVTEX_LAUNCH_CHECK();
    return 0;
}

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
//
// Stride 1 (-m 1 / -m 2) is bound by the FP32 pipe and instruction issue, not by HBM: 40 FMAs per output = 0.32 ms of
// pure FMA-pipe time at N = 16000 against 0.47 ms of HBM time.  Round 2: (1) the FFMA2 tap pairs are built on the HOST
// and passed as aligned kernel parameters (TapPairs) — assembling them in the kernel cost two UMOVs per odd pair, more
// feeding instructions than FMAs; (2) a CTA-uniform all-outputs-inside path without range checks, pow inlined:
// 1151 M -> 697 M warp instructions, 1.11 -> 0.75 ms (0.42 -> 0.62 of the HBM roofline); (3) the SYMMETRIC form below
// halves the work when D1 is a distance matrix (0.64-0.69 ms at stride 1, 1.67x at stride 4).  The shared-memory staged
// sliding-window design the north star names (cp.async ring, tap pairs in registers, also warp-specialised) was
// built, bit-identical, and SLOWER (1.28-1.40 ms: 124 live registers leave 12-16 warps per SM); archived under
// profiles/r02_filter_sliding_window_experiment.*; DESIGN.md §4.2.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int FT = 128;      // threads per CTA = diagonals per CTA

struct Taps64 { float w[64]; };
struct TapsBig { float w[960]; };
// Packed-FMA operand pairs (w[q], w[q-1]), q = 0..FS (zero outside [0, FS)), built on the host and passed as
// 8-byte-aligned kernel parameters: FFMA2 then reads each pair straight from a uniform-register pair filled by one
// LDCU.  Building the pairs in the kernel from the scalar taps made ptxas assemble every odd pair with two UMOVs
// (366 UMOV + 293 LDCU for 328 FFMA2 per thread: more feeding instructions than FMAs).
struct TapPairs { float2 wp[66]; };

// Out-of-line so the 8/16 calls per thread do not multiply the (already fully unrolled) code size:
// the first version stalled on instruction fetch (ncu: stalled_no_instruction was the top reason).
__device__ __noinline__ float pow_pos_call(float x, float p) { return pow_pos(x, p); }

// SYM: D1 is symmetric (every distance matrix is), hence so is D2 — bit for bit, because D2[a,b] and D2[b,a] add the
// same values in the same order.  Only outputs on or above the diagonal are computed (a thread's outputs all have
// b - a = b0 - a_base, so ownership is per thread; CTAs entirely below the diagonal exit at once) and the strictly
// upper ones are mirrored through a shared-memory tile so that the mirrored rows are written as 64-byte runs:
// half the loads, FMAs and pows (stride 1 is FP32-pipe bound) and half the D1 bytes (stride 4 is HBM bound).
// RES: D1 is given as S residue-class matrices (avtex_diag_filter_pow_res): input element (g, h) with g = h (mod S) lives
// in plane g % S at [g / S, h / S]; the walk along a diagonal visits the planes round-robin, in the same tap order.
template <int FS, int S, int FR, bool PACKED, bool SYM, bool RES>
__global__ void __launch_bounds__(FT)
diag_filter_kernel(const float *__restrict__ D1, int64_t ld1, int64_t plane, int64_t in_row0, int64_t in_rows, const Taps64 taps,
                   const TapPairs pairs, int64_t a0, int64_t rows_out, int64_t m, float *__restrict__ D2, int64_t ld2,
                   float *__restrict__ D3, int64_t ld3, float p, double *sum, unsigned long long *nnz) {
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    constexpr int TR = FT + FR - 1;                                         // mirrored rows touched by one CTA
    constexpr int TP = FR + 1;                                              // tile pitch (conflict-free writes)
    __shared__ float tile2[SYM ? TR * TP : 1];
    __shared__ float tile3[SYM ? TR * TP : 1];
    constexpr int T = (FR - 1) * S + FS;
    const int64_t n_in = (m - 1) * S + FS;                                  // valid input rows / cols
    int bx = blockIdx.x, by = blockIdx.y;
    if (SYM) {
        // Band y needs the column blocks x >= y / (FT / FR) only (the others lie below the diagonal).  The grid is the
        // triangle folded in two: grid row f serves band f with its first n1 CTAs and band GY-1-f with the rest, so
        // hardly any CTA is launched just to exit.
        constexpr int R = FT / FR;
        const int GX = int((m + FR - 1 + FT - 1) / FT), GY = int((rows_out + FR - 1) / FR);
        const int y2 = GY - 1 - by, n1 = GX - by / R;
        if (bx < n1) {
            bx += by / R;
        } else {
            bx -= n1;
            if (y2 <= by || bx >= GX - y2 / R) return;
            bx += y2 / R;
            by = y2;
        }
    }
    const int64_t a_base = a0 + int64_t(by) * FR;                           // first output row of the band
    const int64_t b_blk = int64_t(bx) * FT - (FR - 1);                      // output col of thread 0's output 0
    if (SYM && b_blk + FT + FR - 2 < a_base) return;                        // the whole CTA lies below the diagonal
    const int64_t b0 = b_blk + threadIdx.x;
    const bool owner = !SYM || b0 >= a_base;                                // this thread's outputs have b >= a
    const bool mirror = SYM && b0 > a_base;                                 // ... b > a: also stored as D2[b, a]
    const int64_t grow = a_base * S;                                        // global input row at t = 0
    const int64_t gcol = b0 * S;
    const float *src = RES ? D1 + (a_base - in_row0) * ld1 + b0 : D1 + (grow - in_row0) * ld1 + gcol;   // RES: class rows
    const int64_t step = ld1 + 1;
    const int64_t in_end = RES ? (in_row0 + in_rows) * S : in_row0 + in_rows;   // one past the last (global) input row held
    // address of the input element at diagonal step t
    auto at = [&](int t) -> const float * { return RES ? src + (t % S) * plane + (t / S) * step : src + t * step; };
    float acc[FR];
#pragma unroll
    for (int i = 0; i < FR; ++i) acc[i] = 0.f;
    // CTA-uniform: every load of every thread is in range -> no per-load predicates (all CTAs except
    // those on the matrix border)
    const bool interior = (b_blk * S >= 0) && ((b_blk + FT - 1) * S + T - 1 < n_in) &&
                          (grow + T - 1 < in_end) && (grow + T - 1 < n_in);
    if (!owner) {
        // nothing to compute: the mirror image of these outputs is owned by another CTA
    } else if (interior && PACKED) {
        // Packed fp32x2 FMA (Blackwell FFMA2): outputs 2p and 2p+1 share one instruction.  At step t they
        // need taps q = t - 2p and q - 1: the pair (w[q], w[q-1]) comes from the host-built TapPairs (zero outside
        // [0, FS)).  Same operation order per output as the scalar path -> identical bits, half the FMA issue slots.
        float2 acc2[FR / 2];
#pragma unroll
        for (int p2 = 0; p2 < FR / 2; ++p2) acc2[p2] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float x = __ldg(at(t));
            const float2 x2 = make_float2(x, x);
#pragma unroll
            for (int p2 = 0; p2 < FR / 2; ++p2) {
                const int q = t - 2 * p2;
                if (q >= 0 && q <= FS) acc2[p2] = __ffma2_rn(pairs.wp[q], x2, acc2[p2]);
            }
        }
#pragma unroll
        for (int p2 = 0; p2 < FR / 2; ++p2) { acc[2 * p2] = acc2[p2].x; acc[2 * p2 + 1] = acc2[p2].y; }
    } else if (interior) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float x = __ldg(at(t));
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const int kk = t - i * S;
                if (kk >= 0 && kk < FS) acc[i] = fmaf(taps.w[kk], x, acc[i]);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const bool ok = (gcol + t >= 0) && (gcol + t < n_in) && (grow + t < in_end);
            const float x = ok ? __ldg(at(t)) : 0.f;
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const int kk = t - i * S;
                if (kk >= 0 && kk < FS) acc[i] = fmaf(taps.w[kk], x, acc[i]);
            }
        }
    }
    float s = 0.f;                                                          // fp32 partial over this thread's FR outputs
    int z = 0;
    float *o2 = D2 + (a_base - a0) * ld2 + b0;
    float *o3 = (D3 != nullptr) ? D3 + (a_base - a0) * ld3 + b0 : nullptr;
    float *m2 = tile2 + threadIdx.x * TP, *m3 = tile3 + threadIdx.x * TP;   // element i of this thread -> tile[tid + i][i]
    // CTA-uniform: every output of every thread is inside the matrix -> no per-output range checks, pow inlined
    // (the hot path then is ~16 KB of straight-line code; border CTAs keep the out-of-line pow)
    const bool all_out = interior && (a_base + FR <= a0 + rows_out) && (b_blk >= 0) && (b_blk + FT + FR - 2 < m);
    if (!owner) {
    } else if (all_out) {
        const int64_t st2 = ld2 + 1, st3 = ld3 + 1;
        if (o3 != nullptr) {
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const float pw = pow_pos(acc[i], p);
                *o2 = acc[i];
                *o3 = pw;
                o2 += st2;
                o3 += st3;
                if (SYM) { m2[i * (TP + 1)] = acc[i]; m3[i * (TP + 1)] = pw; }
                s += acc[i];
                z += (acc[i] != 0.f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                *o2 = acc[i];
                o2 += st2;
                if (SYM) m2[i * (TP + 1)] = acc[i];
                s += acc[i];
                z += (acc[i] != 0.f);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < FR; ++i) {
            const int64_t a = a_base + i, b = b0 + i;
            if (a < a0 + rows_out && b >= 0 && b < m) {
                o2[i * (ld2 + 1)] = acc[i];
                if (SYM) m2[i * (TP + 1)] = acc[i];
                if (o3 != nullptr) {
                    const float pw = pow_pos_call(acc[i], p);
                    o3[i * (ld3 + 1)] = pw;
                    if (SYM) m3[i * (TP + 1)] = pw;
                }
                s += acc[i];
                z += (acc[i] != 0.f);
            }
        }
    }
    if (SYM) {
        if (mirror) { s *= 2.f; z *= 2; }                                   // the mirrored copies count too
        __syncthreads();
        // mirrored rows b = b_blk + r, columns a_base .. a_base + FR - 1: element i of row r came from thread r - i
        const int64_t a_end = a0 + rows_out;
        const bool vec_ok = (ld2 % 4 == 0) && (ld3 % 4 == 0) && (FR % 4 == 0) &&
                            (((reinterpret_cast<uintptr_t>(D2) | reinterpret_cast<uintptr_t>(D3)) & 15) == 0);
        for (int idx = threadIdx.x; idx < TR * (FR / 4); idx += FT) {
            const int r = idx / (FR / 4), i0 = (idx % (FR / 4)) * 4;
            const int64_t b = b_blk + r;
            if (b < 0 || b >= m) continue;
            const float *t2 = tile2 + r * TP + i0, *t3 = tile3 + r * TP + i0;
            float *d2 = D2 + (b - a0) * ld2 + a_base + i0;
            float *d3 = (D3 != nullptr) ? D3 + (b - a0) * ld3 + a_base + i0 : nullptr;
            const bool full = vec_ok && (r - (i0 + 3) >= 0) && (r - i0 < FT) && (b - (i0 + 3) > a_base) &&
                              (a_base + i0 + 3 < a_end);
            if (full) {
                *reinterpret_cast<float4 *>(d2) = make_float4(t2[0], t2[1], t2[2], t2[3]);
                if (d3 != nullptr) *reinterpret_cast<float4 *>(d3) = make_float4(t3[0], t3[1], t3[2], t3[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i0 + e;
                    if (r - i >= 0 && r - i < FT && b - i > a_base && a_base + i < a_end) {
                        d2[e] = t2[e];
                        if (d3 != nullptr) d3[e] = t3[e];
                    }
                }
            }
        }
    }
    if (sum != nullptr) {
        const double sd = block_reduce((double)s, 0.0, OpAdd<double>(), sred);
        const unsigned long long zd = block_reduce((unsigned long long)z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, sd); atomicAdd(nnz, zd); }
    }
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

template <int FS, int S, int FR = filter_r<S>()>
void launch_fast(const float *D1, int64_t ld1, int64_t plane, int64_t in_row0, int64_t in_rows, const float *h_w, int64_t a0,
                 int64_t rows_out, int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                 double *sum, unsigned long long *nnz, bool symmetric, cudaStream_t st) {
    Taps64 taps;
    for (int i = 0; i < 64; ++i) taps.w[i] = (i < FS) ? h_w[i] : 0.f;
    TapPairs pairs;
    for (int q = 0; q < 66; ++q) pairs.wp[q] = make_float2(q < FS ? h_w[q] : 0.f, (q >= 1 && q <= FS) ? h_w[q - 1] : 0.f);
    dim3 grid((unsigned)((m + FR - 1 + FT - 1) / FT), (unsigned)((rows_out + FR - 1) / FR));
    static const bool packed_off = []() { const char *e = getenv("AVTEX_FILTER_FFMA2"); return e != nullptr && e[0] == '0'; }();
#define AVTEX_FILTER_LAUNCH(PACKED_, SYM_)                                                                         \
    do {                                                                                                            \
        if (plane != 0)                                                                                             \
            diag_filter_kernel<FS, S, FR, PACKED_, SYM_, (S > 1)><<<grid, FT, 0, st>>>(                             \
                D1, ld1, plane, in_row0, in_rows, taps, pairs, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz);     \
        else                                                                                                        \
            diag_filter_kernel<FS, S, FR, PACKED_, SYM_, false><<<grid, FT, 0, st>>>(                               \
                D1, ld1, 0, in_row0, in_rows, taps, pairs, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz);         \
    } while (0)
    const bool packed = (S == 1) && !packed_off;
    if (symmetric) {                                                 // folded triangle (see the kernel)
        static_assert(FT % FR == 0, "the symmetric grid needs FT to be a multiple of FR");
        const int GX = (int)grid.x, GY = (int)grid.y, R = FT / FR;
        int w = 0;
        for (int f = 0; f < (GY + 1) / 2; ++f) {
            const int y2 = GY - 1 - f;
            const int n = (GX - f / R) + (y2 > f ? (GX - y2 / R > 0 ? GX - y2 / R : 0) : 0);
            w = n > w ? n : w;
        }
        grid = dim3((unsigned)w, (unsigned)((GY + 1) / 2));
    }
    if (symmetric) { if (packed) AVTEX_FILTER_LAUNCH(S == 1, true); else AVTEX_FILTER_LAUNCH(false, true); }
    else           { if (packed) AVTEX_FILTER_LAUNCH(S == 1, false); else AVTEX_FILTER_LAUNCH(false, false); }
#undef AVTEX_FILTER_LAUNCH
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

namespace {

int diag_filter_impl(const float *D1, int64_t ld1, int64_t plane, int64_t in_row0, int64_t in_rows, const float *h_w, int fs,
                     int stride,
                     int64_t a0, int64_t rows_out, int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                     double *sum, unsigned long long *nnz, bool symmetric, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(fs >= 1 && fs <= 960 && stride >= 1, "diag_filter: fs=%d stride=%d unsupported", fs, stride);
    AVTEX_REQUIRE(m >= 1 && rows_out >= 1 && a0 >= 0 && a0 + rows_out <= m && ld2 >= m,
                  "diag_filter: bad output shape a0=%lld rows=%lld m=%lld", (long long)a0,
                  (long long)rows_out, (long long)m);
    if (plane != 0) {                                                // residue-class planes; in_row0 / in_rows in CLASS rows
        const int64_t need_cols = ((m - 1) * stride + fs + stride - 1) / stride;     // columns of every class matrix
        const int64_t last_row = a0 + rows_out - 1 + (fs - 1) / stride;              // last class row read
        AVTEX_REQUIRE(stride >= 2 && in_row0 >= 0 && in_row0 <= a0 && in_row0 + in_rows > last_row &&
                          ld1 >= need_cols && plane >= in_rows * ld1,
                      "diag_filter (residue form): needs stride >= 2 and class rows [%lld, %lld] x %lld columns",
                      (long long)a0, (long long)last_row, (long long)need_cols);
    } else {
        AVTEX_REQUIRE(ld1 >= (m - 1) * stride + fs, "diag_filter: ld1=%lld too small", (long long)ld1);
        AVTEX_REQUIRE(in_row0 >= 0 && in_row0 <= a0 * stride &&
                          in_row0 + in_rows >= (a0 + rows_out - 1) * stride + fs,
                      "diag_filter: D1 rows [%lld, %lld) do not cover the rows needed", (long long)in_row0,
                      (long long)(in_row0 + in_rows));
    }
    AVTEX_REQUIRE((sum == nullptr) == (nnz == nullptr), "diag_filter: sum and nnz go together");
    AVTEX_REQUIRE(D3 == nullptr || ld3 >= m, "diag_filter: ld3 too small");
    AVTEX_REQUIRE(!symmetric || (a0 == 0 && rows_out == m && in_row0 == 0),
                  "diag_filter: the symmetric form needs the whole matrix (a0 = 0, rows_out = m, in_row0 = 0)");
    cudaStream_t st = as_stream(stream);
    const int key = fs * 100 + stride;
#define AVTEX_FAST(FS_, S_)                                                                            \
    case FS_ * 100 + S_:                                                                               \
        AVTEX_REQUIRE((rows_out + filter_r<S_>() - 1) / filter_r<S_>() <= 65535, "diag_filter: too many row bands"); \
        launch_fast<FS_, S_>(D1, ld1, plane, in_row0, in_rows, h_w, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz, symmetric, st); \
        break;
    switch (key) {
        AVTEX_FAST(40, 1)
        AVTEX_FAST(40, 4)
        AVTEX_FAST(16, 1)
        AVTEX_FAST(16, 4)
        AVTEX_FAST(8, 1)
        default: {                                                   // any (fs, stride): no symmetric shortcut
            AVTEX_REQUIRE(plane == 0, "diag_filter (residue form): no register-resident kernel for fs=%d stride=%d", fs,
                          stride);
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

}  // namespace

extern "C" int avtex_diag_filter_pow(const float *D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const float *h_w,
                                     int fs, int stride, int64_t a0, int64_t rows_out, int64_t m,
                                     float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                                     double *sum, unsigned long long *nnz, int device, void *stream) {
    return diag_filter_impl(D1, ld1, 0, in_row0, in_rows, h_w, fs, stride, a0, rows_out, m, D2, ld2, D3, ld3, p, sum, nnz,
                            false, device, stream);
}

extern "C" int avtex_diag_filter_pow_sym(const float *D1, int64_t ld1, int64_t n_rows, const float *h_w, int fs, int stride,
                                         int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                                         double *sum, unsigned long long *nnz, int device, void *stream) {
    return diag_filter_impl(D1, ld1, 0, 0, n_rows, h_w, fs, stride, 0, m, m, D2, ld2, D3, ld3, p, sum, nnz, true, device,
                            stream);
}

extern "C" int avtex_diag_filter_pow_res(const float *D1r, int64_t ld1, int64_t plane, int64_t in_row0, int64_t in_rows,
                                         const float *h_w, int fs, int stride, int64_t a0, int64_t rows_out, int64_t m,
                                         float *D2, int64_t ld2, float *D3, int64_t ld3, float p, double *sum,
                                         unsigned long long *nnz, int symmetric, int device, void *stream) {
    AVTEX_REQUIRE(plane > 0, "diag_filter (residue form): plane stride must be positive");
    return diag_filter_impl(D1r, ld1, plane, in_row0, in_rows, h_w, fs, stride, a0, rows_out, m, D2, ld2, D3, ld3, p, sum,
                            nnz, symmetric != 0, device, stream);
}

// K6 / K7 — contrastive synthesis step at the embedding boundary.
//   K6  L2-normalise rows (cvt/models/models.py:351,412) and one query against all L windows,
//       out[w] = <q^, t^_w> / temp (models.py:416-417; driving-audio term :433-439,457).
//       Streaming GEMV: 4*L*D bytes per step, HBM-bound; one warp per window row, 128-bit loads.
//   K7  the selection block of cvt/validate.py:524-527,554,558,568 over the target list
//       ids = [pos] ++ ascending(rest) (validate.py:369-378), fused into one CTA: sums, alpha-mix,
//       max, threshold, renormalise, ordered compaction of the survivors.
// The reference's "probabilities" are logits divided by their plain sum, and the draw is uniform
// over survivors; both are reproduced, not corrected.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT)
l2_normalize_rows_kernel(const float *__restrict__ x, int64_t ld, int64_t dim, float *__restrict__ y,
                         int64_t ldy) {
    __shared__ double dred[32];
    const float *src = x + int64_t(blockIdx.x) * ld;
    float *dst = y + int64_t(blockIdx.x) * ldy;
    double ss = 0.0;
    for (int64_t k = threadIdx.x; k < dim; k += NT) {
        const double v = (double)src[k];
        ss += v * v;
    }
    ss = block_reduce(ss, 0.0, OpAdd<double>(), dred);
    const float denom = fmaxf((float)sqrt(ss), 1e-12f);          // F.normalize: x / max(||x||, eps)
    for (int64_t k = threadIdx.x; k < dim; k += NT) dst[k] = __fdiv_rn(src[k], denom);
}

// One warp per window; 8 windows per CTA.  q^ is re-read through L1 (it is shared by every warp).
__global__ void __launch_bounds__(NT)
cosine_scores_kernel(const float *__restrict__ tn, int64_t ld, int64_t rows, int64_t dim,
                     const float *__restrict__ qn, float temp, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = int64_t(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
    if (w >= rows) return;
    const float *row = tn + w * ld;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const bool vec = (((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(qn)) & 15) == 0);
    const int64_t dv = vec ? (dim & ~int64_t(3)) : 0;
    int64_t k = int64_t(lane) * 4;
    // eight independent 128-bit streaming loads in flight per lane (a 9 KB row is only 18 per lane: with
    // four in flight the kernel reached 0.65 of the HBM peak at D = 2304)
    for (; k + 7 * 128 < dv; k += 8 * 128) {
        float4 t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = ld_stream_f4(row + k + u * 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(qn + k + u * 128));
            float &acc = (u & 3) == 0 ? acc0 : (u & 3) == 1 ? acc1 : (u & 3) == 2 ? acc2 : acc3;
            acc = fmaf(t[u].x, q.x, acc); acc = fmaf(t[u].y, q.y, acc);
            acc = fmaf(t[u].z, q.z, acc); acc = fmaf(t[u].w, q.w, acc);
        }
    }
    for (; k + 1 * 128 < dv; k += 2 * 128) {
        const float4 t0 = ld_stream_f4(row + k), t1 = ld_stream_f4(row + k + 128);
        const float4 q0 = __ldg(reinterpret_cast<const float4 *>(qn + k)),
                     q1 = __ldg(reinterpret_cast<const float4 *>(qn + k + 128));
        acc0 = fmaf(t0.x, q0.x, acc0); acc0 = fmaf(t0.y, q0.y, acc0); acc0 = fmaf(t0.z, q0.z, acc0); acc0 = fmaf(t0.w, q0.w, acc0);
        acc1 = fmaf(t1.x, q1.x, acc1); acc1 = fmaf(t1.y, q1.y, acc1); acc1 = fmaf(t1.z, q1.z, acc1); acc1 = fmaf(t1.w, q1.w, acc1);
    }
    for (; k < dv; k += 128) {
        const float4 t0 = ld_stream_f4(row + k);
        const float4 q0 = __ldg(reinterpret_cast<const float4 *>(qn + k));
        acc2 = fmaf(t0.x, q0.x, acc2); acc2 = fmaf(t0.y, q0.y, acc2); acc2 = fmaf(t0.z, q0.z, acc2); acc2 = fmaf(t0.w, q0.w, acc2);
    }
    for (int64_t s = dv + lane; s < dim; s += 32) acc3 = fmaf(row[s], qn[s], acc3);
    const float dot = warp_sum((acc0 + acc1) + (acc2 + acc3));
    if (lane == 0) out[w] = __fdiv_rn(dot, temp);
}

constexpr int SELT = 1024;

__global__ void __launch_bounds__(SELT)
select_step_kernel(const float *__restrict__ o, const float *__restrict__ a, int64_t L, int64_t q,
                   float alpha, float oma, float th, int *__restrict__ choices, int *__restrict__ n_choices,
                   float *__restrict__ vals) {
    __shared__ double dred[32];
    __shared__ float fred[32];
    __shared__ int wcount[SELT / 32];
    __shared__ int base_s;
    const int64_t pos = (q + 1 < L - 1) ? q + 1 : L - 1;
    const bool q_in_list = (q == L - 1);                 // then pos == q and the query is target 0
    // pass A: plain sums of the logits over the target list (validate.py:524,526)
    double so = 0.0, sa = 0.0;
    for (int64_t w = threadIdx.x; w < L; w += SELT) {
        if (w == q && !q_in_list) continue;
        so += (double)o[w];
        if (a != nullptr) sa += (double)a[w];
    }
    const float So = (float)block_reduce(so, 0.0, OpAdd<double>(), dred);
    const float Sa = (float)block_reduce(sa, 0.0, OpAdd<double>(), dred);
    auto mixed = [&](int64_t w) -> float {
        const float on = __fdiv_rn(o[w], So);
        if (a == nullptr) return on;
        return __fadd_rn(__fmul_rn(alpha, on), __fmul_rn(oma, __fdiv_rn(a[w], Sa)));   // validate.py:527
    };
    // pass B: max
    float mx = -INFINITY;
    for (int64_t w = threadIdx.x; w < L; w += SELT) {
        if (w == q && !q_in_list) continue;
        mx = fmaxf(mx, mixed(w));
    }
    mx = block_reduce(mx, -INFINITY, OpMax(), fred);
    const float cut = __fsub_rn(mx, __fmul_rn(th, mx));                                 // validate.py:554
    // pass C: sum of the thresholded vector
    double sk = 0.0;
    for (int64_t w = threadIdx.x; w < L; w += SELT) {
        if (w == q && !q_in_list) continue;
        const float v = mixed(w);
        if (!(v < cut)) sk += (double)v;
    }
    const float Sk = (float)block_reduce(sk, 0.0, OpAdd<double>(), dred);
    // pass D: renormalise survivors (validate.py:558) and compact them in target-list order
    if (threadIdx.x == 0) {
        int b = 0;
        const float v = mixed(pos);
        const float r = (v < cut) ? 0.f : v;
        const float f = (r != 0.f) ? __fdiv_rn(r, Sk) : 0.f;
        if (f != 0.f) { choices[0] = (int)pos; b = 1; }
        base_s = b;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t w0 = 0; w0 < L; w0 += SELT) {
        const int64_t w = w0 + threadIdx.x;
        float f = 0.f;
        const bool in_list = (w < L) && !(w == q && !q_in_list);
        if (in_list) {
            const float v = mixed(w);
            const float r = (v < cut) ? 0.f : v;
            f = (r != 0.f) ? __fdiv_rn(r, Sk) : 0.f;
            if (vals != nullptr) vals[w] = f;
        }
        const bool take = in_list && (w != pos) && (f != 0.f);
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (lane == 0) wcount[wid] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int i = 0; i < wid; ++i) off += wcount[i];
        if (take) choices[off + __popc(bal & ((1u << lane) - 1u))] = (int)w;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < SELT / 32; ++i) t += wcount[i];
            base_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_choices = base_s;
}

// sims[w] = <x_w, d> / (||x_w|| * ||d||)   (fp64 accumulation), one warp per row.
__global__ void __launch_bounds__(NT)
audio_sims_kernel(const float *__restrict__ x, int64_t ld, int64_t rows, int64_t dim,
                  const float *__restrict__ d, float *__restrict__ sims) {
    const int lane = threadIdx.x & 31;
    const int64_t w = int64_t(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
    if (w >= rows) return;
    const float *row = x + w * ld;
    double dot = 0.0, sx = 0.0, sd = 0.0;
    for (int64_t k = lane; k < dim; k += 32) {
        const double xv = row[k], dv = d[k];
        dot += xv * dv; sx += xv * xv; sd += dv * dv;
    }
    dot = warp_sum(dot); sx = warp_sum(sx); sd = warp_sum(sd);
    if (lane == 0) {
        const double den = fmax(sqrt(sx), 1e-12) * fmax(sqrt(sd), 1e-12);
        sims[w] = (float)(dot / den);
    }
}

// First index whose similarity is strictly greater than every earlier one, starting from 0
// (validate.py:236-240: `if sim > max_sim`, max_sim initialised to 0, q_id to 0).
__global__ void __launch_bounds__(SELT)
first_argmax_kernel(const float *__restrict__ sims, int64_t rows, int *__restrict__ out) {
    __shared__ float fred[32];
    __shared__ int ired[32];
    float mx = 0.f;
    for (int64_t w = threadIdx.x; w < rows; w += SELT) mx = fmaxf(mx, sims[w]);
    mx = block_reduce(mx, 0.f, OpMax(), fred);
    int best = 0x7fffffff;
    if (mx > 0.f)
        for (int64_t w = threadIdx.x; w < rows; w += SELT)
            if (sims[w] == mx && (int)w < best) best = (int)w;
    struct OpMinI { __device__ int operator()(int p, int r) const { return p < r ? p : r; } };
    best = block_reduce(best, 0x7fffffff, OpMinI(), ired);
    if (threadIdx.x == 0) *out = (best == 0x7fffffff) ? 0 : best;
}

}  // namespace

extern "C" int avtex_l2_normalize_rows(const float *x, int64_t ld, int64_t rows, int64_t dim, float *y,
                                       int64_t ldy, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && dim >= 1 && ld >= dim && ldy >= dim, "l2_normalize_rows: bad shape");
    l2_normalize_rows_kernel<<<(unsigned)rows, NT, 0, as_stream(stream)>>>(x, ld, dim, y, ldy);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_cosine_scores(const float *tn, int64_t ld, int64_t rows, int64_t dim,
                                   const float *qn, float temp, float *out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && dim >= 1 && ld >= dim, "cosine_scores: bad shape");
    const unsigned grid = (unsigned)((rows + NT / 32 - 1) / (NT / 32));
    cosine_scores_kernel<<<grid, NT, 0, as_stream(stream)>>>(tn, ld, rows, dim, qn, temp, out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_select_step(const float *o, const float *a, int64_t L, int64_t q, float alpha,
                                 float one_minus_alpha, float th, int *choices, int *n_choices,
                                 float *vals, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(L >= 2 && q >= 0 && q < L && L < (int64_t(1) << 31), "select_step: bad L=%lld q=%lld",
                  (long long)L, (long long)q);
    select_step_kernel<<<1, SELT, 0, as_stream(stream)>>>(o, a, L, q, alpha, one_minus_alpha, th, choices,
                                                          n_choices, vals);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_audio_start(const float *x, int64_t ld, int64_t rows, int64_t dim, const float *d,
                                 float *sims_ws, int *out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(rows >= 1 && dim >= 1 && ld >= dim && rows < (int64_t(1) << 31), "audio_start: bad shape");
    const unsigned grid = (unsigned)((rows + NT / 32 - 1) / (NT / 32));
    audio_sims_kernel<<<grid, NT, 0, as_stream(stream)>>>(x, ld, rows, dim, d, sims_ws);
    AVTEX_LAUNCH_CHECK();
    first_argmax_kernel<<<1, SELT, 0, as_stream(stream)>>>(sims_ws, rows, out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

// K6 / K7 — contrastive synthesis step at the embedding boundary.
//   K6  L2-normalise rows (cvt/models/models.py:351,412) and one query against all L windows,
//       out[w] = <q^, t^_w> / temp (models.py:416-417; driving-audio term :433-439,457).
//       Streaming GEMV: 4*L*D bytes per step, HBM-bound; one warp per window row, 128-bit loads.
//   K7  the selection block of cvt/validate.py:524-527,554,558,568 over the target list
//       ids = [pos] ++ ascending(rest) (validate.py:369-378), fused into one CTA: sums, alpha-mix,
//       max, threshold, renormalise, ordered compaction of the survivors.
// The reference's "probabilities" are logits divided by their plain sum, and the draw is uniform
// over survivors; both are reproduced, not corrected.
#include <cooperative_groups.h>
#include <float.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;
constexpr int SELT = 1024;

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

// ------------------------------------------------------------------ one launch per synthesis step
// K6 + K7 fused: scores of the query against all L windows (+ the driving-audio scores) and their plain sums by
// all CTAs (rows handed out dynamically); the LAST CTA to finish (atomic ticket) does the alpha-mix, max,
// threshold, survivor sum and the ORDERED survivor list — one ordinary launch, no grid barrier, instead of three
// launches, a device-to-host copy and a stream synchronisation per step (round 1: 0.10 ms per step for a 39 us
// GEMV; a first cooperative version with three grid barriers still took 0.08 ms).  The count and the first
// `host_cap` survivors are written straight into MAPPED PINNED host memory, followed by the step's sequence
// number: the host polls that word.
// Same arithmetic as cosine_scores_kernel + select_step_kernel (explicit roundings, fp64 sums cast to fp32), so
// the survivor lists stay bit-identical to the reference restatement (cvt/validate.py:524-527,554,558,568).
struct SynthStepArgs {
    const float *tn; int64_t ld, L, dim;          // normalised window table
    const float *qn;                               // the query row
    const float *sn; int64_t lds, dimA;            // normalised source-audio table (nullable)
    const float *dn;                               // the driving-audio row
    float temp, alpha, oma, th;
    int64_t q;
    float *o, *a, *v;                              // [L] scratch: logits, audio logits, mixed values
    double *acc;                                   // [2][4] by step parity: sum o, sum a, sum kept, (unused)
    unsigned int *mx;                              // [2][2] by step parity: dynamic row counter, CTA ticket
    int *counts;                                   // (unused)
    int *choices, *n_choices;                      // device copy of the survivor list
    float *vals;                                   // nullable [L]
    volatile int *host;                            // mapped pinned: [seq, n, choices[0..host_cap)]
    int host_cap, seq, parity;
};

__device__ __forceinline__ float warp_row_dot(const float *__restrict__ row, const float *__restrict__ qn, int64_t dim,
                                              int lane) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const bool vec = (((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(qn)) & 15) == 0);
    const int64_t dv = vec ? (dim & ~int64_t(3)) : 0;
    int64_t k = int64_t(lane) * 4;
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
    return warp_sum((acc0 + acc1) + (acc2 + acc3));
}

// Selection over the L logits by ONE CTA of SELT threads (the last CTA of the step to finish its rows):
// mixed values + max, survivor sum, ordered compaction.  Each warp owns a contiguous slice of windows, so the
// ordered compaction needs one prefix over 32 warp counts instead of a barrier per 1024 elements.
__device__ void synthesis_select(const SynthStepArgs &p, double *acc, int64_t q, int seq) {
    __shared__ double dred[32];
    __shared__ float fred[32];
    __shared__ int wcnt[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t L = p.L;
    const int64_t pos = (q + 1 < L - 1) ? q + 1 : L - 1;
    const bool q_in_list = (q == L - 1);
    const bool audio = (p.sn != nullptr);
    const float So = (float)__ldcg(acc + 0), Sa = (float)__ldcg(acc + 1);
    // mixed values (validate.py:524-527) and their maximum
    float mx = -INFINITY;
    for (int64_t w = threadIdx.x; w < L; w += blockDim.x) {
        if (w == q && !q_in_list) continue;
        const float on = __fdiv_rn(__ldcg(p.o + w), So);
        const float val = audio ? __fadd_rn(__fmul_rn(p.alpha, on), __fmul_rn(p.oma, __fdiv_rn(__ldcg(p.a + w), Sa))) : on;
        p.v[w] = val;
        mx = fmaxf(mx, val);
    }
    mx = block_reduce(mx, -INFINITY, OpMax(), fred);                  // (also orders the p.v writes for this CTA)
    const float cut = __fsub_rn(mx, __fmul_rn(p.th, mx));             // validate.py:554
    // survivor sum + survivors per warp slice
    const int64_t seg = (L + nw - 1) / nw;
    const int64_t lo = int64_t(wid) * seg, hi = (lo + seg < L) ? lo + seg : L;
    double sk = 0.0;
    int cnt = 0;
    for (int64_t w = lo + lane; w < hi; w += 32) {
        if (w == q && !q_in_list) continue;
        const float val = p.v[w];
        if (!(val < cut)) {
            sk += (double)val;
            cnt += (val != 0.f) && (w != pos);
        }
    }
    cnt = warp_sum(cnt);
    if (lane == 0) wcnt[wid] = cnt;
    sk = block_reduce(sk, 0.0, OpAdd<double>(), dred);                // (barrier: wcnt visible)
    const float Sk = (float)sk;
    const float vpos = p.v[pos];
    const int pos_in = (!(vpos < cut) && vpos != 0.f) ? 1 : 0;
    int base = pos_in, total = pos_in;
    for (int i = 0; i < nw; ++i) {
        const int c = wcnt[i];
        if (i < wid) base += c;
        total += c;
    }
    if (threadIdx.x == 0) {
        if (pos_in) { p.choices[0] = (int)pos; if (p.host_cap > 0) p.host[2] = (int)pos; }
        *p.n_choices = total;
    }
    // renormalise (validate.py:558) and compact in target-list order: [pos] ++ ascending(rest)
    for (int64_t w0 = lo; w0 < hi; w0 += 32) {
        const int64_t w = w0 + lane;
        const bool in_list = (w < hi) && !(w == q && !q_in_list);
        bool keep = false;
        if (in_list) {
            const float val = p.v[w];
            keep = !(val < cut) && val != 0.f;
            if (p.vals != nullptr) p.vals[w] = keep ? __fdiv_rn(val, Sk) : 0.f;
        }
        const bool take = keep && (w != pos);
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int slot = base + __popc(bal & ((1u << lane) - 1u));
            p.choices[slot] = (int)w;
            if (slot < p.host_cap) p.host[2 + slot] = (int)w;
        }
        base += __popc(bal);
    }
    if (p.host == nullptr) {                                           // device-resident loop: the list stays here
        __threadfence();
        __syncthreads();
        return;
    }
    // publish: the list, then the count, then the sequence word the host polls
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        p.host[1] = total;
        __threadfence_system();
        p.host[0] = seq;
    }
}

// Logits of one query against all L windows (+ audio logits) and their plain sums over the target list, by all
// CTAs.  CTA c owns the contiguous rows [c*L/G, (c+1)*L/G); inside the CTA the warps take rows from a SHARED-memory
// ticket, one at a time: a static warp assignment leaves a 4-or-5 rows split (16 % idle), and a single global
// ticket serialises on one L2 atomic unit (20000 same-address atomics = the whole 60 us of the first version).
// Same operation order per row as cosine_scores_kernel.  Adds this CTA's sums to acc[0], acc[1].
__device__ __forceinline__ void synthesis_scores(const SynthStepArgs &p, const float *qn, const float *dn, int64_t q,
                                                 double *acc, int *ticket_s, double *dred) {
    const int lane = threadIdx.x & 31;
    const int64_t L = p.L;
    const bool q_in_list = (q == L - 1);
    const bool audio = (p.sn != nullptr);
    const int64_t lo = L * blockIdx.x / gridDim.x, hi = L * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) *ticket_s = 0;
    __syncthreads();
    double so = 0.0, sa = 0.0;
    int next = 0;
    if (lane == 0) next = atomicAdd(ticket_s, 1);
    for (;;) {
        const int64_t w = lo + __shfl_sync(0xffffffffu, next, 0);
        if (w >= hi) break;
        if (lane == 0) next = atomicAdd(ticket_s, 1);                  // the next ticket travels under this row's loads
        const float dot = warp_row_dot(p.tn + w * p.ld, qn, p.dim, lane);
        float ov = 0.f, av = 0.f;
        if (lane == 0) { ov = __fdiv_rn(dot, p.temp); p.o[w] = ov; }
        if (audio) {
            const float da = warp_row_dot(p.sn + w * p.lds, dn, p.dimA, lane);
            if (lane == 0) { av = __fdiv_rn(da, p.temp); p.a[w] = av; }
        }
        if (lane == 0 && !(w == q && !q_in_list)) { so += (double)ov; sa += (double)av; }
    }
    so = block_reduce(so, 0.0, OpAdd<double>(), dred);
    sa = block_reduce(sa, 0.0, OpAdd<double>(), dred);
    if (threadIdx.x == 0) { atomicAdd(acc + 0, so); atomicAdd(acc + 1, sa); }
}

__global__ void __launch_bounds__(SELT)
synthesis_step_kernel(const SynthStepArgs p) {
    __shared__ double dred[32];
    __shared__ int ticket_s;
    __shared__ int last_s;
    double *acc = p.acc + 4 * p.parity, *acc_next = p.acc + 4 * (1 - p.parity);
    unsigned int *ctr = p.mx + 2 * p.parity, *ctr_next = p.mx + 2 * (1 - p.parity);   // [1]: CTA ticket
    if (blockIdx.x == 0 && threadIdx.x < 4) {                          // the other parity's scratch is idle: re-arm it
        acc_next[threadIdx.x] = 0.0;
        if (threadIdx.x < 2) ctr_next[threadIdx.x] = 0u;
    }
    synthesis_scores(p, p.qn, p.dn, p.q, acc, &ticket_s, dred);
    if (threadIdx.x == 0) {
        __threadfence();                                               // my rows' logits + sums before my ticket
        last_s = (atomicAdd(ctr + 1, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last_s) return;
    __threadfence();                                                   // acquire side of the ticket
    synthesis_select(p, acc, p.q, p.seq);
}

// ------------------------------------------------------------------ the WHOLE synthesis loop in one launch
// numpy's legacy generator on the device.  The reference draws the next window with `np.random.choice(choices)`
// (cvt/validate.py:570) = RandomState.randint(0, n) = one masked-rejection draw of MT19937 32-bit outputs
// (numpy/random/src/distributions: random_bounded_uint64_fill with use_masked; n == 1 consumes nothing).  With
// the generator state imported from np.random.get_state() and exported back afterwards, the device draws the
// SAME numbers the host would have, so the chosen sequence stays bit-identical to the reference while the host
// leaves the loop entirely: no launch, no copy and no poll per step.
struct Mt19937 {
    uint32_t key[624];
    int pos;
};
__host__ __device__ inline void mt19937_gen(Mt19937 *st) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
    uint32_t *mt = st->key;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
    }
    for (; kk < 623; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
    }
    y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
    st->pos = 0;
}
__host__ __device__ inline uint32_t mt19937_next32(Mt19937 *st) {
    if (st->pos >= 624) mt19937_gen(st);
    uint32_t y = st->key[st->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
// RandomState.randint(0, n), n >= 1, n - 1 < 2^32 (numpy legacy: masked rejection on 32-bit draws)
__host__ __device__ inline uint32_t legacy_randint(Mt19937 *st, uint32_t n) {
    const uint32_t rng = n - 1u;
    if (rng == 0u) return 0u;
    if (rng == 0xFFFFFFFFu) return mt19937_next32(st);
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t val;
    do { val = mt19937_next32(st) & mask; } while (val > rng);
    return val;
}

struct SynthLoopArgs {
    SynthStepArgs s;                               // tables, scratch (qn / dn / q / seq / parity unused)
    const float *qn_table; int64_t ldq;            // normalised query table [L, dim]
    const float *dn_table; int64_t ldd;            // normalised driving table [>= n_steps + 1, dimA] (nullable)
    int64_t q_start;
    int n_steps;
    Mt19937 *mt;                                   // device copy of numpy's generator state (in / out)
    int *q_ids, *nz;                               // [n_steps] out
    int64_t *q_cur;                                // device scalar: the current query (scratch)
};

__global__ void __launch_bounds__(SELT)
synthesis_loop_kernel(const SynthLoopArgs a) {
    __shared__ double dred[32];
    __shared__ int ticket_s;
    cg::grid_group grid = cg::this_grid();
    SynthStepArgs p = a.s;
    p.host_cap = 0;                                                    // nothing goes to the host during the loop
    p.host = nullptr;
    int64_t q = a.q_start;
    for (int step = 0; step < a.n_steps; ++step) {
        const int parity = step & 1;
        double *acc = p.acc + 4 * parity, *acc_next = p.acc + 4 * (1 - parity);
        if (blockIdx.x == 0 && threadIdx.x < 4) acc_next[threadIdx.x] = 0.0;
        // driving-audio example `iter_count` = step + 1 (cvt/validate.py:417)
        const float *dn = a.dn_table != nullptr ? a.dn_table + int64_t(step + 1) * a.ldd : nullptr;
        synthesis_scores(p, a.qn_table + q * a.ldq, dn, q, acc, &ticket_s, dred);
        grid.sync();
        if (blockIdx.x == 0) {
            synthesis_select(p, acc, q, 0);
            if (threadIdx.x == 0) {
                const int n = *p.n_choices;
                const uint32_t r = legacy_randint(a.mt, (uint32_t)n);  // == np.random.choice(choices)
                const int64_t nq = p.choices[r];
                a.q_ids[step] = (int)nq;
                a.nz[step] = n;
                *a.q_cur = nq;
            }
        }
        grid.sync();
        q = *reinterpret_cast<volatile int64_t *>(a.q_cur);
    }
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

extern "C" int avtex_synthesis_step(const float *tn, int64_t ld, int64_t L, int64_t dim, const float *qn,
                                    const float *sn, int64_t lds, int64_t dimA, const float *dn, float temp,
                                    int64_t q, float alpha, float one_minus_alpha, float th, float *ws_f32,
                                    double *ws_acc, unsigned int *ws_max, int *ws_counts, int ws_counts_len,
                                    int *choices, int *n_choices, float *vals, int *host_out, int host_cap,
                                    int seq, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(L >= 2 && q >= 0 && q < L && L < (int64_t(1) << 31) && dim >= 1 && ld >= dim,
                  "synthesis_step: bad L=%lld q=%lld dim=%lld", (long long)L, (long long)q, (long long)dim);
    AVTEX_REQUIRE((sn == nullptr) == (dn == nullptr) && (sn == nullptr || (dimA >= 1 && lds >= dimA)),
                  "synthesis_step: audio table and driving row go together");
    AVTEX_REQUIRE(ws_f32 != nullptr && ws_acc != nullptr && ws_max != nullptr && ws_counts != nullptr &&
                      choices != nullptr && n_choices != nullptr && host_out != nullptr && host_cap >= 0,
                  "synthesis_step: workspace / output pointers must not be NULL");
    int sms = 0, cc = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    int64_t grid = (int64_t)sms;                     // one 1024-thread CTA per SM: 32 warps each stream the table
    const int64_t need = (L + SELT / 32 - 1) / (SELT / 32);
    if (grid > need) grid = need;
    (void)ws_counts_len;
    SynthStepArgs p;
    p.tn = tn; p.ld = ld; p.L = L; p.dim = dim; p.qn = qn; p.sn = sn; p.lds = lds; p.dimA = dimA; p.dn = dn;
    p.temp = temp; p.alpha = alpha; p.oma = one_minus_alpha; p.th = th; p.q = q;
    p.o = ws_f32; p.a = ws_f32 + L; p.v = ws_f32 + 2 * L;
    p.acc = ws_acc; p.mx = ws_max; p.counts = ws_counts; p.choices = choices; p.n_choices = n_choices; p.vals = vals;
    p.host = host_out; p.host_cap = host_cap; p.seq = seq; p.parity = seq & 1;
    synthesis_step_kernel<<<(unsigned)grid, SELT, 0, as_stream(stream)>>>(p);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_synthesis_loop(const float *tn, int64_t ld, int64_t L, int64_t dim, const float *qn_table,
                                    int64_t ldq, const float *sn, int64_t lds, int64_t dimA, const float *dn_table,
                                    int64_t ldd, float temp, float alpha, float one_minus_alpha, float th,
                                    int64_t q_start, int n_steps, float *ws_f32, double *ws_acc, int *choices,
                                    int *n_choices, void *mt_state, int64_t *q_scratch, int *q_ids, int *nz,
                                    int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(L >= 2 && q_start >= 0 && q_start < L && L < (int64_t(1) << 31) && dim >= 1 && ld >= dim && ldq >= dim,
                  "synthesis_loop: bad L=%lld q_start=%lld dim=%lld", (long long)L, (long long)q_start, (long long)dim);
    AVTEX_REQUIRE((sn == nullptr) == (dn_table == nullptr) && (sn == nullptr || (dimA >= 1 && lds >= dimA && ldd >= dimA)),
                  "synthesis_loop: audio table and driving table go together");
    AVTEX_REQUIRE(n_steps >= 1 && ws_f32 != nullptr && ws_acc != nullptr && choices != nullptr && n_choices != nullptr &&
                      mt_state != nullptr && q_scratch != nullptr && q_ids != nullptr && nz != nullptr,
                  "synthesis_loop: workspace / output pointers must not be NULL");
    int sms = 0, cc = 0, per_sm = 0, coop = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    AVTEX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    AVTEX_REQUIRE(coop != 0, "synthesis_loop: device does not support cooperative launch");
    AVTEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, synthesis_loop_kernel, SELT, 0));
    AVTEX_REQUIRE(per_sm >= 1, "synthesis_loop: kernel does not fit on an SM");
    int64_t grid = (int64_t)sms;                     // one 1024-thread CTA per SM
    const int64_t need = (L + SELT / 32 - 1) / (SELT / 32);
    if (grid > need) grid = need;
    SynthLoopArgs a;
    SynthStepArgs &p = a.s;
    p.tn = tn; p.ld = ld; p.L = L; p.dim = dim; p.qn = nullptr; p.sn = sn; p.lds = lds; p.dimA = dimA; p.dn = nullptr;
    p.temp = temp; p.alpha = alpha; p.oma = one_minus_alpha; p.th = th; p.q = q_start;
    p.o = ws_f32; p.a = ws_f32 + L; p.v = ws_f32 + 2 * L;
    p.acc = ws_acc; p.mx = nullptr; p.counts = nullptr; p.choices = choices; p.n_choices = n_choices; p.vals = nullptr;
    p.host = nullptr; p.host_cap = 0; p.seq = 0; p.parity = 0;
    a.qn_table = qn_table; a.ldq = ldq; a.dn_table = dn_table; a.ldd = ldd; a.q_start = q_start; a.n_steps = n_steps;
    a.mt = static_cast<Mt19937 *>(mt_state); a.q_ids = q_ids; a.nz = nz; a.q_cur = q_scratch;
    void *args[] = {(void *)&a};
    AVTEX_CUDA(cudaLaunchCooperativeKernel((void *)synthesis_loop_kernel, dim3((unsigned)grid), dim3(SELT), args, 0,
                                           as_stream(stream)));
    return 0;
}

// Host test hook: `count` draws of RandomState.randint(0, n[i]) from the state (key[624], *pos), updated in place —
// the same code the device loop runs (tests/test_host_cpu.py checks it against numpy itself).
extern "C" int avtex_mt19937_randint_host(uint32_t *key, int *pos, const uint32_t *n, int count, uint32_t *out) {
    AVTEX_REQUIRE(key != nullptr && pos != nullptr && n != nullptr && out != nullptr && count >= 0 && *pos >= 0 && *pos <= 624,
                  "mt19937_randint_host: bad arguments");
    Mt19937 st;
    memcpy(st.key, key, sizeof(st.key));
    st.pos = *pos;
    for (int i = 0; i < count; ++i) {
        AVTEX_REQUIRE(n[i] >= 1, "mt19937_randint_host: n must be >= 1");
        out[i] = legacy_randint(&st, n[i]);
    }
    memcpy(key, st.key, sizeof(st.key));
    *pos = st.pos;
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

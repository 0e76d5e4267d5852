// K3 / K4 — future cost ("anticipating the future", classic/q_learning.py:36-51) in vector form.
//
// The reference materialises X^{t+1}[i,:] = D3[i,:] + alpha * m^t with m^t_j = min_{k!=j} X^t[j,k] and
// rebuilds an M x M mask M times per sweep.  Because every row receives the SAME vector, the state
// of the iteration is the M-vector m alone:
//     m^{t+1}_j = min_{k != j} ( D3[j,k] + fl(alpha * m^t_k) )  (j >= 1),   m_0 fixed,
// so one sweep is ONE streaming read of D3 (4 M^2 bytes, HBM-bound; L2-resident when 4 M^2 < ~100 MB).
// min and a single rounded add are order-free, so every sweep is bit-identical to the reference
// given the same D3.  The reference's stop test needs eps = mean((X^{t+1}-X^t)^2); its numerator is
// accumulated (fp64) in the NEXT sweep's read, so no extra pass over D3 is spent on it.
//
// fl(alpha*m) and the add are written with __fmul_rn / __fadd_rn: nvcc must not contract them into
// an FMA, the reference rounds twice (`alpha * mins`, then `D3[i] + ...`).
#include <cooperative_groups.h>
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int ST = 256;

struct SweepAcc {
    float mn;
    double e;
};

// One element: x = D3 + fl(alpha*m_prev); row minimum off the diagonal; eps numerator against
// x_prev = D3 + fl(alpha*m_prev2).  a / a2 are the already loaded m_prev[k] / m_prev2[k].
template <bool USE_A, bool EPS, bool HAVE2>
__device__ __forceinline__ void sweep_elem(float v, float a, float a2, int64_t k, int64_t j, float alpha,
                                           SweepAcc &acc) {
    float x = v;
    if (USE_A) x = __fadd_rn(v, __fmul_rn(alpha, a));
    if (k != j) acc.mn = fminf(acc.mn, x);
    if (EPS) {
        const float xp = HAVE2 ? __fadd_rn(v, __fmul_rn(alpha, a2)) : v;
        const float d = __fsub_rn(x, xp);
        acc.e += (double)__fmul_rn(d, d);
    }
}

// Four elements.  The sweep with the eps numerator is not far from instruction-issue bound at HBM speed
// (5.8 elements per clock per SM), so: the diagonal test is made once per float4 (only the one vector of a row
// that holds D3[j,j] takes the per-element path), and the four squared differences are added in fp32 before the
// single conversion + fp64 add (4 fp32 squares: <= 2 ulp, far inside the 1e-6 the eps trail is compared at).
template <bool USE_A, bool EPS, bool HAVE2>
__device__ __forceinline__ void sweep_vec4(const float4 v, const float *__restrict__ mp,
                                           const float *__restrict__ mp2, int64_t k, int64_t j, float alpha,
                                           SweepAcc &acc) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a;
    // the m vectors are rewritten between sweeps by other CTAs (and peer GPUs) of the fused kernels: coherent
    // loads, cached in L1 — the grid barrier between sweeps invalidates L1 (see ld_ca_f4)
    if (USE_A) a = ld_ca_f4(mp + k);
    if (EPS && HAVE2) a2 = ld_ca_f4(mp2 + k);
    float4 x = v;
    if (USE_A) {
        x.x = __fadd_rn(v.x, __fmul_rn(alpha, a.x));
        x.y = __fadd_rn(v.y, __fmul_rn(alpha, a.y));
        x.z = __fadd_rn(v.z, __fmul_rn(alpha, a.z));
        x.w = __fadd_rn(v.w, __fmul_rn(alpha, a.w));
    }
    if ((unsigned long long)(j - k) < 4ull) {                 // this vector holds the diagonal element
        if (k + 0 != j) acc.mn = fminf(acc.mn, x.x);
        if (k + 1 != j) acc.mn = fminf(acc.mn, x.y);
        if (k + 2 != j) acc.mn = fminf(acc.mn, x.z);
        if (k + 3 != j) acc.mn = fminf(acc.mn, x.w);
    } else {
        acc.mn = fminf(fminf(acc.mn, x.x), fminf(fminf(x.y, x.z), x.w));
    }
    if (EPS) {
        float4 xp = v;
        if (HAVE2) {
            xp.x = __fadd_rn(v.x, __fmul_rn(alpha, a2.x));
            xp.y = __fadd_rn(v.y, __fmul_rn(alpha, a2.y));
            xp.z = __fadd_rn(v.z, __fmul_rn(alpha, a2.z));
            xp.w = __fadd_rn(v.w, __fmul_rn(alpha, a2.w));
        }
        const float d0 = __fsub_rn(x.x, xp.x), d1 = __fsub_rn(x.y, xp.y), d2 = __fsub_rn(x.z, xp.z),
                    d3 = __fsub_rn(x.w, xp.w);
        acc.e += (double)fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
    }
}

template <bool USE_A, bool EPS, bool HAVE2>
__device__ __forceinline__ void sweep_row(const float *__restrict__ row, int64_t m, int64_t j,
                                          const float *__restrict__ mp, const float *__restrict__ mp2,
                                          float alpha, SweepAcc &acc) {
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int64_t mv = vec ? (m & ~int64_t(3)) : 0;
    int64_t k = int64_t(threadIdx.x) * 4;
    // four independent 128-bit streaming loads in flight per thread (HBM latency hiding)
    for (; k + 3 * ST * 4 < mv; k += 4 * ST * 4) {
        const float4 v0 = ld_stream_f4(row + k), v1 = ld_stream_f4(row + k + ST * 4),
                     v2 = ld_stream_f4(row + k + 2 * ST * 4), v3 = ld_stream_f4(row + k + 3 * ST * 4);
        sweep_vec4<USE_A, EPS, HAVE2>(v0, mp, mp2, k, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v1, mp, mp2, k + ST * 4, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v2, mp, mp2, k + 2 * ST * 4, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v3, mp, mp2, k + 3 * ST * 4, j, alpha, acc);
    }
    for (; k < mv; k += ST * 4) sweep_vec4<USE_A, EPS, HAVE2>(ld_stream_f4(row + k), mp, mp2, k, j, alpha, acc);
    for (int64_t s = mv + threadIdx.x; s < m; s += ST)
        sweep_elem<USE_A, EPS, HAVE2>(row[s], USE_A ? ld_ca_f(mp + s) : 0.f, (EPS && HAVE2) ? ld_ca_f(mp2 + s) : 0.f, s, j, alpha, acc);
}

// One WARP per row (rows of a few KB, matrix L2-resident): lane-strided float4 loads, no block-level reduction.
template <bool USE_A, bool EPS, bool HAVE2>
__device__ __forceinline__ void sweep_row_warp(const float *__restrict__ row, int64_t m, int64_t j,
                                               const float *__restrict__ mp, const float *__restrict__ mp2,
                                               float alpha, SweepAcc &acc, int lane) {
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int64_t mv = vec ? (m & ~int64_t(3)) : 0;
    int64_t k = int64_t(lane) * 4;
    for (; k + 3 * 128 < mv; k += 4 * 128) {
        const float4 v0 = ld_stream_f4(row + k), v1 = ld_stream_f4(row + k + 128), v2 = ld_stream_f4(row + k + 256),
                     v3 = ld_stream_f4(row + k + 384);
        sweep_vec4<USE_A, EPS, HAVE2>(v0, mp, mp2, k, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v1, mp, mp2, k + 128, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v2, mp, mp2, k + 256, j, alpha, acc);
        sweep_vec4<USE_A, EPS, HAVE2>(v3, mp, mp2, k + 384, j, alpha, acc);
    }
    for (; k < mv; k += 128) sweep_vec4<USE_A, EPS, HAVE2>(ld_stream_f4(row + k), mp, mp2, k, j, alpha, acc);
    for (int64_t s = mv + lane; s < m; s += 32)
        sweep_elem<USE_A, EPS, HAVE2>(row[s], USE_A ? ld_ca_f(mp + s) : 0.f, (EPS && HAVE2) ? ld_ca_f(mp2 + s) : 0.f, s, j, alpha, acc);
}

// One pass of the small-matrix fused kernels over this CTA's rows, a warp per row.  At M = 1241 (C2: D3 = 6 MB, in L2)
// a CTA-wide row left 3/4 of the threads without data and cost two block reductions per row; with a warp per row the
// grid shrinks to M/8 CTAs, which also makes the grid barrier between the sweeps cheaper — that barrier, not the
// data, is the cost of a 6 MB sweep.  `emit(j, min)` is called by lane 0 of the row's warp; returns the CTA's eps
// numerator to every thread.
template <bool USE_A, bool EPS, bool HAVE2, typename Emit>
__device__ __forceinline__ double warp_rows_pass(const float *__restrict__ D3, int64_t ld, int64_t row0, int64_t rows,
                                                 int64_t m, const float *mp, const float *mp2, float alpha, double *dred,
                                                 Emit emit) {
    const int lane = threadIdx.x & 31;
    SweepAcc acc{INFINITY, 0.0};
    for (int64_t jl = int64_t(blockIdx.x) * (ST / 32) + (threadIdx.x >> 5); jl < rows; jl += int64_t(gridDim.x) * (ST / 32)) {
        const int64_t j = row0 + jl;
        const float *row = D3 + jl * ld;
        acc.mn = INFINITY;
        if (!USE_A || j == 0) sweep_row_warp<false, false, false>(row, m, j, mp, mp2, alpha, acc, lane);
        else sweep_row_warp<true, EPS, HAVE2>(row, m, j, mp, mp2, alpha, acc, lane);
        float mn = acc.mn;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) emit(j, mn);
    }
    return EPS ? block_reduce(acc.e, 0.0, OpAdd<double>(), dred) : 0.0;
}

// 6 CTAs per SM (<= 42 registers): with the compiler's own 51 this kernel fell from 0.31 to 0.40 ms per pass at M = 19961
__global__ void __launch_bounds__(ST, 6)
future_cost_sweep_kernel(const float *__restrict__ D3, int64_t ld, int64_t row0, int64_t m,
                         const float *__restrict__ mp, const float *__restrict__ mp2, float alpha,
                         float *__restrict__ m_new, double *eps_sum) {
    __shared__ float fred[32];
    __shared__ double dred[32];
    const int64_t j = row0 + blockIdx.x;
    const float *row = D3 + int64_t(blockIdx.x) * ld;
    const bool use_a = (mp != nullptr) && (j >= 1);          // row 0 is never updated (q_learning.py:42)
    const bool do_eps = (eps_sum != nullptr) && use_a;
    SweepAcc acc{INFINITY, 0.0};
    if (!use_a) sweep_row<false, false, false>(row, m, j, mp, mp2, alpha, acc);
    else if (!do_eps) sweep_row<true, false, false>(row, m, j, mp, mp2, alpha, acc);
    else if (mp2 == nullptr) sweep_row<true, true, false>(row, m, j, mp, mp2, alpha, acc);
    else sweep_row<true, true, true>(row, m, j, mp, mp2, alpha, acc);
    const float mn = block_reduce(acc.mn, INFINITY, OpMin(), fred);
    if (threadIdx.x == 0) m_new[j] = mn;
    if (do_eps) {
        const double e = block_reduce(acc.e, 0.0, OpAdd<double>(), dred);
        if (threadIdx.x == 0) atomicAdd(eps_sum, e);
    }
}

__global__ void __launch_bounds__(ST)
future_cost_finalize_kernel(const float *__restrict__ D3, int64_t ld, int64_t row0, int64_t m,
                            const float *__restrict__ mvec, float alpha, float *__restrict__ out,
                            int64_t ld_out, double *sum, unsigned long long *nnz) {
    __shared__ double sred[32];
    __shared__ unsigned long long nred[32];
    const int64_t j = row0 + blockIdx.x;
    const float *row = D3 + int64_t(blockIdx.x) * ld;
    float *dst = out + int64_t(blockIdx.x) * ld_out;
    const bool use_a = (j >= 1);
    double s = 0.0;
    unsigned long long z = 0;
    const bool vec = (((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
    const int64_t mv = vec ? (m & ~int64_t(3)) : 0;
    auto one = [&](float4 v, int64_t k) {
        if (use_a) {
            const float4 a = *reinterpret_cast<const float4 *>(mvec + k);
            v.x = __fadd_rn(v.x, __fmul_rn(alpha, a.x));
            v.y = __fadd_rn(v.y, __fmul_rn(alpha, a.y));
            v.z = __fadd_rn(v.z, __fmul_rn(alpha, a.z));
            v.w = __fadd_rn(v.w, __fmul_rn(alpha, a.w));
        }
        *reinterpret_cast<float4 *>(dst + k) = v;
        s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
        z += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
    };
    int64_t k = int64_t(threadIdx.x) * 4;
    for (; k + 3 * ST * 4 < mv; k += 4 * ST * 4) {
        const float4 v0 = ld_stream_f4(row + k), v1 = ld_stream_f4(row + k + ST * 4),
                     v2 = ld_stream_f4(row + k + 2 * ST * 4), v3 = ld_stream_f4(row + k + 3 * ST * 4);
        one(v0, k); one(v1, k + ST * 4); one(v2, k + 2 * ST * 4); one(v3, k + 3 * ST * 4);
    }
    for (; k < mv; k += ST * 4) one(ld_stream_f4(row + k), k);
    for (int64_t k = mv + threadIdx.x; k < m; k += ST) {
        float v = row[k];
        if (use_a) v = __fadd_rn(v, __fmul_rn(alpha, mvec[k]));
        dst[k] = v;
        s += (double)v;
        z += (v != 0.f);
    }
    if (sum != nullptr) {
        s = block_reduce(s, 0.0, OpAdd<double>(), sred);
        z = block_reduce(z, 0ull, OpAdd<unsigned long long>(), nred);
        if (threadIdx.x == 0) { atomicAdd(sum, s); atomicAdd(nnz, z); }
    }
}

// All sweeps in one cooperative launch: persistent CTAs stride over the rows, a grid barrier separates
// the sweeps, and every CTA evaluates the reference's stop rule from the same fp64 numerator.
// At M = 1241 (D3 = 6 MB, L2-resident) a sweep is ~5 us of work; the per-sweep launch + host read of
// eps it replaces cost ~25 us each.
__global__ void __launch_bounds__(ST)
future_cost_fused_kernel(const float *__restrict__ D3, int64_t ld, int64_t m, float alpha, float eps_stop,
                         int max_sweeps, float *mbuf, int64_t mpad, double *eps_trail, int *info, float *m_out) {
    __shared__ float fred[32];
    __shared__ double dred[32];
    cg::grid_group grid = cg::this_grid();
    float *buf[3] = {mbuf, mbuf + mpad, mbuf + 2 * mpad};
    // pass 0: m^0 = off-diagonal row minima of D3
    for (int64_t j = blockIdx.x; j < m; j += gridDim.x) {
        SweepAcc acc{INFINITY, 0.0};
        sweep_row<false, false, false>(D3 + j * ld, m, j, nullptr, nullptr, alpha, acc);
        const float mn = block_reduce(acc.mn, INFINITY, OpMin(), fred);
        if (threadIdx.x == 0) buf[0][j] = mn;
    }
    grid.sync();
    int cur = 0, prev2 = -1;
    for (int p = 1; p <= max_sweeps; ++p) {
        const int out = 3 - cur - (prev2 < 0 ? (cur == 0 ? 1 : 0) : prev2);      // the buffer that is neither cur nor prev2
        const float *mp = buf[cur];
        const float *mp2 = prev2 < 0 ? nullptr : buf[prev2];
        double e_blk = 0.0;
        for (int64_t j = blockIdx.x; j < m; j += gridDim.x) {
            SweepAcc acc{INFINITY, 0.0};
            const float *row = D3 + j * ld;
            if (j == 0) sweep_row<false, false, false>(row, m, j, mp, mp2, alpha, acc);
            else if (mp2 == nullptr) sweep_row<true, true, false>(row, m, j, mp, mp2, alpha, acc);
            else sweep_row<true, true, true>(row, m, j, mp, mp2, alpha, acc);
            const float mn = block_reduce(acc.mn, INFINITY, OpMin(), fred);
            if (threadIdx.x == 0) buf[out][j] = mn;
            e_blk += block_reduce(acc.e, 0.0, OpAdd<double>(), dred);
        }
        if (threadIdx.x == 0 && e_blk != 0.0) atomicAdd(eps_trail + p, e_blk);
        grid.sync();
        const double num = *reinterpret_cast<volatile double *>(eps_trail + p);
        const float eps = (float)(num / ((double)m * (double)m));
        if (!(eps > eps_stop)) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = p; info[1] = cur; }
            // the converged vector goes to a fixed place, so the host can launch the finalize kernel
            // without first reading which of the three rotating buffers holds it
            if (m_out != nullptr)
                for (int64_t k = int64_t(blockIdx.x) * ST + threadIdx.x; k < m; k += int64_t(gridDim.x) * ST)
                    m_out[k] = __ldcg(buf[cur] + k);
            return;
        }
        prev2 = cur;
        cur = out;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = 0; info[1] = cur; }
}

// future_cost_fused_kernel for small (L2-resident) matrices: a warp per row (warp_rows_pass).  Same contract.
__global__ void __launch_bounds__(ST)
future_cost_fused_small_kernel(const float *__restrict__ D3, int64_t ld, int64_t m, float alpha, float eps_stop,
                               int max_sweeps, float *mbuf, int64_t mpad, double *eps_trail, int *info, float *m_out) {
    __shared__ double dred[32];
    cg::grid_group grid = cg::this_grid();
    float *buf[3] = {mbuf, mbuf + mpad, mbuf + 2 * mpad};
    {
        float *dst = buf[0];
        warp_rows_pass<false, false, false>(D3, ld, 0, m, m, nullptr, nullptr, alpha, dred,
                                            [dst](int64_t j, float v) { dst[j] = v; });
    }
    grid.sync();
    int cur = 0, prev2 = -1;
    for (int p = 1; p <= max_sweeps; ++p) {
        const int out = 3 - cur - (prev2 < 0 ? (cur == 0 ? 1 : 0) : prev2);
        const float *mp = buf[cur];
        const float *mp2 = prev2 < 0 ? nullptr : buf[prev2];
        float *dst = buf[out];
        const double e_blk =
            mp2 == nullptr ? warp_rows_pass<true, true, false>(D3, ld, 0, m, m, mp, mp2, alpha, dred,
                                                               [dst](int64_t j, float v) { dst[j] = v; })
                           : warp_rows_pass<true, true, true>(D3, ld, 0, m, m, mp, mp2, alpha, dred,
                                                              [dst](int64_t j, float v) { dst[j] = v; });
        if (threadIdx.x == 0 && e_blk != 0.0) atomicAdd(eps_trail + p, e_blk);
        grid.sync();
        const double num = *reinterpret_cast<volatile double *>(eps_trail + p);
        const float eps = (float)(num / ((double)m * (double)m));
        if (!(eps > eps_stop)) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = p; info[1] = cur; }
            if (m_out != nullptr)
                for (int64_t k = int64_t(blockIdx.x) * ST + threadIdx.x; k < m; k += int64_t(gridDim.x) * ST)
                    m_out[k] = __ldcg(buf[cur] + k);
            return;
        }
        prev2 = cur;
        cur = out;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = 0; info[1] = cur; }
}

// ------------------------------------------------------------------ multi-GPU fused future cost
// Row-sharded version of future_cost_fused_kernel: every GPU runs one cooperative kernel over its own
// rows; after each sweep the freshly computed row minima are PUSHED into every peer's copy of the
// m vector (peer-mapped symmetric memory, plain stores over NVLink), the per-rank eps numerators are
// pushed into per-rank slots, and a flag barrier (monotonic counters written by the peers, spun on
// locally) closes the sweep.  This is the all-gather of the per-row minima the algorithm needs
// (BASELINE.json north_star) done inside the kernel: ~10 us per sweep instead of a kernel launch, two
// NCCL collectives and a host read of eps (~70 us).  Every rank adds the eps slots in rank order, so
// all ranks take the same stop decision from bit-identical numbers.
struct FcPeerArgs {
    const float *D3;
    int64_t ld, row0, rows, m, mpad;
    float alpha, eps_stop;
    int max_sweeps, rank, world;
    unsigned int epoch_base;
    float *mbuf[8];              // rank r's 3 * mpad floats
    double *epsbuf[8];           // rank r's (max_sweeps + 1) * world slots
    unsigned int *flags[8];      // rank r's `world` counters
    double *eps_local;           // [max_sweeps + 1], zeroed: this rank's numerators (atomicAdd target)
    double *eps_trail;           // [max_sweeps + 1] out: summed numerators
    int *info;                   // [0] sweeps (0: not converged)  [1] buffer index  [2] error (1: peer timeout)
    float *m_out;                // nullable [mpad]: the converged vector (local)
    long long timeout_cycles;
};

// Closes one sweep across the GPUs: thread r of CTA 0 handles peer r (eps slot, flag, spin), so the
// NVLink round trips to the peers overlap instead of queueing behind each other.  A peer that does
// not show up within the bound makes every CTA of this rank leave the kernel with info[2] = 1 — no
// __trap(), which would poison the CUDA context; the peers then time out the same way.
__device__ __forceinline__ bool fc_peer_exchange(cg::grid_group &grid, const FcPeerArgs &a, int step, bool with_eps) {
    grid.sync();                                     // all local rows done, all pushes issued
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        const int r = threadIdx.x;
        const unsigned int target = a.epoch_base + (unsigned int)step + 1u;
        __threadfence_system();                      // cumulative: orders every CTA's pushes before the flag
        if (with_eps) {
            const double mine = *reinterpret_cast<volatile double *>(a.eps_local + step);
            *reinterpret_cast<volatile double *>(a.epsbuf[r] + (int64_t)step * a.world + a.rank) = mine;
            __threadfence_system();
        }
        *reinterpret_cast<volatile unsigned int *>(a.flags[r] + a.rank) = target;
        const long long t0 = clock64();
        while ((int)(*reinterpret_cast<volatile unsigned int *>(a.flags[a.rank] + r) - target) < 0) {
            if (clock64() - t0 > a.timeout_cycles) {
                atomicExch(a.info + 2, 1);
                break;
            }
        }
        __threadfence_system();
    }
    grid.sync();                                     // the gathered vector / eps slots may now be read
    return *reinterpret_cast<volatile int *>(a.info + 2) == 0;
}

__global__ void __launch_bounds__(ST)
future_cost_fused_peer_kernel(const FcPeerArgs a) {
    __shared__ float fred[32];
    __shared__ double dred[32];
    cg::grid_group grid = cg::this_grid();
    const float *local = a.mbuf[a.rank];
    const int64_t row_end = a.row0 + a.rows;
    auto push = [&](int buf, int64_t j, float v) {
        for (int r = 0; r < a.world; ++r) a.mbuf[r][(int64_t)buf * a.mpad + j] = v;
    };
    for (int64_t j = a.row0 + blockIdx.x; j < row_end; j += gridDim.x) {
        SweepAcc acc{INFINITY, 0.0};
        sweep_row<false, false, false>(a.D3 + (j - a.row0) * a.ld, a.m, j, nullptr, nullptr, a.alpha, acc);
        const float mn = block_reduce(acc.mn, INFINITY, OpMin(), fred);
        if (threadIdx.x == 0) push(0, j, mn);
    }
    if (!fc_peer_exchange(grid, a, 0, false)) return;
    int cur = 0, prev2 = -1;
    for (int p = 1; p <= a.max_sweeps; ++p) {
        const int out = 3 - cur - (prev2 < 0 ? (cur == 0 ? 1 : 0) : prev2);
        const float *mp = local + (int64_t)cur * a.mpad;
        const float *mp2 = prev2 < 0 ? nullptr : local + (int64_t)prev2 * a.mpad;
        double e_blk = 0.0;
        for (int64_t j = a.row0 + blockIdx.x; j < row_end; j += gridDim.x) {
            SweepAcc acc{INFINITY, 0.0};
            const float *row = a.D3 + (j - a.row0) * a.ld;
            if (j == 0) sweep_row<false, false, false>(row, a.m, j, mp, mp2, a.alpha, acc);
            else if (mp2 == nullptr) sweep_row<true, true, false>(row, a.m, j, mp, mp2, a.alpha, acc);
            else sweep_row<true, true, true>(row, a.m, j, mp, mp2, a.alpha, acc);
            const float mn = block_reduce(acc.mn, INFINITY, OpMin(), fred);
            if (threadIdx.x == 0) push(out, j, mn);
            e_blk += block_reduce(acc.e, 0.0, OpAdd<double>(), dred);
        }
        if (threadIdx.x == 0 && e_blk != 0.0) atomicAdd(a.eps_local + p, e_blk);
        if (!fc_peer_exchange(grid, a, p, true)) return;
        double num = 0.0;
        for (int r = 0; r < a.world; ++r)
            num += *reinterpret_cast<volatile double *>(a.epsbuf[a.rank] + (int64_t)p * a.world + r);
        if (blockIdx.x == 0 && threadIdx.x == 0) a.eps_trail[p] = num;
        const float eps = (float)(num / ((double)a.m * (double)a.m));
        if (!(eps > a.eps_stop)) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { a.info[0] = p; a.info[1] = cur; }
            if (a.m_out != nullptr)
                for (int64_t k = int64_t(blockIdx.x) * ST + threadIdx.x; k < a.m; k += int64_t(gridDim.x) * ST)
                    a.m_out[k] = __ldcg(mp + k);
            return;
        }
        prev2 = cur;
        cur = out;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.info[0] = 0; a.info[1] = cur; }
}

// future_cost_fused_peer_kernel for small shards: a warp per row.  Same contract and exchange protocol.
__global__ void __launch_bounds__(ST)
future_cost_fused_peer_small_kernel(const FcPeerArgs a) {
    __shared__ double dred[32];
    cg::grid_group grid = cg::this_grid();
    const float *local = a.mbuf[a.rank];
    auto push = [&a](int buf, int64_t j, float v) {
        for (int r = 0; r < a.world; ++r) a.mbuf[r][(int64_t)buf * a.mpad + j] = v;
    };
    warp_rows_pass<false, false, false>(a.D3, a.ld, a.row0, a.rows, a.m, nullptr, nullptr, a.alpha, dred,
                                        [&push](int64_t j, float v) { push(0, j, v); });
    if (!fc_peer_exchange(grid, a, 0, false)) return;
    int cur = 0, prev2 = -1;
    for (int p = 1; p <= a.max_sweeps; ++p) {
        const int out = 3 - cur - (prev2 < 0 ? (cur == 0 ? 1 : 0) : prev2);
        const float *mp = local + (int64_t)cur * a.mpad;
        const float *mp2 = prev2 < 0 ? nullptr : local + (int64_t)prev2 * a.mpad;
        const double e_blk =
            mp2 == nullptr ? warp_rows_pass<true, true, false>(a.D3, a.ld, a.row0, a.rows, a.m, mp, mp2, a.alpha, dred,
                                                               [&push, out](int64_t j, float v) { push(out, j, v); })
                           : warp_rows_pass<true, true, true>(a.D3, a.ld, a.row0, a.rows, a.m, mp, mp2, a.alpha, dred,
                                                              [&push, out](int64_t j, float v) { push(out, j, v); });
        if (threadIdx.x == 0 && e_blk != 0.0) atomicAdd(a.eps_local + p, e_blk);
        if (!fc_peer_exchange(grid, a, p, true)) return;
        double num = 0.0;
        for (int r = 0; r < a.world; ++r)
            num += *reinterpret_cast<volatile double *>(a.epsbuf[a.rank] + (int64_t)p * a.world + r);
        if (blockIdx.x == 0 && threadIdx.x == 0) a.eps_trail[p] = num;
        const float eps = (float)(num / ((double)a.m * (double)a.m));
        if (!(eps > a.eps_stop)) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { a.info[0] = p; a.info[1] = cur; }
            if (a.m_out != nullptr)
                for (int64_t k = int64_t(blockIdx.x) * ST + threadIdx.x; k < a.m; k += int64_t(gridDim.x) * ST)
                    a.m_out[k] = __ldcg(mp + k);
            return;
        }
        prev2 = cur;
        cur = out;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.info[0] = 0; a.info[1] = cur; }
}

// Rows of a few KB in an L2-sized matrix / shard: the warp-per-row kernels.
inline bool fc_small(int64_t rows, int64_t m) { return m <= 8192 && rows * m * 4 <= (int64_t(96) << 20); }

}  // namespace

extern "C" int avtex_future_cost_sweep(const float *D3, int64_t ld, int64_t row0, int64_t rows, int64_t m,
                                       const float *m_prev, const float *m_prev2, float alpha,
                                       float *m_new, double *eps_sum, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(m >= 2 && rows >= 1 && row0 >= 0 && row0 + rows <= m && ld >= m,
                  "future_cost_sweep: bad shape row0=%lld rows=%lld m=%lld ld=%lld", (long long)row0,
                  (long long)rows, (long long)m, (long long)ld);
    AVTEX_REQUIRE(m_prev != nullptr || m_prev2 == nullptr, "future_cost_sweep: m_prev2 without m_prev");
    AVTEX_REQUIRE(((reinterpret_cast<uintptr_t>(m_prev) | reinterpret_cast<uintptr_t>(m_prev2)) & 15) == 0,
                  "future_cost_sweep: m vectors must be 16-byte aligned");
    future_cost_sweep_kernel<<<(unsigned)rows, ST, 0, as_stream(stream)>>>(D3, ld, row0, m, m_prev, m_prev2,
                                                                           alpha, m_new, eps_sum);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_future_cost_finalize(const float *D3, int64_t ld, int64_t row0, int64_t rows,
                                          int64_t m, const float *mvec, float alpha, float *D3_new,
                                          int64_t ld_out, double *sum, unsigned long long *nnz,
                                          int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(m >= 1 && rows >= 1 && row0 >= 0 && row0 + rows <= m && ld >= m && ld_out >= m,
                  "future_cost_finalize: bad shape row0=%lld rows=%lld m=%lld", (long long)row0,
                  (long long)rows, (long long)m);
    AVTEX_REQUIRE(mvec != nullptr && (reinterpret_cast<uintptr_t>(mvec) & 15) == 0,
                  "future_cost_finalize: mvec must be non-NULL and 16-byte aligned");
    AVTEX_REQUIRE((sum == nullptr) == (nnz == nullptr), "future_cost_finalize: sum and nnz go together");
    future_cost_finalize_kernel<<<(unsigned)rows, ST, 0, as_stream(stream)>>>(D3, ld, row0, m, mvec, alpha,
                                                                              D3_new, ld_out, sum, nnz);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_future_cost_fused(const float *D3, int64_t ld, int64_t m, float alpha, float eps_stop,
                                       int max_sweeps, float *mbuf, int64_t mpad, double *eps_trail, int *info,
                                       float *m_out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(m >= 2 && ld >= m && mpad >= m && mpad % 4 == 0 && max_sweeps >= 1,
                  "future_cost_fused: bad shape m=%lld ld=%lld mpad=%lld", (long long)m, (long long)ld, (long long)mpad);
    AVTEX_REQUIRE((reinterpret_cast<uintptr_t>(mbuf) & 15) == 0, "future_cost_fused: mbuf must be 16-byte aligned");
    int sms = 0, cc = 0, per_sm = 0, coop = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    AVTEX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    AVTEX_REQUIRE(coop != 0, "future_cost_fused: device does not support cooperative launch");
    const bool small = fc_small(m, m);
    const void *kernel = small ? (const void *)future_cost_fused_small_kernel : (const void *)future_cost_fused_kernel;
    AVTEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ST, 0));
    AVTEX_REQUIRE(per_sm >= 1, "future_cost_fused: kernel does not fit on an SM");
    int64_t grid = (int64_t)sms * per_sm;              // full occupancy: the sweeps are HBM/L2 streaming
    const int64_t work = small ? (m + ST / 32 - 1) / (ST / 32) : m;          // CTAs that have a row (or 8) to process
    if (grid > work) grid = work;
    void *args[] = {(void *)&D3, (void *)&ld, (void *)&m, (void *)&alpha, (void *)&eps_stop, (void *)&max_sweeps,
                    (void *)&mbuf, (void *)&mpad, (void *)&eps_trail, (void *)&info, (void *)&m_out};
    AVTEX_CUDA(cudaLaunchCooperativeKernel(kernel, dim3((unsigned)grid), dim3(ST), args, 0, as_stream(stream)));
    return 0;
}

extern "C" int avtex_future_cost_fused_peer(const float *D3, int64_t ld, int64_t row0, int64_t rows, int64_t m,
                                            float alpha, float eps_stop, int max_sweeps, int rank, int world,
                                            float *const *h_mbuf, int64_t mpad, double *const *h_epsbuf,
                                            unsigned int *const *h_flags, unsigned int epoch_base,
                                            double *eps_local, double *eps_trail, int *info, float *m_out,
                                            int max_ctas, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "future_cost_fused_peer: bad rank %d / world %d", rank, world);
    AVTEX_REQUIRE(m >= 2 && rows >= 1 && row0 >= 0 && row0 + rows <= m && ld >= m && mpad >= m && mpad % 4 == 0 &&
                      max_sweeps >= 1,
                  "future_cost_fused_peer: bad shape row0=%lld rows=%lld m=%lld", (long long)row0, (long long)rows, (long long)m);
    FcPeerArgs a;
    a.D3 = D3; a.ld = ld; a.row0 = row0; a.rows = rows; a.m = m; a.mpad = mpad;
    a.alpha = alpha; a.eps_stop = eps_stop; a.max_sweeps = max_sweeps; a.rank = rank; a.world = world;
    a.epoch_base = epoch_base;
    for (int r = 0; r < 8; ++r) {
        a.mbuf[r] = r < world ? h_mbuf[r] : nullptr;
        a.epsbuf[r] = r < world ? h_epsbuf[r] : nullptr;
        a.flags[r] = r < world ? h_flags[r] : nullptr;
    }
    a.eps_local = eps_local; a.eps_trail = eps_trail; a.info = info; a.m_out = m_out;
    a.timeout_cycles = 20000000000LL;            // ~10 s: host-side skew is absorbed by the caller's barrier
    int sms = 0, cc = 0, per_sm = 0, coop = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    AVTEX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    AVTEX_REQUIRE(coop != 0, "future_cost_fused_peer: device does not support cooperative launch");
    const bool small = fc_small(rows, m);
    const void *kernel = small ? (const void *)future_cost_fused_peer_small_kernel : (const void *)future_cost_fused_peer_kernel;
    AVTEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ST, 0));
    AVTEX_REQUIRE(per_sm >= 1, "future_cost_fused_peer: kernel does not fit on an SM");
    int64_t grid = (int64_t)sms * per_sm;
    // max_ctas > 0: several ranks share ONE device ("virtual ranks", tests): their cooperative kernels must
    // all be resident at the same time or the flag barrier would never close
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (max_ctas < 0) grid = grid / (-max_ctas) > 0 ? grid / (-max_ctas) : 1;      // -G: an equal share for each of G ranks
    const int64_t work = small ? (rows + ST / 32 - 1) / (ST / 32) : rows;
    if (grid > work) grid = work;
    void *args[] = {(void *)&a};
    AVTEX_CUDA(cudaLaunchCooperativeKernel(kernel, dim3((unsigned)grid), dim3(ST), args, 0, as_stream(stream)));
    return 0;
}

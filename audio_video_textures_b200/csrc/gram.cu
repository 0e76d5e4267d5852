// K1 — pairwise frame L2 distances as a Gram contraction on the 5th-gen tensor cores.
//
//   D[r,c] = sqrt( n_r + n_c - 2 <x_r, x_c> )        replaces classic/computeD1.py:50-56, 58-96
//
// x are byte frames: either the raw uint8 rows (unsigned operands, K0 = frame_norms_u8) or the centred
// int8 copy written by pack.cu (signed operands); n are their exact sums of squares.  tcgen05.mma kind::i8
// accumulates <x_r, x_c> in TMEM as int32 (modulo 2^32 for the unsigned form), the epilogue evaluates
// n_r + n_c - 2g modulo 2^32, and the result is the EXACT integer d^2 whenever the true d^2 < 2^32: no
// cancellation error, duplicate frames give exactly 0 (the reference's sigma counts `nonzero(D1)`), and the
// only rounding left is one fp32 sqrt.  The host guarantees the domain through
// d <= ||x_r - 128|| + ||x_c - 128||, i.e. 4 * max centred norm < 2^32, and otherwise routes to direct.cu.
//
// Work is a JOB LIST (GramJob: a rectangle of the matrix with a direct and/or a transposed destination, the
// latter possibly in PEER memory); the tiles of all jobs are enumerated in L2-sized groups.  Two kernels:
//   gram_l2_s8_2cta_kernel (default)  cluster of 2 CTAs, tcgen05.mma.cta_group::2, 256 x 256 tile:
//       warp 0 (both CTAs)   TMA producer: its 128 A rows + its 128 of the 256 B rows per 128-byte K block into a
//                            6-stage ring; completion bytes are credited to the leader's mbarrier
//       warp 1 (leader)      one lane issues 4 MMAs (M256 N256 K32) per stage; multicast tcgen05.commit frees
//                            the stage in both CTAs and signals both epilogues; accumulators double-buffered
//                            in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//       warps 2-5 (both)     epilogue: tcgen05.ld 32 lanes x 32 columns, fused norm + sqrt, 32 x 32 transpose
//                            through shared memory so every global store writes one full 128 B line (direct
//                            rows and transposed rows), fp32/int partial sigma statistics per chunk
//   gram_l2_s8_kernel (AVTEX_GRAM_MODE=1cta)  one CTA per SM, 128 x 256 tile, 4-stage 48 KB ring; kept as the
//       measured baseline of the 2-CTA design (shared-memory bandwidth bound: 57.6 % vs 68 % tensor pipe).
// Every mbarrier wait is bounded (trap, never hang).  DESIGN.md §4.1 has the measurements.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace {

constexpr int BM = 128;                 // tile rows  (UMMA M)
constexpr int BN = 256;                 // tile cols  (UMMA N)
constexpr int BKB = 128;                // K bytes per stage = one 128B swizzle atom
constexpr int UMMA_KB = 32;             // K bytes per tcgen05.mma (kind::i8: K = 32)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BKB;       // 16 KB
constexpr int B_BYTES = BN * BKB;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GROUP_M = 16;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int EPI_PITCH = 33;                              // epilogue transpose tile pitch (floats)
constexpr int EPI_SMEM_BYTES = 4 * 32 * EPI_PITCH * 4;     // one 32 x 33 tile per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_SMEM_BYTES;

// kind::i8 instruction descriptor (cute/arch/mma_sm100_desc.hpp bit layout):
//   c_format[4,6)=2 (S32) | a_format[7,10)=1 (s8) | b_format[10,13)=1 (s8) | a/b major = K (0)
//   n_dim[17,23) = N>>3 | m_dim[24,29) = M>>4
//   a_format / b_format: 1 = signed 8-bit (centred frames from pack.cu), 0 = unsigned (raw frames)
constexpr uint32_t make_idesc(int m, bool is_signed) {
    return (2u << 4) | ((is_signed ? 1u : 0u) << 7) | ((is_signed ? 1u : 0u) << 10) | (uint32_t(BN >> 3) << 17) |
           (uint32_t(m >> 4) << 24);
}

// One rectangle of the distance matrix: A rows [row0, row0+rows) x output columns [col0, col0+cols).
//   D  (nullable): direct destination,      D [(r - d_row0)  * ldd + c]
//   DT (nullable): transposed destination,  DT[(c - dt_row0) * ldt + r]   -- may be PEER memory (NVLink)
//   symmetric: rows == cols range; only tiles touching the upper triangle are computed, the direct store
//   takes r <= c and the transposed store r < c (the mirror image).
struct GramJob {
    int64_t row0, rows, col0, cols;
    float *D;
    int64_t d_row0, ldd;
    float *DT;
    int64_t dt_row0, ldt;
    int symmetric, count_stats, TM, TN, tile_begin, k_off;   // k_off: first operand byte column of this job's K range
    int64_t sq_off, sq_stride;                              // norm of job row a = sqnorm[sq_off + a * sq_stride]
};
constexpr int MAX_JOBS = 32;      // 8 ranks x 4 residue classes; GramArgs stays below the 32 KB launch-parameter limit (CUDA >= 12.1)

struct GramArgs {
    int64_t n, kp;
    const int64_t *sqnorm;
    double *sum;
    unsigned long long *nnz;
    uint32_t idesc;                     // kind::i8 instruction descriptor (signed or unsigned operands)
    int num_jobs, num_tiles;
    int group_m;                        // schedule group (row-tiles) of the 2-CTA kernel
    int l2_hint;                        // TMA L2 policy of the operand loads: 0 none, 1 A evict_last, 2 A and B evict_last
    unsigned long long *clock_probe;    // nullable: [0] SM cycles, [1] ns spent by CTA 0's first epilogue warp
    // K0 fused into this launch (2-CTA kernel, raw u8 frames): the epilogue warps have nothing to do until the first
    // tiles' MMAs retire (0.43 ms at K = 150528), so they compute the norms meanwhile — norm_units byte rows of
    // norm_len bytes, norm_pitch apart, results to norm_out (= sqnorm) / norm_max; norm_sync (zeroed by the caller)
    // counts the warps that have finished, and no epilogue reads sqnorm before all of them have.
    const uint8_t *norm_src;
    int64_t norm_units, norm_len, norm_pitch;
    int64_t *norm_out;
    unsigned long long *norm_max;
    unsigned int *norm_sync;
    GramJob jobs[MAX_JOBS];
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("avtex gram: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128B-swizzled operand tile: rows are 128 B, 8-row groups are 1024 B apart (SBO),
// LBO is ignored for swizzled K-major layouts (set to 1), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ epilogue (shared by both kernels)
struct EpiStats {
    double s;
    unsigned long long z;
};

// One warp drains its 32 TMEM lanes (rows r0 .. r0+31 of the output) x BN accumulator columns:
// d^2 = n_r + n_c - 2g (exact, mod 2^32), fp32 sqrt.  Each lane owns one output ROW, so a direct
// store would write 4-16 B per row per instruction: partial 32 B sectors that L2 has to fill from
// DRAM (measured: 3.3 GB of extra DRAM reads for a 1.6 GB D1).  The 32 x 32 chunk is therefore
// transposed through a per-warp shared-memory tile (33-float pitch, conflict-free both ways) and
// every global store instruction writes one full 128 B line; the mirrored store D[c][r] is
// coalesced as it is (lanes <-> consecutive r).  Symmetric mode handles r <= c only.
// `release()` is called by lane 0 once the warp's last TMEM read has completed, so the MMA warp can
// reuse the accumulator while the stores drain.

// Output stores.  ST = 0: plain; 1: st.global.cs (streaming, evict-first): D1 is written once and not read by this
// kernel, so its lines should not displace the operand tiles that every wave re-reads from L2.
template <int ST>
__device__ __forceinline__ void st_out(float *p, float v) {
    if (ST == 0) *p = v;
    else asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int ST = 0, typename Release>
__device__ __forceinline__ void epilogue_tile(const GramArgs &args, const GramJob &job, uint32_t t_base, int64_t r0,
                                              int64_t c0, int lane, float *tile, Release release, EpiStats &st,
                                              int cc_begin = 0, int cc_end = BN / 32) {
    const int64_t row_end = job.row0 + job.rows, col_end = job.col0 + job.cols;
    const int64_t r = r0 + lane;
    const bool r_ok = r < row_end;
    // __ldcg, not __ldg: with the fused norms the vector is written by other SMs during this very launch
    const uint32_t nr = r_ok ? (uint32_t)__ldcg(args.sqnorm + job.sq_off + r * job.sq_stride) : 0u;
    int64_t rows_here = row_end - r0;                              // valid rows of this warp's 32
    if (rows_here > 32) rows_here = 32;
    const bool sym = job.symmetric != 0;
    // every stored element is counted once; D and DT hold the same values, so when both exist the
    // direct loop counts with weight 2 and the transposed loop not at all (a diagonal entry is 0)
    const int w_direct = (job.D != nullptr) ? ((job.DT != nullptr) ? 2 : 1) : 0;
    const int w_trans = (job.D != nullptr) ? 0 : 1;
#pragma unroll 1
    for (int cc = cc_begin; cc < cc_end; ++cc) {
        uint32_t g[32];
        tmem_ld32(t_base + cc * 32, g);
        if (cc == cc_end - 1) {                                   // all of this warp's reads are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release();
        }
        const int64_t cbase = c0 + cc * 32;
        if (cbase >= col_end || rows_here <= 0) continue;
        if (sym && cbase + 31 < r0) continue;                     // whole chunk below the diagonal for this warp
        const uint32_t nc_lane = (cbase + lane < col_end)
                                     ? (uint32_t)__ldcg(args.sqnorm + job.sq_off + (cbase + lane) * job.sq_stride) : 0u;
        // warp-uniform: all 32 x 32 elements are valid and (symmetric) strictly above the diagonal ->
        // no per-element predicates (the per-element branches of the general path were the critical path
        // of the whole kernel at K = 12288: the epilogue, not the MMAs, set the tile time)
        const bool interior = (rows_here == 32) && (cbase + 32 <= col_end) && (!sym || r0 + 31 < cbase);
        __syncwarp();                                             // previous chunk's tile reads are finished
        float cs = 0.f;                                           // fp32 / int partials per chunk, fp64 once per chunk
        int cz = 0;
        float *dt = (job.DT != nullptr) ? job.DT + (cbase - job.dt_row0) * job.ldt + r : nullptr;
        if (interior) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t nc = __shfl_sync(0xffffffffu, nc_lane, j);
                const float d = __fsqrt_rn(__uint2float_rn(nr + nc - 2u * g[j]));      // d^2 exact mod 2^32
                tile[lane * EPI_PITCH + j] = d;
                if (dt != nullptr) {
                    st_out<ST>(dt, d);                            // transposed store: one 128 B line per warp
                    dt += job.ldt;
                }
                if (w_trans) { cs += d; cz += (d != 0.f); }
            }
            __syncwarp();
            if (job.D != nullptr) {
                float *dst = job.D + (r0 - job.d_row0) * job.ldd + cbase + lane;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float d = tile[i * EPI_PITCH + lane];
                    st_out<ST>(dst, d);                           // direct store: one 128 B line per warp
                    dst += job.ldd;
                    cs += d;
                    cz += (d != 0.f);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t nc = __shfl_sync(0xffffffffu, nc_lane, j);
                const float d = __fsqrt_rn(__uint2float_rn(nr + nc - 2u * g[j]));
                tile[lane * EPI_PITCH + j] = d;
                const int64_t c = cbase + j;
                if (dt != nullptr && r_ok && c < col_end && (!sym || r < c)) {
                    st_out<ST>(dt + j * job.ldt, d);
                    if (w_trans) { cs += d; cz += (d != 0.f); }
                }
            }
            __syncwarp();
            if (job.D != nullptr) {
                const int64_t c = cbase + lane;
                const bool c_ok = c < col_end;
                float *dst = job.D + (r0 - job.d_row0) * job.ldd + c;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < rows_here && c_ok && (!sym || r0 + i <= c)) {
                        const float d = tile[i * EPI_PITCH + lane];
                        st_out<ST>(dst + i * job.ldd, d);
                        if (!sym || r0 + i < c || job.DT == nullptr) { cs += d; cz += (d != 0.f); }
                    }
                }
            }
        }
        if (job.count_stats) {
            const int w = (job.D != nullptr) ? w_direct : 1;
            st.s += (double)(w * cs);
            st.z += (unsigned long long)(w * cz);
        }
    }
}

// tile index over all jobs -> (job, local tile index)
__device__ __forceinline__ int find_job(const GramArgs &args, int t, int *local) {
    int j = 0;
    while (j + 1 < args.num_jobs && t >= args.jobs[j + 1].tile_begin) ++j;
    *local = t - args.jobs[j].tile_begin;
    return j;
}

__device__ __forceinline__ void epilogue_flush(const GramArgs &args, EpiStats &st, int lane) {
    if (args.sum != nullptr) {
        const double s = warp_sum(st.s);
        const unsigned long long z = warp_sum(st.z);
        if (lane == 0) { atomicAdd(args.sum, s); atomicAdd(args.nnz, z); }
    }
}

// ------------------------------------------------------------------ tile schedule
// Row-tiles are taken in groups of GROUP_M; inside a group the order is column-major, so a wave of
// CTAs shares <= GROUP_M A-tiles and ~148/GROUP_M B-tiles: a roughly square super-tile minimises the
// rows a wave must pull through L2 (measured at N=20000, K=12288: DRAM reads 12x the operand size
// with groups of 4 x 256 rows).
// Symmetric mode keeps tile (tm, tn) iff it intersects the upper triangle: tn >= tm / 2.
__host__ __device__ inline int sym_group_count(int g, int TM, int TN, int *gm_out) {
    const int first = g * GROUP_M;
    const int gm = (TM - first < GROUP_M) ? (TM - first) : GROUP_M;
    const int tn0 = first / 2;
    int cnt = 0;
    for (int j = 0; j < GROUP_M / 2; ++j)
        if (tn0 + j < TN) cnt += (gm < 2 * j + 2) ? gm : (2 * j + 2);
    const int full = TN - (tn0 + GROUP_M / 2);
    if (full > 0) cnt += full * gm;
    *gm_out = gm;
    return cnt;
}

__host__ __device__ inline void decode_tile(int t, int TM, int TN, int symmetric, int *tm, int *tn) {
    if (!symmetric) {
        const int per_group = GROUP_M * TN;
        const int g = t / per_group;
        const int first = g * GROUP_M;
        const int gm = (TM - first < GROUP_M) ? (TM - first) : GROUP_M;
        const int rem = t - g * per_group;
        *tn = rem / gm;
        *tm = first + rem % gm;
        return;
    }
    for (int g = 0;; ++g) {
        int gm;
        const int cnt = sym_group_count(g, TM, TN, &gm);
        if (t < cnt) {
            const int first = g * GROUP_M, tn0 = first / 2;
            for (int j = 0; j < GROUP_M / 2; ++j) {
                if (tn0 + j >= TN) break;
                const int c = (gm < 2 * j + 2) ? gm : (2 * j + 2);
                if (t < c) { *tm = first + t; *tn = tn0 + j; return; }
                t -= c;
            }
            *tm = first + t % gm;
            *tn = tn0 + GROUP_M / 2 + t / gm;
            return;
        }
        t -= cnt;
    }
}

int count_tiles(int TM, int TN, int symmetric) {
    if (!symmetric) return TM * TN;
    int total = 0, gm;
    for (int g = 0; g * GROUP_M < TM; ++g) total += sym_group_count(g, TM, TN, &gm);
    return total;
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_l2_s8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const GramArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                     // SWIZZLE_128B needs 1024 B alignment
    const uint32_t bars = tiles + STAGES * STAGE_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
    const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = int((args.kp + BKB - 1) / BKB);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
                int tm, tn, lt;
                const GramJob &job = args.jobs[find_job(args, t, &lt)];
                decode_tile(lt, job.TM, job.TN, job.symmetric, &tm, &tn);
                const int row_a = int(job.row0) + tm * BM, row_b = int(job.col0) + tn * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = tiles + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(full_bar + 8 * stage, STAGE_BYTES);
                    tma_load_2d(sa, &map_a, full_bar + 8 * stage, job.k_off + kb * BKB, row_a);
                    tma_load_2d(sb, &map_b, full_bar + 8 * stage, job.k_off + kb * BKB, row_b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);          // epilogue drained this TMEM buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);               // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = tiles + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < BKB / UMMA_KB; ++k)
                        umma_s8(d_tmem, da + uint64_t(k * (UMMA_KB >> 4)), db + uint64_t(k * (UMMA_KB >> 4)),
                                args.idesc, (kb | k) != 0);
                    umma_commit(empty_bar + 8 * stage);                   // frees the smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar + 8 * acc);                         // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                                     // TMEM lane quarter this warp may read
        float *epi_tile = reinterpret_cast<float *>(smem_raw + (bars + 256 - raw)) + (warp - 2) * 32 * EPI_PITCH;
        int acc = 0;
        uint32_t acc_phase = 0;
        EpiStats st{0.0, 0ull};
        for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
            int tm, tn, lt;
            const GramJob &job = args.jobs[find_job(args, t, &lt)];
            decode_tile(lt, job.TM, job.TN, job.symmetric, &tm, &tn);
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN);
            const uint32_t release = tempty_bar + 8 * acc;
            epilogue_tile(args, job, t_base, job.row0 + int64_t(tm) * BM + quarter * 32, job.col0 + int64_t(tn) * BN,
                          lane, epi_tile, [release]() { mbar_arrive(release); }, st);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        epilogue_flush(args, st, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ 2-CTA (cta_group::2) variant
// A CTA pair (cluster of 2, same TPC) computes one 256 x 256 tile: each CTA holds 128 rows of A and
// 128 of the 256 B rows in its own shared memory, the leader issues tcgen05.mma.cta_group::2 (M = 256)
// and each CTA's TMEM receives its own 128 output rows.  Per CTA and per 128-byte K block this halves
// the operand bytes TMA must write and the tensor core must read from shared memory (32 KB instead
// of 48 KB for the same 512 MMA cycles): the 1-CTA kernel is shared-memory-bandwidth bound
// (16 KB A + 32 KB B read + 48 KB TMA fill per 512 MMA cycles > 128 B/cycle).
constexpr int BM2 = 256;                              // rows per cluster tile
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = 2 * A_BYTES;             // A half (128 rows) + B half (128 rows)
// Row-tiles per schedule group.  A wave of 74 CTA pairs then touches G A-tiles and ~74/G B-tiles; every B tile is
// re-read once per group of rows, so DRAM reads ~ operand bytes x ceil(TM / G).  8 for long K (a 256-row operand
// tile is 38 MB at K = 150528: only the current K window of each tile lives in L2 anyway), 16 for K <= 16384
// (3 MB tiles: 16 + 5 of them fit in L2; measured at K = 12288 with G = 8: 4.2x the operand bytes from DRAM).
constexpr int GROUP_M2_DEFAULT = 8;
// 8 epilogue warps (two per TMEM lane quarter, 128 accumulator columns each): with four, one warp per scheduler
// had to drain 128 x 256 outputs per tile alone, and at K = 12288 (25 us of MMAs per tile) the epilogue, not the
// tensor pipe, set the tile time (tensor pipe 44-49 % active, L2 35 %: neither memory nor math bound)
constexpr int EPI_WARPS2 = 8;
constexpr int NUM_THREADS2 = 64 + 32 * EPI_WARPS2;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + EPI_WARPS2 * 32 * EPI_PITCH * 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}"
        ::"r"(bar), "r"(cta) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, completion bytes are credited to the
// leader CTA's mbarrier (peer bit of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_hint(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                                     uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {          // arrives in both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_s8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// 256 x 256 tiles; symmetric mode keeps (tm, tn) iff tn >= tm.  Groups of GROUP_M2 row-tiles,
// column-major inside a group (same L2 argument as the 1-CTA schedule).
__host__ __device__ inline int sym2_group_count(int g, int TM, int TN, int *gm_out, int GROUP_M2 = GROUP_M2_DEFAULT) {
    const int first = g * GROUP_M2;
    const int gm = (TM - first < GROUP_M2) ? (TM - first) : GROUP_M2;
    int cnt = 0;
    for (int j = 0; j < GROUP_M2; ++j)
        if (first + j < TN) cnt += (gm < j + 1) ? gm : (j + 1);
    const int full = TN - (first + GROUP_M2);
    if (full > 0) cnt += full * gm;
    *gm_out = gm;
    return cnt;
}

__host__ __device__ inline void decode_tile2(int t, int TM, int TN, int symmetric, int *tm, int *tn,
                                             int GROUP_M2 = GROUP_M2_DEFAULT) {
    if (!symmetric) {
        const int per_group = GROUP_M2 * TN;
        const int g = t / per_group;
        const int first = g * GROUP_M2;
        const int gm = (TM - first < GROUP_M2) ? (TM - first) : GROUP_M2;
        const int rem = t - g * per_group;
        *tn = rem / gm;
        *tm = first + rem % gm;
        return;
    }
    for (int g = 0;; ++g) {
        int gm;
        const int cnt = sym2_group_count(g, TM, TN, &gm, GROUP_M2);
        if (t < cnt) {
            const int first = g * GROUP_M2;
            for (int j = 0; j < GROUP_M2; ++j) {
                if (first + j >= TN) break;
                const int c = (gm < j + 1) ? gm : (j + 1);
                if (t < c) { *tm = first + t; *tn = first + j; return; }
                t -= c;
            }
            *tm = first + t % gm;
            *tn = first + GROUP_M2 + t / gm;
            return;
        }
        t -= cnt;
    }
}

// The same decode for a caller whose tile indices only grow (every role of the kernel walks t, t + stride, ...):
// the cursor remembers the group reached so far, so the search over groups is amortised O(1) per tile.  Scanning
// from group 0 for every tile (decode_tile2) costs ~25 x 16 loop iterations at N = 100000 — about 3 us of
// dependent integer work in the single TMA-producer thread at every tile boundary, longer than the 2.2 us of
// operand stages the ring buffers ahead.
struct TileCursor {
    int job = -1, g = 0, base = 0;
};
__device__ __forceinline__ void decode_tile2_cursor(TileCursor &cur, int job_id, int t, int TM, int TN, int symmetric,
                                                    int *tm, int *tn, int GROUP_M2) {
    if (!symmetric) { decode_tile2(t, TM, TN, 0, tm, tn, GROUP_M2); return; }
    if (cur.job != job_id) { cur.job = job_id; cur.g = 0; cur.base = 0; }
    int gm, cnt = sym2_group_count(cur.g, TM, TN, &gm, GROUP_M2);
    while (t - cur.base >= cnt) {
        cur.base += cnt;
        ++cur.g;
        cnt = sym2_group_count(cur.g, TM, TN, &gm, GROUP_M2);
    }
    int r = t - cur.base;
    const int first = cur.g * GROUP_M2;
    for (int j = 0; j < GROUP_M2; ++j) {
        if (first + j >= TN) break;
        const int c = (gm < j + 1) ? gm : (j + 1);
        if (r < c) { *tm = first + r; *tn = first + j; return; }
        r -= c;
    }
    *tm = first + r % gm;
    *tn = first + GROUP_M2 + r / gm;
}

int count_tiles2(int TM, int TN, int symmetric, int GROUP_M2 = GROUP_M2_DEFAULT) {
    if (!symmetric) return TM * TN;
    int total = 0, gm;
    for (int g = 0; g * GROUP_M2 < TM; ++g) total += sym2_group_count(g, TM, TN, &gm, GROUP_M2);
    return total;
}

// 16 raw bytes: sum of squares and sum (for the centred norm) by dp4a
__device__ __forceinline__ void norm_accum16(const uint4 v, unsigned long long &sq, unsigned long long &sm) {
    unsigned int q = 0, t = 0;
    q = __dp4a(v.x, v.x, q); q = __dp4a(v.y, v.y, q); q = __dp4a(v.z, v.z, q); q = __dp4a(v.w, v.w, q);
    t = __dp4a(v.x, 0x01010101u, t); t = __dp4a(v.y, 0x01010101u, t);
    t = __dp4a(v.z, 0x01010101u, t); t = __dp4a(v.w, 0x01010101u, t);
    sq += q;
    sm += t;
}

// K0 inside K1 (see GramArgs): epilogue warp `ew` of `EW` takes the byte rows ew, ew + EW, ...; one warp streams one
// row with 8 x 512 B in flight.  Returns when EVERY epilogue warp of the grid has published its norms (all CTAs of
// this persistent launch are co-resident: one per SM).
__device__ __forceinline__ void fused_norms(const GramArgs &args, int ew, int EW, int lane) {
    unsigned long long wmax = 0;
    const int64_t len = args.norm_len, kv = len & ~int64_t(15);
    for (int64_t u = ew; u < args.norm_units; u += EW) {
        const uint8_t *src = args.norm_src + u * args.norm_pitch;
        unsigned long long sq = 0, sm = 0;
        int64_t c = int64_t(lane) * 16;
        for (; c + 7 * 512 < kv; c += 8 * 512) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const uint4 *>(src + c + i * 512));
#pragma unroll
            for (int i = 0; i < 8; ++i) norm_accum16(v[i], sq, sm);
        }
        for (; c < kv; c += 512) norm_accum16(__ldg(reinterpret_cast<const uint4 *>(src + c)), sq, sm);
        for (int64_t t = kv + lane; t < len; t += 32) {
            const unsigned int x = src[t];
            sq += x * x;
            sm += x;
        }
        sq = warp_sum(sq);
        sm = warp_sum(sm);
        if (lane == 0) {
            args.norm_out[u] = (int64_t)sq;
            const unsigned long long centred = sq + 16384ull * (unsigned long long)len - 256ull * sm;
            wmax = centred > wmax ? centred : wmax;
        }
    }
    if (lane == 0) {
        if (args.norm_max != nullptr && wmax != 0) atomicMax(args.norm_max, wmax);
        __threadfence();
        atomicAdd(args.norm_sync, 1u);
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned int *>(args.norm_sync) < (unsigned int)EW) {
            if (clock64() - t0 > 8000000000LL) {
                printf("avtex gram: fused-norms barrier timeout (block %d)\n", blockIdx.x);
                __trap();
            }
        }
        __threadfence();
    }
    __syncwarp();
}

template <int ST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
gram_l2_s8_2cta_kernel(const __grid_constant__ CUtensorMap map, const GramArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + STAGES2 * STAGE2_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES2;
    const uint32_t tfull_bar = bars + 16 * STAGES2, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int KB = int((args.kp + BKB - 1) / BKB);
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
        for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar + 8 * s, 2); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 2 * EPI_WARPS2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                                   // peer barriers initialised, both TMEM allocations done
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint64_t pol_last = 0;
            if (args.l2_hint) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            TileCursor cur;
            for (int t = cluster_id; t < args.num_tiles; t += num_clusters) {
                int tm, tn, lt;
                const int jid = find_job(args, t, &lt);
                const GramJob &job = args.jobs[jid];
                decode_tile2_cursor(cur, jid, lt, job.TM, job.TN, job.symmetric, &tm, &tn, args.group_m);
                int row_a = int(job.row0) + tm * BM2 + int(rank) * BM;
                int row_b = int(job.col0) + tn * BN + int(rank) * (BN / 2);
                if (row_a >= args.n) row_a = 0;           // fully out-of-range half: load anything, stores are masked
                if (row_b >= args.n) row_b = 0;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = tiles + stage * STAGE2_BYTES, sb = sa + A_BYTES;
                    if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * STAGE2_BYTES);
                    else mbar_arrive_cluster(full_bar + 8 * stage, 0);
                    const int kc = job.k_off + kb * BKB;
                    if (args.l2_hint >= 1) tma_load_2d_2sm_hint(sa, &map, full_bar + 8 * stage, kc, row_a, pol_last);
                    else tma_load_2d_2sm(sa, &map, full_bar + 8 * stage, kc, row_a);
                    if (args.l2_hint >= 2) tma_load_2d_2sm_hint(sb, &map, full_bar + 8 * stage, kc, row_b, pol_last);
                    else tma_load_2d_2sm(sb, &map, full_bar + 8 * stage, kc, row_b);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = cluster_id; t < args.num_tiles; t += num_clusters) {
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);          // both CTAs' epilogues drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);               // both CTAs' TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = tiles + stage * STAGE2_BYTES, sb = sa + A_BYTES;
                    const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < BKB / UMMA_KB; ++k)
                        umma_s8_2cta(d_tmem, da + uint64_t(k * (UMMA_KB >> 4)), db + uint64_t(k * (UMMA_KB >> 4)),
                                     args.idesc, (kb | k) != 0);
                    umma_commit_2cta(empty_bar + 8 * stage);              // frees the stage in both CTAs
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(tfull_bar + 8 * acc);                    // accumulators complete in both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs: own 128 rows) =====================
        // warp w may read TMEM lanes 32*(w%4)..+31: warps 2-5 take accumulator columns 0..127, warps 6-9 128..255
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        float *epi_tile = reinterpret_cast<float *>(smem_raw + (bars + 256 - raw)) + (warp - 2) * 32 * EPI_PITCH;
        int acc = 0;
        uint32_t acc_phase = 0;
        EpiStats st{0.0, 0ull};
        // in-kernel clock probe: SM cycles and wall nanoseconds over this CTA's whole tile loop give the
        // SM clock the kernel actually ran at (NVML's 10 ms samples cannot see a 1 ms kernel)
        const bool probe = (args.clock_probe != nullptr) && blockIdx.x == 0 && warp == 2 && lane == 0;   // (one of 8 epilogue warps)
        long long c0 = 0;
        unsigned long long g0 = 0;
        if (probe) { c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0)); }
        if (args.norm_src != nullptr) fused_norms(args, int(blockIdx.x) * EPI_WARPS2 + (warp - 2), int(gridDim.x) * EPI_WARPS2, lane);
        TileCursor cur;
        for (int t = cluster_id; t < args.num_tiles; t += num_clusters) {
            int tm, tn, lt;
            const int jid = find_job(args, t, &lt);
            const GramJob &job = args.jobs[jid];
            decode_tile2_cursor(cur, jid, lt, job.TM, job.TN, job.symmetric, &tm, &tn, args.group_m);
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN);
            const uint32_t release = tempty_bar + 8 * acc;
            epilogue_tile<ST>(args, job, t_base, job.row0 + int64_t(tm) * BM2 + int64_t(rank) * BM + quarter * 32,
                          job.col0 + int64_t(tn) * BN, lane, epi_tile,
                          [release]() { mbar_arrive_cluster(release, 0); }, st, half * (BN / 64), (half + 1) * (BN / 64));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (probe) {
            unsigned long long g1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            args.clock_probe[0] = (unsigned long long)(clock64() - c0);
            args.clock_probe[1] = g1 - g0;
        }
        epilogue_flush(args, st, lane);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn *out) {
    static std::atomic<EncodeTiledFn> cached{nullptr};      // idempotent: a racing second lookup stores the same pointer
    if (cached.load(std::memory_order_acquire) == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        AVTEX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        AVTEX_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                      "cuTensorMapEncodeTiled not available from the driver");
        cached.store(reinterpret_cast<EncodeTiledFn>(fn), std::memory_order_release);
    }
    *out = cached.load(std::memory_order_acquire);
    return 0;
}

int make_map(EncodeTiledFn enc, CUtensorMap *map, const void *base, int64_t n, int64_t k_extent, int64_t pitch,
             int box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)k_extent, (cuuint64_t)n};       // columns beyond k_extent read as 0
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)BKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVTEX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

struct NormFusion {
    int64_t units = 0, pitch = 0;              // units == 0: norms are already in sqnorm
    unsigned long long *max_centred = nullptr;
    unsigned int *sync = nullptr;
};

int launch_gram_jobs(const void *operand, bool is_signed, int64_t n, int64_t k_extent, int64_t pitch,
                     const int64_t *sqnorm, const AvtexGramJob *jobs, int num_jobs, double *sum,
                     unsigned long long *nnz, unsigned long long *clock_probe, int device, void *stream,
                     const NormFusion &nf = NormFusion()) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(n >= 1 && n < (int64_t(1) << 30) && k_extent >= 1 && k_extent < (int64_t(1) << 31),
                  "gram_l2: bad shape n=%lld k=%lld", (long long)n, (long long)k_extent);
    AVTEX_REQUIRE(pitch >= k_extent && pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(operand) & 15) == 0,
                  "gram_l2: operand rows must be 16-byte aligned (pitch %lld)", (long long)pitch);
    AVTEX_REQUIRE(num_jobs >= 1 && num_jobs <= MAX_JOBS, "gram_l2: 1..%d jobs per launch (got %d)", MAX_JOBS, num_jobs);
    AVTEX_REQUIRE((sum == nullptr) == (nnz == nullptr), "gram_l2: sum and nnz go together");
    int cc = 0, sms = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    AVTEX_REQUIRE(cc == 100, "gram_l2: needs an sm_100 device (tcgen05 kind::i8), got cc %d", cc);
    const char *mode = getenv("AVTEX_GRAM_MODE");                  // "1cta" selects the single-CTA kernel
    const bool two_cta = !(mode != nullptr && mode[0] == '1');
    const int bm = two_cta ? BM2 : BM;

    GramArgs a;
    a.n = n; a.kp = k_extent; a.sqnorm = sqnorm; a.sum = sum; a.nnz = nnz;
    a.idesc = make_idesc(bm, is_signed);
    a.num_jobs = num_jobs;
    a.clock_probe = clock_probe;
    a.norm_src = nullptr; a.norm_units = 0; a.norm_len = 0; a.norm_pitch = 0; a.norm_out = nullptr; a.norm_max = nullptr;
    a.norm_sync = nullptr;
    if (nf.units > 0) {
        AVTEX_REQUIRE(two_cta && !is_signed, "gram_l2: fused norms need the 2-CTA kernel and raw uint8 frames");
        AVTEX_REQUIRE(nf.sync != nullptr && nf.pitch >= k_extent && nf.pitch % 16 == 0,
                      "gram_l2: fused norms need a zeroed sync word and 16-byte aligned frame rows");
        a.norm_src = static_cast<const uint8_t *>(operand);
        a.norm_units = nf.units; a.norm_len = k_extent; a.norm_pitch = nf.pitch;
        a.norm_out = const_cast<int64_t *>(sqnorm); a.norm_max = nf.max_centred; a.norm_sync = nf.sync;
    }
    a.group_m = (k_extent <= 16384) ? 16 : GROUP_M2_DEFAULT;
    a.l2_hint = 0;
    int st_mode = 0;
    if (const char *e = getenv("AVTEX_GRAM_GROUP")) { const int g = atoi(e); if (g >= 1 && g <= 64) a.group_m = g; }
    if (const char *e = getenv("AVTEX_GRAM_HINT")) a.l2_hint = atoi(e);
    if (const char *e = getenv("AVTEX_GRAM_ST")) st_mode = atoi(e);
    int total = 0;
    int64_t map_k = k_extent;                                       // columns the tensor map must expose
    for (int j = 0; j < num_jobs; ++j) {
        const AvtexGramJob &in = jobs[j];
        AVTEX_REQUIRE(in.rows >= 1 && in.cols >= 1 && in.row0 >= 0 && in.col0 >= 0 && in.row0 + in.rows <= n &&
                          in.col0 + in.cols <= n,
                      "gram_l2: job %d rectangle [%lld,+%lld) x [%lld,+%lld) outside the %lld frames", j,
                      (long long)in.row0, (long long)in.rows, (long long)in.col0, (long long)in.cols, (long long)n);
        AVTEX_REQUIRE(!in.symmetric || (in.row0 == in.col0 && in.rows == in.cols),
                      "gram_l2: job %d is symmetric but its row and column ranges differ", j);
        AVTEX_REQUIRE(in.D != nullptr || in.DT != nullptr, "gram_l2: job %d has no destination", j);
        AVTEX_REQUIRE(in.D == nullptr || (in.ldd >= in.col0 + in.cols && in.d_row0 <= in.row0),
                      "gram_l2: job %d direct destination too small", j);
        AVTEX_REQUIRE(in.DT == nullptr || (in.ldt >= in.row0 + in.rows && in.dt_row0 <= in.col0),
                      "gram_l2: job %d transposed destination too small", j);
        GramJob &o = a.jobs[j];
        o.row0 = in.row0; o.rows = in.rows; o.col0 = in.col0; o.cols = in.cols;
        o.D = in.D; o.d_row0 = in.d_row0; o.ldd = in.ldd;
        o.DT = in.DT; o.dt_row0 = in.dt_row0; o.ldt = in.ldt;
        o.symmetric = in.symmetric ? 1 : 0;
        o.count_stats = (in.count_stats && sum != nullptr) ? 1 : 0;
        o.TM = int((in.rows + bm - 1) / bm);
        o.TN = int((in.cols + BN - 1) / BN);
        o.tile_begin = total;
        AVTEX_REQUIRE(in.k_off >= 0 && in.k_off + k_extent <= pitch && (in.k_off == 0 || k_extent % BKB == 0) &&
                          in.sq_off >= 0 && in.sq_stride >= 0,
                      "gram_l2: job %d residue fields (k_off %lld needs K %% 128 == 0 and k_off + K <= ld)", j,
                      (long long)in.k_off);
        o.k_off = (int)in.k_off;
        o.sq_off = in.sq_off;
        o.sq_stride = in.sq_stride > 0 ? in.sq_stride : 1;
        if (in.k_off + k_extent > map_k) map_k = in.k_off + k_extent;
        total += two_cta ? count_tiles2(o.TM, o.TN, o.symmetric, a.group_m) : count_tiles(o.TM, o.TN, o.symmetric);
    }
    a.num_tiles = total;

    EncodeTiledFn enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    if (two_cta) {
        CUtensorMap map;
        if (int rc = make_map(enc, &map, operand, n, map_k, pitch, BM)) return rc;
        // cudaFuncSetAttribute is idempotent: the atomic flag only skips repeats, two racing host threads both succeed
        static std::atomic<bool> attr2_set[64];
        if (device >= 64 || !attr2_set[device].load(std::memory_order_acquire)) {
            AVTEX_CUDA(cudaFuncSetAttribute(gram_l2_s8_2cta_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
            AVTEX_CUDA(cudaFuncSetAttribute(gram_l2_s8_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
            if (device < 64) attr2_set[device].store(true, std::memory_order_release);
        }
        int clusters = sms / 2;
        if (a.num_tiles < clusters) clusters = a.num_tiles;
        if (st_mode == 1) gram_l2_s8_2cta_kernel<1><<<2 * clusters, NUM_THREADS2, SMEM2_BYTES, as_stream(stream)>>>(map, a);
        else gram_l2_s8_2cta_kernel<0><<<2 * clusters, NUM_THREADS2, SMEM2_BYTES, as_stream(stream)>>>(map, a);
        AVTEX_LAUNCH_CHECK();
        return 0;
    }
    CUtensorMap map_a, map_b;
    if (int rc = make_map(enc, &map_a, operand, n, map_k, pitch, BM)) return rc;
    if (int rc = make_map(enc, &map_b, operand, n, map_k, pitch, BN)) return rc;
    static std::atomic<bool> attr_set[64];
    if (device >= 64 || !attr_set[device].load(std::memory_order_acquire)) {
        AVTEX_CUDA(cudaFuncSetAttribute(gram_l2_s8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        if (device < 64) attr_set[device].store(true, std::memory_order_release);
    }
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    gram_l2_s8_kernel<<<grid, NUM_THREADS, SMEM_BYTES, as_stream(stream)>>>(map_a, map_b, a);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

// The two classic shapes as one job: the full symmetric matrix, or a row block against all columns.
int launch_gram(const void *operand, bool is_signed, int64_t n, int64_t k_extent, int64_t pitch, const int64_t *sqnorm,
                int64_t row0, int64_t rows, int symmetric, float *D, int64_t ldd, double *sum, unsigned long long *nnz,
                int device, void *stream) {
    AVTEX_REQUIRE(rows >= 1 && row0 >= 0 && row0 + rows <= n && ldd >= n, "gram_l2: bad row block [%lld,+%lld)",
                  (long long)row0, (long long)rows);
    AVTEX_REQUIRE(!symmetric || (row0 == 0 && rows == n), "gram_l2: symmetric mode needs the full matrix");
    AvtexGramJob job;
    job.row0 = row0; job.rows = rows; job.col0 = 0; job.cols = n;
    job.D = D; job.d_row0 = row0; job.ldd = ldd;
    job.DT = symmetric ? D : nullptr; job.dt_row0 = 0; job.ldt = ldd;
    job.symmetric = symmetric ? 1 : 0;
    job.count_stats = 1;
    job.k_off = 0; job.sq_off = 0; job.sq_stride = 1;
    return launch_gram_jobs(operand, is_signed, n, k_extent, pitch, sqnorm, &job, 1, sum, nnz, nullptr, device, stream);
}

}  // namespace

extern "C" int avtex_gram_l2_s8(const int8_t *packed, int64_t n, int64_t kp, const int64_t *sqnorm,
                                int64_t row0, int64_t rows, int symmetric, float *D, int64_t ldd,
                                double *sum, unsigned long long *nnz, int device, void *stream) {
    return launch_gram(packed, true, n, kp, kp, sqnorm, row0, rows, symmetric, D, ldd, sum, nnz, device, stream);
}

extern "C" int avtex_gram_l2_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, const int64_t *sqnorm,
                                int64_t row0, int64_t rows, int symmetric, float *D, int64_t ldd,
                                double *sum, unsigned long long *nnz, int device, void *stream) {
    return launch_gram(frames, false, n, k, ld, sqnorm, row0, rows, symmetric, D, ldd, sum, nnz, device, stream);
}

extern "C" int avtex_gram_l2_jobs(const void *operand, int operand_signed, int64_t n, int64_t k, int64_t ld,
                                  const int64_t *sqnorm, const AvtexGramJob *h_jobs, int num_jobs, double *sum,
                                  unsigned long long *nnz, unsigned long long *clock_probe, int device, void *stream) {
    return launch_gram_jobs(operand, operand_signed != 0, n, k, ld, sqnorm, h_jobs, num_jobs, sum, nnz, clock_probe, device,
                            stream);
}

extern "C" int avtex_gram_l2_jobs_fused_norms(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t norm_units,
                                              int64_t norm_pitch, int64_t *sqnorm, unsigned long long *max_centred,
                                              unsigned int *sync_zeroed, const AvtexGramJob *h_jobs, int num_jobs,
                                              double *sum, unsigned long long *nnz, int device, void *stream) {
    AVTEX_REQUIRE(norm_units >= 1 && sqnorm != nullptr, "gram_l2 (fused norms): norm_units >= 1 and a sqnorm vector");
    NormFusion nf;
    nf.units = norm_units; nf.pitch = norm_pitch; nf.max_centred = max_centred; nf.sync = sync_zeroed;
    return launch_gram_jobs(frames, false, n, k, ld, sqnorm, h_jobs, num_jobs, sum, nnz, nullptr, device, stream, nf);
}

extern "C" int avtex_gram_tile_schedule2(int TM, int TN, int symmetric, int *tm_out, int *tn_out, int capacity) {
    const int total = count_tiles2(TM, TN, symmetric);
    if (tm_out == nullptr) return total;
    for (int t = 0; t < total && t < capacity; ++t) decode_tile2(t, TM, TN, symmetric, &tm_out[t], &tn_out[t]);
    return total;
}

extern "C" int avtex_gram_tile_schedule2g(int TM, int TN, int symmetric, int group, int *tm_out, int *tn_out, int capacity) {
    if (group < 1 || group > 64) return -1;
    const int total = count_tiles2(TM, TN, symmetric, group);
    if (tm_out == nullptr) return total;
    for (int t = 0; t < total && t < capacity; ++t) decode_tile2(t, TM, TN, symmetric, &tm_out[t], &tn_out[t], group);
    return total;
}

// Exposed for tests: the tile schedule must cover every needed tile exactly once.
extern "C" int avtex_gram_tile_schedule(int TM, int TN, int symmetric, int *tm_out, int *tn_out, int capacity) {
    const int total = count_tiles(TM, TN, symmetric);
    if (tm_out == nullptr) return total;
    for (int t = 0; t < total && t < capacity; ++t) decode_tile(t, TM, TN, symmetric, &tm_out[t], &tn_out[t]);
    return total;
}

// K0 — frame packing: [n,k] u8 / integer-valued f32  ->  centred s8 [n,kp] + exact squared norms.
// HBM-bound: reads n*k*(1|4) B, writes n*kp B + 8n B.  One CTA per frame row, 128-bit accesses.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int PACK_THREADS = 256;

// 16 centred bytes from 16 raw u8: x - 128 == x ^ 0x80 reinterpreted as s8.
__device__ __forceinline__ uint4 centre16(uint4 v) {
    v.x ^= 0x80808080u; v.y ^= 0x80808080u; v.z ^= 0x80808080u; v.w ^= 0x80808080u;
    return v;
}

__device__ __forceinline__ int sq4(unsigned int w) {            // sum of squares of 4 packed s8
    return __dp4a((int)w, (int)w, 0);
}

__global__ void __launch_bounds__(PACK_THREADS)
pack_u8_kernel(const uint8_t *__restrict__ in, int64_t k, int64_t ld, int8_t *__restrict__ out,
               int64_t kp, int64_t *__restrict__ sqnorm, unsigned long long *__restrict__ max_sqnorm) {
    __shared__ unsigned long long scratch[32];
    const int64_t row = blockIdx.x;
    const uint8_t *src = in + row * ld;
    int8_t *dst = out + row * kp;
    unsigned long long acc = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int64_t kv = vec ? (k & ~int64_t(15)) : 0;
    for (int64_t c = int64_t(threadIdx.x) * 16; c < kv; c += int64_t(PACK_THREADS) * 16) {
        uint4 v = centre16(*reinterpret_cast<const uint4 *>(src + c));
        acc += (unsigned)(sq4(v.x) + sq4(v.y) + sq4(v.z) + sq4(v.w));   // <= 16*16384
        *reinterpret_cast<uint4 *>(dst + c) = v;
    }
    for (int64_t c = kv + threadIdx.x; c < kp; c += PACK_THREADS) {
        int v = (c < k) ? int(src[c]) - 128 : 0;
        acc += (unsigned)(v * v);
        dst[c] = (int8_t)v;
    }
    acc = block_reduce(acc, 0ull, OpAdd<unsigned long long>(), scratch);
    if (threadIdx.x == 0) {
        sqnorm[row] = (int64_t)acc;
        if (max_sqnorm != nullptr) atomicMax(max_sqnorm, acc);
    }
}

__global__ void __launch_bounds__(PACK_THREADS)
pack_f32_kernel(const float *__restrict__ in, int64_t k, int64_t ld, int8_t *__restrict__ out,
                int64_t kp, int64_t *__restrict__ sqnorm, int *__restrict__ flags,
                unsigned long long *__restrict__ max_sqnorm) {
    __shared__ unsigned long long scratch[32];
    const int64_t row = blockIdx.x;
    const float *src = in + row * ld;
    int8_t *dst = out + row * kp;
    unsigned long long acc = 0;
    int bad = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int64_t kv = vec ? (k & ~int64_t(15)) : 0;
    for (int64_t c = int64_t(threadIdx.x) * 16; c < kv; c += int64_t(PACK_THREADS) * 16) {
        unsigned int w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 f = ld_stream_f4(src + c + 4 * q);
            float e[4] = {f.x, f.y, f.z, f.w};
            unsigned int pk = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float r = rintf(e[b]);
                bad |= !(r == e[b] && r >= 0.f && r <= 255.f);
                int v = (int)r - 128;
                acc += (unsigned)(v * v);
                pk |= (unsigned)(v & 0xff) << (8 * b);
            }
            w[q] = pk;
        }
        *reinterpret_cast<uint4 *>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int64_t c = kv + threadIdx.x; c < kp; c += PACK_THREADS) {
        int v = 0;
        if (c < k) {
            float e = src[c], r = rintf(e);
            bad |= !(r == e && r >= 0.f && r <= 255.f);
            v = (int)r - 128;
        }
        acc += (unsigned)(v * v);
        dst[c] = (int8_t)v;
    }
    acc = block_reduce(acc, 0ull, OpAdd<unsigned long long>(), scratch);
    if (threadIdx.x == 0) {
        sqnorm[row] = (int64_t)acc;
        if (max_sqnorm != nullptr) atomicMax(max_sqnorm, acc);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, 1);
}

// Norms only (no packed copy): the Gram kernel can contract the raw unsigned bytes directly
// (kind::i8 with unsigned operands, see gram.cu), which needs sum x^2 per frame for the epilogue and the
// CENTRED norm sum (x-128)^2 = sum x^2 - 256 sum x + 128^2 k for the d^2 < 2^32 domain guard.
// Results can be written to up to 8 destinations (the same vector on every GPU of the box, through
// peer-mapped pointers): the row-sharded path computes 1/G of the norms per rank and PUSHES them to its
// peers from this kernel instead of running an all-gather + all-reduce afterwards.
struct NormDst {
    int64_t *sqnorm[8];                 // base of the full [N] vector on destination d
    unsigned long long *max_centred[8]; // nullable
    int count;
};

__device__ __forceinline__ void norms_accum16(const uint4 v, unsigned long long &sq, unsigned long long &sm) {
    unsigned int q = 0, t = 0;
    q = __dp4a(v.x, v.x, q); q = __dp4a(v.y, v.y, q); q = __dp4a(v.z, v.z, q); q = __dp4a(v.w, v.w, q);
    t = __dp4a(v.x, 0x01010101u, t); t = __dp4a(v.y, 0x01010101u, t);
    t = __dp4a(v.z, 0x01010101u, t); t = __dp4a(v.w, 0x01010101u, t);
    sq += q;
    sm += t;
}

__device__ __forceinline__ void norms_store(const NormDst &dst, int64_t grow, int64_t k, unsigned long long sq,
                                            unsigned long long sm) {
    const unsigned long long centred = sq + 16384ull * (unsigned long long)k - 256ull * sm;
    for (int d = 0; d < dst.count; ++d) {
        dst.sqnorm[d][grow] = (int64_t)sq;
        if (dst.max_centred[d] != nullptr) atomicMax(dst.max_centred[d], centred);
    }
}

// One CTA streams one row at a time (256 threads x 16 B = 4 KB of CONTIGUOUS bytes per load instruction, four
// loads in flight per thread) and strides over the rows persistently.  Measured on 12 KB rows (a 64x64 RGB
// frame): splitting the CTA over several rows (32 / 64 / 128 threads per row, more bytes in flight) was SLOWER
// (0.59 vs 0.70 of the HBM peak) — many short concurrent streams cost DRAM locality.
__global__ void __launch_bounds__(PACK_THREADS)
frame_norms_u8_kernel(const uint8_t *__restrict__ in, int64_t n, int64_t k, int64_t ld, int64_t row0, const NormDst dst) {
    __shared__ unsigned long long scratch[32];
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
        const uint8_t *src = in + row * ld;
        unsigned long long sq = 0, sm = 0;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const int64_t kv = vec ? (k & ~int64_t(15)) : 0;
        constexpr int64_t W = int64_t(PACK_THREADS) * 16;
        int64_t c = int64_t(threadIdx.x) * 16;
        for (; c + 3 * W < kv; c += 4 * W) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(src + c)),
                        v1 = __ldg(reinterpret_cast<const uint4 *>(src + c + W)),
                        v2 = __ldg(reinterpret_cast<const uint4 *>(src + c + 2 * W)),
                        v3 = __ldg(reinterpret_cast<const uint4 *>(src + c + 3 * W));
            norms_accum16(v0, sq, sm); norms_accum16(v1, sq, sm); norms_accum16(v2, sq, sm); norms_accum16(v3, sq, sm);
        }
        if (c + W < kv) {                                   // 2 or 3 remaining: both/all in flight together
            const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(src + c)),
                        v1 = __ldg(reinterpret_cast<const uint4 *>(src + c + W));
            uint4 v2 = make_uint4(0, 0, 0, 0);
            if (c + 2 * W < kv) v2 = __ldg(reinterpret_cast<const uint4 *>(src + c + 2 * W));
            norms_accum16(v0, sq, sm); norms_accum16(v1, sq, sm); norms_accum16(v2, sq, sm);
        } else if (c < kv) {
            norms_accum16(__ldg(reinterpret_cast<const uint4 *>(src + c)), sq, sm);
        }
        for (int64_t u = kv + threadIdx.x; u < k; u += PACK_THREADS) {
            const unsigned int v = src[u];
            sq += v * v;
            sm += v;
        }
        sq = block_reduce(sq, 0ull, OpAdd<unsigned long long>(), scratch);
        sm = block_reduce(sm, 0ull, OpAdd<unsigned long long>(), scratch);
        if (threadIdx.x == 0) norms_store(dst, row0 + row, k, sq, sm);
    }
}

}  // namespace

static int check_pack(int64_t n, int64_t k, int64_t ld, int64_t kp) {
    AVTEX_REQUIRE(n > 0 && k > 0 && ld >= k, "pack_frames: bad shape n=%lld k=%lld ld=%lld",
                  (long long)n, (long long)k, (long long)ld);
    AVTEX_REQUIRE(kp >= k && kp % 128 == 0, "pack_frames: kp=%lld must be >= k and a multiple of 128",
                  (long long)kp);
    AVTEX_REQUIRE(n < (int64_t(1) << 31), "pack_frames: n too large");
    return 0;
}

extern "C" int avtex_pack_frames_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld,
                                    int8_t *packed, int64_t kp, int64_t *sqnorm,
                                    unsigned long long *max_sqnorm, int device, void *stream) {
    AVTEX_ENTER(device);
    if (int rc = check_pack(n, k, ld, kp)) return rc;
    pack_u8_kernel<<<(unsigned)n, PACK_THREADS, 0, as_stream(stream)>>>(frames, k, ld, packed, kp, sqnorm, max_sqnorm);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_pack_frames_f32(const float *frames, int64_t n, int64_t k, int64_t ld,
                                     int8_t *packed, int64_t kp, int64_t *sqnorm, int *flags,
                                     unsigned long long *max_sqnorm, int device, void *stream) {
    AVTEX_ENTER(device);
    if (int rc = check_pack(n, k, ld, kp)) return rc;
    AVTEX_REQUIRE(flags != nullptr, "pack_frames_f32: flags must not be NULL");
    pack_f32_kernel<<<(unsigned)n, PACK_THREADS, 0, as_stream(stream)>>>(frames, k, ld, packed, kp, sqnorm, flags, max_sqnorm);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

static int launch_norms(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t row0, const NormDst &dst,
                        int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(n > 0 && k > 0 && ld >= k && n < (int64_t(1) << 31) && row0 >= 0,
                  "frame_norms_u8: bad shape n=%lld k=%lld ld=%lld", (long long)n, (long long)k, (long long)ld);
    int sms = 0, cc = 0;
    if (int rc = avtex_device_info(device, &sms, &cc)) return rc;
    const int64_t resident = int64_t(sms) * 8;              // 2048 threads per SM / 256
    const unsigned grid = (unsigned)(n < resident ? n : resident);
    frame_norms_u8_kernel<<<grid, PACK_THREADS, 0, as_stream(stream)>>>(frames, n, k, ld, row0, dst);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_frame_norms_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t *sqnorm,
                                    unsigned long long *max_centred, int device, void *stream) {
    NormDst dst;
    for (int d = 0; d < 8; ++d) { dst.sqnorm[d] = nullptr; dst.max_centred[d] = nullptr; }
    dst.sqnorm[0] = sqnorm; dst.max_centred[0] = max_centred; dst.count = 1;
    return launch_norms(frames, n, k, ld, 0, dst, device, stream);
}

extern "C" int avtex_frame_norms_u8_push(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t row0,
                                         int64_t *const *h_sqnorm, unsigned long long *const *h_max_centred,
                                         int num_dst, int device, void *stream) {
    AVTEX_REQUIRE(num_dst >= 1 && num_dst <= 8 && h_sqnorm != nullptr, "frame_norms_u8_push: 1..8 destinations (got %d)", num_dst);
    NormDst dst;
    for (int d = 0; d < 8; ++d) {
        dst.sqnorm[d] = d < num_dst ? h_sqnorm[d] : nullptr;
        dst.max_centred[d] = (d < num_dst && h_max_centred != nullptr) ? h_max_centred[d] : nullptr;
        AVTEX_REQUIRE(d >= num_dst || dst.sqnorm[d] != nullptr, "frame_norms_u8_push: destination %d is NULL", d);
    }
    dst.count = num_dst;
    return launch_norms(frames, n, k, ld, row0, dst, device, stream);
}

// ---------------------------------------------------------------- (f1) window construction
// out[w, t, :] = rows[idx[w * win + t], :]   (idx < 0: zero row — the reference zero-pads its chunks,
// cvt/utils/utils.py:252).  One CTA per output row; 128-bit copies when the row size and both bases allow it.
// HBM-bound gather: the frames of a window are contiguous in the clip, so consecutive CTAs read consecutive rows.
namespace {
__global__ void __launch_bounds__(PACK_THREADS)
gather_rows_kernel(const uint8_t *__restrict__ rows, int64_t row_bytes, int64_t pitch, int64_t n_rows,
                   const int *__restrict__ idx, uint8_t *__restrict__ out) {
    const int64_t o = blockIdx.x;
    const int src_row = idx[o];
    uint8_t *dst = out + o * row_bytes;
    const bool live = src_row >= 0 && src_row < n_rows;
    const uint8_t *src = rows + (live ? int64_t(src_row) : 0) * pitch;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const int64_t nv = vec ? (row_bytes & ~int64_t(15)) : 0;
    for (int64_t c = int64_t(threadIdx.x) * 16; c < nv; c += int64_t(PACK_THREADS) * 16) {
        const uint4 v = live ? __ldg(reinterpret_cast<const uint4 *>(src + c)) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(dst + c) = v;
    }
    for (int64_t c = nv + threadIdx.x; c < row_bytes; c += PACK_THREADS) dst[c] = live ? src[c] : uint8_t(0);
}
}  // namespace

extern "C" int avtex_gather_rows(const void *rows, int64_t row_bytes, int64_t pitch_bytes, int64_t n_rows,
                                 const int *idx, int64_t n_out, void *out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(row_bytes >= 1 && pitch_bytes >= row_bytes && n_rows >= 1 && n_out >= 1 && n_out < (int64_t(1) << 31) &&
                      rows != nullptr && idx != nullptr && out != nullptr,
                  "gather_rows: bad arguments row_bytes=%lld n_rows=%lld n_out=%lld", (long long)row_bytes,
                  (long long)n_rows, (long long)n_out);
    gather_rows_kernel<<<(unsigned)n_out, PACK_THREADS, 0, as_stream(stream)>>>(
        static_cast<const uint8_t *>(rows), row_bytes, pitch_bytes, n_rows, idx, static_cast<uint8_t *>(out));
    AVTEX_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- (f4) output frame assembly
// out[o] = video[ids[o]] with the reference's progress bar painted over rows [H-25, H-10): black, and red
// (255, 0, 0) in columns [mark_lo[o], mark_hi[o]) (empty range = no marker).  One CTA per output frame; the bar
// rows are written after the copy by the same CTA.  replaces: the per-frame host loop of
// classic/video_textures.py:215-226 and cvt/validate.py:622-634 (np.array(frames[idx]) + PIL per frame).
namespace {
__global__ void __launch_bounds__(PACK_THREADS)
assemble_frames_kernel(const uint8_t *__restrict__ video, int64_t n_frames, int h, int w, const int *__restrict__ ids,
                       const int *__restrict__ mark_lo, const int *__restrict__ mark_hi, int draw_bar,
                       uint8_t *__restrict__ out) {
    const int64_t o = blockIdx.x;
    const int64_t frame_bytes = int64_t(h) * w * 3;
    const int id = ids[o];
    const uint8_t *src = video + int64_t(id >= 0 && id < n_frames ? id : 0) * frame_bytes;
    uint8_t *dst = out + o * frame_bytes;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    const int64_t nv = vec ? (frame_bytes & ~int64_t(15)) : 0;
    for (int64_t c = int64_t(threadIdx.x) * 16; c < nv; c += int64_t(PACK_THREADS) * 16)
        *reinterpret_cast<uint4 *>(dst + c) = __ldg(reinterpret_cast<const uint4 *>(src + c));
    for (int64_t c = nv + threadIdx.x; c < frame_bytes; c += PACK_THREADS) dst[c] = src[c];
    if (!draw_bar || h < 25) return;
    __syncthreads();
    const int lo = mark_lo[o], hi = mark_hi[o];
    for (int i = threadIdx.x; i < 15 * w; i += PACK_THREADS) {
        const int r = h - 25 + i / w, c = i % w;
        uint8_t *px = dst + (int64_t(r) * w + c) * 3;
        px[0] = (c >= lo && c < hi) ? 255 : 0;
        px[1] = 0;
        px[2] = 0;
    }
}
}  // namespace

extern "C" int avtex_assemble_frames(const uint8_t *video, int64_t n_frames, int h, int w, const int *ids,
                                     const int *mark_lo, const int *mark_hi, int draw_bar, int64_t n_out,
                                     uint8_t *out, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(n_frames >= 1 && h >= 1 && w >= 1 && n_out >= 1 && n_out < (int64_t(1) << 31) && video != nullptr &&
                      ids != nullptr && out != nullptr && (!draw_bar || (mark_lo != nullptr && mark_hi != nullptr)),
                  "assemble_frames: bad arguments");
    assemble_frames_kernel<<<(unsigned)n_out, PACK_THREADS, 0, as_stream(stream)>>>(video, n_frames, h, w, ids, mark_lo,
                                                                                   mark_hi, draw_bar, out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

"""Host-side engine: thin, allocation-explicit wrappers over the C ABI (include/avtex.h).

Everything here takes / returns CUDA tensors and launches on the current torch stream; torch is
used for device memory and streams only.  Row-range arguments (`row0`, `rows`, `a0`, ...) exist so
the same calls serve the row-sharded multi-GPU path (dist.py).  No function here falls back to
PyTorch arithmetic: a missing library or a failed kernel raises.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

F32_EPS_STOP = 10e-3            # classic/q_learning.py:39  `while eps > 10e-3`


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(x) -> np.float32:
    """Python float / 0-dim tensor -> fp32 exactly as an ATen scalar operand is cast."""
    if isinstance(x, torch.Tensor):
        return np.float32(x.detach().cpu().item())
    return np.float32(x)


def empty_matrix(rows: int, cols: int, device, dtype=torch.float32) -> torch.Tensor:
    """[rows, cols] view of a buffer whose leading dimension is padded to 128 bytes, so every row
    starts 16-byte aligned and the kernels' 128-bit paths apply to all rows (an odd M such as 1241
    or 19961 would otherwise leave 3 of 4 rows on the scalar path)."""
    per = 128 // torch.empty(0, dtype=dtype).element_size()
    ld = (cols + per - 1) // per * per
    return torch.empty((rows, ld), dtype=dtype, device=device)[:, :cols]


# --------------------------------------------------------------------------- stats / sigma
def new_stats(device) -> torch.Tensor:
    """16-byte accumulator: [fp64 sum | uint64 nnz]."""
    return torch.zeros(2, dtype=torch.float64, device=device)


def _stats_ptrs(stats):
    if stats is None:
        return None, None
    base = stats.data_ptr()
    return C.c_void_p(base), C.c_void_p(base + 8)


def read_stats(stats: torch.Tensor):
    """(sum: float, nnz: int) — one D2H copy + sync."""
    h = stats.cpu()
    return float(h[0]), int(h.view(torch.int64)[1])


def sigma_from_stats(total: float, nnz: int, sigma_factor) -> np.float32:
    """sigma = f * (sum(D) / nnz) in fp32 (classic/computeD1.py:240-241): the fp32 sum divided by
    the count cast to fp32, then multiplied by fp32(f)."""
    return _f32(sigma_factor) * (np.float32(total) / np.float32(nnz))


def sum_nnz(D: torch.Tensor, stats: torch.Tensor | None = None) -> torch.Tensor:
    stats = new_stats(D.device) if stats is None else stats
    s, z = _stats_ptrs(stats)
    _lib.call("avtex_sum_nnz", _lib.ptr(D), D.shape[0], D.shape[1], D.stride(0), s, z, _dev(D), _stream(D))
    return stats


# --------------------------------------------------------------------------- K0 / K1
GRAM_MAX_SQNORM = (1 << 32) // 4  # (sqrt(n_r)+sqrt(n_c))^2 <= 4 max(n) must stay below 2^32


@dataclass
class PackedFrames:
    """Gram operand: either the raw uint8 frames themselves (`signed` False, no copy) or the centred
    int8 copy written by K0 (`signed` True; float inputs, or byte rows that are not 16-byte aligned)."""
    packed: torch.Tensor          # [N, K] uint8 (raw) or [N, Kp] int8 (centred, zero padded)
    sqnorm: torch.Tensor          # [N] int64: sum x^2 of the operand rows
    k: int
    flags: torch.Tensor           # int64[2] on device: [non-byte values seen, max centred sqnorm]
    signed: bool = True
    _checked: tuple | None = None
    norms_pending: bool = False   # raw uint8 frames whose norms the NEXT Gram launch computes itself (pack_frames(defer_norms=True))

    def validate(self):
        """(ok, reason): is the tensor-core Gram path exact for these frames?  One 16-byte D2H read,
        deferred so that avtex_gram_l2_s8 can already be running (it is launched speculatively; its
        result must be discarded when this returns False)."""
        if self._checked is None:
            bad, mx = (int(v) for v in self.flags.cpu())
            if bad:
                self._checked = (False, "frames are not integer-valued bytes")
            elif mx >= GRAM_MAX_SQNORM:
                self._checked = (False, "squared norms too large for the mod-2^32 epilogue")
            else:
                self._checked = (True, "")
        return self._checked

    @property
    def exact_ok(self) -> bool:
        return self.validate()[0]

    @property
    def reason(self) -> str:
        return self.validate()[1]


def pack_frames(frames: torch.Tensor, defer_norms: bool = False) -> PackedFrames:
    """K0.  frames: CUDA tensor [N, ...] uint8 or float (integer-valued 0..255).  No host sync.
    defer_norms: for raw uint8 frames launch nothing now — the next gram_l2 / gram_l2_residues / gram_l2_jobs on the
    result computes the norms inside the Gram launch (its epilogue warps are idle until the first tiles are done)."""
    x = frames.reshape(frames.shape[0], -1)
    if x.stride(-1) != 1:
        x = x.contiguous()
    n, k = x.shape
    kp = (k + 127) // 128 * 128
    sqnorm = torch.empty(n, dtype=torch.int64, device=x.device)
    flags = torch.zeros(2, dtype=torch.int64, device=x.device)
    dev, st = _dev(x), _stream(x)
    mx = C.c_void_p(flags.data_ptr() + 8)
    if x.dtype == torch.uint8 and x.stride(0) % 16 == 0 and x.data_ptr() % 16 == 0:
        # raw bytes feed the tensor cores directly: only the norms are computed (one read of the frames)
        if not defer_norms:
            _lib.call("avtex_frame_norms_u8", _lib.ptr(x), n, k, x.stride(0), _lib.ptr(sqnorm), mx, dev, st)
        pf = PackedFrames(x, sqnorm, k, flags, signed=False, norms_pending=bool(defer_norms))
        if kp * 128 * 128 < GRAM_MAX_SQNORM:
            pf._checked = (True, "")
        return pf
    packed = torch.empty((n, kp), dtype=torch.int8, device=x.device)
    if x.dtype == torch.uint8:
        _lib.call("avtex_pack_frames_u8", _lib.ptr(x), n, k, x.stride(0), _lib.ptr(packed), kp,
                  _lib.ptr(sqnorm), mx, dev, st)
    elif x.dtype == torch.float32:
        _lib.call("avtex_pack_frames_f32", _lib.ptr(x), n, k, x.stride(0), _lib.ptr(packed), kp,
                  _lib.ptr(sqnorm), _lib.ptr(flags), mx, dev, st)
    else:
        raise TypeError(f"frames dtype {x.dtype} not supported (uint8 or float32)")
    pf = PackedFrames(packed, sqnorm, k, flags, signed=True)
    if kp * 128 * 128 < GRAM_MAX_SQNORM and x.dtype == torch.uint8:
        pf._checked = (True, "")                           # no byte frame of this size can leave the domain
    return pf


def frame_norms_rows(frames_u8: torch.Tensor, row0: int, rows: int, sqnorm: torch.Tensor, flags: torch.Tensor):
    """K0 on a row range of raw uint8 frames [N, K] (row pitch multiple of 16): fills sqnorm[row0:row0+rows]
    and raises flags[1] to the maximum centred norm.  Used by the sharded path, where each rank computes only
    its own slice of the norms."""
    part = frames_u8[row0:row0 + rows]
    _lib.call("avtex_frame_norms_u8", _lib.ptr(part), rows, frames_u8.shape[1], frames_u8.stride(0),
              C.c_void_p(sqnorm.data_ptr() + 8 * row0), C.c_void_p(flags.data_ptr() + 8), _dev(frames_u8),
              _stream(frames_u8))


def frame_norms_push(frames_u8: torch.Tensor, row0: int, rows: int, sqnorm_ptrs: list, max_ptrs: list | None):
    """K0 on rows [row0, row0+rows) of raw uint8 frames, results written to EVERY destination vector in
    `sqnorm_ptrs` (raw device pointers of the full [N] int64 vector on each GPU, peer-mapped) and the centred
    maximum raised in `max_ptrs`: the row-sharded path pushes its slice of the norms to all peers from the
    kernel instead of all-gathering afterwards."""
    part = frames_u8[row0:row0 + rows]
    n_dst = len(sqnorm_ptrs)
    sq = (C.c_void_p * n_dst)(*sqnorm_ptrs)
    mx = (C.c_void_p * n_dst)(*max_ptrs) if max_ptrs is not None else None
    _lib.call("avtex_frame_norms_u8_push", _lib.ptr(part), rows, frames_u8.shape[1], frames_u8.stride(0), row0,
              sq, mx, n_dst, _dev(frames_u8), _stream(frames_u8))


def gram_l2(pf: PackedFrames, row0: int = 0, rows: int | None = None, symmetric: bool | None = None,
            stats: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """K1 on tensor cores.  Returns D[rows, N] for global rows [row0, row0+rows)."""
    n, kp = pf.packed.shape
    rows = n - row0 if rows is None else rows
    if symmetric is None:
        symmetric = (row0 == 0 and rows == n)
    D = empty_matrix(rows, n, pf.packed.device) if out is None else out
    if pf.norms_pending:                                   # norms computed inside this launch (job-list entry point)
        gram_l2_jobs(pf, [dict(row0=row0, rows=rows, col0=0, cols=n, symmetric=1 if symmetric else 0, count_stats=1,
                               D=D.data_ptr(), d_row0=row0, ldd=D.stride(0),
                               DT=D.data_ptr() if symmetric else None, dt_row0=0, ldt=D.stride(0))], stats)
        if symmetric:
            mark_symmetric(D)
        return D
    s, z = _stats_ptrs(stats)
    if pf.signed:
        _lib.call("avtex_gram_l2_s8", _lib.ptr(pf.packed), n, kp, _lib.ptr(pf.sqnorm), row0, rows,
                  1 if symmetric else 0, _lib.ptr(D), D.stride(0), s, z, _dev(D), _stream(D))
    else:
        _lib.call("avtex_gram_l2_u8", _lib.ptr(pf.packed), n, pf.k, pf.packed.stride(0), _lib.ptr(pf.sqnorm),
                  row0, rows, 1 if symmetric else 0, _lib.ptr(D), D.stride(0), s, z, _dev(D), _stream(D))
    if symmetric:
        mark_symmetric(D)
    return D


def gram_job_array(jobs: list):
    """ctypes array of AvtexGramJob from a list of dicts (fields of the struct; D / DT raw device pointers).  Job lists
    that do not change between calls (the sharded step's) are built once and passed to gram_l2_jobs as is."""
    arr = (_lib.GramJob * len(jobs))()
    for dst, j in zip(arr, jobs):
        dst.row0, dst.rows, dst.col0, dst.cols = j["row0"], j["rows"], j["col0"], j["cols"]
        dst.D, dst.d_row0, dst.ldd = j.get("D"), j.get("d_row0", 0), j.get("ldd", 0)
        dst.DT, dst.dt_row0, dst.ldt = j.get("DT"), j.get("dt_row0", 0), j.get("ldt", 0)
        dst.symmetric, dst.count_stats = int(j.get("symmetric", 0)), int(j.get("count_stats", 0))
        dst.k_off, dst.sq_off, dst.sq_stride = j.get("k_off", 0), j.get("sq_off", 0), j.get("sq_stride", 1)
    return arr


def gram_l2_jobs(pf: PackedFrames, jobs, stats: torch.Tensor | None = None, device=None,
                 clock_probe: torch.Tensor | None = None, n: int | None = None, ld: int | None = None):
    """K1, general form: `jobs` is a list of dicts with the fields of AvtexGramJob (D / DT are raw device
    pointers, possibly into a peer GPU's symmetric-memory buffer).  `clock_probe`: int64[2] on the device,
    receives (SM cycles, nanoseconds) of CTA 0's tile loop (bench.py's in-kernel clock measurement)."""
    arr = jobs if isinstance(jobs, C.Array) else gram_job_array(jobs)      # a prebuilt array skips ~2 us per field
    frames_n = pf.packed.shape[0]
    n = frames_n if n is None else n                       # n / ld overridden by the residue-class view [N/s, s*K]
    s, z = _stats_ptrs(stats)
    k_extent = pf.packed.shape[1] if pf.signed else pf.k
    if pf.norms_pending:
        if clock_probe is not None:
            raise ValueError("clock_probe and deferred norms cannot be combined")
        sync = torch.zeros(1, dtype=torch.int32, device=pf.packed.device)
        _lib.call("avtex_gram_l2_jobs_fused_norms", _lib.ptr(pf.packed), n, k_extent,
                  pf.packed.stride(0) if ld is None else ld, frames_n, pf.packed.stride(0), _lib.ptr(pf.sqnorm),
                  C.c_void_p(pf.flags.data_ptr() + 8), _lib.ptr(sync), arr, len(jobs), s, z, _dev(pf.packed),
                  _stream(pf.packed))
        pf.norms_pending = False
        return
    _lib.call("avtex_gram_l2_jobs", _lib.ptr(pf.packed), 1 if pf.signed else 0, n, k_extent,
              pf.packed.stride(0) if ld is None else ld,
              _lib.ptr(pf.sqnorm), arr, len(jobs), s, z, _lib.ptr(clock_probe), _dev(pf.packed), _stream(pf.packed))


def pairwise_l2_from_host(frames: torch.Tensor, device=None, stats: torch.Tensor | None = None, chunks: int = 8):
    """D1 for frames that live in HOST memory (uint8, or float32 as the reference's own call hands them over,
    classic/video_textures.py:245,266): the H2D copy is cut into row chunks on a copy stream and, as each chunk
    lands, its norms and its part of the symmetric Gram are computed on the main stream (job list: the chunk's
    diagonal block + the rectangle against all earlier rows, stored direct and transposed).  The tensor cores
    work underneath the PCIe transfer, so the end-to-end time is the copy time plus the last chunk's tiles.
    float32 frames travel as they are (4 bytes per value: the copy is 4x longer) through two staging buffers and
    are packed to centred int8 per chunk on the device; whether they were integer-valued bytes is known when the
    last chunk has been packed (PackedFrames.exact_ok).  Returns (D1, PackedFrames) or None when the frames are
    not eligible (then use pairwise_l2)."""
    if frames.is_cuda or frames.dtype not in (torch.uint8, torch.float32):
        return None
    x = frames.reshape(frames.shape[0], -1)
    n, k = x.shape
    is_f32 = frames.dtype == torch.float32
    if not x.is_contiguous() or n < 512 or (k % 16 != 0 and not is_f32):
        return None
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    kp = (k + 127) // 128 * 128
    step = max(256, (-(-n // chunks) + 255) // 256 * 256)
    bounds = [(c0, min(n, c0 + step)) for c0 in range(0, n, step)]
    sqnorm = torch.empty(n, dtype=torch.int64, device=dev)
    flags = torch.zeros(2, dtype=torch.int64, device=dev)
    if is_f32:
        buf = torch.empty((n, kp), dtype=torch.int8, device=dev)                 # the packed operand
        staging = [torch.empty((step, k), dtype=torch.float32, device=dev) for _ in range(2)]
        pf = PackedFrames(buf, sqnorm, k, flags, signed=True)
    else:
        buf = torch.empty((n, k), dtype=torch.uint8, device=dev)
        pf = PackedFrames(buf, sqnorm, k, flags, signed=False)
    D1 = empty_matrix(n, n, dev)
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    copy_stream.wait_stream(main)
    mx = C.c_void_p(flags.data_ptr() + 8)
    d_ptr, ld = D1.data_ptr(), D1.stride(0)
    consumed = []                                               # float path: "staging buffer i has been packed"
    for i, (c0, c1) in enumerate(bounds):
        with torch.cuda.stream(copy_stream):
            if is_f32:
                if i >= 2:
                    copy_stream.wait_event(consumed[i - 2])
                staging[i % 2][:c1 - c0].copy_(x[c0:c1], non_blocking=True)
            else:
                buf[c0:c1].copy_(x[c0:c1], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(copy_stream)
        main.wait_event(landed)
        if is_f32:
            part = staging[i % 2][:c1 - c0]
            _lib.call("avtex_pack_frames_f32", _lib.ptr(part), c1 - c0, k, part.stride(0),
                      C.c_void_p(buf.data_ptr() + c0 * kp), kp, C.c_void_p(sqnorm.data_ptr() + 8 * c0),
                      _lib.ptr(flags), mx, _dev(buf), _stream(buf))
            ev = torch.cuda.Event()
            ev.record(main)
            consumed.append(ev)
        else:
            part = buf[c0:c1]
            _lib.call("avtex_frame_norms_u8", _lib.ptr(part), c1 - c0, k, buf.stride(0),
                      C.c_void_p(sqnorm.data_ptr() + 8 * c0), mx, _dev(buf), _stream(buf))
        jobs = [dict(row0=c0, rows=c1 - c0, col0=c0, cols=c1 - c0, symmetric=1, count_stats=1,
                     D=d_ptr, d_row0=0, ldd=ld, DT=d_ptr, dt_row0=0, ldt=ld)]
        if c0 > 0:
            jobs.append(dict(row0=0, rows=c0, col0=c0, cols=c1 - c0, symmetric=0, count_stats=1,
                             D=d_ptr, d_row0=0, ldd=ld, DT=d_ptr, dt_row0=0, ldt=ld))
        gram_l2_jobs(pf, jobs, stats)
    buf.record_stream(copy_stream)
    if is_f32:
        for t in staging:
            t.record_stream(copy_stream)
    elif kp * 128 * 128 < GRAM_MAX_SQNORM:
        pf._checked = (True, "")
    mark_symmetric(D1)
    return D1, pf


def pairdist_direct(frames: torch.Tensor, row0: int = 0, rows: int | None = None,
                    stats: torch.Tensor | None = None) -> torch.Tensor:
    """K1 fallback: direct difference in fp32 (any float features, or uint8)."""
    x = frames.reshape(frames.shape[0], -1)
    if x.stride(-1) != 1:
        x = x.contiguous()
    n, k = x.shape
    rows = n - row0 if rows is None else rows
    D = empty_matrix(rows, n, x.device)
    s, z = _stats_ptrs(stats)
    name = {torch.float32: "avtex_pairdist_direct_f32", torch.uint8: "avtex_pairdist_direct_u8"}.get(x.dtype)
    if name is None:
        raise TypeError(f"frames dtype {x.dtype} not supported (uint8 or float32)")
    _lib.call(name, _lib.ptr(x), n, k, x.stride(0), row0, rows, _lib.ptr(D), D.stride(0), s, z, _dev(D), _stream(D))
    return D


def pairwise_l2(frames: torch.Tensor, row0: int = 0, rows: int | None = None,
                stats: torch.Tensor | None = None, method: str = "auto"):
    """D1 rows [row0, row0+rows).  method: auto | gram | direct.  Returns (D, method_used)."""
    if method not in ("auto", "gram", "direct"):
        raise ValueError(method)
    if method != "direct" and frames.dtype in (torch.uint8, torch.float32):
        pf = pack_frames(frames)
        D = gram_l2(pf, row0, rows, stats=stats)           # speculative: validated right below
        if pf.exact_ok:
            return D, "gram"
        if method == "gram":
            raise _lib.AvtexError(f"gram path not applicable: {pf.reason}")
        if stats is not None:
            stats.zero_()
    return pairdist_direct(frames if frames.dtype == torch.uint8 else frames.float(), row0, rows, stats), "direct"


# --------------------------------------------------------------------------- K5
def transition_probs(D: torch.Tensor, sigma, shift: int = 1, rows_out: int | None = None,
                     threshold: float | None = None, want_P: bool = True, want_counts: bool = False,
                     Pn_out: torch.Tensor | None = None):
    """P (and thresholded P_new) for output rows [0, rows_out) from source rows min(i+shift, rows-1).
    `Pn_out`: caller-provided [rows_out, cols] destination of P_new (e.g. a peer-mapped symmetric buffer)."""
    rows_in, cols = D.shape
    rows_out = rows_in if rows_out is None else rows_out
    P = empty_matrix(rows_out, cols, D.device) if want_P else None
    Pn = None
    if threshold is not None:
        Pn = empty_matrix(rows_out, cols, D.device) if Pn_out is None else Pn_out[:rows_out, :cols]
    counts = torch.empty(rows_out, dtype=torch.int32, device=D.device) if want_counts else None
    th = _f32(threshold) if threshold is not None else np.float32(-1.0)
    _lib.call("avtex_transition_probs", _lib.ptr(D), D.stride(0), rows_in, cols, C.c_float(_f32(sigma)), shift,
              rows_out, _lib.ptr(P), P.stride(0) if P is not None else 0, C.c_float(th), _lib.ptr(Pn),
              Pn.stride(0) if Pn is not None else 0, _lib.ptr(counts), _dev(D), _stream(D))
    return P, Pn, counts


_PINNED: dict = {"ring": [None] * 2, "next": 0}


def _pinned(nbytes: int) -> torch.Tensor:
    """Page-locked staging memory of at least nbytes from a ring of 2 cached buffers (grown geometrically, both sized
    on the first call — a page-locked allocation costs ~6 ms): results exported through it stay valid until the
    second following export."""
    ring = _PINNED["ring"]
    i = _PINNED["next"]
    _PINNED["next"] = (i + 1) % len(ring)
    if ring[i] is None or ring[i].numel() < nbytes:
        size = max(nbytes, 2 * max((b.numel() for b in ring if b is not None), default=0), 1 << 20)
        for j in range(len(ring)):
            if ring[j] is None or ring[j].numel() < nbytes:
                ring[j] = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    return ring[i]


def csr_from_matrix(P: torch.Tensor, counts: torch.Tensor | None = None):
    """Ascending non-zero columns per row -> (rowptr int64 numpy, colidx int32 numpy).
    Matrices up to 16 MB (M <= 2048) are compacted into a full-capacity index buffer and brought to the host with
    ONE asynchronous copy into page-locked memory and one sync; the arrays returned are VIEWS of that page-locked
    buffer (valid until the second following export — copy them to keep them longer).  Round 1 sized the list with an
    `.item()`, copied it with a pageable `.cpu()` and again into numpy: 1.5 ms at C2, more than the whole device
    pass; a fresh 3 MB numpy copy alone costs 1 ms (page faults).  Larger matrices size the list first."""
    rows, cols = P.shape
    dev = P.device
    if counts is None:
        counts = torch.empty(rows, dtype=torch.int32, device=dev)
        _lib.call("avtex_row_nnz", _lib.ptr(P), P.stride(0), rows, cols, _lib.ptr(counts), _dev(P), _stream(P))
    cap = rows * cols
    small = cap * 4 <= (16 << 20)
    head = (rows + 1) * 8
    if small:
        # [rowptr int64 (rows + 1) | colidx int32 (capacity)] in one device buffer
        both = torch.empty(head + cap * 4, dtype=torch.uint8, device=dev)
        rowptr = both[:head].view(torch.int64)
        colidx = both[head:].view(torch.int32)
        rowptr[0] = 0
        torch.cumsum(counts, 0, out=rowptr[1:])
        _lib.call("avtex_csr_fill", _lib.ptr(P), P.stride(0), rows, cols, _lib.ptr(rowptr), _lib.ptr(colidx),
                  _dev(P), _stream(P))
        host = _pinned(both.numel())[:both.numel()]
        host.copy_(both, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        rp = host[:head].view(torch.int64).numpy()
        return rp, host[head:head + int(rp[-1]) * 4].view(torch.int32).numpy()
    rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=rowptr[1:])
    total = int(rowptr[-1].item())
    colidx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    _lib.call("avtex_csr_fill", _lib.ptr(P), P.stride(0), rows, cols, _lib.ptr(rowptr), _lib.ptr(colidx),
              _dev(P), _stream(P))
    host = _pinned(head + total * 4)
    host[:head].copy_(rowptr.view(torch.uint8), non_blocking=True)
    host[head:head + total * 4].copy_(colidx[:total].view(torch.uint8), non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host[:head].view(torch.int64).numpy(), host[head:head + total * 4].view(torch.int32).numpy()


class SurvivorRows:
    """The survivor lists (ascending non-zero columns per row of P3_new) the sampling walk draws from, fetched ON
    DEMAND: `rows[i]` compacts row i on the device and copies only that list to the host (cached).  The reference
    does the same per step (`P[this_frame].nonzero().cpu()`, classic/video_textures.py:76-78); at N = 100000 about
    half of the 6.2e8 entries survive the default threshold, so copying every list (1.2 GB) for a 900-frame walk
    would dominate the whole pipeline.  `row_fn(i)` returns a CUDA fp32 view of row i (local shard or a peer-mapped
    row of another GPU's shard)."""

    def __init__(self, n_rows: int, row_fn):
        self.n_rows, self._row_fn, self._cache, self.fetched = n_rows, row_fn, {}, 0

    def __len__(self):
        return self.n_rows

    def __getitem__(self, i: int) -> np.ndarray:
        i = int(i)
        got = self._cache.get(i)
        if got is None:
            row = self._row_fn(i)
            got = torch.nonzero(row).view(-1).to(torch.int32).cpu().numpy()
            self._cache[i] = got
            self.fetched += 1
        return got

    @classmethod
    def from_matrix(cls, P: torch.Tensor) -> "SurvivorRows":
        return cls(P.shape[0], lambda i: P[i])


# --------------------------------------------------------------------------- K2
@functools.lru_cache(maxsize=64)
def _binomial_taps_cached(filter_size: int) -> np.ndarray:
    taps = np.asarray((np.poly1d([0.5, 0.5]) ** (filter_size - 1)).coeffs, dtype=np.float32).reshape(-1)
    taps.setflags(write=False)
    return taps


def binomial_taps(filter_size: int) -> np.ndarray:
    """classic/computeD2.py:34 — coeffs((0.5 x + 0.5)^(fs-1)) evaluated in float64, cast to fp32.  Cached: the
    polynomial power costs ~0.6 ms of host time at fs = 40, which sat in front of every filter launch."""
    return _binomial_taps_cached(int(filter_size))


def filtered_size(n: int, filter_size: int, stride: int) -> int:
    return (n - filter_size) // stride + 1


# Distance matrices this module produced itself (symmetric by construction: the Gram epilogue stores D[r,c] and
# D[c,r] from the same register) are remembered by object identity + torch's in-place version counter, so that K2
# may use its symmetric form without the caller promising anything.  A matrix that was modified in place, copied,
# re-sliced or supplied by the user is simply not found and takes the general kernel.
_SYMMETRIC: list = []


def mark_symmetric(t: torch.Tensor) -> torch.Tensor:
    import weakref
    _SYMMETRIC[:] = [e for e in _SYMMETRIC[-15:] if e[0]() is not None]
    _SYMMETRIC.append((weakref.ref(t), t._version, t.data_ptr(), tuple(t.shape), t.stride()))
    return t


def known_symmetric(t: torch.Tensor) -> bool:
    for ref, ver, ptr, shape, stride in _SYMMETRIC:
        if ref() is t and t._version == ver and t.data_ptr() == ptr and tuple(t.shape) == shape and t.stride() == stride:
            return True
    return False


def diag_filter(D1: torch.Tensor, filter_size: int, stride: int = 1, p: float | None = None,
                m: int | None = None, a0: int = 0, rows_out: int | None = None, in_row0: int = 0,
                stats: torch.Tensor | None = None, taps: np.ndarray | None = None, symmetric: bool | None = None):
    """K2.  D1 holds global rows [in_row0, in_row0 + D1.shape[0]) and all columns.
    Returns (D2[rows_out, m], D3 | None).
    symmetric: D1 == D1.T bit for bit (the whole square matrix): only the upper triangle is computed and mirrored
    (half the D1 bytes, FMAs and pows).  None = yes iff D1 is a distance matrix this module produced
    (`known_symmetric`); row shards and caller-supplied matrices take the general kernel."""
    n_cols = D1.shape[1]
    m = filtered_size(n_cols, filter_size, stride) if m is None else m
    rows_out = m - a0 if rows_out is None else rows_out
    taps = binomial_taps(filter_size) if taps is None else np.ascontiguousarray(taps, dtype=np.float32)
    if taps.shape[0] != filter_size:
        raise ValueError("taps length != filter_size")
    whole = (a0 == 0 and rows_out == m and in_row0 == 0 and D1.shape[0] == n_cols)
    if symmetric is None:
        symmetric = whole and known_symmetric(D1)
    elif symmetric and not whole:
        raise ValueError("symmetric=True needs the whole square matrix")
    D2 = empty_matrix(rows_out, m, D1.device)
    D3 = empty_matrix(rows_out, m, D1.device) if p is not None else None
    s, z = _stats_ptrs(stats)
    tp = taps.ctypes.data_as(C.POINTER(C.c_float))
    pf = C.c_float(_f32(p if p is not None else 1.0))
    if symmetric:
        _lib.call("avtex_diag_filter_pow_sym", _lib.ptr(D1), D1.stride(0), D1.shape[0], tp, filter_size, stride, m,
                  _lib.ptr(D2), D2.stride(0), _lib.ptr(D3), D3.stride(0) if D3 is not None else 0, pf, s, z,
                  _dev(D1), _stream(D1))
        mark_symmetric(D2)
    else:
        _lib.call("avtex_diag_filter_pow", _lib.ptr(D1), D1.stride(0), in_row0, D1.shape[0], tp, filter_size, stride,
                  a0, rows_out, m, _lib.ptr(D2), D2.stride(0), _lib.ptr(D3), D3.stride(0) if D3 is not None else 0,
                  pf, s, z, _dev(D1), _stream(D1))
    return D2, D3


# --------------------------------------------------------------------------- stride-s pipelines: residue classes
# A stride-s filter reads D1[i, j] only where i = j (mod s):  D2[a,b] = sum_k w[k] D1[s*a + k, s*b + k].  Those
# entries are exactly the s Gram matrices of the frames of one residue class each (N/s x N/s), i.e. 1/s of the
# distance matrix and 1/s of the tensor-core work (1/4 at the reference's default -stride 4).  When the caller
# wants D2 / D3 and not D1 itself (the classic++ pipeline after compute_D1's own outputs: video_textures.py:265-284
# uses P3_new only), K1 computes the s class matrices in ONE launch (job list over the clip viewed as [N/s, s*K]) and
# K2 walks the planes round-robin in the same tap order, so D2 and D3 are bit-identical to the full-D1 path.
RESIDUE_FAST = {(40, 4), (16, 4)}          # (fs, stride) pairs with a register-resident K2 instantiation


def residue_eligible(pf: PackedFrames, filter_size: int, stride: int) -> bool:
    n = pf.packed.shape[0]
    kb = pf.packed.shape[1] if pf.signed else pf.k
    return ((filter_size, stride) in RESIDUE_FAST and n % stride == 0 and n // stride >= 1 and kb % 128 == 0
            and pf.packed.stride(0) == kb and pf.packed.stride(1) == 1)


def gram_l2_residues(pf: PackedFrames, stride: int) -> torch.Tensor:
    """K1 on the residue classes: returns D1r [stride, N/stride, ld] fp32 with
    D1r[r, a, b] = d(frame stride*a + r, frame stride*b + r), every plane symmetric.  One launch."""
    n = pf.packed.shape[0]
    nc = n // stride
    kb = pf.packed.shape[1] if pf.signed else pf.k
    ldc = (nc + 31) // 32 * 32
    D1r = torch.empty((stride, nc, ldc), dtype=torch.float32, device=pf.packed.device)
    jobs = []
    for r in range(stride):
        ptr = D1r.data_ptr() + r * nc * ldc * 4
        jobs.append(dict(row0=0, rows=nc, col0=0, cols=nc, symmetric=1, count_stats=0, D=ptr, d_row0=0, ldd=ldc,
                         DT=ptr, dt_row0=0, ldt=ldc, k_off=r * kb, sq_off=r, sq_stride=stride))
    gram_l2_jobs(pf, jobs, n=nc, ld=stride * kb)
    return D1r


def diag_filter_residues(D1r: torch.Tensor, n_frames: int, filter_size: int, stride: int, p: float | None = None,
                         stats: torch.Tensor | None = None, taps: np.ndarray | None = None, symmetric: bool | None = None,
                         a0: int = 0, rows_out: int | None = None, in_row0: int = 0):
    """K2 on the residue-class planes of `gram_l2_residues`.  Returns (D2 [rows_out, M], D3 | None), bit-identical to
    diag_filter on the full matrix.  Row shards: the planes hold class rows [in_row0, in_row0 + D1r.shape[1])."""
    m = filtered_size(n_frames, filter_size, stride)
    rows_out = m - a0 if rows_out is None else rows_out
    whole = a0 == 0 and rows_out == m and in_row0 == 0
    symmetric = whole if symmetric is None else symmetric
    if symmetric and not whole:
        raise ValueError("symmetric=True needs the whole matrix")
    taps = binomial_taps(filter_size) if taps is None else np.ascontiguousarray(taps, dtype=np.float32)
    D2 = empty_matrix(rows_out, m, D1r.device)
    D3 = empty_matrix(rows_out, m, D1r.device) if p is not None else None
    s, z = _stats_ptrs(stats)
    _lib.call("avtex_diag_filter_pow_res", _lib.ptr(D1r), D1r.stride(1), D1r.stride(0), in_row0, D1r.shape[1],
              taps.ctypes.data_as(C.POINTER(C.c_float)), filter_size, stride, a0, rows_out, m, _lib.ptr(D2), D2.stride(0),
              _lib.ptr(D3), D3.stride(0) if D3 is not None else 0, C.c_float(_f32(p if p is not None else 1.0)), s, z,
              1 if symmetric else 0, _dev(D1r), _stream(D1r))
    if symmetric:
        mark_symmetric(D2)
    return D2, D3


def distance_filter(frames: torch.Tensor, filter_size: int, stride: int, p: float | None = 0.7,
                    stats: torch.Tensor | None = None, allow_residues: bool = True):
    """frames -> (D2, D3, how) without handing out D1: K0 + K1 + K2 with the residue-class shortcut when the stride,
    the clip and the filter allow it (`how` = "residues"), else the full distance matrix ("gram" / "direct")."""
    if allow_residues and stride >= 2 and frames.dtype in (torch.uint8, torch.float32):
        pf = pack_frames(frames, defer_norms=True)          # raw uint8: the norms are computed inside the Gram launch
        if not residue_eligible(pf, filter_size, stride):
            pf = None                                       # (a deferred raw-uint8 pack has launched nothing)
        if pf is not None:
            D1r = gram_l2_residues(pf, stride)              # speculative: the exactness domain is checked right below
            if pf.exact_ok:
                D2, D3 = diag_filter_residues(D1r, frames.shape[0], filter_size, stride, p=p, stats=stats)
                return D2, D3, "residues"
    D1, how = pairwise_l2(frames if frames.dtype in (torch.uint8, torch.float32) else frames.float())
    D2, D3 = diag_filter(D1, filter_size, stride, p=p, stats=stats)
    return D2, D3, how


class PipelineGraph:
    """frames -> (D2, D3, converged future cost, D3_new) as ONE CUDA-graph launch, for a fixed clip shape.

    The classic++ pass is 5-6 kernels of which all but the Gram take 10-100 us: launched one by one from Python the
    GPU idles between them (~0.06-0.1 ms per pass at C2, 4 % of the full pass and 11 % of the residue-class pass).
    Every launch of libavtex takes its stream explicitly and none synchronises, so the whole pass is capturable:
    `PipelineGraph(n, k, fs, stride)` captures it once (after one eager warm-up pass that loads the kernels and sets
    their attributes), `graph(frames)` copies the clip into the static input buffer and replays.  The outputs are
    static tensors, overwritten by every replay.  `how` = "residues" (stride-s residue-class pipeline) or "gram"."""

    def __init__(self, n: int, k: int, filter_size: int, stride: int, p: float = 0.7, alpha: float = 0.997,
                 device="cuda", residues: bool | None = None, want_D1: bool = False):
        self.fs, self.stride, self.p, self.alpha = filter_size, stride, p, alpha
        self.frames = torch.zeros((n, k), dtype=torch.uint8, device=device)
        pf = pack_frames(self.frames)
        if pf.signed:
            raise _lib.AvtexError("PipelineGraph needs 16-byte aligned uint8 rows (K % 16 == 0)")
        eligible = stride >= 2 and residue_eligible(pf, filter_size, stride) and not want_D1
        self.how = "residues" if (eligible if residues is None else residues and eligible) else "gram"
        side = torch.cuda.Stream(self.frames.device)
        side.wait_stream(torch.cuda.current_stream(self.frames.device))
        with torch.cuda.stream(side):                      # eager warm-up: module load, function attributes
            self._run()
        torch.cuda.current_stream(self.frames.device).wait_stream(side)
        torch.cuda.synchronize(self.frames.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()
        self._fc_static = self.fc._pending                 # (info, eps trail, m, None) of the captured launch
        self._result = None

    def _run(self):
        pf = pack_frames(self.frames, defer_norms=True)    # K0 runs inside the Gram launch
        self.pf = pf
        if self.how == "residues":
            self.D1 = gram_l2_residues(pf, self.stride)
            self.D2, self.D3 = diag_filter_residues(self.D1, self.frames.shape[0], self.fs, self.stride, p=self.p)
        else:
            self.D1 = gram_l2(pf)
            self.D2, self.D3 = diag_filter(self.D1, self.fs, self.stride, p=self.p, symmetric=True)
        self.fc = future_cost_fused(self.D3, self.alpha)
        self.stats = new_stats(self.frames.device)
        self.D3_new = future_cost_finalize(self.D3, self.fc.mvec, self.alpha, stats=self.stats)

    def __call__(self, frames: torch.Tensor | None = None) -> "PipelineGraph":
        """Replays the pass (on `frames` when given: uint8 [N, ...] of the captured shape, device or pinned host).
        Results: .D2, .D3, .D3_new, .stats (sigma3 statistics), .n_sweeps / .eps_trail (read lazily)."""
        if frames is not None:
            self.frames.copy_(frames.reshape(self.frames.shape), non_blocking=True)
        self.graph.replay()
        self._result = FutureCostResult(self.fc.mvec, pending=self._fc_static)     # static device buffers, refilled
        return self

    @property
    def n_sweeps(self) -> int:
        return self._result.n_sweeps

    @property
    def eps_trail(self):
        return self._result.eps_trail


# --------------------------------------------------------------------------- K3 / K4
class FutureCostResult:
    """Result of the future-cost iteration.  `mvec` ([M (padded)] fp32: D3_new = D3 + fl(alpha*mvec) on rows >= 1)
    is available immediately (stream-ordered); `n_sweeps` (== number of `Eps:` lines the reference prints),
    `eps_trail` and `passes` (streaming reads of D3 performed = n_sweeps + 1) of the fused kernels are read from
    the device lazily, so a caller that only needs `mvec` launches the next kernel without a host sync."""

    def __init__(self, mvec, n_sweeps=None, eps_trail=None, passes=0, pending=None):
        self.mvec = mvec
        self._n_sweeps, self._eps_trail, self._passes, self._pending = n_sweeps, eps_trail, passes, pending

    def _resolve(self):
        if self._pending is not None:
            info, trail, m, on_fail = self._pending
            self._pending = None
            h = [int(v) for v in info.cpu()]
            if len(h) > 2 and h[2] != 0:
                raise RuntimeError("future cost: a peer GPU did not reach the sweep barrier (timeout)")
            if h[0] == 0:
                if on_fail is None:
                    raise RuntimeError("future cost did not converge")
                other = on_fail()
                self.mvec.copy_(other.mvec[:self.mvec.shape[0]])
                self._n_sweeps, self._eps_trail, self._passes = other.n_sweeps, other.eps_trail, other.passes
                return
            self._n_sweeps = h[0]
            self._eps_trail = [float(np.float32(v / (float(m) * float(m)))) for v in trail[1:h[0] + 1].cpu()]
            self._passes = h[0] + 1

    @property
    def n_sweeps(self) -> int:
        self._resolve()
        return self._n_sweeps

    @property
    def eps_trail(self) -> list:
        self._resolve()
        return self._eps_trail

    @property
    def passes(self) -> int:
        self._resolve()
        return self._passes


def future_cost(D3: torch.Tensor, alpha: float = 0.997, row0: int = 0, m: int | None = None,
                exchange=None, pad_to: int | None = None, verbose: bool = False,
                max_sweeps: int = 10000) -> FutureCostResult:
    """K3 loop.  D3 holds global rows [row0, row0+rows) (rows may include a halo row: pass `rows`
    via slicing before the call).  `exchange(m_vec, eps_buf)` (multi-GPU) must all-gather the
    per-row minima in place and all-reduce the eps numerator; None on a single GPU.
    One host read of eps per sweep is inherent: the reference's stop rule is data dependent.
    """
    rows = D3.shape[0]
    m = D3.shape[1] if m is None else m
    length = m if pad_to is None else pad_to
    dev, st = _dev(D3), _stream(D3)
    bufs = [torch.zeros(length, dtype=torch.float32, device=D3.device) for _ in range(3)]
    eps_buf = torch.zeros(1, dtype=torch.float64, device=D3.device)
    alpha32 = C.c_float(_f32(alpha))

    def sweep(prev, prev2, out, eps):
        _lib.call("avtex_future_cost_sweep", _lib.ptr(D3), D3.stride(0), row0, rows, m, _lib.ptr(prev),
                  _lib.ptr(prev2), alpha32, _lib.ptr(out), _lib.ptr(eps), dev, st)

    sweep(None, None, bufs[0], None)                       # pass 0: m^0 = row minima of D3
    if exchange is not None:
        exchange(bufs[0], None)
    cur, prev2 = bufs[0], None
    free = [bufs[1], bufs[2]]
    trail = []
    for p in range(1, max_sweeps + 1):
        out = free.pop()
        eps_buf.zero_()
        sweep(cur, prev2, out, eps_buf)                    # m^p and the eps numerator of sweep p
        if exchange is not None:
            exchange(out, eps_buf)
        eps = np.float32(eps_buf.item() / (float(m) * float(m)))
        trail.append(float(eps))
        if verbose:
            print("Eps:", f"tensor({eps:.4f}, device='{D3.device}')")
        if not (eps > np.float32(F32_EPS_STOP)):
            return FutureCostResult(cur, p, trail, p + 1)
        if prev2 is not None:
            free.append(prev2)
        prev2, cur = cur, out
    raise RuntimeError("future cost did not converge")


def future_cost_fused(D3: torch.Tensor, alpha: float = 0.997, verbose: bool = False,
                      max_sweeps: int = 4096) -> FutureCostResult:
    """K3, all sweeps in one cooperative launch (single GPU): no host round trip per sweep, and none after the
    launch either — the converged vector lands in a fixed buffer, sweep count / eps trail are read on demand."""
    m = D3.shape[1]
    mpad = (m + 31) // 32 * 32
    mbuf = torch.zeros(4 * mpad, dtype=torch.float32, device=D3.device)          # 3 rotating vectors + the result
    trail = torch.zeros(max_sweeps + 1, dtype=torch.float64, device=D3.device)
    info = torch.zeros(4, dtype=torch.int32, device=D3.device)
    m_out = mbuf[3 * mpad:]
    _lib.call("avtex_future_cost_fused", _lib.ptr(D3), D3.stride(0), m, C.c_float(_f32(alpha)),
              C.c_float(np.float32(F32_EPS_STOP)), max_sweeps, _lib.ptr(mbuf), mpad, _lib.ptr(trail),
              _lib.ptr(info), _lib.ptr(m_out), _dev(D3), _stream(D3))
    res = FutureCostResult(m_out[:m], pending=(info, trail, m, None))
    if verbose:
        for e in res.eps_trail:
            print("Eps:", f"tensor({e:.4f}, device='{D3.device}')")
    return res


def pow_matrix(D: torch.Tensor, p: float) -> torch.Tensor:
    """D ** p (classic/q_learning.py:34) into a fresh padded matrix."""
    out = empty_matrix(D.shape[0], D.shape[1], D.device)
    _lib.call("avtex_pow_matrix", _lib.ptr(D), D.stride(0), D.shape[0], D.shape[1], C.c_float(_f32(p)),
              _lib.ptr(out), out.stride(0), _dev(D), _stream(D))
    return out


def future_cost_finalize(D3: torch.Tensor, mvec: torch.Tensor, alpha: float = 0.997, row0: int = 0,
                         m: int | None = None, stats: torch.Tensor | None = None) -> torch.Tensor:
    rows = D3.shape[0]
    m = D3.shape[1] if m is None else m
    out = empty_matrix(rows, m, D3.device)
    s, z = _stats_ptrs(stats)
    _lib.call("avtex_future_cost_finalize", _lib.ptr(D3), D3.stride(0), row0, rows, m, _lib.ptr(mvec),
              C.c_float(_f32(alpha)), _lib.ptr(out), out.stride(0), s, z, _dev(D3), _stream(D3))
    return out


# --------------------------------------------------------------------------- K6 / K7
def l2_normalize_rows(x: torch.Tensor) -> torch.Tensor:
    x = x if x.stride(-1) == 1 else x.contiguous()
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    _lib.call("avtex_l2_normalize_rows", _lib.ptr(x), x.stride(0), x.shape[0], x.shape[1], _lib.ptr(y),
              y.stride(0), _dev(x), _stream(x))
    return y


def cosine_scores(tn: torch.Tensor, qn: torch.Tensor, temp: float, out: torch.Tensor | None = None):
    out = torch.empty(tn.shape[0], dtype=torch.float32, device=tn.device) if out is None else out
    _lib.call("avtex_cosine_scores", _lib.ptr(tn), tn.stride(0), tn.shape[0], tn.shape[1], _lib.ptr(qn),
              C.c_float(_f32(temp)), _lib.ptr(out), _dev(tn), _stream(tn))
    return out


def select_step(o: torch.Tensor, a: torch.Tensor | None, q: int, alpha: float, threshold: float,
                choices: torch.Tensor, n_choices: torch.Tensor, vals: torch.Tensor | None = None):
    """K7.  `choices` (int32 [L]) / `n_choices` (int32 [1]) are caller-owned device buffers."""
    _lib.call("avtex_select_step", _lib.ptr(o), _lib.ptr(a), o.shape[0], int(q), C.c_float(_f32(alpha)),
              C.c_float(np.float32(1.0 - float(alpha))), C.c_float(_f32(threshold)), _lib.ptr(choices),
              _lib.ptr(n_choices), _lib.ptr(vals), _dev(o), _stream(o))


class SynthesisWorkspace:
    """Scratch of avtex_synthesis_step for tables of L windows, incl. the mapped pinned result buffer."""

    def __init__(self, L: int, device, host_cap: int = 4096):
        self.L, self.host_cap = L, min(L, host_cap)
        self.f32 = torch.empty(3 * L, dtype=torch.float32, device=device)
        self.acc = torch.zeros(8, dtype=torch.float64, device=device)
        self.mx = torch.zeros(4, dtype=torch.int32, device=device)
        self.counts = torch.zeros(4096, dtype=torch.int32, device=device)
        self.sel = torch.zeros(L + 1, dtype=torch.int32, device=device)         # [count | choices...]
        self.host = torch.zeros(2 + self.host_cap, dtype=torch.int32).pin_memory()
        self.host_np = self.host.numpy()
        self.seq = 0
        self.mt = torch.zeros(625, dtype=torch.int32, device=device)            # numpy MT19937 key[624] + pos
        self.q_scratch = torch.zeros(1, dtype=torch.int64, device=device)

    def reset(self):
        """Back to the freshly allocated state (the step kernel and the loop kernel keep different parities)."""
        self.acc.zero_()
        self.mx.zero_()
        self.host_np[:2] = 0
        self.seq = 0


def synthesis_step(ws: SynthesisWorkspace, tn: torch.Tensor, qrow: torch.Tensor, q: int, temp: float, alpha: float,
                   threshold: float, sn: torch.Tensor | None = None, drow: torch.Tensor | None = None,
                   vals: torch.Tensor | None = None) -> np.ndarray:
    """K6 + K7 in ONE cooperative launch; returns the survivor window ids (int32, target-list order).  The kernel
    writes the list into mapped pinned memory and the host polls the sequence word: no D2H copy call, no
    stream synchronisation."""
    ws.seq += 1
    seq = ws.seq
    _lib.call("avtex_synthesis_step", _lib.ptr(tn), tn.stride(0), tn.shape[0], tn.shape[1], _lib.ptr(qrow),
              _lib.ptr(sn), sn.stride(0) if sn is not None else 0, sn.shape[1] if sn is not None else 0,
              _lib.ptr(drow), C.c_float(_f32(temp)), int(q), C.c_float(_f32(alpha)),
              C.c_float(np.float32(1.0 - float(alpha))), C.c_float(_f32(threshold)), _lib.ptr(ws.f32),
              _lib.ptr(ws.acc), _lib.ptr(ws.mx), _lib.ptr(ws.counts), ws.counts.shape[0],
              C.c_void_p(ws.sel.data_ptr() + 4), _lib.ptr(ws.sel), _lib.ptr(vals), _lib.ptr(ws.host), ws.host_cap,
              seq, _dev(tn), _stream(tn))
    h = ws.host_np
    spins = 0
    stream = torch.cuda.current_stream(tn.device)
    while h[0] != seq:
        spins += 1
        if spins % 20000 == 0 and stream.query() and h[0] != seq:
            torch.cuda.synchronize()                       # surfaces a launch / execution error if there was one
            if h[0] != seq:
                raise _lib.AvtexError("synthesis_step: the kernel finished without publishing its result")
    n = int(h[1])
    if n <= ws.host_cap:
        return h[2:2 + n].copy()
    return ws.sel[1:n + 1].cpu().numpy()


def mt19937_randint_host(key: np.ndarray, pos: int, ns) -> tuple:
    """Host copy of the device generator (test hook): draws RandomState.randint(0, n) for every n in `ns` from the
    MT19937 state (key uint32[624], pos).  Returns (draws, new_key, new_pos)."""
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    ns = np.ascontiguousarray(ns, dtype=np.uint32)
    out = np.zeros(len(ns), dtype=np.uint32)
    p = C.c_int(int(pos))
    _lib.call("avtex_mt19937_randint_host", key.ctypes.data_as(C.c_void_p), C.byref(p), ns.ctypes.data_as(C.c_void_p),
              len(ns), out.ctypes.data_as(C.c_void_p))
    return out, key, p.value


def synthesis_loop(ws: SynthesisWorkspace, tn: torch.Tensor, qn: torch.Tensor, q_start: int, n_steps: int, temp: float,
                   alpha: float, threshold: float, sn: torch.Tensor | None = None, dn: torch.Tensor | None = None):
    """The whole synthesis loop in ONE cooperative launch (avtex_synthesis_loop): numpy's global MT19937 state goes
    to the device, every step's `np.random.choice` is drawn there by the same algorithm, and the advanced state is
    written back with np.random.set_state — the host's random stream continues exactly where the reference's loop
    would have left it.  Returns (q_ids int32[n_steps], nz_counts int32[n_steps]) as numpy arrays."""
    if dn is not None and dn.shape[0] < n_steps + 1:
        raise IndexError(f"driving table has {dn.shape[0]} rows, step {n_steps} reads row {n_steps}")   # as the reference's indexing would
    kind, key, pos, has_gauss, cached = np.random.get_state()
    if kind != "MT19937":
        raise RuntimeError("numpy's global generator is not MT19937")
    host_state = np.empty(625, dtype=np.uint32)
    host_state[:624] = key
    host_state[624] = pos
    ws.reset()
    ws.mt.copy_(torch.from_numpy(host_state.view(np.int32)))
    out = torch.empty(2 * n_steps, dtype=torch.int32, device=tn.device)
    _lib.call("avtex_synthesis_loop", _lib.ptr(tn), tn.stride(0), tn.shape[0], tn.shape[1], _lib.ptr(qn), qn.stride(0),
              _lib.ptr(sn), sn.stride(0) if sn is not None else 0, sn.shape[1] if sn is not None else 0,
              _lib.ptr(dn), dn.stride(0) if dn is not None else 0, C.c_float(_f32(temp)), C.c_float(_f32(alpha)),
              C.c_float(np.float32(1.0 - float(alpha))), C.c_float(_f32(threshold)), int(q_start), int(n_steps),
              _lib.ptr(ws.f32), _lib.ptr(ws.acc), C.c_void_p(ws.sel.data_ptr() + 4), _lib.ptr(ws.sel), _lib.ptr(ws.mt),
              _lib.ptr(ws.q_scratch), _lib.ptr(out), C.c_void_p(out.data_ptr() + 4 * n_steps), _dev(tn), _stream(tn))
    h = out.cpu().numpy()
    st = ws.mt.cpu().numpy().view(np.uint32)
    np.random.set_state((kind, st[:624].copy(), int(st[624]), has_gauss, cached))
    ws.reset()
    return h[:n_steps].copy(), h[n_steps:].copy()


def audio_start(x: torch.Tensor, d: torch.Tensor) -> int:
    """a10: start window = first arg-max of the cosine similarity to the first driving example."""
    x = x if x.stride(-1) == 1 else x.contiguous()
    ws = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    out = torch.zeros(1, dtype=torch.int32, device=x.device)
    _lib.call("avtex_audio_start", _lib.ptr(x), x.stride(0), x.shape[0], x.shape[1], _lib.ptr(d.contiguous()),
              _lib.ptr(ws), _lib.ptr(out), _dev(x), _stream(x))
    return int(out.item())


def gram_tile_schedule(TM: int, TN: int, symmetric: bool, two_cta: bool = False, group: int | None = None):
    """Test hook (host only, no GPU): visiting order of the Gram tiles (128x256, or 256x256 for 2-CTA)."""
    lib = _lib.load()
    if group is not None:
        total = lib.avtex_gram_tile_schedule2g(TM, TN, 1 if symmetric else 0, group, None, None, 0)
        tm = (C.c_int * total)()
        tn = (C.c_int * total)()
        lib.avtex_gram_tile_schedule2g(TM, TN, 1 if symmetric else 0, group, tm, tn, total)
        return list(zip(tm, tn))
    fn = lib.avtex_gram_tile_schedule2 if two_cta else lib.avtex_gram_tile_schedule
    total = fn(TM, TN, 1 if symmetric else 0, None, None, 0)
    tm = (C.c_int * total)()
    tn = (C.c_int * total)()
    fn(TM, TN, 1 if symmetric else 0, tm, tn, total)
    return list(zip(tm, tn))

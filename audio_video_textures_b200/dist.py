"""Row-sharded classic++ transition matrix over the GPUs of one box (one process per GPU).

Rank r owns rows [a0, a1) of every M x M matrix (D2, D3, D3_new, P3) and the D1 rows those need
(`a*stride + k`, plus one halo output row for the row shift of P, over-computed locally instead of
exchanged: ~fs extra D1 rows out of N/G).  Frames are replicated (N*K bytes, 1.2 GB at N = 100k,
64x64).  The ONLY data-path exchange is the one the algorithm has: after each future-cost sweep
the per-row minima (M fp32) are all-gathered and the eps numerator (one fp64) is all-reduced
(BASELINE.json north_star).  sigma needs one more all-reduce of (sum, nnz).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import engine


@dataclass(frozen=True)
class ShardPlan:
    n: int            # frames
    m: int            # filtered size
    world: int
    rank: int
    shard: int        # rows per rank (last rank may own fewer)
    a0: int           # first owned D2/D3 row
    a1: int           # one past the last owned row
    a1h: int          # one past the last computed row (owned + 1 halo row for the P shift)
    r_lo: int         # first D1 row computed
    r_hi: int         # one past the last D1 row computed

    @property
    def padded(self) -> int:
        return self.shard * self.world


def plan_shards(n: int, filter_size: int, stride: int, world: int, rank: int) -> ShardPlan:
    m = engine.filtered_size(n, filter_size, stride)
    shard = (-(-m // world) + 3) // 4 * 4          # multiple of 4 rows: keeps every shard's D1 rows 16-byte aligned
    a0 = rank * shard
    a1 = min(m, a0 + shard)
    if a0 >= m:
        raise ValueError(f"rank {rank} of {world} would own no rows of the {m} x {m} matrix")
    a1h = min(m, a1 + 1)
    r_hi = n if a1 == m else (a1h - 1) * stride + filter_size      # the last rank keeps the unused tail rows too
    return ShardPlan(n, m, world, rank, shard, a0, a1, a1h, a0 * stride, r_hi)


def load_frames_sharded(host_frames: torch.Tensor, rank: int, world: int, device, group=None) -> torch.Tensor:
    """Replicates a HOST-resident uint8 clip on every GPU without sending it over every PCIe link: each
    rank copies only its 1/G slice host->device, then the slices are all-gathered over NVLink (in place).
    Copying the whole clip on all ranks at once is bound by the host side (8 x 2.1 GB took 95 ms)."""
    x = host_frames.reshape(host_frames.shape[0], -1)
    n, k = x.shape
    per = -(-n // world)
    full = torch.empty((per * world, k), dtype=x.dtype, device=device)
    lo, hi = rank * per, min(n, (rank + 1) * per)
    if hi > lo:
        full[lo:hi].copy_(x[lo:hi], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per], group=group)
    return full[:n]


def pack_frames_sharded(frames: torch.Tensor, rank: int, world: int, group=None) -> engine.PackedFrames:
    """K0 for the replicated clip: every rank computes the norms of 1/G of the frames and the [N] int64 vector is
    all-gathered (113 KB at N = 14144) instead of every rank reading all N*K bytes again."""
    x = frames.reshape(frames.shape[0], -1)
    n, k = x.shape
    if world == 1 or x.dtype != torch.uint8 or x.stride(0) % 16 != 0 or x.data_ptr() % 16 != 0 or x.stride(1) != 1:
        return engine.pack_frames(frames)
    per = -(-n // world)
    sqnorm = torch.zeros(per * world, dtype=torch.int64, device=x.device)
    flags = torch.zeros(2, dtype=torch.int64, device=x.device)
    lo, hi = rank * per, min(n, (rank + 1) * per)
    if hi > lo:
        engine.frame_norms_rows(x, lo, hi - lo, sqnorm, flags)
    dist.all_gather_into_tensor(sqnorm, sqnorm[rank * per:(rank + 1) * per], group=group)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    pf = engine.PackedFrames(x, sqnorm[:n], k, flags, signed=False)
    if (k + 127) // 128 * 128 * 128 * 128 < engine.GRAM_MAX_SQNORM:
        pf._checked = (True, "")
    return pf


def make_exchange(plan: ShardPlan, group=None):
    """Returns exchange(mvec, eps_buf): in-place all-gather of the per-row minima (each rank wrote
    its own rows of the padded vector) + all-reduce of the eps numerator."""
    if plan.world == 1:
        return None
    backend = dist.get_backend(group)

    def exchange(mvec: torch.Tensor, eps_buf):
        mine = mvec[plan.rank * plan.shard:(plan.rank + 1) * plan.shard]
        if backend == "nccl":
            dist.all_gather_into_tensor(mvec, mine, group=group)          # in place: NCCL allows it
        else:
            chunks = list(mvec.view(plan.world, plan.shard).unbind(0))
            dist.all_gather(chunks, mine.clone(), group=group)
        if eps_buf is not None:
            dist.all_reduce(eps_buf, op=dist.ReduceOp.SUM, group=group)

    return exchange


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """(fp64 sum, uint64 nnz) summed over ranks.  The count travels as int64 bits in a second tensor
    because a fp64 SUM would not add integer bit patterns."""
    s = stats[:1].clone()
    z = stats.view(torch.int64)[1:].clone()
    dist.all_reduce(s, group=group)
    dist.all_reduce(z, group=group)
    out = stats.clone()
    out[:1] = s
    out.view(torch.int64)[1:] = z
    return out


def pair_split(lo: int, hi: int) -> int:
    """Column split point of core_j for the pair (i < j): rank i computes core_i x [lo, mid) and rank j
    computes [mid, hi) x core_i, each pushing the transpose to the other, so both do half of the rectangle
    whatever the number of ranks.  Aligned to the 256-wide tile."""
    mid = lo + (hi - lo) // 2
    mid = lo + (mid - lo + 255) // 256 * 256
    return min(mid, hi)


class SymmetricShardWorkspace:
    """D1 row shards in torch symmetric memory (peer-mapped over NVLink / NVSwitch).

    Row sharding alone forfeits the factor-2 symmetry of the distance matrix: rank I needs D1[rows_I, :],
    and D1[rows_I, cols_J] is the transpose of D1[rows_J, cols_I] that rank J needs.  Here each off-diagonal
    rectangle is split between the two ranks; the same tcgen05 kernel that computes a tile stores it
    locally and pushes the transposed tile straight into the peer's shard (coalesced 128-byte NVLink
    stores from the epilogue), so every rank does N^2/(2G) pairs and no separate exchange step exists.
    """

    def __init__(self, n: int, filter_size: int, stride: int, rank: int, world: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.n, self.fs, self.stride, self.rank, self.world = n, filter_size, stride, rank, world
        self.plans = [plan_shards(n, filter_size, stride, world, r) for r in range(world)]
        self.plan = self.plans[rank]
        self.ld = (n + 31) // 32 * 32
        rows_max = max(p.r_hi - p.r_lo for p in self.plans)
        group = dist.group.WORLD if group is None else group
        self.buf = symm_mem.empty((rows_max, self.ld), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.group = group

    def core(self, r: int):
        """Disjoint cover of the frame rows: the rows whose distances rank r is responsible for."""
        p = self.plans[r]
        return p.a0 * self.stride, (self.n if r == self.world - 1 else p.a1 * self.stride)

    def D1(self) -> torch.Tensor:
        p = self.plan
        return self.buf[:p.r_hi - p.r_lo, :self.n]

    def jobs(self):
        me, p = self.rank, self.plan
        lo, hi = self.core(me)
        mine = dict(D=self.ptrs[me], d_row0=p.r_lo, ldd=self.ld)
        out = [dict(row0=lo, rows=hi - lo, col0=lo, cols=hi - lo, symmetric=1, count_stats=1,
                    DT=self.ptrs[me], dt_row0=p.r_lo, ldt=self.ld, **mine)]
        # peers are visited in rotated order (me+1, me+2, ...): at any moment every rank pushes into a
        # DIFFERENT destination, instead of all ranks hammering rank 0's NVLink ingress first, then rank 1's...
        for step in range(1, self.world):
            other = (me + step) % self.world
            olo, ohi = self.core(other)
            peer = dict(DT=self.ptrs[other], dt_row0=self.plans[other].r_lo, ldt=self.ld)
            if me < other:                                  # my rows x the first half of the peer's columns
                mid = pair_split(olo, ohi)
                if mid > olo:
                    out.append(dict(row0=lo, rows=hi - lo, col0=olo, cols=mid - olo, symmetric=0, count_stats=1,
                                    **peer, **mine))
            else:                                           # the second half of my rows x the peer's columns
                mid = pair_split(lo, hi)
                if hi > mid:
                    out.append(dict(row0=mid, rows=hi - mid, col0=olo, cols=ohi - olo, symmetric=0, count_stats=1,
                                    **peer, **mine))
        return out

    def halo_rows(self) -> int:
        """Rows past this rank's core that its filter outputs read: the next rank's first rows."""
        return self.plan.r_hi - self.core(self.rank)[1]

    def gram(self, pf: engine.PackedFrames, stats=None) -> torch.Tensor:
        """Fills this rank's D1 shard (its own tiles + the tiles peers push).  Stream-ordered barriers on
        both sides: peers must be done reading the previous contents, and done pushing, respectively."""
        self.hdl.barrier(channel=0)
        engine.gram_l2_jobs(pf, self.jobs(), stats)
        self.hdl.barrier(channel=1)
        halo = self.halo_rows()
        if halo > 0:
            # the halo rows are the first core rows of the next rank: one small peer copy over NVLink
            # (fs rows) instead of fs-row MMA tiles that would waste 216 of their 256 rows
            p, nxt = self.plan, self.rank + 1
            src = self.hdl.get_buffer(nxt, (halo, self.ld), torch.float32)
            lo = self.core(self.rank)[1] - p.r_lo
            self.buf[lo:lo + halo].copy_(src)
        return self.D1()


class PeerFutureCost:
    """Symmetric-memory workspace + driver of avtex_future_cost_fused_peer: the whole row-sharded
    future-cost iteration in one cooperative kernel per GPU, exchanging the row minima by peer stores."""
    MAX_SWEEPS = 254

    def __init__(self, m: int, rank: int, world: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.m, self.rank, self.world = m, rank, world
        self.mpad = (m + 31) // 32 * 32
        self.n_m = 6 * self.mpad * 4                                   # two sets of three fp32 vectors
        self.n_eps = (self.MAX_SWEEPS + 1) * world * 8
        total = self.n_m + self.n_eps + 256
        group = dist.group.WORLD if group is None else group
        self.buf = symm_mem.empty((total,), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        self.hdl.barrier(channel=0)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self.calls = 0
        self.device = device

    def run(self, D3_own: torch.Tensor, row0: int, alpha: float = 0.997, verbose: bool = False):
        import ctypes as C
        import numpy as np
        from . import _lib
        rows = D3_own.shape[0]
        parity = self.calls % 2
        arr = C.c_void_p * self.world
        mptr = arr(*[b + parity * 3 * self.mpad * 4 for b in self.base])
        eptr = arr(*[b + self.n_m for b in self.base])
        fptr = arr(*[b + self.n_m + self.n_eps for b in self.base])
        eps_local = torch.zeros(self.MAX_SWEEPS + 1, dtype=torch.float64, device=self.device)
        trail = torch.zeros(self.MAX_SWEEPS + 1, dtype=torch.float64, device=self.device)
        info = torch.zeros(2, dtype=torch.int32, device=self.device)
        epoch = (self.calls * (self.MAX_SWEEPS + 2)) & 0xFFFFFFFF
        _lib.call("avtex_future_cost_fused_peer", _lib.ptr(D3_own), D3_own.stride(0), row0, rows, self.m,
                  C.c_float(engine._f32(alpha)), C.c_float(np.float32(engine.F32_EPS_STOP)), self.MAX_SWEEPS,
                  self.rank, self.world, mptr, self.mpad, eptr, fptr, C.c_uint(epoch), _lib.ptr(eps_local),
                  _lib.ptr(trail), _lib.ptr(info), engine._dev(D3_own), engine._stream(D3_own))
        self.calls += 1
        n_sweeps, idx = (int(v) for v in info.cpu())
        if n_sweeps == 0:
            raise RuntimeError("future cost did not converge")
        m_all = self.buf[:self.n_m].view(torch.float32)
        off = (parity * 3 + idx) * self.mpad
        eps = [float(np.float32(v / (float(self.m) ** 2))) for v in trail[1:n_sweeps + 1].cpu()]
        if verbose:
            for e in eps:
                print("Eps:", f"tensor({e:.4f})")
        return engine.FutureCostResult(m_all[off:off + self.m], n_sweeps, eps, n_sweeps + 1)


@dataclass
class ShardResult:
    plan: ShardPlan
    D1: torch.Tensor          # rows [r_lo, r_hi)
    D2: torch.Tensor          # rows [a0, a1h)
    D3: torch.Tensor
    D3_new: torch.Tensor      # rows [a0, a1h)
    n_sweeps: int
    eps_trail: list
    sigma: float | None = None
    P3: torch.Tensor | None = None        # rows [a0, a1)
    P3_new: torch.Tensor | None = None
    counts: torch.Tensor | None = None
    launches: int = 0
    stage_ms: dict | None = None          # CUDA-event stage times when AVTEX_DIST_TIMING=1


def classic_sharded(frames: torch.Tensor, filter_size: int, stride: int, rank: int, world: int,
                    p: float = 0.7, alpha: float = 0.997, sigma_factor=None, threshold=None, group=None,
                    packed: engine.PackedFrames | None = None,
                    workspace: SymmetricShardWorkspace | None = None,
                    peer_fc: PeerFutureCost | None = None) -> ShardResult:
    """Distance + filter + converged future cost (+ sigma3 / P3 / P3_new when sigma_factor is given) for
    this rank's rows.  `frames`: the full [N, ...] uint8 clip on this rank's device.  With a
    SymmetricShardWorkspace the distance stage uses the symmetry across ranks (peer pushes over NVLink);
    without one every rank computes its full row block locally."""
    import os
    timing = os.environ.get("AVTEX_DIST_TIMING") == "1"
    marks = []

    def mark(name):
        if timing:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    n = frames.shape[0]
    plan = plan_shards(n, filter_size, stride, world, rank)
    mark("start")
    pf = pack_frames_sharded(frames, rank, world, group) if packed is None else packed
    mark("norms")
    if not pf.exact_ok:
        raise engine._lib.AvtexError(f"sharded path needs byte frames inside the Gram domain: {pf.reason}")
    if workspace is not None and world > 1:
        D1 = workspace.gram(pf)                            # symmetric across ranks: transposed tiles pushed to peers
    else:
        D1 = engine.gram_l2(pf, plan.r_lo, plan.r_hi - plan.r_lo, symmetric=False)
    mark("gram")
    D2, D3 = engine.diag_filter(D1, filter_size, stride, p=p, m=plan.m, a0=plan.a0,
                                rows_out=plan.a1h - plan.a0, in_row0=plan.r_lo)
    own = plan.a1 - plan.a0
    mark("filter")
    if peer_fc is not None and world > 1:
        fc = peer_fc.run(D3[:own], plan.a0, alpha)         # all sweeps + exchanges inside one kernel per GPU
        n_fc_launches = 1
    else:
        fc = engine.future_cost(D3[:own], alpha, row0=plan.a0, m=plan.m, exchange=make_exchange(plan, group),
                                pad_to=plan.padded)
        n_fc_launches = fc.passes
    res = ShardResult(plan, D1, D2, D3, None, fc.n_sweeps, fc.eps_trail)
    res.launches = 1 + 1 + 1 + n_fc_launches + 1
    mark("future_cost")
    res.D3_new = engine.future_cost_finalize(D3, fc.mvec, alpha, row0=plan.a0, m=plan.m)
    mark("finalize")
    if timing:
        torch.cuda.synchronize()
        res.stage_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
    if sigma_factor is not None:
        stats = engine.sum_nnz(res.D3_new[:own])
        if world > 1:
            stats = allreduce_stats(stats, group)
        res.sigma = engine.sigma_from_stats(*engine.read_stats(stats), sigma_factor)
        res.P3, res.P3_new, res.counts = engine.transition_probs(
            res.D3_new, res.sigma, shift=1, rows_out=own, threshold=threshold, want_counts=threshold is not None)
        res.launches += 2
    return res


def gather_survivors(res: ShardResult, group=None):
    """CSR of P3_new over all ranks, assembled on every rank (the walk runs on rank 0's host)."""
    rowptr, colidx = engine.csr_from_matrix(res.P3_new, res.counts)
    if res.plan.world == 1:
        return rowptr, colidx
    parts = [None] * res.plan.world
    dist.all_gather_object(parts, (rowptr, colidx), group=group)
    import numpy as np
    counts = np.concatenate([np.diff(rp) for rp, _ in parts])
    full_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    return full_ptr, np.concatenate([ci for _, ci in parts])

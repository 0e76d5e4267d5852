"""Row-sharded classic++ transition matrix over the GPUs of one box (one process per GPU).

Rank r owns rows [a0, a1) of every M x M matrix (D2, D3, D3_new, P3) and the D1 rows those need
(`a*stride + k`, plus one halo output row for the row shift of P).  Frames are replicated (N*K bytes,
1.2 GB at N = 100k, 64x64).  The ONLY data-path exchange the algorithm has is the per-sweep all-gather of
the row minima + the eps numerator (BASELINE.json north_star); it and the two exchanges
the sharding adds (frame norms, transposed Gram tiles) are done by the kernels themselves through
peer-mapped pointers:

  * K0   every rank computes the norms of 1/G of the frames and PUSHES them into every peer's vector
  * K1   symmetric Gram: each off-diagonal rectangle is split between the two ranks that need it; the tile
         kernel stores it locally and pushes the transpose into the peer's shard (NVLink stores)
  * K3   one cooperative kernel per GPU runs all sweeps; minima / eps are pushed to the peers, a flag
         barrier closes each sweep

`RankStep` holds one rank's phases; `classic_sharded` runs them for a real rank (torch symmetric memory,
stream-ordered barriers between phases), `VirtualBox` runs G of them on ONE device (peer pointers = plain
local buffers, ranks executed phase by phase) so that the 8-rank job lists, halo plans and the flag
protocol are parity-tested on a single-GPU box.  The NCCL form of the same algorithm (`make_exchange`,
`--no_symmetric`) remains as the fallback when symmetric memory is unavailable and as the gloo-testable
host logic.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine


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


def core_rows(plans: list, r: int, stride: int):
    """Disjoint cover of the frame rows: the D1 rows rank r is responsible for computing."""
    p = plans[r]
    return p.a0 * stride, (p.n if r == p.world - 1 else p.a1 * stride)


def pair_split(lo: int, hi: int, big_first: bool = True) -> int:
    """Column split point of core_j for the pair (i < j): rank i computes core_i x [lo, mid) and rank j
    computes [mid, hi) x core_i, each pushing the transpose to the other, so both do half of the rectangle
    whatever the number of ranks.  Aligned to the 256-wide tile; `big_first` says which side takes the larger
    part when the range is not a multiple of two tiles (alternated over the pairs so that no rank collects all
    the larger halves: with the split always rounded up, rank 0 of 8 did 224 tiles and rank 7 175)."""
    half = (hi - lo) // 2
    up = min(hi, lo + (half + 255) // 256 * 256)
    if -(-(up - lo) // 256) == -(-(hi - up) // 256):       # rounding up already gives both sides the same tile count
        return up
    mid = up if big_first else lo + half // 256 * 256
    return max(lo, min(mid, hi))


def _big_first(i: int, j: int) -> bool:
    return (i + j) % 2 == 1


def symmetric_jobs(plans: list, me: int, ptrs: list, ld: int, stride: int) -> list:
    """Gram job list of rank `me` (pure function of the plans and the shard base pointers): its own
    symmetric diagonal block plus, per peer, its half of the off-diagonal rectangle with the transposed
    destination in the PEER's shard.  Peers are visited in rotated order (me+1, me+2, ...): at any moment
    every rank pushes into a DIFFERENT destination instead of all ranks hammering one GPU's NVLink ingress."""
    world = len(plans)
    p = plans[me]
    lo, hi = core_rows(plans, me, stride)
    mine = dict(D=ptrs[me], d_row0=p.r_lo, ldd=ld)
    out = [dict(row0=lo, rows=hi - lo, col0=lo, cols=hi - lo, symmetric=1, count_stats=1,
                DT=ptrs[me], dt_row0=p.r_lo, ldt=ld, **mine)]
    for step in range(1, world):
        other = (me + step) % world
        olo, ohi = core_rows(plans, other, stride)
        peer = dict(DT=ptrs[other], dt_row0=plans[other].r_lo, ldt=ld)
        if me < other:                                  # my rows x the first half of the peer's columns
            mid = pair_split(olo, ohi, _big_first(me, other))
            if mid > olo:
                out.append(dict(row0=lo, rows=hi - lo, col0=olo, cols=mid - olo, symmetric=0, count_stats=1,
                                **peer, **mine))
        else:                                           # the second half of my rows x the peer's columns
            mid = pair_split(lo, hi, _big_first(other, me))
            if hi > mid:
                out.append(dict(row0=mid, rows=hi - mid, col0=olo, cols=ohi - olo, symmetric=0, count_stats=1,
                                **peer, **mine))
    return out


def halo_sources(plans: list, me: int, stride: int) -> list:
    """Rows past rank `me`'s core that its filter outputs read, as (owner rank, first row, rows) pieces taken
    from the owners' CORE rows only.  When shard*stride < filter_size (small M over many ranks) the halo spans
    several following ranks; reading a neighbour's own halo region instead would race with its halo copy."""
    need_lo, need_hi = core_rows(plans, me, stride)[1], plans[me].r_hi
    out = []
    r = me + 1
    while need_lo < need_hi:
        if r >= len(plans):
            raise ValueError("halo rows beyond the last rank's core")
        clo, chi = core_rows(plans, r, stride)
        take = min(need_hi, chi) - need_lo
        if take > 0:
            out.append((r, need_lo, take))
            need_lo += take
        r += 1
    return out


# --------------------------------------------------------------------------- NCCL form (fallback / gloo tests)
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


def load_frames_pushed(host_frames: torch.Tensor, ws: "SymmetricShardWorkspace", chunks: int = 4) -> torch.Tensor:
    """load_frames_sharded without NCCL and without waiting for the whole slice: every rank copies its 1/G of the
    pinned host clip in `chunks` pieces and, as each piece lands, PUSHES it into all peers' copies of the clip through
    peer-mapped symmetric memory (plain device copies over NVLink, peers visited in rotated order), so the NVLink
    replication runs underneath the PCIe transfer instead of after it.  Stream-ordered symmetric-memory barriers fence
    the buffers: nobody writes into a peer that may still be reading the previous clip, nobody computes before every
    push has landed.  Returns the [N, K] uint8 clip on this rank's device (a view of the symmetric buffer)."""
    x = host_frames.reshape(host_frames.shape[0], -1)
    n, k = x.shape
    world, rank = ws.world, ws.rank
    buf = ws.frames_buffer(n, k)
    per = buf.shape[0] // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    main = torch.cuda.current_stream(ws.device)
    ws._hdl_frames.barrier(channel=0)                   # every rank is done with the previous clip (stream-ordered)
    ws._copy_stream.wait_stream(main)
    ws._push_stream.wait_stream(main)
    step = max(1, -(-(hi - lo) // chunks))
    for c0 in range(lo, hi, step):
        c1 = min(hi, c0 + step)
        with torch.cuda.stream(ws._copy_stream):
            buf[c0:c1].copy_(x[c0:c1], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(ws._copy_stream)
        with torch.cuda.stream(ws._push_stream):
            ws._push_stream.wait_event(landed)
            for s_ in range(1, world):
                ws._frame_peers[(rank + s_) % world][c0:c1].copy_(buf[c0:c1], non_blocking=True)
    main.wait_stream(ws._copy_stream)
    main.wait_stream(ws._push_stream)
    ws._hdl_frames.barrier(channel=1)                   # all pushes of all ranks have landed
    return buf[:n]


def pack_frames_sharded(frames: torch.Tensor, rank: int, world: int, group=None) -> engine.PackedFrames:
    """K0 for the replicated clip, NCCL form: every rank computes the norms of 1/G of the frames and the [N]
    int64 vector is all-gathered.  (The symmetric-memory path pushes them from the kernel instead.)"""
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


# --------------------------------------------------------------------------- peer-mapped workspaces
class PeerBuffers:
    """What one rank sees of the box: for every rank the base pointers of its D1 shard, norm vectors and
    future-cost scratch, plus stream-ordered barriers.  Two providers: torch symmetric memory over NVLink
    (`SymmetricShardWorkspace`) and plain local buffers shared by virtual ranks (`VirtualBox`)."""
    MAX_SWEEPS = 2046

    def __init__(self, n, filter_size, stride, rank, world, device, residues: bool = False):
        self.n, self.fs, self.stride, self.rank, self.world, self.device = n, filter_size, stride, rank, world, device
        self.plans = [plan_shards(n, filter_size, stride, world, r) for r in range(world)]
        self.plan = self.plans[rank]
        # residues: the shards hold the `stride` residue-class planes (engine.gram_l2_residues) instead of D1 rows.  In
        # class coordinates the pipeline is a stride-1 filter of fs/stride taps per plane over N/stride frames, so
        # the SAME planner describes it (cplans: class rows owned / computed / haloed by every rank) and the output
        # row ranges (a0, a1, a1h) coincide with the full plan's.
        self.residues = bool(residues)
        if self.residues:
            if (filter_size, stride) not in engine.RESIDUE_FAST or n % stride or filter_size % stride:
                raise ValueError(f"residue-class shards need (fs, stride) in {sorted(engine.RESIDUE_FAST)} and "
                                 f"N % stride == 0 (got fs={filter_size}, stride={stride}, N={n})")
            self.cplans = [plan_shards(n // stride, filter_size // stride, 1, world, r) for r in range(world)]
            assert all((c.m, c.a0, c.a1, c.a1h) == (p.m, p.a0, p.a1, p.a1h) for c, p in zip(self.cplans, self.plans))
            self.ld = (n // stride + 31) // 32 * 32
            self.rows_c = max(c.r_hi - c.r_lo for c in self.cplans)           # class rows per plane and rank
            self.rows_max = stride * self.rows_c                              # the buffer stacks the planes
        else:
            self.cplans = None
            self.ld = (n + 31) // 32 * 32
            self.rows_max = max(p.r_hi - p.r_lo for p in self.plans)
        self.m = self.plan.m
        self.mpad = (self.m + 31) // 32 * 32
        self.npad = (n + 31) // 32 * 32
        # byte layout of the small per-rank buffer: [2 x sqnorm int64 npad | 2 x flags int64[2] | 2 x 3 m-vectors |
        # 2 x eps slots | sweep flags]
        self.off_sq = 0
        self.off_fl = self.off_sq + 2 * self.npad * 8
        self.off_m = self.off_fl + 2 * 16
        self.off_eps = self.off_m + 6 * self.mpad * 4
        self.n_eps = (self.MAX_SWEEPS + 1) * world * 8
        self.off_sweep = self.off_eps + 2 * self.n_eps
        self.small_bytes = self.off_sweep + 256
        self.calls = 0
        # filled by the provider: d1 (local [rows_max, ld] fp32), small (local uint8), d1_ptrs / small_ptrs (ints)
        self.d1 = self.small = None
        self.d1_ptrs = self.small_ptrs = None

    def barrier(self, channel: int):
        raise NotImplementedError

    def peer_d1_rows(self, r: int, row_off: int, rows: int) -> torch.Tensor:
        raise NotImplementedError

    # -- views into this rank's small buffer
    def sqnorm(self, parity: int) -> torch.Tensor:
        return self.small[self.off_sq + parity * self.npad * 8: self.off_sq + (parity + 1) * self.npad * 8].view(torch.int64)

    def flags(self, parity: int) -> torch.Tensor:
        return self.small[self.off_fl + parity * 16: self.off_fl + (parity + 1) * 16].view(torch.int64)

    def D1(self) -> torch.Tensor:
        p = self.plan
        return self.d1[:p.r_hi - p.r_lo, :self.n]

    def planes(self) -> torch.Tensor:
        """residues: this rank's class planes [stride, class rows held, ld] (plane stride = rows_c * ld)."""
        c = self.cplans[self.rank]
        return self.d1.view(self.stride, self.rows_c, self.ld)[:, :c.r_hi - c.r_lo]

    def peer_plane_rows(self, r: int, plane: int, row_off: int, rows: int) -> torch.Tensor:
        return self.peer_d1_rows(r, plane * self.rows_c + row_off, rows)


class SymmetricShardWorkspace(PeerBuffers):
    """Shards in torch symmetric memory (peer-mapped over NVLink / NVSwitch).

    Row sharding alone forfeits the factor-2 symmetry of the distance matrix: rank I needs D1[rows_I, :],
    and D1[rows_I, cols_J] is the transpose of D1[rows_J, cols_I] that rank J needs.  Here each off-diagonal
    rectangle is split between the two ranks; the same tcgen05 kernel that computes a tile stores it
    locally and pushes the transposed tile straight into the peer's shard (coalesced 128-byte NVLink
    stores from the epilogue), so every rank does N^2/(2G) pairs and no separate exchange step exists.
    """

    def __init__(self, n: int, filter_size: int, stride: int, rank: int, world: int, device, group=None,
                 residues: bool = False):
        import torch.distributed._symmetric_memory as symm_mem
        super().__init__(n, filter_size, stride, rank, world, device, residues)
        group = dist.group.WORLD if group is None else group
        self.group = group
        self.d1 = symm_mem.empty((self.rows_max, self.ld), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.d1, group.group_name)
        self.d1_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.small = symm_mem.empty((self.small_bytes,), dtype=torch.uint8, device=device)
        self.small.zero_()
        self.hdl_small = symm_mem.rendezvous(self.small, group.group_name)
        self.small_ptrs = [int(p) for p in self.hdl_small.buffer_ptrs]
        self.hdl.barrier(channel=0)                    # every rank's zero fill is done before anyone pushes

    def barrier(self, channel: int):
        self.hdl.barrier(channel=channel)

    def peer_d1_rows(self, r: int, row_off: int, rows: int) -> torch.Tensor:
        buf = self.hdl.get_buffer(r, (self.rows_max, self.ld), torch.float32)
        return buf[row_off:row_off + rows]

    def frames_buffer(self, n: int, k: int) -> torch.Tensor:
        """The replicated clip [N padded to world * ceil(N / world), K] uint8 in symmetric memory (allocated on first
        use, collectively), plus every rank's view of every peer's copy (`_frame_peers`)."""
        per = -(-n // self.world)
        shape = (per * self.world, k)
        if getattr(self, "_frames", None) is None or tuple(self._frames.shape) != shape:
            import torch.distributed._symmetric_memory as symm_mem
            self._frames = symm_mem.empty(shape, dtype=torch.uint8, device=self.device)
            self._hdl_frames = symm_mem.rendezvous(self._frames, self.group.group_name)
            self._frame_peers = [self._hdl_frames.get_buffer(r, shape, torch.uint8) for r in range(self.world)]
            self._copy_stream = torch.cuda.Stream(self.device)
            self._push_stream = torch.cuda.Stream(self.device)
        return self._frames

    def p3n(self) -> torch.Tensor:
        """This rank's P3_new shard [shard, mpad] in symmetric memory (allocated on first use, collectively)."""
        if getattr(self, "_p3n", None) is None:
            import torch.distributed._symmetric_memory as symm_mem
            self._p3n = symm_mem.empty((self.plan.shard, self.mpad), dtype=torch.float32, device=self.device)
            self._hdl_p3n = symm_mem.rendezvous(self._p3n, self.group.group_name)
        return self._p3n

    def peer_p3n(self, r: int) -> torch.Tensor:
        self.p3n()
        return self._hdl_p3n.get_buffer(r, (self.plan.shard, self.mpad), torch.float32)


class RankStep:
    """One rank's classic++ step, phase by phase.  Phases of different ranks are separated by barriers (real
    ranks) or by program order (virtual ranks on one device)."""

    def __init__(self, pb: PeerBuffers, frames: torch.Tensor, p: float = 0.7, alpha: float = 0.997,
                 max_ctas: int = 0, timing: bool = False):
        self.pb, self.frames, self.p, self.alpha, self.max_ctas = pb, frames, p, alpha, max_ctas
        x = frames.reshape(frames.shape[0], -1)
        if x.dtype != torch.uint8 or x.stride(1) != 1 or x.stride(0) % 16 != 0 or x.data_ptr() % 16 != 0:
            raise engine._lib.AvtexError("sharded path needs contiguous uint8 frames with 16-byte aligned rows")
        self.x = x
        self.parity = pb.calls % 2
        self.call = pb.calls
        pb.calls += 1
        self.marks = [] if timing else None
        self.res = None

    def mark(self, name):
        if self.marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    # -- K0: my slice of the norms, pushed to every rank's vector
    def norms(self):
        pb = self.pb
        n, k = self.x.shape
        per = -(-n // pb.world)
        lo, hi = pb.rank * per, min(n, (pb.rank + 1) * per)
        sq_off = pb.off_sq + self.parity * pb.npad * 8
        fl_off = pb.off_fl + self.parity * 16
        self.static_ok = (k + 127) // 128 * 128 * 128 * 128 < engine.GRAM_MAX_SQNORM
        # the other parity's max flag is idle now (last read at the end of the previous call, next pushed into
        # after this call's barriers): reset it for the next call
        pb.flags(1 - self.parity).zero_()
        if hi > lo:
            engine.frame_norms_push(self.x, lo, hi - lo, [b + sq_off for b in pb.small_ptrs],
                                    None if self.static_ok else [b + fl_off + 8 for b in pb.small_ptrs])
        self.pf = engine.PackedFrames(self.x, pb.sqnorm(self.parity)[:n], k, pb.flags(self.parity), signed=False)
        if self.static_ok:
            self.pf._checked = (True, "")

    # -- K1: my tiles, transposes pushed into the peers' shards
    def gram(self):
        pb = self.pb
        n, k = self.x.shape
        # the job list depends on the plans, the shard pointers and K only: built once per workspace (building the
        # dicts and the ctypes array took ~0.1 ms of host time per step, during which the GPU sat idle behind the
        # short norms kernel)
        cache = pb.__dict__.setdefault("_job_cache", {})
        arr = cache.get(k)
        if not pb.residues:
            if arr is None:
                arr = cache[k] = engine.gram_job_array(symmetric_jobs(pb.plans, pb.rank, pb.d1_ptrs, pb.ld, pb.stride))
            engine.gram_l2_jobs(self.pf, arr)
            return
        # one job list per residue-class plane (the same symmetric split, in class rows), all in ONE launch
        if k % 128 or self.x.stride(0) != k:
            raise engine._lib.AvtexError("residue-class shards need dense rows with K % 128 == 0")
        if arr is None:
            plane_bytes = pb.rows_c * pb.ld * 4
            jobs = []
            for r in range(pb.stride):
                for j in symmetric_jobs(pb.cplans, pb.rank, [b + r * plane_bytes for b in pb.d1_ptrs], pb.ld, 1):
                    jobs.append(dict(j, k_off=r * k, sq_off=r, sq_stride=pb.stride))
            arr = cache[k] = engine.gram_job_array(jobs)
        engine.gram_l2_jobs(self.pf, arr, n=n // pb.stride, ld=pb.stride * k)

    def halo(self):
        pb = self.pb
        if pb.residues:
            c = pb.cplans[pb.rank]
            mine = pb.d1.view(pb.stride, pb.rows_c, pb.ld)
            for owner, row, rows in halo_sources(pb.cplans, pb.rank, 1):
                for r in range(pb.stride):
                    src = pb.peer_plane_rows(owner, r, row - pb.cplans[owner].r_lo, rows)
                    mine[r, row - c.r_lo: row - c.r_lo + rows].copy_(src)
            self.D1 = pb.planes()
            return
        p = pb.plan
        for owner, row, rows in halo_sources(pb.plans, pb.rank, pb.stride):
            src = pb.peer_d1_rows(owner, row - pb.plans[owner].r_lo, rows)
            pb.d1[row - p.r_lo: row - p.r_lo + rows].copy_(src)
        self.D1 = pb.D1()

    # -- K2
    def filter(self):
        pb = self.pb
        p = pb.plan
        if pb.residues:
            self.D2, self.D3 = engine.diag_filter_residues(self.D1, pb.n, pb.fs, pb.stride, p=self.p, a0=p.a0,
                                                           rows_out=p.a1h - p.a0, in_row0=pb.cplans[pb.rank].r_lo,
                                                           symmetric=False)
            return
        self.D2, self.D3 = engine.diag_filter(self.D1, pb.fs, pb.stride, p=self.p, m=p.m, a0=p.a0,
                                              rows_out=p.a1h - p.a0, in_row0=p.r_lo)

    # -- K3: all sweeps + their exchanges in one cooperative kernel
    def future_cost(self):
        pb = self.pb
        p = pb.plan
        own = p.a1 - p.a0
        dev = pb.device
        arr = C.c_void_p * pb.world
        mptr = arr(*[b + pb.off_m + self.parity * 3 * pb.mpad * 4 for b in pb.small_ptrs])
        eptr = arr(*[b + pb.off_eps + self.parity * pb.n_eps for b in pb.small_ptrs])
        fptr = arr(*[b + pb.off_sweep for b in pb.small_ptrs])
        scratch = torch.zeros(2 * (pb.MAX_SWEEPS + 1), dtype=torch.float64, device=dev)
        eps_local, trail = scratch[:pb.MAX_SWEEPS + 1], scratch[pb.MAX_SWEEPS + 1:]
        info = torch.zeros(4, dtype=torch.int32, device=dev)
        self.m_out = torch.empty(pb.mpad, dtype=torch.float32, device=dev)
        epoch = (self.call * (pb.MAX_SWEEPS + 2)) & 0xFFFFFFFF
        D3_own = self.D3[:own]
        _lib.call("avtex_future_cost_fused_peer", _lib.ptr(D3_own), D3_own.stride(0), p.a0, own, p.m,
                  C.c_float(engine._f32(self.alpha)), C.c_float(np.float32(engine.F32_EPS_STOP)), pb.MAX_SWEEPS,
                  pb.rank, pb.world, mptr, pb.mpad, eptr, fptr, C.c_uint(epoch), _lib.ptr(eps_local),
                  _lib.ptr(trail), _lib.ptr(info), _lib.ptr(self.m_out), self.max_ctas, engine._dev(D3_own),
                  engine._stream(D3_own))
        self.fc = engine.FutureCostResult(self.m_out[:p.m], pending=(info, trail, p.m, None))

    # -- K4 (+ sigma statistics of my rows)
    def finalize(self, want_stats: bool):
        p = self.pb.plan
        own = p.a1 - p.a0
        self.stats = engine.new_stats(self.pb.device) if want_stats else None
        if want_stats and p.a1h > p.a1:
            # statistics cover the OWNED rows only: the halo row is finalized by a second, stat-less launch
            self.D3_new = engine.empty_matrix(p.a1h - p.a0, p.m, self.pb.device)
            _finalize_into(self.D3[:own], self.fc.mvec, self.alpha, p.a0, p.m, self.D3_new[:own], self.stats)
            _finalize_into(self.D3[own:], self.fc.mvec, self.alpha, p.a1, p.m, self.D3_new[own:], None)
        else:
            self.D3_new = engine.future_cost_finalize(self.D3, self.fc.mvec, self.alpha, row0=p.a0, m=p.m,
                                                      stats=self.stats)

    def result(self) -> "ShardResult":
        res = ShardResult(self.pb.plan, self.D1, self.D2, self.D3, self.D3_new, self.fc)
        res.launches = 1 + 1 + 1 + 1 + 1
        if self.marks is not None:
            torch.cuda.synchronize()
            res.stage_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks[:-1], self.marks[1:])}
        return res


def _finalize_into(D3, mvec, alpha, row0, m, out, stats):
    s, z = engine._stats_ptrs(stats)
    _lib.call("avtex_future_cost_finalize", _lib.ptr(D3), D3.stride(0), row0, D3.shape[0], m, _lib.ptr(mvec),
              C.c_float(engine._f32(alpha)), _lib.ptr(out), out.stride(0), s, z, engine._dev(D3), engine._stream(D3))


@dataclass
class ShardResult:
    plan: ShardPlan
    D1: torch.Tensor          # rows [r_lo, r_hi)
    D2: torch.Tensor          # rows [a0, a1h)
    D3: torch.Tensor
    D3_new: torch.Tensor      # rows [a0, a1h)
    fc: engine.FutureCostResult | None = None
    sigma: float | None = None
    P3: torch.Tensor | None = None        # rows [a0, a1)
    P3_new: torch.Tensor | None = None
    counts: torch.Tensor | None = None
    launches: int = 0
    stage_ms: dict | None = None          # CUDA-event stage times (timing=True / AVTEX_DIST_TIMING=1)

    @property
    def n_sweeps(self) -> int:            # reads the device on first use (one small D2H)
        return self.fc.n_sweeps

    @property
    def eps_trail(self) -> list:
        return self.fc.eps_trail


def _probabilities(res: ShardResult, stats: torch.Tensor, sigma_factor, threshold, Pn_out=None):
    own = res.plan.a1 - res.plan.a0
    res.sigma = engine.sigma_from_stats(*engine.read_stats(stats), sigma_factor)
    res.P3, res.P3_new, res.counts = engine.transition_probs(
        res.D3_new, res.sigma, shift=1, rows_out=own, threshold=threshold, want_counts=threshold is not None,
        Pn_out=Pn_out)
    res.launches += 1


def classic_sharded(frames: torch.Tensor, filter_size: int, stride: int, rank: int, world: int,
                    p: float = 0.7, alpha: float = 0.997, sigma_factor=None, threshold=None, group=None,
                    packed: engine.PackedFrames | None = None,
                    workspace: SymmetricShardWorkspace | None = None, timing: bool | None = None) -> ShardResult:
    """Distance + filter + converged future cost (+ sigma3 / P3 / P3_new when sigma_factor is given) for
    this rank's rows.  `frames`: the full [N, ...] uint8 clip on this rank's device.  With a
    SymmetricShardWorkspace every exchange happens inside the kernels (module docstring) and the host never
    synchronises before the step's results are read; without one every rank computes its full row block
    locally and the future cost exchanges through NCCL (`make_exchange`)."""
    import os
    if timing is None:
        timing = os.environ.get("AVTEX_DIST_TIMING") == "1"
    n = frames.shape[0]
    plan = plan_shards(n, filter_size, stride, world, rank)
    if workspace is not None and world > 1:
        st = RankStep(workspace, frames, p, alpha, timing=timing)
        st.mark("start")
        st.norms()
        st.mark("norms")
        workspace.barrier(0)               # norms of all ranks have landed; peers finished reading my old shard
        st.gram()
        workspace.barrier(1)               # all pushed tiles have landed
        st.halo()
        st.mark("gram")
        st.filter()
        st.mark("filter")
        st.future_cost()
        st.mark("future_cost")
        st.finalize(want_stats=sigma_factor is not None)
        st.mark("finalize")
        res = st.result()
        if not st.static_ok and not st.pf.exact_ok:       # deferred domain check (the Gram ran speculatively)
            raise engine._lib.AvtexError(f"sharded path needs byte frames inside the Gram domain: {st.pf.reason}")
        if sigma_factor is not None and res.fc._n_sweeps is None:
            # results are about to be consumed: settle convergence now (first host read of the step).  A clip
            # that needs more sweeps than the in-kernel cap finishes through the NCCL loop (all ranks take this
            # branch together: the stop decision is identical everywhere).
            try:
                res.fc._resolve()
            except RuntimeError as exc:
                if "did not converge" not in str(exc):
                    raise
                own = plan.a1 - plan.a0
                res.fc = engine.future_cost(st.D3[:own], alpha, row0=plan.a0, m=plan.m,
                                            exchange=make_exchange(plan, group), pad_to=plan.padded)
                st.fc = res.fc
                st.finalize(want_stats=True)
                res.D3_new = st.D3_new
        stats = st.stats
    else:
        marks = [] if timing else None

        def mark(name):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        pf = pack_frames_sharded(frames, rank, world, group) if packed is None else packed
        mark("norms")
        D1 = engine.gram_l2(pf, plan.r_lo, plan.r_hi - plan.r_lo, symmetric=False)
        mark("gram")
        D2, D3 = engine.diag_filter(D1, filter_size, stride, p=p, m=plan.m, a0=plan.a0,
                                    rows_out=plan.a1h - plan.a0, in_row0=plan.r_lo)
        own = plan.a1 - plan.a0
        mark("filter")
        fc = engine.future_cost(D3[:own], alpha, row0=plan.a0, m=plan.m, exchange=make_exchange(plan, group),
                                pad_to=plan.padded)
        mark("future_cost")
        stats = engine.new_stats(frames.device) if sigma_factor is not None else None
        if stats is not None and plan.a1h > plan.a1:
            D3_new = engine.empty_matrix(plan.a1h - plan.a0, plan.m, frames.device)
            _finalize_into(D3[:own], fc.mvec, alpha, plan.a0, plan.m, D3_new[:own], stats)
            _finalize_into(D3[own:], fc.mvec, alpha, plan.a1, plan.m, D3_new[own:], None)
        else:
            D3_new = engine.future_cost_finalize(D3, fc.mvec, alpha, row0=plan.a0, m=plan.m, stats=stats)
        mark("finalize")
        res = ShardResult(plan, D1, D2, D3, D3_new, fc)
        res.launches = 1 + 1 + 1 + fc.passes + 1
        if marks is not None:
            torch.cuda.synchronize()
            res.stage_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
        if not pf.exact_ok:
            raise engine._lib.AvtexError(f"sharded path needs byte frames inside the Gram domain: {pf.reason}")
    if sigma_factor is not None:
        if world > 1:
            stats = allreduce_stats(stats, group)
        # with a workspace P3_new goes into peer-mapped memory, so the host that runs the walk can read any rank's rows
        Pn_out = workspace.p3n() if (workspace is not None and world > 1 and threshold is not None) else None
        _probabilities(res, stats, sigma_factor, threshold, Pn_out)
    return res


def sharded_survivor_rows(res: ShardResult, workspace: "SymmetricShardWorkspace") -> engine.SurvivorRows:
    """Survivor lists over ALL ranks' P3_new shards, fetched on demand through peer-mapped memory (one-sided: the
    owning rank does nothing).  Call on the rank whose host runs the walk, after `workspace.barrier(2)` on every
    rank (all shards complete)."""
    plan = res.plan
    views = [workspace.peer_p3n(r) for r in range(plan.world)]

    def row(i):
        r = min(i // plan.shard, plan.world - 1)
        return views[r][i - r * plan.shard, :plan.m]

    return engine.SurvivorRows(plan.m, row)


def gather_survivors(res: ShardResult, group=None):
    """CSR of P3_new over all ranks, assembled on every rank (the walk runs on rank 0's host): row counts
    and column lists travel as two padded NCCL all-gathers on the device, then ONE D2H copy."""
    world = res.plan.world
    if world == 1:
        return engine.csr_from_matrix(res.P3_new, res.counts)
    dev = res.P3_new.device
    plan = res.plan
    own = plan.a1 - plan.a0
    counts = torch.zeros(plan.shard, dtype=torch.int32, device=dev)
    counts[:own] = res.counts
    all_counts = torch.empty(world * plan.shard, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    per_rank = all_counts.view(world, plan.shard).sum(1)
    cap = int(per_rank.max().item())                               # one small sync: the padded length
    rowptr = torch.zeros(own + 1, dtype=torch.int64, device=dev)
    torch.cumsum(res.counts, 0, out=rowptr[1:])
    colidx = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    _lib.call("avtex_csr_fill", _lib.ptr(res.P3_new), res.P3_new.stride(0), own, res.P3_new.shape[1],
              _lib.ptr(rowptr), _lib.ptr(colidx), engine._dev(res.P3_new), engine._stream(res.P3_new))
    all_cols = torch.empty(world * max(cap, 1), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_cols, colidx, group=group)
    h_counts = all_counts.cpu().numpy().reshape(world, plan.shard)
    h_cols = all_cols.cpu().numpy().reshape(world, max(cap, 1))
    counts_list, cols_list = [], []
    for r in range(world):
        a0 = r * plan.shard
        own_r = min(plan.m, a0 + plan.shard) - a0
        c = h_counts[r, :own_r]
        counts_list.append(c)
        cols_list.append(h_cols[r, :int(c.sum())])
    full_counts = np.concatenate(counts_list)
    full_ptr = np.concatenate(([0], np.cumsum(full_counts, dtype=np.int64))).astype(np.int64)
    return full_ptr, np.concatenate(cols_list)


# --------------------------------------------------------------------------- virtual ranks on one device
class _VirtualRankBuffers(PeerBuffers):
    def __init__(self, box: "VirtualBox", rank: int):
        super().__init__(box.n, box.fs, box.stride, rank, box.world, box.device, box.residues)
        self.box = box

    def barrier(self, channel: int):
        pass                                          # program order on one stream separates the phases

    def peer_d1_rows(self, r: int, row_off: int, rows: int) -> torch.Tensor:
        return self.box.ranks[r].d1[row_off:row_off + rows]


class VirtualBox:
    """G "virtual ranks" on ONE device: every rank has its own shard buffers (plain allocations), the peer
    pointers handed to the kernels are the other ranks' local buffers, and the phases of a step run rank by
    rank in program order.  The same job lists, halo plans, norm pushes and — with the G cooperative kernels
    co-resident on G streams — the same flag barrier as on a real 8-GPU box, so tests/test_gpu_virtual.py
    checks the G = 2 / 4 / 8 paths bit-for-bit against the single-GPU pipeline on a 1-GPU machine."""

    def __init__(self, n: int, filter_size: int, stride: int, world: int, device, residues: bool = False):
        self.n, self.fs, self.stride, self.world, self.device = n, filter_size, stride, world, device
        self.residues = residues
        self.ranks = [_VirtualRankBuffers(self, r) for r in range(world)]
        for pb in self.ranks:
            pb.d1 = torch.empty((pb.rows_max, pb.ld), dtype=torch.float32, device=device)
            pb.small = torch.zeros(pb.small_bytes, dtype=torch.uint8, device=device)
        for pb in self.ranks:
            pb.d1_ptrs = [q.d1.data_ptr() for q in self.ranks]
            pb.small_ptrs = [q.small.data_ptr() for q in self.ranks]
        self.streams = [torch.cuda.Stream(device) for _ in range(world)]

    def step(self, frames: torch.Tensor, sigma_factor=None, threshold=None, p: float = 0.7, alpha: float = 0.997):
        """Returns the list of per-rank ShardResults (sigma / P3 / P3_new filled when sigma_factor is given)."""
        steps = [RankStep(pb, frames, p, alpha, max_ctas=-self.world) for pb in self.ranks]
        for phase in ("norms", "gram", "halo", "filter"):
            for st in steps:
                getattr(st, phase)()
        # the G cooperative kernels spin on each other's flags: they must run CONCURRENTLY (one stream each,
        # grids capped so that all are resident)
        main = torch.cuda.current_stream(self.device)
        done = []
        for st, stream in zip(steps, self.streams):
            stream.wait_stream(main)
            with torch.cuda.stream(stream):
                st.future_cost()
                ev = torch.cuda.Event()
                ev.record(stream)
                done.append(ev)
        for ev in done:
            main.wait_event(ev)
        out = []
        for st in steps:
            st.finalize(want_stats=sigma_factor is not None)
            out.append(st.result())
        if sigma_factor is not None:
            total = torch.zeros(2, dtype=torch.float64, device=self.device)
            total[0] = sum(st.stats[0] for st in steps)
            total.view(torch.int64)[1] = sum(st.stats.view(torch.int64)[1] for st in steps)
            for res in out:
                _probabilities(res, total, sigma_factor, threshold)
        return out


def virtual_survivors(results: list):
    """CSR of P3_new assembled from the virtual ranks' shards."""
    # csr_from_matrix returns views of a small ring of page-locked buffers: copy, more shards than buffers may follow
    ptrs, cols = zip(*((rp.copy(), ci.copy()) for rp, ci in (engine.csr_from_matrix(r.P3_new, r.counts) for r in results)))
    counts = np.concatenate([np.diff(p) for p in ptrs])
    return np.concatenate(([0], np.cumsum(counts))).astype(np.int64), np.concatenate(cols)

#!/usr/bin/env python
"""bench.py — the classic++ transition-matrix hot path (distance + temporal filter + converged
future cost) on synthetic video, per BASELINE.json: frame-pairs/s at N frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c5]

N = 1   workload c2 = configs[1]: classic++ (-m 3, -fs 40, -stride 4) on a synthetic 5000-frame
        224x224 RGB clip, one B200.
N > 1   (torchrun, one rank per GPU) the same clip shape with N_frames = 5000*sqrt(N): per-GPU
        frame-pairs are constant ("weak"), rows sharded, all-gather of the per-row minima per sweep.
One "step" = pack (K0) -> tcgen05 Gram + L2 epilogue (K1) -> diagonal filter + pow (K2) -> future-cost
sweeps to convergence (K3, host reads eps each sweep) -> finalize (K4).  `value` is timed with the byte
frames resident in HBM; `e2e` goes through the reference-named entry points (compute_D1 / compute_D2 /
q_learning + the survivor lists for the walk) from PINNED HOST frames, copies inside the timed region.
`--impl reference` / `cpu_baseline` time the CPU oracle port of the reference algorithm on a bounded
sample (see cpu_reference()).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: frames, H, W, model_type, filter_size, stride
    "c1": dict(n=300, h=64, w=64, m=1, fs=40, stride=1),
    "c2": dict(n=5000, h=224, w=224, m=3, fs=40, stride=4),
    "c5": dict(n=100000, h=64, w=64, m=3, fs=40, stride=4),
    # not a BASELINE config: M = 19961 makes every M x M matrix 1.6 GB, far beyond L2 — used to
    # measure the HBM-bound kernels (filter, sweep, finalize, probabilities) against the HBM roofline
    "hbm": dict(n=20000, h=64, w=64, m=1, fs=40, stride=1),
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE gram_l2_s8_2cta_kernel launch at C2 from the committed
# ncu capture (profiles/r01_launches_c2_final.summary.txt: 1699 MB read + 73 MB written; operands are 753 MB,
# the 100 MB D1 mostly stays in L2)
GRAM_DRAM_BYTES_C2 = 1.771e9
METRIC = "frame-pairs/s (distance + temporal filter + converged future-cost)"
L2_FLUSH_BYTES = 256 << 20


def workload_name(wl):
    """The same string in both arms (`ours` and `--impl reference`)."""
    return (f"classic++ -m {wl['m']} -fs {wl['fs']} -stride {wl['stride']}: {wl['n']} frames "
            f"{wl['h']}x{wl['w']} RGB (K={wl['h'] * wl['w'] * 3})")


def scaled_workload(name, world):
    wl = dict(WORKLOADS[name])
    if world > 1 and name == "c2":
        wl["n"] = int(round(wl["n"] * math.sqrt(world) / 4)) * 4          # weak scaling: pairs per GPU fixed
    return wl


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], source="MEASURED_PEAKS.json (burst)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv is not None:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(wl, budget_s=20.0, D1_host=None, seed=0):
    """The reference's own algorithm on the host cores (oracle port; kind = "port").

    D1: the literal block algorithm of classic/computeD1.py:58-96 (repeat -> view -> torch.norm,
    bs = 48) on as many 48x48 blocks as fit in ~budget_s, extrapolated to all ceil(N/48)^2 blocks
    (each block costs the same).  D2 + future cost: in full on an N x N distance matrix (the GPU's D1
    when given, else a synthetic symmetric one), with the row minima computed once per sweep — the
    same arithmetic as the reference's O(M^3) loop, which could not finish at this size.
    Returns dict(value=pairs/s, seconds=..., sample=...).
    """
    from audio_video_textures_b200.synth import synth_video
    from oracle import classic as oc
    torch.set_num_threads(os.cpu_count())
    n, h, w, fs, stride = wl["n"], wl["h"], wl["w"], wl["fs"], wl["stride"]
    bs = 48
    sample_frames = synth_video(2 * bs, h, w, seed=seed).float()
    t0 = time.perf_counter()
    blocks = 0
    while True:
        _, done = oc.pairwise_l2_reference_blocks(sample_frames, bs, max_blocks=4)
        blocks += done
        if time.perf_counter() - t0 > budget_s * 0.6 or blocks >= 64:
            break
    t_block = (time.perf_counter() - t0) / blocks
    n_blocks = math.ceil(n / bs) ** 2
    t_d1 = t_block * n_blocks
    if D1_host is None:
        g = torch.Generator().manual_seed(seed)
        a = torch.rand(n, n, generator=g) * 1000.0
        D1_host = (a + a.T).fill_diagonal_(0.0)
    f = torch.tensor(4.5, dtype=torch.float32)
    t1 = time.perf_counter()
    D2 = oc.compute_D2(D1_host, f, fs, stride)[0]
    D3_new, trail = oc.future_cost(D2 ** 0.7)
    t_rest = time.perf_counter() - t1
    total = t_d1 + t_rest
    return dict(value=n * n / total, seconds=total, cores=os.cpu_count(), kind="port",
                sample=(f"D1: {blocks} of {n_blocks} 48x48 blocks of the reference block algorithm timed "
                        f"({t_block:.3f} s/block) and extrapolated ({t_d1:.0f} s); D2 + future cost "
                        f"({len(trail)} sweeps, vectorised row-min) in full at M={D2.shape[0]}: {t_rest:.2f} s"))


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch.distributed as dist

    from audio_video_textures_b200 import dist as avdist
    from audio_video_textures_b200 import engine
    from audio_video_textures_b200.synth import synth_video_cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = scaled_workload(args.workload, world)
    n, fs, stride = wl["n"], wl["fs"], wl["stride"]
    k = wl["h"] * wl["w"] * 3
    frames = synth_video_cuda(n, wl["h"], wl["w"], seed=0, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    peaks = load_peaks()
    gram_ms, step_ms = [], []
    state = {}
    workspace = peer_fc = None
    comm_note = ""
    if world > 1 and not args.no_symmetric:
        try:
            workspace = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev)
            peer_fc = avdist.PeerFutureCost(avdist.plan_shards(n, fs, stride, world, rank).m, rank, world, dev)
            ok = torch.ones(1, device=dev)
        except Exception as exc:                      # no peer-mapped memory on this box: NCCL path of the same algorithm
            workspace = peer_fc = None
            ok = torch.zeros(1, device=dev)
            comm_note = f"symmetric memory unavailable ({type(exc).__name__}): NCCL all-gather path"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            workspace = peer_fc = None

    def one_step(timed: bool):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        pf = engine.pack_frames(frames) if world == 1 else avdist.pack_frames_sharded(frames, rank, world)
        if world == 1:
            stats = engine.new_stats(dev)
            ev[1].record()
            D1 = engine.gram_l2(pf, stats=stats)
            ev[2].record()
            D2, D3 = engine.diag_filter(D1, fs, stride, p=0.7)
            fc = engine.future_cost_fused(D3, 0.997)
            D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997)
            launches = 1 + 1 + 1 + 1 + 1
            if not pf.exact_ok:
                raise RuntimeError(pf.reason)
            state.update(D1=D1, D3n=D3n, sweeps=fc.n_sweeps, m=D3.shape[0], rows=n)
        else:
            ev[1].record()
            res = avdist.classic_sharded(frames, fs, stride, rank, world, packed=pf, workspace=workspace,
                                         peer_fc=peer_fc)
            ev[2].record()       # (gram is the first kernel after ev[1]; the sharded call is timed as a whole)
            launches = res.launches
            if res.stage_ms is not None and rank == 0 and timed:
                print("stage_ms", {k: round(v, 3) for k, v in res.stage_ms.items()}, file=sys.stderr)
            state.update(D3n=res.D3_new, sweeps=res.n_sweeps, m=res.plan.m, rows=res.plan.r_hi - res.plan.r_lo)
        ev[3].record()
        torch.cuda.synchronize()
        if timed:
            step_ms.append(ev[0].elapsed_time(ev[3]))
            if world == 1:
                gram_ms.append(ev[1].elapsed_time(ev[2]))
        return launches

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        one_step(False)
        flush.fill_(1)
    sync_all()
    launches = 0
    wall0 = time.perf_counter()
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            flush.fill_(1)                      # L2 flush (256 MB > 126 MB L2) between timed steps
            torch.cuda.synchronize()
            launches += one_step(True)
    sync_all()
    wall = time.perf_counter() - wall0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    value = n * n * args.steps / total_s

    # ---- end to end through the reference-named entry points, host frames in pinned memory
    e2e = None
    if world == 1:
        from audio_video_textures_b200.classic.computeD1 import compute_D1
        from audio_video_textures_b200.classic.computeD2 import compute_D2
        from audio_video_textures_b200.classic.q_learning import q_learning
        host = frames.cpu().pin_memory()
        f = torch.tensor(4.5, dtype=torch.float32)
        times, d2h = [], 0
        for it in range(args.warmup + max(3, args.steps // 4)):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                D1, P1, s1 = compute_D1(host, f, "RGB", slow=True, batch_size=48)
                if wl["m"] in (1, 2):
                    D2, P2, s2, _ = compute_D2(D1, f, filter_size=fs)
                else:
                    D2, P2, s2, _ = compute_D2(D1, f, filter_size=fs, stride=stride)
                D3n, P3, P3n, s3 = q_learning(D2, f, thresholding=0.08)
            rowptr, colidx = engine.csr_from_matrix(P3n)            # what the walk consumes (D2H)
            sig = s3.item()
            torch.cuda.synchronize()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
            d2h = rowptr.nbytes + colidx.nbytes + 3 * 4
        e2e = {"value": n * n / float(np.mean(times)), "unit": "frame-pairs/s",
               "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * float(np.mean(times)),
               "includes": "compute_D1+compute_D2+q_learning (P1,P2,P3,P3_new, sigmas) + survivor CSR D2H"}

    elif world > 1:
        # every rank copies the (replicated) clip from its own pinned host buffer, then runs its shard
        host = frames.reshape(n, -1).cpu().pin_memory()
        times = []
        for it in range(args.warmup + max(3, args.steps // 4)):
            flush.fill_(1)
            sync_all()
            t0 = time.perf_counter()
            dev_frames = avdist.load_frames_sharded(host, rank, world, dev)     # 1/G over PCIe + NVLink all-gather
            res = avdist.classic_sharded(dev_frames, fs, stride, rank, world, workspace=workspace, peer_fc=peer_fc)
            sync_all()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": n * n / float(t.item()), "unit": "frame-pairs/s", "h2d_bytes_per_step": int(host.numel()),
               "d2h_bytes_per_step": 8 * world, "ms_per_step": 1e3 * float(t.item()),
               "includes": "each rank copies 1/G of the pinned host clip, NVLink all-gather replicates it, then the "
                           "sharded distance/filter/future-cost; D3_new shards stay on device"}

    if rank != 0:
        return
    m = state["m"]
    out = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None,
        "dtype": "u8 (exact int32 tensor-core Gram) + fp32", "data": "synthetic",
        "config": {"workload": workload_name(wl), "name": args.workload,
                   "detail": f"M={m}, {state['sweeps']} future-cost sweeps to eps <= 0.01",
                   "l2": "256 MB L2 flush between timed steps",
                   "sharding": "single GPU" if world == 1 else
                   (f"rows over {world} ranks, N=5000*sqrt(G)" if args.workload == "c2" else f"rows over {world} ranks") +
                   ("" if world == 1 or args.no_symmetric else "; symmetric Gram, transposed tiles pushed to peer shards over NVLink; future cost "
                                                                 "fused with its all-gather (peer stores + flag barrier)")},
        "gpu_launches": launches, "wall_s": wall,
    }
    if world == 1:
        g_ms = float(np.mean(gram_ms))
        flops = 2.0 * k * n * n
        ach = flops / (g_ms * 1e-3) / 1e12
        out["roofline"] = {
            "kernel": "gram_l2_s8_2cta_kernel (tcgen05 kind::i8, cta_group::2, TMA-fed, symmetric tile schedule)", "bound": "tensor",
            "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "traffic": GRAM_DRAM_BYTES_C2 if args.workload == "c2" else None, "ms": g_ms,
            "executed_tops": 0.5 * ach * (1.0 + 1.0 / math.ceil(n / 256)),     # upper-triangle 256x256 tiles only
            "peak_i8_nominal_tops": 4500.0, "share_of_step": g_ms * len(gram_ms) / sum(step_ms),
            "note": ("algorithmic flops 2*K*N^2 over the event-timed launch; the symmetric schedule executes "
                     "~half of them and kind::i8 runs at twice the bf16 rate, so frac is quoted against the "
                     f"measured bf16 peak from {peaks['source']} and can exceed 1")}
        D1_host = state["D1"].cpu() if n <= 8000 else None
        out["cpu_baseline"] = {kk: vv for kk, vv in cpu_reference(wl, args.cpu_budget, D1_host).items()
                               if kk != "seconds"}
        out["cpu_baseline"]["unit"] = "frame-pairs/s"
    out["e2e"] = e2e
    if comm_note:
        out["config"]["sharding"] += "; " + comm_note
    out["clocks"] = clocks.summary()
    print(json.dumps(out))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = scaled_workload(args.workload, args.gpus)        # the same clip the GPU arm uses at this --gpus
    vals = []
    info = None
    for it in range(args.warmup + args.steps):
        info = cpu_reference(wl, budget_s=max(4.0, args.cpu_budget / max(1, args.steps)), seed=it)
        if it >= args.warmup:
            vals.append(info["value"])
    v = float(np.mean(vals))
    k = wl["h"] * wl["w"] * 3
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wl["n"] ** 2 / v,
           "higher_is_better": True, "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None,
           "dtype": "fp32", "data": "synthetic",
           "config": {"workload": workload_name(wl), "name": args.workload},
           "cpu_baseline": {"value": v, "unit": "frame-pairs/s", "cores": info["cores"], "kind": "port",
                            "sample": info["sample"]},
           "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no_symmetric", action="store_true",
                    help="N>1: plain row shards (every rank computes its full row block) instead of peer pushes")
    ap.add_argument("--cpu_budget", type=float, default=20.0, help="seconds of CPU work for the baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.warmup + args.steps > 6:          # keep the CPU arm within minutes
            args.warmup, args.steps = min(args.warmup, 1), min(args.steps, 3)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

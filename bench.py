#!/usr/bin/env python
"""bench.py — the classic++ transition-matrix hot path (distance + temporal filter + converged
future cost) on synthetic video, per BASELINE.json: frame-pairs/s at N frames; synth frames/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c5|hbm]

Main line
  N = 1   workload c2 = configs[1]: classic++ (-m 3, -fs 40, -stride 4) on a synthetic 5000-frame
          224x224 RGB clip, one B200.
  N > 1   (torchrun, one rank per GPU) the same clip shape with N_frames = 5000*sqrt(N): per-GPU
          frame-pairs are constant ("weak"); rows sharded; every exchange (norms, transposed Gram tiles,
          per-sweep row minima) is a peer store from inside the kernels (dist.py).
  One "step" = norms (K0) -> tcgen05 Gram + L2 epilogue (K1) -> diagonal filter + pow (K2) -> all future-cost
  sweeps in one cooperative kernel (K3) -> finalize (K4).  `value` is timed with the byte frames resident in
  HBM: at N = 1 the K steps are enqueued back to back between ONE synchronize on each side, every step with its
  own CUDA-event pair on the stream and a 256 MB L2 flush in front of it (outside the pair); at N > 1 the ranks are
  re-aligned (synchronize + barrier) before every step so that one rank's flush is not charged to its peers; at N = 1 a step is one
  CUDA-graph replay (engine.PipelineGraph, K0 fused into the Gram launch), the eager launches are reported as
  extra.eager.  `e2e` goes through the reference-named entry points from PINNED HOST frames and includes sigma3 / P3 /
  P3_new, the survivor lists the walk needs copied back to the host and the 900-frame walk itself — the SAME
  scope at every N.
Extra records on the same JSON line (all measured in this run)
  roofline      the Gram kernel: executed int8 ops against an int8 peak MEASURED here (cuBLASLt int8 GEMM and a
                long-K run of the kernel itself), algorithmic ops separately, the SM clock measured INSIDE the
                kernel (clock64 / globaltimer), DRAM traffic from the committed ncu launch list
  roofline_hbm  (N = 1) the HBM-bound kernels at M = 19961 (1.6 GB matrices, far beyond L2)
  extra.c5      configs[4], the north-star size: 100000 frames 64x64, -m 3, at THIS N (strong scaling from
                the N = 1 run of the same clip), with per-stage times (max over ranks) and its own e2e
  extra.parity  (N > 1) shards of the row-sharded pipeline == the single-GPU pipeline on a 3000-frame clip
  extra.residue_pipeline   the same pass through the residue-class pipeline (stride 4: the tensor cores evaluate 1/8
                of the pairs, D2 / D3 / D3_new bit-identical, no D1) — beside the headline, never instead of it
  extra.synth   (N = 1) `synth frames/s`: the classic walk over the survivor lists, and contrastive synthesis
                C3 / C4 (20000 windows, D = 2304, A = 128 / 12288) ms per step and frames/s
  stages        (N > 1) norms / gram / filter / future_cost / finalize of the main workload, max over ranks
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
METRIC = "frame-pairs/s (distance + temporal filter + converged future-cost)"
L2_FLUSH_BYTES = 256 << 20
NOMINAL_I8_TOPS = 4500.0            # dense int8 at the 1965 MHz boost clock
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_gram_traffic.json")


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
        self.power = []
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
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

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
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(wl, budget_s=20.0, D1_host=None, seed=0):
    """The reference's own algorithm on the host cores (oracle port; kind = "port"), one bounded sample.

    D1: the literal block algorithm of classic/computeD1.py:58-96 (repeat -> view -> torch.norm, bs = 48) on
    as many 48x48 blocks as fit in ~0.6*budget_s; every block of the full clip costs the same, so the
    sample's rate carries over.  D2 + future cost: in full on an N x N distance matrix (the GPU's D1 when
    given, else a synthetic symmetric one) with the row minima computed once per sweep — the same arithmetic
    as the reference's O(M^3) loop, which could not finish at this size; its time is charged to the sample
    in proportion to the sampled pairs.  Returns dict(value = sampled pairs / their time, wall_s = what this
    call really took, seconds_full = implied time of one full pass, ...).
    """
    from audio_video_textures_b200.synth import synth_video
    from oracle import classic as oc
    torch.set_num_threads(os.cpu_count())
    n, h, w, fs, stride = wl["n"], wl["h"], wl["w"], wl["fs"], wl["stride"]
    bs = 48
    wall0 = time.perf_counter()
    sample_frames = synth_video(2 * bs, h, w, seed=seed).float()
    t0 = time.perf_counter()
    blocks = 0
    while True:
        _, done = oc.pairwise_l2_reference_blocks(sample_frames, bs, max_blocks=4)
        blocks += done
        if time.perf_counter() - t0 > budget_s * 0.6 or blocks >= 64:
            break
    t_blocks = time.perf_counter() - t0
    n_blocks = math.ceil(n / bs) ** 2
    if D1_host is None:
        g = torch.Generator().manual_seed(seed)
        a = torch.rand(n, n, generator=g) * 1000.0
        D1_host = (a + a.T).fill_diagonal_(0.0)
    f = torch.tensor(4.5, dtype=torch.float32)
    t1 = time.perf_counter()
    D2 = oc.compute_D2(D1_host, f, fs, stride)[0]
    D3_new, trail = oc.future_cost(D2 ** 0.7)
    t_rest = time.perf_counter() - t1
    frac = blocks / n_blocks                                  # share of the clip's pairs the sample covers
    sample_pairs = frac * n * n
    sample_s = t_blocks + t_rest * frac
    return dict(value=sample_pairs / sample_s, seconds_full=sample_s / frac, wall_s=time.perf_counter() - wall0,
                cores=os.cpu_count(), kind="port",
                sample=(f"D1: {blocks} of {n_blocks} 48x48 blocks of the reference block algorithm "
                        f"({t_blocks / blocks:.3f} s/block, every block costs the same) + the matching share of "
                        f"D2 + future cost ({len(trail)} sweeps, vectorised row-min, run in full at M={D2.shape[0]}: "
                        f"{t_rest:.2f} s); one full pass would take {sample_s / frac:.0f} s"))


# ------------------------------------------------------------------------------------ helpers (GPU arm)
def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def measure_i8_peak(dev):
    """An int8 tensor-core peak measured on THIS box, two ways: cuBLASLt's int8 GEMM (torch._int_mm,
    8192^3) and a long-K run of the repo's own Gram kernel (296 full 256x256 tiles = 4 per CTA pair,
    K = 65536: the epilogue is 0.4 % of the tile time).  Best of 10 each, CUDA events."""
    from audio_video_textures_b200 import engine
    out = {}
    try:
        a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
        best = 1e9
        for i in range(13):
            ev = _events(2)
            ev[0].record()
            torch._int_mm(a, b)
            ev[1].record()
            torch.cuda.synchronize()
            if i >= 3:
                best = min(best, ev[0].elapsed_time(ev[1]))
        out["cublaslt_int8_8192_tops"] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    except Exception as exc:                                   # pragma: no cover - depends on the torch build
        out["cublaslt_int8_8192_tops"] = None
        out["cublaslt_note"] = f"torch._int_mm unavailable: {type(exc).__name__}"
    n, k, rows = 9472, 65536, 2048
    x = torch.randint(0, 255, (n, k), dtype=torch.uint8, device=dev)
    pf = engine.pack_frames(x)
    D = engine.empty_matrix(rows, n, dev)
    job = [dict(row0=0, rows=rows, col0=0, cols=n, symmetric=0, count_stats=0, D=D.data_ptr(), d_row0=0,
                ldd=D.stride(0))]
    best = 1e9
    for i in range(13):
        ev = _events(2)
        ev[0].record()
        engine.gram_l2_jobs(pf, job)
        ev[1].record()
        torch.cuda.synchronize()
        if i >= 3:
            best = min(best, ev[0].elapsed_time(ev[1]))
    out["own_kernel_longK_tops"] = 2.0 * rows * n * k / (best * 1e-3) / 1e12
    out["own_kernel_longK_shape"] = f"{rows}x{n} outputs, K={k} (296 tiles of 256x256)"
    del x, pf, D
    vals = [v for v in (out["cublaslt_int8_8192_tops"], out["own_kernel_longK_tops"]) if v]
    out["peak_tops"] = max(vals)
    out["peak_source"] = ("cuBLASLt int8 8192^3" if out["peak_tops"] == out.get("cublaslt_int8_8192_tops")
                          else "own kernel, long K")
    return out


def gram_roofline(frames, wl, gram_ms, step_share, dev, peaks):
    """The tensor-bound kernel's line: executed and algorithmic int8 ops of ONE launch over its event-timed
    duration, against the int8 peak measured here; in-kernel SM clock; DRAM traffic of the committed capture."""
    from audio_video_textures_b200 import engine
    n = wl["n"]
    k = wl["h"] * wl["w"] * 3
    kp = (k + 127) // 128 * 128
    tiles_1d = math.ceil(n / 256)
    tiles = tiles_1d * (tiles_1d + 1) // 2                                  # upper-triangle 256 x 256 tiles
    executed = tiles * 256.0 * 256.0 * kp * 2.0
    algorithmic = 2.0 * k * n * n
    t = gram_ms * 1e-3
    # SM clock inside the kernel: clock64 / globaltimer deltas of CTA 0's tile loop (median of 5 launches)
    pf = engine.pack_frames(frames)
    D1 = engine.empty_matrix(n, n, dev)
    probe = torch.zeros(2, dtype=torch.int64, device=dev)
    job = [dict(row0=0, rows=n, col0=0, cols=n, symmetric=1, count_stats=0, D=D1.data_ptr(), d_row0=0,
                ldd=D1.stride(0), DT=D1.data_ptr(), dt_row0=0, ldt=D1.stride(0))]
    mhz = []
    for _ in range(5):
        engine.gram_l2_jobs(pf, job, clock_probe=probe)
        cyc, ns = (int(v) for v in probe.cpu())
        if ns > 0:
            mhz.append(1e3 * cyc / ns)
    del D1
    i8 = measure_i8_peak(dev)
    traffic = None
    traffic_note = "no committed capture for this workload"
    if os.path.exists(TRAFFIC_FILE):
        with open(TRAFFIC_FILE) as f:
            tr = json.load(f)
        if tr.get("workload") == workload_name(wl):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            traffic_note = tr.get("source", "")
    ex_tops = executed / t / 1e12
    clk = float(np.median(mhz)) if mhz else None
    return {
        "kernel": "gram_l2_s8_2cta_kernel (tcgen05 kind::i8, cta_group::2, TMA-fed, symmetric tile schedule)",
        "bound": "tensor", "unit": "TOP/s (int8)", "ms": gram_ms,
        "achieved": ex_tops, "peak": i8["peak_tops"], "frac": ex_tops / i8["peak_tops"],
        "peak_source": f"measured in this run: {i8['peak_source']}", "i8_peak_measurements": i8,
        "executed_ops": executed, "algorithmic_ops": algorithmic, "algorithmic_tops": algorithmic / t / 1e12,
        "tiles_executed": tiles, "tiles_full_matrix": tiles_1d * tiles_1d,
        "sm_mhz_in_kernel": clk,
        "frac_of_nominal_at_measured_clock": (ex_tops / (NOMINAL_I8_TOPS * clk / 1965.0)) if clk else None,
        "frac_of_nominal_i8_4500": ex_tops / NOMINAL_I8_TOPS,
        "bf16_peak_measured_tflops": peaks["bf16_tflops"],
        "traffic": traffic, "traffic_source": traffic_note, "operand_bytes": float(n) * k,
        "share_of_step": step_share,
        "note": ("achieved = executed int8 ops (upper-triangle tiles, K padded to 128) / event-timed launch; peak = the "
                 "larger of the two int8 rates measured in this run; algorithmic = 2*K*N^2 reported separately; "
                 "sm_mhz_in_kernel = clock64/globaltimer inside the kernel (NVML's sampling cannot see a ~1 ms launch). "
                 "The measured peak is this same tensor pipe at the same power-limited clock, so frac sits at 1.00 +- 0.01 "
                 "and says 'no better int8 rate was obtainable on this box' (cuBLASLt's is lower); the distance to the "
                 "NOMINAL peak is frac_of_nominal_i8_4500 (clock) and frac_of_nominal_at_measured_clock (pipe occupancy)")}


def hbm_rooflines(dev, peaks):
    """The HBM-bound kernels at M = 19961 (every matrix 1.6 GB): algorithmic bytes / event-timed launch against
    the measured copy bandwidth.  L2 (126 MB) is irrelevant at this size; median of 5 launches each."""
    from audio_video_textures_b200 import engine
    from audio_video_textures_b200.synth import synth_video_cuda
    wl = WORKLOADS["hbm"]
    n, fs = wl["n"], wl["fs"]
    k = wl["h"] * wl["w"] * 3
    frames = synth_video_cuda(n, wl["h"], wl["w"], seed=1, device=dev)
    out = []

    def timed(fn, reps=5):
        ms = []
        for _ in range(reps + 1):
            ev = _events(2)
            ev[0].record()
            r = fn()
            ev[1].record()
            torch.cuda.synchronize()
            ms.append(ev[0].elapsed_time(ev[1]))
        return float(np.median(ms[1:])), r

    def line(kernel, nbytes, ms, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "ms": ms, "algorithmic_bytes": nbytes, "what": what})

    x = frames.reshape(n, -1)
    sq = torch.empty(n, dtype=torch.int64, device=dev)
    fl = torch.zeros(2, dtype=torch.int64, device=dev)
    ms, _ = timed(lambda: engine.frame_norms_rows(x, 0, n, sq, fl))
    line("frame_norms_u8_kernel (12 KB rows)", float(n) * k, ms, f"K0: {n} rows of {k} B read once")
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf)
    m = engine.filtered_size(n, fs, 1)
    ms, (D2, D3) = timed(lambda: engine.diag_filter(D1, fs, 1, p=0.7, symmetric=False))
    line("diag_filter_kernel<40,1,16,general> (stride 1; FP32-pipe bound, see DESIGN 4.2)", 4.0 * n * n + 8.0 * m * m, ms,
         f"K2 -m 1/2, any D1: read D1 {n}^2, write D2 + D3 {m}^2")
    del D2, D3
    ms, (D2, D3) = timed(lambda: engine.diag_filter(D1, fs, 1, p=0.7, symmetric=True))
    line("diag_filter_kernel<40,1,16,symmetric> (stride 1; the form compute_D2 uses on compute_D1's matrix)",
         2.0 * n * n + 8.0 * m * m, ms,
         f"K2 -m 1/2, symmetric D1: read the upper triangle of D1 {n}^2, write D2 + D3 {m}^2 "
         f"(= {(4.0 * n * n + 8.0 * m * m) / (ms * 1e-3) / 1e9 / peaks['hbm_gbs']:.3f} of HBM counted in the general kernel's bytes)")
    del D2
    mv = torch.zeros((m + 31) // 32 * 32, dtype=torch.float32, device=dev)
    out_m = torch.empty_like(mv)
    import ctypes as C
    from audio_video_textures_b200 import _lib

    def sweep():
        _lib.call("avtex_future_cost_sweep", _lib.ptr(D3), D3.stride(0), 0, m, m, _lib.ptr(mv), None,
                  C.c_float(0.997), _lib.ptr(out_m), None, engine._dev(D3), engine._stream(D3))
    ms, _ = timed(sweep)
    line("future_cost_sweep_kernel", 4.0 * m * m, ms, f"K3: one sweep = one read of D3 {m}^2")
    ms, D3n = timed(lambda: engine.future_cost_finalize(D3, mv[:m]))
    line("future_cost_finalize_kernel", 8.0 * m * m, ms, "K4: read D3, write D3_new")
    del D3
    ms, _ = timed(lambda: engine.transition_probs(D3n, 3000.0, threshold=0.08, want_counts=True))
    line("transition_probs_kernel<cache,1024>", 12.0 * m * m, ms, "K5: read D3_new once, write P3 and P3_new")
    del D3n, D1
    # stride-4 filter on the C5 row-shard shape (what one of 8 ranks runs at N = 100000): 12539 D1 rows x 100000
    n5, rows5 = 100000, 12536 + 40
    D1s = torch.empty((rows5, n5), dtype=torch.float32, device=dev).uniform_(1.0, 2.0)
    m5 = engine.filtered_size(n5, fs, 4)
    rows_out = (rows5 - fs) // 4 + 1
    ms, _ = timed(lambda: engine.diag_filter(D1s, fs, 4, p=0.7, m=m5, a0=0, rows_out=rows_out, in_row0=0))
    line("diag_filter_kernel<40,4,8> (stride 4, C5 shard)", 4.0 * rows5 * n5 + 8.0 * rows_out * m5, ms,
         f"K2 -m 3: read {rows5} x {n5} D1 rows, write D2 + D3 {rows_out} x {m5}")
    return out


def synth_records(dev, state_c2):
    """BASELINE.json metric, second half: synth frames/s.  Classic: the sampling walk over the survivor lists
    (host loop, numpy legacy RNG; classic/video_textures.py:43-209).  Contrastive: the -e synthesis loop at the
    embedding boundary for C3 / C4 (cvt/validate.py:324-572), 30 s of video at 30 fps."""
    from audio_video_textures_b200 import engine
    from audio_video_textures_b200.classic.video_textures import texture_walk
    from audio_video_textures_b200.contrastive.validate import synthesize
    from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings
    rec = {}
    P3n, counts = state_c2["P3n"], state_c2["counts"]
    t0 = time.perf_counter()
    rowptr, colidx = engine.csr_from_matrix(P3n, counts)
    csr_s = time.perf_counter() - t0
    walks = {}
    for mode in (3, 1):
        np.random.seed(0)
        t0 = time.perf_counter()
        frames_list, jumps = texture_walk((rowptr, colidx), mode, 30, 30, 4, 40)
        dt = time.perf_counter() - t0
        walks[f"m{mode}"] = {"frames": len(frames_list), "frames_per_s": len(frames_list) / (dt + csr_s),
                             "walk_ms": 1e3 * dt, "jump_count": int(jumps)}
    rec["classic_walk_c2"] = {"survivor_csr_ms": 1e3 * csr_s, "nnz": int(rowptr[-1]), **walks,
                              "note": "frames/s = emitted frames / (GPU survivor compaction + D2H + host walk); -nvl 30 at 30 fps"}
    L, D = 20000, 2304
    emb = synth_embeddings(L, D, seed=0, device=dev)
    for name, A in (("c3", 0), ("c4_a128", 128), ("c4_a12288", 12288)):
        kw = {}
        if A:
            kw = dict(alpha=0.5, q_audio=synth_audio_features(L, A, seed=0, device=dev),
                      da_source=synth_audio_features(L, A, seed=1, device=dev),
                      da_driving=synth_audio_features(160, A, seed=2, device=dev))
        args = dict(temp=0.1, threshold=0.3, fps=30, new_video_length=30, window=15, stride=6)
        np.random.seed(0)
        synthesize(emb, **args, **kw)                                      # warm-up (tables, first launches)
        torch.cuda.synchronize()
        best = None
        for rep in range(3):
            np.random.seed(0)
            t0 = time.perf_counter()
            res = synthesize(emb, **args, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        steps = len(res["q_ids"])
        # loop only: table normalisation excluded by timing the steps through a prepared state
        from audio_video_textures_b200.contrastive.validate import SynthesisState
        st = SynthesisState(emb, None, kw.get("q_audio"), None, kw.get("da_source"), kw.get("da_driving"))
        torch.cuda.synchronize()
        np.random.seed(0)
        q, t0 = res["start"], time.perf_counter()
        for it in range(1, steps + 1):
            ch = st.step(q, it, 0.1, 0.5, 0.3)
            q = int(np.random.choice(ch))
        per_step_launch_s = time.perf_counter() - t0
        loop_s = 1e9
        for rep in range(3):                                             # the product path: the whole loop in one kernel
            np.random.seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            engine.synthesis_loop(st.ws, st.tn, st.qn, res["start"], steps, 0.1, 0.5, 0.3, st.sn, st.dn)
            loop_s = min(loop_s, time.perf_counter() - t0)
        row_bytes = 4.0 * L * (D + A) + (4.0 * L * A if A else 0.0)
        rec[name] = {"config": f"contrastive synthesis -e -th 0.3 -temp 0.1{' -m 2 -alpha 0.5' if A else ''}: L={L} D={D} A={A}",
                     "steps": steps, "frames": len(res["frame_ids"]), "ms_per_step": 1e3 * loop_s / steps,
                     "frames_per_s": len(res["frame_ids"]) / loop_s, "window_pairs_per_s": steps * L / loop_s,
                     "ms_per_step_incl_table_setup": 1e3 * best / steps,
                     "ms_per_step_one_launch_per_step_host_draw": 1e3 * per_step_launch_s / steps,
                     "how": "whole loop in one persistent cooperative kernel, np.random.choice drawn on the device from numpy's MT19937 state",
                     "hbm_bytes_per_step": row_bytes, "hbm_frac_of_step": row_bytes / (loop_s / steps) / 1e9 / load_peaks()["hbm_gbs"],
                     "launches_per_loop": 1}
        del kw, st
    return rec


def _walk_sharded(avdist, engine, res, workspace, wl, rank):
    """The walk of the e2e legs at N > 1: rank 0's host draws from survivor lists fetched on demand out of the
    owning ranks' P3_new shards (peer-mapped symmetric memory); without a workspace the lists are all-gathered.
    Returns the bytes copied device -> host on rank 0."""
    from audio_video_textures_b200.classic.video_textures import texture_walk
    if workspace is None:
        rowptr, colidx = avdist.gather_survivors(res)
        if rank == 0:
            np.random.seed(0)
            texture_walk((rowptr, colidx), wl["m"], 30, 30, wl["stride"], wl["fs"])
        return rowptr.nbytes + colidx.nbytes
    workspace.barrier(2)                                   # every rank's P3_new shard is complete
    if rank != 0:
        return 0
    rows = avdist.sharded_survivor_rows(res, workspace)
    np.random.seed(0)
    texture_walk(rows, wl["m"], 30, 30, wl["stride"], wl["fs"])
    return int(sum(v.nbytes for v in rows._cache.values()))


def residue_step_record(frames, fs, stride, flush, steps=5, warm=2):
    """The same pass (norms, distances, filter, converged future cost, finalize) through the RESIDUE-CLASS pipeline
    (engine.distance_filter): a stride-s filter reads D1[i,j] only where i = j (mod s), so K1 computes the s class
    Gram matrices (1/s of the pairs, one launch) and K2 walks the planes; D2 / D3 / D3_new are bit-identical to the
    full-D1 pass (tests/test_gpu_classic.py::test_residue_class_pipeline_is_bit_identical), D1 / P1 are not produced.
    Reported NEXT to the headline, never instead of it."""
    from audio_video_textures_b200 import engine
    names = ["norms", "gram_residues", "filter", "future_cost", "finalize"]

    def one():
        ev = _events(6)
        ev[0].record()
        pf = engine.pack_frames(frames)
        ev[1].record()
        D1r = engine.gram_l2_residues(pf, stride)
        ev[2].record()
        D2, D3 = engine.diag_filter_residues(D1r, frames.shape[0], fs, stride, p=0.7)
        ev[3].record()
        fc = engine.future_cost_fused(D3, 0.997)
        ev[4].record()
        engine.future_cost_finalize(D3, fc.mvec, 0.997)
        ev[5].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[5]), [ev[i].elapsed_time(ev[i + 1]) for i in range(5)], fc.n_sweeps

    pf = engine.pack_frames(frames)
    if not (engine.residue_eligible(pf, fs, stride) and pf.exact_ok):
        return {"eligible": False}
    for _ in range(warm):
        one()
    ms, st = [], []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t, s_, sweeps = one()
        ms.append(t)
        st.append(s_)
    n = frames.shape[0]
    eager = float(np.mean(ms))
    step, launch = eager, "eager launches"
    try:                                                     # the same pass as one CUDA-graph replay
        g = engine.PipelineGraph(n, frames[0].numel(), fs, stride, residues=True, device=frames.device)
        g.frames.copy_(frames.reshape(n, -1))
        gm = []
        for it in range(warm + steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ev = _events(2)
            ev[0].record()
            g()
            ev[1].record()
            torch.cuda.synchronize()
            if it >= warm:
                gm.append(ev[0].elapsed_time(ev[1]))
        step, launch = float(np.mean(gm)), "one CUDA-graph replay per step"
        del g
    except Exception as exc:
        launch = f"eager launches (graph capture failed: {type(exc).__name__}: {exc})"
    return {"eligible": True, "ms_per_step": step, "launch": launch, "ms_per_step_eager": eager,
            "value": n * n / (step * 1e-3), "unit": "frame-pairs/s",
            "stages_ms": {k: float(v) for k, v in zip(names, np.median(np.array(st), axis=0))}, "sweeps": int(sweeps),
            "pairs_computed_fraction": 1.0 / (2 * stride),
            "note": "frame-pairs/s counts all N^2 pairs of the clip as the headline does; the tensor cores evaluate "
                    "N^2/(2*stride) of them (the residue classes, upper triangles). Same D2/D3/D3_new bits as the headline pass."}


def sharded_residue_record(frames, n, fs, stride, rank, world, dev, flush, steps=5, warm=2):
    """residue_step_record for N > 1: the row-sharded step on residue-class planes (dist.py, `residues=True`):
    world x stride Gram jobs in one launch per rank, transposes pushed to the peers, per-plane halos."""
    import torch.distributed as dist

    from audio_video_textures_b200 import dist as avdist
    ws = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev, residues=True)
    names = ["norms", "gram", "filter", "future_cost", "finalize"]
    for _ in range(warm):
        avdist.classic_sharded(frames, fs, stride, rank, world, workspace=ws)
    evs = []
    for _ in range(steps):                               # ranks re-aligned before every step (see run_ours)
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        ev = _events(2)
        ev[0].record()
        avdist.classic_sharded(frames, fs, stride, rank, world, workspace=ws)
        ev[1].record()
        evs.append(ev)
    torch.cuda.synchronize()
    dist.barrier()
    ms = [a.elapsed_time(b) for a, b in evs]
    runs = []
    for _ in range(3):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        res = avdist.classic_sharded(frames, fs, stride, rank, world, workspace=ws, timing=True)
        runs.append([res.stage_ms[k] for k in names])
    t = torch.tensor([sum(ms)] + list(np.median(np.array(runs), axis=0)), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step = float(t[0].item()) / steps
    return {"eligible": True, "ms_per_step": step, "value": n * n / (step * 1e-3), "unit": "frame-pairs/s",
            "stages_ms_max_over_ranks": {k: float(v) for k, v in zip(names, t[1:].tolist())},
            "sweeps": int(res.n_sweeps), "pairs_computed_fraction": 1.0 / (2 * stride),
            "note": "row-sharded step on residue-class planes: the tensor cores evaluate N^2/(2*stride) pairs; same "
                    "D2/D3/D3_new bits as the headline pass (extra.parity.residue_class_shards)"}


def c5_record(args, dev, rank, world, peaks):
    """configs[4] at THIS N: 100000 frames 64x64, -m 3 -fs 40 -stride 4 (M = 24991).  Strong scaling: the same
    clip at every N, rows sharded over the ranks (N = 1: the single-GPU path, D1 = 40 GB resident)."""
    import torch.distributed as dist

    from audio_video_textures_b200 import dist as avdist
    from audio_video_textures_b200 import engine
    from audio_video_textures_b200.synth import synth_video_cuda
    wl = WORKLOADS["c5"]
    n, fs, stride = wl["n"], wl["fs"], wl["stride"]
    frames = synth_video_cuda(n, wl["h"], wl["w"], seed=0, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    steps, warm = max(3, min(args.steps, 5)), 2
    ws = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev) if world > 1 else None
    stage_names = ["norms", "gram", "filter", "future_cost", "finalize"]

    def one(timing):
        if world == 1:
            ev = _events(6)
            ev[0].record()
            pf = engine.pack_frames(frames)
            ev[1].record()
            D1 = engine.gram_l2(pf)
            ev[2].record()
            D2, D3 = engine.diag_filter(D1, fs, stride, p=0.7)
            ev[3].record()
            fc = engine.future_cost_fused(D3, 0.997)
            ev[4].record()
            D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997)
            ev[5].record()
            torch.cuda.synchronize()
            return ev[0].elapsed_time(ev[5]), {s: ev[i].elapsed_time(ev[i + 1]) for i, s in enumerate(stage_names)}, fc.n_sweeps
        ev = _events(2)
        ev[0].record()
        res = avdist.classic_sharded(frames, fs, stride, rank, world, workspace=ws, timing=timing)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), res.stage_ms, res.n_sweeps

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(warm):
        one(False)
    ms = []
    for _ in range(steps):
        flush.fill_(1)
        sync_all()
        ms.append(one(False)[0])
    stage_runs = []
    for _ in range(3):
        flush.fill_(1)
        sync_all()
        _, st, sweeps = one(True)
        stage_runs.append([st[s] for s in stage_names])
    tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    stg = torch.tensor(np.median(np.array(stage_runs), axis=0), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(stg, op=dist.ReduceOp.MAX)
    residue = None
    try:
        if world == 1:
            residue = residue_step_record(frames, fs, stride, flush, steps=3, warm=1)
        else:
            residue = sharded_residue_record(frames, n, fs, stride, rank, world, dev, flush, steps=5, warm=2)
    except Exception as exc:
        residue = {"error": f"{type(exc).__name__}: {exc}"}
    # end to end: pinned host clip -> device (1/G per rank + NVLink all-gather) -> pipeline incl. sigma3 / P3_new ->
    # survivor lists on the host
    host = frames.reshape(n, -1).cpu().pin_memory()
    del frames
    e2e_ms, d2h = [], 0
    for it in range(3):
        flush.fill_(1)
        sync_all()
        t0 = time.perf_counter()
        dev_frames = (avdist.load_frames_pushed(host, ws) if ws is not None else
                      avdist.load_frames_sharded(host, rank, world, dev))
        if world == 1:
            pf = engine.pack_frames(dev_frames)
            D1 = engine.gram_l2(pf)
            D2, D3 = engine.diag_filter(D1, fs, stride, p=0.7)
            fc = engine.future_cost_fused(D3, 0.997)
            stats = engine.new_stats(dev)
            D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997, stats=stats)
            sigma = engine.sigma_from_stats(*engine.read_stats(stats), 4.5)
            P3, P3n, counts = engine.transition_probs(D3n, sigma, threshold=0.08, want_counts=True)
            rows = engine.SurvivorRows.from_matrix(P3n)     # 3e8 survivors: only the visited rows travel
            from audio_video_textures_b200.classic.video_textures import texture_walk
            np.random.seed(0)
            texture_walk(rows, wl["m"], 30, 30, stride, fs)
            d2h = int(sum(v.nbytes for v in rows._cache.values()))
            del D1, D2, D3, D3n, P3, P3n
        else:
            res = avdist.classic_sharded(dev_frames, fs, stride, rank, world, sigma_factor=4.5, threshold=0.08,
                                         workspace=ws)
            d2h = _walk_sharded(avdist, engine, res, ws, wl, rank)
            del res
        sync_all()
        if it >= 1:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
    t_e2e = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    step_ms = float(tot.item()) / steps
    stages = {s: float(v) for s, v in zip(stage_names, stg.tolist())}
    return {"workload": workload_name(wl), "n_gpus": world, "scaling": "strong", "steps": steps, "warmup": warm,
            "ms_per_step": step_ms, "value": n * n / (step_ms * 1e-3), "unit": "frame-pairs/s",
            "sweeps": int(sweeps), "M": engine.filtered_size(n, fs, stride),
            "stages_ms_max_over_ranks": stages, "residue_pipeline": residue,
            "e2e": {"ms_per_step": float(t_e2e.item()), "value": n * n / (float(t_e2e.item()) * 1e-3),
                    "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": int(d2h),
                    "includes": "pinned host clip -> HBM, norms, Gram, filter, future cost, sigma3, P3, P3_new, the 900-frame "
                                "-m 3 walk over survivor lists fetched on demand"},
            "l2": "256 MB L2 flush between timed steps"}


def parity_record(dev, rank, world):
    """Sharded == single-GPU on a 3000-frame clip (M = 741), through the same symmetric workspace path the timed
    steps use; every rank checks its own shard; MIN over ranks."""
    import torch.distributed as dist

    from audio_video_textures_b200 import dist as avdist
    from audio_video_textures_b200 import selfcheck
    from audio_video_textures_b200.synth import synth_video
    n, fs, stride = 3000, 40, 4
    frames = synth_video(n, 32, 32, seed=3).to(dev)
    ws = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev)
    for _ in range(2):
        res = avdist.classic_sharded(frames, fs, stride, rank, world, sigma_factor=4.5, threshold=0.08, workspace=ws)
    single = selfcheck.single_gpu_pipeline(frames, fs, stride, 4.5, 0.08)
    ok = selfcheck.shard_equals_single(res, single)
    keys = sorted(ok)
    flags = torch.tensor([int(ok[k]) for k in keys], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    failed = [k for k, v in zip(keys, flags.tolist()) if not v]
    # the residue-class shards (stride 4: 1/4 of the pairs) against the same single-GPU FULL-D1 pipeline
    ws_r = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev, residues=True)
    for _ in range(2):
        res_r = avdist.classic_sharded(frames, fs, stride, rank, world, sigma_factor=4.5, threshold=0.08, workspace=ws_r)
    ok_r = selfcheck.shard_equals_single(res_r, single)
    flags = torch.tensor([int(ok_r[k]) for k in keys], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    failed += ["residues:" + k for k, v in zip(keys, flags.tolist()) if not v]
    return {"status": "bit-exact" if not failed else "MISMATCH: " + ",".join(failed),
            "residue_class_shards": "bit-exact against the full-D1 single-GPU pipeline" if not any(
                f.startswith("residues:") for f in failed) else "MISMATCH",
            "clip": f"{n} frames 32x32, -fs {fs} -stride {stride}, M={res.plan.m}", "ranks": world,
            "checked": {"bit_exact": ["D1", "D2", "D3n", "sweeps", "survivors"], "rtol_1e-6": ["eps", "sigma"],
                        "rtol_1e-5": ["P3"]},
            "sweeps": int(res.n_sweeps)}


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
    frames = synth_video_cuda(n, wl["h"], wl["w"], seed=0, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    peaks = load_peaks()
    gram_ms, step_ms, pending = [], [], []
    state = {}
    workspace = None
    comm_note = ""
    if world > 1 and not args.no_symmetric:
        try:
            workspace = avdist.SymmetricShardWorkspace(n, fs, stride, rank, world, dev)
            ok = torch.ones(1, device=dev)
        except Exception as exc:                      # no peer-mapped memory on this box: NCCL path of the same algorithm
            workspace = None
            ok = torch.zeros(1, device=dev)
            comm_note = f"symmetric memory unavailable ({type(exc).__name__}): NCCL all-gather path"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            workspace = None

    def one_step(timed: bool):
        ev = _events(4)
        ev[0].record()
        if world == 1:
            pf = engine.pack_frames(frames)
            ev[1].record()
            D1 = engine.gram_l2(pf)
            ev[2].record()
            D2, D3 = engine.diag_filter(D1, fs, stride, p=0.7)
            fc = engine.future_cost_fused(D3, 0.997)
            D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997)
            launches = 1 + 1 + 1 + 1 + 1
            state.update(D1=D1, D3n=D3n, fc=fc, m=D3.shape[0])
        else:
            ev[1].record()
            res = avdist.classic_sharded(frames, fs, stride, rank, world, workspace=workspace)
            ev[2].record()
            launches = res.launches
            state.update(D3n=res.D3_new, fc=res.fc, m=res.plan.m)
        ev[3].record()
        if timed:
            pending.append(ev)                              # read after the ONE sync that closes the timed region
        else:
            torch.cuda.synchronize()
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
        # The K steps are enqueued back to back and bracketed by ONE synchronisation (+ barrier) on each side, as the
        # contract says; every step has its own CUDA-event pair on the stream (the 256 MB L2 flush between steps is
        # outside the pairs).  Synchronising after every step instead adds the host's launch preparation (~0.1 ms of
        # Python per step, with the GPU idle behind it) to a 1.5 ms step.
        for _ in range(args.steps):
            flush.fill_(1)                      # L2 flush (256 MB > 126 MB L2) between timed steps
            if world > 1:
                # N > 1: ranks are re-aligned before every step.  Free-running ranks advance at the pace of the
                # slowest one INCLUDING its in-stream flush, which the faster ranks' event pairs would then absorb as
                # barrier wait inside the step (measured at 8 GPUs: 1.98 ms instead of 1.88).
                sync_all()
            launches += one_step(True)
        sync_all()
        for ev in pending:
            step_ms.append(ev[0].elapsed_time(ev[3]))
            if world == 1:
                gram_ms.append(ev[1].elapsed_time(ev[2]))
        # N = 1: the same pass as ONE CUDA-graph replay per step (engine.PipelineGraph, full D1): no host gaps between
        # the five kernels.  This is the headline when the capture succeeds; the eager loop above supplies the Gram's
        # own event pair for the roofline and is reported as extra.eager.
        graph_ms, graph_note = None, None
        if world == 1 and not args.no_graph:
            try:
                g = engine.PipelineGraph(n, frames[0].numel(), fs, stride, residues=False, device=dev)
                g.frames.copy_(frames.reshape(n, -1))
                for _ in range(args.warmup):
                    g()
                    flush.fill_(1)
                torch.cuda.synchronize()
                gev = []
                for _ in range(args.steps):
                    flush.fill_(1)
                    ev = _events(2)
                    ev[0].record()
                    g()
                    ev[1].record()
                    gev.append(ev)
                torch.cuda.synchronize()
                graph_ms = [a.elapsed_time(b) for a, b in gev]
                if not (torch.equal(g.D3_new, state["D3n"]) and g.n_sweeps == int(state["fc"].n_sweeps)):
                    raise RuntimeError("graph replay differs from the eager pass")
                del g
            except Exception as exc:
                graph_ms, graph_note = None, f"{type(exc).__name__}: {exc}"
    sync_all()
    wall = time.perf_counter() - wall0
    sweeps = int(state["fc"].n_sweeps)
    eager_step_ms = list(step_ms)
    if graph_ms is not None:
        step_ms[:] = graph_ms
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    value = n * n * args.steps / total_s

    # ---- per-stage times of the sharded step (max over ranks), measured on separate steps
    stages = None
    if world > 1:
        names = ["norms", "gram", "filter", "future_cost", "finalize"]
        runs = []
        for _ in range(5):
            flush.fill_(1)
            sync_all()
            res = avdist.classic_sharded(frames, fs, stride, rank, world, workspace=workspace, timing=True)
            runs.append([res.stage_ms[s] for s in names])
        stg = torch.tensor(np.median(np.array(runs), axis=0), dtype=torch.float64, device=dev)
        dist.all_reduce(stg, op=dist.ReduceOp.MAX)
        stages = {s: float(v) for s, v in zip(names, stg.tolist())}

    # ---- end to end from pinned host frames; same scope at every N: ... sigma3, P3_new, survivor lists on the host
    host = frames.reshape(n, -1).cpu().pin_memory()
    f = torch.tensor(4.5, dtype=torch.float32)
    times, d2h = [], 0
    reps = args.warmup + max(3, args.steps // 4)
    if world == 1:
        from audio_video_textures_b200.classic.computeD1 import compute_D1
        from audio_video_textures_b200.classic.computeD2 import compute_D2
        from audio_video_textures_b200.classic.q_learning import LAST, q_learning
        from audio_video_textures_b200.classic.video_textures import texture_walk
        host4 = host.view(n, wl["h"], wl["w"], 3)
        for it in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                D1, P1, s1 = compute_D1(host4, f, "RGB", slow=True, batch_size=48)
                if wl["m"] in (1, 2):
                    D2, P2, s2, _ = compute_D2(D1, f, filter_size=fs)
                else:
                    D2, P2, s2, _ = compute_D2(D1, f, filter_size=fs, stride=stride)
                D3n, P3, P3n, s3 = q_learning(D2, f, thresholding=0.08)
            rowptr, colidx = engine.csr_from_matrix(P3n, LAST["counts"])          # what the walk consumes (D2H)
            np.random.seed(0)
            walk, _ = texture_walk((rowptr, colidx), wl["m"], 30, 30, stride, fs)
            sig = s3.item()
            torch.cuda.synchronize()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
            d2h = rowptr.nbytes + colidx.nbytes + 3 * 4
        state.update(P3n=P3n, counts=LAST["counts"])
        t_e2e = float(np.median(times))          # median: one page-locked allocation or host hiccup is not the path
        includes = ("compute_D1+compute_D2+q_learning (P1,P2,P3,P3_new, sigmas) + survivor lists D2H + the "
                    f"{len(walk)}-frame -m {wl['m']} walk")
    else:
        for it in range(reps):
            flush.fill_(1)
            sync_all()
            t0 = time.perf_counter()
            if workspace is not None:                                           # 1/G over PCIe, NVLink pushes underneath
                dev_frames = avdist.load_frames_pushed(host, workspace)
            else:
                dev_frames = avdist.load_frames_sharded(host, rank, world, dev)  # 1/G over PCIe + NCCL all-gather
            res = avdist.classic_sharded(dev_frames, fs, stride, rank, world, sigma_factor=f, threshold=0.08,
                                         workspace=workspace)
            d2h = _walk_sharded(avdist, engine, res, workspace, wl, rank)
            sync_all()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = torch.tensor([float(np.median(times))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
        includes = ("each rank copies 1/G of the pinned host clip and pushes every piece to its peers over NVLink as it "
                    "lands (NCCL all-gather without symmetric memory); sharded norms / "
                    "Gram / filter / future cost; sigma3 (all-reduce), P3, P3_new; the 900-frame walk on rank 0 over "
                    "survivor lists read on demand from the owning ranks' shards (peer-mapped)")
    e2e = {"value": n * n / t_e2e, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(host.numel()),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e, "includes": includes}
    del host

    extra = {}
    if world > 1 and workspace is not None and not args.skip_extra:
        extra["parity"] = parity_record(dev, rank, world)
    roof = hbm = cpu = None
    if world == 1 and rank == 0:
        g_ms = float(np.mean(gram_ms))
        roof = gram_roofline(frames, wl, g_ms, g_ms * len(gram_ms) / sum(eager_step_ms), dev, peaks)
        if not args.skip_extra:
            extra["synth"] = synth_records(dev, state)
        D1_host = state["D1"].cpu() if n <= 8000 else None
        state.clear()
        if not args.skip_extra:
            hbm = hbm_rooflines(dev, peaks)
        cpu = {kk: vv for kk, vv in cpu_reference(wl, args.cpu_budget, D1_host).items() if kk != "wall_s"}
        cpu["unit"] = "frame-pairs/s"
    if not args.skip_extra and stride >= 2 and (world == 1 or (workspace is not None and n % stride == 0)):
        try:
            if world == 1:
                extra["residue_pipeline"] = residue_step_record(frames, fs, stride, flush, steps=max(5, args.steps // 2))
            else:
                extra["residue_pipeline"] = sharded_residue_record(frames, n, fs, stride, rank, world, dev, flush,
                                                                   steps=max(5, args.steps // 2))
        except Exception as exc:
            extra["residue_pipeline"] = {"error": f"{type(exc).__name__}: {exc}"}
    state.clear()
    del frames
    torch.cuda.empty_cache()
    if not args.skip_extra and args.workload == "c2":
        try:
            extra["c5"] = c5_record(args, dev, rank, world, peaks)
        except Exception as exc:                                            # never lose the main line to the extra record
            extra["c5"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        return
    out = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None,
        "dtype": "u8 (exact int32 tensor-core Gram) + fp32", "data": "synthetic",
        "config": {"workload": workload_name(wl), "name": args.workload,
                   "detail": f"M={engine.filtered_size(n, fs, stride)}, {sweeps} future-cost sweeps to eps <= 0.01",
                   "l2": "256 MB L2 flush in front of every timed step (in-stream, outside the step's event pair)",
                   "sharding": "single GPU" if world == 1 else
                   (f"rows over {world} ranks, N=5000*sqrt(G)" if args.workload == "c2" else f"rows over {world} ranks") +
                   ("" if args.no_symmetric or workspace is None else
                    "; norms, transposed Gram tiles and per-sweep row minima pushed to peer shards over NVLink from inside the kernels")},
        "gpu_launches": launches, "wall_s": wall,
    }
    if world == 1:
        out["config"]["launch"] = ("one CUDA-graph replay per step (engine.PipelineGraph; results checked equal to the "
                                   "eager pass)" if graph_ms is not None else
                                   "eager launches" + (f" (graph capture failed: {graph_note})" if graph_note else ""))
        extra["eager"] = {"ms_per_step": float(np.mean(eager_step_ms)),
                          "note": "the same five launches issued one by one from Python (host gaps included); "
                                  "roofline.ms and share_of_step come from this loop"}
    if roof is not None:
        out["roofline"] = roof
    if hbm is not None:
        out["roofline_hbm"] = hbm
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if stages is not None:
        out["stages"] = stages
    out["e2e"] = e2e
    out["extra"] = extra
    if "parity" in extra:
        out["parity"] = extra["parity"]["status"]
    if comm_note:
        out["config"]["sharding"] += "; " + comm_note
    out["clocks"] = clocks.summary()
    print(json.dumps(out))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = scaled_workload(args.workload, args.gpus)        # the same clip the GPU arm uses at this --gpus
    total = args.warmup + args.steps
    budget = min(args.cpu_budget, max(1.5, 150.0 / total))  # the whole run stays within a few minutes
    vals, walls = [], []
    info = None
    for it in range(total):
        info = cpu_reference(wl, budget_s=budget, seed=it)
        if it >= args.warmup:
            vals.append(info["value"])
            walls.append(info["wall_s"])
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
           "higher_is_better": True, "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None,
           "dtype": "fp32", "data": "synthetic",
           "config": {"workload": workload_name(wl), "name": args.workload},
           "cpu_baseline": {"value": v, "unit": "frame-pairs/s", "cores": info["cores"], "kind": "port",
                            "sample": info["sample"]},
           "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "full_pass_s": info["seconds_full"],
           "note": ("each step times a bounded sample of the workload (ms_per_step is the sample's wall time); value = "
                    "sampled frame-pairs / their time, which equals N^2 / full-pass time because every 48x48 block of "
                    "the reference algorithm costs the same"),
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no_symmetric", action="store_true",
                    help="N>1: plain row shards + NCCL exchange instead of peer pushes from the kernels")
    ap.add_argument("--skip_extra", action="store_true", help="main line only (no c5 / synth / hbm / parity records)")
    ap.add_argument("--no_graph", action="store_true", help="N = 1: time the eager launches instead of the CUDA-graph replay")
    ap.add_argument("--cpu_budget", type=float, default=20.0, help="seconds of CPU work for the baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

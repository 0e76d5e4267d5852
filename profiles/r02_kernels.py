"""Kernel-level driver for ncu captures and A/B timing (round 2).
    python profiles/r02_kernels.py filter1 [N]     stride-1 filter + pow at N frames (default 16000 -> M = 15961)
    python profiles/r02_kernels.py filter4         stride-4 filter on the C5 row-shard shape (12576 x 100000 D1 rows)
    python profiles/r02_kernels.py norms           K0 on the C5 clip (100000 rows of 12288 B)
    python profiles/r02_kernels.py gram5 [rows]    K1 on a C5 row shard (rows x 100000, K = 12288)
    python profiles/r02_kernels.py synth           150 fused synthesis steps at C3 with per-step kernel time
Environment switches read by the library: AVTEX_FILTER_S1=0 (old stride-1 kernel), AVTEX_NORMS_G=32|64|128|256.
Prints CUDA-event medians; run under ncu for per-launch counters."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_embeddings, synth_video_cuda

what = sys.argv[1]
HBM = 6547.8


def timed(fn, reps=5):
    ms = []
    for _ in range(reps + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        r = fn()
        ev[1].record()
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    return float(np.median(ms[2:])), r


if what == "filter1":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
    D1 = engine.empty_matrix(n, n, "cuda")
    D1.uniform_(100.0, 20000.0)
    m = n - 39
    for stats in (False, True):
        st = engine.new_stats("cuda") if stats else None
        ms, _ = timed(lambda: engine.diag_filter(D1, 40, 1, p=0.7, stats=st))
        b = 4.0 * n * n + 8.0 * m * m
        print(f"filter1 N={n} stats={stats} s1={os.environ.get('AVTEX_FILTER_S1', '1')}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s  {b / ms / 1e6 / HBM:.3f} of HBM")
elif what == "filter1s":
    # symmetric form of the stride-1 filter (D1 produced by the Gram kernel: known symmetric)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
    A = torch.empty((n, n), dtype=torch.float32, device="cuda").uniform_(100.0, 20000.0)
    D1 = engine.empty_matrix(n, n, "cuda")
    D1.copy_(torch.triu(A) + torch.triu(A, 1).T)
    del A
    m = n - 39
    for stats in (False, True):
        st = engine.new_stats("cuda") if stats else None
        ms, _ = timed(lambda: engine.diag_filter(D1, 40, 1, p=0.7, stats=st, symmetric=True))
        b_full, b_sym = 4.0 * n * n + 8.0 * m * m, 2.0 * n * n + 8.0 * m * m
        print(f"filter1 SYMMETRIC N={n} stats={stats}: {ms:.3f} ms  {b_sym / ms / 1e6:.0f} GB/s of the bytes it needs "
              f"({b_sym / ms / 1e6 / HBM:.3f} of HBM); against the general kernel's bytes {b_full / ms / 1e6 / HBM:.3f}")
elif what == "filter4s":
    # symmetric form, stride 4, the single-GPU C5 shape scaled to 40000 frames
    n = 40000
    A = torch.empty((n, n), dtype=torch.float32, device="cuda").uniform_(100.0, 20000.0)
    D1 = engine.empty_matrix(n, n, "cuda")
    D1.copy_(torch.triu(A) + torch.triu(A, 1).T)
    del A
    m = (n - 40) // 4 + 1
    for sym in (False, True):
        ms, _ = timed(lambda: engine.diag_filter(D1, 40, 4, p=0.7, symmetric=sym))
        b = (2.0 if sym else 4.0) * n * n + 8.0 * m * m
        print(f"filter4 N={n} symmetric={sym}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s  {b / ms / 1e6 / HBM:.3f} of HBM (own bytes)")
elif what == "filter4":
    rows, n = 12576, 100000
    D1 = torch.empty((rows, n), dtype=torch.float32, device="cuda").uniform_(100.0, 20000.0)
    m = (n - 40) // 4 + 1
    ro = (rows - 40) // 4 + 1
    ms, _ = timed(lambda: engine.diag_filter(D1, 40, 4, p=0.7, m=m, a0=0, rows_out=ro, in_row0=0))
    b = 4.0 * rows * n + 8.0 * ro * m
    print(f"filter4 shard: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s  {b / ms / 1e6 / HBM:.3f} of HBM")
elif what == "norms":
    frames = synth_video_cuda(100000, 64, 64, seed=0)
    x = frames.reshape(100000, -1)
    sq = torch.empty(100000, dtype=torch.int64, device="cuda")
    fl = torch.zeros(2, dtype=torch.int64, device="cuda")
    ms, _ = timed(lambda: engine.frame_norms_rows(x, 0, 100000, sq, fl))
    b = float(x.numel())
    print(f"norms G={os.environ.get('AVTEX_NORMS_G', 'default')}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s  {b / ms / 1e6 / HBM:.3f} of HBM")
elif what == "gram5":
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 12536
    frames = synth_video_cuda(100000, 64, 64, seed=0)
    pf = engine.pack_frames(frames)
    D = engine.empty_matrix(rows, 100000, "cuda")
    ms, _ = timed(lambda: engine.gram_l2(pf, 0, rows, symmetric=False, out=D), reps=3)
    ops = 2.0 * rows * 100000 * 12288
    print(f"gram5 {rows} x 100000 K=12288: {ms:.3f} ms  {ops / ms / 1e9:.0f} TOP/s")
elif what == "gramsym":
    # symmetric Gram (the single-GPU C5 shape, scaled down): n frames 64x64, K = 12288, direct + transposed stores
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    frames = synth_video_cuda(n, 64, 64, seed=0)
    pf = engine.pack_frames(frames)
    D = engine.empty_matrix(n, n, "cuda")
    ms, _ = timed(lambda: engine.gram_l2(pf, out=D), reps=3)
    tiles = (-(-n // 256)) * (-(-n // 256) + 1) // 2
    ops = 2.0 * tiles * 256 * 256 * 12288
    cfg = " ".join(f"{k}={os.environ[k]}" for k in ("AVTEX_GRAM_GROUP", "AVTEX_GRAM_HINT", "AVTEX_GRAM_ST") if k in os.environ)
    print(f"gramsym n={n} K=12288 [{cfg}]: {ms:.3f} ms  {ops / ms / 1e9:.0f} TOP/s executed ({tiles} tiles)")
elif what == "gramfused":
    # C2 Gram with K0 fused into the launch (norms by the idle epilogue warps) vs separate K0 + K1
    frames = synth_video_cuda(5000, 224, 224, seed=0)
    def sep():
        return engine.gram_l2(engine.pack_frames(frames))
    def fused():
        return engine.gram_l2(engine.pack_frames(frames, defer_norms=True))
    ms_s, D_s = timed(sep, reps=5)
    ms_f, D_f = timed(fused, reps=5)
    assert torch.equal(D_s, D_f)
    print(f"gramfused C2: K0 + K1 separate {ms_s:.3f} ms; K0 inside K1 {ms_f:.3f} ms")
elif what == "gramjobs":
    # job list of rank `me` of an 8-rank C5 step, run on ONE GPU: the peer destinations all point at one local scratch
    # shard, so the time is everything except the NVLink transport of the pushed tiles
    from audio_video_textures_b200 import dist as D
    me = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    world, n, fs, stride = 8, 100000, 40, 4
    frames = synth_video_cuda(n, 64, 64, seed=0)
    pf = engine.pack_frames(frames)
    plans = [D.plan_shards(n, fs, stride, world, r) for r in range(world)]
    ld = (n + 31) // 32 * 32
    rows_max = max(p.r_hi - p.r_lo for p in plans)
    mine = torch.empty((rows_max, ld), dtype=torch.float32, device="cuda")
    peer = torch.empty((rows_max, ld), dtype=torch.float32, device="cuda")
    ptrs = [mine.data_ptr() if r == me else peer.data_ptr() for r in range(world)]
    jobs = D.symmetric_jobs(plans, me, ptrs, ld, stride)
    tiles = sum((-(-j["rows"] // 256)) * (-(-j["cols"] // 256)) if not j["symmetric"] else
                (-(-j["rows"] // 256)) * (-(-j["rows"] // 256) + 1) // 2 for j in jobs)
    for mode in ("push", "nopush"):
        jl = jobs if mode == "push" else [dict(j, DT=(j["DT"] if j["symmetric"] else None)) for j in jobs]
        ms, _ = timed(lambda: engine.gram_l2_jobs(pf, jl), reps=5)
        print(f"gramjobs rank {me}/8 C5 [{mode}]: {ms:.3f} ms  {tiles} tiles  {2.0 * tiles * 65536 * 12288 / ms / 1e9:.0f} TOP/s executed")
elif what == "fc":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    frames = synth_video_cuda(n, 64, 64, seed=0)
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf)
    D2, D3 = engine.diag_filter(D1, 40, 1, p=0.7)
    del D1, D2
    m = D3.shape[0]
    ms, fc = timed(lambda: engine.future_cost_fused(D3), reps=3)
    passes = fc.passes
    b = 4.0 * m * m * passes
    print(f"future_cost_fused M={m}: {ms:.3f} ms for {passes} passes = {ms / passes:.3f} ms/pass  {b / ms / 1e6:.0f} GB/s  {b / ms / 1e6 / HBM:.3f} of HBM")
    ms2, fc2 = timed(lambda: engine.future_cost(D3), reps=2)
    print(f"future_cost (one launch per sweep, host reads eps) M={m}: {ms2:.3f} ms for {fc2.passes} passes")
    assert torch.equal(fc.mvec, fc2.mvec[:m]) and fc.n_sweeps == fc2.n_sweeps
elif what == "synth":
    from audio_video_textures_b200.contrastive.validate import SynthesisState
    L, D = 20000, 2304
    emb = synth_embeddings(L, D, seed=0, device="cuda")
    st = SynthesisState(emb)
    np.random.seed(0)
    q = 10
    kern, wall = [], []
    for it in range(1, 160):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t0 = time.perf_counter()
        ev[0].record()
        ch = st.step(q, it, 0.1, 0.5, 0.3)
        ev[1].record()
        t1 = time.perf_counter()
        q = int(np.random.choice(ch))
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        kern.append(ev[0].elapsed_time(ev[1]) * 1e3)
        wall.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6))
    w = np.array(wall[10:])
    print(f"synth C3: kernel (events) median {np.median(kern[10:]):.1f} us; host step() {np.median(w[:, 0]):.1f} us; "
          f"np.random.choice {np.median(w[:, 1]):.1f} us")
    for n_steps in (149, 600):
        best = 1e9
        for rep in range(4):
            np.random.seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            q_ids, nz = engine.synthesis_loop(st.ws, st.tn, st.qn, 10, n_steps, 0.1, 0.5, 0.3)
            best = min(best, time.perf_counter() - t0)
        print(f"synth C3 device loop: {n_steps} steps in {best * 1e3:.3f} ms = {best / n_steps * 1e6:.1f} us/step "
              f"({4.0 * L * D / (best / n_steps) / 1e9 / HBM:.3f} of HBM)")

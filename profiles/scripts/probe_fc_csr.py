import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda

def ev_ms(fn, reps=3):
    out = []
    for _ in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); r = fn(); e[1].record(); torch.cuda.synchronize()
        out.append(e[0].elapsed_time(e[1]))
    return out, r

what = sys.argv[1]
if what == "fc5":
    n = 100000
    frames = synth_video_cuda(n, 64, 64, seed=0)
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf)
    D2, D3 = engine.diag_filter(D1, 40, 4, p=0.7)
    torch.cuda.synchronize()
    print("M", D3.shape, D3.stride())
    ms, fc = ev_ms(lambda: engine.future_cost_fused(D3, 0.997))
    print("fc after pipeline, D1 resident:", ms, "passes", fc.passes)
    del D1, D2
    torch.cuda.empty_cache()
    ms, fc = ev_ms(lambda: engine.future_cost_fused(D3, 0.997))
    print("fc, D1 freed:", ms, "passes", fc.passes)
    D3b = engine.empty_matrix(D3.shape[0], D3.shape[1], "cuda"); D3b.copy_(D3)
    ms, fc = ev_ms(lambda: engine.future_cost_fused(D3b, 0.997))
    print("fc on a copy:", ms, "passes", fc.passes, "eps", fc.eps_trail[:3])
    D3c = engine.empty_matrix(D3.shape[0], D3.shape[1], "cuda"); D3c.uniform_(100.0, 2000.0)
    ms, fc = ev_ms(lambda: engine.future_cost_fused(D3c, 0.997))
    print("fc on uniform random values:", ms, "passes", fc.passes)
else:
    n = 5000
    frames = synth_video_cuda(n, 224, 224, seed=0)
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf)
    D2, D3 = engine.diag_filter(D1, 40, 4, p=0.7)
    fc = engine.future_cost_fused(D3, 0.997)
    st = engine.new_stats("cuda")
    D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997, stats=st)
    sigma = engine.sigma_from_stats(*engine.read_stats(st), 4.5)
    P3, P3n, counts = engine.transition_probs(D3n, sigma, threshold=0.08, want_counts=True)
    torch.cuda.synchronize()
    for _ in range(4):
        t0 = time.perf_counter(); rp, ci = engine.csr_from_matrix(P3n, counts); t1 = time.perf_counter()
        print(f"csr_from_matrix {1e3 * (t1 - t0):.3f} ms  nnz {len(ci)}")
    import ctypes as C
    from audio_video_textures_b200 import _lib
    rows, cols = P3n.shape
    both = torch.empty((rows + 1) * 8 + rows * cols * 4, dtype=torch.uint8, device="cuda")
    rowptr = both[:(rows + 1) * 8].view(torch.int64); colidx = both[(rows + 1) * 8:].view(torch.int32)
    for name, fn in (("cumsum", lambda: torch.cumsum(counts, 0, out=rowptr[1:])),
                     ("fill", lambda: _lib.call("avtex_csr_fill", _lib.ptr(P3n), P3n.stride(0), rows, cols, _lib.ptr(rowptr), _lib.ptr(colidx), engine._dev(P3n), engine._stream(P3n))),
                     ("copy", lambda: engine._pinned(both.numel())[:both.numel()].copy_(both, non_blocking=True))):
        ms, _ = ev_ms(fn, 4)
        print(name, ms)

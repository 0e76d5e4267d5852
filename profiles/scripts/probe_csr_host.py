import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_video_textures_b200 import engine, _lib
torch.manual_seed(0)
M = 1241
P = engine.empty_matrix(M, M, "cuda"); P.uniform_(0, 1); P[P < 0.52] = 0
counts = (P != 0).sum(1).to(torch.int32)
torch.cuda.synchronize()
for rep in range(4):
    T = [time.perf_counter()]
    def tick(): T.append(time.perf_counter())
    rows, cols = P.shape; dev = P.device
    cap = rows * cols; head = (rows + 1) * 8
    both = torch.empty(head + cap * 4, dtype=torch.uint8, device=dev); tick()
    rowptr = both[:head].view(torch.int64); colidx = both[head:].view(torch.int32); tick()
    rowptr[0] = 0; tick()
    torch.cumsum(counts, 0, out=rowptr[1:]); tick()
    _lib.call("avtex_csr_fill", _lib.ptr(P), P.stride(0), rows, cols, _lib.ptr(rowptr), _lib.ptr(colidx), engine._dev(P), engine._stream(P)); tick()
    host = engine._pinned(both.numel())[:both.numel()]; tick()
    host.copy_(both, non_blocking=True); tick()
    torch.cuda.current_stream(dev).synchronize(); tick()
    rp = host[:head].view(torch.int64).numpy().copy(); tick()
    total = int(rp[-1]); ci = host[head:head + total * 4].view(torch.int32).numpy().copy(); tick()
    names = ["empty", "views", "rowptr0", "cumsum", "fill", "pinned", "copy_", "sync", "rp", "ci"]
    print(rep, " ".join(f"{n}={1e3*(b-a):.3f}" for n, a, b in zip(names, T[:-1], T[1:])), f"total={1e3*(T[-1]-T[0]):.3f} nnz={total}")

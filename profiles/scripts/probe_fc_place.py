"""Does the future-cost time depend on where / when the 2.5 GB matrix was allocated?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_video_textures_b200 import engine

def ev(fn, reps=3):
    out = []
    for _ in range(reps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); r = fn(); e[1].record(); torch.cuda.synchronize(); out.append(e[0].elapsed_time(e[1]))
    return float(np.median(out[1:])), r

mode = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 24991
if mode == "big_first":
    big = torch.empty(40 * (1 << 30), dtype=torch.uint8, device="cuda")
    big.fill_(1)
elif mode == "big_freed":
    big = torch.empty(40 * (1 << 30), dtype=torch.uint8, device="cuda")
    big.fill_(1)
    del big
    torch.cuda.empty_cache()
D3 = engine.empty_matrix(M, M, "cuda")
D3.uniform_(100.0, 2000.0)
ms, fc = ev(lambda: engine.future_cost_fused(D3, 0.997))
print(f"{mode} M={M}: {ms:.3f} ms, {fc.passes} passes, {ms / fc.passes:.3f} ms/pass, {4.0 * M * M * fc.passes / ms / 1e6 / 6547.8:.3f} of HBM")

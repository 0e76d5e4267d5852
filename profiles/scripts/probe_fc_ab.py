"""A/B of the fused future cost: the library in the tree vs a copy built from the previous future_cost.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_video_textures_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1] == "prev":
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libavtex_prev.so")
from audio_video_textures_b200 import engine
def ev(fn, reps=20):
    out = []
    for _ in range(reps + 3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); r = fn(); e[1].record(); torch.cuda.synchronize(); out.append(e[0].elapsed_time(e[1]))
    return float(np.median(out[3:])), r
for M in (1241, 3527, 8000, 19961):
    D3 = engine.empty_matrix(M, M, "cuda"); D3.uniform_(100.0, 2000.0)
    ms, fc = ev(lambda: engine.future_cost_fused(D3, 0.997), 20 if M < 9000 else 5)
    print(f"{sys.argv[1] if len(sys.argv) > 1 else 'new'} M={M}: {ms*1e3:.1f} us, {fc.passes} passes, {ms*1e3/fc.passes:.2f} us/pass")

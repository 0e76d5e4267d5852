"""Profiling driver: ONE rank's share of a row-sharded workload on a single GPU, communication
replaced by local stand-ins (the kernels and their sizes are what a rank of the real run executes).
    python profiles/run_shard.py c5 8 [rank] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_video_textures_b200 import dist as avd
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
from bench import WORKLOADS

name, world = sys.argv[1], int(sys.argv[2])
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
wl = WORKLOADS[name]
n, fs, s = wl["n"], wl["fs"], wl["stride"]
frames = synth_video_cuda(n, wl["h"], wl["w"], seed=0)
plan = avd.plan_shards(n, fs, s, world, rank)
for _ in range(reps):
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf, plan.r_lo, plan.r_hi - plan.r_lo, symmetric=False)
    D2, D3 = engine.diag_filter(D1, fs, s, p=0.7, m=plan.m, a0=plan.a0, rows_out=plan.a1h - plan.a0, in_row0=plan.r_lo)
    own = plan.a1 - plan.a0
    fc = engine.future_cost(D3[:own], 0.997, row0=plan.a0, m=plan.m, pad_to=plan.padded, max_sweeps=3)  \
        if False else None
    # three sweeps' worth of the sweep kernel (convergence needs the other ranks' rows)
    mv = torch.zeros(plan.padded, device="cuda")
    import ctypes as C
    from audio_video_textures_b200 import _lib
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    eps = torch.zeros(1, dtype=torch.float64, device="cuda")
    for it in range(3):
        _lib.call("avtex_future_cost_sweep", _lib.ptr(D3), D3.stride(0), plan.a0, own, plan.m, _lib.ptr(mv),
                  _lib.ptr(mv), C.c_float(0.997), _lib.ptr(mv.clone()), _lib.ptr(eps), 0, st)
    D3n = engine.future_cost_finalize(D3, mv, 0.997, row0=plan.a0, m=plan.m)
    torch.cuda.synchronize()
print("done", name, world, rank, plan)

"""Profiling driver: runs the classic++ stages a few times on a named workload (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = dict(WORKLOADS[name])
if len(sys.argv) > 3:
    wl["n"] = int(sys.argv[3])
frames = synth_video_cuda(wl["n"], wl["h"], wl["w"], seed=0)
residues = os.environ.get("AVTEX_STAGE_RESIDUES") == "1"       # the residue-class pipeline instead of the full D1
for _ in range(reps):
    if residues:
        D2, D3, how = engine.distance_filter(frames, wl["fs"], wl["stride"], p=0.7)
        assert how == "residues"
        fc = engine.future_cost_fused(D3, 0.997)
        D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997)
        torch.cuda.synchronize()
        continue
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf, stats=engine.new_stats(frames.device))
    D2, D3 = engine.diag_filter(D1, wl["fs"], wl["stride"], p=0.7)
    fc = engine.future_cost(D3, 0.997)
    stats = engine.new_stats(frames.device)
    D3n = engine.future_cost_finalize(D3, fc.mvec, 0.997, stats=stats)
    sigma = engine.sigma_from_stats(*engine.read_stats(stats), 4.5)
    P3, P3n, counts = engine.transition_probs(D3n, sigma, threshold=0.08, want_counts=True)
    torch.cuda.synchronize()
print("done", name, wl["n"], fc.n_sweeps)

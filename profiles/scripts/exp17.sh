cd /root/repo
python - <<'P'
import torch, numpy as np, time
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
def ev(fn, reps=4):
    out=[]
    for _ in range(reps+2):
        e=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); r=fn(); e[1].record(); torch.cuda.synchronize(); out.append(e[0].elapsed_time(e[1]))
    return [round(x,3) for x in out], r
n=100000
frames = synth_video_cuda(n,64,64,seed=0)
pf = engine.pack_frames(frames)
ms, D1r = ev(lambda: engine.gram_l2_residues(pf,4),2); print("residue gram", ms)
for sym in (True, False):
    ms,_ = ev(lambda: engine.diag_filter_residues(D1r, n, 40, 4, p=0.7, symmetric=sym)); print("residue filter sym",sym, ms)
    ms,_ = ev(lambda: engine.diag_filter_residues(D1r, n, 40, 4, p=None, symmetric=sym)); print("residue filter (no D3) sym",sym, ms)
ms,_ = ev(lambda: engine.pack_frames(frames)); print("pack", ms)
P

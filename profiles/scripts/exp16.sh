cd /root/repo
python -m pytest tests/test_gpu_classic.py -x -q -m gpu 2>&1 | tail -3
python - <<'P'
import torch, numpy as np, time
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
def ev(fn, reps=5):
    out=[]
    for _ in range(reps+2):
        e=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); r=fn(); e[1].record(); torch.cuda.synchronize(); out.append(e[0].elapsed_time(e[1]))
    return float(np.median(out[2:])), r
for name,(n,h,w) in {"c2":(5000,224,224),"c5_25k":(25000,64,64),"c5":(100000,64,64)}.items():
    frames = synth_video_cuda(n,h,w,seed=0)
    pf = engine.pack_frames(frames)
    ok = engine.residue_eligible(pf,40,4)
    ms_r,_ = ev(lambda: engine.gram_l2_residues(pf,4),3)
    ms_all,_ = ev(lambda: engine.distance_filter(frames,40,4),3)
    print(f"{name}: eligible {ok}; residue gram {ms_r:.3f} ms; norms+gram+filter (residues) {ms_all:.3f} ms")
    if n <= 25000:
        ms_f,_ = ev(lambda: engine.distance_filter(frames,40,4,allow_residues=False),3)
        print(f"{name}: norms+gram+filter (full D1) {ms_f:.3f} ms")
    del frames, pf
    torch.cuda.empty_cache()
P

import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
from audio_video_textures_b200.classic.computeD1 import compute_D1
from audio_video_textures_b200.classic.computeD2 import compute_D2
from audio_video_textures_b200.classic.q_learning import LAST, q_learning
from audio_video_textures_b200.classic.video_textures import texture_walk
n=5000
frames = synth_video_cuda(n,224,224,seed=0)
host = frames.reshape(n,-1).cpu().pin_memory().view(n,224,224,3)
f = torch.tensor(4.5)
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
for it in range(6):
    flush.fill_(1); torch.cuda.synchronize()
    T=[time.perf_counter()]
    with contextlib.redirect_stdout(io.StringIO()):
        D1,P1,s1 = compute_D1(host, f, "RGB", slow=True, batch_size=48); T.append(time.perf_counter())
        torch.cuda.synchronize(); T.append(time.perf_counter())
        D2,P2,s2,_ = compute_D2(D1, f, filter_size=40, stride=4); T.append(time.perf_counter())
        D3n,P3,P3n,s3 = q_learning(D2, f, thresholding=0.08); T.append(time.perf_counter())
    rp,ci = engine.csr_from_matrix(P3n, LAST["counts"]); T.append(time.perf_counter())
    np.random.seed(0); walk,_ = texture_walk((rp,ci),3,30,30,4,40); T.append(time.perf_counter())
    sig = s3.item(); torch.cuda.synchronize(); T.append(time.perf_counter())
    names=["compute_D1(host call)","sync(D1 done)","compute_D2","q_learning","csr","walk","final sync"]
    print(it, " ".join(f"{k}={1e3*(b-a):.3f}" for k,a,b in zip(names,T[:-1],T[1:])), f"total={1e3*(T[-1]-T[0]):.3f}")

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_video_textures_b200 import engine
n = 16000
D1 = engine.empty_matrix(n, n, "cuda"); torch.manual_seed(0); D1.copy_(torch.rand(n, n, device="cuda") * 1000)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def run():
    ts = []
    for i in range(6):
        ev[0].record(); D2, D3 = engine.diag_filter(D1, 40, 1, p=0.7); ev[1].record(); torch.cuda.synchronize()
        if i >= 2: ts.append(ev[0].elapsed_time(ev[1]))
    return sum(ts) / len(ts), D2, D3
t, D2, D3 = run()
print("filter<40,1> ms", round(t, 3), "GB/s", round((4 * n * n + 8 * D2.shape[0] ** 2) / t / 1e6, 1), "mode", os.environ.get("AVTEX_FILTER_FFMA2", "packed"))
torch.save((D2[:64].cpu(), D3[:64].cpu()), "/tmp/f_%s.pt" % os.environ.get("AVTEX_FILTER_FFMA2", "1"))

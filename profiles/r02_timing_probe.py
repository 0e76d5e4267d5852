"""Why do event-timed kernels run slower than the same launches under ncu?  Times the stride-4 filter on the C5
shard shape three ways (single launch per event pair, 10 back-to-back launches, a 2 s loop with nvidia-smi
sampling SM clock and power at 20 ms) with PREALLOCATED outputs."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_video_textures_b200 import _lib, engine

rows, n, fs, s = 12576, 100000, 40, 4
mode = sys.argv[1] if len(sys.argv) > 1 else "real"
D1 = torch.empty((rows, n), dtype=torch.float32, device="cuda")
if mode == "real":
    D1.uniform_(100.0, 20000.0)
else:
    D1.fill_(1000.0)
m = (n - fs) // s + 1
ro = (rows - fs) // s + 1
D2 = engine.empty_matrix(ro, m, "cuda")
D3 = engine.empty_matrix(ro, m, "cuda")
taps = engine.binomial_taps(fs)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch():
    _lib.call("avtex_diag_filter_pow", _lib.ptr(D1), D1.stride(0), 0, rows, taps.ctypes.data_as(C.POINTER(C.c_float)), fs, s,
              0, ro, m, _lib.ptr(D2), D2.stride(0), _lib.ptr(D3), D3.stride(0), C.c_float(0.7), None, None, 0, st)


def ev_pair():
    return [torch.cuda.Event(enable_timing=True) for _ in range(2)]


for _ in range(3):
    launch()
torch.cuda.synchronize()
single = []
for _ in range(10):
    e = ev_pair()
    e[0].record(); launch(); e[1].record()
    torch.cuda.synchronize()
    single.append(e[0].elapsed_time(e[1]))
e = ev_pair()
e[0].record()
for _ in range(10):
    launch()
e[1].record()
torch.cuda.synchronize()
print(f"{mode}: single launches (ms): {[round(v, 3) for v in single]}")
print(f"{mode}: 10 back to back: {e[0].elapsed_time(e[1]) / 10:.3f} ms each")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader",
                      "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.perf_counter()
e = ev_pair()
e[0].record()
k = 0
while time.perf_counter() - t0 < 2.0:
    for _ in range(20):
        launch()
    k += 20
    torch.cuda.synchronize()
e[1].record()
torch.cuda.synchronize()
p.terminate()
lines = p.stdout.read().strip().split("\n")
print(f"{mode}: 2 s loop: {e[0].elapsed_time(e[1]) / k:.3f} ms each over {k} launches")
print("nvidia-smi samples (sm MHz, mem MHz, W, reasons):", lines[::8][:14])

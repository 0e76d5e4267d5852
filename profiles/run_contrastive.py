"""Contrastive synthesis (BASELINE configs 3 and 4) timing + parity against the CPU oracle.
    python profiles/run_contrastive.py [L] [D] [A] [steps_cpu]
Prints one JSON line per configuration: ms/step, window-pairs/s, GEMV GB/s (CUDA events), and the
oracle's CPU time per step on the same inputs (first `steps_cpu` steps) with the chosen windows compared."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_video_textures_b200 import engine
from audio_video_textures_b200.contrastive.validate import SynthesisState, synthesize
from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings
from oracle import contrastive as oc

L = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2304
A_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 128, 12288]
steps_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 5
HBM = 6547.8
fps, nvl, W, S, temp, th = 30, 30, 15, 6, 0.1, 0.3

emb = synth_embeddings(L, D, seed=0, device="cuda")
for A in A_list:
    kw = {}
    if A:
        qa = synth_audio_features(L, A, seed=0, device="cuda")
        das = synth_audio_features(L, A, seed=1, device="cuda")
        dad = synth_audio_features(160, A, seed=2, device="cuda")
        kw = dict(alpha=0.5, q_audio=qa, da_source=das, da_driving=dad)
    np.random.seed(0)
    synthesize(emb, temp=temp, threshold=th, fps=fps, new_video_length=nvl, window=W, stride=S, **kw)   # warm-up
    torch.cuda.synchronize()
    np.random.seed(0)
    t0 = time.perf_counter()
    res = synthesize(emb, temp=temp, threshold=th, fps=fps, new_video_length=nvl, window=W, stride=S, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = len(res["q_ids"])
    # GEMV alone, CUDA events, table larger than L2 or L2 flushed
    st = SynthesisState(emb, None, kw.get("q_audio"), None, kw.get("da_source"), kw.get("da_driving"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for i in range(13):
        flush.fill_(1)
        ev[0].record()
        engine.cosine_scores(st.tn, st.qn[100 + i], temp, out=st.o)
        ev[1].record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(ev[0].elapsed_time(ev[1]))
    gemv_bytes = 4.0 * L * st.tn.shape[1]
    # CPU oracle on the same inputs, first steps only (bounded), sequence compared
    cpu_kw = {k: v.cpu() for k, v in kw.items() if torch.is_tensor(v)}
    if A:
        cpu_kw["alpha"] = 0.5
    np.random.seed(0)
    t1 = time.perf_counter()
    fps_cpu = 1
    want = oc.synthesize(emb.cpu(), temp, th, 150, fps_cpu, W + (steps_cpu - 1) * S, W, S,
                         q_start=res["start"], return_debug=True, **cpu_kw)
    cpu_dt = (time.perf_counter() - t1) / len(want["q_ids"])
    n_cmp = len(want["q_ids"])
    same = res["q_ids"][:n_cmp] == want["q_ids"]
    print(json.dumps({
        "config": f"contrastive synthesis L={L} D={D} A={A} (-e -th {th} -temp {temp}" + (" -m 2 -alpha 0.5)" if A else ")"),
        "steps": steps, "ms_per_step": 1e3 * dt / steps, "window_pairs_per_s": steps * L / dt,
        "gemv_ms": float(np.mean(ms)), "gemv_GBps": gemv_bytes / (np.mean(ms) * 1e-3) / 1e9,
        "gemv_frac_of_measured_hbm": gemv_bytes / (np.mean(ms) * 1e-3) / 1e9 / HBM,
        "cpu_oracle_ms_per_step": 1e3 * cpu_dt, "cpu_cores": os.cpu_count(), "cpu_steps_timed": n_cmp,
        "first_steps_identical_to_oracle": bool(same), "oracle_min_margin": float(min(want["margins"])),
        "nz_mean": float(np.mean(res["nz_counts"])), "jump_count": res["jump_count"]}))

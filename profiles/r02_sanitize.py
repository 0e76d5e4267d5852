"""Small run of every hand-rolled synchronisation protocol for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck  python profiles/r02_sanitize.py
    compute-sanitizer --tool racecheck python profiles/r02_sanitize.py
  * gram_l2_s8_2cta_kernel: symmetric full matrix (ragged 777 frames), a row block, a 3-job list with direct +
    transposed destinations (the multi-GPU shape), signed and unsigned operands  -> TMA / mbarrier / TMEM pipeline
  * future_cost_fused_kernel (cooperative, grid barrier between sweeps)
  * future_cost_fused_peer_kernel on 2 and 4 VIRTUAL ranks (flag barrier between concurrently resident kernels)
  * synthesis_step_kernel (dynamic row tickets + last-CTA selection, mapped pinned result)
  * filter (general, symmetric-mirror and residue-plane forms) / finalize / probabilities / CSR compaction
  * residue-class Gram jobs (k_off / strided norms, 4 symmetric jobs in one launch) and the 8-virtual-rank residue shards
Results are checked against each other so a sanitizer-induced slowdown cannot hide a wrong answer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_video_textures_b200 import dist as avd
from audio_video_textures_b200 import engine, selfcheck
from audio_video_textures_b200.synth import synth_embeddings, synth_video

frames = synth_video(777, 12, 12, seed=1).cuda()
pf = engine.pack_frames(frames)
D1 = engine.gram_l2(pf)
blk = engine.gram_l2(pf, 300, 200)
assert torch.equal(blk, D1[300:500])
out = engine.empty_matrix(777, 777, "cuda").zero_()
ptr, ld = out.data_ptr(), out.stride(0)
jobs = [dict(row0=0, rows=400, col0=0, cols=400, symmetric=1, count_stats=0, D=ptr, d_row0=0, ldd=ld, DT=ptr, dt_row0=0, ldt=ld),
        dict(row0=400, rows=377, col0=400, cols=377, symmetric=1, count_stats=0, D=ptr, d_row0=0, ldd=ld, DT=ptr, dt_row0=0, ldt=ld),
        dict(row0=0, rows=400, col0=400, cols=377, symmetric=0, count_stats=0, D=ptr, d_row0=0, ldd=ld, DT=ptr, dt_row0=0, ldt=ld)]
engine.gram_l2_jobs(pf, jobs)
assert torch.equal(out, D1)
pfd = engine.pack_frames(frames, defer_norms=True)             # K0 fused into the Gram launch (counter barrier)
assert torch.equal(engine.gram_l2(pfd), D1) and torch.equal(pfd.sqnorm, pf.sqnorm)
pfs = engine.pack_frames(frames.float())                       # centred s8 operand path
assert torch.equal(engine.gram_l2(pfs), D1)
single = selfcheck.single_gpu_pipeline(frames, 16, 1, 4.5, 0.08)
loop = engine.future_cost(single["D3"])
assert torch.equal(loop.mvec[:single["D3"].shape[0]], single["fc"].mvec) and loop.n_sweeps == single["fc"].n_sweeps
rowptr, colidx = engine.csr_from_matrix(single["P3n"], single["counts"])
for world in (2, 4):
    box = avd.VirtualBox(777, 16, 1, world, frames.device)
    for _ in range(2):
        res = box.step(frames, sigma_factor=4.5, threshold=0.08)
    for r in res:
        ok = selfcheck.shard_equals_single(r, single)
        assert all(ok.values()), ok
# symmetric filter form (shared-memory mirror tile) and the residue-class pipeline (K % 128 == 0, N % 4 == 0)
S2, S3 = engine.diag_filter(D1, 16, 1, p=0.7, symmetric=True)
assert torch.equal(S2, single["D2"]) and torch.equal(S3, single["D3"])
fr4 = synth_video(776, 16, 8, seed=2).cuda()                   # K = 384
pf4 = engine.pack_frames(fr4)
full = engine.diag_filter(engine.gram_l2(pf4), 40, 4, p=0.7, symmetric=False)
D1r = engine.gram_l2_residues(pf4, 4)
for sym in (False, True):
    R2, R3 = engine.diag_filter_residues(D1r, 776, 40, 4, p=0.7, symmetric=sym)
    assert torch.equal(R2, full[0]) and torch.equal(R3, full[1])
single4 = selfcheck.single_gpu_pipeline(fr4, 40, 4, 4.5, 0.08)
box = avd.VirtualBox(776, 40, 4, 8, fr4.device, residues=True)
for r in box.step(fr4, sigma_factor=4.5, threshold=0.08):
    ok = selfcheck.shard_equals_single(r, single4)
    assert all(ok.values()), ok
emb = synth_embeddings(600, 96, seed=0).cuda()
tn = engine.l2_normalize_rows(emb)
ws = engine.SynthesisWorkspace(600, "cuda")
o = engine.cosine_scores(tn, tn[10], 0.1)
sel = torch.zeros(601, dtype=torch.int32, device="cuda")
engine.select_step(o, None, 10, 0.5, 0.3, sel[1:], sel[:1], None)
for _ in range(3):
    got = engine.synthesis_step(ws, tn, tn[10], 10, 0.1, 0.5, 0.3)
    assert np.array_equal(got, sel[1:int(sel[0]) + 1].cpu().numpy())
torch.cuda.synchronize()
print("sanitize driver ok")

"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz.

Run in the BUILD CONTAINER (needs /root/reference):  python -m oracle.make_golden

Every matrix / scalar stored under a `ref_` key is the output of the UNMODIFIED reference
function (compute_D1 / compute_D2 / q_learning / ContrastivePredictionTemporal.forward) imported
through oracle/ref_shim.py on the stored seeded input.  Keys under `walk_` / `synth_` come from
the restated walk / selection loop (the reference's own `audio_video_texture` and `validate`
cannot be imported) applied to the reference's P3_new / logits.
The inputs are stored too: torch's CPU `sin`/`randn` are not guaranteed bit-stable across CPU
types, so fixtures must not be regenerated from the seed on another machine.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings, synth_video
from oracle import classic, contrastive, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CLASSIC_CASES = [
    # name, N, H, W, fs, stride, model_type, threshold, sigma_factor, walk (fps, nvl)
    dict(name="classic_small_m1", N=150, H=16, W=16, fs=8, stride=1, m=1, th=0.08, f=4.5, fps=30, nvl=10),
    dict(name="classic_ragged_m2", N=333, H=12, W=20, fs=16, stride=1, m=2, th=0.08, f=4.52, fps=30, nvl=10),
    dict(name="classic_stride_m3", N=520, H=8, W=8, fs=40, stride=4, m=3, th=0.08, f=4.55, fps=30, nvl=10),
    dict(name="classic_c1", N=300, H=64, W=64, fs=40, stride=1, m=1, th=0.08, f=4.5, fps=30, nvl=30),
]


def _pick_seed(case, min_margin=2e-5, max_seed=20):
    """Screen seeds: the survivor sets must not hinge on an element closer than `min_margin`
    (relative) to a threshold cut."""
    cD1, cD2, ql = ref_shim.load_classic()
    best = None
    for seed in range(max_seed):
        video = synth_video(case["N"], case["H"], case["W"], seed=seed)
        frames = video.float()
        f = torch.tensor(case["f"], dtype=torch.float32)
        D1 = classic.pairwise_l2(frames)
        D2 = classic.compute_D2(D1, f, case["fs"], case["stride"])[0]
        D3_new, P3, P3_new, sigma = classic.q_learning(D2, f, thresholding=case["th"])
        mg = classic.threshold_margin(P3, case["th"])
        if best is None or mg > best[1]:
            best = (seed, mg)
        if mg >= min_margin:
            return seed, mg
    return best


def make_classic(case):
    cD1, cD2, ql = ref_shim.load_classic()
    seed, margin = _pick_seed(case)
    video = synth_video(case["N"], case["H"], case["W"], seed=seed)
    frames = video.float()
    f = torch.tensor(case["f"], dtype=torch.float32)
    D1, P1, s1 = cD1(frames, f, "RGB", slow=True, batch_size=48)
    if case["m"] in (1, 2):
        D2, P2, s2, bfilt = cD2(D1, f, filter_size=case["fs"])
    else:
        D2, P2, s2, bfilt = cD2(D1, f, filter_size=case["fs"], stride=case["stride"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        D3_new, P3, P3_new, s3 = ql(D2, f, thresholding=case["th"])
    n_sweeps = buf.getvalue().count("Eps:")
    # restated pieces, bit-compared with the reference here
    o = classic.q_learning(D2, f, thresholding=case["th"], return_trail=True)
    assert torch.equal(o[0], D3_new) and torch.equal(o[2], P3_new) and len(o[4]) == n_sweeps
    assert torch.equal(classic.pairwise_l2(frames), D1)
    rows, cols = torch.nonzero(P3_new, as_tuple=True)
    counts = torch.bincount(rows, minlength=P3_new.shape[0])
    np.random.seed(seed)
    wl, jc = classic.walk(P3_new, case["m"], case["fps"], case["nvl"], case["stride"], case["fs"])
    out = dict(
        seed=seed, margin=margin, video=video.numpy(), sigma_factor=np.float32(case["f"]),
        fs=case["fs"], stride=case["stride"], model_type=case["m"], threshold=case["th"],
        fps=case["fps"], nvl=case["nvl"],
        ref_D1=D1.numpy(), ref_sigma1=s1.numpy(), ref_D2=D2.numpy(), ref_sigma2=s2.numpy(),
        ref_D3_new=D3_new.numpy(), ref_sigma3=s3.numpy(), ref_P3=P3.numpy(),
        ref_filter_diag=torch.diagonal(bfilt.view(case["fs"], case["fs"])).numpy(),
        ref_n_sweeps=n_sweeps, eps_trail=np.array(o[4], dtype=np.float64),
        ref_P3new_rowptr=torch.cat((torch.zeros(1, dtype=torch.long), counts.cumsum(0))).numpy(),
        ref_P3new_cols=cols.numpy().astype(np.int32),
        ref_P1_row0=P1[0].numpy(), ref_P2_row0=P2[0].numpy(),
        walk_frames=np.array(wl, dtype=np.int64), walk_jump_count=jc,
    )
    path = os.path.join(OUT, case["name"] + ".npz")
    np.savez_compressed(path, **out)
    print(f"{case['name']}: seed {seed} margin {margin:.3g} sweeps {n_sweeps} M {D2.shape[0]} "
          f"nnz/row {counts.float().mean():.1f} jumps {jc} -> {os.path.getsize(path)/1e6:.2f} MB")


def make_contrastive():
    L, D, A, temp, th, mbs = 96, 48, 24, 0.1, 0.3, 16
    best = None
    for seed in range(20):
        emb = synth_embeddings(L, D, seed=seed)
        np.random.seed(seed)
        res = contrastive.synthesize(emb, temp, th, mbs, fps=30, new_video_length=4, window=15,
                                     stride=6, q_start=10, return_debug=True)
        mg = min(res["margins"])
        if best is None or mg > best[1]:
            best = (seed, mg)
        if mg > 2e-5:
            break
    seed = best[0]
    emb = synth_embeddings(L, D, seed=seed)
    model = ref_shim.load_contrastive_model(temp, mbs)
    # reference forward on every chunk of two query steps
    ref_logits = []
    for q in (10, L - 1):
        ids = contrastive.target_order(q, L)
        T = emb[ids]
        n_chunks = -(-len(ids) // mbs)
        padded = torch.zeros((n_chunks * mbs, D))
        padded[:len(ids)] = T
        outs = [ref_shim.reference_chunk_scores(model, emb[q], padded[c * mbs:(c + 1) * mbs])
                for c in range(n_chunks)]
        ref = torch.cat(outs)[:len(ids)]
        mine, _ = contrastive.step_scores(emb[q], emb, ids, temp, mbs)
        assert torch.equal(ref, mine), float((ref - mine).abs().max())
        ref_logits.append(ref.numpy())
    np.random.seed(seed)
    r1 = contrastive.synthesize(emb, temp, th, mbs, fps=30, new_video_length=4, window=15, stride=6,
                                q_start=10, return_debug=True)
    # audio-conditioned (model_type 2 + driving audio)
    qa = synth_audio_features(L - 5, A, seed=seed)            # shorter than L: exercises the index clamp
    das = synth_audio_features(L - 5, A, seed=seed + 1)
    dad = synth_audio_features(40, A, seed=seed + 2)
    start = contrastive.start_segment(das, dad[0])
    np.random.seed(seed)
    r2 = contrastive.synthesize(emb, temp, th, mbs, fps=30, new_video_length=4, window=15, stride=6,
                                q_start=start, alpha=0.5, q_audio=qa, da_source=das, da_driving=dad,
                                return_debug=True)
    path = os.path.join(OUT, "contrastive_small.npz")
    np.savez_compressed(
        path, seed=seed, emb=emb.numpy(), temp=temp, th=th, mbs=mbs, fps=30, nvl=4, window=15, stride=6,
        ref_logits_q10=ref_logits[0], ref_logits_qlast=ref_logits[1],
        synth_q_ids=np.array(r1["q_ids"]), synth_frame_ids=np.array(r1["frame_ids"]),
        synth_nz=np.array(r1["nz_counts"]), synth_jumps=r1["jump_count"], synth_margin=min(r1["margins"]),
        q_audio=qa.numpy(), da_source=das.numpy(), da_driving=dad.numpy(), audio_start=start,
        synth2_q_ids=np.array(r2["q_ids"]), synth2_frame_ids=np.array(r2["frame_ids"]),
        synth2_nz=np.array(r2["nz_counts"]), synth2_jumps=r2["jump_count"], synth2_margin=min(r2["margins"]),
    )
    print(f"contrastive_small: seed {seed} steps {len(r1['q_ids'])} margin {min(r1['margins']):.3g} / "
          f"{min(r2['margins']):.3g} nz {np.mean(r1['nz_counts']):.1f}/{np.mean(r2['nz_counts']):.1f} "
          f"start {start} -> {os.path.getsize(path)/1e6:.2f} MB")


def main():
    if not ref_shim.available():
        sys.exit("needs /root/reference (build container)")
    os.makedirs(OUT, exist_ok=True)
    for case in CLASSIC_CASES:
        make_classic(case)
    make_contrastive()


if __name__ == "__main__":
    main()

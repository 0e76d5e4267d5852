"""The `-e` synthesis loop of contrastive_video_textures/validate.py:218-242, 324-690 restricted to
its hot-path lines, at the embedding boundary (per-window embeddings in, window/frame ids out).

Per step the reference re-encodes ALL L target windows through a 3D CNN on every GPU
(validate.py:442-493).  Here the normalised embedding tables stay resident in HBM and one step is
    K6  o = <q^, T^> / temp            streaming GEMV, 4 L D bytes
    K6  a = <d^_step, S^> / temp       (audio-conditioned only)
    K7  o/sum, a/sum, alpha-mix, threshold at max - th*max, renormalise, ordered survivor list
all inside ONE cooperative launch (avtex_synthesis_step) that writes the survivor list into mapped pinned
memory; the uniform draw stays `np.random.choice` on the host so the chosen sequence is bit-identical to the
reference under the same seed.
`mini_batchsize` (-mbs) only shapes the reference's DataParallel chunks (validate.py:409-411,
utils.py:208-260); at the embedding boundary the result layout is contiguous, so it is accepted
and ignored.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import engine

HOST_CAP = 4096          # survivors copied with the count in one D2H; more triggers a second copy


def start_segment(da_source: torch.Tensor, driving0: torch.Tensor) -> int:
    """validate.py:218-242 — 10 without driving audio; else first arg-max cosine similarity (> 0)."""
    return engine.audio_start(da_source.float(), driving0.float().reshape(-1))


class SynthesisState:
    """Resident tables + scratch for the step kernels."""

    def __init__(self, t_emb, q_emb=None, q_audio=None, t_audio=None, da_source=None, da_driving=None):
        dev = t_emb.device
        L = t_emb.shape[0]
        self.L = L

        def clamp_rows(a):                 # audio tables may be shorter than L: index min(i, La-1)
            if a.shape[0] >= L:
                return a[:L]
            idx = torch.clamp(torch.arange(L, device=dev), max=a.shape[0] - 1)
            return a[idx]

        t = t_emb.float()
        q = t if q_emb is None else q_emb.float()
        if q_audio is not None:            # model_type 2: cat(video, audio) before normalising
            ta = clamp_rows((q_audio if t_audio is None else t_audio).float())
            qa = clamp_rows(q_audio.float())
            same = q_emb is None and t_audio is None
            t = torch.cat((t, ta), dim=1)
            q = t if same else torch.cat((q, qa), dim=1)
        self.tn = engine.l2_normalize_rows(t)
        self.qn = self.tn if q is t else engine.l2_normalize_rows(q)
        self.sn = self.dn = None
        if da_driving is not None:
            self.sn = engine.l2_normalize_rows(clamp_rows(da_source.float()))
            self.dn = engine.l2_normalize_rows(da_driving.float())
        self.ws = engine.SynthesisWorkspace(L, dev, HOST_CAP)
        self.vals = None
        self.launches_per_step = 1

    def step(self, q_id: int, iter_count: int, temp, alpha, threshold, want_vals=False):
        """Returns the survivor window ids (numpy int32) in the reference's target-list order.  ONE kernel launch
        (scores + audio scores + mix + threshold + ordered survivor list); the list arrives in mapped pinned
        memory, the host polls its sequence word."""
        if want_vals and self.vals is None:
            self.vals = torch.zeros(self.L, dtype=torch.float32, device=self.tn.device)
        return engine.synthesis_step(self.ws, self.tn, self.qn[q_id], q_id, temp, alpha, threshold,
                                     self.sn, self.dn[iter_count] if self.dn is not None else None,
                                     self.vals if want_vals else None)


def planned_steps(max_length, window: int, stride: int, subsample_rate: int = 1, max_steps=None) -> int:
    """Number of iterations of `while len(new_frames) < max_length` (cvt/validate.py:324): the first step emits
    `window` frames, every later one `stride` (validate.py:581-612), whatever is chosen — so the trip count is
    known before the loop starts, which is what lets the whole loop run on the device."""
    n_frames, steps = 0, 0
    while n_frames < max_length and (max_steps is None or steps < max_steps):
        n_frames += (window if steps == 0 else stride) * subsample_rate
        steps += 1
    return steps


def synthesize(t_emb, temp=0.1, threshold=0.0, fps=30, new_video_length=30, window=20, stride=4,
               q_start=None, alpha=0.5, mini_batchsize=150, q_emb=None, q_audio=None, t_audio=None,
               da_source=None, da_driving=None, subsample_rate=1, max_steps=None, device_loop=True):
    """Runs the synthesis loop.  Tensors are CUDA (or are moved there).  Consumes the numpy global RNG
    once per step — with `device_loop` (default) ON THE DEVICE, inside one persistent kernel, from numpy's own
    generator state, which is handed back afterwards (engine.synthesis_loop); `device_loop=False` keeps one
    launch per step with the draw on the host.  Both give the same sequence.
    Returns dict(q_ids, frame_ids, jump_count, nz_counts, start)."""
    dev = t_emb.device if t_emb.is_cuda else torch.device("cuda")
    mv = lambda x: None if x is None else x.to(dev)
    st = SynthesisState(mv(t_emb), mv(q_emb), mv(q_audio), mv(t_audio), mv(da_source), mv(da_driving))
    if q_start is None:
        q_start = 10 if da_driving is None else start_segment(mv(da_source), mv(da_driving)[0])
    W, S = window, stride
    max_length = math.ceil(fps) * new_video_length
    if da_driving is not None:                                      # validate.py:260-263: a short driving clip caps the output
        max_length = min(max_length, np.ceil(fps) * np.floor(len(da_driving) * S + W))
    q_id, p_q_id, iter_count, n_frames, jump_count = q_start, -1, 1, 0, 0
    q_ids, frame_ids, nz_counts = [], [], []
    chosen = None
    if device_loop:
        n_steps = planned_steps(max_length, W, S, subsample_rate, max_steps)
        if n_steps > 0:
            chosen, nz = engine.synthesis_loop(st.ws, st.tn, st.qn, q_start, n_steps, temp, alpha, threshold, st.sn, st.dn)
    while n_frames < max_length and (max_steps is None or len(q_ids) < max_steps):
        if chosen is not None:
            nz_counts.append(int(nz[len(q_ids)]))
            q_id = int(chosen[len(q_ids)])
        else:
            choices = st.step(q_id, iter_count, temp, alpha, threshold)
            nz_counts.append(len(choices))
            q_id = int(np.random.choice(choices))                   # validate.py:570-572
        if p_q_id == -1:                                            # validate.py:581-612
            diff = range(q_id * S, q_id * S + W)
        else:
            if q_id != p_q_id + 1:
                jump_count += 1
            diff = range(q_id * S + (W - S), q_id * S + W)
        frame_ids.extend(diff)
        n_frames += len(diff) * subsample_rate
        q_ids.append(q_id)
        iter_count += 1
        p_q_id = q_id
    return dict(q_ids=q_ids, frame_ids=frame_ids, jump_count=jump_count, nz_counts=nz_counts, start=q_start)


def validate(args, t_emb, q_emb=None, q_audio=None, t_audio=None, da_source=None, da_driving=None):
    """`validate(model, args, ...)` of the reference, at the embedding boundary: reads
    args.{temp, threshold, alpha, mini_batchsize, fps, new_video_length, window, stride, subsample_rate}."""
    return synthesize(t_emb, temp=args.temp, threshold=args.threshold, fps=args.fps,
                      new_video_length=args.new_video_length, window=args.window, stride=args.stride,
                      alpha=args.alpha, mini_batchsize=args.mini_batchsize, q_emb=q_emb, q_audio=q_audio,
                      t_audio=t_audio, da_source=da_source, da_driving=da_driving,
                      subsample_rate=getattr(args, "subsample_rate", 1))

"""(f1) Window construction and the per-window embedding cache in front of the synthesis loop.

The reference re-encodes ALL L target windows through its 3D CNN on EVERY synthesis step
(cvt/validate.py:380-395, 442-493; cvt/utils/utils.py:233-260; cvt/models/models.py:355-402): ~150 x L encoder
passes for a 30 s texture.  The window embeddings do not depend on the step, so here they are computed ONCE:
the clip stays on the device, the [B, W, ...] window tensors are assembled by a gather kernel straight from it
(avtex_gather_rows) and pushed through the caller's encoder in batches; the resulting [L, D] table is what
`validate.synthesize` consumes (L2-normalised once, resident in HBM).

The reference's own per-step target construction is reproduced as an INDEX PLAN (`reference_step_plan`) for
parity checking, including its defects: target frames are the union of the target
windows' frames in first-seen order — the positive's frames first — and chunk `c` starts at c*S*(mbs-1) while
emitting mbs windows, so the "windows" the encoder sees drift and are not the clip's true windows.  The cache
uses the TRUE windows w*S .. w*S+W (what the embeddings mean); `reference_step_embeddings` re-encodes the
reference's layout for one step when bit-level comparison with the reference's logits is wanted.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib, engine


def num_windows(n_frames: int, window: int, stride: int) -> int:
    """L = number of windows w with w*S + W <= n_frames."""
    return 0 if n_frames < window else (n_frames - window) // stride + 1


def window_plan(L: int, window: int, stride: int) -> np.ndarray:
    """True windows: plan[w, t] = w*S + t."""
    return (np.arange(L, dtype=np.int32)[:, None] * stride + np.arange(window, dtype=np.int32)[None, :]).astype(np.int32)


def reference_target_frames(q_id: int, L: int, window: int, stride: int):
    """cvt/validate.py:369-388: target segment ids [pos] ++ ascending(rest) and the frame ids of their windows,
    duplicates removed keeping the first occurrence."""
    pos_id = min(q_id + 1, L - 1)
    mask = np.ones(L, dtype=bool)
    mask[[q_id, pos_id]] = False
    seg = np.concatenate((np.array([pos_id]), np.arange(L)[mask]), axis=0)
    frames = (seg[:, None] * stride + np.arange(window)[None, :]).reshape(-1)
    _, first = np.unique(frames, return_index=True)
    return seg, frames[np.sort(first)]


def reference_step_plan(q_id: int, L: int, window: int, stride: int, mini_batchsize: int):
    """Index plan of ONE reference step: which clip frame every slot of every encoder window holds.
    Returns (target_segment_ids, plan int32 [n_chunks * mbs, W]) with -1 for the zero padding:
      chunk c = target_frames[c*S*(mbs-1) : c*S*(mbs-1) + mbs*S + W]       (cvt/utils/utils.py:248-258)
      window i of a chunk = chunk[i*S : i*S + W], i < mbs                  (cvt/models/models.py:355-362)
    Logits of window (c, i) land at output position c*mbs + i (cvt/validate.py:481-493)."""
    seg, tframes = reference_target_frames(q_id, L, window, stride)
    n_in = len(tframes)
    total_segments = math.ceil((n_in - window) / stride)
    chunk_size = mini_batchsize * stride + window
    batch_size = math.ceil(total_segments / mini_batchsize)
    plan = np.full((batch_size * mini_batchsize, window), -1, dtype=np.int32)
    for c in range(batch_size):
        start = c * stride * (mini_batchsize - 1)
        chunk = np.full(chunk_size, -1, dtype=np.int64)
        end = min(start + chunk_size, n_in)
        chunk[:end - start] = tframes[start:end]
        for i in range(mini_batchsize):
            w = chunk[i * stride: i * stride + window]
            plan[c * mini_batchsize + i, :len(w)] = w
    return seg, plan


def gather_windows(frames: torch.Tensor, plan) -> torch.Tensor:
    """frames: CUDA [T, ...] (any dtype, contiguous rows); plan: int32 [n, W] of frame ids (-1 = zeros).
    Returns [n, W, ...] assembled on the device by avtex_gather_rows."""
    if not frames.is_cuda:
        raise ValueError("frames must live on the device (the clip is uploaded once)")
    x = frames.reshape(frames.shape[0], -1)
    if x.stride(1) != 1:
        x = x.contiguous()
    plan_t = torch.as_tensor(np.ascontiguousarray(plan, dtype=np.int32)).to(frames.device)
    n, w = plan_t.shape
    row_bytes = x.shape[1] * x.element_size()
    out = torch.empty((n, w) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    _lib.call("avtex_gather_rows", _lib.ptr(x), row_bytes, x.stride(0) * x.element_size(), x.shape[0], _lib.ptr(plan_t),
              n * w, _lib.ptr(out), engine._dev(x), engine._stream(x))
    return out


class EmbeddingCache:
    """Per-window embedding table [L, D], computed once.

    frames   CUDA tensor [T, ...] — the whole clip on the device
    encoder  callable: [B, W, ...] window batch -> [B, D] embeddings (e.g. the reference's q/t encoder stack;
             the 3D CNNs themselves are feature producers outside this repo)
    """

    def __init__(self, frames: torch.Tensor, encoder, window: int, stride: int, batch: int = 64):
        self.frames, self.encoder, self.window, self.stride, self.batch = frames, encoder, window, stride, batch
        self.L = num_windows(frames.shape[0], window, stride)
        if self.L < 2:
            raise ValueError("clip too short for two windows")
        self.encoder_calls = 0
        self.table = self._encode(window_plan(self.L, window, stride))

    def _encode(self, plan) -> torch.Tensor:
        out = []
        with torch.no_grad():
            for b0 in range(0, len(plan), self.batch):
                wins = gather_windows(self.frames, plan[b0:b0 + self.batch])
                out.append(self.encoder(wins).reshape(wins.shape[0], -1).float())
                self.encoder_calls += 1
        return torch.cat(out, dim=0)

    def reference_step_embeddings(self, q_id: int, mini_batchsize: int):
        """The embeddings the REFERENCE would compute at one step (its scrambled chunk layout), re-encoded:
        returns (target_segment_ids, q_embedding [D], t_embeddings [n_chunks*mbs, D])."""
        seg, plan = reference_step_plan(q_id, self.L, self.window, self.stride, mini_batchsize)
        q = self._encode(window_plan(self.L, self.window, self.stride)[q_id:q_id + 1])[0]
        return seg, q, self._encode(plan)

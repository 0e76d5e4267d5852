"""Drop-in for baselines/classic_video_textures/computePaudio.py:6-18 (the classic-side audio prior;
SURVEY.md §8(f) item 3 — the same cosine-similarity kernels as the contrastive driving-audio term)."""
from __future__ import annotations

import torch

from .. import engine


def compute_Paudio(t_audio_eg: torch.Tensor, driving_audio: torch.Tensor) -> torch.Tensor:
    """p_audio[i] = cos(normalize(driving_audio), normalize(t_audio_eg[i])) / (sum + 1e-6).

    t_audio_eg [T, A], driving_audio [A].  The reference normalises both operands and then takes a
    CosineSimilarity of the already unit vectors (a second normalisation that changes the value by ~1e-7);
    here the two are one dot product of the normalised rows."""
    s_a = t_audio_eg if t_audio_eg.is_cuda else t_audio_eg.cuda()
    d_a = driving_audio if driving_audio.is_cuda else driving_audio.cuda()
    sn = engine.l2_normalize_rows(s_a.float().reshape(s_a.shape[0], -1))
    dn = engine.l2_normalize_rows(d_a.float().reshape(1, -1))
    cos = engine.cosine_scores(sn, dn[0], 1.0)
    return cos / (cos.sum() + 1e-6)

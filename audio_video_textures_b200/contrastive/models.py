"""Similarity tail of `ContrastivePredictionTemporal.forward`
(contrastive_video_textures/models/models.py:351-352, 412-417, 419-457) on libavtex kernels.

The encoders of the reference class are feature producers outside the hot path (SURVEY.md §2.1
rows 9-11); this module starts where they end: per-window embeddings.
"""
from __future__ import annotations

import torch

from .. import engine


def similarity_tail(q: torch.Tensor, t: torch.Tensor, temp: float) -> torch.Tensor:
    """q [B, D], t [B, T, D] -> output [B, T] = bmm(normalize(q), normalize(t)^T) / temp
    (models.py:351-352, 412-417)."""
    B, T, D = t.shape
    qn = engine.l2_normalize_rows(q.reshape(B, D).float())
    tn = engine.l2_normalize_rows(t.reshape(B * T, D).float()).view(B, T, D)
    out = torch.empty((B, T), dtype=torch.float32, device=t.device)
    for b in range(B):
        engine.cosine_scores(tn[b], qn[b], temp, out=out[b])
    return out


def audio_tail(d_a: torch.Tensor, s_a: torch.Tensor, temp: float) -> torch.Tensor:
    """d_a [B, A], s_a [B, T, A] -> output_a [B, 1, T] (models.py:433-439, 457)."""
    return similarity_tail(d_a, s_a, temp).unsqueeze(1)


class ContrastivePredictionTemporal(torch.nn.Module):
    """Same constructor keywords as the reference class (models.py:234-249) for the arguments the
    tail uses; `forward` takes EMBEDDINGS where the reference takes frames:
        q_f [B, D], t_f [B, T, D], q_audio_eg [B, A] / t_audio_eg [B, T, A] (model_type 2: concatenated
        before the normalisation, models.py:347,408), driving_audio [B, A'] with `da_model` applied to
        (t_audio_eg, driving_audio) when given (models.py:424-431), else raw features (:445-455)."""

    def __init__(self, q_image_enc_model=None, t_image_enc_model=None, audio_enc_model=None, model_type=1,
                 fc_dim=0, temp=0.1, window=20, stride=2, threshold=0.20, mini_batchsize=20, dropout=0.5,
                 enc_arch="precomputed", img_size=224):
        super().__init__()
        self.temp, self.window, self.stride = temp, window, stride
        self.threshold, self.mini_batchsize, self.model_type = threshold, mini_batchsize, model_type

    def forward(self, q_f, t_f, q_audio_eg=None, t_audio_eg=None, is_inference=False, driving_audio=None,
                da_model=None, da_feats=None, cam_viz=False):
        if self.model_type == 2:
            q = torch.cat((q_f, q_audio_eg), dim=1)
            t = torch.cat((t_f, t_audio_eg), dim=2)
        else:
            q, t = q_f, t_f
        output = similarity_tail(q, t, self.temp)
        if driving_audio is None:
            return output
        s_a, d_a = t_audio_eg, driving_audio
        if da_model is not None:
            B, T = s_a.shape[:2]
            s_a = da_model(s_a.reshape(B * T, -1)).view(B, T, -1)
            d_a = da_model(d_a)
        return output, audio_tail(d_a, s_a, self.temp)

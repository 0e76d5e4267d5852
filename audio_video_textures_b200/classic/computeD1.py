"""Drop-in for baselines/classic_video_textures/computeD1.py (RGB branch, :27-96 and the tail :240-247)."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine


def _to_device(frames: torch.Tensor) -> torch.Tensor:
    if frames.is_cuda:
        return frames
    if not torch.cuda.is_available():
        raise RuntimeError("audio_video_textures_b200 needs a CUDA device (B200); there is no CPU path")
    return frames.cuda(non_blocking=True)


def tail(D: torch.Tensor, sigma_factor, stats: torch.Tensor | None = None, threshold=None, want_P=True):
    """sigma / P tail shared by computeD1.py:240-245, computeD2.py:44-50, q_learning.py:53-59.
    Returns (P, P_new, sigma 0-dim CUDA fp32 tensor, counts)."""
    if stats is None:
        stats = engine.sum_nnz(D)
    total, nnz = engine.read_stats(stats)
    sigma = engine.sigma_from_stats(total, nnz, sigma_factor)
    P, P_new, counts = engine.transition_probs(D, sigma, shift=1, threshold=threshold, want_P=want_P,
                                               want_counts=threshold is not None)
    return P, P_new, torch.tensor(sigma, dtype=torch.float32, device=D.device), counts


def compute_D1(
    frames: torch.Tensor,
    sigma_factor: float,
    feats: str = "L2",
    audio: np.ndarray = None,
    sr: int = 0,
    fps: int = 30,
    slow: bool = True,
    batch_size: int = 128,
):
    """Pairwise frame L2 distances, sigma1 and the shifted row-stochastic P1.

    frames: [N, H, W, C] (any trailing layout) float32 or uint8, CPU or CUDA.  `slow` / `batch_size`
    are the reference's tiling knobs (computeD1.py:49,58-63); the per-pair value does not depend on
    them, so they are accepted and ignored.  Only feats == "RGB" is on the hot path (the ResNet /
    VGGish branches :98-238 are feature producers that need pretrained weights).
    Returns (D1 [N,N], P1 [N,N], sigma) as CUDA fp32 tensors, like the reference.
    """
    if feats != "RGB":
        raise NotImplementedError(
            f"feats={feats!r}: only the RGB branch of compute_D1 is implemented (SURVEY.md §2.1 row 1)")
    if not torch.cuda.is_available():
        raise RuntimeError("audio_video_textures_b200 needs a CUDA device (B200); there is no CPU path")
    stats = engine.new_stats(torch.device("cuda", torch.cuda.current_device()))
    streamed = engine.pairwise_l2_from_host(frames, stats=stats) if not frames.is_cuda else None
    if streamed is not None and streamed[1].exact_ok:      # uint8 host frames: Gram overlapped with the H2D copy
        D1 = streamed[0]
    else:
        x = _to_device(frames)
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        stats.zero_()
        D1, _ = engine.pairwise_l2(x, stats=stats)
    P1, _, sigma, _ = tail(D1, sigma_factor, stats)
    return D1, P1, sigma

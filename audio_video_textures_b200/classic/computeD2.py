"""Drop-in for baselines/classic_video_textures/computeD2.py:21-52."""
from __future__ import annotations

import torch

from .. import engine
from .computeD1 import tail


def compute_D2(D1: torch.Tensor, sigma_factor: float, filter_size: int = 16, stride: int = 1):
    """Diagonal binomial temporal filter ('valid', optional stride), sigma2 and P2.

    Returns (D2 [M,M], P2 [M,M], sigma, binomial_filter [1,1,fs,fs]) — the dense fs x fs filter
    tensor is returned only because the caller plots it (video_textures.py:304-311); the kernel
    uses its fs diagonal taps.
    """
    if not D1.is_cuda:
        D1 = D1.cuda()
    if D1.stride(-1) != 1:
        D1 = D1.contiguous()
    taps = engine.binomial_taps(filter_size)
    stats = engine.new_stats(D1.device)
    D2, _ = engine.diag_filter(D1, filter_size, stride, stats=stats, taps=taps)
    P2, _, sigma, _ = tail(D2, sigma_factor, stats)
    binomial_filter = torch.diag(torch.from_numpy(taps)).to(D1.device).view(1, 1, filter_size, filter_size)
    return D2, P2, sigma, binomial_filter

"""Drop-in for baselines/classic_video_textures/q_learning.py:27-68."""
from __future__ import annotations

import torch

from .. import engine
from .computeD1 import tail

LAST = {}     # diagnostics of the most recent call (sweeps, eps trail): the reference only prints them


def q_learning(
    D2: torch.Tensor,
    sigma_factor: float,
    p: float = 0.7,
    alpha: float = 0.997,
    thresholding: float = 0.75,
):
    """Future-cost fixed point on D2**p, sigma3, P3 and the thresholded (not renormalised) P3_new.

    Prints one `Eps:` line per sweep and `Non Zero in P3:` like the reference (:51, :66).
    Returns (D3_new, P3, P3_new, sigma); D2 is not modified.
    """
    if not D2.is_cuda:
        D2 = D2.cuda()
    if D2.stride(-1) != 1:
        D2 = D2.contiguous()
    D3 = engine.pow_matrix(D2, p)                                      # q_learning.py:34
    fc = engine.future_cost_fused(D3, alpha, verbose=True)
    stats = engine.new_stats(D2.device)
    D3_new = engine.future_cost_finalize(D3, fc.mvec, alpha, stats=stats)
    P3, P3_new, sigma, counts = tail(D3_new, sigma_factor, stats, threshold=thresholding)
    LAST.update(n_sweeps=fc.n_sweeps, eps_trail=fc.eps_trail, counts=counts)
    print("Non Zero in P3:", int(counts[0].item()))
    return D3_new, P3, P3_new, sigma

"""Self-consistency check used by bench.py (`parity` field of the multi-GPU lines) and the multi-GPU tests:
a rank's shard of the row-sharded pipeline must equal the same rows of the single-GPU pipeline computed on
that rank's own device — bit for bit for D1 / D2 / D3_new / sweep count / survivor lists (integer Gram,
order-free min and one rounded add), to 1e-6 for sigma (fp64 partial sums are added in a different order).
No oracle involved: both sides are the CUDA product path."""
from __future__ import annotations

import numpy as np
import torch

from . import engine


def single_gpu_pipeline(frames: torch.Tensor, filter_size: int, stride: int, sigma_factor=4.5, threshold=0.08,
                        p: float = 0.7, alpha: float = 0.997) -> dict:
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf)
    D2, D3 = engine.diag_filter(D1, filter_size, stride, p=p)
    fc = engine.future_cost_fused(D3, alpha)
    stats = engine.new_stats(frames.device)
    D3n = engine.future_cost_finalize(D3, fc.mvec, alpha, stats=stats)
    sigma = engine.sigma_from_stats(*engine.read_stats(stats), sigma_factor)
    P3, P3n, counts = engine.transition_probs(D3n, sigma, threshold=threshold, want_counts=True)
    return dict(pf=pf, D1=D1, D2=D2, D3=D3, D3n=D3n, fc=fc, sigma=sigma, P3=P3, P3n=P3n, counts=counts)


def shard_equals_single(res, single: dict) -> dict:
    """`res`: a dist.ShardResult with sigma / P3 / P3_new filled.  Returns {quantity: bool}."""
    p = res.plan
    own = p.a1 - p.a0
    if res.D1.dim() == 3:                     # residue-class planes [stride, class rows from a0, ld]
        s_ = res.D1.shape[0]
        nc = single["D1"].shape[0] // s_
        d1_ok = all(torch.equal(res.D1[r, :, :nc], single["D1"][r::s_, r::s_][p.a0:p.a0 + res.D1.shape[1]])
                    for r in range(s_))
    else:
        d1_ok = torch.equal(res.D1, single["D1"][p.r_lo:p.r_hi])
    return dict(
        D1=d1_ok,
        D2=torch.equal(res.D2, single["D2"][p.a0:p.a1h]),
        D3n=torch.equal(res.D3_new, single["D3n"][p.a0:p.a1h]),
        sweeps=res.n_sweeps == single["fc"].n_sweeps,
        eps=res.eps_trail == single["fc"].eps_trail or
        bool(np.allclose(res.eps_trail, single["fc"].eps_trail, rtol=1e-6)),
        sigma=abs(float(res.sigma) - float(single["sigma"])) <= 1e-6 * float(single["sigma"]),
        P3=torch.allclose(res.P3, single["P3"][p.a0:p.a1], rtol=1e-5, atol=0),
        survivors=torch.equal(res.counts, single["counts"][p.a0:p.a1]) and
        torch.equal(res.P3_new != 0, single["P3n"][p.a0:p.a1] != 0))

"""GPU tests at BASELINE.json's full C2 size (5000 frames, 224x224 RGB, -m 3 -fs 40 -stride 4), where the
CPU oracle cannot run in test time: size-independent properties + sampled exactness."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    from audio_video_textures_b200 import engine
    from audio_video_textures_b200.synth import synth_video_cuda
    frames = synth_video_cuda(5000, 224, 224, seed=0)
    stats = engine.new_stats(frames.device)
    pf = engine.pack_frames(frames)
    D1 = engine.gram_l2(pf, stats=stats)
    assert pf.exact_ok
    return engine, frames, pf, D1, stats


def test_c2_distance_properties_and_sampled_exactness(c2):
    eng, frames, pf, D1, stats = c2
    n = frames.shape[0]
    assert torch.equal(D1, D1.T)                                   # symmetric, bit for bit
    assert float(D1.diagonal().abs().max()) == 0.0
    total, nnz = eng.read_stats(stats)
    assert nnz == n * n - n                                        # no duplicate frames in the synthetic clip
    np.testing.assert_allclose(total, D1.double().sum().item(), rtol=2e-7)
    # exact integer d^2 on the CPU for random pairs (int64 arithmetic on the bytes)
    gen = torch.Generator().manual_seed(0)
    idx = torch.randint(0, n, (64, 2), generator=gen)
    x = frames.reshape(n, -1)
    for i, j in idx.tolist():
        d2 = int(((x[i].cpu().to(torch.int64) - x[j].cpu().to(torch.int64)) ** 2).sum())
        want = np.float32(np.sqrt(np.float32(d2)))                 # same rounding sequence as the epilogue
        assert float(D1[i, j]) == float(want), (i, j, float(D1[i, j]), float(want))
    # a row block computed separately (what a shard does) is identical
    blk = eng.gram_l2(pf, 1234, 700)
    assert torch.equal(blk, D1[1234:1934])
    # triangle inequality on random triples (a property the wrap-around arithmetic would break)
    t = torch.randint(0, n, (2000, 3), generator=gen).cuda()
    a, b, c = D1[t[:, 0], t[:, 1]], D1[t[:, 1], t[:, 2]], D1[t[:, 0], t[:, 2]]
    assert bool((c <= (a + b) * (1 + 1e-6)).all())


def test_c2_filter_future_cost_probabilities_properties(c2):
    eng, frames, pf, D1, stats = c2
    fs, stride = 40, 4
    D2, D3 = eng.diag_filter(D1, fs, stride, p=0.7)
    m = D2.shape[0]
    assert m == 1241 and torch.equal(D2, D2.T)                     # a symmetric D1 filtered along diagonals stays symmetric
    # spot-check the filter against a float64 evaluation
    w = torch.from_numpy(eng.binomial_taps(fs)).double().cuda()
    gen = torch.Generator().manual_seed(1)
    for a, b in torch.randint(0, m, (32, 2), generator=gen).tolist():
        k = torch.arange(fs, device="cuda")
        want = float((w * D1[a * stride + k, b * stride + k].double()).sum())
        assert abs(float(D2[a, b]) - want) <= 1e-5 * want + 1e-6
    fused = eng.future_cost_fused(D3)
    loop = eng.future_cost(D3)
    assert fused.n_sweeps == loop.n_sweeps and torch.equal(fused.mvec, loop.mvec[:m])
    assert fused.eps_trail[-1] <= 0.01 < fused.eps_trail[0]
    # fixed-point residual of the converged vector: m_j = min_{k != j}(D3[j,k] + alpha*m_k) up to the stop rule
    mv = fused.mvec
    X = D3 + np.float32(0.997) * mv[None, :]
    X.fill_diagonal_(float("inf"))
    resid = (X.min(1)[0] - mv)[1:]
    assert float((resid ** 2).mean()) <= 0.011
    stats3 = eng.new_stats("cuda")
    D3n = eng.future_cost_finalize(D3, mv, stats=stats3)
    assert torch.equal(D3n[0], D3[0]) and torch.equal(D3n[1:], D3[1:] + np.float32(0.997) * mv[None, :])
    sigma = eng.sigma_from_stats(*eng.read_stats(stats3), 4.5)
    P3, P3n, counts = eng.transition_probs(D3n, sigma, threshold=0.08, want_counts=True)
    np.testing.assert_allclose(P3.sum(1).cpu().numpy(), 1.0, rtol=2e-6)
    assert bool((P3n.max(1)[0] == P3.max(1)[0]).all())             # the row maximum always survives
    assert bool(((P3n == 0) | (P3n == P3)).all())                  # survivors keep their value (not renormalised)
    assert torch.equal(counts.long(), (P3n != 0).sum(1))
    rowptr, cols = eng.csr_from_matrix(P3n, counts)
    assert rowptr[-1] == int(counts.sum()) and np.all(np.diff(cols)[np.diff(np.repeat(np.arange(m), np.diff(rowptr))) == 0] > 0)
    # threshold monotone: a larger threshold keeps a superset
    _, P3n_wide, _ = eng.transition_probs(D3n, sigma, threshold=0.2)
    assert bool(((P3n != 0) <= (P3n_wide != 0)).all())

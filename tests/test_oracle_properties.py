"""CPU property tests (hypothesis) of the oracle restatement - the invariants the GPU tests rely on."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import classic, contrastive


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 24), k=st.integers(1, 40), seed=st.integers(0, 10_000))
def test_distance_symmetric_zero_diagonal_and_exact_for_bytes(n, k, seed):
    x = torch.randint(0, 256, (n, k), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
    D = classic.pairwise_l2(x.float())
    assert torch.equal(D, D.T) and float(D.diagonal().abs().max()) == 0.0
    np.testing.assert_allclose(D.numpy(), classic.pairwise_l2_exact_u8(x).numpy(), rtol=1e-6)


@settings(max_examples=20, deadline=None)
@given(n=st.integers(12, 40), fs=st.integers(1, 8), stride=st.integers(1, 3), shift=st.integers(0, 5))
def test_filter_of_shifted_identity_is_binomial_band(n, fs, stride, shift):
    """D1 = 1 on the diagonal offset by `shift*stride`: the filter must return sum(w) = 1 on the band and 0 off it."""
    if n < fs + stride * (shift + 1):
        return
    D1 = torch.zeros(n, n)
    idx = torch.arange(n - shift * stride)
    D1[idx, idx + shift * stride] = 1.0
    D2 = classic.compute_D2(D1, torch.tensor(4.5), fs, stride)[0]
    m = D2.shape[0]
    want = torch.zeros(m, m)
    i = torch.arange(m - shift) if shift < m else torch.arange(0)
    want[i, i + shift] = 1.0
    np.testing.assert_allclose(D2.numpy(), want.numpy(), atol=1e-6)
    np.testing.assert_allclose(classic.diag_filter_sequential(D1, fs, stride).numpy(), D2.numpy(), atol=1e-6)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(m=st.integers(3, 30), seed=st.integers(0, 10_000))
def test_future_cost_reaches_its_fixed_point(m, seed):
    D3 = torch.rand(m, m, generator=torch.Generator().manual_seed(seed)) * 10 + 0.1
    out, trail = classic.future_cost(D3)
    assert trail[-1] <= 0.01 and all(e > 0.01 for e in trail[:-1])
    assert torch.equal(out[0], D3[0])                                  # row 0 is never updated (q_learning.py:42)
    mins = classic.row_min_offdiag(out)
    nxt = D3[1:] + 0.997 * mins
    # one more sweep changes (almost) nothing: the stop rule bounds the mean over all m rows (row 0 never moves),
    # this mean runs over the m-1 updated rows, and a min over k is non-expansive only in the sup norm
    assert float(((nxt - out[1:]) ** 2).mean()) <= 0.01 * m / (m - 1) * 1.5
    lit, trail2 = classic.future_cost(D3, faithful=True)
    assert torch.equal(lit, out) and trail2 == trail


@settings(max_examples=20, deadline=None)
@given(m=st.integers(2, 30), seed=st.integers(0, 10_000), t1=st.floats(0.0, 0.5), t2=st.floats(0.0, 0.5))
def test_threshold_survivors_monotone_in_threshold(m, seed, t1, t2):
    P = torch.rand(m, m, generator=torch.Generator().manual_seed(seed))
    P = P / P.sum(1, keepdim=True)
    lo, hi = sorted((t1, t2))
    a, b = classic.threshold_rows(P, lo) != 0, classic.threshold_rows(P, hi) != 0
    assert bool((a <= b).all()) and bool(a.any(1).all())               # superset; the row maximum always survives


@settings(max_examples=20, deadline=None)
@given(L=st.integers(2, 40), q=st.integers(0, 39))
def test_target_order_is_a_permutation_with_the_positive_first(L, q):
    q = q % L
    ids = contrastive.target_order(q, L)
    pos = min(q + 1, L - 1)
    assert ids[0] == pos and list(ids[1:]) == sorted(set(range(L)) - {q, pos})
    assert len(ids) == (L if q == L - 1 else L - 1)

"""GPU parity tests (B200): every kernel of the classic path through the C ABI against the golden
vectors of the unmodified reference and against the CPU oracle on the same inputs.

Tolerances: integer / index work bit-exact; order-free fp32 work (future-cost sweeps, thresholds on
given probabilities) bit-exact; everything that involves a reduction order or exp/pow is held to
rtol 1e-4 (BASELINE.json north_star), most of it to 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import CLASSIC_GOLDEN, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from audio_video_textures_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return engine


def _csr_of(P):
    rows, cols = torch.nonzero(P, as_tuple=True)
    counts = torch.bincount(rows, minlength=P.shape[0])
    return torch.cat((torch.zeros(1, dtype=torch.long), counts.cumsum(0))).numpy(), cols.numpy()


# ----------------------------------------------------------------------------- K0 / K1
@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_direct_distance_matches_reference(eng, name):
    g = load_golden(name)
    video = torch.from_numpy(g["video"]).cuda()
    stats = eng.new_stats(video.device)
    D = eng.pairdist_direct(video.float(), stats=stats)
    np.testing.assert_allclose(D.cpu().numpy(), g["ref_D1"], rtol=1e-5)
    Du8 = eng.pairdist_direct(video)
    assert torch.equal(D, Du8)
    from oracle import classic
    np.testing.assert_allclose(D.cpu().numpy(), classic.pairwise_l2_exact_u8(video.cpu()).numpy(), rtol=2e-7)
    total, nnz = eng.read_stats(stats)
    assert nnz == int((g["ref_D1"] != 0).sum())
    np.testing.assert_allclose(total, g["ref_D1"].astype(np.float64).sum(), rtol=1e-5)


@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_pack_frames_exact(eng, name):
    g = load_golden(name)
    video = torch.from_numpy(g["video"])
    x = video.reshape(video.shape[0], -1)
    raw = eng.pack_frames(video.cuda())
    assert raw.exact_ok and not raw.signed and raw.packed.dtype == torch.uint8          # bytes used as they are
    assert torch.equal(raw.sqnorm.cpu(), (x.to(torch.int64) ** 2).sum(1))
    mis = torch.empty(x.shape[0], x.shape[1] + 1, dtype=torch.uint8, device="cuda")[:, :x.shape[1]]
    mis.copy_(x)                                                                       # pitch not 16-byte aligned
    pf = eng.pack_frames(mis)
    assert pf.exact_ok and pf.signed
    ref = (x.to(torch.int16) - 128).to(torch.int8)
    assert torch.equal(pf.packed[:, :x.shape[1]].cpu(), ref)
    assert int(pf.packed[:, x.shape[1]:].abs().sum()) == 0
    assert torch.equal(pf.sqnorm.cpu(), (ref.to(torch.int64) ** 2).sum(1))
    pf32 = eng.pack_frames(video.float().cuda())
    assert pf32.exact_ok and torch.equal(pf32.packed, pf.packed) and torch.equal(pf32.sqnorm, pf.sqnorm)
    bad = video.float().cuda()
    bad[3, 0, 0, 0] = 17.5
    assert not eng.pack_frames(bad).exact_ok


@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_gram_distance_exact_and_within_tolerance(eng, name):
    from oracle import classic
    g = load_golden(name)
    video = torch.from_numpy(g["video"])
    pf = eng.pack_frames(video.cuda())
    stats = eng.new_stats("cuda")
    D = eng.gram_l2(pf, stats=stats)
    torch.cuda.synchronize()
    Dh = D.cpu()
    exact = classic.pairwise_l2_exact_u8(video)
    # d^2 is an exact integer on the GPU; only the final fp32 sqrt rounds
    np.testing.assert_allclose(Dh.numpy(), exact.numpy(), rtol=2e-7, atol=0)
    assert torch.equal(Dh, Dh.T) and float(Dh.diagonal().abs().max()) == 0.0
    assert np.array_equal(Dh.numpy() == 0, g["ref_D1"] == 0)
    np.testing.assert_allclose(Dh.numpy(), g["ref_D1"], rtol=RTOL)
    total, nnz = eng.read_stats(stats)
    assert nnz == int((Dh != 0).sum())
    np.testing.assert_allclose(total, Dh.double().sum().item(), rtol=2e-7)   # fp32 partials per 32-chunk
    # centred-int8 operand path (float frames) gives the identical matrix
    pf8 = eng.pack_frames(video.float().cuda())
    assert pf8.signed and not pf.signed
    assert torch.equal(eng.gram_l2(pf8).cpu(), Dh)
    # row-block (non-symmetric) mode reproduces the same values: what a row shard computes
    n = video.shape[0]
    r0, rows = n // 3, n // 2
    Db = eng.gram_l2(pf, r0, rows)
    assert torch.equal(Db.cpu(), Dh[r0:r0 + rows])
    # against the direct-difference kernel on the same device
    Dd = eng.pairdist_direct(video.cuda())
    assert torch.equal(Dh, Dd.cpu())                         # byte frames: both paths give the exact d^2


def test_streamed_host_frames_equal_resident(eng):
    """compute_D1 on uint8 HOST frames (chunked H2D overlapped with the Gram jobs) == device-resident path."""
    from audio_video_textures_b200.classic.computeD1 import compute_D1
    from audio_video_textures_b200.synth import synth_video
    video = synth_video(1300, 16, 16, seed=4)
    f = torch.tensor(4.5)
    for host in (video, video.pin_memory()):
        D1h, P1h, s1h = compute_D1(host, f, "RGB")
        D1d, P1d, s1d = compute_D1(video.cuda(), f, "RGB")
        assert torch.equal(D1h, D1d)
        np.testing.assert_allclose(s1h.item(), s1d.item(), rtol=1e-6)
        np.testing.assert_allclose(P1h.cpu().numpy(), P1d.cpu().numpy(), rtol=1e-5)
    res = eng.pairwise_l2_from_host(video, chunks=3)
    assert res is not None and torch.equal(res[0], D1d)
    assert eng.pairwise_l2_from_host(video[:100]) is None


def test_streamed_float_host_frames_equal_resident(eng):
    """The reference's own call hands FLOAT frames to compute_D1 (classic/video_textures.py:245,266): float32 host
    frames go through the same overlapped chunked path (staged, packed to centred int8 per chunk on the device) and
    give the identical matrix; non-byte float features are detected after the last chunk and take the direct
    fp32 kernel instead."""
    from audio_video_textures_b200.classic.computeD1 import compute_D1
    from audio_video_textures_b200.synth import synth_video
    video = synth_video(1300, 16, 16, seed=4)
    f = torch.tensor(4.5)
    D1d, P1d, s1d = compute_D1(video.cuda(), f, "RGB")
    for host in (video.float(), video.float().pin_memory()):
        for chunks in (8, 3, 5):
            res = eng.pairwise_l2_from_host(host, chunks=chunks)
            assert res is not None and res[1].exact_ok and torch.equal(res[0], D1d)
        D1h, P1h, s1h = compute_D1(host, f, "RGB")
        assert torch.equal(D1h, D1d)
        np.testing.assert_allclose(s1h.item(), s1d.item(), rtol=1e-6)
    feats = torch.randn(600, 96)                                    # real-valued features: not bytes
    res = eng.pairwise_l2_from_host(feats)
    assert res is not None and not res[1].exact_ok
    D, _, _ = compute_D1(feats, f, "RGB")
    np.testing.assert_allclose(D.cpu().numpy(), torch.cdist(feats.double(), feats.double()).float().numpy(), rtol=2e-5, atol=1e-5)


def test_gram_edge_shapes(eng):
    """Ragged N (not a multiple of the 128 x 256 tile), K not a multiple of 128, duplicate frames,
    a single K block, and N smaller than one tile."""
    from oracle import classic
    gen = torch.Generator().manual_seed(5)
    for n, k in [(1, 7), (2, 128), (130, 100), (257, 384), (300, 1000), (515, 129)]:
        x = torch.randint(0, 256, (n, k), dtype=torch.uint8, generator=gen)
        if n > 5:
            x[5] = x[2]                        # duplicate -> exact zero off the diagonal
        exact = classic.pairwise_l2_exact_u8(x)
        for frames in (x.cuda(), x.float().cuda()):          # raw-u8 operand (when aligned) and centred-s8 operand
            pf = eng.pack_frames(frames)
            D = eng.gram_l2(pf).cpu()
            np.testing.assert_allclose(D.numpy(), exact.numpy(), rtol=2e-7)
            if n > 5:
                assert D[5, 2] == 0 and D[2, 5] == 0


def test_gram_domain_guard_and_extreme_values(eng):
    """The epilogue evaluates d^2 modulo 2^32.  The host wrapper only takes the tensor-core path when
    4*max(n) < 2^32, which bounds every prefix of <x,y> by 2^30 (Cauchy-Schwarz) and d^2 by 2^32, so
    nothing can wrap; black-vs-white frames at K = 150528 (d^2 = K*255^2 > 2^32) must be routed to the
    direct kernel, and large-but-legal norms at the same K must stay exact."""
    from oracle import classic
    k = 150528
    x = torch.zeros((4, k), dtype=torch.uint8)
    x[1] = 255
    x[2] = 200
    x[3, ::2] = 255
    pf = eng.pack_frames(x.cuda())
    assert not pf.exact_ok
    with pytest.raises(Exception):
        eng.pairwise_l2(x.cuda(), method="gram")
    D, used = eng.pairwise_l2(x.cuda())
    assert used == "direct"
    np.testing.assert_allclose(D.cpu().numpy(), classic.pairwise_l2_exact_u8(x).numpy(), rtol=2e-7)
    y = torch.full((3, k), 128, dtype=torch.uint8)
    y[0] = 128 + 84                                           # n = K*84^2 = 1.06e9: 4n just below 2^32
    y[1] = 128 + 80
    y[2] = 128 - 80
    for frames in (y.cuda(), y.float().cuda()):
        # raw-u8 operand: <y0,y1> = K*212*208 = 6.6e9 wraps the int32 accumulator (mod 2^32) and must
        # still give the exact distance; centred-s8 operand: no wrap
        pf = eng.pack_frames(frames)
        assert pf.exact_ok
        Dg = eng.gram_l2(pf).cpu()
        np.testing.assert_allclose(Dg.numpy(), classic.pairwise_l2_exact_u8(y).numpy(), rtol=2e-7)


# ----------------------------------------------------------------------------- K2
@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_diag_filter_matches_reference(eng, name):
    from oracle import classic
    g = load_golden(name)
    fs, stride = int(g["fs"]), int(g["stride"])
    D1 = torch.from_numpy(g["ref_D1"]).cuda()
    stats = eng.new_stats("cuda")
    D2, D3 = eng.diag_filter(D1, fs, stride, p=0.7, stats=stats)
    np.testing.assert_allclose(D2.cpu().numpy(), g["ref_D2"], rtol=1e-5)
    seq = classic.diag_filter_sequential(torch.from_numpy(g["ref_D1"]), fs, stride)
    np.testing.assert_allclose(D2.cpu().numpy(), seq.numpy(), rtol=2e-6)
    np.testing.assert_allclose(D3.cpu().numpy(), (D2.cpu() ** 0.7).numpy(), rtol=1e-5)
    total, nnz = eng.read_stats(stats)
    assert nnz == int((D2 != 0).sum())
    np.testing.assert_allclose(total, D2.double().sum().item(), rtol=1e-8)    # fp32 partials inside a band, fp64 across
    np.testing.assert_array_equal(eng.binomial_taps(fs), g["ref_filter_diag"])
    # row-sharded call with a halo'd D1 block gives the same rows
    m = D2.shape[0]
    a0, rows = m // 3, m // 4
    lo, hi = a0 * stride, (a0 + rows - 1) * stride + fs
    part, _ = eng.diag_filter(D1[lo:hi], fs, stride, m=m, a0=a0, rows_out=rows, in_row0=lo)
    assert torch.equal(part, D2[a0:a0 + rows])


@pytest.mark.parametrize("fs,stride,n", [(5, 2, 61), (3, 3, 40), (70, 1, 200), (1, 1, 33), (16, 4, 100),
                                         (8, 1, 9), (40, 4, 43)])
def test_diag_filter_generic_and_edge(eng, fs, stride, n):
    from oracle import classic
    gen = torch.Generator().manual_seed(fs * 100 + stride)
    D1 = torch.rand(n, n, generator=gen) * 100
    D2, _ = eng.diag_filter(D1.cuda(), fs, stride)
    ref = classic.compute_D2(D1, torch.tensor(4.5), fs, stride)[0]
    assert D2.shape == ref.shape
    np.testing.assert_allclose(D2.cpu().numpy(), ref.numpy(), rtol=1e-5)


@pytest.mark.parametrize("fs,stride,n", [(40, 1, 300), (40, 1, 777), (40, 4, 1241), (40, 4, 523), (16, 1, 130),
                                         (16, 4, 401), (8, 1, 64), (40, 1, 41), (40, 4, 40), (5, 2, 61)])
def test_diag_filter_symmetric_form_is_bit_identical(eng, fs, stride, n):
    """K2's symmetric form (upper triangle computed, mirrored through shared memory) against the general kernel on
    a symmetric D1: D2 and D3 bit for bit, nnz equal, sum to fp32-partial accuracy.  Also: the automatic choice
    (`known_symmetric`) only fires for matrices the engine produced and drops out after an in-place edit."""
    gen = torch.Generator().manual_seed(fs * 1000 + stride * 10 + n)
    A = torch.rand(n, n, generator=gen) * 3000
    D1 = eng.empty_matrix(n, n, "cuda")
    D1.copy_(((A + A.T) / 2).fill_diagonal_(0.0))
    assert torch.equal(D1, D1.T)
    st_g, st_s = eng.new_stats("cuda"), eng.new_stats("cuda")
    G2, G3 = eng.diag_filter(D1, fs, stride, p=0.7, stats=st_g, symmetric=False)
    S2, S3 = eng.diag_filter(D1, fs, stride, p=0.7, stats=st_s, symmetric=True)
    assert torch.equal(G2, S2) and torch.equal(G3, S3)
    assert torch.equal(S2, S2.T)
    (tg, zg), (ts, zs) = eng.read_stats(st_g), eng.read_stats(st_s)
    assert zg == zs
    np.testing.assert_allclose(ts, tg, rtol=1e-6)
    S2only, none = eng.diag_filter(D1, fs, stride, symmetric=True)
    assert none is None and torch.equal(S2only, G2)
    # automatic choice
    assert not eng.known_symmetric(D1)                      # filled by the caller: unknown provenance
    frames = torch.randint(0, 256, (max(n, 64), 8, 8, 3), dtype=torch.uint8, generator=gen).cuda()
    Dg = eng.gram_l2(eng.pack_frames(frames))
    assert eng.known_symmetric(Dg) and torch.equal(Dg, Dg.T)
    assert not eng.known_symmetric(Dg[:-1])                 # another view: not recognised
    Dg[0, 1] += 1.0
    assert not eng.known_symmetric(Dg)                      # modified in place: no longer trusted


@pytest.mark.parametrize("fs,stride,n,hw", [(40, 4, 1000, 16), (16, 4, 516, 16), (40, 4, 44, 16), (40, 4, 2048, 16)])
def test_residue_class_pipeline_is_bit_identical(eng, fs, stride, n, hw):
    """Stride-s pipelines read D1[i,j] only where i = j (mod s): the s residue-class Gram matrices (one launch, 1/s of
    the pairs) + the plane-walking filter must give the SAME bits as the full distance matrix + filter."""
    gen = torch.Generator().manual_seed(n + fs)
    base = torch.randint(0, 256, (1, hw, hw, 3), dtype=torch.uint8, generator=gen)
    drift = torch.randint(-20, 21, (n, hw, hw, 3), generator=gen)
    frames = (base.int() + torch.cumsum(drift, 0) // 8).clamp(0, 255).to(torch.uint8).cuda()   # K = hw*hw*3: 768 / 192... 
    pf = eng.pack_frames(frames)
    if not eng.residue_eligible(pf, fs, stride):
        pytest.skip("clip not eligible (K % 128)")
    D1 = eng.gram_l2(pf)
    st_f, st_r = eng.new_stats("cuda"), eng.new_stats("cuda")
    F2, F3 = eng.diag_filter(D1, fs, stride, p=0.7, stats=st_f, symmetric=False)
    D1r = eng.gram_l2_residues(pf, stride)
    for r in range(stride):
        nc = n // stride
        assert torch.equal(D1r[r, :, :nc], D1[r::stride, r::stride])
    for sym in (False, True):
        st_r.zero_()
        R2, R3 = eng.diag_filter_residues(D1r, n, fs, stride, p=0.7, stats=st_r, symmetric=sym)
        assert torch.equal(R2, F2) and torch.equal(R3, F3), sym
        (tf, zf), (tr, zr) = eng.read_stats(st_f), eng.read_stats(st_r)
        assert zf == zr
        np.testing.assert_allclose(tr, tf, rtol=1e-6)
    m = F2.shape[0]
    if m >= 12:                                           # row shard of the planes (class rows lo .. hi)
        a0, rows = m // 3, m // 4
        lo, hi = a0, a0 + rows - 1 + (fs - 1) // stride + 1
        part, part3 = eng.diag_filter_residues(D1r[:, lo:hi], n, fs, stride, p=0.7, a0=a0, rows_out=rows, in_row0=lo)
        assert torch.equal(part, F2[a0:a0 + rows]) and torch.equal(part3, F3[a0:a0 + rows])
    D2, D3, how = eng.distance_filter(frames, fs, stride)
    assert how == "residues" and torch.equal(D2, F2) and torch.equal(D3, F3)
    D2b, _, how_b = eng.distance_filter(frames[:n - 1], fs, stride)                  # N % s != 0: full path
    assert how_b == "gram" and D2b.shape[0] == eng.filtered_size(n - 1, fs, stride)


@pytest.mark.parametrize("stride,residues", [(4, True), (4, False), (1, False)])
def test_pipeline_graph_replays_equal_eager(eng, stride, residues):
    """engine.PipelineGraph: the whole pass captured in one CUDA graph; replays on different clips must equal the
    eager pipeline bit for bit (D2, D3, D3_new, sweep count, eps trail, sigma statistics)."""
    from audio_video_textures_b200 import selfcheck
    from audio_video_textures_b200.synth import synth_video
    n, h, w, fs = 600, 16, 16, 16
    g = eng.PipelineGraph(n, h * w * 3, fs, stride, residues=residues)
    assert g.how == ("residues" if residues else "gram")
    for seed in (1, 2, 1):
        frames = synth_video(n, h, w, seed=seed).cuda()
        single = selfcheck.single_gpu_pipeline(frames, fs, stride, 4.5, 0.08)
        g(frames)
        assert torch.equal(g.D2, single["D2"]) and torch.equal(g.D3, single["D3"]) and torch.equal(g.D3_new, single["D3n"])
        assert g.n_sweeps == single["fc"].n_sweeps and g.eps_trail == single["fc"].eps_trail
        sigma = eng.sigma_from_stats(*eng.read_stats(g.stats), 4.5)
        np.testing.assert_allclose(float(sigma), float(single["sigma"]), rtol=1e-6)


@pytest.mark.parametrize("n,h,w", [(777, 12, 12), (1000, 16, 16), (2048, 8, 8), (300, 20, 28)])
def test_norms_fused_into_the_gram_launch(eng, n, h, w):
    """pack_frames(defer_norms=True): K0 runs inside the Gram launch (idle epilogue warps).  Norms, the centred
    maximum, D1 (symmetric, row block, job list) and the residue-class planes must equal the separate-kernel path."""
    from audio_video_textures_b200.synth import synth_video
    frames = synth_video(n, h, w, seed=n).cuda()
    ref = eng.pack_frames(frames)
    D1 = eng.gram_l2(ref)
    pf = eng.pack_frames(frames, defer_norms=True)
    assert pf.norms_pending and not pf.signed
    D1f = eng.gram_l2(pf)
    assert not pf.norms_pending
    assert torch.equal(pf.sqnorm, ref.sqnorm) and torch.equal(pf.flags, ref.flags)
    assert torch.equal(D1f, D1) and eng.known_symmetric(D1f)
    pf2 = eng.pack_frames(frames, defer_norms=True)
    blk = eng.gram_l2(pf2, 100, 150)
    assert torch.equal(blk, D1[100:250]) and torch.equal(pf2.sqnorm, ref.sqnorm)
    if eng.residue_eligible(ref, 40, 4):
        pf3 = eng.pack_frames(frames, defer_norms=True)
        assert torch.equal(eng.gram_l2_residues(pf3, 4), eng.gram_l2_residues(ref, 4))
        assert torch.equal(pf3.sqnorm, ref.sqnorm) and torch.equal(pf3.flags, ref.flags)
    # black-vs-white frames: the centred maximum must flag the clip as outside the exact domain in both paths
    bw = torch.zeros((64, 224, 224, 3), dtype=torch.uint8, device="cuda")
    bw[::2] = 255
    a, b = eng.pack_frames(bw), eng.pack_frames(bw, defer_norms=True)
    eng.gram_l2(b)
    assert torch.equal(a.flags, b.flags) and a.exact_ok == b.exact_ok


def test_fused_pow_accuracy(eng):
    """D3 = D2 ** p is evaluated by a split-exponent exp2/log2 (common.cuh: pow_pos) instead of powf;
    it must stay within a few ulp of the exact power over the whole dynamic range, incl. 0."""
    gen = torch.Generator().manual_seed(0)
    n = 512
    x = torch.exp(torch.rand(n, n, generator=gen, dtype=torch.float64) * 60 - 30).float()   # 1e-13 .. 1e13
    x[0, :8] = torch.tensor([0.0, 1.0, 2.0, 0.5, 1e-38, 3e38, 1e-45, 7.0])
    for p in (0.7, 0.5, 1.0, 2.0, 0.123):
        _, D3 = eng.diag_filter(x.cuda(), 1, 1, p=p, taps=[1.0])
        want = torch.from_numpy(np.power(x.double().numpy(), float(np.float32(p))))
        got = D3.cpu().double()
        ok = want > 0
        rel = ((got - want).abs() / want)[ok & torch.isfinite(want) & (want < 3e38) & (want > 1e-37)]
        assert float(rel.max()) < 6e-7, (p, float(rel.max()))
        assert float(D3[0, 0]) == 0.0


# ----------------------------------------------------------------------------- K3 / K4
@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_future_cost_bit_exact(eng, name):
    from oracle import classic
    g = load_golden(name)
    D3 = torch.from_numpy(g["ref_D2"]) ** 0.7                  # same D3 on both sides
    want, trail = classic.future_cost(D3)
    fc = eng.future_cost(D3.cuda())
    assert fc.n_sweeps == int(g["ref_n_sweeps"]) == len(trail)
    stats = eng.new_stats("cuda")
    got = eng.future_cost_finalize(D3.cuda(), fc.mvec, stats=stats)
    assert torch.equal(got.cpu(), want)                        # min / single rounded add: order-free
    np.testing.assert_array_equal(got.cpu().numpy(), g["ref_D3_new"])
    np.testing.assert_allclose(fc.eps_trail, trail, rtol=1e-5, atol=1e-12)
    for e in fc.eps_trail:                                     # stop decision is not marginal
        assert abs(e - 0.01) / 0.01 > 1e-3
    # all sweeps in one cooperative launch: same vector, same sweep count, same eps trail
    ff = eng.future_cost_fused(D3.cuda())
    assert ff.n_sweeps == fc.n_sweeps and torch.equal(ff.mvec, fc.mvec[:ff.mvec.shape[0]])
    np.testing.assert_allclose(ff.eps_trail, trail, rtol=1e-5, atol=1e-12)
    total, nnz = eng.read_stats(stats)
    sigma = eng.sigma_from_stats(total, nnz, g["sigma_factor"])
    np.testing.assert_allclose(sigma, g["ref_sigma3"], rtol=1e-6)


def test_future_cost_wide_rows(eng):
    """M = 4500: exercises the 4-deep unrolled 128-bit loop, its remainder loop and the scalar tail."""
    from oracle import classic
    gen = torch.Generator().manual_seed(9)
    M = 4503
    D3 = (torch.rand(M, M, generator=gen) * 40 + 0.5)
    want, trail = classic.future_cost(D3)
    d = D3.cuda()
    fc = eng.future_cost(d)
    assert fc.n_sweeps == len(trail)
    assert torch.equal(eng.future_cost_finalize(d, fc.mvec).cpu(), want)
    np.testing.assert_allclose(fc.eps_trail, trail, rtol=1e-5, atol=1e-12)
    ff = eng.future_cost_fused(d)
    assert ff.n_sweeps == len(trail) and torch.equal(eng.future_cost_finalize(d, ff.mvec).cpu(), want)
    # probabilities at the same width: shared-memory row cache path vs torch
    sigma = np.float32(25.0)
    P, Pn, counts = eng.transition_probs(d, sigma, threshold=0.08, want_counts=True)
    E = torch.exp(-D3 / torch.tensor(sigma))
    E = torch.cat((E[1:], E[-1:]), 0)
    Pw = E / E.sum(1, keepdim=True)
    np.testing.assert_allclose(P.cpu().numpy(), Pw.numpy(), rtol=1e-5)
    assert torch.equal(counts.cpu().long(), (Pn != 0).sum(1).cpu())


def test_future_cost_unaligned_rows_and_row_blocks(eng):
    """Odd M (rows not 16-byte aligned -> scalar path) and the row-sharded call pattern."""
    from oracle import classic
    gen = torch.Generator().manual_seed(3)
    M = 203
    D3 = (torch.rand(M, M, generator=gen) * 50 + 1)
    want, trail = classic.future_cost(D3)
    d = D3.cuda()
    fc = eng.future_cost(d)
    assert fc.n_sweeps == len(trail)
    assert torch.equal(eng.future_cost_finalize(d, fc.mvec).cpu(), want)
    # one sweep computed in two row blocks == one block
    m0 = torch.zeros(M, device="cuda")
    import ctypes as C
    from audio_video_textures_b200 import _lib
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for r0, rows in ((0, 100), (100, 103)):
        blk = d[r0:r0 + rows]
        _lib.call("avtex_future_cost_sweep", _lib.ptr(blk), blk.stride(0), r0, rows, M, None, None,
                  C.c_float(0.997), _lib.ptr(m0), None, 0, st)
    assert torch.equal(m0.cpu(), classic.row_min_offdiag(D3))


# ----------------------------------------------------------------------------- K5 + CSR + walk
@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_transition_probs_threshold_and_walk(eng, name):
    from audio_video_textures_b200.classic.video_textures import texture_walk
    g = load_golden(name)
    th = float(g["threshold"])
    D3n = torch.from_numpy(g["ref_D3_new"]).cuda()
    P3, P3n, counts = eng.transition_probs(D3n, g["ref_sigma3"], threshold=th, want_counts=True)
    np.testing.assert_allclose(P3.cpu().numpy(), g["ref_P3"], rtol=1e-5)
    np.testing.assert_allclose(P3.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    rowptr, cols = eng.csr_from_matrix(P3n, counts)
    np.testing.assert_array_equal(rowptr, g["ref_P3new_rowptr"])     # survivor sets: bit-exact
    np.testing.assert_array_equal(cols, g["ref_P3new_cols"])
    rp2, c2 = eng.csr_from_matrix(P3n)                                # count kernel path
    np.testing.assert_array_equal(rp2, rowptr)
    kept = P3n != 0
    assert torch.equal(P3n[kept], P3[kept])                           # not renormalised (q_learning.py:64)
    np.random.seed(int(g["seed"]))
    frames, jumps = texture_walk(P3n, int(g["model_type"]), int(g["fps"]), int(g["nvl"]), int(g["stride"]),
                                 int(g["fs"]))
    np.testing.assert_array_equal(np.array(frames), g["walk_frames"])
    assert jumps == int(g["walk_jump_count"])


# ----------------------------------------------------------------------------- drop-in, end to end
@pytest.mark.parametrize("name", CLASSIC_GOLDEN)
def test_dropin_pipeline_end_to_end(eng, name, capsys):
    """video -> compute_D1 -> compute_D2 -> q_learning -> walk through the reference-named entry
    points; values rtol 1e-4, survivor sets and the emitted frame sequence bit-exact."""
    from audio_video_textures_b200.classic import q_learning as ql_mod
    from audio_video_textures_b200.classic.computeD1 import compute_D1
    from audio_video_textures_b200.classic.computeD2 import compute_D2
    from audio_video_textures_b200.classic.q_learning import q_learning
    from audio_video_textures_b200.classic.video_textures import texture_walk
    g = load_golden(name)
    f = torch.tensor(float(g["sigma_factor"]), dtype=torch.float32)
    fs, stride, th, m = int(g["fs"]), int(g["stride"]), float(g["threshold"]), int(g["model_type"])
    frames = torch.from_numpy(g["video"]).float()                 # what a float `read_data` would hand over
    D1, P1, s1 = compute_D1(frames, f, "RGB", slow=True, batch_size=48)
    assert D1.is_cuda and P1.is_cuda and s1.is_cuda and s1.dim() == 0
    np.testing.assert_allclose(D1.cpu().numpy(), g["ref_D1"], rtol=RTOL)
    np.testing.assert_allclose(s1.item(), g["ref_sigma1"], rtol=RTOL)
    np.testing.assert_allclose(P1[0].cpu().numpy(), g["ref_P1_row0"], rtol=RTOL)
    if m in (1, 2):
        D2, P2, s2, bf = compute_D2(D1, f, filter_size=fs)
    else:
        D2, P2, s2, bf = compute_D2(D1, f, filter_size=fs, stride=stride)
    assert tuple(bf.shape) == (1, 1, fs, fs)
    np.testing.assert_allclose(D2.cpu().numpy(), g["ref_D2"], rtol=RTOL)
    np.testing.assert_allclose(s2.item(), g["ref_sigma2"], rtol=RTOL)
    np.testing.assert_allclose(P2[0].cpu().numpy(), g["ref_P2_row0"], rtol=RTOL)
    D2_before = D2.clone()
    D3n, P3, P3n, s3 = q_learning(D2, f, thresholding=th)
    assert torch.equal(D2, D2_before)
    out = capsys.readouterr().out
    assert out.count("Eps:") == int(g["ref_n_sweeps"]) and "Non Zero in P3:" in out
    np.testing.assert_allclose(D3n.cpu().numpy(), g["ref_D3_new"], rtol=RTOL)
    np.testing.assert_allclose(s3.item(), g["ref_sigma3"], rtol=RTOL)
    np.testing.assert_allclose(P3.cpu().numpy(), g["ref_P3"], rtol=RTOL)
    rowptr, cols = eng.csr_from_matrix(P3n)
    np.testing.assert_array_equal(rowptr, g["ref_P3new_rowptr"])
    np.testing.assert_array_equal(cols, g["ref_P3new_cols"])
    np.random.seed(int(g["seed"]))
    frames_out, jumps = texture_walk(P3n, m, int(g["fps"]), int(g["nvl"]), stride, fs)
    np.testing.assert_array_equal(np.array(frames_out), g["walk_frames"])
    assert jumps == int(g["walk_jump_count"]) and ql_mod.LAST["n_sweeps"] == int(g["ref_n_sweeps"])


# ----------------------------------------------------------------------------- CLI orchestration
@pytest.mark.parametrize("model_type,stride,n,seed", [(1, 1, 300, 3), (2, 2, 300, 3), (3, 4, 520, 5)])
def test_cli_main_sigma_sweep_matches_oracle(eng, model_type, stride, n, seed, capsys):
    """`video_textures.main` (the -m 1/2/3 entry point) on a synthetic clip: for every sigma factor of the
    reference's sweep the emitted frame sequence equals the oracle's under the same numpy seed."""
    from audio_video_textures_b200.classic import video_textures as vt
    from audio_video_textures_b200.synth import synth_video
    from oracle import classic as oc
    h, w, fs, th, nvl = 8, 8, 16, 0.08, 2                       # clips screened offline for threshold margins
    args = vt.build_parser().parse_args(
        ["-m", str(model_type), "-fs", str(fs), "-stride", str(stride), "-t", str(th), "-nvl", str(nvl),
         "--synthetic", f"{n},{h},{w},{seed}"])
    video = synth_video(n, h, w, seed=seed)
    D1 = oc.pairwise_l2(video.float())
    eff_stride = 1 if model_type in (1, 2) else stride
    sweep, checked = list(vt.SIGMAS), 0
    try:
        for value in sweep:                                   # one sigma factor per call: independent RNG streams
            vt.SIGMAS[:] = [value]
            np.random.seed(5)
            res = vt.main(args, "synthetic")
            out = capsys.readouterr().out
            assert out.count("Frames list:") == 1 and "Eps:" in out and "Non Zero in P3:" in out
            f = torch.tensor(value, dtype=torch.float32)
            D2 = oc.compute_D2(D1, f, fs, eff_stride)[0]
            D3n, P3, P3n, s3 = oc.q_learning(D2, f, thresholding=th)
            np.testing.assert_allclose(res["sigmas"][0], float(s3), rtol=1e-5)
            np.random.seed(5)
            want, jumps = oc.walk(P3n, model_type, args.fps, nvl, stride, fs)
            m_rows = P3.shape[0]
            visited = sorted({r for f0 in want for r in (f0, min(f0 + stride, m_rows - 1)) if r < m_rows} | {100})
            if oc.threshold_margin(P3, th, rows=visited) < 1e-5:
                continue                                      # a visited row's survivor set hinges on an element at the cut
            assert res["sequences"][0] == want and res["jump_counts"][0] == jumps
            checked += 1
    finally:
        vt.SIGMAS[:] = sweep
    assert checked >= 2

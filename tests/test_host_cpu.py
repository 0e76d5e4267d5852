"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, host-side logic
(tile schedule, shard plan, sigma arithmetic, walk over survivor lists, argument surfaces) and the
N>1 exchange over gloo with world_size 2."""
import os
import re
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "avtex.h")).read()
    return sorted(set(re.findall(r"AVTEX_API\s+[\w\s\*]+?\b(avtex_\w+)\s*\(", text)))


def test_cabi_library_exports_every_declared_symbol():
    import ctypes
    from audio_video_textures_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/avtex.h but not exported"
    assert set(_lib.SIGNATURES) | {"avtex_last_error"} == set(names)
    assert lib.avtex_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.avtex_last_error(), bytes)
    assert isinstance(ctypes.CDLL(_lib.LIB_PATH), ctypes.CDLL)


def test_ctypes_signatures_match_header_prototypes():
    """Every prototype in include/avtex.h has as many parameters as the ctypes binding declares (ABI drift)."""
    from audio_video_textures_b200 import _lib
    text = open(os.path.join(ROOT, "include", "avtex.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"AVTEX_API\s+[\w\s\*]+?\b(avtex_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 25
    for name, params in protos:
        params = params.strip()
        n_params = 0 if params in ("", "void") else len(params.split(","))
        if name == "avtex_last_error":
            assert n_params == 0
            continue
        assert name in _lib.SIGNATURES, name
        assert len(_lib.SIGNATURES[name]) == n_params, (name, n_params, len(_lib.SIGNATURES[name]))


def test_product_path_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audio_video_textures_b200.classic.computeD1 import compute_D1
    with pytest.raises(RuntimeError, match="no CPU path"):
        compute_D1(torch.zeros(4, 2, 2, 3), 4.5, "RGB")
    import audio_video_textures_b200 as pkg
    src_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


@pytest.mark.parametrize("TM,TN", [(1, 1), (2, 1), (5, 3), (16, 8), (17, 9), (40, 20), (98, 391), (782, 391)])
def test_gram_tile_schedule_symmetric_covers_upper_triangle_once(TM, TN):
    from audio_video_textures_b200 import engine
    if TM > 2 * TN:
        pytest.skip("not a square matrix tiling")
    s = engine.gram_tile_schedule(TM, TN, True)
    need = {(tm, tn) for tm in range(TM) for tn in range(TN) if tn * 256 + 255 >= tm * 128}
    assert len(s) == len(set(s)) and set(s) == need


@pytest.mark.parametrize("TM,TN", [(1, 1), (3, 7), (8, 2), (9, 2), (98, 391)])
def test_gram_tile_schedule_rowblock_covers_all_once(TM, TN):
    from audio_video_textures_b200 import engine
    s = engine.gram_tile_schedule(TM, TN, False)
    assert sorted(s) == [(a, b) for a in range(TM) for b in range(TN)]
    # groups of 8 row tiles, column-major inside a group (L2 footprint of a wave)
    assert s[:min(16, TM)] == [(i, 0) for i in range(min(16, TM))]


def test_binomial_taps_and_sigma_arithmetic():
    from audio_video_textures_b200 import engine
    from oracle import classic
    for fs in (1, 2, 8, 16, 40, 64):
        np.testing.assert_array_equal(engine.binomial_taps(fs), classic.binomial_weights(fs).numpy())
    g = torch.Generator().manual_seed(0)
    D = torch.rand(300, 300, generator=g) * 1e4
    D.fill_diagonal_(0)
    f = torch.tensor(4.52, dtype=torch.float32)
    nnz = torch.nonzero(D).size(0)
    want = f * (D.sum() / nnz)
    got = engine.sigma_from_stats(float(D.double().sum()), nnz, f)
    np.testing.assert_allclose(got, want.item(), rtol=1e-6)
    assert isinstance(got, np.float32)
    assert engine.filtered_size(5000, 40, 4) == 1241 and engine.filtered_size(300, 40, 1) == 261


@pytest.mark.parametrize("name", ["classic_small_m1", "classic_ragged_m2", "classic_stride_m3"])
def test_walk_over_survivor_lists_matches_golden(name):
    """texture_walk consumes (rowptr, colidx) -- exactly what the GPU compaction returns."""
    from audio_video_textures_b200.classic.video_textures import texture_walk
    g = load_golden(name)
    np.random.seed(int(g["seed"]))
    frames, jumps = texture_walk((g["ref_P3new_rowptr"], g["ref_P3new_cols"]), int(g["model_type"]),
                                 int(g["fps"]), int(g["nvl"]), int(g["stride"]), int(g["fs"]))
    np.testing.assert_array_equal(np.array(frames), g["walk_frames"])
    assert jumps == int(g["walk_jump_count"])


def test_argument_surfaces_match_reference_defaults():
    from audio_video_textures_b200.classic.video_textures import build_parser as classic_parser
    from audio_video_textures_b200.contrastive.main import build_parser as cvt_parser
    a = classic_parser().parse_args([])
    assert (a.model_type, a.feats, a.filter_size, a.batch_size, a.stride, a.new_video_length, a.threshold,
            a.fps, a.sr, a.SF, a.slow, a.interpolation) == (1, "RGB", 40, 64, 4, 30, 0.08, 30, 22050, 3, False, True)
    a = classic_parser().parse_args("-m 3 -bs 48 -fs 16 -stride 2 -t 0.1 -s -nvl 10".split())
    assert (a.model_type, a.batch_size, a.filter_size, a.stride, a.threshold, a.slow, a.new_video_length) == \
        (3, 48, 16, 2, 0.1, True, 10)
    c = cvt_parser().parse_args([])
    assert (c.model_type, c.temp, c.threshold, c.alpha, c.mini_batchsize, c.window, c.stride,
            c.new_video_length, c.evaluate, c.da_feats) == (1, 0.1, 0.0, 0.5, 150, 20, 4, 30, False, "VGG")
    c = cvt_parser().parse_args("-e -th 0.3 -temp 0.1 -alpha 0.5 -mbs 100 -m 2".split())
    assert (c.evaluate, c.threshold, c.temp, c.alpha, c.mini_batchsize, c.model_type) == (True, 0.3, 0.1, 0.5, 100, 2)


@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 9, 20, 391])
def test_gram_tile_schedule_2cta(T):
    from audio_video_textures_b200 import engine
    s = engine.gram_tile_schedule(T, T, True, two_cta=True)
    assert len(s) == len(set(s)) and set(s) == {(a, b) for a in range(T) for b in range(T) if b >= a}
    s = engine.gram_tile_schedule(max(1, T // 3), T, False, two_cta=True)
    assert sorted(s) == [(a, b) for a in range(max(1, T // 3)) for b in range(T)]
    for group in (16, 5):                                  # the short-K schedule (16) and an odd size
        s = engine.gram_tile_schedule(T, T, True, group=group)
        assert len(s) == len(set(s)) and set(s) == {(a, b) for a in range(T) for b in range(T) if b >= a}
        s = engine.gram_tile_schedule(max(1, T // 3), T, False, group=group)
        assert sorted(s) == [(a, b) for a in range(max(1, T // 3)) for b in range(T)]


@pytest.mark.parametrize("n,fs,stride,world", [(5000, 40, 4, 8), (300, 40, 1, 2), (100000, 40, 4, 8), (520, 40, 4, 3)])
def test_shard_plan_covers_rows_and_halos(n, fs, stride, world):
    from audio_video_textures_b200 import dist as avd
    from audio_video_textures_b200 import engine
    m = engine.filtered_size(n, fs, stride)
    owned = []
    for r in range(world):
        p = avd.plan_shards(n, fs, stride, world, r)
        assert p.m == m and p.a0 == r * p.shard and p.a0 < p.a1 <= m
        assert p.a1h == min(m, p.a1 + 1)                              # one halo row for the P shift
        assert p.r_lo == p.a0 * stride and p.r_hi <= n
        assert p.r_hi == (n if p.a1 == m else (p.a1h - 1) * stride + fs)
        assert p.padded >= m and p.padded % world == 0
        owned.extend(range(p.a0, p.a1))
    assert owned == list(range(m))


@pytest.mark.parametrize("n,fs,stride,world", [(5000, 40, 4, 8), (100000, 40, 4, 8), (7072, 40, 4, 2), (2052, 16, 4, 3)])
def test_residue_class_shard_geometry(n, fs, stride, world):
    """The residue-class shards are planned in CLASS coordinates (N/s frames, fs/s taps, stride 1) by the same
    planner: output row ranges must coincide with the full plan's, the class rows held must cover what the
    plane-walking filter reads, and the buffer must stack `stride` planes."""
    from audio_video_textures_b200 import dist as avd
    for r in range(world):
        pb = avd.PeerBuffers(n, fs, stride, r, world, "cpu", residues=True)
        p, c = pb.plan, pb.cplans[r]
        assert (c.m, c.a0, c.a1, c.a1h) == (p.m, p.a0, p.a1, p.a1h)
        assert c.r_lo == p.a0 and c.r_hi >= (p.a1h - 1) + (fs - 1) // stride + 1        # last class row read + 1
        assert c.r_hi <= n // stride
        assert pb.rows_max == stride * pb.rows_c and pb.ld % 32 == 0 and pb.ld >= n // stride
    with pytest.raises(ValueError):
        avd.PeerBuffers(n + 1, fs, stride, 0, world, "cpu", residues=True)               # N % stride != 0
    with pytest.raises(ValueError):
        avd.PeerBuffers(n, fs, 1, 0, world, "cpu", residues=True)                        # stride 1: nothing to skip


@pytest.mark.parametrize("n,fs,stride,world", [(7072, 40, 4, 2), (12000, 40, 4, 8), (2051, 40, 4, 3), (333, 16, 1, 2),
                                               (300, 40, 1, 8), (100000, 40, 4, 8),
                                               (25000, 10, 1, 8), (1768, 10, 1, 2)])        # last two: class coordinates
def test_symmetric_shard_jobs_cover_every_needed_element_once(n, fs, stride, world):
    """Host logic of the peer-push scheme (no GPU, no process group): over all ranks, the direct and the
    transposed destinations of the job lists cover every (row, col) of every rank's CORE rows exactly once, and
    the halo pieces come from the owners' core rows only (also when shard*stride < fs: (300, 40, 1, 8))."""
    from audio_video_textures_b200 import dist as avd
    plans = [avd.plan_shards(n, fs, stride, world, r) for r in range(world)]
    ld = (n + 31) // 32 * 32
    ptrs = [1000 + r for r in range(world)]                  # stand-in "pointers" = rank ids
    big = n > 20000                                          # 100k: interval bookkeeping only (no n x n cover arrays)
    cover = None if big else [np.zeros((p.r_hi - p.r_lo, n), dtype=np.int16) for p in plans]
    work = []
    for r in range(world):
        jobs = avd.symmetric_jobs(plans, r, ptrs, ld, stride)
        assert len(jobs) <= 16
        pairs = 0
        for j in jobs:
            rs, cs = slice(j["row0"], j["row0"] + j["rows"]), slice(j["col0"], j["col0"] + j["cols"])
            assert j["ldd"] == ld and j["ldt"] == ld
            tri = None
            if j["symmetric"]:
                assert j["row0"] == j["col0"] and j["rows"] == j["cols"]
                tri = None if big else np.triu(np.ones((j["rows"], j["cols"]), dtype=np.int16))
            owner_d, owner_t = j["D"] - 1000, j["DT"] - 1000
            assert owner_d == r
            lo_t, hi_t = avd.core_rows(plans, owner_t, stride)
            assert lo_t <= cs.start and cs.stop <= hi_t        # pushes land in the destination's CORE rows only
            if not big:
                blk = cover[owner_d][rs.start - j["d_row0"]: rs.stop - j["d_row0"], cs]
                blk += 1 if tri is None else tri
                blk = cover[owner_t][cs.start - j["dt_row0"]: cs.stop - j["dt_row0"], rs]
                blk += 1 if tri is None else (np.triu(np.ones((j["rows"], j["cols"]), dtype=np.int16), 1)).T
            pairs += j["rows"] * j["cols"] // (2 if j["symmetric"] else 1)
        work.append(pairs)
    for r in range(world):
        lo, hi = avd.core_rows(plans, r, stride)
        core = hi - plans[r].r_lo
        if not big:
            c = cover[r]
            assert c[:core].min() == 1 and c[:core].max() == 1 and (c[core:] == 0).all()
        # halo: contiguous pieces that start where the core ends, each inside its owner's core
        need = hi
        for owner, row, rows in avd.halo_sources(plans, r, stride):
            olo, ohi = avd.core_rows(plans, owner, stride)
            assert owner > r and row == need and olo <= row and row + rows <= ohi
            need += rows
        assert need == plans[r].r_hi
        if r + 1 < world:
            assert plans[r + 1].r_lo == hi
    assert sum(work) >= n * n // 2 - n                                    # nothing is left uncomputed
    assert max(work) <= 1.25 * (n * n / (2 * world)) + 300 * n           # balanced up to halos and tile rounding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, D3_np, alpha, out_dir):
    """Emulates the sharded future-cost loop of dist.classic_sharded with torch-CPU arithmetic in
    place of the kernels (test only) and the REAL exchange / stats all-reduce over gloo."""
    import torch.distributed as dist
    from audio_video_textures_b200 import dist as avd
    from audio_video_textures_b200 import engine
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D3 = torch.from_numpy(D3_np)
    M = D3.shape[0]
    shard = -(-M // world)
    plan = avd.ShardPlan(0, M, world, rank, shard, rank * shard, min(M, (rank + 1) * shard),
                         min(M, (rank + 1) * shard + 1), 0, 0)
    exchange = avd.make_exchange(plan)
    rows = D3[plan.a0:plan.a1]
    a32 = torch.tensor(alpha, dtype=torch.float32)

    def sweep(prev, prev2, out, eps):
        X = rows if prev is None else rows + a32 * prev[:M]
        if prev is not None and plan.a0 == 0:
            X = X.clone(); X[0] = rows[0]
        Y = X.clone()
        idx = torch.arange(plan.a0, plan.a1)
        Y[idx - plan.a0, idx] = float("inf")
        out[plan.a0:plan.a1] = Y.min(1)[0]
        if eps is not None:
            Xp = rows if prev2 is None else rows + a32 * prev2[:M]
            if plan.a0 == 0:
                Xp = Xp.clone(); Xp[0] = rows[0]
            eps += ((X - Xp) ** 2).double().sum()

    bufs = [torch.zeros(plan.padded) for _ in range(3)]
    eps = torch.zeros(1, dtype=torch.float64)
    sweep(None, None, bufs[0], None)
    exchange(bufs[0], None)
    cur, prev2, free, n_sweeps = bufs[0], None, [bufs[1], bufs[2]], 0
    for p in range(1, 100):
        out = free.pop()
        eps.zero_()
        sweep(cur, prev2, out, eps)
        exchange(out, eps)
        e = np.float32(eps.item() / (M * M))
        if not (e > np.float32(engine.F32_EPS_STOP)):
            n_sweeps = p
            break
        if prev2 is not None:
            free.append(prev2)
        prev2, cur = cur, out
    D3_new = rows + a32 * cur[:M]
    if plan.a0 == 0:
        D3_new[0] = rows[0]
    stats = torch.zeros(2, dtype=torch.float64)
    stats[0] = D3_new.double().sum()
    stats.view(torch.int64)[1] = int((D3_new != 0).sum())
    tot = avd.allreduce_stats(stats)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), D3_new=D3_new.numpy(), n_sweeps=n_sweeps,
             total=float(tot[0]), nnz=int(tot.view(torch.int64)[1]), a0=plan.a0)
    dist.destroy_process_group()


def test_row_sharded_future_cost_exchange_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import classic
    g = load_golden("classic_small_m1")
    D3 = torch.from_numpy(g["ref_D2"]) ** 0.7
    want, trail = classic.future_cost(D3)
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, D3.numpy(), 0.997, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    got = np.concatenate([p["D3_new"] for p in parts])
    np.testing.assert_array_equal(got, want.numpy())                  # sharded == single process, bit for bit
    assert all(int(p["n_sweeps"]) == len(trail) for p in parts)
    np.testing.assert_allclose(parts[0]["total"], want.double().sum().item(), rtol=1e-12)
    assert int(parts[0]["nnz"]) == int((want != 0).sum()) == int(parts[1]["nnz"])


def test_device_generator_reproduces_numpy_legacy_randint():
    """The MT19937 + masked-rejection randint the synthesis loop kernel runs on the GPU (same source, host build)
    against numpy itself: np.random.choice(a) == a[randint(0, len(a))], state handed over and back."""
    from audio_video_textures_b200 import engine
    for seed in (0, 1, 12345):
        np.random.seed(seed)
        np.random.rand(seed % 7)                                   # start from an arbitrary position in the stream
        kind, key, pos, has_gauss, cached = np.random.get_state()
        rs = np.random.RandomState(seed + 99)
        ns = np.concatenate([rs.randint(1, 300, 700), [1, 1, 2, 3, 255, 256, 257, 65535, 65536, 2 ** 31 - 1, 4_000_000_000]])
        want = np.array([np.random.choice(np.arange(n)) if n < 10 ** 6 else np.random.randint(0, n) for n in ns], dtype=np.uint64)
        got, new_key, new_pos = engine.mt19937_randint_host(key, pos, ns)
        np.testing.assert_array_equal(got.astype(np.uint64), want)
        k2, key2, pos2 = np.random.get_state()[:3]
        assert pos2 == new_pos and np.array_equal(key2, new_key)   # 700+ draws cross at least one 624-word refill
        np.random.set_state((kind, new_key, new_pos, has_gauss, cached))
        a, b = np.random.randint(0, 1000), np.random.rand()
        np.random.set_state((kind, key2, pos2, has_gauss, cached))
        assert a == np.random.randint(0, 1000) and b == np.random.rand()


def test_symmetric_provenance_registry():
    """engine.known_symmetric: only the very tensor object that was marked, and only while nobody wrote to it."""
    from audio_video_textures_b200 import engine
    a = torch.zeros(8, 8)
    assert not engine.known_symmetric(a)
    engine.mark_symmetric(a)
    assert engine.known_symmetric(a)
    assert not engine.known_symmetric(a[:4]) and not engine.known_symmetric(a.clone())
    a[0, 1] = 3.0                                           # in-place edit bumps torch's version counter
    assert not engine.known_symmetric(a)
    for _ in range(40):                                     # the registry is bounded and drops dead tensors
        engine.mark_symmetric(torch.zeros(2, 2))
    assert len(engine._SYMMETRIC) <= 16


def test_residue_eligibility_and_job_array():
    from audio_video_textures_b200 import engine, _lib
    def pf(n, k, pitch=None, signed=False):
        pitch = k if pitch is None else pitch
        buf = torch.zeros(n * pitch, dtype=torch.int8 if signed else torch.uint8)
        x = buf.as_strided((n, k), (pitch, 1))
        return engine.PackedFrames(x, torch.zeros(n, dtype=torch.int64), k, torch.zeros(2, dtype=torch.int64), signed=signed)
    assert engine.residue_eligible(pf(1000, 768), 40, 4) and engine.residue_eligible(pf(1000, 768), 16, 4)
    assert not engine.residue_eligible(pf(1001, 768), 40, 4)          # N % stride
    assert not engine.residue_eligible(pf(1000, 192), 40, 4)          # K % 128
    assert not engine.residue_eligible(pf(1000, 768, pitch=784), 40, 4)   # padded rows: the [N/s, s*K] view needs dense rows
    assert not engine.residue_eligible(pf(1000, 768), 40, 1) and not engine.residue_eligible(pf(1000, 768), 5, 2)
    arr = engine.gram_job_array([dict(row0=1, rows=2, col0=3, cols=4, symmetric=0, count_stats=1, D=1000, d_row0=1, ldd=64,
                                      DT=None, k_off=768, sq_off=3, sq_stride=4),
                                 dict(row0=0, rows=8, col0=0, cols=8, symmetric=1, D=2000, ldd=32, DT=2000, ldt=32)])
    assert isinstance(arr, _lib.C.Array) and len(arr) == 2
    assert (arr[0].k_off, arr[0].sq_off, arr[0].sq_stride, arr[0].DT) == (768, 3, 4, None)
    assert (arr[1].k_off, arr[1].sq_off, arr[1].sq_stride, arr[1].symmetric, arr[1].DT) == (0, 0, 1, 1, 2000)


def test_walk_draw_is_numpy_choice():
    """texture_walk draws with a[np.random.randint(0, len(a))]; the reference with np.random.choice(a)
    (classic/video_textures.py:78): same value AND same generator state afterwards, for every list length."""
    lists = [np.arange(n, dtype=np.int32) * 3 + 1 for n in (1, 2, 3, 5, 600, 1000, 7, 1, 33, 4096, 100000, 2 ** 16 + 1)]
    np.random.seed(11)
    a = [int(np.random.choice(x)) for x in lists * 40]
    sa = np.random.get_state()
    np.random.seed(11)
    b = [int(x[np.random.randint(0, len(x))]) for x in lists * 40]
    sb = np.random.get_state()
    assert a == b and sa[2] == sb[2] and np.array_equal(sa[1], sb[1])


def test_planned_steps_matches_the_loop():
    from audio_video_textures_b200.contrastive.validate import planned_steps
    for max_length, W, S, ssr in ((900, 15, 6, 1), (900, 20, 4, 1), (15, 15, 6, 1), (16, 15, 6, 1), (0, 15, 6, 1), (630, 15, 6, 2)):
        n_frames, steps = 0, 0
        while n_frames < max_length:
            n_frames += (W if steps == 0 else S) * ssr
            steps += 1
        assert planned_steps(max_length, W, S, ssr) == steps
    assert planned_steps(900, 15, 6, 1, max_steps=7) == 7


def test_contrastive_parser_accepts_every_reference_flag():
    """A full reference command line (every flag of contrastive_video_textures/main.py:41-296) parses unchanged."""
    from audio_video_textures_b200.contrastive.main import build_parser
    line = ("-ea slowfast -m 2 -vdata v -adata a -pdata p -fdata f -dadata d -vl a b -fps 25 -subsample 2 -temp 0.2 -th 0.3 "
            "-l2 -nintp -size 112 -negs 10 -w 16 -train_stride 2 -stride 3 -nvl 20 -alpha 0.7 -SF 3 -long -fb --epochs 5 "
            "--size 112 --start_epoch 2 -bs 8 -mbs 100 -lr 0.1 --lr_steps 10 --momentum 0.8 --wd 0.001 -j 2 -p 1 -lf 2 "
            "--resume x.pth -e -da t1 t2 -daf Contrastive -daf_resume c1 c2 -ve -vf 3 --logdir l --logname n -rf r --ckpt c")
    a = build_parser().parse_args(line.split())
    assert (a.enc_arch, a.model_type, a.subsample_rate, a.temp, a.threshold, a.l2, a.interpolation, a.img_size) == \
        ("slowfast", 2, 2, 0.2, 0.3, False, False, 112)
    assert (a.driving_audio, a.da_feats, a.daf_resume, a.evaluate, a.mini_batchsize, a.alpha, a.weight_decay) == \
        (["t1", "t2"], "Contrastive", ["c1", "c2"], True, 100, 0.7, 0.001)
    d = build_parser().parse_args([])
    assert (d.lr, d.epochs, d.workers, d.ckpt, d.train_stride, d.n_negs, d.long, d.frames_bar) == \
        (10e-3, 60, 4, "./ckpt", 4, 20, False, False)

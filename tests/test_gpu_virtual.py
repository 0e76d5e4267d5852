"""Multi-GPU parity on ONE GPU: G "virtual ranks" (dist.VirtualBox) run the row-sharded classic++ pipeline with
the same Gram job lists (transposed tiles pushed into the other ranks' shards), norm pushes, halo plans and the
in-kernel flag barrier of the fused future cost as a real G-GPU box; every rank's shard must equal the single-GPU
result bit for bit (SURVEY.md §4 "virtual ranks on one GPU").  Covers G = 8, ragged last shards, the
shard*stride < filter_size halo case and repeated calls (buffer parities / flag epochs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(n, h, w, fs, stride, world, calls=1, residues=False):
    from audio_video_textures_b200 import dist as avd
    from audio_video_textures_b200 import engine, selfcheck
    from audio_video_textures_b200.classic.video_textures import texture_walk
    from audio_video_textures_b200.synth import synth_video
    frames = synth_video(n, h, w, seed=2).cuda()
    box = avd.VirtualBox(n, fs, stride, world, frames.device, residues=residues)
    for _ in range(calls):
        results = box.step(frames, sigma_factor=4.5, threshold=0.08)
    torch.cuda.synchronize()
    single = selfcheck.single_gpu_pipeline(frames, fs, stride, 4.5, 0.08)
    for r, res in enumerate(results):
        ok = selfcheck.shard_equals_single(res, single)
        assert all(ok.values()), (r, world, ok)
    rowptr, colidx = avd.virtual_survivors(results)
    rp1, ci1 = engine.csr_from_matrix(single["P3n"], single["counts"])
    assert np.array_equal(rowptr, rp1) and np.array_equal(colidx, ci1)
    np.random.seed(0)
    a = texture_walk((rowptr, colidx), 1, 30, 5, stride, fs)
    np.random.seed(0)
    b = texture_walk((rp1, ci1), 1, 30, 5, stride, fs)
    assert a == b
    return results


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,h,w,fs,stride", [(900, 16, 16, 40, 4), (2051, 8, 8, 40, 4), (1500, 12, 20, 16, 1)])
def test_virtual_ranks_equal_single_gpu(n, h, w, fs, stride, world):
    _check(n, h, w, fs, stride, world)


def test_virtual_ranks_halo_spans_several_ranks():
    """N = 300 (C1's frame count), fs = 40, stride 1, 8 ranks: shard*stride = 36 rows < fs, so rank r's 40 halo rows
    run through the whole core of rank r+1 into rank r+2's — the case the round-1 halo copy raced on."""
    from audio_video_textures_b200 import dist as avd
    plans = [avd.plan_shards(300, 40, 1, 8, r) for r in range(8)]
    assert any(len(avd.halo_sources(plans, r, 1)) > 1 for r in range(7))
    _check(300, 12, 20, 40, 1, 8)


def test_virtual_ranks_repeated_calls_and_c5_shape():
    """8 ranks on the C5 layout (64x64 frames, -m 3) at a reduced frame count, three calls on the same buffers."""
    res = _check(6000, 64, 64, 40, 4, 8, calls=3)
    assert res[0].plan.m == 1491 and res[0].n_sweeps >= 2


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("n,h,w,fs", [(1200, 16, 16, 40), (2052, 16, 8, 16), (6000, 64, 64, 40)])
def test_virtual_ranks_residue_class_shards(n, h, w, fs, world):
    """The sharded pipeline on residue-class planes (stride 4: K1 computes 1/4 of the pairs; 8 ranks x 4 classes = 32
    Gram jobs in one launch, per-plane halos, the row-shard form of the plane-walking filter): every shard equals the
    single-GPU FULL-D1 pipeline bit for bit, two calls on the same buffers."""
    _check(n, h, w, fs, 4, world, calls=2, residues=True)


def test_lazy_survivor_rows_equal_bulk_csr():
    """engine.SurvivorRows (per-row device compaction on demand, what the large-N walk uses) == the bulk CSR."""
    from audio_video_textures_b200 import engine, selfcheck
    from audio_video_textures_b200.classic.video_textures import texture_walk
    from audio_video_textures_b200.synth import synth_video
    frames = synth_video(900, 16, 16, seed=4).cuda()
    single = selfcheck.single_gpu_pipeline(frames, 40, 1, 4.5, 0.08)
    rowptr, colidx = engine.csr_from_matrix(single["P3n"], single["counts"])
    rows = engine.SurvivorRows.from_matrix(single["P3n"])
    for i in (0, 1, 100, 500, len(rows) - 1):
        np.testing.assert_array_equal(rows[i], colidx[rowptr[i]:rowptr[i + 1]])
    for mode in (1, 2, 3):
        np.random.seed(3)
        a = texture_walk((rowptr, colidx), mode, 30, 10, 1, 40)
        np.random.seed(3)
        b = texture_walk(engine.SurvivorRows.from_matrix(single["P3n"]), mode, 30, 10, 1, 40)
        assert a == b

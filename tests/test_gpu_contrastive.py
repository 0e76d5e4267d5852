"""GPU parity tests for the contrastive synthesis path (K6/K7) against the golden vectors produced
by the unmodified ContrastivePredictionTemporal.forward and against the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from audio_video_textures_b200 import engine
    assert torch.cuda.is_available()
    return engine


def test_l2_normalize_rows(eng):
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(77, 2304, generator=gen)
    x[5] = 0                                           # zero row: x / max(0, 1e-12) = 0
    y = eng.l2_normalize_rows(x.cuda()).cpu()
    np.testing.assert_allclose(y.numpy(), F.normalize(x, dim=1).numpy(), rtol=1e-6, atol=1e-12)
    assert float(y[5].abs().max()) == 0.0


def test_scores_match_reference_forward(eng):
    from oracle import contrastive as oc
    g = load_golden("contrastive_small")
    emb = torch.from_numpy(g["emb"])
    L = emb.shape[0]
    tn = eng.l2_normalize_rows(emb.cuda())
    for q, key in ((10, "ref_logits_q10"), (L - 1, "ref_logits_qlast")):
        o = eng.cosine_scores(tn, tn[q], float(g["temp"])).cpu()
        ids = oc.target_order(q, L)
        np.testing.assert_allclose(o[ids].numpy(), g[key], rtol=1e-5, atol=2e-6)


def test_similarity_tail_matches_oracle_chunk(eng):
    from audio_video_textures_b200.contrastive.models import ContrastivePredictionTemporal, similarity_tail
    from oracle import contrastive as oc
    gen = torch.Generator().manual_seed(1)
    q = torch.randn(3, 200, generator=gen)
    t = torch.randn(3, 17, 200, generator=gen)
    want = oc.similarity_chunk(q, t, 0.07)
    got = similarity_tail(q.cuda(), t.cuda(), 0.07).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=2e-6)
    # model_type 2 + driving audio through the class (embedding boundary)
    qa, ta, da = torch.rand(3, 24, generator=gen), torch.rand(3, 17, 24, generator=gen), torch.rand(3, 24, generator=gen)
    m = ContrastivePredictionTemporal(model_type=2, temp=0.1)
    out, out_a = m(q.cuda(), t.cuda(), q_audio_eg=qa.cuda(), t_audio_eg=ta.cuda(), driving_audio=da.cuda())
    want = oc.similarity_chunk(torch.cat((q, qa), 1), torch.cat((t, ta), 2), 0.1)
    want_a = oc.audio_chunk(da, ta, 0.1)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out_a.cpu().numpy(), want_a.numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("L,q", [(96, 10), (96, 95), (96, 0), (96, 94), (5000, 1234), (2, 0), (2, 1), (1025, 1024)])
@pytest.mark.parametrize("audio", [False, True])
def test_select_step_matches_oracle(eng, L, q, audio):
    """Same logits on both sides -> same survivor list (bit-exact), values rtol 1e-5."""
    from oracle import contrastive as oc
    gen = torch.Generator().manual_seed(L * 7 + q)
    o_nat = torch.randn(L, generator=gen) * 2 + 6          # logits in natural window order
    a_nat = torch.rand(L, generator=gen) * 9 + 0.5 if audio else None
    th, alpha = 0.3, 0.5
    ids = oc.target_order(q, L)
    out, choices, mixed = oc.mix_and_select(o_nat[ids], a_nat[ids] if audio else None, alpha, th)
    if oc.select_margin(mixed, th) < 1e-5:
        pytest.skip("fixture too close to the threshold cut")
    sel = torch.zeros(L + 1, dtype=torch.int32, device="cuda")
    vals = torch.zeros(L, device="cuda")
    eng.select_step(o_nat.cuda(), a_nat.cuda() if audio else None, q, alpha, th, sel[1:], sel[:1], vals)
    n = int(sel[0])
    got = sel[1:n + 1].cpu().numpy()
    np.testing.assert_array_equal(got, ids[choices.numpy()])
    np.testing.assert_allclose(vals.cpu()[ids].numpy(), out.numpy(), rtol=1e-5, atol=0)


@pytest.mark.parametrize("L,D,A,q", [(96, 64, 0, 10), (96, 64, 32, 95), (5000, 256, 0, 1234), (5000, 256, 128, 0),
                                     (1025, 100, 0, 1024), (2, 8, 0, 1), (3000, 2304, 128, 77)])
def test_fused_synthesis_step_equals_separate_kernels(eng, L, D, A, q):
    """avtex_synthesis_step (one cooperative launch, result in mapped pinned memory) == avtex_cosine_scores +
    avtex_select_step: same logits bit for bit, same survivor list, same renormalised values; repeated calls
    alternate the accumulator parity; a survivor list longer than the pinned window falls back to a D2H copy."""
    from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings
    tn = eng.l2_normalize_rows(synth_embeddings(L, D, seed=L + q).cuda())
    sn = dn = None
    if A:
        sn = eng.l2_normalize_rows(synth_audio_features(L, A, seed=1).cuda())
        dn = eng.l2_normalize_rows(synth_audio_features(4, A, seed=2).cuda())
    th, alpha, temp = 0.3, 0.5, 0.1
    o = eng.cosine_scores(tn, tn[q], temp)
    a = eng.cosine_scores(sn, dn[1], temp) if A else None
    sel = torch.zeros(L + 1, dtype=torch.int32, device="cuda")
    vals = torch.zeros(L, device="cuda")
    eng.select_step(o, a, q, alpha, th, sel[1:], sel[:1], vals)
    want = sel[1:int(sel[0]) + 1].cpu().numpy()
    for cap in (4096, 3):
        ws = eng.SynthesisWorkspace(L, "cuda", host_cap=cap)
        v2 = torch.zeros(L, device="cuda")
        for rep in range(3):                                   # parity 1, 0, 1
            got = eng.synthesis_step(ws, tn, tn[q], q, temp, alpha, th, sn, dn[1] if A else None, v2)
            np.testing.assert_array_equal(got, want)
        assert torch.equal(ws.f32[:L], o)
        if A:
            assert torch.equal(ws.f32[L:2 * L], a)
        mask = torch.ones(L, dtype=torch.bool)
        if q != L - 1:
            mask[q] = False
        assert torch.equal(v2.cpu()[mask], vals.cpu()[mask])


@pytest.mark.parametrize("audio", [False, True])
def test_device_loop_equals_host_loop_and_hands_the_generator_back(eng, audio):
    """engine.synthesis_loop (the whole loop in one persistent kernel, np.random.choice drawn on the device from
    numpy's own MT19937 state) == one launch per step with the draw on the host: same windows, same survivor
    counts, and numpy's global generator ends in the SAME state (the next host draw is identical)."""
    from audio_video_textures_b200.contrastive.validate import synthesize
    from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings
    emb = synth_embeddings(3000, 256, seed=3).cuda()
    kw = {}
    if audio:
        kw = dict(alpha=0.5, q_audio=synth_audio_features(3000, 64, seed=0).cuda(),
                  da_source=synth_audio_features(3000, 64, seed=1).cuda(),
                  da_driving=synth_audio_features(200, 64, seed=2).cuda())
    args = dict(temp=0.1, threshold=0.3, fps=30, new_video_length=20, window=15, stride=6)
    np.random.seed(21)
    host = synthesize(emb, device_loop=False, **args, **kw)
    after_host = np.random.randint(0, 1 << 30)
    np.random.seed(21)
    dev = synthesize(emb, device_loop=True, **args, **kw)
    after_dev = np.random.randint(0, 1 << 30)
    assert dev["q_ids"] == host["q_ids"] and dev["nz_counts"] == host["nz_counts"]
    assert dev["frame_ids"] == host["frame_ids"] and dev["jump_count"] == host["jump_count"]
    assert after_dev == after_host
    assert len(dev["q_ids"]) >= 98


def test_synthesis_sequences_bit_exact(eng):
    from audio_video_textures_b200.contrastive.validate import start_segment, synthesize
    g = load_golden("contrastive_small")
    emb = torch.from_numpy(g["emb"]).cuda()
    kw = dict(temp=float(g["temp"]), threshold=float(g["th"]), fps=int(g["fps"]), new_video_length=int(g["nvl"]),
              window=int(g["window"]), stride=int(g["stride"]), mini_batchsize=int(g["mbs"]))
    np.random.seed(int(g["seed"]))
    r = synthesize(emb, q_start=10, **kw)
    np.testing.assert_array_equal(r["q_ids"], g["synth_q_ids"])
    np.testing.assert_array_equal(r["frame_ids"], g["synth_frame_ids"])
    np.testing.assert_array_equal(r["nz_counts"], g["synth_nz"])
    assert r["jump_count"] == int(g["synth_jumps"])
    qa, das, dad = (torch.from_numpy(g[k]).cuda() for k in ("q_audio", "da_source", "da_driving"))
    assert start_segment(das, dad[0]) == int(g["audio_start"])
    np.random.seed(int(g["seed"]))
    r2 = synthesize(emb, alpha=0.5, q_audio=qa, da_source=das, da_driving=dad, **kw)
    assert r2["start"] == int(g["audio_start"])
    np.testing.assert_array_equal(r2["q_ids"], g["synth2_q_ids"])
    np.testing.assert_array_equal(r2["frame_ids"], g["synth2_frame_ids"])
    np.testing.assert_array_equal(r2["nz_counts"], g["synth2_nz"])


def test_synthesis_medium_against_oracle(eng):
    """L = 3000, D = 256: the CUDA loop and the CPU oracle choose the same windows under one seed."""
    from audio_video_textures_b200.contrastive.validate import synthesize
    from audio_video_textures_b200.synth import synth_embeddings
    from oracle import contrastive as oc
    emb = synth_embeddings(3000, 256, seed=3)
    np.random.seed(11)
    want = oc.synthesize(emb, 0.1, 0.3, 150, 30, 10, 15, 6, q_start=10, return_debug=True)
    if min(want["margins"]) < 1e-5:
        pytest.skip("fixture too close to the threshold cut")
    np.random.seed(11)
    got = synthesize(emb.cuda(), temp=0.1, threshold=0.3, fps=30, new_video_length=10, window=15, stride=6,
                     q_start=10)
    assert got["q_ids"] == want["q_ids"] and got["frame_ids"] == want["frame_ids"]
    assert got["jump_count"] == want["jump_count"] and got["nz_counts"] == want["nz_counts"]


def _full_size_case(A, nvl, seed):
    """BASELINE configs[2]/[3] at full size (L = 20000 windows, D = 2304 [+ A audio dims]) against
    oracle.contrastive.synthesize on the same tables: chosen windows, emitted frames, survivor counts and jump
    count bit-exact under one numpy seed; the oracle's minimum threshold margin is printed.
    Reference: cvt/models/models.py:351-352,412-457; cvt/validate.py:369-378,524-572."""
    from audio_video_textures_b200.contrastive.validate import synthesize
    from audio_video_textures_b200.synth import synth_audio_features, synth_embeddings
    from oracle import contrastive as oc
    L, D = 20000, 2304
    emb = synth_embeddings(L, D, seed=0, device="cuda")
    kw, cpu_kw = {}, {}
    if A:
        qa = synth_audio_features(L, A, seed=0, device="cuda")
        das = synth_audio_features(L, A, seed=1, device="cuda")
        dad = synth_audio_features(160, A, seed=2, device="cuda")
        kw = dict(alpha=0.5, q_audio=qa, da_source=das, da_driving=dad)
        cpu_kw = dict(alpha=0.5, q_audio=qa.cpu(), da_source=das.cpu(), da_driving=dad.cpu())
    np.random.seed(seed)
    got = synthesize(emb, temp=0.1, threshold=0.3, fps=30, new_video_length=nvl, window=15, stride=6, **kw)
    np.random.seed(seed)
    want = oc.synthesize(emb.cpu(), 0.1, 0.3, 150, 30, nvl, 15, 6, q_start=got["start"], return_debug=True, **cpu_kw)
    print(f"L={L} D={D} A={A}: {len(want['q_ids'])} steps, oracle min threshold margin {min(want['margins']):.3e}, "
          f"mean survivors {np.mean(want['nz_counts']):.1f}, jumps {want['jump_count']}")
    if min(want["margins"]) < 1e-5:
        pytest.skip("fixture too close to the threshold cut")
    assert got["q_ids"] == want["q_ids"] and got["frame_ids"] == want["frame_ids"]
    assert got["nz_counts"] == want["nz_counts"] and got["jump_count"] == want["jump_count"]
    return got


def test_c3_full_size_synthesis_matches_oracle(eng):
    got = _full_size_case(A=0, nvl=30, seed=5)           # -e -th 0.3 -temp 0.1, 30 s at 30 fps: 150 steps
    assert len(got["q_ids"]) >= 148


def test_c4_full_size_audio_conditioned_a128(eng):
    got = _full_size_case(A=128, nvl=30, seed=6)         # -m 2 -alpha 0.5, canonical 128-d VGGish: 150 steps
    assert len(got["q_ids"]) >= 148


def test_c4_full_size_audio_conditioned_a12288(eng):
    got = _full_size_case(A=12288, nvl=8, seed=7)        # reference-faithful 12288-d conv map: 39 steps (CPU time)
    assert len(got["q_ids"]) >= 38


def test_cli_main_synthesis_mode(eng, capsys):
    """`contrastive.main -e -th -temp -mbs` on synthetic embeddings: same windows as the oracle."""
    from audio_video_textures_b200.contrastive import main as cm
    from audio_video_textures_b200.synth import synth_embeddings
    from oracle import contrastive as oc
    args = cm.build_parser().parse_args("-e -th 0.3 -temp 0.1 -mbs 100 -nvl 3 --synthetic 400,64,2 --seed 9".split())
    res = cm.main(args)
    assert "Chosen windows:" in capsys.readouterr().out
    emb = synth_embeddings(400, 64, seed=2, device="cuda").cpu()
    np.random.seed(9)
    want = oc.synthesize(emb, 0.1, 0.3, 100, 30, 3, 15, 6, q_start=10, return_debug=True)
    if min(want["margins"]) > 1e-5:
        assert res["q_ids"] == want["q_ids"] and res["jump_count"] == want["jump_count"]


def test_compute_Paudio_dropin(eng):
    """classic/computePaudio.py: the audio prior over T source examples."""
    from audio_video_textures_b200.classic.computePaudio import compute_Paudio
    from audio_video_textures_b200.synth import synth_audio_features
    from oracle import classic as oc
    t_a = synth_audio_features(300, 128, seed=4)
    d = synth_audio_features(3, 128, seed=5)[1]
    want = oc.compute_Paudio(t_a, d)
    got = compute_Paudio(t_a.cuda(), d.cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(float(got.sum()), 1.0, rtol=1e-5)

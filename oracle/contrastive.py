"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (torch-CPU + numpy) of the contrastive synthesis hot path of
medhini/audio-video-textures AT THE EMBEDDING BOUNDARY (per-window embeddings are inputs;
the 3D-CNN / VGGish encoders are feature producers and out of scope):

    similarity tail   contrastive_video_textures/models/models.py:351-352, 412-417
    driving-audio     contrastive_video_textures/models/models.py:419-457
    target ordering   contrastive_video_textures/validate.py:369-378
    chunk layout      contrastive_video_textures/validate.py:409-411, 442-493, 522
    normalise / mix   contrastive_video_textures/validate.py:524-527
    select / sample   contrastive_video_textures/validate.py:554-572
    emitted frames    contrastive_video_textures/validate.py:581-612
    start search      contrastive_video_textures/validate.py:218-242

Parity status: "parity unpinned" by the reference (no tests / vectors).  `similarity_chunk`
is pinned against the UNMODIFIED `ContrastivePredictionTemporal.forward` imported in the build
container with identity encoders (oracle/ref_shim.py, tests/golden/contrastive_*.npz);
`validate()` itself cannot run (video decode, undefined args.vcam), so the selection block is
a line-by-line restatement.
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch
import torch.nn.functional as F


def target_order(q_id: int, L: int) -> np.ndarray:
    """validate.py:369-378 — [pos] ++ ascending(all ids minus {q, pos}), pos = min(q+1, L-1)."""
    all_segment_ids = np.arange(L)
    pos_id = min(q_id + 1, L - 1)
    mask = np.ones(L, dtype=bool)
    mask[[q_id, pos_id]] = False
    n_ids = all_segment_ids[mask, ...]
    return np.concatenate((np.array([pos_id]), n_ids), axis=0)


def frame_level_step_logits(frames: torch.Tensor, q_id: int, L: int, W: int, S: int, mbs: int, temp: float, embed):
    """One synthesis step from FRAMES, as the reference runs it (num_gpus = 1): validate.py:329 (query window),
    :365-388 (target frames: union of the target windows in first-seen order), utils.py:233-260
    (split_into_overlapping_segments incl. its start = idx*S*(mbs-1)), models.py:355-362 (mbs windows per chunk),
    :351-352,412-417 (similarity), validate.py:481-493,522 (result layout).  `embed(windows [n, W, ...]) -> [n, D]`
    stands for the encoder stack.  Returns (target_segment_ids, logits [n_targets])."""
    ids = target_order(q_id, L)
    tf = []
    for i in ids:
        tf.extend(list(np.arange(i * S, i * S + W)))
    tf = np.array(tf)
    _, first = np.unique(tf, return_index=True)
    tf = tf[np.sort(first)]
    t_video = frames[tf]
    num_inputs = t_video.size(0)
    total_segments = math.ceil((num_inputs - W) / S)
    chunk_size = mbs * S + W
    batch_size = math.ceil(total_segments / mbs)
    batched = torch.zeros(*([batch_size, chunk_size] + list(t_video.size()[1:])), dtype=t_video.dtype)
    for idx in range(batch_size):
        start = idx * S * (mbs - 1)
        end = min(start + chunk_size, num_inputs)
        batched[idx, :end - start] = t_video[start:end]
    q = embed(frames[q_id * S: q_id * S + W].unsqueeze(0))
    output = torch.zeros(len(ids), dtype=torch.float32)
    num_valid = len(ids)
    for itr in range(batch_size):
        wins = torch.stack([batched[itr, i * S: i * S + W] for i in range(mbs)])
        b_out = similarity_chunk(q, embed(wins).unsqueeze(0), temp).view(-1)
        take = min(num_valid, mbs)
        output[itr * mbs: itr * mbs + take] = b_out[:take]
        num_valid -= mbs
    return ids, output


def similarity_chunk(q: torch.Tensor, t: torch.Tensor, temp: float) -> torch.Tensor:
    """models.py:351-352,412-417 for one replica: q [B,D], t [B,T,D] -> [B,T]."""
    q = F.normalize(q, dim=1).unsqueeze(1)
    t = F.normalize(t, dim=2).permute(0, 2, 1)
    output = torch.bmm(q, t).squeeze(1)
    output /= temp
    return output


def audio_chunk(d_a: torch.Tensor, s_a: torch.Tensor, temp: float) -> torch.Tensor:
    """models.py:433-439,457 (same math in the raw-log-mel branch :445-455): d_a [B,A],
    s_a [B,T,A] -> output_a [B,1,T]."""
    s_a = F.normalize(s_a, dim=2).permute(0, 2, 1)
    d_a = F.normalize(d_a, dim=1).unsqueeze(1)
    output_a = torch.bmm(d_a, s_a)
    output_a /= temp
    return output_a


def step_scores(q_emb, t_emb, ids, temp, mbs, num_gpus=1, d_a=None, s_a=None):
    """Scores of one synthesis step in the reference's chunk layout.

    Targets are taken in `ids` order, split into chunks of `mbs` (zero padded, utils.py:208-230),
    `num_gpus` chunks per model call (validate.py:442-445), each replica scoring its chunk
    with `similarity_chunk`; results land at itr*G*mbs with the num_valid tail
    (validate.py:481-493, 522).  Returns (output, output_a | None), CPU fp32, length len(ids).
    """
    n = len(ids)
    T = t_emb[ids]
    n_chunks = math.ceil(n / mbs)
    padded = torch.zeros((n_chunks, mbs, T.shape[1]), dtype=T.dtype)
    padded.view(-1, T.shape[1])[:n] = T
    output = torch.zeros(n, dtype=torch.float32)
    output_a = None
    if d_a is not None:
        S = s_a[ids]
        padded_a = torch.zeros((n_chunks, mbs, S.shape[1]), dtype=S.dtype)
        padded_a.view(-1, S.shape[1])[:n] = S
        output_a = torch.zeros(n, dtype=torch.float32)
    num_valid = n
    for itr in range(math.ceil(n_chunks / num_gpus)):
        b_t = padded[itr * num_gpus: itr * num_gpus + num_gpus]
        B = b_t.shape[0]
        b_q = q_emb.unsqueeze(0).repeat(B, 1)
        b_out = similarity_chunk(b_q, b_t, temp)
        lo = itr * num_gpus * mbs
        take = min(num_valid, num_gpus * mbs)
        output[lo: lo + take] = b_out.contiguous().view(-1)[:take]
        if d_a is not None:
            b_s = padded_a[itr * num_gpus: itr * num_gpus + num_gpus]
            b_out_a = audio_chunk(d_a.unsqueeze(0).repeat(B, 1), b_s, temp)
            output_a[lo: lo + take] = b_out_a.view(-1)[:take]
        num_valid -= mbs * num_gpus
    return output, output_a


def mix_and_select(output, output_a, alpha, threshold):
    """validate.py:524-527, 554, 558, 568.  Returns (output after renorm, choices, pre-threshold)."""
    output = output.clone()
    output /= output.sum()
    if output_a is not None:
        output_a = output_a.clone()
        output_a /= output_a.sum()
        output = alpha * output + (1 - alpha) * output_a
    mixed = output.clone()
    output[output < (output.max() - threshold * output.max())] = 0.0
    output[torch.nonzero(output).view(-1)] /= output.sum()
    choices = output.nonzero().view(-1)
    return output, choices, mixed


def select_margin(mixed: torch.Tensor, threshold: float) -> float:
    """Relative distance of the closest element to the threshold cut (fixture screening)."""
    cut = mixed.max() - threshold * mixed.max()
    return float(((mixed - cut).abs() / cut.abs()).min())


def start_segment(audio_eg: torch.Tensor, driving_eg0: torch.Tensor) -> int:
    """validate.py:222-240 — arg-max cosine similarity with strict '>' (first max wins, init 0)."""
    q_id = 0
    max_sim = 0
    driving_eg = driving_eg0
    for choice in range(audio_eg.shape[0]):
        source_eg = F.normalize(audio_eg[choice].view(-1), dim=0)
        driving_eg = F.normalize(driving_eg.view(-1), dim=0)
        sim = torch.nn.CosineSimilarity(dim=0)(source_eg, driving_eg)
        if sim > max_sim:
            q_id = copy.deepcopy(choice)
            max_sim = max(sim, max_sim)
    return int(q_id)


def synthesize(t_emb, temp, threshold, mbs, fps, new_video_length, window, stride,
               q_start=10, alpha=0.5, num_gpus=1, q_audio=None, da_source=None, da_driving=None,
               subsample_rate=1, return_debug=False):
    """The `while len(new_frames) < max_length` loop of validate.py:324-690 restricted to the
    hot-path lines, at the embedding boundary.

    t_emb      [L, D]   per-window video embeddings (query and target encoders share them here)
    q_audio    [La, A]  model-side audio embeddings (model_type 2: concatenated before the
                        normalisation, models.py:347,408); index clamped to La-1 (validate.py:346,400)
    da_source  [La, A'] driving-audio-model features of the source windows (models.py:425-427)
    da_driving [>=steps+1, A'] driving audio features; step uses row `iter_count` (validate.py:417)
    Consumes the numpy global RNG once per step.  Returns dict(q_ids, frame_ids, jump_count, ...).
    """
    L = t_emb.shape[0]
    W, S = window, stride
    max_length = math.ceil(fps) * new_video_length
    if da_driving is not None:                                      # validate.py:260-263
        max_length = min(max_length, np.ceil(fps) * np.floor(len(da_driving) * S + W))
    if q_audio is not None:
        max_a = q_audio.shape[0] - 1
        a_idx = [min(i, max_a) for i in range(L)]
        emb = torch.cat((t_emb, q_audio[a_idx]), dim=1)
    else:
        emb = t_emb
    s_a = None
    if da_driving is not None:
        max_a = da_source.shape[0] - 1
        s_a = da_source[[min(i, max_a) for i in range(L)]]
    q_id = q_start
    p_q_id = -1
    iter_count = 1
    n_frames = 0
    q_ids, frame_ids, nz_counts, margins = [], [], [], []
    jump_count = 0
    while n_frames < max_length:
        ids = target_order(q_id, L)
        d_a = da_driving[iter_count] if da_driving is not None else None
        output, output_a = step_scores(emb[q_id], emb, ids, temp, mbs, num_gpus, d_a, s_a)
        output, choices, mixed = mix_and_select(output, output_a, alpha, threshold)
        margins.append(select_margin(mixed, threshold))
        nz_counts.append(len(choices))
        rdm_id = np.random.choice(choices.numpy())
        q_id = int(ids[rdm_id])
        if p_q_id == -1:
            diff = list(range(q_id * S, q_id * S + W))
        else:
            if q_id != p_q_id + 1:
                jump_count += 1
            diff = list(range(q_id * S + (W - S), q_id * S + W))
        frame_ids.extend(diff)
        n_frames += len(diff) * subsample_rate
        q_ids.append(q_id)
        iter_count += 1
        p_q_id = q_id
    res = dict(q_ids=q_ids, frame_ids=frame_ids, jump_count=jump_count, nz_counts=nz_counts)
    if return_debug:
        res["margins"] = margins
    return res

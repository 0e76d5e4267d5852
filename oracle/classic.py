"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (torch-CPU + numpy, the same arithmetic libraries the reference
uses) of the classic transition-matrix path of medhini/audio-video-textures:

    compute_D1   baselines/classic_video_textures/computeD1.py:47-96, 240-247
    compute_D2   baselines/classic_video_textures/computeD2.py:21-52
    q_learning   baselines/classic_video_textures/q_learning.py:27-68
    walk         baselines/classic_video_textures/video_textures.py:32-211

Parity status: the reference has no tests or golden vectors ("parity unpinned" by
the reference itself).  This restatement is pinned instead against the UNMODIFIED
reference functions imported in the build container (oracle/ref_shim.py,
oracle/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import copy

import numpy as np
import torch


# --------------------------------------------------------------------------- tail
def sigma_and_probs(D: torch.Tensor, sigma_factor):
    """computeD1.py:240-245 == computeD2.py:44-50 == q_learning.py:53-59.

    nnz = #nonzero; sigma = f * (sum(D) / nnz); E = exp(-D / sigma);
    P[i] = E[i+1] (last row duplicated); P /= rowsum.
    """
    non_zero_count = torch.nonzero(D).size(0)
    sigma = sigma_factor * (D.sum() / non_zero_count)
    P = torch.exp(-D / sigma)
    P = torch.cat((P[1:, :], P[-1, :].unsqueeze(0)), dim=0)
    P = P / P.sum(1, keepdim=True)
    return P, sigma


# --------------------------------------------------------------------------- D1
def pairwise_l2(frames: torch.Tensor, block: int = 64) -> torch.Tensor:
    """D1[i,j] = ||x_i - x_j||_2 by direct difference (computeD1.py:64-94).

    Same per-pair op as the reference (`torch.norm(A - B, dim=2)` over the flattened
    pixels, fp32), re-blocked so it does not need the reference's two bs*bs*K `repeat`
    temporaries.  The per-pair value does not depend on the blocking.
    """
    x = frames.reshape(frames.shape[0], -1).to(torch.float32)
    n = x.shape[0]
    D1 = torch.empty((n, n), dtype=torch.float32)
    for i in range(0, n, block):
        a = x[i:i + block]
        for j in range(0, n, block):
            b = x[j:j + block]
            D1[i:i + block, j:j + block] = torch.norm(a.unsqueeze(1) - b.unsqueeze(0), dim=2)
    return D1


def pairwise_l2_reference_blocks(frames: torch.Tensor, batch_size: int, max_blocks=None):
    """Literal block algorithm of computeD1.py:58-96 (repeat -> view -> norm -> permute).

    Used (a) on small inputs to show the re-blocked `pairwise_l2` is identical and
    (b) by bench.py's CPU baseline, on a bounded number of blocks (`max_blocks`).
    Returns (D1, blocks_done); untouched entries stay at the reference's init value 1.
    """
    n = len(frames)
    D1 = torch.ones((n, n))
    done = 0
    for i in range(0, n, batch_size):
        frames_batch_A = frames[i:min(i + batch_size, n)]
        for j in range(0, n, batch_size):
            bj = min(batch_size, n - j)
            bi = min(batch_size, n - i)
            feats_A = frames_batch_A.unsqueeze(0).repeat(bj, 1, 1, 1, 1).view(bj, bi, -1)
            frames_batch_B = frames[j:min(j + batch_size, n)]
            feats_B = frames_batch_B.unsqueeze(1).repeat(1, bi, 1, 1, 1).view(bj, bi, -1)
            D_mini = torch.norm(feats_A - feats_B, dim=2)
            D1[i:i + batch_size, j:j + batch_size] = D_mini.permute(1, 0)
            done += 1
            if max_blocks is not None and done >= max_blocks:
                return D1, done
    return D1, done


def pairwise_l2_exact_u8(video_u8: torch.Tensor) -> torch.Tensor:
    """Exact integer d^2 for uint8 frames, sqrt in fp64, rounded to fp32.

    Not a reference restatement: an error-free yardstick used by tests to show that
    both the reference's fp32 accumulation and the GPU s8 tensor-core path sit within
    the stated tolerance of the true value.
    """
    x = video_u8.reshape(video_u8.shape[0], -1).to(torch.float64)
    n2 = (x * x).sum(1)
    d2 = n2[:, None] + n2[None, :] - 2.0 * (x @ x.T)          # exact: integers < 2^53
    return d2.clamp_min_(0).sqrt_().to(torch.float32)


def compute_D1(frames: torch.Tensor, sigma_factor, feats: str = "RGB", audio=None, sr: int = 0,
               fps: int = 30, slow: bool = True, batch_size: int = 128):
    """computeD1.py:27-36,47-96,240-247 (RGB branch only; other `feats` are feature producers)."""
    if feats != "RGB":
        raise NotImplementedError("oracle covers the RGB branch (SURVEY.md §2.1 row 1)")
    D1 = pairwise_l2(frames)
    P1, sigma = sigma_and_probs(D1, sigma_factor)
    return D1, P1, sigma


def feature_mode_D1(image_feats: torch.Tensor, feats: str = "ResNet", audio_feats=None, fps: int = 30,
                    slow: bool = False, batch_size: int = 128) -> torch.Tensor:
    """D1 of the non-RGB modes given the producers' outputs (the ResNet-18 / VGGish networks themselves are
    feature producers): computeD1.py:105-116 (ResNet dense), :117-148 (ResNet tiled), :155-192 (ResNet_VGGish
    dense), :194-236 (ResNet_VGGish tiled).  Reproduces the tiled loops' skipped blocks (`range(0, N - bs, bs)`,
    `continue` on ragged blocks) and their initial values (ones / zeros), and the `.repeat(fps, 1)` tiling of the
    per-second audio features.  The tiled ResNet branch re-normalises A on every column block (idempotent up to
    an ulp); it is normalised once here."""
    import torch.nn.functional as F
    x = image_feats
    if feats == "ResNet_VGGish":
        n = int(len(x) / fps) * fps
        a = audio_feats[: int(len(image_feats) / fps)].repeat(fps, 1)
        x = torch.cat((x[:n], a), dim=1)
    n = len(x)
    normalise = (feats == "ResNet") or not slow
    if not slow:
        A = x.unsqueeze(0).repeat(n, 1, 1)
        B = x.unsqueeze(1).repeat(1, n, 1)
        if normalise:
            A, B = F.normalize(A, dim=2), F.normalize(B, dim=2)
        return torch.norm(A - B, dim=2)
    D1 = torch.ones((n, n)) if feats == "ResNet" else torch.zeros((n, n))
    for i in range(0, n, batch_size):
        fa = x[i:i + batch_size].unsqueeze(0).repeat(batch_size, 1, 1)
        for j in range(0, n - batch_size, batch_size):
            fb = x[j:j + batch_size].unsqueeze(1).repeat(1, batch_size, 1)
            if fa.shape != fb.shape:
                continue
            a_, b_ = (F.normalize(fa, dim=2), F.normalize(fb, dim=2)) if normalise else (fa, fb)
            D1[i:i + batch_size, j:j + batch_size] = torch.norm(a_ - b_, dim=2).permute(1, 0)
    return D1


# --------------------------------------------------------------------------- D2
def binomial_weights(filter_size: int) -> torch.Tensor:
    """computeD2.py:34 — coeffs((0.5x+0.5)^(fs-1)) in float64, cast to fp32."""
    return torch.tensor((np.poly1d([0.5, 0.5]) ** (filter_size - 1)).coeffs, dtype=torch.float32)


def compute_D2(D1: torch.Tensor, sigma_factor, filter_size: int = 16, stride: int = 1):
    """computeD2.py:21-52 — conv2d with diag(binomial) kernel, 'valid', optional stride."""
    binomial_filter = torch.tensor(
        np.diag((np.poly1d([0.5, 0.5]) ** (filter_size - 1)).coeffs), dtype=torch.float32)
    D2 = D1.view(1, 1, D1.shape[0], D1.shape[0])
    binomial_filter = binomial_filter.view(1, 1, filter_size, filter_size)
    D2 = torch.nn.functional.conv2d(D2, binomial_filter, stride=stride)
    D2 = D2.view(D2.shape[2], D2.shape[3])
    P2, sigma = sigma_and_probs(D2, sigma_factor)
    return D2, P2, sigma, binomial_filter


def diag_filter_sequential(D1: torch.Tensor, filter_size: int, stride: int = 1) -> torch.Tensor:
    """Same filter as compute_D2 written as an explicit k-ordered fp32 sum (the order the
    CUDA kernel uses); differs from conv2d only by summation order (<= few ulp)."""
    w = binomial_weights(filter_size)
    n = D1.shape[0]
    m = (n - filter_size) // stride + 1
    out = torch.zeros((m, m), dtype=torch.float32)
    span = (m - 1) * stride + 1
    for k in range(filter_size):
        out += w[k] * D1[k:k + span:stride, k:k + span:stride]
    return out


# --------------------------------------------------------------------------- future cost
def row_min_offdiag(X: torch.Tensor) -> torch.Tensor:
    """q_learning.py:43-46 — min over k != j of X[j,k] (exact, order-free)."""
    Y = X.clone()
    Y.fill_diagonal_(float("inf"))
    return Y.min(dim=1)[0]


def future_cost(D3: torch.Tensor, alpha: float = 0.997, faithful: bool = False, verbose=False):
    """q_learning.py:36-51.  Returns (D3_new, eps_trail).

    faithful=True runs the literal O(M^3) loop (mask rebuilt and `mins` recomputed for every
    row); the default computes `mins` once per sweep, which is the same arithmetic because
    `mins` only reads D3_old.
    """
    eps = 10000
    D3_new = copy.deepcopy(D3)
    trail = []
    while eps > 10e-3:
        D3_old = copy.deepcopy(D3_new)
        if faithful:
            for i in range(D3.shape[0] - 1, 0, -1):
                mask = np.ones((D3.shape[0], D3.shape[1]), dtype=bool)
                np.fill_diagonal(mask, False)
                mins = D3_old[mask, ...].view(D3.shape[0], -1).min(axis=1)[0]
                D3_new[i] = D3[i] + alpha * mins
        else:
            mins = row_min_offdiag(D3_old)
            D3_new[1:] = D3[1:] + alpha * mins
        eps = ((D3_new - D3_old) ** 2).mean()
        trail.append(float(eps))
        if verbose:
            print("Eps:", eps)
    return D3_new, trail


def threshold_rows(P3: torch.Tensor, thresholding: float) -> torch.Tensor:
    """q_learning.py:61-64 — zero entries below max - th*max per row; NOT renormalised."""
    P3_new = copy.deepcopy(P3)
    for i in range(len(P3_new)):
        P3_new[i][P3_new[i] < (P3_new[i].max() - thresholding * P3_new[i].max())] = 0.0
    return P3_new


def q_learning(D2: torch.Tensor, sigma_factor, p: float = 0.7, alpha: float = 0.997,
               thresholding: float = 0.75, faithful: bool = False, return_trail: bool = False):
    """q_learning.py:27-68."""
    D3 = D2 ** p
    D3_new, trail = future_cost(D3, alpha, faithful=faithful)
    P3, sigma = sigma_and_probs(D3_new, sigma_factor)
    P3_new = threshold_rows(P3, thresholding)
    if return_trail:
        return D3_new, P3, P3_new, sigma, trail
    return D3_new, P3, P3_new, sigma


def threshold_margin(P3: torch.Tensor, thresholding: float, rows=None) -> float:
    """Smallest relative distance of any element to its row's cut.

    Fixtures whose margin is below ~10x the value tolerance cannot be expected to give
    bit-identical survivor sets across CPU/GPU `exp` implementations.
    """
    mx = P3.max(dim=1, keepdim=True)[0]
    cut = mx - thresholding * mx
    rel = ((P3 - cut).abs() / cut)
    if rows is not None:
        rel = rel[rows]
    return float(rel.min())


# --------------------------------------------------------------------------- audio prior
def compute_Paudio(t_audio_eg: torch.Tensor, driving_audio: torch.Tensor) -> torch.Tensor:
    """computePaudio.py:6-18."""
    import torch.nn.functional as F
    s_a = F.normalize(t_audio_eg, dim=1)
    d_a = F.normalize(driving_audio, dim=0).unsqueeze(0)
    cos = torch.nn.CosineSimilarity(dim=1)
    p_audio = cos(d_a.repeat([s_a.shape[0], 1]), s_a)
    return p_audio / (p_audio.sum() + 1e-6)


# --------------------------------------------------------------------------- walk
def walk(P, model_type: int, fps: int, new_video_length: int, stride: int, filter_size: int,
         start: int = 100):
    """Sampling walk of video_textures.py:43-46 and
    m1 :48-51,73-81,103,120 ; m2 :131-133,146-158,169 ; m3 :170-172,185-197,209.

    Consumes numpy's GLOBAL legacy RNG exactly like the reference (one np.random.choice
    per draw).  Returns (new_frames_list, jump_count).  Reproduces the mode-3 defect:
    `this_frame` is only advanced after the loop.
    """
    P = torch.as_tensor(P)
    target_len = fps * new_video_length
    jump_count = 0
    if model_type == 1:
        this_frame = start
        out = [start]
        while len(out) < target_len:
            next_frame = np.random.choice(P[this_frame].nonzero().view(-1).numpy())
            if next_frame != this_frame + 1:
                jump_count += 1
            out.append(next_frame)
            this_frame = copy.deepcopy(next_frame)
    elif model_type == 2:
        this_frame = start
        out = list(np.arange(this_frame, this_frame + stride))
        this_frame += stride
        while len(out) < target_len:
            next_frame = np.random.choice(P[this_frame].nonzero().view(-1).numpy())
            if next_frame != this_frame + 1:
                jump_count += 1
            out.extend(list(np.arange(next_frame, min(next_frame + stride, P.shape[0]))))
            this_frame = min(next_frame + stride, P.shape[0] - 1)
    else:
        this_frame = start
        out = list(np.arange(this_frame, this_frame + filter_size))
        while len(out) < target_len:
            next_frame = np.random.choice(P[this_frame].nonzero().view(-1).numpy())
            if next_frame != this_frame + 1:
                jump_count += 1
            out.extend(list(np.arange(this_frame * stride + (filter_size - stride),
                                      this_frame * stride + filter_size)))
        this_frame = next_frame  # noqa: F841  (reference: outside the while)
    return [int(v) for v in out], jump_count


# --------------------------------------------------------------------------- frame assembly
def assemble_frames(frames, new_frames_list, half: int = 4, floor_div: bool = True):
    """video_textures.py:215-221 (classic: half-width 4, `//`) / validate.py:622-631 (contrastive: half-width 3,
    `/`): the chosen frames with the black progress bar and the red position marker painted in."""
    frames = np.asarray(frames)
    out = []
    for frame_idx in new_frames_list:
        frames_bar = np.zeros((15, frames.shape[-2], 3))
        frame_n = int(frame_idx * frames.shape[-2] // len(frames)) if floor_div else int(frame_idx * frames.shape[-2] / len(frames))
        frames_bar[:, frame_n - half: frame_n + half, :] = [255, 0, 0]
        frame_arr = np.array(frames[frame_idx])
        frame_arr[-25:-10, :, :] = frames_bar
        out.append(frame_arr)
    return np.stack(out)

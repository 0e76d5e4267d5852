"""Drop-in for baselines/classic_video_textures/computeD1.py (RGB branch, :27-96 and the tail :240-247)."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine


def _to_device(frames: torch.Tensor) -> torch.Tensor:
    if frames.is_cuda:
        return frames
    if not torch.cuda.is_available():
        raise RuntimeError("audio_video_textures_b200 needs a CUDA device (B200); there is no CPU path")
    return frames.cuda(non_blocking=True)


def tail(D: torch.Tensor, sigma_factor, stats: torch.Tensor | None = None, threshold=None, want_P=True):
    """sigma / P tail shared by computeD1.py:240-245, computeD2.py:44-50, q_learning.py:53-59.
    Returns (P, P_new, sigma 0-dim CUDA fp32 tensor, counts)."""
    if stats is None:
        stats = engine.sum_nnz(D)
    total, nnz = engine.read_stats(stats)
    sigma = engine.sigma_from_stats(total, nnz, sigma_factor)
    P, P_new, counts = engine.transition_probs(D, sigma, shift=1, threshold=threshold, want_P=want_P,
                                               want_counts=threshold is not None)
    return P, P_new, torch.tensor(sigma, dtype=torch.float32, device=D.device), counts


# Feature producers of the non-RGB modes (classic/computeD1.py:98-103, 151-167): an ImageNet ResNet-18 trunk
# (torchvision, pretrained) and torch.hub's VGGish.  Both need downloaded weights, which is why they are injectable:
# set these to callables (frames [N,3,H,W] -> [N, C] CUDA fp32;  (audio, sr) -> [seconds, A]) or leave None to let
# compute_D1 try the reference's own loaders.  The distance arithmetic after them is on libavtex kernels.
IMAGE_FEATURES = None
AUDIO_FEATURES = None


def _image_features(frames: torch.Tensor, batch_size: int) -> torch.Tensor:
    fn = IMAGE_FEATURES
    if fn is None:
        try:
            import torchvision.models as models
            net = torch.nn.Sequential(*list(models.resnet18(pretrained=True).children())[:-1]).cuda().eval()
        except Exception as exc:
            raise RuntimeError("feats='ResNet*' needs the pretrained torchvision ResNet-18 (not available here: "
                               f"{type(exc).__name__}); set classic.computeD1.IMAGE_FEATURES to a feature callable") from exc
        fn = lambda x: net(x).view(x.shape[0], -1)                 # noqa: E731
    out = []
    with torch.no_grad():
        for i in range(0, len(frames), max(1, batch_size)):
            part = _to_device(frames[i:i + batch_size]).float()
            out.append(fn(part).reshape(part.shape[0], -1))
    return torch.cat(out, dim=0).float().contiguous()


def _audio_features(audio, sr: int) -> torch.Tensor:
    fn = AUDIO_FEATURES
    if fn is None:
        try:
            vggish = torch.hub.load("harritaylor/torchvggish", "vggish")
            vggish.eval()
            fn = vggish.forward
        except Exception as exc:
            raise RuntimeError("feats='ResNet_VGGish' needs torch.hub's VGGish (not available here: "
                               f"{type(exc).__name__}); set classic.computeD1.AUDIO_FEATURES to a feature callable") from exc
    with torch.no_grad():
        return torch.as_tensor(fn(audio, sr)).float().cuda()


def _feature_distances(feats: torch.Tensor, normalise: bool, slow: bool, batch_size: int, fill: float,
                       stats: torch.Tensor) -> torch.Tensor:
    """D1 of the feature modes.  Dense (`slow=False`, computeD1.py:105-116 / 174-192): all pairs.  Tiled
    (`slow=True`, :117-148 / 194-236): the reference only visits FULL bs x bs blocks with column start
    j < N - bs and skips ragged row blocks (`continue` on a shape mismatch), leaving its initial value
    (`fill`: ones for ResNet, zeros for ResNet_VGGish) everywhere else — reproduced, not fixed.  One thing is NOT
    reproducible by anyone: the tiled ResNet loop re-normalises block A on every column block but B only once
    (:135-136), so the reference's diagonal is rounding noise (exact 0 or ~1e-8) and its `nonzero` count — hence
    sigma — moves by up to N entries of N^2 with the arithmetic library; here the diagonal is exactly 0."""
    x = engine.l2_normalize_rows(feats) if normalise else feats.contiguous()
    n = x.shape[0]
    D1 = engine.pairdist_direct(x)
    if slow:
        rows_done = (n // batch_size) * batch_size
        cols_done = batch_size * len(range(0, n - batch_size, batch_size))
        D1[rows_done:, :] = fill                                   # (plumbing: constant fill of the untouched blocks)
        D1[:, cols_done:] = fill
    engine.sum_nnz(D1, stats)
    return D1


def compute_D1(
    frames: torch.Tensor,
    sigma_factor: float,
    feats: str = "L2",
    audio: np.ndarray = None,
    sr: int = 0,
    fps: int = 30,
    slow: bool = True,
    batch_size: int = 128,
):
    """Pairwise frame L2 distances, sigma1 and the shifted row-stochastic P1.

    frames: [N, H, W, C] (any trailing layout) float32 or uint8, CPU or CUDA.  In the RGB mode `slow` /
    `batch_size` are pure tiling knobs (computeD1.py:49,58-63): the per-pair value does not depend on them, so
    they are accepted and ignored.  feats == "ResNet" / "ResNet_VGGish" (:98-238): features come from the
    injectable producers above, the (normalised) feature distances from the direct fp32 kernel; there `slow`
    DOES change the result (the reference's tiled loops skip blocks) and is honoured.
    Returns (D1 [N,N], P1 [N,N], sigma) as CUDA fp32 tensors, like the reference.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("audio_video_textures_b200 needs a CUDA device (B200); there is no CPU path")
    stats = engine.new_stats(torch.device("cuda", torch.cuda.current_device()))
    if feats == "ResNet":                                          # computeD1.py:98-148
        image = _image_features(frames, batch_size)
        D1 = _feature_distances(image, True, slow, batch_size, 1.0, stats)
        P1, _, sigma, _ = tail(D1, sigma_factor, stats)
        return D1, P1, sigma
    if feats == "ResNet_VGGish":                                   # computeD1.py:150-236
        audio_feats = _audio_features(audio, sr)
        audio_feats = audio_feats[: int(len(frames) / fps)].repeat(fps, 1)     # :155-156 (tiles the block fps times)
        print("Shape of audio feats:", audio_feats.shape)
        frames = frames[: int(len(frames) / fps) * fps]                        # :160
        image = _image_features(frames, batch_size)
        if not slow:
            print("Shape of image feats:", image.shape)
        joint = torch.cat((image, audio_feats.to(image.device)), dim=1)        # plumbing: concatenation
        if not slow:
            print("Shape of joint feats:", joint.shape)
        D1 = _feature_distances(joint, not slow, slow, batch_size, 0.0, stats)  # tiled branch is NOT normalised (:225-226)
        P1, _, sigma, _ = tail(D1, sigma_factor, stats)
        return D1, P1, sigma
    if feats != "RGB":
        raise NotImplementedError(f"feats={feats!r}: the reference defines RGB, ResNet and ResNet_VGGish "
                                  "(anything else is an UnboundLocalError there)")
    streamed = engine.pairwise_l2_from_host(frames, stats=stats) if not frames.is_cuda else None
    if streamed is not None and streamed[1].exact_ok:      # uint8 host frames: Gram overlapped with the H2D copy
        D1 = streamed[0]
    else:
        x = _to_device(frames)
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        stats.zero_()
        D1, _ = engine.pairwise_l2(x, stats=stats)
    P1, _, sigma, _ = tail(D1, sigma_factor, stats)
    return D1, P1, sigma

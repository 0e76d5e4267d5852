"""Deterministic synthetic inputs for the transition-matrix path.

The reference ships no data, no seeds and no `read_data` (classic/video_textures.py:26
imports a module that does not exist), so every parity/bench number in this repo is
quoted on the generators below.  They are pure functions of (seed, shape).

CPU generators use ``torch.Generator().manual_seed(seed)`` and are bit-reproducible
across machines; the ``*_cuda`` variants use the device generator and are only used
for the large bench workloads where a CPU draw would take minutes.
"""
from __future__ import annotations

import math

import torch

PERIOD = 37.0
AMPLITUDE = 100.0
NOISE = 8.0


def synth_video(n_frames: int, height: int, width: int, seed: int = 0,
                period: float = PERIOD, noise: float = NOISE) -> torch.Tensor:
    """Smooth periodic uint8 video ``[N, H, W, 3]`` (CPU).

    x_t = clamp(128 + 100*sin(2*pi*(t/period + phase)) + noise*randn, 0, 255).round()
    Draw order on the generator: ``rand(H,W,3)`` (phase) then ``randn(N,H,W,3)``.
    """
    g = torch.Generator().manual_seed(seed)
    phase = torch.rand(height, width, 3, generator=g)
    eps = torch.randn(n_frames, height, width, 3, generator=g)
    t = torch.arange(n_frames, dtype=torch.float32).view(-1, 1, 1, 1)
    x = 128.0 + AMPLITUDE * torch.sin(2.0 * math.pi * (t / period + phase)) + noise * eps
    return x.clamp_(0.0, 255.0).round_().to(torch.uint8)


def synth_video_cuda(n_frames: int, height: int, width: int, seed: int = 0,
                     device: str | torch.device = "cuda", period: float = PERIOD,
                     noise: float = NOISE, chunk: int = 256) -> torch.Tensor:
    """Same distribution as :func:`synth_video`, generated on the device in chunks."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    phase = torch.rand(height, width, 3, generator=g, device=dev)
    out = torch.empty(n_frames, height, width, 3, dtype=torch.uint8, device=dev)
    for s in range(0, n_frames, chunk):
        e = min(s + chunk, n_frames)
        t = torch.arange(s, e, dtype=torch.float32, device=dev).view(-1, 1, 1, 1)
        x = 128.0 + AMPLITUDE * torch.sin(2.0 * math.pi * (t / period + phase))
        x += noise * torch.randn(e - s, height, width, 3, generator=g, device=dev)
        out[s:e] = x.clamp_(0.0, 255.0).round_().to(torch.uint8)
    return out


def _smooth_walk(n: int, dim: int, g: torch.Generator, device=None, decay: float = 0.9):
    """Low-pass random walk over the window index: neighbours are similar."""
    steps = torch.randn(n, dim, generator=g, device=device)
    walk = torch.empty_like(steps)
    acc = torch.zeros(dim, device=device)
    scale = math.sqrt(1.0 - decay * decay)
    for i in range(n):
        acc = decay * acc + scale * steps[i]
        walk[i] = acc
    return walk


def synth_embeddings(n_windows: int, dim: int = 2304, seed: int = 0, walk_gain: float = 3.0,
                     device=None) -> torch.Tensor:
    """SlowFast-shape per-window embeddings ``[L, D]`` fp32 (not normalised)."""
    dev = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.randn(n_windows, dim, generator=g, device=dev)
    return base + walk_gain * _smooth_walk(n_windows, dim, g, dev)


def synth_audio_features(n_rows: int, dim: int, seed: int = 0, walk_gain: float = 2.0,
                         device=None) -> torch.Tensor:
    """VGGish-shape non-negative (post-ReLU) audio features ``[rows, A]`` fp32.

    A = 12288 is the reference-faithful conv-map size (cvt/models/audio_models/vggish.py:42-46),
    A = 128 the canonical VGGish embedding (cvt/utils/vggish_params.py:23).
    """
    dev = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=dev).manual_seed(seed + 7919)
    base = torch.randn(n_rows, dim, generator=g, device=dev)
    return torch.relu(base + walk_gain * _smooth_walk(n_rows, dim, g, dev))

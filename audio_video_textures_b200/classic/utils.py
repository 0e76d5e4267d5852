"""Shim for the module the reference imports but does not ship
(`from utils import Logger, read_data, save_video`, classic/video_textures.py:26)."""
from __future__ import annotations

import os

import numpy as np
import torch

from ..synth import synth_video


class Logger:
    """No-op stand-in for the tensorboardX logger (figures are outside the hot path)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def read_data(args, video_name: str):
    """Returns (input_frames, video, fps, audio, sr, extra) like the missing reference function.

    Sources: `--synthetic N,H,W[,seed]` (deterministic clip of synth.py), a `<vdata>/<name>.pt|.npy`
    uint8 tensor [N,H,W,3], or (when torchvision can decode it) `<vdata>/<name>.mp4`.
    `input_frames` is the uint8 video itself: the distance kernels consume bytes directly, which is
    numerically identical to the `video.float()` a float reader would hand to compute_D1.
    """
    spec = getattr(args, "synthetic", None)
    if spec or args.vdata is None:
        parts = [int(v) for v in (spec or "300,64,64,0").split(",")]
        n, h, w = parts[:3]
        seed = parts[3] if len(parts) > 3 else 0
        video = synth_video(n, h, w, seed=seed)
        return video, video, args.fps, None, args.sr, None
    base = os.path.join(args.vdata, video_name)
    if os.path.exists(base + ".pt"):
        video = torch.load(base + ".pt")
    elif os.path.exists(base + ".npy"):
        video = torch.from_numpy(np.load(base + ".npy"))
    else:
        import torchvision.io as io
        video, _, info = io.read_video(base + ".mp4", pts_unit="sec")
        args.fps = int(round(info.get("video_fps", args.fps)))
    video = video.to(torch.uint8)
    return video, video, args.fps, None, args.sr, None


def write_frames(frames, frame_ids, folder: str):
    from PIL import Image
    os.makedirs(folder, exist_ok=True)
    for count, idx in enumerate(frame_ids):
        Image.fromarray(np.asarray(frames[idx])).save(os.path.join(folder, "{:04d}.png".format(count + 1)))


def save_video(*a, **k):
    """ffmpeg muxing is I/O glue outside the hot path."""
    return None

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


def marker_columns(frame_ids, width: int, n_source: int, half: int = 4, floor_div: bool = True):
    """Column range of the red position marker for every output frame, with the reference's slice semantics:
    classic (video_textures.py:217-218): frame_n = int(idx * W // N), columns [frame_n-4, frame_n+4);
    contrastive (validate.py:629-630): frame_n = int(idx * W / N), half-width 3.  A negative slice start wraps
    around in numpy and yields an EMPTY slice there (no marker for the first frames) — reproduced."""
    lo, hi = [], []
    for idx in frame_ids:
        fn = int(idx * width // n_source) if floor_div else int(idx * width / n_source)
        a, b = fn - half, fn + half
        if a < 0:
            a += width                                              # numpy: negative start counts from the end
        b = min(b, width)
        lo.append(a)
        hi.append(b if b > a else a)
    return np.asarray(lo, dtype=np.int32), np.asarray(hi, dtype=np.int32)


def assemble_frames(video, frame_ids, bar: bool = True, half: int = 4, floor_div: bool = True) -> torch.Tensor:
    """(f4) The output clip [n_out, H, W, 3] uint8 for the chosen frame ids, gathered (and the progress bar painted)
    on the device in one launch instead of one host copy + PIL image per frame (video_textures.py:215-226,
    validate.py:622-634).  `video`: uint8 [N, H, W, 3], CPU or CUDA."""
    import ctypes as C

    from .. import _lib, engine
    v = video if video.is_cuda else video.cuda()
    v = v.contiguous()
    n, h, w, _ = v.shape
    ids = torch.as_tensor(np.asarray(frame_ids, dtype=np.int32)).to(v.device)
    lo, hi = marker_columns(frame_ids, w, n, half, floor_div)
    d_lo, d_hi = torch.from_numpy(lo).to(v.device), torch.from_numpy(hi).to(v.device)
    out = torch.empty((len(ids), h, w, 3), dtype=torch.uint8, device=v.device)
    _lib.call("avtex_assemble_frames", _lib.ptr(v), n, h, w, _lib.ptr(ids), _lib.ptr(d_lo), _lib.ptr(d_hi), 1 if bar else 0,
              len(ids), _lib.ptr(out), engine._dev(v), engine._stream(v))
    return out


def write_frames(frames, frame_ids, folder: str, bar: bool = True):
    """PNG files 0001.png ... of the chosen frames (video_textures.py:212-226), assembled on the GPU, one D2H copy."""
    from PIL import Image
    os.makedirs(folder, exist_ok=True)
    clip = assemble_frames(frames, frame_ids, bar=bar).cpu().numpy()
    for count in range(len(frame_ids)):
        Image.fromarray(clip[count]).save(os.path.join(folder, "{:04d}.png".format(count + 1)))


def save_video(*a, **k):
    """ffmpeg muxing is I/O glue outside the hot path."""
    return None

"""Argument surface of contrastive_video_textures/main.py:41-296 for the `-e` synthesis mode and its
dispatch (:429-438, :508-518): same flags, defaults and meanings for everything the hot path reads.
Training, checkpoints and encoders are outside the hot path; embeddings are loaded or synthesised.

    python -m audio_video_textures_b200.contrastive.main -e -th 0.3 -temp 0.1 -mbs 100 \
        --embeddings emb.pt            # [L, D] per-window embeddings
    ... -m 2 -da drive -alpha 0.5 --audio_embeddings a.pt --driving_embeddings d.pt
"""
from __future__ import annotations

import argparse
import math

import numpy as np
import torch

from ..synth import synth_audio_features, synth_embeddings
from .validate import validate


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="PyTorch Video Textures")
    p.add_argument("--enc_arch", "-ea", metavar="ARCH", default="resnet18", help="model architecture")
    p.add_argument("--model_type", "-m", default=1, type=int, help="(1) Video Textures (2) Audio Video Textures")
    p.add_argument("--vdata", "-vdata", default=None, type=str, help="Path to video dataset")
    p.add_argument("--adata", "-adata", default=None, type=str, help="Path to audio")
    p.add_argument("--dadata", "-dadata", default="audio/target", type=str, help="Path to driving audio")
    p.add_argument("--video_list", "-vl", default=None, type=str, nargs="+", help="list of input videos")
    p.add_argument("--fps", "-fps", default=30, type=int, help="frame rate of input video")
    p.add_argument("--subsample_rate", "-ssr", default=1, type=int)
    p.add_argument("--temp", "-temp", default=0.1, type=float, help="Temperature value")
    p.add_argument("--threshold", "-th", default=0.0, type=float, help="Threshold value")
    p.add_argument("--interpolation", "-nintp", default=True, action="store_false")
    p.add_argument("--window", "-w", default=20, type=int, help="Size of temporal window")
    p.add_argument("--stride", "-stride", default=4, type=int, help="Stride length")
    p.add_argument("--new_video_length", "-nvl", default=30, type=int, help="Length of new video")
    p.add_argument("--alpha", "-alpha", default=0.5, type=float)
    p.add_argument("--SF", "-SF", default=5, type=int)
    p.add_argument("--batch_size", "-bs", default=32, type=int)
    p.add_argument("--mini_batchsize", "-mbs", default=150, type=int)
    p.add_argument("--evaluate", "-e", dest="evaluate", action="store_true")
    p.add_argument("--driving_audio", "-da", default=None, type=str)
    p.add_argument("--da_feats", "-daf", default="VGG", type=str)
    p.add_argument("--daf_resume", "-daf_resume", default="", type=str)
    p.add_argument("--results_folder", "-rf", default="results", type=str)
    p.add_argument("--logdir", default="./logs")
    p.add_argument("--logname", default="exp")
    # embedding-boundary inputs (additions; synthetic tables when omitted)
    p.add_argument("--embeddings", default=None, type=str, help=".pt/.npy [L, D] per-window embeddings")
    p.add_argument("--audio_embeddings", default=None, type=str, help=".pt/.npy [La, A]")
    p.add_argument("--driving_embeddings", default=None, type=str, help=".pt/.npy [steps+1, A]")
    p.add_argument("--synthetic", default="20000,2304,0", type=str, help="L,D[,seed] when no --embeddings")
    p.add_argument("--seed", default=None, type=int, help="np.random.seed before the loop")
    return p


def _load(path):
    if path.endswith(".npy"):
        return torch.from_numpy(np.load(path))
    return torch.load(path)


def main(args):
    if not args.evaluate:
        raise SystemExit("only the -e synthesis mode is on the hot path (training is out of scope)")
    # main.py:515-516 — window / stride follow the frame rate
    args.window = math.ceil(args.fps / 2)
    args.stride = math.ceil(args.fps / 5)
    if args.embeddings:
        emb = _load(args.embeddings)
    else:
        parts = [int(v) for v in args.synthetic.split(",")]
        emb = synth_embeddings(parts[0], parts[1], seed=parts[2] if len(parts) > 2 else 0, device="cuda")
    q_audio = da_source = da_driving = None
    if args.model_type == 2 or args.driving_audio is not None:
        a = _load(args.audio_embeddings) if args.audio_embeddings else synth_audio_features(emb.shape[0], 128, device="cuda")
        if args.model_type == 2:
            q_audio = a
        if args.driving_audio is not None:
            da_source = a
            steps = math.ceil(args.fps) * args.new_video_length // args.stride + 4
            da_driving = _load(args.driving_embeddings) if args.driving_embeddings else \
                synth_audio_features(steps, a.shape[1], seed=1, device="cuda")
    if args.seed is not None:
        np.random.seed(args.seed)
    res = validate(args, emb.cuda(), q_audio=q_audio, da_source=da_source, da_driving=da_driving)
    print("Start:", res["start"])
    print("Chosen windows:", res["q_ids"])
    print("jump_count:", res["jump_count"])
    return res


if __name__ == "__main__":
    main(build_parser().parse_args())

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
    """EVERY flag of contrastive_video_textures/main.py:41-296, same names, short forms, defaults and types, so a
    reference command line parses unchanged.  The synthesis path reads -m -temp -th -alpha -mbs -fps -subsample -w
    -stride -nvl -e -da; the training / checkpoint / data-loader / logging flags are accepted and ignored (that
    machinery is outside the hot path)."""
    p = argparse.ArgumentParser(description="PyTorch Video Textures")
    p.add_argument("--enc_arch", "-ea", metavar="ARCH", default="resnet18", help="model architecture")
    p.add_argument("--model_type", "-m", default=1, type=int, help="(1) Video Textures (2) Audio Video Textures")
    p.add_argument("--vdata", "-vdata", default=None, type=str, help="Path to video dataset")
    p.add_argument("--adata", "-adata", default=None, type=str, help="Path to audio")
    p.add_argument("--pdata", "-pdata", default=None, type=str, help="Path to poses")
    p.add_argument("--fdata", "-fdata", default=None, type=str, help="Path to flow")
    p.add_argument("--dadata", "-dadata", default="audio/target", type=str, help="Path to driving audio dataset")
    p.add_argument("--video_list", "-vl", default=None, type=str, nargs="+", help="list of input videos")
    p.add_argument("--fps", "-fps", default=30, type=int, help="frame rate of input video")
    p.add_argument("--subsample_rate", "-subsample", default=1, type=int, help="rate for subsampling the video")
    p.add_argument("--temp", "-temp", default=0.1, type=float, help="Temperature value")
    p.add_argument("--threshold", "-th", default=0.0, type=float, help="Threshold value")
    p.add_argument("--l2", "-l2", default=True, action="store_false", help="To use l2 norm or not")
    p.add_argument("--interpolation", "-nintp", default=True, action="store_false", help="Interpolate frames at eval")
    p.add_argument("--img_size", "-size", default=224, type=int, help="resize image to this size")
    p.add_argument("--n_negs", "-negs", default=20, type=int, help="Number negative frames to use when training")
    p.add_argument("--window", "-w", default=20, type=int, help="Size of temporal window")
    p.add_argument("--train_stride", "-train_stride", default=4, type=int, help="Stride length")
    p.add_argument("--stride", "-stride", default=4, type=int, help="Stride length")
    p.add_argument("--new_video_length", "-nvl", default=30, type=int, help="Length of new video")
    p.add_argument("--alpha", "-alpha", default=0.5, type=float, help="alpha for validation to control driving audio")
    p.add_argument("--SF", "-SF", default=5, type=int, help="slomo factor N")
    p.add_argument("-long", "--long", dest="long", default=False, action="store_true")
    p.add_argument("-fb", "--frames_bar", dest="frames_bar", default=False, action="store_true", help="Visualize transitions.")
    p.add_argument("--epochs", default=60, type=int, metavar="N", help="number of total epochs to run")
    p.add_argument("--size", default=224, type=int, metavar="N", help="primary image input size")
    p.add_argument("--start_epoch", default=None, type=int, metavar="N", help="manual epoch number (useful on restarts)")
    p.add_argument("--batch_size", "-bs", default=32, type=int, metavar="N", help="mini-batch size (default: 32)")
    p.add_argument("--mini_batchsize", "-mbs", default=150, type=int, help="mini-batch size for target frames")
    p.add_argument("--lr", "-lr", default=10e-3, type=float, metavar="LR", help="initial learning rate")
    p.add_argument("--lr_steps", default=30, type=int, metavar="LRSteps", help="epochs to decay learning rate by 10")
    p.add_argument("--momentum", default=0.9, type=float, metavar="M", help="momentum")
    p.add_argument("--weight_decay", "--wd", default=0.0001, type=float, metavar="W", help="weight decay (default: 1e-4)")
    p.add_argument("--workers", "-j", default=4, type=int, metavar="N", help="number of data loading workers")
    p.add_argument("--print_freq", "-p", default=5, type=int, metavar="N", help="print frequency")
    p.add_argument("--log_freq", "-lf", default=10, type=int, metavar="N", help="frequency to write in tensorboard")
    p.add_argument("--resume", default="", type=str, metavar="PATH", help="path to latest checkpoint (default: none)")
    p.add_argument("-e", "--evaluate", dest="evaluate", action="store_true", help="evaluate model on validation set")
    p.add_argument("-da", "--driving_audio", default=None, type=str, nargs="+", help="list of target audios")
    p.add_argument("-daf", "--da_feats", default="VGG", type=str, help="type of feats for audio conditioning")
    p.add_argument("-daf_resume", "--daf_resume", default="", type=str, nargs="+", help="List of paths to best VideoForAudio ckpt")
    p.add_argument("-ve", "--visualize_evaluate", dest="visualize_evaluate", action="store_true")
    p.add_argument("-vf", "--val_freq", default=5, type=int, metavar="VF", help="frequency to call validate during train")
    p.add_argument("--logdir", default="./logs", help="folder to output tensorboard logs")
    p.add_argument("--logname", default="exp", help="name of the experiment for checkpoints and logs")
    p.add_argument("-rf", "--results_folder", default="results", type=str, help="folder for result videos")
    p.add_argument("--ckpt", default="./ckpt", help="folder to output checkpoints")
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

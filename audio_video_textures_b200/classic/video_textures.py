"""Drop-in for baselines/classic_video_textures/video_textures.py: the sampling walk (:32-241), the
sigma sweep of `main` (:244-454) restricted to the hot path, and the CLI argument surface (:457-551).

Differences that are deliberate and documented:
  * the reference's `utils.{Logger,read_data,save_video}` module does not exist; `utils.py` here is
    the shim (synthetic / tensor / directory-of-frames reader, optional PNG writer);
  * SuperSloMo interpolation, tensorboard figures and ffmpeg are I/O glue outside the hot path;
  * modes 2/3 crash in the reference after the walk (`new_frames_intp` NameError, :231) — here they
    return the same frame list the reference prints just before crashing, including the mode-3 defect
    that `this_frame` never advances inside the loop (:209).
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from .. import engine
from .computeD1 import compute_D1
from .computeD2 import compute_D2
from .q_learning import q_learning

SIGMAS = [4.45, 4.5, 4.52, 4.55, 4.58]        # video_textures.py:250


def texture_walk(P, model_type: int, fps: int, new_video_length: int, stride: int, filter_size: int,
                 start: int = 100):
    """The walk of video_textures.py:43-209 over the survivor lists of P (non-zeros per row).

    The reference calls `P[this].nonzero().cpu()` (a device sync) on every step; here the ordered
    survivor lists of ALL rows are compacted once on the GPU (avtex_row_nnz / avtex_csr_fill) and
    the walk runs on the host — or, for matrices whose survivor lists run to gigabytes, `P` is an
    engine.SurvivorRows that fetches only the visited rows.  The draw itself stays `np.random.choice` on numpy's global legacy
    generator, so the sequence is bit-identical to the reference under the same `np.random.seed`.
    Returns (new_frames_list, jump_count).
    """
    if isinstance(P, engine.SurvivorRows):                 # lazy: only the rows the walk visits are compacted / copied
        n_rows = len(P)
        survivors = P.__getitem__
    else:
        if isinstance(P, tuple):
            rowptr, colidx = P
        else:
            rowptr, colidx = engine.csr_from_matrix(P if P.is_cuda else P.cuda())
        n_rows = len(rowptr) - 1

        def survivors(i):
            return colidx[rowptr[i]:rowptr[i + 1]]

    def draw(i):
        # np.random.choice(a) on the legacy global generator IS a[np.random.randint(0, len(a))] — the same single
        # bounded draw (tests/test_host_cpu.py::test_walk_draw_is_numpy_choice) — minus ~4 us of argument handling per step
        a = survivors(i)
        return int(a[np.random.randint(0, len(a))])

    new_video_length = fps * new_video_length
    jump_count = 0
    if model_type == 1:
        this_frame = start
        new_frames_list = [start]
        while len(new_frames_list) < new_video_length:
            next_frame = draw(this_frame)
            if next_frame != this_frame + 1:
                jump_count += 1
            new_frames_list.append(next_frame)
            this_frame = next_frame
    elif model_type == 2:
        this_frame = start
        new_frames_list = list(range(this_frame, this_frame + stride))
        this_frame += stride
        while len(new_frames_list) < new_video_length:
            next_frame = draw(this_frame)
            if next_frame != this_frame + 1:
                jump_count += 1
            new_frames_list.extend(range(next_frame, min(next_frame + stride, n_rows)))
            this_frame = min(next_frame + stride, n_rows - 1)
    else:
        this_frame = start
        new_frames_list = list(range(this_frame, this_frame + filter_size))
        while len(new_frames_list) < new_video_length:
            next_frame = draw(this_frame)
            if next_frame != this_frame + 1:
                jump_count += 1
            new_frames_list.extend(range(this_frame * stride + (filter_size - stride),
                                         this_frame * stride + filter_size))
        this_frame = next_frame  # noqa: F841 — reference: assigned after the loop only
    return new_frames_list, jump_count


def audio_video_texture(args, P, frames, output_video_folder, output_video_folder_intp, audio=None,
                        output_audio_file="", intp_model=None):
    """Same signature as the reference (:32-41).  Reads args.{fps,new_video_length,model_type,stride,
    filter_size}; consumes the numpy global RNG; prints `Frames list:`; returns jump_count.
    Frames are written as PNGs (reference :213-226, without the red progress bar) when a folder is
    given; interpolated output and the wav file are outside the hot path."""
    new_frames_list, jump_count = texture_walk(P, args.model_type, args.fps, args.new_video_length,
                                               args.stride, args.filter_size)
    print("Frames list: ", new_frames_list)
    if output_video_folder:
        from .utils import write_frames
        write_frames(frames, new_frames_list, output_video_folder)
        print("Written_{}".format(output_video_folder))
    audio_video_texture.last_frames = new_frames_list
    return jump_count


def main(args, video_name: str):
    """Hot-path part of the reference `main` (:244-454): for each sigma factor of the sweep (:250) produce
    P3_new and walk it.  D1, D2 and the converged future cost do not depend on sigma, so they are computed
    ONCE; per sigma factor only sigma3 / P3 / P3_new and the walk are re-evaluated (the reference recomputes
    everything five times, :265-284).  Prints the reference's `Eps:` / `Non Zero in P3:` / `Frames list:` lines.
    Returns dict(sigmas, jump_counts, sequences)."""
    from .utils import read_data
    input_frames, video, args.fps, audio, args.sr, _ = read_data(args, video_name)
    if args.feats != "RGB":
        raise NotImplementedError("only -f RGB is on the hot path")
    frames = input_frames if input_frames.is_cuda else input_frames.cuda(non_blocking=True)
    stride = 1 if args.model_type in (1, 2) else args.stride            # video_textures.py:276-283
    # only P3_new is consumed below, so D1 itself need not exist: with -stride s the filter reads D1[i,j] only where
    # i = j (mod s), and engine.distance_filter computes just those residue-class blocks (1/s of the pairs; same bits)
    D2, D3, used = engine.distance_filter(frames, args.filter_size, stride, p=0.7)
    fc = engine.future_cost_fused(D3, 0.997, verbose=True)
    stats = engine.new_stats(D3.device)
    D3_new = engine.future_cost_finalize(D3, fc.mvec, 0.997, stats=stats)
    total, nnz = engine.read_stats(stats)
    jump_counts, new_sigmas, sequences = [], [], []
    for value in torch.tensor(SIGMAS, dtype=torch.float32):
        sigma = engine.sigma_from_stats(total, nnz, value)
        P3, P3_new, counts = engine.transition_probs(D3_new, sigma, threshold=args.threshold, want_P=False,
                                                     want_counts=True)
        print("Non Zero in P3:", int(counts[0].item()))
        new_sigmas.append(float(sigma))
        out_dir = None
        if getattr(args, "write_frames", False):
            out_dir = os.path.join(args.results_folder, "{}_{}_{:.4f}".format(video_name, args.model_type, float(sigma)))
        jump_counts.append(audio_video_texture(args, engine.csr_from_matrix(P3_new, counts), video, out_dir, None,
                                               audio, ""))
        sequences.append(audio_video_texture.last_frames)
    return dict(sigmas=new_sigmas, jump_counts=jump_counts, sequences=sequences, n_sweeps=fc.n_sweeps,
                distance_path=used)


def build_parser() -> argparse.ArgumentParser:
    """Argument surface of video_textures.py:457-549 (same flags, defaults and meanings)."""
    parser = argparse.ArgumentParser(description="=Video Textures")
    parser.add_argument("--model_type", "-m", default=1, type=int,
                        help="(1) Classic (2) Classic + (3) Classic ++")
    parser.add_argument("--vdata", "-vdata", default=None, type=str, help="Path to video dataset")
    parser.add_argument("--adata", "-adata", default=None, type=str, help="Path to audio dataset")
    parser.add_argument("--video_list", "-vl", default=None, type=str, nargs="+", help="list of input videos")
    parser.add_argument("--feats", "-f", default="RGB", type=str, help="Features to use")
    parser.add_argument("--slow", "-s", dest="slow", action="store_true", help="set false for large videos")
    parser.add_argument("--fps", "-fps", default=30, type=int, help="frame rate of input video")
    parser.add_argument("--sr", "-sr", default=22050, type=int, help="rate of input audio")
    parser.add_argument("--filter_size", "-fs", default=40, type=int, help="filter size of gaussian filter")
    parser.add_argument("--batch_size", "-bs", default=64, type=int, help="mini batch size")
    parser.add_argument("--stride", "-stride", default=4, type=int, help="stride")
    parser.add_argument("--new_video_length", "-nvl", default=30, type=int, help="frame rate of input video")
    parser.add_argument("--interpolation", "-nintp", default=True, action="store_false",
                        help="Interpolate frames at eval")
    parser.add_argument("--SF", "-SF", default=3, type=int, help="slomo factor N")
    parser.add_argument("--sigma", "-sigma", default=0.5, type=float, help="Sigma value")
    parser.add_argument("--threshold", "-t", default=0.08, type=float, help="Threshold value for P")
    parser.add_argument("-rf", "--results_folder", default="results_classic", type=str,
                        help="folder for result videos")
    parser.add_argument("--logdir", default="./logs", help="folder to output tensorboard logs")
    parser.add_argument("--logname", default="exp_classic", help="name of the experiment")
    # additions (none mandatory): synthetic input when no -vdata is given, optional PNG output
    parser.add_argument("--synthetic", default=None, type=str,
                        help="N,H,W[,seed]: use the deterministic synthetic clip instead of -vdata")
    parser.add_argument("--write_frames", action="store_true", help="write the chosen frames as PNGs")
    return parser


if __name__ == "__main__":
    args = build_parser().parse_args()
    print(args)
    if args.video_list is None:
        if args.vdata is not None and os.path.isdir(args.vdata):
            args.video_list = [f.split(".")[0] for f in sorted(os.listdir(args.vdata)) if not f.startswith(".")]
        else:
            args.video_list = ["synthetic"]
    for itr, video_name in enumerate(args.video_list):
        print("Starting video {}".format(video_name))
        res = main(args, video_name)
        print("sigmas", res["sigmas"], "jump_counts", res["jump_counts"])

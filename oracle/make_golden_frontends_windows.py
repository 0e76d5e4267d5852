"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for (f1) window construction and (f2) the classic feature
modes, produced by the UNMODIFIED reference code (imported through oracle/ref_shim.py) in the build container."""
from __future__ import annotations

import importlib.util
import math
import os
import sys

import numpy as np
import torch

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_windows():
    """(f1) One synthesis step at FRAME level, as cvt/validate.py:327-329, 365-395, 442-493 runs it, with the
    reference's own `split_into_overlapping_segments` and `ContrastivePredictionTemporal.forward` (identity 3D
    encoder + the class's AdaptiveAvgPool3d: a window's embedding is the mean of its frames' channel vectors)."""
    ru = _load(os.path.join(ref_shim.CVT, "utils", "utils.py"), "ref_cvt_utils")
    W, S, mbs, C, temp = 5, 2, 6, 16, 0.1
    L = 23
    T = (L - 1) * S + W
    g = torch.Generator().manual_seed(11)
    frames = torch.randn(T, C, 1, 1, generator=g) + 0.5 * torch.cumsum(torch.randn(T, C, 1, 1, generator=g), 0) / 4
    model = ref_shim.load_contrastive_model(temp, mbs, model_type=1, window=W, stride=S)
    out = dict(frames=frames.numpy(), W=W, S=S, mbs=mbs, L=L, temp=temp)
    for q_id in (3, 0, L - 1, 11):
        with torch.no_grad():
            qf_t = frames[q_id * S: q_id * S + W].unsqueeze(0)                          # validate.py:329
            pos_id = min((q_id + 1), L - 1)                                               # :366
            mask = np.ones(L, dtype=bool)                                                 # :369-371
            mask[[q_id, pos_id]] = False
            n_ids = np.arange(L)[mask, ...]
            target_segment_ids = np.concatenate((np.array([pos_id]), n_ids), axis=0)      # :374
            target_frame_ids = []                                                         # :378-380
            for i in target_segment_ids:
                target_frame_ids.extend(list(np.arange(i * S, i * S + W)))
            target_frame_ids = np.array(target_frame_ids)                                 # :383-385
            _, idxs = np.unique(target_frame_ids, return_index=True)
            target_frame_ids = target_frame_ids[np.sort(idxs)]
            t_video = frames[target_frame_ids]                                            # :388
            t_video_chunks, _ = ru.split_into_overlapping_segments(t_video, mbs, W, S)    # :390-392
            num_valid = len(target_segment_ids)                                           # split_into_batches' num_inputs
            output = torch.zeros(len(target_segment_ids), dtype=torch.float32)            # :422
            for itr in range(math.ceil(len(t_video_chunks) / 1)):                         # :442 (num_gpus = 1)
                b_tf_t = t_video_chunks[itr: itr + 1]
                b_output = model(qf_t, b_tf_t, is_inference=True)                         # :472-479
                lo = itr * mbs
                take = min(num_valid, mbs)
                output[lo: lo + take] = b_output.contiguous().view(-1)[:take]             # :481-493
                num_valid -= mbs                                                          # :522
        out[f"q{q_id}_segment_ids"] = target_segment_ids
        out[f"q{q_id}_frame_ids"] = target_frame_ids
        out[f"q{q_id}_chunks_shape"] = np.asarray(t_video_chunks.shape)
        out[f"q{q_id}_ref_logits"] = output.numpy()
        print("windows q", q_id, "targets", len(target_segment_ids), "chunks", tuple(t_video_chunks.shape))
    np.savez_compressed(os.path.join(OUT, "frontend_windows.npz"), **out)


class _ToyResNet(torch.nn.Module):
    """Stand-in for torchvision.models.resnet18(pretrained=True) (weights are not available offline): what matters to
    classic/computeD1.py:98-150 is `list(children())[:-1]` -> [N, C, 1, 1] features."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.conv = torch.nn.Conv2d(3, 24, 3, stride=2)
        with torch.no_grad():
            self.conv.weight.copy_(torch.randn(self.conv.weight.shape, generator=g) * 0.2)
            self.conv.bias.copy_(torch.randn(24, generator=g) * 0.1)
        self.relu = torch.nn.ReLU()
        self.pool = torch.nn.AdaptiveAvgPool2d(1)
        self.fc = torch.nn.Linear(24, 10)


def make_features():
    """(f2) compute_D1 feats="ResNet" (dense and tiled "slow" branch) with the reference's own code; the feature
    producer is a seeded toy network patched in for the unavailable pretrained ResNet-18."""
    cD1, cD2, ql = ref_shim.load_classic()
    import computeD1 as ref_mod
    toy = _ToyResNet()
    ref_mod.models.resnet18 = lambda pretrained=True: toy
    g = torch.Generator().manual_seed(3)
    n = 70
    frames = torch.rand(n, 3, 12, 12, generator=g) + 0.3 * torch.sin(torch.arange(n).float().view(-1, 1, 1, 1) / 5)
    with torch.no_grad():
        feats = torch.nn.Sequential(*list(toy.children())[:-1])(frames).view(n, -1)
    f = torch.tensor(4.5)
    out = dict(frames=frames.numpy(), image_feats=feats.numpy(), f=4.5)
    D1, P1, s = cD1(frames, f, "ResNet", slow=False)
    out["ref_D1_dense"], out["ref_P1_dense"], out["ref_sigma_dense"] = D1.numpy(), P1.numpy(), float(s)
    D1s, P1s, ss = cD1(frames, f, "ResNet", slow=True, batch_size=16)
    out["ref_D1_slow16"], out["ref_sigma_slow16"] = D1s.numpy(), float(ss)
    print("features: dense D1", D1.shape, "sigma", float(s), "| slow bs=16 untouched entries (== 1):", int((D1s == 1).sum()))

    # feats="ResNet_VGGish": torch.hub's VGGish is patched the same way (a seeded per-second toy embedding)
    class _ToyVGGish:
        def eval(self):
            return self

        def forward(self, audio, sr):
            a = torch.as_tensor(audio, dtype=torch.float32)
            secs = a.shape[0] // sr
            blocks = a[: secs * sr].view(secs, sr)
            basis = torch.randn(sr, 12, generator=torch.Generator().manual_seed(9)) / sr ** 0.5
            return torch.relu(blocks @ basis)

    ref_mod.torch.hub.load = lambda *a, **k: _ToyVGGish()
    fps, sr = 10, 400
    audio = torch.randn(7 * sr + 37, generator=g).numpy()
    audio_feats = _ToyVGGish().forward(audio, sr)
    out["audio"], out["audio_feats"], out["fps"], out["sr"] = audio, audio_feats.numpy(), fps, sr
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        D1j, P1j, sj = cD1(frames, f, "ResNet_VGGish", audio=audio, sr=sr, fps=fps, slow=False)
        D1js, P1js, sjs = cD1(frames, f, "ResNet_VGGish", audio=audio, sr=sr, fps=fps, slow=True, batch_size=16)
    out["ref_D1_joint_dense"], out["ref_sigma_joint_dense"] = D1j.numpy(), float(sj)
    out["ref_D1_joint_slow16"], out["ref_sigma_joint_slow16"] = D1js.numpy(), float(sjs)
    print("joint: dense", D1j.shape, "slow zeros", int((D1js == 0).sum()))
    np.savez_compressed(os.path.join(OUT, "frontend_features.npz"), **out)

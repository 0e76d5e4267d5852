"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference modules from /root/reference (build container only; the
GPU box has no /root/reference, so nothing run there may import this file).  Recipe from
SURVEY.md Appendix A: stub the reference's dead top-level imports and make `.cuda()` the
identity on a CPU-only host.  No reference source is copied.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF = os.environ.get("AVTEX_REFERENCE", "/root/reference")
CLASSIC = os.path.join(REF, "baselines", "classic_video_textures")
CVT = os.path.join(REF, "contrastive_video_textures")


def available() -> bool:
    return os.path.isdir(CLASSIC)


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    return mod


def _patch_cuda():
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def load_classic():
    """Returns (compute_D1, compute_D2, q_learning) — the reference's own functions."""
    for name in ["librosa", "IPython", "IPython.display"]:
        _stub(name)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.pyplot")
    _patch_cuda()
    if CLASSIC not in sys.path:
        sys.path.insert(0, CLASSIC)
    from computeD1 import compute_D1
    from computeD2 import compute_D2
    from q_learning import q_learning
    return compute_D1, compute_D2, q_learning


class _Identity3D(torch.nn.Module):
    """Toy encoder: input (B, C, window, H, W) with window=H=W=1 -> (B, C, 1, 1, 1); after the
    reference's AdaptiveAvgPool3d the embedding IS the input row."""

    def forward(self, x):
        return x


def load_contrastive_model(temp: float, mini_batchsize: int, model_type: int = 1, window: int = 1, stride: int = 1):
    """Builds the reference ContrastivePredictionTemporal with identity encoders
    (default window=1, stride=1, 1x1 'frames' whose channel vector is the embedding; with window > 1 the class's
    own AdaptiveAvgPool3d makes a window's embedding the mean of its frames' channel vectors)."""
    for name in ["librosa", "resampy", "soundfile", "ipdb", "imageio", "IPython", "IPython.display"]:
        _stub(name)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.pyplot")
    _stub("tensorboardX", SummaryWriter=object)
    _stub("slowfast")
    _stub("slowfast.utils")
    _stub("slowfast.utils.parser", load_config=lambda *a, **k: None, parse_args=lambda *a, **k: None)
    _stub("slowfast.visualization")
    _stub("slowfast.visualization.predictor", ActionPredictor=object)
    _stub("slowfast.visualization.utils", process_cv2_inputs=lambda *a, **k: None)
    _patch_cuda()
    # the classic dir also has modules named `utils`-less; make sure cvt/ wins for `models`, `utils`
    for m in ["models", "utils"]:
        sys.modules.pop(m, None)
    if CVT in sys.path:
        sys.path.remove(CVT)
    sys.path.insert(0, CVT)
    from models import ContrastivePredictionTemporal
    model = ContrastivePredictionTemporal(
        _Identity3D(), _Identity3D(), None, model_type, fc_dim=0, temp=temp, window=window, stride=stride,
        threshold=0.0, mini_batchsize=mini_batchsize, enc_arch="toy")
    model.eval()
    return model


def reference_chunk_scores(model, q_emb: torch.Tensor, t_chunk: torch.Tensor) -> torch.Tensor:
    """Runs the reference forward on one chunk: q_emb [D], t_chunk [mbs, D] -> [mbs] logits."""
    D = q_emb.numel()
    q_f = q_emb.view(1, 1, D, 1, 1)                      # (B, window, C, H, W)
    t_f = t_chunk.view(1, t_chunk.shape[0], D, 1, 1)      # (B, chunk_frames, C, H, W); eval path re-windows
    with torch.no_grad():
        out = model(q_f, t_f, is_inference=True)
    return out.view(-1)

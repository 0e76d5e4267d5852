"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for the "next" rows (front ends either side of the hot path).

Run in the BUILD CONTAINER (needs /root/reference):  python -m oracle.make_golden_frontends
Writes tests/golden/frontend_*.npz; every `ref_` key is the output of the UNMODIFIED reference code.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

from oracle import audio as oa
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_ref_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_audio():
    """(f3) reference utils/mel_features.py + vggish_utils.waveform_to_examples on a seeded waveform."""
    utils_dir = os.path.join(ref_shim.CVT, "utils")
    pkg = types.ModuleType("refutils")
    pkg.__path__ = [utils_dir]
    sys.modules["refutils"] = pkg
    for stub in ("resampy", "soundfile"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    mel = _load_ref_module(os.path.join(utils_dir, "mel_features.py"), "refutils.mel_features")
    sys.modules["refutils.mel_features"] = mel
    params = _load_ref_module(os.path.join(utils_dir, "vggish_params.py"), "refutils.vggish_params")
    sys.modules["refutils.vggish_params"] = params
    vu = _load_ref_module(os.path.join(utils_dir, "vggish_utils.py"), "refutils.vggish_utils")
    out = {}
    for name, secs, ch, seed in (("mono", 3.0, 1, 0), ("stereo", 2.2, 2, 1)):
        wave = oa.synth_waveform(secs, seed=seed, channels=ch)
        ref_examples = vu.waveform_to_examples(wave, 16000)
        ref_logmel = mel.log_mel_spectrogram(wave if ch == 1 else wave.mean(axis=1), audio_sample_rate=16000, log_offset=0.01,
                                             window_length_secs=0.025, hop_length_secs=0.010, num_mel_bins=64,
                                             lower_edge_hertz=125, upper_edge_hertz=7500)
        got = oa.waveform_to_examples(wave, 16000)
        assert np.array_equal(np.asarray(got), np.asarray(ref_examples)), "oracle/audio.py differs from the reference"
        out[f"{name}_wave"] = wave
        # the examples are overlapping row windows of ref_logmel (pure indexing): store their shape and a checksum
        assert np.array_equal(np.asarray(ref_examples), oa.frame(np.asarray(ref_logmel), 100, 10))
        out[f"{name}_ref_examples_shape"] = np.asarray(ref_examples.shape)
        out[f"{name}_ref_examples_sum"] = np.float64(np.asarray(ref_examples).sum())
        out[f"{name}_ref_logmel"] = np.asarray(ref_logmel, dtype=np.float64)
        print(name, "examples", ref_examples.shape, "logmel range", float(ref_logmel.min()), float(ref_logmel.max()))
    np.savez_compressed(os.path.join(OUT, "frontend_audio.npz"), **out)


if __name__ == "__main__":
    assert ref_shim.available(), "needs /root/reference"
    which = sys.argv[1:] or ["audio", "windows", "features"]
    if "audio" in which:
        make_audio()
    if "windows" in which:
        from oracle.make_golden_frontends_windows import make_windows
        make_windows()
    if "features" in which:
        from oracle.make_golden_frontends_windows import make_features
        make_features()

"""(f3) Audio front end of contrastive_video_textures: waveform -> log-mel examples, on the GPU.

Same function names and arguments as the reference's `utils/vggish_utils.py` / `utils/mel_features.py`
(TF-VGGish front end): `waveform_to_examples(data, sample_rate)` returns [num_examples, 100, 64] patches
(1.0 s windows, 0.1 s hop — cvt/utils/vggish_params.py:35-36) of the log-mel spectrogram (25 ms periodic-Hann
STFT, 10 ms hop, 64 mel bands 125-7500 Hz, log(mel + 0.01)).  The tables (window, mel matrix) are computed on
the host with the reference's own float64 formulas; the spectrogram itself runs in libavtex (csrc/audio.cu, fp64
arithmetic like numpy's).  Resampling (resampy) is a separate producer: audio must arrive at 16 kHz.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib, engine

# cvt/utils/vggish_params.py:21-36
SAMPLE_RATE = 16000
STFT_WINDOW_LENGTH_SECONDS = 0.025
STFT_HOP_LENGTH_SECONDS = 0.010
NUM_MEL_BINS = 64
MEL_MIN_HZ = 125
MEL_MAX_HZ = 7500
LOG_OFFSET = 0.01
EXAMPLE_WINDOW_SECONDS = 1.0
EXAMPLE_HOP_SECONDS = 0.1

_MEL_BREAK_FREQUENCY_HERTZ = 700.0
_MEL_HIGH_FREQUENCY_Q = 1127.0


def periodic_hann(window_length: int) -> np.ndarray:
    """mel_features.py:49-69."""
    return 0.5 - (0.5 * np.cos(2 * np.pi / window_length * np.arange(window_length)))


def hertz_to_mel(frequencies_hertz):
    """mel_features.py:101-110 (HTK formula)."""
    return _MEL_HIGH_FREQUENCY_Q * np.log(1.0 + (frequencies_hertz / _MEL_BREAK_FREQUENCY_HERTZ))


def spectrogram_to_mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, audio_sample_rate=8000,
                              lower_edge_hertz=125.0, upper_edge_hertz=3800.0) -> np.ndarray:
    """mel_features.py:113-185: triangular mel weights [num_spectrogram_bins, num_mel_bins], DC row zeroed."""
    nyquist_hertz = audio_sample_rate / 2.0
    if lower_edge_hertz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_edge_hertz)
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f" % (lower_edge_hertz, upper_edge_hertz))
    if upper_edge_hertz > nyquist_hertz:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f" % (upper_edge_hertz, nyquist_hertz))
    bins_mel = hertz_to_mel(np.linspace(0.0, nyquist_hertz, num_spectrogram_bins))
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    weights = np.empty((num_spectrogram_bins, num_mel_bins))
    for i in range(num_mel_bins):
        lower, center, upper = edges[i:i + 3]
        weights[:, i] = np.maximum(0.0, np.minimum((bins_mel - lower) / (center - lower), (upper - bins_mel) / (upper - center)))
    weights[0, :] = 0.0
    return weights


def log_mel_spectrogram(data, audio_sample_rate=8000, log_offset=0.0, window_length_secs=0.025,
                        hop_length_secs=0.010, device=None, **kwargs) -> torch.Tensor:
    """mel_features.py:188-223 on the GPU.  data: 1-D (mono) or [n_samples, channels] waveform (numpy / tensor).
    Returns a CUDA fp32 tensor [num_frames, num_mel_bins]."""
    if not torch.cuda.is_available():
        raise RuntimeError("audio_video_textures_b200 needs a CUDA device (B200); there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    wave = torch.as_tensor(np.asarray(data) if not torch.is_tensor(data) else data).to(dev, torch.float64).contiguous()
    channels = 1 if wave.dim() == 1 else wave.shape[1]
    n_samples = wave.shape[0]
    win = int(round(audio_sample_rate * window_length_secs))
    hop = int(round(audio_sample_rate * hop_length_secs))
    fft_length = 2 ** int(np.ceil(np.log(win) / np.log(2.0)))
    n_frames = 1 + int(np.floor((n_samples - win) / hop))                      # mel_features.frame
    if n_frames < 1:
        raise ValueError("waveform shorter than one analysis window")
    mel = spectrogram_to_mel_matrix(num_spectrogram_bins=fft_length // 2 + 1, audio_sample_rate=audio_sample_rate, **kwargs)
    n_mel = mel.shape[1]
    d_window = torch.from_numpy(periodic_hann(win)).to(dev)
    d_mel = torch.from_numpy(np.ascontiguousarray(mel)).to(dev)
    out = torch.empty((n_frames, n_mel), dtype=torch.float32, device=dev)
    _lib.call("avtex_logmel", _lib.ptr(wave), n_samples, channels, win, hop, fft_length, _lib.ptr(d_window),
              _lib.ptr(d_mel), n_mel, C.c_double(log_offset), _lib.ptr(out), n_frames, engine._dev(out), engine._stream(out))
    return out


def frame(data: torch.Tensor, window_length: int, hop_length: int) -> torch.Tensor:
    """mel_features.frame (:22-46) for a [rows, bands] CUDA feature matrix -> [num_frames, window_length, bands]."""
    n_rows, bands = data.shape
    n = 1 + int(np.floor((n_rows - window_length) / hop_length))
    if n < 1:
        return torch.empty((0, window_length, bands), dtype=torch.float32, device=data.device)
    data = data.contiguous()
    out = torch.empty((n, window_length, bands), dtype=torch.float32, device=data.device)
    _lib.call("avtex_frame_examples", _lib.ptr(data), n_rows, bands, window_length, hop_length, _lib.ptr(out), n,
              engine._dev(out), engine._stream(out))
    return out


def waveform_to_examples(data, sample_rate, device=None) -> torch.Tensor:
    """vggish_utils.py:27-69.  Returns CUDA fp32 [num_examples, 100, 64]."""
    if sample_rate != SAMPLE_RATE:
        raise NotImplementedError(f"resampling {sample_rate} -> {SAMPLE_RATE} Hz (resampy) is a separate producer; "
                                  "pass 16 kHz audio")
    log_mel = log_mel_spectrogram(data, audio_sample_rate=SAMPLE_RATE, log_offset=LOG_OFFSET,
                                  window_length_secs=STFT_WINDOW_LENGTH_SECONDS, hop_length_secs=STFT_HOP_LENGTH_SECONDS,
                                  num_mel_bins=NUM_MEL_BINS, lower_edge_hertz=MEL_MIN_HZ, upper_edge_hertz=MEL_MAX_HZ,
                                  device=device)
    features_sample_rate = 1.0 / STFT_HOP_LENGTH_SECONDS
    example_window_length = int(round(EXAMPLE_WINDOW_SECONDS * features_sample_rate))
    example_hop_length = int(round(EXAMPLE_HOP_SECONDS * features_sample_rate))
    return frame(log_mel, example_window_length, example_hop_length)

"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (numpy float64, the reference's own arithmetic library) of the audio front end of
contrastive_video_textures:

    frame                 utils/mel_features.py:22-46
    periodic_hann         utils/mel_features.py:49-69
    stft_magnitude        utils/mel_features.py:72-93
    mel matrix            utils/mel_features.py:113-185
    log_mel_spectrogram   utils/mel_features.py:188-223
    waveform_to_examples  utils/vggish_utils.py:27-69   (parameters utils/vggish_params.py:21-36)

Parity status: pinned against the UNMODIFIED reference modules imported in the build container
(oracle/make_golden_frontends.py asserts bit-equality and stores tests/golden/frontend_audio.npz;
tests/test_oracle_golden.py re-checks the restatement against that fixture anywhere).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE, WIN_S, HOP_S, N_MEL, MEL_LO, MEL_HI, LOG_OFFSET = 16000, 0.025, 0.010, 64, 125, 7500, 0.01
EXAMPLE_WINDOW_SECONDS, EXAMPLE_HOP_SECONDS = 1.0, 0.1


def frame(data, window_length, hop_length):
    num_samples = data.shape[0]
    num_frames = 1 + int(np.floor((num_samples - window_length) / hop_length))
    shape = (num_frames, window_length) + data.shape[1:]
    strides = (data.strides[0] * hop_length,) + data.strides
    return np.lib.stride_tricks.as_strided(data, shape=shape, strides=strides)


def periodic_hann(window_length):
    return 0.5 - (0.5 * np.cos(2 * np.pi / window_length * np.arange(window_length)))


def stft_magnitude(signal, fft_length, hop_length, window_length):
    frames = frame(signal, window_length, hop_length)
    return np.abs(np.fft.rfft(frames * periodic_hann(window_length), int(fft_length)))


def hertz_to_mel(f):
    return 1127.0 * np.log(1.0 + (f / 700.0))


def mel_matrix(num_mel_bins, num_spectrogram_bins, audio_sample_rate, lower_edge_hertz, upper_edge_hertz):
    nyquist = audio_sample_rate / 2.0
    bins_mel = hertz_to_mel(np.linspace(0.0, nyquist, num_spectrogram_bins))
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    w = np.empty((num_spectrogram_bins, num_mel_bins))
    for i in range(num_mel_bins):
        lo, ce, up = edges[i:i + 3]
        w[:, i] = np.maximum(0.0, np.minimum((bins_mel - lo) / (ce - lo), (up - bins_mel) / (up - ce)))
    w[0, :] = 0.0
    return w


def log_mel_spectrogram(data, audio_sample_rate=SAMPLE_RATE, log_offset=LOG_OFFSET, window_length_secs=WIN_S,
                        hop_length_secs=HOP_S, num_mel_bins=N_MEL, lower_edge_hertz=MEL_LO, upper_edge_hertz=MEL_HI):
    win = int(round(audio_sample_rate * window_length_secs))
    hop = int(round(audio_sample_rate * hop_length_secs))
    fft_length = 2 ** int(np.ceil(np.log(win) / np.log(2.0)))
    spec = stft_magnitude(data, fft_length, hop, win)
    mel = np.dot(spec, mel_matrix(num_mel_bins, spec.shape[1], audio_sample_rate, lower_edge_hertz, upper_edge_hertz))
    return np.log(mel + log_offset)


def waveform_to_examples(data, sample_rate):
    if len(data.shape) > 1:
        data = np.mean(data, axis=1)
    assert sample_rate == SAMPLE_RATE, "resampling is outside the oracle (resampy)"
    log_mel = log_mel_spectrogram(data)
    rate = 1.0 / HOP_S
    return frame(log_mel, int(round(EXAMPLE_WINDOW_SECONDS * rate)), int(round(EXAMPLE_HOP_SECONDS * rate)))


def synth_waveform(seconds: float, seed: int = 0, channels: int = 1) -> np.ndarray:
    """Deterministic test signal: a few drifting partials + noise + a silent stretch, in [-1, 1]."""
    rs = np.random.RandomState(seed)
    n = int(seconds * SAMPLE_RATE)
    t = np.arange(n) / SAMPLE_RATE
    x = np.zeros((n, channels))
    for c in range(channels):
        for f0, amp in ((220.0, 0.4), (880.0, 0.25), (3520.0, 0.1)):
            x[:, c] += amp * np.sin(2 * np.pi * (f0 * (1 + 0.1 * c)) * t * (1 + 0.05 * np.sin(2 * np.pi * 0.7 * t)))
        x[:, c] += 0.05 * rs.randn(n)
    x[n // 3: n // 3 + SAMPLE_RATE // 4] *= 1e-4                     # near silence: the log offset matters here
    x = np.clip(x, -1.0, 1.0)
    return x[:, 0] if channels == 1 else x

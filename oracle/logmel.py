"""float64 numpy restatement of the Whisper log-mel frontend -- TEST INFRASTRUCTURE ONLY.

Follows HF ``WhisperFeatureExtractor._torch_extract_fbank_features``
(transformers/models/whisper/feature_extraction_whisper.py:135-164, ctor :69-103) which the
reference calls through ``processor(audio["array"], sampling_rate=..., return_tensors="pt")``
(data_utils.py:56-58).  Pinned against the live HF extractor by
tests/golden/make_golden.py -> tests/golden/logmel_*.npz.
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
SAMPLE_RATE = 16000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, lin)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    lin = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), lin)


def mel_filter_bank_slaney(n_mels: int, n_freqs: int = 1 + N_FFT // 2, fmin: float = 0.0,
                           fmax: float = 8000.0, sr: int = SAMPLE_RATE) -> np.ndarray:
    """[n_freqs, n_mels] float64 triangular filters, slaney mel scale + slaney area norm."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freqs)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, n_samples: int = 480000) -> np.ndarray:
    """audio float32 [B, n] (or [n]) -> float64 [B, n_mels, n_samples // 160].

    pad/truncate to n_samples; reflect-pad 200; periodic hann(400); 400-pt DFT every 160
    samples; drop the last frame; |.|^2; mel; log10(clamp 1e-10); max(., utterance max - 8);
    (. + 4) / 4."""
    a = np.atleast_2d(np.asarray(audio, dtype=np.float32)).astype(np.float64)
    B = a.shape[0]
    if a.shape[1] < n_samples:
        a = np.pad(a, ((0, 0), (0, n_samples - a.shape[1])))
    a = a[:, :n_samples]
    pad = N_FFT // 2
    a = np.pad(a, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = n_samples // HOP  # after dropping the last of 1 + n_samples // HOP
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)
    # hann window as torch builds it: float32 values
    window = window.astype(np.float32).astype(np.float64)
    fb = mel_filter_bank_slaney(n_mels).astype(np.float32).astype(np.float64)  # [201, n_mels]
    out = np.empty((B, n_mels, n_frames), dtype=np.float64)
    for b in range(B):
        frames = a[b][idx] * window[None, :]
        spec = np.fft.rfft(frames, axis=1)
        power = spec.real ** 2 + spec.imag ** 2            # [frames, 201]
        mel = power @ fb                                   # [frames, n_mels]
        logm = np.log10(np.maximum(mel, 1e-10)).T
        logm = np.maximum(logm, logm.max() - 8.0)
        out[b] = (logm + 4.0) / 4.0
    return out

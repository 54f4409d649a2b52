"""Log-mel frontend: drop-in for the ``WhisperFeatureExtractor`` call the reference makes in
``map_to_feats`` (data_utils.py:55-59): ``processor(audio["array"], sampling_rate=...,
return_tensors="pt").input_features`` -> float32 [B, n_mels, 3000].

The STFT / mel / log pipeline runs in one sm_100a kernel (csrc/logmel.cu); only the constant mel
filterbank table is built on the host (once, float64, slaney scale + slaney norm exactly as
transformers.audio_utils.mel_filter_bank does for Whisper).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import functional as F


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)),
                    3.0 * f / 200.0)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)


def whisper_mel_filters(n_mels: int, n_fft: int = 400, sampling_rate: int = 16000) -> np.ndarray:
    """[1 + n_fft // 2, n_mels] float32 (feature_extraction_whisper.py:95-103)."""
    n_freqs = 1 + n_fft // 2
    hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(8000.0), n_mels + 2))
    fft_freqs = np.linspace(0, sampling_rate // 2, n_freqs)
    diff = np.diff(hz)
    slopes = hz[None, :] - fft_freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    fb *= (2.0 / (hz[2:n_mels + 2] - hz[:n_mels]))[None, :]
    return fb.astype(np.float32)


class BatchFeature(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class LogMelFrontend:
    """Callable with the WhisperFeatureExtractor arguments the reference uses."""

    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 chunk_length: int = 30, n_fft: int = 400, device: Union[str, torch.device] = "cuda"):
        if hop_length != 160 or n_fft != 400:
            raise NotImplementedError("the CUDA frontend implements Whisper's n_fft=400 / hop=160")
        self.feature_size, self.sampling_rate = feature_size, sampling_rate
        self.hop_length, self.chunk_length, self.n_fft = hop_length, chunk_length, n_fft
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.device = torch.device(device)
        self.mel_filters = whisper_mel_filters(feature_size, n_fft, sampling_rate)
        self._filters_dev: Optional[torch.Tensor] = None

    def _filters(self) -> torch.Tensor:
        if self._filters_dev is None:
            self._filters_dev = torch.from_numpy(self.mel_filters).to(self.device)
        return self._filters_dev

    def features_from_device_audio(self, audio: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio float32 [B, L] already on the GPU -> float32 [B, n_mels, frames] on the GPU."""
        return F.log_mel(audio, self._filters(), self.n_samples, lengths)

    def __call__(self, raw_speech, sampling_rate: Optional[int] = None, return_tensors: Optional[str] = "pt",
                 **unused) -> BatchFeature:
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"sampling rate {sampling_rate} != {self.sampling_rate}")
        if isinstance(raw_speech, torch.Tensor) and raw_speech.is_cuda:
            a = raw_speech if raw_speech.dim() == 2 else raw_speech[None]
            feats = self.features_from_device_audio(a.float().contiguous())
        else:
            seqs: List[np.ndarray]
            if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 2:
                seqs = [r for r in raw_speech]
            elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) and not np.isscalar(raw_speech[0]):
                seqs = [np.asarray(r, dtype=np.float32) for r in raw_speech]
            else:
                seqs = [np.asarray(raw_speech, dtype=np.float32)]
            L = min(max(len(s) for s in seqs), self.n_samples)
            host = torch.zeros((len(seqs), L), dtype=torch.float32, pin_memory=True)
            lens = torch.empty((len(seqs),), dtype=torch.int32)
            for i, s in enumerate(seqs):
                n = min(len(s), L)
                host[i, :n] = torch.from_numpy(np.ascontiguousarray(s[:n], dtype=np.float32))
                lens[i] = n
            feats = self.features_from_device_audio(host.to(self.device, non_blocking=True), lens.to(self.device))
        if return_tensors == "np":
            feats = feats.cpu().numpy()
        elif return_tensors == "pt_cuda":
            pass
        else:
            feats = feats.cpu()   # HF returns CPU tensors; the reference moves them later
        return BatchFeature(input_features=feats)

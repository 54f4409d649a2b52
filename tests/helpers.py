"""Shared test helpers (synthetic inputs per SURVEY.md section 8d)."""
import numpy as np


def synth_audio(idx: int, n: int = 480000) -> np.ndarray:
    """Seeded gaussian noise, sigma 0.1, 16 kHz (utterance `idx`)."""
    return (np.random.RandomState(1000 + idx).randn(n).astype(np.float32) * 0.1).astype(np.float32)

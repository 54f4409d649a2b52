"""Synthetic-input harness for the hot path: random-init HF Whisper of the
named sizes, seeded 30 s audio, the scheme swaps of the BASELINE.json configs, a stub processor
(no tokenizer files exist offline) and the utterance-sharded evaluation loop.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import dynamic, quanto, swap, tally
from .frontend import LogMelFrontend

# SURVEY.md Appendix B
WHISPER_SIZES = {
    "tiny":     dict(d_model=384,  heads=6,  ffn=1536, enc=4,  dec=4,  mels=80,  vocab=51865),
    "base":     dict(d_model=512,  heads=8,  ffn=2048, enc=6,  dec=6,  mels=80,  vocab=51865),
    "small":    dict(d_model=768,  heads=12, ffn=3072, enc=12, dec=12, mels=80,  vocab=51865),
    "medium":   dict(d_model=1024, heads=16, ffn=4096, enc=24, dec=24, mels=80,  vocab=51865),
    "large-v3": dict(d_model=1280, heads=20, ffn=5120, enc=32, dec=32, mels=128, vocab=51866),
}
SCHEMES = ("fp16", "llm_int8", "bnb_nf4", "bnb_nf4_direct", "quanto_int8", "quanto_int4", "dynamic_int8")


def whisper_config(size: str, **overrides):
    from transformers import WhisperConfig
    s = WHISPER_SIZES[size]
    kw = dict(vocab_size=s["vocab"], num_mel_bins=s["mels"], d_model=s["d_model"],
              encoder_layers=s["enc"], decoder_layers=s["dec"], encoder_attention_heads=s["heads"],
              decoder_attention_heads=s["heads"], encoder_ffn_dim=s["ffn"], decoder_ffn_dim=s["ffn"])
    kw.update(overrides)
    return WhisperConfig(**kw)


def build_model(size: str, seed: int = 0, device=None, **overrides):
    """Random-init (HF init, normal sigma 0.02) fp32 master copy, identical on every rank.  device: build (and
    draw the random weights) directly on that device -- seconds instead of a minute for medium / large-v3; the
    values then come from the device's generator (still identical across ranks of one GPU type)."""
    from transformers import WhisperForConditionalGeneration
    torch.manual_seed(seed)
    if device is not None:
        with torch.device(device):
            model = WhisperForConditionalGeneration(whisper_config(size, **overrides)).eval()
    else:
        model = WhisperForConditionalGeneration(whisper_config(size, **overrides)).eval()
    model.config.forced_decoder_ids = None
    return model


def synth_audio(idx: int, n: int = 480000) -> np.ndarray:
    """Utterance `idx`: seeded gaussian noise, sigma 0.1, 16 kHz."""
    return (np.random.RandomState(1000 + idx).randn(n).astype(np.float32) * 0.1).astype(np.float32)


def synth_reference(idx: int, n_words: int = 60) -> str:
    rng = np.random.RandomState(5000 + idx)
    return " ".join(f"t{v}" for v in rng.randint(0, 51000, size=n_words))


def global_l1_prune(model: nn.Module, amount: float) -> nn.Module:
    """prune.global_unstructured(all nn.Linear weights, L1Unstructured, amount) + prune.remove
    (pruning/baseline_scripts/unstructured_L1_baseline.py:500-502,525): zeros baked into .weight."""
    import torch.nn.utils.prune as prune
    params = [(m, "weight") for m in model.modules() if type(m) is nn.Linear]
    prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=amount)
    for m, name in params:
        prune.remove(m, name)
    return model


def apply_scheme(model: nn.Module, scheme: str, device, threshold: float = 6.0) -> nn.Module:
    """Swap the linears as the reference flow for `scheme` does and move the model to `device`."""
    device = torch.device(device)
    if scheme == "fp16":
        return model.half().to(device)
    if scheme == "llm_int8":      # HF load_in_8bit: fp16 model, proj_out kept fp16 (config 2)
        model = model.half()
        swap.replace_with_bnb_linear(model, load_in_8bit=True, llm_int8_threshold=threshold)
        return model.to(device)
    if scheme == "bnb_nf4":       # HF load_in_4bit nf4, fp16 compute, proj_out kept fp16 (config 3)
        model = model.half()
        swap.replace_with_bnb_linear(model, load_in_4bit=True, bnb_4bit_compute_dtype=torch.float16,
                                     bnb_4bit_quant_type="nf4")
        return model.to(device)
    if scheme == "bnb_nf4_direct":  # the reference's own convert_model_to_4bit: fp32, incl. proj_out
        swap.convert_model_to_4bit(model, compute_dtype=torch.float32, quant_type="nf4", double_quant=False)
        return model.to(device)
    if scheme == "quanto_int8":   # model_utils.py:126-137 order: quantize, freeze, then .to(device)
        quanto.quantize(model, weights=quanto.qint8)
        quanto.freeze(model)
        return model.to(device)
    if scheme == "quanto_int4":   # model_utils.py "quanto_int4": group-wise affine uint4 weights
        quanto.quantize(model, weights=quanto.qint4)
        quanto.freeze(model)
        return model.to(device)
    if scheme == "quanto_int8_fp16":
        model = model.half()
        quanto.quantize(model, weights=quanto.qint8)
        quanto.freeze(model)
        return model.to(device)
    if scheme == "dynamic_int8":  # GPU twin of torch quantize_dynamic (config 1)
        dynamic.quantize_dynamic(model, {nn.Linear}, dtype=torch.qint8, inplace=True)
        return model.to(device)
    raise ValueError(f"unknown scheme {scheme}")


class _StubTokenizer:
    def normalize(self, text: str) -> str:
        return text


class StubProcessor:
    """Processor stand-in: log-mel via the CUDA frontend, pseudo-word decode."""

    def __init__(self, n_mels: int = 80, device="cuda", chunk_length: int = 30):
        self.feature_extractor = LogMelFrontend(feature_size=n_mels, device=device, chunk_length=chunk_length)
        self.tokenizer = _StubTokenizer()

    def __call__(self, audio, sampling_rate=16000, return_tensors="pt", **kw):
        return self.feature_extractor(audio, sampling_rate=sampling_rate, return_tensors=return_tensors, **kw)

    _words: List[str] = []

    def _table(self, n: int) -> List[str]:
        if len(self._words) < n:
            StubProcessor._words = [f"t{i}" for i in range(max(n, 51866))]
        return self._words

    def decode(self, ids, **kw) -> str:
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        words = self._table(max(ids, default=0) + 1)
        return " ".join([words[i] for i in ids])

    _word_array = None

    def batch_decode(self, ids, **kw) -> List[str]:
        if isinstance(ids, torch.Tensor) and ids.dim() == 2 and ids.numel() > 0:
            # one fancy-index over an object array of the pseudo-words, then one join per row
            a = ids.detach().cpu().numpy()
            words = self._table(int(a.max()) + 1)
            if StubProcessor._word_array is None or len(StubProcessor._word_array) != len(words):
                StubProcessor._word_array = np.array(words, dtype=object)
            return [" ".join(r) for r in StubProcessor._word_array[a]]
        rows = ids.tolist() if hasattr(ids, "tolist") else [list(r) for r in ids]
        words = self._table(max((max(r, default=0) for r in rows), default=0) + 1)
        return [" ".join([words[i] for i in r]) for r in rows]


@torch.no_grad()
def greedy_generate(model, features: torch.Tensor, new_tokens: int) -> torch.Tensor:
    """model.generate exactly as transcribe_batch calls it (data_utils.py:152), with the decode
    length pinned (random weights never emit EOS): greedy, min = max = new_tokens."""
    return model.generate(features, do_sample=False, num_beams=1, min_new_tokens=new_tokens,
                          max_new_tokens=new_tokens)


def model_dtype(model) -> torch.dtype:
    return next(model.parameters()).dtype

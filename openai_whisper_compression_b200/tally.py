"""WER / CER tallies: the integer edit-distance core of ``evaluate.load("wer"/"cer").compute``
(evaluation.py:110-116; jiwer -> rapidfuzz Levenshtein) on the GPU, plus the cross-rank sum.

A tally is int64[4] = {word errors, reference words, char errors, reference chars}; corpus WER
= 100 * t[0] / t[1], CER = 100 * t[2] / t[3] -- the same corpus-level definition `evaluate` uses
(sum of S+D+I over sum of reference lengths).  Under torch.distributed (one process per GPU,
utterance-sharded) the only collective on the whole path is one all_reduce(SUM) of this tensor
over NCCL.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import functional as F


def _pack(seqs):
    """list of int32 arrays -> (concatenated ids, int64 offsets [P+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    flat = np.concatenate(seqs).astype(np.int32, copy=False) if len(seqs) else np.zeros((0,), np.int32)
    return flat, off


def _codepoints(text: str) -> np.ndarray:
    return np.frombuffer(text.encode("utf-32-le"), dtype=np.uint32).astype(np.int32)


def tally_on_device(references: Sequence[str], predictions: Sequence[str], device) -> torch.Tensor:
    """int64[4] tally on `device`; word ids are per-pair vocab indices, chars are code points."""
    if len(references) != len(predictions):
        raise ValueError("references and predictions differ in length")
    dev = torch.device(device)
    if len(references) == 0:
        return torch.zeros(4, dtype=torch.int64, device=dev)
    rw, hw, rc, hc = [], [], [], []
    vocab = {}                      # one id space for the whole batch (ids only need to be consistent)
    for ref, hyp in zip(references, predictions):
        rw.append(np.fromiter((vocab.setdefault(w, len(vocab)) for w in ref.split()), dtype=np.int32))
        hw.append(np.fromiter((vocab.setdefault(w, len(vocab)) for w in hyp.split()), dtype=np.int32))
        rc.append(_codepoints(ref))
        hc.append(_codepoints(hyp))
    out = torch.zeros(4, dtype=torch.int64, device=dev)
    for slot, (rs, hs) in enumerate(((rw, hw), (rc, hc))):
        r, ro = _pack(rs)
        h, ho = _pack(hs)
        if max(max((len(s) for s in rs), default=0), max((len(s) for s in hs), default=0)) > 4096:
            raise ValueError("sequence longer than 4096 ids")
        d = F.edit_distance(torch.from_numpy(r).to(dev), torch.from_numpy(ro).to(dev),
                            torch.from_numpy(h).to(dev), torch.from_numpy(ho).to(dev))
        out[2 * slot] = d.sum()
        out[2 * slot + 1] = int(ro[-1])
    return out


def all_reduce_tally(t: torch.Tensor) -> torch.Tensor:
    """SUM over ranks (NCCL on GPUs, gloo in the CPU tests); identity without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def rates(t: torch.Tensor) -> dict:
    v = [int(x) for x in t.tolist()]
    return {"WER": 100.0 * v[0] / max(v[1], 1), "CER": 100.0 * v[2] / max(v[3], 1)}


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous utterance shard of rank `rank`: [r*n/W, (r+1)*n/W)."""
    return range((rank * n) // world, ((rank + 1) * n) // world)


class _Metric:
    """Stand-in for evaluate.load("wer"|"cer"): .compute(references=, predictions=) -> fraction."""

    def __init__(self, name: str):
        if name not in ("wer", "cer"):
            raise ValueError(f"unknown metric {name}")
        self.name = name

    def compute(self, references: List[str], predictions: List[str]) -> float:
        dev = torch.device("cuda", torch.cuda.current_device())
        t = tally_on_device(references, predictions, dev).tolist()
        num, den = (t[0], t[1]) if self.name == "wer" else (t[2], t[3])
        return num / max(den, 1)


def load_metric(name: str, *a, **k) -> _Metric:
    return _Metric(name)

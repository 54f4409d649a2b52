"""WER / CER tallies: the integer edit-distance core of ``evaluate.load("wer"/"cer").compute``
(evaluation.py:110-116; jiwer -> rapidfuzz Levenshtein) on the GPU, plus the cross-rank sum.

A tally is int64[4] = {word errors, reference words, char errors, reference chars}; corpus WER
= 100 * t[0] / t[1], CER = 100 * t[2] / t[3] -- the same corpus-level definition `evaluate` uses
(sum of S+D+I over sum of reference lengths).  Under torch.distributed (one process per GPU,
utterance-sharded) the only collective on the whole path is one all_reduce(SUM) of this tensor
over NCCL.
"""
from __future__ import annotations

from itertools import chain, count
from typing import List, Sequence

import numpy as np
import torch

from . import functional as F


_REF_CACHE = {}


def _reference_side(references: Sequence[str]):
    """Word ids, vocabulary, code points and lengths of the references.  An evaluation scores the SAME references
    against new hypotheses batch after batch (evaluation.py:96-116), so this half is cached per list object."""
    key = (id(references), len(references))
    hit = _REF_CACHE.get(key)
    if hit is not None and hit[0] is references:
        return hit[1]
    rw = [r.split() for r in references]
    n_rw = sum(map(len, rw))
    vocab = {}
    words = np.fromiter(map(vocab.setdefault, chain.from_iterable(rw), count()), dtype=np.int32, count=n_rw)
    chars = np.frombuffer("".join(references).encode("utf-32-le"), dtype=np.int32)
    lens = (np.fromiter(map(len, rw), dtype=np.int64, count=len(rw)),
            np.fromiter(map(len, references), dtype=np.int64, count=len(references)))
    side = (words, vocab, n_rw, chars, lens)
    if len(_REF_CACHE) > 8:
        _REF_CACHE.clear()
    _REF_CACHE[key] = (references, side)
    return side


def _host_pack(references: Sequence[str], predictions: Sequence[str]):
    """Host side of a tally: 2P sequence pairs -- P (reference, hypothesis) word-id pairs, then P code-point
    pairs -- as (ids int32 [ref ids..., hyp ids...], offsets int64 [ref offsets (2P+1), hyp offsets (2P+1)],
    number of reference ids).  Word ids index one vocabulary for the whole batch (they only need to be
    consistent within a pair).  Written for the host cost to stay small next to a 100 ms GPU step: one pass of
    dict look-ups over the hypothesis words (the reference half is cached), one utf-32 encode of the text."""
    P = len(references)
    ref_words, ref_vocab, n_rw, ref_chars, (rw_len, rc_len) = _reference_side(references)
    hw = [h.split() for h in predictions]
    n_hw = sum(map(len, hw))
    # id of a word = position of its first occurrence (dict.setdefault driven by map(): no Python-level loop); the
    # hypotheses continue the references' vocabulary
    vocab = dict(ref_vocab)
    hyp_words = np.fromiter(map(vocab.setdefault, chain.from_iterable(hw), count(n_rw)), dtype=np.int32, count=n_hw)
    hyp_chars = np.frombuffer("".join(predictions).encode("utf-32-le"), dtype=np.int32)
    n_rc = int(ref_chars.shape[0])
    ref_len = np.concatenate([rw_len, rc_len])
    hyp_len = np.fromiter(chain(map(len, hw), map(len, predictions)), dtype=np.int64, count=2 * P)
    if max(int(ref_len.max(initial=0)), int(hyp_len.max(initial=0))) > 4096:
        raise ValueError("sequence longer than 4096 ids")
    ids = np.concatenate([ref_words, ref_chars, hyp_words, hyp_chars])
    off = np.zeros(2 * (2 * P + 1), dtype=np.int64)
    np.cumsum(ref_len, out=off[1:2 * P + 1])
    np.cumsum(hyp_len, out=off[2 * P + 2:])
    return ids, off, n_rw + n_rc, n_rw, n_rc


def _to_device(a: np.ndarray, dev, non_blocking: bool) -> torch.Tensor:
    t = torch.from_numpy(a)
    if non_blocking:
        # pinned staging + asynchronous copy: the host does not wait for the stream's earlier work (the caching host
        # allocator keeps the staging block alive until the copy has run)
        return t.pin_memory().to(dev, non_blocking=True)
    return t.to(dev)


def tally_on_device(references: Sequence[str], predictions: Sequence[str], device, non_blocking: bool = False) -> torch.Tensor:
    """int64[4] tally on `device`; word ids are per-batch vocab indices, chars are code points.  One edit-distance
    launch over the 2P pairs, two host->device copies.  non_blocking: nothing here waits for the device (a caller that
    overlaps the tally with other GPU work reads the result later)."""
    if len(references) != len(predictions):
        raise ValueError("references and predictions differ in length")
    dev = torch.device(device)
    P = len(references)
    if P == 0:
        return torch.zeros(4, dtype=torch.int64, device=dev)
    ids, off, n_ref, n_rw, n_rc = _host_pack(references, predictions)
    ids_d = _to_device(ids, dev, non_blocking)
    off_d = _to_device(off, dev, non_blocking)
    d = F.edit_distance(ids_d[:n_ref], off_d[:2 * P + 1], ids_d[n_ref:], off_d[2 * P + 1:])
    errs = d.view(2, P).sum(1)       # (the kernel's -1 sentinel for over-long pairs cannot occur: _host_pack refuses them)
    out = _to_device(np.array([0, n_rw, 0, n_rc], dtype=np.int64), dev, non_blocking)
    out[0::2] = errs
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
        # evaluate's wer / cer hand the strings to jiwer, whose default transforms strip the ends and collapse runs
        # of white space (wer: then split on spaces; cer: characters of the cleaned string), and jiwer refuses an
        # empty reference
        references = [" ".join(str(r).split()) for r in references]
        predictions = [" ".join(str(p).split()) for p in predictions]
        if any(len(r) == 0 for r in references):
            raise ValueError("one or more references are empty strings")
        dev = torch.device("cuda", torch.cuda.current_device())
        t = tally_on_device(references, predictions, dev).tolist()
        num, den = (t[0], t[1]) if self.name == "wer" else (t[2], t[3])
        return num / den


def load_metric(name: str, *a, **k) -> _Metric:
    return _Metric(name)

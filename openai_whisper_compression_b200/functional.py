"""Tensor-level wrappers over the C ABI (include/whisperq.h).

Names and argument meaning mirror the third-party routines the reference's module swaps reach
(bitsandbytes.functional.*, optimum.quanto quantizers, torch dynamic quantization), so that the
parity tests read like tests of those libraries.  PyTorch only provides device memory and the
stream here; every computation is a hand-written sm_100a kernel in libwhisperq.so.  There is no
CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_QT = {"nf4": 0, "fp4": 1}


class _Stats:
    """Kernel-launch accounting (bench.py's `gpu_launches`) and optional CUDA-event timing of the
    GEMM launches (bench.py's live roofline).  `launches` counts kernels of libwhisperq.so only."""

    def __init__(self):
        self.launches = 0
        self.profile_min_rows = None      # set to an int to time GEMM calls with M >= this
        self.records = []                 # (kind, M, N, K, start_event, end_event)

    def reset(self):
        self.launches = 0
        self.records = []


STATS = _Stats()
# Linear8bitLt calls with at most this many rows take the fused single-launch kernel (the kernel
# supports up to 64; every CTA re-quantizes all rows, which stops paying off beyond ~16 rows)
SMALL_M_ROWS = 16


class _Timed:
    def __init__(self, kind, M, N, K):
        self.on = STATS.profile_min_rows is not None and M >= STATS.profile_min_rows
        self.key = (kind, M, N, K)

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.e1.record()
            STATS.records.append((*self.key, self.e0, self.e1))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("openai_whisper_compression_b200 runs on CUDA (sm_100a) only; got a CPU tensor "
                               "and there is no CPU fallback")
        if not t.is_contiguous():
            raise RuntimeError("expected a contiguous tensor")


# ----------------------------------------------------------------------------------------------
# bitsandbytes 4-bit
# ----------------------------------------------------------------------------------------------
def quantize_4bit(w: torch.Tensor, blocksize: int = 64, quant_type: str = "nf4"
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """bitsandbytes.functional.quantize_4bit -> (packed uint8 [(n+1)//2, 1], absmax f32 [n/bs])."""
    w = w.contiguous()
    _need_cuda(w)
    n = w.numel()
    packed = torch.empty(((n + 1) // 2, 1), dtype=torch.uint8, device=w.device)
    absmax = torch.empty(((n + blocksize - 1) // blocksize,), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().wq_quant_4bit(_ptr(w), _DT[w.dtype], n, blocksize, _QT[quant_type], _ptr(packed),
                                             _ptr(absmax), _stream()), "wq_quant_4bit")
    STATS.launches += 1
    return packed, absmax


def dequantize_4bit(packed: torch.Tensor, absmax: torch.Tensor, shape, blocksize: int = 64,
                    quant_type: str = "nf4", dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """bitsandbytes.functional.dequantize_4bit."""
    _need_cuda(packed, absmax)
    out = torch.empty(tuple(shape), dtype=dtype, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(_lib.load().wq_dequant_4bit(_ptr(packed), _ptr(absmax), out.numel(), blocksize, _QT[quant_type],
                                               _ptr(out), _DT[dtype], _stream()), "wq_dequant_4bit")
    STATS.launches += 1
    return out


_DYNAMIC_MAP = {}


def dynamic_map(device) -> torch.Tensor:
    """bitsandbytes create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): the 256-entry
    code book of blockwise 8-bit quantization (a constant table, built once per device with the
    library's own formula)."""
    dev = torch.device(device)
    t = _DYNAMIC_MAP.get(dev)
    if t is None:
        data = []
        for i in range(7):
            boundaries = torch.linspace(0.1, 1, 2 ** i + 1)
            means = (boundaries[:-1] + boundaries[1:]) / 2.0
            data += ((10 ** (-6 + i)) * means).tolist()
            data += (-(10 ** (-6 + i)) * means).tolist()
        data += [0, 1.0]
        data.sort()
        t = _DYNAMIC_MAP[dev] = torch.tensor(data, dtype=torch.float32, device=dev)
    return t


def quantize_absmax_double(absmax: torch.Tensor):
    """Nested quantization of the 4-bit statistics (quantize_4bit(compress_statistics=True)).

    Returns (q uint8 [n], absmax2 f32 [ceil(n/256)], offset f32 [1], absmax_deq f32 [n])."""
    absmax = absmax.contiguous()
    _need_cuda(absmax)
    n = absmax.numel()
    code = dynamic_map(absmax.device)
    q = torch.empty((n,), dtype=torch.uint8, device=absmax.device)
    a2 = torch.empty(((n + 255) // 256,), dtype=torch.float32, device=absmax.device)
    off = torch.empty((1,), dtype=torch.float32, device=absmax.device)
    deq = torch.empty((n,), dtype=torch.float32, device=absmax.device)
    with torch.cuda.device(absmax.device):
        _lib.check(_lib.load().wq_quant_absmax_double(_ptr(absmax), n, _ptr(code), _ptr(q), _ptr(a2), _ptr(off),
                                                      _ptr(deq), _stream()), "wq_quant_absmax_double")
    STATS.launches += 2
    return q, a2, off, deq


def dequantize_absmax_double(q: torch.Tensor, absmax2: torch.Tensor, offset: torch.Tensor) -> torch.Tensor:
    """dequantize_blockwise(absmax, state2) + offset -> fp32 statistics."""
    _need_cuda(q, absmax2, offset)
    n = q.numel()
    out = torch.empty((n,), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().wq_dequant_absmax_double(_ptr(q), _ptr(absmax2), _ptr(dynamic_map(q.device)),
                                                        _ptr(offset), n, _ptr(out), _stream()),
                   "wq_dequant_absmax_double")
    STATS.launches += 1
    return out


# decode-shaped calls of the weight-only schemes with at most this many rows stream the packed weights through the
# CUDA-core GEMV (gemv_wq.cu) instead of the tcgen05 pipeline (WQ_GEMV_ROWS=0 disables: A/B measurements)
import os as _os
GEMV_ROWS = int(_os.environ.get("WQ_GEMV_ROWS", "32"))


def _gemv_weightonly(x2: torch.Tensor, mode: int, w: torch.Tensor, s0: torch.Tensor, s1, group: int, quant_type: int,
                     bias, y: torch.Tensor, N: int, K: int) -> None:
    with torch.cuda.device(x2.device):
        _lib.check(_lib.load().wq_gemv_weightonly(_ptr(x2), _DT[x2.dtype], x2.shape[0], K, mode, _ptr(w), _ptr(s0), _ptr(s1),
                                                  group, quant_type, _ptr(bias), _ptr(y), N, _stream()),
                   "wq_gemv_weightonly")
    STATS.launches += 1


def _gemv_ok(x2: torch.Tensor, out_dtype) -> bool:
    return 0 < x2.shape[0] <= GEMV_ROWS and out_dtype == x2.dtype and x2.dtype in (torch.float16, torch.bfloat16)


def _gemv_f32_ok(x2: torch.Tensor, out_dtype) -> bool:
    """fp32 rows of the reference's fp32 flows (the reference evaluates in batches of 16, quantization.py:33): one GEMV
    launch replaces the fp32 -> fp16 cast pass + tensor-core tile.  Only up to GEMV_ROWS rows: splitting 64 rows into
    two launches measured SLOWER than cast + GEMM (whisper-medium quanto, 64 utterances: 8.9 vs 7.1 ms per token --
    at 32 rows the CUDA-core FMAs, not the weight bytes, bound the GEMV)."""
    return 0 < x2.shape[0] <= GEMV_ROWS and out_dtype == torch.float32 and x2.dtype == torch.float32


def _gemv_f32(x2: torch.Tensor, mode: int, w, s0, s1, group: int, quant_type: int, bias, y: torch.Tensor, N: int, K: int):
    _gemv_weightonly(x2, mode, w, s0, s1, group, quant_type, bias, y, N, K)


def _operand(x2: torch.Tensor) -> torch.Tensor:
    """fp32 activations go to the tensor cores as fp16 (DESIGN.md "Numerics"): the one cast pass of the fp32 flows."""
    return x2.to(torch.float16) if x2.dtype == torch.float32 else x2


def _dest(out: Optional[torch.Tensor], M: int, N: int, dtype, device, what: str) -> torch.Tensor:
    """[M, N] destination of a GEMM: a fresh tensor, or the caller's contiguous buffer of M*N elements."""
    if out is None:
        return torch.empty((M, N), dtype=dtype, device=device)
    if out.dtype != dtype or out.numel() != M * N or not out.is_contiguous() or not out.is_cuda:
        raise RuntimeError(f"{what}: out must be a contiguous CUDA tensor of M*N elements in the output dtype")
    return out.view(M, N)


def gemm_w4a16(x: torch.Tensor, packed: torch.Tensor, absmax: torch.Tensor, N: int, K: int,
               bias: Optional[torch.Tensor] = None, quant_type: str = "nf4",
               out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Linear4bit.forward hot path: y = x @ dequantize_4bit(W).T + bias, fused (blocksize 64)."""
    x2 = x.reshape(-1, K).contiguous()
    _need_cuda(x2, packed, absmax, bias)
    out_dtype = out_dtype or x2.dtype
    y = _dest(out, x2.shape[0], N, out_dtype, x.device, "gemm_w4a16")
    if _gemv_ok(x2, out_dtype):
        _gemv_weightonly(x2, 0, packed, absmax, None, 0, _QT[quant_type], bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    if _gemv_f32_ok(x2, out_dtype):
        _gemv_f32(x2, 0, packed, absmax, None, 0, _QT[quant_type], bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    x2 = _operand(x2)
    with torch.cuda.device(x.device), _Timed("w4a16", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_w4a16(_ptr(x2), _DT[x2.dtype], _ptr(packed), _ptr(absmax), _QT[quant_type],
                                             _ptr(bias), _ptr(y), _DT[out_dtype], x2.shape[0], N, K, _stream()),
                   "wq_gemm_w4a16")
    STATS.launches += 1
    return y.reshape(*x.shape[:-1], N)


# ----------------------------------------------------------------------------------------------
# bitsandbytes LLM.int8
# ----------------------------------------------------------------------------------------------
_SLOT = 0


class scratch_slot:
    """Context manager selecting which copy of the per-device scratch (outlier flags, row counters) the wrappers
    below hand to the kernels.  The flags are written by a producer kernel and consumed / cleared by the GEMM that
    follows it on the same stream; two streams working on two half-batches at the same time (fastgen's two-stream
    decode step) therefore need a copy each."""

    def __init__(self, slot: int):
        self.slot = int(slot)

    def __enter__(self):
        global _SLOT
        self.prev, _SLOT = _SLOT, self.slot
        return self

    def __exit__(self, *exc):
        global _SLOT
        _SLOT = self.prev
        return False


class OutlierState:
    """Per-device (and per scratch slot) buffers for the outlier bookkeeping (flags are self-cleaning)."""

    _cache = {}

    def __init__(self, device: torch.device, cols: int):
        self.col_flags = torch.zeros((cols + 2,), dtype=torch.int32, device=device)   # flags, any, counter
        self.outlier_cols = torch.empty((cols,), dtype=torch.int32, device=device)
        self.n_outliers = torch.zeros((1,), dtype=torch.int32, device=device)

    @classmethod
    def get(cls, device: torch.device, cols: int) -> "OutlierState":
        key = (device.index, cols, _SLOT)
        st = cls._cache.get(key)
        if st is None:
            st = cls._cache[key] = cls(device, cols)
        return st


def int8_vectorwise_quant(a: torch.Tensor, threshold: float = 0.0, state: Optional[OutlierState] = None,
                          finalize: bool = True):
    """bitsandbytes.functional.int8_vectorwise_quant.

    Returns (CA int8, row_stats f32, state).  With threshold > 0 the outlier columns stay on the
    device -- no host synchronisation.  finalize=True reproduces the library's outputs exactly
    (CA[:, outlier_cols] zeroed, state.outlier_cols[: state.n_outliers] filled, flags cleared);
    finalize=False leaves the raw flags in state.col_flags for gemm_llmint8 to consume (the module
    forward path: two kernel launches per linear)."""
    a2 = a.reshape(-1, a.shape[-1])
    if a2.dtype != torch.float16:
        a2 = a2.to(torch.float16)
    a2 = a2.contiguous()
    _need_cuda(a2)
    rows, cols = a2.shape
    ca = torch.empty((rows, cols), dtype=torch.int8, device=a.device)
    stats = torch.empty((rows,), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        if threshold > 0.0:
            state = state or OutlierState.get(a.device, cols)
            _lib.check(lib.wq_quant_i8_rowwise_bnb(_ptr(a2), rows, cols, float(threshold), _ptr(ca), _ptr(stats),
                                                   _ptr(state.col_flags), _stream()), "wq_quant_i8_rowwise_bnb")
            STATS.launches += 1
            if finalize:
                _lib.check(lib.wq_outlier_columns(_ptr(state.col_flags), rows, cols, _ptr(ca),
                                                  _ptr(state.outlier_cols), _ptr(state.n_outliers), _stream()),
                           "wq_outlier_columns")
                STATS.launches += 2
        else:
            state = None
            _lib.check(lib.wq_quant_i8_rowwise_bnb(_ptr(a2), rows, cols, 0.0, _ptr(ca), _ptr(stats), None,
                                                   _stream()), "wq_quant_i8_rowwise_bnb")
            STATS.launches += 1
    return ca.reshape(a.shape), stats, state


def gemm_llmint8(ca: torch.Tensor, sca: torch.Tensor, cb: torch.Tensor, scb: torch.Tensor,
                 bias: Optional[torch.Tensor] = None, a_f16: Optional[torch.Tensor] = None,
                 state: Optional[OutlierState] = None, out: Optional[torch.Tensor] = None,
                 keep_flags: bool = False, residual: Optional[torch.Tensor] = None,
                 clamp_abs: float = 0.0, a_pre_gelu: bool = False) -> torch.Tensor:
    """int8_linear_matmul + int8_mm_dequant (+ mixed-precision outlier decomposition when `state`
    carries raw flags from int8_vectorwise_quant(..., finalize=False)), one kernel; fp16 [M, N].
    out: optional contiguous fp16 [M, N] destination.  keep_flags: leave the outlier flags set because
    another GEMM consumes the same quantized rows next (the last consumer clears them).  residual / clamp_abs: the
    layer's residual connection in the epilogue, y = clamp(fp16(linear) + residual) (fc2 of a Whisper layer).
    a_pre_gelu: the int8 rows came from gelu_quant(..., store_h=False) and a_f16 is the tensor before the GELU."""
    ca2 = ca.reshape(-1, ca.shape[-1])
    _need_cuda(ca2, sca, cb, scb, bias, a_f16)
    M, K = ca2.shape
    N = cb.shape[0]
    if out is None:
        y = torch.empty((M, N), dtype=torch.float16, device=ca.device)
    else:
        if out.dtype != torch.float16 or out.numel() != M * N or not out.is_contiguous() or not out.is_cuda:
            raise RuntimeError("gemm_llmint8: out must be a contiguous CUDA fp16 tensor with M*N elements")
        y = out.view(M, N)
    if bias is not None and bias.dtype != torch.float32:
        if bias.dtype != torch.float16:
            raise RuntimeError("gemm_llmint8: bias must be fp16 (or its exact fp32 widening)")
        bias = bias.float()          # exact; modules pass a cached fp32 copy instead
    with torch.cuda.device(ca.device), _Timed("llmint8+res" if residual is not None else "llmint8", M, N, K):
        if residual is not None:
            residual = residual.reshape(M, N)
            _need_cuda(residual)
            if residual.dtype != torch.float16:
                raise RuntimeError("gemm_llmint8: the residual must be fp16 [M, N]")
        _lib.check(_lib.load().wq_gemm_llmint8_residual(
            _ptr(ca2), _ptr(sca), _ptr(cb), _ptr(scb), _ptr(bias), _ptr(y), M, N, K,
            _ptr(a_f16) if state is not None else None,
            _ptr(state.col_flags) if state is not None else None, 1 if keep_flags else 0, _ptr(residual),
            float(clamp_abs), 1 if a_pre_gelu else 0, _stream()), "wq_gemm_llmint8")
    STATS.launches += 1
    return y


def linear8bitlt(x: torch.Tensor, cb: torch.Tensor, scb: torch.Tensor, bias: Optional[torch.Tensor],
                 threshold: float) -> torch.Tensor:
    """bnb.matmul(x, Int8Params, state) for has_fp16_weights=False: quantize + fused GEMM."""
    a = x.reshape(-1, x.shape[-1])
    if a.dtype != torch.float16:
        a = a.to(torch.float16)
    a = a.contiguous()
    M, K = a.shape
    if 0 < M <= SMALL_M_ROWS and K % 16 == 0 and 64 * K + K + 272 <= 200 * 1024:
        # decode-shaped call: quantize + int8 GEMV + dequant (+ outliers) in one launch
        _need_cuda(a, cb, scb, bias)
        if bias is not None and bias.dtype != torch.float32:
            bias = bias.float()
        N = cb.shape[0]
        y = torch.empty((M, N), dtype=torch.float16, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().wq_linear_llmint8_small(_ptr(a), M, K, float(threshold), _ptr(cb), _ptr(scb),
                                                           _ptr(bias), _ptr(y), N, _stream()),
                       "wq_linear_llmint8_small")
        STATS.launches += 1
        return y.reshape(*x.shape[:-1], N).to(x.dtype)
    ca, sca, st = int8_vectorwise_quant(a, threshold, finalize=False)
    y = gemm_llmint8(ca, sca, cb, scb, bias, a if st is not None else None, st)
    return y.reshape(*x.shape[:-1], cb.shape[0]).to(x.dtype)


# ----------------------------------------------------------------------------------------------
# producers fused with the LLM.int8 quantizer (rowops.cu, attn_decode.cu)
# ----------------------------------------------------------------------------------------------
def _quant_outputs(rows: int, cols: int, device, threshold: Optional[float]):
    """(ca, sca, state) buffers for a fused producer; all None when threshold is None (no quantization)."""
    if threshold is None:
        return None, None, None
    ca = torch.empty((rows, cols), dtype=torch.int8, device=device)
    sca = torch.empty((rows,), dtype=torch.float32, device=device)
    state = OutlierState.get(device, cols) if threshold > 0.0 else None
    return ca, sca, state


def add_layernorm_quant(x: torch.Tensor, delta: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor,
                        eps: float, threshold: Optional[float] = None, h_out: Optional[torch.Tensor] = None):
    """x' = x + delta (delta may be None), h = layer_norm(x') and, with a threshold, the Linear8bitLt row
    quantization of h -- one launch.  Returns (x', h, (ca, sca, state) | None); shapes follow x.  h_out: optional
    contiguous destination for h (same shape and dtype as x)."""
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    d2 = None if delta is None else delta.reshape(-1, cols)
    _need_cuda(x2, d2, weight, bias)
    rows = x2.shape[0]
    x_out = torch.empty_like(x2) if d2 is not None else x2
    if h_out is None:
        h = torch.empty_like(x2)
    else:
        if h_out.shape != x2.shape or h_out.dtype != x2.dtype or not h_out.is_contiguous():
            raise RuntimeError("add_layernorm_quant: h_out must be a contiguous tensor shaped like x")
        h = h_out
    ca, sca, state = _quant_outputs(rows, cols, x.device, threshold)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().wq_add_layernorm_quant(
            _ptr(x2), _ptr(d2), _DT[x2.dtype], _ptr(weight), _ptr(bias), float(eps), rows, cols,
            _ptr(x_out) if d2 is not None else None, _ptr(h), float(threshold or 0.0), _ptr(ca), _ptr(sca),
            _ptr(state.col_flags) if state is not None else None, _stream()), "wq_add_layernorm_quant")
    STATS.launches += 1
    quant = None if threshold is None else (ca, sca, state)
    return x_out.view(x.shape), h.view(x.shape), quant


def gelu_quant(x: torch.Tensor, threshold: Optional[float] = None, store_h: bool = True):
    """h = gelu(x) (erf form) and, with a threshold, the Linear8bitLt row quantization of h -- one launch.
    Returns (h, (ca, sca, state) | None).  store_h=False (needs a threshold): only the int8 rows are written and h is
    None -- the consuming gemm_llmint8 takes x itself with a_pre_gelu=True (2 of 5 bytes per element less HBM traffic)."""
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    _need_cuda(x2)
    rows = x2.shape[0]
    if not store_h and threshold is None:
        raise RuntimeError("gelu_quant: store_h=False without a threshold would produce nothing")
    h = torch.empty_like(x2) if store_h else None
    ca, sca, state = _quant_outputs(rows, cols, x.device, threshold)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().wq_gelu_quant(_ptr(x2), _DT[x2.dtype], rows, cols, _ptr(h), float(threshold or 0.0),
                                             _ptr(ca), _ptr(sca),
                                             _ptr(state.col_flags) if state is not None else None, _stream()),
                   "wq_gelu_quant")
    STATS.launches += 1
    return (h.view(x.shape) if h is not None else None), (None if threshold is None else (ca, sca, state))


def self_attn_decode(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scaling: float, k_cache: torch.Tensor,
                     v_cache: torch.Tensor, pos: torch.Tensor, num_heads: int, threshold: Optional[float] = None):
    """One decode step of WhisperAttention self-attention: append k/v at `pos` (device int64 scalar) to the
    [B, t_max, H*64] caches and attend over positions 0..pos.  q/k/v: [B, H*64] rows with a common row
    stride (column blocks of a fused projection are fine).  Returns (out [B, H*64], (ca, sca, state) | None)."""
    B, d = q.shape
    ld = q.stride(0)
    if not (k.shape == q.shape == v.shape and k.stride(0) == ld == v.stride(0)
            and q.stride(1) == k.stride(1) == v.stride(1) == 1):
        raise RuntimeError("self_attn_decode: q, k, v must be [B, H*64] with unit column stride and one row stride")
    if d != num_heads * 64 or k_cache.shape != v_cache.shape or k_cache.shape[0] != B or k_cache.shape[2] != d:
        raise RuntimeError("self_attn_decode: caches must be [B, t_max, H*64] (head_dim 64)")
    if pos.dtype != torch.int64 or not pos.is_cuda:
        raise RuntimeError("self_attn_decode: pos must be a CUDA int64 scalar tensor")
    _need_cuda(k_cache, v_cache)
    out = torch.empty((B, d), dtype=q.dtype, device=q.device)
    ca, sca, state = _quant_outputs(B, d, q.device, threshold)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().wq_self_attn_decode(
            _ptr(q), _ptr(k), _ptr(v), ld, _DT[q.dtype], float(scaling), _ptr(k_cache), _ptr(v_cache), B,
            num_heads, k_cache.shape[1], _ptr(pos), _ptr(out), float(threshold or 0.0), _ptr(ca), _ptr(sca),
            _ptr(state.col_flags) if state is not None else None, _stream()), "wq_self_attn_decode")
    STATS.launches += 1
    return out, (None if threshold is None else (ca, sca, state))


_ROW_COUNTERS = {}
_ROW_COUNTERS_KEEP = []      # outgrown buffers stay alive: captured CUDA graphs may still point at them


def _row_counters(device, rows: int) -> torch.Tensor:
    """Zeroed int32 scratch (self-resetting inside the kernels that use it), one per device, grown on demand."""
    key = (device, _SLOT)
    t = _ROW_COUNTERS.get(key)
    if t is None or t.numel() < rows:
        if t is not None:
            _ROW_COUNTERS_KEEP.append(t)
        t = _ROW_COUNTERS[key] = torch.zeros((max(rows, 1024),), dtype=torch.int32, device=device)
    return t


def cross_attn_decode(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scaling: float, num_heads: int,
                      threshold: Optional[float] = None):
    """One decode step of WhisperAttention cross-attention over cached encoder K/V.  q: [B, H*64] (unit column
    stride); k, v: [B, S, H*64] views with unit column stride, dense in S (stride(0) == S * stride(1)) and one
    common row stride.  Returns (out [B, H*64], (ca, sca, state) | None) like the other fused producers."""
    B, d = q.shape
    if k.shape != v.shape or k.dim() != 3 or k.shape[0] != B or k.shape[2] != d or d != num_heads * 64:
        raise RuntimeError("cross_attn_decode: k, v must be [B, S, H*64] (head_dim 64) matching q [B, H*64]")
    S, ld = k.shape[1], k.stride(1)
    if not (q.stride(1) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and v.stride(1) == ld
            and k.stride(0) == S * ld and v.stride(0) == S * ld):
        raise RuntimeError("cross_attn_decode: unsupported strides")
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("cross_attn_decode: CUDA tensors only (no CPU fallback)")
    out = torch.empty((B, d), dtype=q.dtype, device=q.device)
    ca, sca, state = _quant_outputs(B, d, q.device, threshold)
    counters = _row_counters(q.device, B) if threshold is not None else None
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().wq_cross_attn_decode(
            _ptr(q), q.stride(0) if B > 1 else d, _DT[q.dtype], float(scaling), _ptr(k), _ptr(v), ld, B, S, num_heads,
            _ptr(out), float(threshold or 0.0), _ptr(ca), _ptr(sca),
            _ptr(state.col_flags) if state is not None else None, _ptr(counters), _stream()), "wq_cross_attn_decode")
    STATS.launches += 1
    return out, (None if threshold is None else (ca, sca, state))


def decode_fused_llmint8(after, before, args, run_after: bool, run_before: bool, run_final: bool, max_ctas: int) -> None:
    """One launch of the persistent decoder-layer kernel (decode_fused.cu).  after / before: _lib.DecodeLayer or None;
    args: _lib.DecodeArgs (the caller keeps every tensor behind the pointers alive)."""
    import ctypes
    dev = torch.cuda.current_device()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().wq_decode_fused_llmint8(
            None if after is None else ctypes.byref(after), None if before is None else ctypes.byref(before),
            ctypes.byref(args), int(run_after), int(run_before), int(run_final), int(max_ctas), _stream()),
            "wq_decode_fused_llmint8")
    STATS.launches += 1


def masked_argmax(logits: torch.Tensor, mask: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """argmax over the last dim of [B, V] fp16/bf16 logits (unit column stride, any 16-byte aligned row
    stride) with the columns where `mask` (bool [V]) is set treated as -inf; torch.argmax tie/NaN rules."""
    B, V = logits.shape
    if logits.stride(1) != 1 or not logits.is_cuda:
        raise RuntimeError("masked_argmax: logits must be a CUDA [B, V] tensor with unit column stride")
    if mask is not None:
        if mask.dtype not in (torch.bool, torch.uint8) or mask.numel() != V:
            raise RuntimeError("masked_argmax: mask must be bool/uint8 [V]")
        _need_cuda(mask)
    if out is None:
        out = torch.empty((B,), dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        ld = logits.stride(0) if B > 1 else -(-V // 8) * 8      # a single row has no meaningful stride
        _lib.check(_lib.load().wq_masked_argmax(_ptr(logits), _DT[logits.dtype], B, V, ld, _ptr(mask), _ptr(out),
                                                _stream()), "wq_masked_argmax")
    STATS.launches += 1
    return out


def gemm_f16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, argmax_keys: Optional[torch.Tensor] = None,
             mask: Optional[torch.Tensor] = None, store: bool = True) -> Optional[torch.Tensor]:
    """Unquantized linear y = x @ w.T + bias on the tcgen05 pipeline (the fp16 / bf16 `proj_out` of the HF
    bitsandbytes flows).  out: optional [M, >= N] destination whose first N columns are written (row stride kept).
    argmax_keys: uint64-as-int64 [M] zero-initialised keys that receive the masked arg-max of every row (see
    argmax_finalize); mask: bool / uint8, padded to whole 128-column tiles.  store=False with argmax_keys: the
    logits are not written at all."""
    N, K = w.shape
    x2 = x.reshape(-1, K)
    _need_cuda(x2, w, bias, argmax_keys, mask)
    if x2.dtype not in (torch.float16, torch.bfloat16) or w.dtype != x2.dtype:
        raise RuntimeError("gemm_f16: x and w must both be float16 or bfloat16")
    M = x2.shape[0]
    y = None
    ldy = 0
    if store:
        if out is None:
            y = torch.empty((M, N), dtype=x2.dtype, device=x.device)
        else:
            if out.dim() != 2 or out.shape[0] != M or out.shape[1] < N or out.stride(1) != 1 or out.dtype != x2.dtype:
                raise RuntimeError("gemm_f16: out must be [M, >= N] with unit column stride in the dtype of x")
            y, ldy = out, (out.stride(0) if M > 1 else out.shape[1])
    elif argmax_keys is None:
        raise RuntimeError("gemm_f16: store=False needs argmax_keys")
    if argmax_keys is not None and (argmax_keys.dtype != torch.int64 or argmax_keys.numel() < M):
        raise RuntimeError("gemm_f16: argmax_keys must be int64 [M]")
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    with torch.cuda.device(x.device), _Timed("f16", M, N, K):
        _lib.check(_lib.load().wq_gemm_f16(_ptr(x2), _DT[x2.dtype], _ptr(w), _ptr(bias), _ptr(y), _DT[x2.dtype], ldy,
                                           M, N, K, _ptr(mask), 0 if mask is None else mask.numel(),
                                           _ptr(argmax_keys), _stream()), "wq_gemm_f16")
    STATS.launches += 1
    return y


def argmax_finalize(keys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Token ids from the arg-max keys gemm_f16 accumulated; zeroes the keys for the next projection."""
    _need_cuda(keys, out)
    M = keys.numel()
    if out is None:
        out = torch.empty((M,), dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.load().wq_argmax_finalize(_ptr(keys), M, _ptr(out), _stream()), "wq_argmax_finalize")
    STATS.launches += 1
    return out


# ----------------------------------------------------------------------------------------------
# optimum-quanto qint8
# ----------------------------------------------------------------------------------------------
def quanto_quantize_qint8(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """quanto AbsmaxOptimizer + SymmetricQuantizer (axis 0): (int8 [N, K], scale f32 [N, 1])."""
    w = w.contiguous()
    _need_cuda(w)
    N, K = w.shape
    q = torch.empty((N, K), dtype=torch.int8, device=w.device)
    scale = torch.empty((N, 1), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().wq_quant_i8_rowwise_quanto(_ptr(w), _DT[w.dtype], N, K, _ptr(q), _ptr(scale),
                                                          _stream()), "wq_quant_i8_rowwise_quanto")
    STATS.launches += 1
    return q, scale


def gemm_w8a16(x: torch.Tensor, wq: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
               out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QLinear.forward hot path: y = (x @ Wq.T) * scale + bias with fp32 accumulation."""
    N, K = wq.shape
    x2 = x.reshape(-1, K).contiguous()
    _need_cuda(x2, wq, scale, bias)
    out_dtype = out_dtype or x2.dtype
    y = _dest(out, x2.shape[0], N, out_dtype, x.device, "gemm_w8a16")
    if _gemv_ok(x2, out_dtype) and K % 8 == 0:
        _gemv_weightonly(x2, 1, wq, scale, None, 0, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    if _gemv_f32_ok(x2, out_dtype) and K % 8 == 0:
        _gemv_f32(x2, 1, wq, scale, None, 0, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    x2 = _operand(x2)
    with torch.cuda.device(x.device), _Timed("w8a16", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_w8a16(_ptr(x2), _DT[x2.dtype], _ptr(wq), _ptr(scale), _ptr(bias), _ptr(y),
                                             _DT[out_dtype], x2.shape[0], N, K, _stream()), "wq_gemm_w8a16")
    STATS.launches += 1
    return y.reshape(*x.shape[:-1], N)


def quanto_quantize_qfloat8(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """quanto AbsmaxOptimizer + SymmetricQuantizer for weights=qfloat8 (e4m3fn, axis 0): (uint8 codes [N, K], scale
    f32 [N, 1] = absmax / 448 rounded to the weight dtype)."""
    w = w.contiguous()
    _need_cuda(w)
    N, K = w.shape
    q = torch.empty((N, K), dtype=torch.uint8, device=w.device)
    scale = torch.empty((N, 1), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().wq_quant_f8_rowwise_quanto(_ptr(w), _DT[w.dtype], N, K, _ptr(q), _ptr(scale), _stream()),
                   "wq_quant_f8_rowwise_quanto")
    STATS.launches += 1
    return q, scale


def gemm_wf8a16(x: torch.Tensor, wq: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QLinear.forward with qfloat8 weights: y = (x @ e4m3(Wq).T) * scale + bias with fp32 accumulation."""
    N, K = wq.shape
    x2 = x.reshape(-1, K).contiguous()
    _need_cuda(x2, wq, scale, bias)
    out_dtype = out_dtype or x2.dtype
    y = _dest(out, x2.shape[0], N, out_dtype, x.device, "gemm_wf8a16")
    if _gemv_ok(x2, out_dtype) and K % 8 == 0:
        _gemv_weightonly(x2, 3, wq, scale, None, 0, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    if _gemv_f32_ok(x2, out_dtype) and K % 8 == 0:
        _gemv_f32(x2, 3, wq, scale, None, 0, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    x2 = _operand(x2)
    with torch.cuda.device(x.device), _Timed("wf8a16", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_wf8a16(_ptr(x2), _DT[x2.dtype], _ptr(wq), _ptr(scale), _ptr(bias), _ptr(y),
                                              _DT[out_dtype], x2.shape[0], N, K, _stream()), "wq_gemm_wf8a16")
    STATS.launches += 1
    return y.reshape(*x.shape[:-1], N)


_ACT_QT = {"qint8": 0, "qfloat8": 1, "qfloat8_e4m3fn": 1}


def quant_act_static(x: torch.Tensor, scale: torch.Tensor, qtype: str, codes: bool = False, grid: bool = False,
                     deq: bool = False):
    """quanto quantize_activation(x, qtype, scale) with a calibrated per-tensor scale (CUDA fp32 scalar tensor).
    Returns (int8 codes | None, fp16 code values | None, dequantized x in its own dtype | None)."""
    x = x.contiguous()
    _need_cuda(x, scale)
    if scale.dtype != torch.float32 or scale.numel() != 1:
        raise RuntimeError("quant_act_static: scale must be a float32 scalar tensor on the device")
    qt = _ACT_QT[qtype]
    c = torch.empty(x.shape, dtype=torch.int8, device=x.device) if codes else None
    g = torch.empty(x.shape, dtype=torch.float16, device=x.device) if grid else None
    d = torch.empty_like(x) if deq else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().wq_quant_act_static(_ptr(x), _DT[x.dtype], x.numel(), _ptr(scale), qt, _ptr(c), _ptr(g),
                                                   _ptr(d), _stream()), "wq_quant_act_static")
    STATS.launches += 1
    return c, g, d


def gemm_w8a8(xq: torch.Tensor, wq: torch.Tensor, out_scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
              out_dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """qint8 activations x qint8 weights (quanto qbytes_int_mm): y = float(int32 acc) * out_scale[n] + bias[n]."""
    N, K = wq.shape
    x2 = xq.reshape(-1, K)
    _need_cuda(x2, wq, out_scale, bias)
    y = torch.empty((x2.shape[0], N), dtype=out_dtype, device=xq.device)
    with torch.cuda.device(xq.device), _Timed("w8a8", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_w8a8(_ptr(x2), _ptr(wq), _ptr(out_scale), _ptr(bias), _ptr(y), _DT[out_dtype],
                                            x2.shape[0], N, K, _stream()), "wq_gemm_w8a8")
    STATS.launches += 1
    return y.reshape(*xq.shape[:-1], N)


def quanto_group_size(in_features: int) -> int:
    """optimum.quanto QModuleMixin group size for qint4/qint2 weights."""
    g = 128
    if in_features > g:
        while in_features % g != 0 and g > 32:
            g -= 32
        if in_features % g == 0:
            return g
    return in_features


def quanto_quantize_qint4(w: torch.Tensor, group: Optional[int] = None, bits: int = 4):
    """quanto MaxOptimizer + AffineQuantizer: (packed uint8 [N, K/2], scale f32 [N, K/g],
    shift f32 [N, K/g], group).  bits = 2 gives quanto's qint2 codes (0..3) in the same one-code-per-nibble
    container, so gemm_u4a16 consumes either."""
    if bits not in (2, 4):
        raise ValueError("bits must be 2 or 4")
    w = w.contiguous()
    _need_cuda(w)
    N, K = w.shape
    g = group or quanto_group_size(K)
    packed = torch.empty((N, K // 2), dtype=torch.uint8, device=w.device)
    scale = torch.empty((N, K // g), dtype=torch.float32, device=w.device)
    shift = torch.empty((N, K // g), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().wq_quant_ubits_group_quanto(_ptr(w), _DT[w.dtype], N, K, g, bits, _ptr(packed),
                                                           _ptr(scale), _ptr(shift), _stream()),
                   "wq_quant_ubits_group_quanto")
    STATS.launches += 1
    return packed, scale, shift, g


def gemm_u4a16(x: torch.Tensor, packed: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, group: int,
               bias: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QLinear.forward with qint4 weights: y = x @ (scale*q - shift).T + bias, fused."""
    N, K = packed.shape[0], packed.shape[1] * 2
    x2 = x.reshape(-1, K).contiguous()
    _need_cuda(x2, packed, scale, shift, bias)
    out_dtype = out_dtype or x2.dtype
    y = _dest(out, x2.shape[0], N, out_dtype, x.device, "gemm_u4a16")
    if _gemv_ok(x2, out_dtype):
        _gemv_weightonly(x2, 2, packed, scale, shift, group, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    if _gemv_f32_ok(x2, out_dtype):
        _gemv_f32(x2, 2, packed, scale, shift, group, 0, bias, y, N, K)
        return y.reshape(*x.shape[:-1], N)
    x2 = _operand(x2)
    with torch.cuda.device(x.device), _Timed("u4a16", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_u4a16(_ptr(x2), _DT[x2.dtype], _ptr(packed), _ptr(scale), _ptr(shift), group,
                                             _ptr(bias), _ptr(y), _DT[out_dtype], x2.shape[0], N, K, _stream()),
                   "wq_gemm_u4a16")
    STATS.launches += 1
    return y.reshape(*x.shape[:-1], N)


# ----------------------------------------------------------------------------------------------
# torch dynamic int8 (GPU twin)
# ----------------------------------------------------------------------------------------------
def torch_quantize_weight(w: torch.Tensor):
    """MinMaxObserver(per_tensor_symmetric, qint8) + quantize_per_tensor.

    Returns (int8 [N, K], scale f32 [1] on device, wsum int32 [N])."""
    w = w.to(torch.float32).contiguous()
    _need_cuda(w)
    N, K = w.shape
    q = torch.empty((N, K), dtype=torch.int8, device=w.device)
    scale = torch.empty((1,), dtype=torch.float32, device=w.device)
    wsum = torch.empty((N,), dtype=torch.int32, device=w.device)
    ws = torch.empty((2,), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().wq_quant_i8_tensor_torch(_ptr(w), N, K, _ptr(q), _ptr(scale), _ptr(wsum), _ptr(ws),
                                                        _stream()), "wq_quant_i8_tensor_torch")
    STATS.launches += 3
    return q, scale, wsum


def torch_quantize_activation(x: torch.Tensor):
    """Dynamic per-tensor quint8 (reduce_range): (uint8 like x, qparams f32 [2] = {scale, zp})."""
    x = x.contiguous()
    _need_cuda(x)
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    qparams = torch.empty((2,), dtype=torch.float32, device=x.device)
    ws = torch.empty((2,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().wq_quant_act_u8_tensor(_ptr(x), _DT[x.dtype], x.numel(), _ptr(q), _ptr(qparams),
                                                      _ptr(ws), _stream()), "wq_quant_act_u8_tensor")
    STATS.launches += 3
    return q, qparams


def gemm_dyn_i8(xq: torch.Tensor, qparams: torch.Tensor, wq: torch.Tensor, w_scale: torch.Tensor,
                wsum: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, K = wq.shape
    x2 = xq.reshape(-1, K)
    _need_cuda(x2, qparams, wq, w_scale, wsum, bias)
    y = torch.empty((x2.shape[0], N), dtype=torch.float32, device=xq.device)
    with torch.cuda.device(xq.device), _Timed("dyn_i8", x2.shape[0], N, K):
        _lib.check(_lib.load().wq_gemm_dyn_i8(_ptr(x2), _ptr(qparams), _ptr(wq), _ptr(w_scale), _ptr(wsum),
                                              _ptr(bias), _ptr(y), x2.shape[0], N, K, _stream()), "wq_gemm_dyn_i8")
    STATS.launches += 1
    return y.reshape(*xq.shape[:-1], N)


# ----------------------------------------------------------------------------------------------
# log-mel frontend and tallies
# ----------------------------------------------------------------------------------------------
def log_mel(audio: torch.Tensor, filters: torch.Tensor, n_samples: int = 480000,
            lengths: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """WhisperFeatureExtractor numerics: audio f32 [B, L] -> [B, n_mels, n_samples // 160]."""
    audio = audio.contiguous()
    _need_cuda(audio, filters, lengths)
    if audio.dtype != torch.float32 or filters.dtype != torch.float32:
        raise RuntimeError("log_mel: audio and filters must be float32")
    B, L = audio.shape
    n_mels = filters.shape[1]
    out = torch.empty((B, n_mels, n_samples // 160), dtype=out_dtype, device=audio.device)
    ws = torch.empty((B + 2 * n_mels,), dtype=torch.int32, device=audio.device)
    with torch.cuda.device(audio.device):
        _lib.check(_lib.load().wq_logmel(_ptr(audio), B, L, _ptr(lengths), n_samples, _ptr(filters), n_mels,
                                         _ptr(out), _DT[out_dtype], _ptr(ws), _stream()), "wq_logmel")
    STATS.launches += 3
    return out


def edit_distance(ref: torch.Tensor, ref_off: torch.Tensor, hyp: torch.Tensor, hyp_off: torch.Tensor
                  ) -> torch.Tensor:
    """Levenshtein distance per pair (ids int32 concatenated, offsets int64 [P+1]) -> int64 [P]."""
    _need_cuda(ref, ref_off, hyp, hyp_off)
    P = ref_off.numel() - 1
    dist = torch.empty((P,), dtype=torch.int64, device=ref.device)
    with torch.cuda.device(ref.device):
        _lib.check(_lib.load().wq_edit_distance(_ptr(ref), _ptr(ref_off), _ptr(hyp), _ptr(hyp_off), P, _ptr(dist),
                                                _stream()), "wq_edit_distance")
    STATS.launches += 1
    return dist

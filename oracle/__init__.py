"""CPU oracle for the compressed-Whisper hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this package, and only as the checker.  The product
package (``openai_whisper_compression_b200``) never imports it and has no CPU fallback.

Bit-exact fp32 / integer arithmetic lives in ``whisperq_oracle.c`` (built by
``oracle/Makefile`` into ``oracle/_build/liboracle.so``); this module wraps it with numpy
and adds the float64 restatements (log-mel, GEMM oracles).

Parity status (see the header of whisperq_oracle.c and DESIGN.md):
  * torch dynamic-int8 and log-mel: pinned against the live torch / HF implementations the
    reference calls (model_utils.py:131-134, data_utils.py:56-58) via tests/golden/.
  * bitsandbytes NF4 / LLM.int8, optimum-quanto qint8: PARITY UNPINNED (libraries absent, the
    reference holds no golden vectors); restated from SURVEY.md Appendix A.  The open points of A.2 were closed on
    a B200 in round 2 (scripts/bnb_open_points.cu, profiles/r02_bnb_open_points.json): the approximate division of
    the row scale DOES change codes, so kernels and oracle now use it (table fixture); mul-then-add and fmaf forms
    of int8_mm_dequant are the same instruction under nvcc's default -fmad=true.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/whisperq_oracle.c with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "whisperq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"])
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_edit_distance.restype = ctypes.c_int64
        _lib.orc_nf4_codebook.restype = ctypes.POINTER(ctypes.c_float)
        _lib.orc_fp4_codebook.restype = ctypes.POINTER(ctypes.c_float)
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


_QT = {"nf4": 0, "fp4": 1}

_FDIVIDEF = None


def fdividef_127_table() -> np.ndarray:
    """__fdividef(127.0f, h) for every fp16 bit pattern h of a positive finite absmax (float32[65536], index = fp16
    bits), dumped on a B200 by scripts/bnb_open_points.cu and committed as tests/golden/fdividef_127_fp16.npz:
    bitsandbytes' int8_vectorwise_quant kernel forms its row scale with that approximate-division intrinsic, which
    differs from the IEEE quotient for 9 185 of the 31 743 absmax values."""
    global _FDIVIDEF
    if _FDIVIDEF is None:
        path = os.path.join(os.path.dirname(_HERE), "tests", "golden", "fdividef_127_fp16.npz")
        _FDIVIDEF = np.ascontiguousarray(np.load(path)["table"], dtype=np.float32)
        assert _FDIVIDEF.shape == (65536,)
    return _FDIVIDEF

NF4_CODE = np.array([lib().orc_nf4_codebook()[i] for i in range(16)], dtype=np.float32)
FP4_CODE = np.array([lib().orc_fp4_codebook()[i] for i in range(16)], dtype=np.float32)


# --------------------------------------------------------------------------------------
# bitsandbytes 4-bit
# --------------------------------------------------------------------------------------
def quantize_4bit(w: np.ndarray, blocksize: int = 64, quant_type: str = "nf4"
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """bitsandbytes.functional.quantize_4bit: returns (packed uint8 [(n+1)//2, 1], absmax f32)."""
    wf = _f32(w).reshape(-1)
    n = wf.size
    packed = np.zeros(((n + 1) // 2, 1), dtype=np.uint8)
    absmax = np.zeros(((n + blocksize - 1) // blocksize,), dtype=np.float32)
    lib().orc_quant_4bit(_p(wf), ctypes.c_int64(n), ctypes.c_int(blocksize),
                         ctypes.c_int(_QT[quant_type]), _p(packed), _p(absmax))
    return packed, absmax


def dequantize_4bit(packed: np.ndarray, absmax: np.ndarray, shape: Sequence[int],
                    blocksize: int = 64, quant_type: str = "nf4",
                    dtype=np.float16) -> np.ndarray:
    """bitsandbytes.functional.dequantize_4bit: code[nibble]*absmax (fp32) rounded to dtype."""
    n = int(np.prod(shape))
    out = np.empty((n,), dtype=np.float32)
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    absmax = _f32(absmax)
    lib().orc_dequant_4bit(_p(packed), _p(absmax), ctypes.c_int64(n), ctypes.c_int(blocksize),
                           ctypes.c_int(_QT[quant_type]), _p(out))
    return out.astype(dtype).reshape(tuple(shape))


def linear4bit_forward(x: np.ndarray, packed, absmax, shape, bias=None, blocksize=64,
                       quant_type="nf4", compute_dtype=np.float16) -> np.ndarray:
    """Linear4bit.forward: F.linear(x.to(cd), dequantize_4bit(W).to(cd), bias.to(cd)).

    Returned in float64 (exact product of the rounded operands) -- the tolerance of the
    comparison models the fp16 accumulate/round of the library GEMM."""
    w = dequantize_4bit(packed, absmax, shape, blocksize, quant_type, compute_dtype)
    xs = np.asarray(x).astype(compute_dtype).astype(np.float64)
    y = xs.reshape(-1, shape[1]) @ w.astype(np.float64).T
    if bias is not None:
        y = y + np.asarray(bias).astype(compute_dtype).astype(np.float64)[None, :]
    return y.reshape(tuple(np.asarray(x).shape[:-1]) + (shape[0],))


def dynamic_map() -> np.ndarray:
    """bitsandbytes.functional.create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    the 256-entry code book of blockwise 8-bit quantization.  Built with torch.linspace exactly as
    the library does (the table is data defined by that formula)."""
    import torch
    data = []
    for i in range(7):
        boundaries = torch.linspace(0.1, 1, 2 ** i + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-6 + i)) * means).tolist()
        data += (-(10 ** (-6 + i)) * means).tolist()
    data += [0, 1.0]
    assert len(data) == 256
    data.sort()
    return torch.tensor(data, dtype=torch.float32).numpy()


def quantize_absmax_double(absmax: np.ndarray):
    """nested quantization of absmax: (q uint8 [n], absmax2 f32 [n/256], offset f32, absmax_deq f32 [n])."""
    a = _f32(absmax).reshape(-1)
    n = a.size
    code = dynamic_map()
    q = np.zeros((n,), dtype=np.uint8)
    a2 = np.zeros(((n + 255) // 256,), dtype=np.float32)
    deq = np.zeros((n,), dtype=np.float32)
    off = ctypes.c_float(0)
    lib().orc_quant_absmax_double(_p(a), ctypes.c_int64(n), _p(code), _p(q), _p(a2), ctypes.byref(off), _p(deq))
    return q, a2, np.float32(off.value), deq


# --------------------------------------------------------------------------------------
# bitsandbytes LLM.int8
# --------------------------------------------------------------------------------------
def int8_vectorwise_quant(a: np.ndarray, threshold: float = 0.0, approx_div: bool = True):
    """bitsandbytes.functional.int8_vectorwise_quant on fp16 input (approx_div: the library's __fdividef row scale,
    through the committed table; False: the IEEE quotient, kept for the sweep test that counts the difference).

    Returns (CA int8 [rows, cols], row_stats f32 [rows], outlier_cols int64 or None)."""
    a16 = np.asarray(a).astype(np.float16)
    af = _f32(a16).reshape(-1, a16.shape[-1])
    rows, cols = af.shape
    out = np.zeros((rows, cols), dtype=np.int8)
    stats = np.zeros((rows,), dtype=np.float32)
    flags = np.zeros((cols,), dtype=np.uint8)
    lib().orc_bnb_int8_vectorwise_quant(_p(af), ctypes.c_int64(rows), ctypes.c_int64(cols),
                                        ctypes.c_float(threshold), _p(out), _p(stats), _p(flags),
                                        _p(fdividef_127_table()) if approx_div else None)
    cols_idx = None
    if threshold > 0.0 and flags.any():
        cols_idx = np.nonzero(flags)[0].astype(np.int64)
        if rows > 1:
            lib().orc_bnb_zero_outlier_cols(_p(out), ctypes.c_int64(rows), ctypes.c_int64(cols),
                                            _p(flags))
    return out.reshape(a16.shape), stats, cols_idx


def igemm_nt(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int8)
    b = np.ascontiguousarray(b, dtype=np.int8)
    M, K = a.shape
    N = b.shape[0]
    c = np.zeros((M, N), dtype=np.int32)
    lib().orc_igemm_nt(_p(a), _p(b), ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int64(K), _p(c))
    return c


def int8_mm_dequant(c32, row_stats, col_stats, bias=None) -> np.ndarray:
    c32 = np.ascontiguousarray(c32, dtype=np.int32)
    M, N = c32.shape
    out = np.empty((M, N), dtype=np.float32)
    b = None if bias is None else _f32(bias)
    lib().orc_bnb_mm_dequant(_p(c32), _p(_f32(row_stats)), _p(_f32(col_stats)), _p(b),
                             ctypes.c_int64(M), ctypes.c_int64(N), _p(out))
    return out.astype(np.float16)


def int8_vectorwise_dequant(cb: np.ndarray, scb: np.ndarray) -> np.ndarray:
    cb = np.ascontiguousarray(cb, dtype=np.int8)
    N, K = cb.shape
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_bnb_vectorwise_dequant(_p(cb), _p(_f32(scb)), ctypes.c_int64(N), ctypes.c_int64(K),
                                     _p(out))
    return out


def linear8bitlt_forward(x: np.ndarray, CB: np.ndarray, SCB: np.ndarray, bias=None,
                         threshold: float = 6.0):
    """bitsandbytes MatMul8bitLt.forward (has_fp16_weights=False).

    Returns (y fp16, y_outlier_part float64 or None).  Without outliers y is bit-exact; with
    outliers y = fp16(fp32(y_int8_fp16) + subA @ subB) where the fp16 addmm's accumulation
    order is unspecified, so callers compare with a 1-ulp(fp16)-scale tolerance."""
    x16 = np.asarray(x).astype(np.float16)
    A = x16.reshape(-1, x16.shape[-1])
    CA, SCA, cols = int8_vectorwise_quant(A, threshold)
    c32 = igemm_nt(CA, CB)
    b16 = None if bias is None else np.asarray(bias).astype(np.float16)
    y = int8_mm_dequant(c32, SCA, SCB, b16)
    extra = None
    if cols is not None and cols.size:
        subA = A[:, cols].astype(np.float64)
        subB = int8_vectorwise_dequant(CB[:, cols], SCB).astype(np.float16).astype(np.float64)
        extra = subA @ subB.T
        y = (y.astype(np.float32) + extra.astype(np.float32)).astype(np.float16)
    return y.reshape(x16.shape[:-1] + (CB.shape[0],)), extra


# --------------------------------------------------------------------------------------
# optimum-quanto qint8
# --------------------------------------------------------------------------------------
def quanto_qint8(w: np.ndarray, dtype=np.float32):
    """quanto AbsmaxOptimizer + SymmetricQuantizer, weights=qint8, axis 0.  dtype float32: the C restatement.  dtype
    float16: the same two expressions (`absmax / 127`, `round(w / scale)`) as torch evaluates them on half tensors --
    each in fp32, the result rounded to half (opmath) -- which is what quanto computes for an fp16 model (the
    reference's static flows, model_utils.py:139-142,185); pinned against live torch in tests/test_oracle_golden.py."""
    if np.dtype(dtype) == np.float16:
        w16 = np.asarray(w).astype(np.float16)
        am = np.abs(w16).max(axis=1).astype(np.float32)
        s = (am / np.float32(127.0)).astype(np.float16)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = (w16.astype(np.float32) / s.astype(np.float32)[:, None]).astype(np.float16).astype(np.float32)
        r = np.where(np.isnan(r), np.float32(0), np.rint(r))
        return np.clip(r, -128, 127).astype(np.int8), s.astype(np.float32).reshape(-1, 1)
    wf = _f32(w)
    N, K = wf.shape
    q = np.zeros((N, K), dtype=np.int8)
    scale = np.zeros((N,), dtype=np.float32)
    lib().orc_quanto_qint8(_p(wf), ctypes.c_int64(N), ctypes.c_int64(K), _p(q), _p(scale))
    return q, scale.reshape(N, 1)


def quanto_qfloat8(w, dtype="float32"):
    """quanto weights=qfloat8 (e4m3fn), axis 0: scale = absmax / 448, data = (w / scale).to(float8_e4m3fn), evaluated
    with torch's own CPU ops in `dtype` (torch IS what quanto calls, so this is the live reference, not a
    restatement).  Returns (uint8 codes, scale float32 [N, 1])."""
    import torch
    t = torch.as_tensor(np.asarray(w, dtype=np.float32)).to(getattr(torch, dtype))
    s = t.abs().amax(dim=1, keepdim=True) / 448.0
    q = torch.clamp(torch.nan_to_num(t / s, nan=0.0), -448.0, 448.0).to(torch.float8_e4m3fn)   # SymmetricQuantizer clamps to finfo
    return q.view(torch.uint8).numpy().copy(), s.float().numpy().copy()


def quanto_quantize_activation(x, scale: float, qtype: str, dtype="float16"):
    """quanto quantize_activation(x, qtype, scale) + dequantize with torch's own CPU ops in `dtype`:
    qint8: clamp(round(x / scale), -128, 127); qfloat8: (x / scale).to(float8_e4m3fn).  Returns (code values as
    float32, dequantized tensor = code * scale in `dtype`, as float32)."""
    import torch
    dt = getattr(torch, dtype)
    t = torch.as_tensor(np.asarray(x, dtype=np.float32)).to(dt)
    s = torch.tensor(scale, dtype=dt)
    if qtype == "qint8":
        g = torch.clamp(torch.round(t / s), -128, 127)
    else:
        g = torch.clamp(t / s, -448.0, 448.0).to(torch.float8_e4m3fn).to(dt)      # clamp to finfo(e4m3fn), then cast
    return g.float().numpy(), (g * s).float().numpy()


def qlinear_forward(x: np.ndarray, q: np.ndarray, scale: np.ndarray, bias=None) -> np.ndarray:
    """QLinear.forward (weights only): matmul(x, Wq.to(x.dtype).t()) * scale + bias, float64."""
    xs = np.asarray(x).astype(np.float64)
    y = xs.reshape(-1, q.shape[1]) @ q.astype(np.float64).T
    y = y * np.asarray(scale).astype(np.float64).reshape(1, -1)
    if bias is not None:
        y = y + np.asarray(bias).astype(np.float64)[None, :]
    return y.reshape(tuple(np.asarray(x).shape[:-1]) + (q.shape[0],))


def quanto_group_size(in_features: int) -> int:
    """quanto QModuleMixin: 128, reduced in steps of 32 until it divides in_features; the whole
    row (per-axis quantization) when in_features <= 128 or nothing divides."""
    g = 128
    if in_features > g:
        while in_features % g != 0 and g > 32:
            g -= 32
        if in_features % g == 0:
            return g
    return in_features


def quanto_qint4(w: np.ndarray, group: Optional[int] = None, bits: int = 4):
    """(codes uint8 [N, K] one per element, scale f32 [N, K/g], shift f32 [N, K/g], group); bits = 2: qint2."""
    wf = _f32(w)
    N, K = wf.shape
    g = group or quanto_group_size(K)
    q = np.zeros((N, K), dtype=np.uint8)
    scale = np.zeros((N, K // g), dtype=np.float32)
    shift = np.zeros((N, K // g), dtype=np.float32)
    lib().orc_quanto_qbits(_p(wf), ctypes.c_int64(N), ctypes.c_int64(K), ctypes.c_int(g), ctypes.c_int(bits), _p(q),
                           _p(scale), _p(shift))
    return q, scale, shift, g


def quanto_qint4_pack(q: np.ndarray) -> np.ndarray:
    """libwhisperq packing: two codes per byte along K, first code in the high nibble."""
    return ((q[:, 0::2] << 4) | q[:, 1::2]).astype(np.uint8)


def quanto_qint4_dequant(q, scale, shift, group) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.uint8)
    N, K = q.shape
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_quanto_qint4_dequant(_p(q), _p(_f32(scale)), _p(_f32(shift)), ctypes.c_int64(N), ctypes.c_int64(K),
                                   ctypes.c_int(group), _p(out))
    return out


# --------------------------------------------------------------------------------------
# torch dynamic int8
# --------------------------------------------------------------------------------------
def torch_weight_qint8(w: np.ndarray):
    wf = _f32(w)
    q = np.zeros(wf.shape, dtype=np.int8)
    s = ctypes.c_float(0)
    lib().orc_torch_weight_qint8(_p(wf), ctypes.c_int64(wf.size), _p(q), ctypes.byref(s))
    return q, float(np.float32(s.value))


def torch_act_quant(x: np.ndarray, reduce_range: bool = True):
    xf = _f32(x)
    q = np.zeros(xf.shape, dtype=np.uint8)
    s = ctypes.c_float(0)
    zp = ctypes.c_int32(0)
    lib().orc_torch_act_quant(_p(xf), ctypes.c_int64(xf.size), ctypes.c_int(int(reduce_range)),
                              _p(q), ctypes.byref(s), ctypes.byref(zp))
    return q, float(np.float32(s.value)), int(zp.value)


def igemm_u8s8_nt(a: np.ndarray, zp: int, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.int8)
    M, K = a.shape
    N = b.shape[0]
    c = np.zeros((M, N), dtype=np.int32)
    lib().orc_igemm_u8s8_nt(_p(a), ctypes.c_int32(zp), _p(b), ctypes.c_int64(M), ctypes.c_int64(N),
                            ctypes.c_int64(K), _p(c))
    return c


def torch_dynamic_linear(x: np.ndarray, wq: np.ndarray, w_scale: float, bias=None) -> np.ndarray:
    """torch.ao.nn.quantized.dynamic.Linear.forward: per-tensor u8 activation (reduce_range),
    u8 x s8 -> s32, y = acc * (s_x * s_w) + bias in fp32."""
    xf = _f32(x)
    xq, sx, zp = torch_act_quant(xf, True)
    acc = igemm_u8s8_nt(xq.reshape(-1, xf.shape[-1]), zp, wq)
    M, N = acc.shape
    out = np.empty((M, N), dtype=np.float32)
    b = None if bias is None else _f32(bias)
    lib().orc_torch_requant(_p(acc), ctypes.c_float(sx), ctypes.c_float(w_scale), _p(b),
                            ctypes.c_int64(M), ctypes.c_int64(N), _p(out))
    return out.reshape(xf.shape[:-1] + (N,))


# --------------------------------------------------------------------------------------
# tallies
# --------------------------------------------------------------------------------------
def edit_distance(ref: Sequence[int], hyp: Sequence[int]) -> int:
    r = np.ascontiguousarray(np.asarray(ref, dtype=np.int32))
    h = np.ascontiguousarray(np.asarray(hyp, dtype=np.int32))
    return int(lib().orc_edit_distance(_p(r), ctypes.c_int64(r.size), _p(h), ctypes.c_int64(h.size)))


def _ids(tokens):
    table = {}
    return [table.setdefault(t, len(table)) for t in tokens], table


def wer_cer_tally(references: Sequence[str], predictions: Sequence[str]) -> np.ndarray:
    """int64[4] = {word errors, ref words, char errors, ref chars}; WER = 100*t[0]/t[1]
    (evaluate.load('wer'/'cer') semantics, evaluation.py:110-116)."""
    t = np.zeros(4, dtype=np.int64)
    for ref, hyp in zip(references, predictions):
        rw, hw = ref.split(), hyp.split()
        vocab = {}
        r = [vocab.setdefault(w, len(vocab)) for w in rw]
        h = [vocab.setdefault(w, len(vocab)) for w in hw]
        t[0] += edit_distance(r, h)
        t[1] += len(rw)
        t[2] += edit_distance([ord(c) for c in ref], [ord(c) for c in hyp])
        t[3] += len(ref)
    return t


from .logmel import log_mel_spectrogram, mel_filter_bank_slaney  # noqa: E402,F401

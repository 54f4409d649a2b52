"""Shared pieces of the producer-fused layer paths (fastgen decode step, fastenc encoder layer): the packed weights
of one or several drop-in linears of ONE scheme, concatenated along the output features (q|k|v, k|v), and the GEMM
call on them.

  int8    bitsandbytes-style Linear8bitLt: the rows arrive already quantized by a producer kernel (rowops.cu /
          attn_decode.cu); LLM.int8 GEMM with the outlier decomposition in its epilogue
  w8a16   optimum-quanto QLinear, qint8 weights
  w4a16   bitsandbytes Linear4bit (NF4 / FP4, blocksize 64, nested statistics resolved to fp32)
  u4a16   optimum-quanto QLinear, qint4 / qint2 weights (group-wise affine)

Concatenation does not change any per-row arithmetic: every scheme quantizes per output row (or per group inside a
row), so a fused GEMM over [Wq; Wk; Wv] gives bit for bit the columns the three separate GEMMs give."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import functional as F


class Packed:
    """Quantized state of 1..n linears of one scheme, concatenated along N."""
    __slots__ = ("kind", "cb", "scb", "bias", "threshold", "out_features", "in_features", "wq", "scale", "shift",
                 "group", "absmax", "quant_type")

    def __init__(self, kind: str):
        self.kind = kind
        self.cb = self.scb = self.bias = self.wq = self.scale = self.shift = self.absmax = None
        self.threshold, self.group, self.quant_type = 0.0, 0, "nf4"


PackedInt8 = Packed     # round-1 name


def small_rows_ok(rows: int, w: "Packed") -> bool:
    """True when Linear8bitLt's single-launch small-row kernel serves this call (functional.linear8bitlt's own rule)."""
    K = w.in_features
    return w.kind == "int8" and 0 < rows <= F.SMALL_M_ROWS and K % 16 == 0 and 64 * K + K + 272 <= 200 * 1024


def is_pow2(x: float) -> bool:
    return x > 0 and math.frexp(x)[0] == 0.5


def _cat_bias(mods, device) -> Optional[torch.Tensor]:
    """fp32 bias over the concatenated rows; a module without bias contributes exact zeros (x + 0 == x and
    fma(x, c, 0) == x * c, so its outputs are unchanged).  None when no module has one."""
    if all(m.bias is None for m in mods):
        return None
    return torch.cat([m.bias.detach().float() if m.bias is not None else
                      torch.zeros(m.out_features, dtype=torch.float32, device=device) for m in mods]).contiguous()


def pack_int8(mods: Sequence[torch.nn.Module]) -> Optional[Packed]:
    """Concatenate the quantized state of `mods` (all bitsandbytes-style Linear8bitLt on CUDA, one threshold);
    None when any of them is something else."""
    from .bnb import Linear8bitLt
    for m in mods:
        if type(m) is not Linear8bitLt or m.state.has_fp16_weights:
            return None
        if m.weight.CB is not None:
            m.init_8bit_state()
        if m.state.CB is None or not m.state.CB.is_cuda:
            return None
        if m.bias is not None and m.bias.dtype != torch.float16:
            m.bias.data = m.bias.data.to(torch.float16)     # what Linear8bitLt.forward does on first use
    w = Packed("int8")
    one = len(mods) == 1
    w.cb = mods[0].state.CB if one else torch.cat([m.state.CB for m in mods], 0).contiguous()
    w.scb = mods[0].state.SCB if one else torch.cat([m.state.SCB for m in mods]).contiguous()
    w.bias = _cat_bias(mods, w.cb.device)
    w.threshold = float(mods[0].state.threshold)
    w.out_features, w.in_features = w.cb.shape
    return w if all(float(m.state.threshold) == w.threshold for m in mods) else None


def pack(mods: Sequence[torch.nn.Module], dtype: torch.dtype) -> Optional[Packed]:
    """Packed weights of `mods` when they are all drop-in linears of one scheme that the fused paths serve for
    activations of `dtype` (fp16: every scheme; bf16: the weight-only schemes); None otherwise."""
    from .bnb import Linear4bit, Linear8bitLt
    from .quanto import QLinear
    if not mods or len({m.in_features for m in mods}) != 1:
        return None
    first = mods[0]
    if type(first) is Linear8bitLt:
        return pack_int8(mods) if dtype == torch.float16 else None
    if dtype not in (torch.float16, torch.bfloat16):
        return None
    K = first.in_features
    if type(first) is Linear4bit:
        if K % 64 != 0:
            return None
        for m in mods:
            qs = getattr(m.weight, "quant_state", None)
            if (type(m) is not Linear4bit or not getattr(m.weight, "bnb_quantized", False) or qs is None
                    or qs.blocksize != 64 or qs.quant_type != first.weight.quant_state.quant_type
                    or not m.weight.data.is_cuda or (m.compute_dtype not in (None, dtype))):
                return None
        w = Packed("w4a16")
        w.quant_type = first.weight.quant_state.quant_type
        w.wq = torch.cat([m.weight.data.view(m.out_features, K // 2) for m in mods], 0).contiguous()
        w.absmax = torch.cat([m.weight.quant_state.effective_absmax().view(m.out_features, K // 64)
                              for m in mods], 0).contiguous()
        w.bias = _cat_bias(mods, w.wq.device)
        w.out_features, w.in_features = w.wq.shape[0], K
        return w
    if type(first) is QLinear:
        for m in mods:
            if (type(m) is not QLinear or not m.frozen or m.weight_qtype is None or m.activation_qtype is not None
                    or m.weight_qtype.name != first.weight_qtype.name or not m._wq.is_cuda
                    or m._group != first._group):
                return None
        if first.weight_qtype.name == "qint8":
            w = Packed("w8a16")
            w.wq = torch.cat([m._wq for m in mods], 0).contiguous()
            w.scale = torch.cat([m._wscale.view(-1) for m in mods]).contiguous()
        elif first.weight_qtype.name in ("qint4", "qint2"):
            w = Packed("u4a16")
            w.wq = torch.cat([m._wq for m in mods], 0).contiguous()
            w.scale = torch.cat([m._wscale for m in mods], 0).contiguous()
            w.shift = torch.cat([m._wshift for m in mods], 0).contiguous()
            w.group = first._group
        else:
            return None
        w.bias = _cat_bias(mods, w.wq.device)
        w.out_features, w.in_features = w.wq.shape[0], K
        return w
    return None


def gemm_int8(quant, a: torch.Tensor, w: Packed, out: Optional[torch.Tensor] = None,
              keep_flags: bool = False, residual: Optional[torch.Tensor] = None, clamp_abs: float = 0.0,
              a_pre_gelu: bool = False) -> torch.Tensor:
    """Linear8bitLt's GEMM on rows that are already quantized: quant = (CA, SCA, outlier state) from a fused
    producer, `a` the fp16 rows they were made from (read only for outlier columns; with a_pre_gelu the rows BEFORE
    the GELU whose output was quantized -- gelu_quant(store_h=False)).  quant None (decode-shaped calls
    with <= 16 rows whose producer did not quantize): the single-launch kernel that quantizes the rows itself and
    multiplies with dp4a (gemv_small.cu; bit-identical results)."""
    if quant is None:
        if residual is not None or out is not None or a_pre_gelu:
            raise RuntimeError("gemm_int8: the small-row kernel takes neither a residual nor a destination")
        return F.linear8bitlt(a, w.cb, w.scb, w.bias, w.threshold)
    ca, sca, state = quant
    return F.gemm_llmint8(ca, sca, w.cb, w.scb, w.bias, a if state is not None else None, state, out=out,
                          keep_flags=keep_flags, residual=residual, clamp_abs=clamp_abs, a_pre_gelu=a_pre_gelu)


def gemm(quant, a: torch.Tensor, w: Packed, out: Optional[torch.Tensor] = None, keep_flags: bool = False) -> torch.Tensor:
    """y = a @ W^T + bias for any packed scheme.  quant: the producer's (CA, SCA, state) for the int8 scheme, ignored
    (None) for the weight-only ones, which read the fp16 / bf16 rows `a`.  out: optional [M, N] destination."""
    if w.kind == "int8":
        return gemm_int8(quant, a, w, out=out, keep_flags=keep_flags)
    if w.kind == "w8a16":
        return F.gemm_w8a16(a, w.wq, w.scale, w.bias, out=out)
    if w.kind == "w4a16":
        return F.gemm_w4a16(a, w.wq, w.absmax, w.out_features, w.in_features, w.bias, w.quant_type, out=out)
    return F.gemm_u4a16(a, w.wq, w.scale, w.shift, w.group, w.bias, out=out)

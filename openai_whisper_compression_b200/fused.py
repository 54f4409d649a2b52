"""Shared pieces of the producer-fused LLM.int8 layer paths (fastgen decode step, fastenc encoder layer):
packed weights of one or several ``Linear8bitLt`` modules and the GEMM call on rows a producer kernel
already quantized (rowops.cu / attn_decode.cu)."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import functional as F


class PackedInt8:
    """int8 weights [sum N, K], fp32 row scales and fp32 bias of 1..n Linear8bitLt modules, concatenated along N."""
    __slots__ = ("cb", "scb", "bias", "threshold", "out_features")


def is_pow2(x: float) -> bool:
    return x > 0 and math.frexp(x)[0] == 0.5


def pack_int8(mods: Sequence[torch.nn.Module]) -> Optional[PackedInt8]:
    """Concatenate the quantized state of `mods` (all bitsandbytes-style Linear8bitLt on CUDA, one threshold);
    None when any of them is something else.  A module without bias contributes exact zeros
    (fma(x, c, 0) == x * c, so its outputs are unchanged)."""
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
    w = PackedInt8()
    one = len(mods) == 1
    w.cb = mods[0].state.CB if one else torch.cat([m.state.CB for m in mods], 0).contiguous()
    w.scb = mods[0].state.SCB if one else torch.cat([m.state.SCB for m in mods]).contiguous()
    if all(m.bias is None for m in mods):
        w.bias = None
    else:
        w.bias = torch.cat([m.bias.detach().float() if m.bias is not None else
                            torch.zeros(m.out_features, dtype=torch.float32, device=w.cb.device)
                            for m in mods]).contiguous()
    w.threshold = float(mods[0].state.threshold)
    w.out_features = w.cb.shape[0]
    return w if all(float(m.state.threshold) == w.threshold for m in mods) else None


def gemm_int8(quant, a: torch.Tensor, w: PackedInt8, out: Optional[torch.Tensor] = None,
              keep_flags: bool = False) -> torch.Tensor:
    """Linear8bitLt's GEMM on rows that are already quantized: quant = (CA, SCA, outlier state) from a fused
    producer, `a` the fp16 rows they were made from (read only for outlier columns)."""
    ca, sca, state = quant
    return F.gemm_llmint8(ca, sca, w.cb, w.scb, w.bias, a if state is not None else None, state, out=out,
                          keep_flags=keep_flags)

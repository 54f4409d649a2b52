"""Drop-in mirror of the bitsandbytes surface the reference (and HF transformers on its behalf)
touches: ``bitsandbytes.nn.{Linear4bit, Params4bit, Linear8bitLt, Int8Params}`` and
``bitsandbytes.functional.{quantize_4bit, dequantize_4bit, int8_vectorwise_quant, ...}``.

Reference call sites
  pruning+quantization/bnb_implementation.py:1093-1118  convert_model_to_4bit (direct swap)
  pruning+quantization/bnb_implementation.py:1216-1221  .to(device) triggers quantization
  model_utils.py:24-49,102-118                          BitsAndBytesConfig(load_in_4bit=...) via HF
  BASELINE.json configs[1]                              Linear8bitLt / LLM.int8, threshold 6.0

Same constructor signatures, attribute names (``weight.quant_state``, ``weight.CB/SCB``,
``state.threshold`` ...), ``isinstance(m, nn.Linear)`` stays true, quantization happens when the
parameter is moved to a CUDA device, and ``state_dict()`` holds only plain tensors.  All
arithmetic runs in libwhisperq.so (sm_100a); there is no CPU implementation -- calling
``forward`` before the module is on a CUDA device raises.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Optional

import torch
from torch import nn

from . import functional as F

_NF4 = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
        -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
        0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
        0.7229568362236023, 1.0]
_FP4 = [0.0, 0.0052083333, 0.6666667, 1.0, 0.33333334, 0.5, 0.16666667, 0.25,
        -0.0, -0.0052083333, -0.6666667, -1.0, -0.33333334, -0.5, -0.16666667, -0.25]


class QuantState:
    """bitsandbytes.functional.QuantState (the fields the reference / HF read)."""

    valid_quant_types = ("fp4", "nf4")

    def __init__(self, absmax, shape=None, code=None, blocksize=None, quant_type=None, dtype=None,
                 offset=None, state2=None):
        self.absmax = absmax
        self.shape = shape
        self.code = code
        self.dtype = dtype
        self.blocksize = blocksize
        self.quant_type = quant_type
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None

    def effective_absmax(self) -> torch.Tensor:
        """fp32 per-block statistics as dequantize_4bit uses them (nested: code[q]*absmax2 + offset)."""
        if not self.nested:
            return self.absmax
        cached = getattr(self, "_absmax_f32", None)
        if cached is None or cached.device != self.absmax.device:
            cached = self._absmax_f32 = F.dequantize_absmax_double(self.absmax, self.state2.absmax, self.offset)
        return cached

    def as_dict(self, packed: bool = False) -> Dict[str, Any]:
        meta = {"quant_type": self.quant_type, "blocksize": self.blocksize,
                "dtype": str(self.dtype).replace("torch.", ""), "shape": tuple(self.shape)}
        out: Dict[str, Any] = {"absmax": self.absmax, "quant_map": self.code}
        if self.nested:
            out["nested_absmax"] = self.state2.absmax
            out["nested_quant_map"] = self.state2.code
            meta.update(nested_blocksize=self.state2.blocksize, nested_dtype="float32",
                        nested_offset=float(self.offset.item()))
        if not packed:
            out.update(meta)
            return out
        blob = json.dumps(meta).encode("utf-8")
        out["quant_state.bitsandbytes__" + self.quant_type] = torch.tensor(list(blob), dtype=torch.uint8)
        return out

    def to(self, device):
        self.absmax = self.absmax.to(device)
        self.code = self.code.to(device)
        if self.nested:
            self.offset = self.offset.to(device)
            self.state2.to(device)
            self._absmax_f32 = None
        return self


def quantize_4bit(A: torch.Tensor, absmax=None, out=None, blocksize: int = 64, compress_statistics: bool = False,
                  quant_type: str = "fp4", quant_storage=torch.uint8):
    """bitsandbytes.functional.quantize_4bit -> (packed uint8 [(n+1)//2, 1], QuantState)."""
    if quant_type not in ("fp4", "nf4"):
        raise NotImplementedError(f"4-bit quantization data type {quant_type} is not implemented.")
    if quant_storage != torch.uint8:
        raise NotImplementedError("only quant_storage=torch.uint8 is implemented")
    packed, am = F.quantize_4bit(A, blocksize, quant_type)
    code = torch.tensor(_NF4 if quant_type == "nf4" else _FP4, dtype=torch.float32, device=A.device)
    if compress_statistics:
        # nested ("double") quantization: offset = mean(absmax), 8-bit blockwise (256) dynamic map
        q, absmax2, offset, absmax_deq = F.quantize_absmax_double(am)
        state2 = QuantState(absmax=absmax2, code=F.dynamic_map(A.device), blocksize=256, dtype=torch.float32)
        qs = QuantState(absmax=q, shape=A.shape, dtype=A.dtype, blocksize=blocksize, code=code,
                        quant_type=quant_type, offset=offset, state2=state2)
        qs._absmax_f32 = absmax_deq
        return packed, qs
    return packed, QuantState(absmax=am, shape=A.shape, dtype=A.dtype, blocksize=blocksize, code=code,
                              quant_type=quant_type)


def dequantize_4bit(A: torch.Tensor, quant_state: Optional[QuantState] = None, absmax=None, out=None,
                    blocksize: int = 64, quant_type: str = "fp4") -> torch.Tensor:
    """bitsandbytes.functional.dequantize_4bit."""
    if quant_state is None:
        raise ValueError("dequantize_4bit needs a quant_state")
    return F.dequantize_4bit(A, quant_state.effective_absmax(), quant_state.shape, quant_state.blocksize,
                             quant_state.quant_type, quant_state.dtype)


def quantize_nf4(A, **kw):
    return quantize_4bit(A, quant_type="nf4", **kw)


def quantize_fp4(A, **kw):
    return quantize_4bit(A, quant_type="fp4", **kw)


def int8_vectorwise_quant(A: torch.Tensor, threshold: float = 0.0):
    """bitsandbytes.functional.int8_vectorwise_quant -> (CA, row_stats, outlier_cols or None).

    Returning the outlier columns as a tensor of the right length needs one host read of the
    count; the module forward path keeps everything on the device instead."""
    ca, stats, st = F.int8_vectorwise_quant(A, threshold)
    cols = None
    if st is not None:
        n = int(st.n_outliers.item())
        if n:
            cols = st.outlier_cols[:n].to(torch.int64).clone()
    return ca, stats, cols


def int8_vectorwise_dequant(A: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    return A * stats.view(-1, 1) * 7.874015718698502e-3


def matmul_4bit(A: torch.Tensor, B: torch.Tensor, quant_state: QuantState, out=None, bias=None) -> torch.Tensor:
    """bnb.matmul_4bit(A, W.t(), quant_state, bias): fused dequant GEMM for every shape."""
    N, K = quant_state.shape
    if quant_state.blocksize != 64:
        raise NotImplementedError("fused 4-bit GEMM supports blocksize 64 (the bitsandbytes default)")
    x = A
    if x.dtype == torch.float32:
        # fp32 compute flow (bnb_implementation.py:1216-1218): tensor cores take fp16 operands,
        # accumulate fp32 and write fp32 (DESIGN.md "Numerics")
        return F.gemm_w4a16(x, B, quant_state.effective_absmax(), N, K,
                            None if bias is None else bias.float(), quant_state.quant_type, torch.float32)
    return F.gemm_w4a16(x, B, quant_state.effective_absmax(), N, K, None if bias is None else bias.float(),
                        quant_state.quant_type)


class Params4bit(nn.Parameter):
    """bitsandbytes.nn.Params4bit: quantizes itself when moved to a CUDA device."""

    def __new__(cls, data=None, requires_grad=False, quant_state=None, blocksize=64, compress_statistics=True,
                quant_type="fp4", quant_storage=torch.uint8, module=None, bnb_quantized=False):
        if data is None:
            data = torch.empty(0)
        self = torch.Tensor._make_subclass(cls, data, requires_grad)
        self.blocksize = blocksize
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_state = quant_state
        self.quant_storage = quant_storage
        self.bnb_quantized = bnb_quantized
        self.module = module
        return self

    def __deepcopy__(self, memo):
        new = type(self).__new__(type(self), self.data.clone(), self.requires_grad, self.quant_state,
                                 self.blocksize, self.compress_statistics, self.quant_type, self.quant_storage,
                                 self.module, self.bnb_quantized)
        return new

    def _quantize(self, device):
        w = self.data.contiguous().to(device)
        packed, qs = quantize_4bit(w, blocksize=self.blocksize, compress_statistics=self.compress_statistics,
                                   quant_type=self.quant_type, quant_storage=self.quant_storage)
        self.data = packed
        self.quant_state = qs
        if self.module is not None:
            self.module.quant_state = qs
        self.bnb_quantized = True
        return self

    def cuda(self, device=None, non_blocking=False):
        return self.to(device="cuda" if device is None else device, non_blocking=non_blocking)

    def to(self, *args, **kwargs):
        device, dtype, non_blocking, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None and device.type == "cuda" and not self.bnb_quantized:
            return self._quantize(device)
        if self.quant_state is not None and device is not None:
            self.quant_state.to(device)
        new = Params4bit(super().to(device=device, dtype=dtype, non_blocking=non_blocking),
                         requires_grad=self.requires_grad, quant_state=self.quant_state,
                         blocksize=self.blocksize, compress_statistics=self.compress_statistics,
                         quant_type=self.quant_type, quant_storage=self.quant_storage, module=self.module,
                         bnb_quantized=self.bnb_quantized)
        return new


class Linear4bit(nn.Linear):
    """bitsandbytes.nn.Linear4bit (NF4 / FP4, blocksize 64)."""

    def __init__(self, input_features, output_features, bias=True, compute_dtype=None, compress_statistics=True,
                 quant_type="fp4", quant_storage=torch.uint8, device=None):
        super().__init__(input_features, output_features, bias, device)
        self.weight = Params4bit(self.weight.data, requires_grad=False, compress_statistics=compress_statistics,
                                 quant_type=quant_type, quant_storage=quant_storage, module=self)
        self.compute_dtype = compute_dtype
        self.compute_type_is_set = compute_dtype is not None
        self.quant_state = None
        self.quant_storage = quant_storage

    def _apply(self, fn, recurse=True):
        # nn.Module._apply rebuilds parameters as plain nn.Parameter(fn(param.data)), which would
        # bypass Params4bit.to(); route the weight through it so that .to("cuda") quantizes and
        # packed codes are only ever moved, never cast.
        w = self.weight
        probe = fn(torch.empty(0, device=w.device, dtype=torch.float32))
        if w.bnb_quantized or probe.device.type == "cuda":
            new_w = w.to(probe.device)
        else:
            new_w = Params4bit(fn(w.data), requires_grad=False, compress_statistics=w.compress_statistics,
                               quant_type=w.quant_type, quant_storage=w.quant_storage, module=self,
                               blocksize=w.blocksize)
        self._parameters["weight"] = new_w
        for k, v in self._parameters.items():
            if k != "weight" and v is not None:
                self._parameters[k] = nn.Parameter(fn(v.data), requires_grad=v.requires_grad)
        for k, b in self._buffers.items():
            if b is not None:
                self._buffers[k] = fn(b)
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        # the reference loads the ORIGINAL fp weights into the new module before .to(device)
        # (bnb_implementation.py:1116): keep them as a not-yet-quantized Params4bit
        w = state_dict.get("weight")
        if w is not None and w.dtype.is_floating_point and not self.weight.bnb_quantized:
            self.weight = Params4bit(w.detach().clone(), requires_grad=False,
                                     compress_statistics=self.weight.compress_statistics,
                                     quant_type=self.weight.quant_type, quant_storage=self.quant_storage,
                                     module=self)
            if self.bias is not None and "bias" in state_dict:
                self.bias.data = state_dict["bias"].detach().clone().to(self.bias.dtype)
            return torch.nn.modules.module._IncompatibleKeys([], [])
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        qs = getattr(self.weight, "quant_state", None)
        if qs is not None:
            for k, v in qs.as_dict(packed=True).items():
                destination[prefix + "weight." + k] = v if keep_vars else v.detach()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w = self.weight
        if not getattr(w, "bnb_quantized", False) or w.quant_state is None:
            raise RuntimeError("Linear4bit has not been quantized: move the module to a CUDA device first "
                               "(quantization runs in the sm_100a library; there is no CPU path)")
        if self.bias is not None and self.bias.dtype != x.dtype:
            self.bias.data = self.bias.data.to(x.dtype)
        if not self.compute_type_is_set:
            self.compute_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else self.compute_dtype
            self.compute_type_is_set = True
        inp_dtype = x.dtype
        if self.compute_dtype is not None:
            x = x.to(self.compute_dtype)
        bias = None if self.bias is None else self.bias.to(self.compute_dtype or x.dtype)
        return matmul_4bit(x, w.data, quant_state=w.quant_state, bias=bias).to(inp_dtype)


class LinearNF4(Linear4bit):
    def __init__(self, input_features, output_features, bias=True, compute_dtype=None, compress_statistics=True,
                 quant_storage=torch.uint8, device=None):
        super().__init__(input_features, output_features, bias, compute_dtype, compress_statistics, "nf4",
                         quant_storage, device)


class LinearFP4(Linear4bit):
    def __init__(self, input_features, output_features, bias=True, compute_dtype=None, compress_statistics=True,
                 quant_storage=torch.uint8, device=None):
        super().__init__(input_features, output_features, bias, compute_dtype, compress_statistics, "fp4",
                         quant_storage, device)


# --------------------------------------------------------------------------------------------------
# LLM.int8
# --------------------------------------------------------------------------------------------------
class MatmulLtState:
    """bitsandbytes.autograd._functions.MatmulLtState (fields HF / callers read)."""

    def __init__(self):
        self.CB = None
        self.SCB = None
        self.idx = None
        self.subB = None
        self.has_fp16_weights = True
        self.threshold = 0.0
        self.is_training = True
        self.memory_efficient_backward = False
        self.use_pool = False


class Int8Params(nn.Parameter):
    """bitsandbytes.nn.Int8Params: row-wise int8 quantization when moved to a CUDA device."""

    def __new__(cls, data=None, requires_grad=True, has_fp16_weights=False, CB=None, SCB=None):
        if data is None:
            data = torch.empty(0)
        obj = torch.Tensor._make_subclass(cls, data, requires_grad)
        obj.CB = CB
        obj.SCB = SCB
        obj.has_fp16_weights = has_fp16_weights
        return obj

    def _quantize(self, device):
        if self.has_fp16_weights:
            return super().to(device)
        B = self.data.contiguous().to(device=device, dtype=torch.float16)
        CB, SCB, _ = F.int8_vectorwise_quant(B, 0.0)
        self.data = CB
        self.CB = CB
        self.SCB = SCB
        return self

    def cuda(self, device=None, non_blocking=False):
        return self.to(device="cuda" if device is None else device)

    def to(self, *args, **kwargs):
        device, dtype, non_blocking, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None and device.type == "cuda" and self.data.dtype != torch.int8:
            return self._quantize(device)
        new = Int8Params(super().to(device=device, dtype=dtype if self.data.dtype != torch.int8 else None,
                                    non_blocking=non_blocking),
                         requires_grad=self.requires_grad, has_fp16_weights=self.has_fp16_weights)
        new.CB = None if self.CB is None else new.data
        new.SCB = None if self.SCB is None else self.SCB.to(new.data.device)
        return new


class Linear8bitLt(nn.Linear):
    """bitsandbytes.nn.Linear8bitLt with has_fp16_weights=False (inference)."""

    def __init__(self, input_features, output_features, bias=True, has_fp16_weights=True, threshold=0.0,
                 index=None, device=None):
        super().__init__(input_features, output_features, bias, device)
        if has_fp16_weights:
            raise NotImplementedError("Linear8bitLt(has_fp16_weights=True) is a training mode; the inference "
                                      "path (HF load_in_8bit) uses has_fp16_weights=False")
        self.state = MatmulLtState()
        self.index = index
        self.state.threshold = threshold
        self.state.has_fp16_weights = has_fp16_weights
        self.weight = Int8Params(self.weight.data, has_fp16_weights=has_fp16_weights,
                                 requires_grad=has_fp16_weights)

    def _apply(self, fn, recurse=True):
        w = self.weight
        probe = fn(torch.empty(0, device=w.device, dtype=torch.float32))
        if w.data.dtype == torch.int8 or probe.device.type == "cuda":
            new_w = w.to(probe.device)
        else:
            new_w = Int8Params(fn(w.data), has_fp16_weights=False, requires_grad=False)
        self._parameters["weight"] = new_w
        if self.state.CB is not None and self.state.CB.device != probe.device:
            self.state.CB = self.state.CB.to(probe.device)
            self.state.SCB = self.state.SCB.to(probe.device)
        for k, v in self._parameters.items():
            if k != "weight" and v is not None:
                self._parameters[k] = nn.Parameter(fn(v.data), requires_grad=v.requires_grad)
        for k, b in self._buffers.items():
            if b is not None:
                self._buffers[k] = fn(b)
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        w = state_dict.get("weight")
        if w is not None and w.dtype.is_floating_point and self.weight.data.dtype != torch.int8:
            self.weight = Int8Params(w.detach().clone(), has_fp16_weights=False, requires_grad=False)
            if self.bias is not None and "bias" in state_dict:
                self.bias.data = state_dict["bias"].detach().clone().to(self.bias.dtype)
            return torch.nn.modules.module._IncompatibleKeys([], [])
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        scb = self.weight.SCB if self.weight.SCB is not None else self.state.SCB
        if scb is not None:
            destination[prefix + "SCB"] = scb if keep_vars else scb.detach()
            destination[prefix + "weight_format"] = torch.tensor(0, dtype=torch.uint8)

    def init_8bit_state(self):
        self.state.CB = self.weight.CB
        self.state.SCB = self.weight.SCB
        self.weight.CB = None
        self.weight.SCB = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.state.is_training = self.training
        if self.weight.CB is not None:
            self.init_8bit_state()
        if self.state.CB is None:
            raise RuntimeError("Linear8bitLt has not been quantized: move the module to a CUDA device first "
                               "(quantization runs in the sm_100a library; there is no CPU path)")
        if self.bias is not None and self.bias.dtype != x.dtype:
            self.bias.data = self.bias.data.to(x.dtype)
        bias = self.bias
        if bias is None or bias.dtype == torch.float16:
            # fused bias (bitsandbytes int8_mm_dequant applies an fp16 bias inside the kernel); the
            # kernel reads its exact fp32 widening, cached per parameter version
            if bias is not None:
                key = (bias.data_ptr(), bias._version)
                if getattr(self, "_bias_f32_key", None) != key:
                    self._bias_f32 = bias.detach().float()
                    self._bias_f32_key = key
                bias = self._bias_f32
            return F.linear8bitlt(x, self.state.CB, self.state.SCB, bias, float(self.state.threshold))
        out = F.linear8bitlt(x, self.state.CB, self.state.SCB, None, float(self.state.threshold))
        return out.add_(bias)

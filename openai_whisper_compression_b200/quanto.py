"""Drop-in mirror of the optimum-quanto 0.2.6 surface the reference imports:
``from optimum.quanto import Calibration, freeze, qfloat8, qint4, qint8, quantize``
(model_utils.py:6; call sites model_utils.py:126-128,185-211,
pruning+quantization/quanto_implementation.py:648-670).

Implemented on the sm_100a library: weights-only ``qint8`` (per-output-channel absmax / 127,
``torch.round`` half-to-even, W8A16 fused GEMM) and weights-only ``qint4`` (group-wise affine uint4,
MaxOptimizer: scale = (max - min) / 15, float shift = -min; fused W4A16 GEMM) and ``qint2`` (the same with
three levels above zero; the 2-bit codes are kept one per nibble in the qint4 container, so the arithmetic is quanto's
and the storage is 4 bits per weight).  ``quantize`` swaps every ``nn.Linear`` (incl.
``proj_out``) for ``QLinear``; ``freeze`` fixes the integer weights.  The reference quantizes a
CPU model and moves it to the device afterwards (model_utils.py:126-137): ``freeze`` on a CPU
module records the request and the integer codes are produced by the CUDA kernel the moment the
module reaches a CUDA device -- there is no CPU arithmetic path.  Not implemented yet (SURVEY.md
section 8f rank 2, raise NotImplementedError): qfloat8 weights, activation quantization + Calibration.
"""
from __future__ import annotations

import fnmatch
from typing import Optional

import torch
from torch import nn

from . import functional as F


class qtype:
    def __init__(self, name: str, is_floating_point: bool, bits: int, dtype, qmin: float, qmax: float):
        self.name, self.is_floating_point, self.bits, self.dtype = name, is_floating_point, bits, dtype
        self.qmin, self.qmax = qmin, qmax

    def __str__(self):
        return f"quanto.{self.name}"

    __repr__ = __str__

    def __hash__(self):
        return hash(self.name)


qint2 = qtype("qint2", False, 2, torch.int8, -2, 1)
qint4 = qtype("qint4", False, 4, torch.int8, -8, 7)
qint8 = qtype("qint8", False, 8, torch.int8, -128, 127)
qfloat8 = qtype("qfloat8_e4m3fn", True, 8, torch.float8_e4m3fn, -448.0, 448.0)
qtypes = {q.name: q for q in (qint2, qint4, qint8, qfloat8)}


class QLinear(nn.Linear):
    """optimum.quanto.nn.QLinear, weights-only qint8.

    state_dict after freeze (quanto 0.2.6 layout): ``weight._data`` int8 [N, K],
    ``weight._scale`` [N, 1], ``bias``, ``input_scale``, ``output_scale``."""

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, weights=None,
                 activations=None, optimizer=None, quantize_input=False):
        super().__init__(in_features, out_features, bias, device, dtype)
        if activations is not None:
            raise NotImplementedError("quanto activation quantization (static, Calibration) is outside the "
                                      "built hot path (SURVEY.md section 8f rank 2)")
        if weights is not None and getattr(weights, "name", None) not in ("qint8", "qint4", "qint2"):
            raise NotImplementedError(f"quanto weights={weights} is not implemented yet; qint8, qint4 and qint2 "
                                      "are (SURVEY.md section 8f rank 2)")
        if weights is not None and weights.name in ("qint4", "qint2") and in_features % 64 != 0:
            raise NotImplementedError("qint4 / qint2 weights need in_features % 64 == 0 in the fused GEMM")
        self.weight_qtype = weights
        self.activation_qtype = activations
        self.optimizer = optimizer
        self.register_buffer("input_scale", torch.ones((), dtype=self.weight.dtype))
        self.register_buffer("output_scale", torch.ones((), dtype=self.weight.dtype))
        self._frozen = False
        self._freeze_pending = False
        self._wq: Optional[torch.Tensor] = None       # int8 [N, K] (qint8) / packed uint8 [N, K/2] (qint4)
        self._wscale: Optional[torch.Tensor] = None   # fp32 [N, 1] (qint8) / [N, K/group] (qint4)
        self._wshift: Optional[torch.Tensor] = None   # fp32 [N, K/group] (qint4)
        self._group = 0
        self._bias_f32: Optional[torch.Tensor] = None

    @classmethod
    def from_module(cls, module: nn.Linear, weights=None, activations=None, optimizer=None):
        q = cls(module.in_features, module.out_features, module.bias is not None, device=module.weight.device,
                dtype=module.weight.dtype, weights=weights, activations=activations, optimizer=optimizer)
        q.weight = module.weight
        q.bias = module.bias
        return q

    # -- freezing -------------------------------------------------------------------------------
    @property
    def frozen(self) -> bool:
        return self._frozen

    def freeze(self):
        if self._frozen or self.weight_qtype is None:
            return
        if self.weight.is_cuda:
            self._quantize_now()
        else:
            self._freeze_pending = True   # codes are produced on arrival at a CUDA device

    def _quantize_now(self):
        w = self.weight.data
        if self.weight_qtype.name in ("qint4", "qint2"):
            # qint2 codes (0..3) live in the qint4 container (one code per nibble): same fused GEMM
            q, scale, self._wshift, self._group = F.quanto_quantize_qint4(w, bits=self.weight_qtype.bits)
        else:
            q, scale = F.quanto_quantize_qint8(w)
        self._wq, self._wscale = q, scale
        self._scale_dtype = w.dtype
        # the float weight is gone after freeze (as in quanto): `weight` now holds the int8 codes
        self.weight = nn.Parameter(q, requires_grad=False)
        self._frozen, self._freeze_pending = True, False

    def _apply(self, fn, recurse=True):
        if self._frozen:
            dev = fn(torch.empty(0, device=self._wq.device, dtype=torch.float32)).device
            self._wq = self._wq.to(dev)
            self._wscale = self._wscale.to(dev)
            if self._wshift is not None:
                self._wshift = self._wshift.to(dev)
            self._parameters["weight"] = nn.Parameter(self._wq, requires_grad=False)
            self._bias_f32 = None
            for k, v in self._parameters.items():
                if k != "weight" and v is not None:
                    self._parameters[k] = nn.Parameter(fn(v.data), requires_grad=v.requires_grad)
            for k, b in self._buffers.items():
                if b is not None:
                    self._buffers[k] = fn(b)
            return self
        super()._apply(fn, recurse)
        if self._freeze_pending and self.weight.is_cuda:
            self._quantize_now()
        return self

    @property
    def qweight(self):
        """(int8 codes, scale) -- quanto exposes a WeightQBytesTensor with ._data / ._scale."""
        if not self._frozen:
            raise RuntimeError("QLinear is not frozen on a CUDA device yet")
        return self._wq, self._wscale.to(self._scale_dtype)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if not self._frozen:
            return super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + "weight._data"] = self._wq
        destination[prefix + "weight._scale"] = self._wscale.to(self._scale_dtype)
        if self._wshift is not None:
            destination[prefix + "weight._shift"] = self._wshift.to(self._scale_dtype)
        if self.bias is not None:
            destination[prefix + "bias"] = self.bias if keep_vars else self.bias.detach()
        destination[prefix + "input_scale"] = self.input_scale
        destination[prefix + "output_scale"] = self.output_scale

    # -- forward --------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.weight_qtype is None:
            return nn.functional.linear(x, self.weight, self.bias)
        if not self._frozen:
            if self.weight.is_cuda:
                # quanto quantizes on the fly until freeze(); the codes are identical, so freeze now
                self._quantize_now()
            else:
                raise RuntimeError("QLinear is on the CPU: move the model to a CUDA device (quantization and "
                                   "the W8A16 GEMM run in the sm_100a library; there is no CPU path)")
        if self.bias is not None and (self._bias_f32 is None or self._bias_f32.device != x.device):
            self._bias_f32 = self.bias.detach().float().contiguous()
        bias = self._bias_f32 if self.bias is not None else None
        if self._wshift is not None:      # qint4
            if x.dtype == torch.float32:
                return F.gemm_u4a16(x.to(torch.float16), self._wq, self._wscale, self._wshift, self._group, bias,
                                    torch.float32)
            return F.gemm_u4a16(x, self._wq, self._wscale, self._wshift, self._group, bias)
        if x.dtype == torch.float32:
            # fp32 flow of the reference (model never .half()-ed, model_utils.py:139-142): operands go
            # to the tensor cores as fp16, accumulate fp32, result written fp32 (DESIGN.md "Numerics")
            return F.gemm_w8a16(x.to(torch.float16), self._wq, self._wscale, bias, torch.float32)
        return F.gemm_w8a16(x, self._wq, self._wscale, bias)


def _match(name: str, patterns) -> bool:
    if patterns is None:
        return False
    patterns = [patterns] if isinstance(patterns, str) else patterns
    return any(fnmatch.fnmatch(name, p) for p in patterns)


def quantize(model: nn.Module, weights=None, activations=None, optimizer=None, include=None, exclude=None):
    """optimum.quanto.quantize: swap every nn.Linear for QLinear in place."""
    if isinstance(weights, str):
        weights = qtypes[weights]
    if isinstance(activations, str):
        activations = qtypes[activations]
    for name, m in list(model.named_modules()):
        if include is not None and not _match(name, include):
            continue
        if exclude is not None and _match(name, exclude):
            continue
        if type(m) is not nn.Linear:
            continue
        q = QLinear.from_module(m, weights=weights, activations=activations, optimizer=optimizer)
        parent = model
        *path, leaf = name.split(".")
        for p in path:
            parent = getattr(parent, p)
        setattr(parent, leaf, q)


def freeze(model: nn.Module):
    """optimum.quanto.freeze."""
    for m in model.modules():
        if isinstance(m, QLinear):
            m.freeze()


class Calibration:
    """optimum.quanto.Calibration -- activation calibration is not part of the built path."""

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        raise NotImplementedError("quanto activation calibration (static quantization) is outside the built "
                                  "hot path (SURVEY.md section 8f rank 2)")

    def __exit__(self, *exc):
        return False

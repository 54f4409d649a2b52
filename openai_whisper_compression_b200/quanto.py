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
module reaches a CUDA device -- there is no CPU arithmetic path.

Round 2 (the reference's six "static_quanto_*" configs, quantization.py:53-86, through
model_utils.py:152-214 apply_static_quantization): ``qfloat8`` (e4m3fn) weights, and static activation quantization --
``quantize(model, weights=..., activations=qint8 | qfloat8)`` additionally swaps ``nn.LayerNorm`` for ``QLayerNorm``,
``with Calibration():`` records per-tensor absmax scales of every quantized module's input and output with a momentum
of 0.9, and ``freeze`` fixes the weights.  Restated from quanto 0.2.6 as remembered (parity unpinned, DESIGN.md
section 4): activations are quantized per tensor with the calibrated ``input_scale`` (or arrive already quantized from a
``QLayerNorm``, whose codes and scale are then used as they are), qint8 x qint8 runs as an int8 GEMM with
``int32 * (input_scale * weight_scale)`` (+ bias), every other combination as the dequantize-then-matmul quanto falls
back to, and every quantized module's output is snapped to its ``output_scale`` grid.
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


class _CalibrationState:
    active = False
    momentum = 0.9


_CAL = _CalibrationState()


def calibrating() -> bool:
    return _CAL.active


def _qmax(qt: "qtype") -> float:
    return 448.0 if qt.is_floating_point else 127.0


def _absmax_scale(t: torch.Tensor, qt: "qtype") -> torch.Tensor:
    """quanto absmax_scale(base, qtype, axis=None): max|t| / qmax, evaluated in the tensor's dtype (calibration only:
    plain torch reductions, not part of the inference path)."""
    return t.detach().abs().max() / _qmax(qt)


def _update_scale(buf: torch.Tensor, new: torch.Tensor) -> None:
    """Calibration.update_scale: the first observation replaces the initial 1, later ones are blended in."""
    new = new.to(buf.dtype)
    if bool(torch.all(buf == 1)):
        buf.copy_(new)
    else:
        buf.copy_(_CAL.momentum * buf + (1.0 - _CAL.momentum) * new)


def _quantize_output(mod, y: torch.Tensor, want_codes: bool) -> torch.Tensor:
    """quantize_output hook of a quanto module with activations: y snapped to the output_scale grid.  The returned
    float tensor is the dequantized value (what the next float op sees); the codes travel with it as an attribute so
    that a quantized consumer uses them and their scale unchanged, as quanto does with an ActivationQBytesTensor."""
    qt = mod.activation_qtype
    if _CAL.active:
        _update_scale(mod.output_scale, _absmax_scale(y, qt))
    s = mod.output_scale.detach().float().view(1)
    codes, grid, deq = F.quant_act_static(y, s, qt.name, codes=want_codes and not qt.is_floating_point, grid=want_codes,
                                          deq=True)
    if want_codes:
        deq._quanto_q = (qt.name, codes, grid, s)
    return deq


class QLayerNorm(nn.LayerNorm):
    """optimum.quanto.nn.QLayerNorm: a float LayerNorm whose OUTPUT is quantized with a calibrated per-tensor scale
    (created by quantize() only when activations are quantized)."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, bias=True, device=None, dtype=None,
                 activations=None):
        super().__init__(normalized_shape, eps, elementwise_affine, bias, device, dtype)
        self.weight_qtype = None
        self.activation_qtype = activations
        ref = self.weight if self.weight is not None else torch.ones((), dtype=dtype or torch.float32)
        self.register_buffer("input_scale", torch.ones((), dtype=ref.dtype, device=ref.device))
        self.register_buffer("output_scale", torch.ones((), dtype=ref.dtype, device=ref.device))

    @classmethod
    def from_module(cls, module: nn.LayerNorm, activations=None):
        q = cls(module.normalized_shape, module.eps, module.elementwise_affine, module.bias is not None,
                device=module.weight.device, dtype=module.weight.dtype, activations=activations)
        q.weight, q.bias = module.weight, module.bias
        return q

    def freeze(self):
        return

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = nn.functional.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if self.activation_qtype is None:
            return y
        if not y.is_cuda:
            raise RuntimeError("QLayerNorm is on the CPU: activation quantization runs in the sm_100a library")
        return _quantize_output(self, y, want_codes=True)


class QLinear(nn.Linear):
    """optimum.quanto.nn.QLinear: qint8 / qint4 / qint2 / qfloat8 weights, optional static qint8 / qfloat8 activations.

    state_dict after freeze (quanto 0.2.6 layout): ``weight._data`` int8 [N, K],
    ``weight._scale`` [N, 1], ``bias``, ``input_scale``, ``output_scale``."""

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, weights=None,
                 activations=None, optimizer=None, quantize_input=False):
        super().__init__(in_features, out_features, bias, device, dtype)
        if activations is not None and getattr(activations, "name", None) not in ("qint8", "qfloat8_e4m3fn"):
            raise NotImplementedError(f"quanto activations={activations}: qint8 and qfloat8 are implemented")
        if weights is not None and getattr(weights, "name", None) not in ("qint8", "qint4", "qint2", "qfloat8_e4m3fn"):
            raise NotImplementedError(f"quanto weights={weights} is not implemented; qint8, qint4, qint2 and qfloat8 are")
        if activations is not None and weights is None:
            raise NotImplementedError("activation quantization without weight quantization is not implemented")
        if weights is not None and weights.name in ("qint4", "qint2") and in_features % 64 != 0:
            raise NotImplementedError("qint4 / qint2 weights need in_features % 64 == 0 in the fused GEMM")
        self.weight_qtype = weights
        self.activation_qtype = activations
        self.optimizer = optimizer
        self.register_buffer("input_scale", torch.ones((), dtype=self.weight.dtype, device=self.weight.device))
        self.register_buffer("output_scale", torch.ones((), dtype=self.weight.dtype, device=self.weight.device))
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
        elif self.weight_qtype.is_floating_point:
            q, scale = F.quanto_quantize_qfloat8(w)          # e4m3fn codes as uint8
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
        if self.activation_qtype is not None:
            return self._forward_quantized_activations(x, bias)
        if self.weight_qtype.is_floating_point:      # qfloat8 weights, float activations
            if x.dtype == torch.float32:
                return F.gemm_wf8a16(x, self._wq, self._wscale.view(-1), bias, torch.float32)
            return F.gemm_wf8a16(x, self._wq, self._wscale.view(-1), bias)
        if self._wshift is not None:      # qint4
            if x.dtype == torch.float32:
                return F.gemm_u4a16(x, self._wq, self._wscale, self._wshift, self._group, bias, torch.float32)
            return F.gemm_u4a16(x, self._wq, self._wscale, self._wshift, self._group, bias)
        if x.dtype == torch.float32:
            # fp32 flow of the reference (model never .half()-ed, model_utils.py:139-142): operands go
            # to the tensor cores as fp16, accumulate fp32, result written fp32 (DESIGN.md "Numerics")
            # (functional casts once for the tensor-core GEMM; decode-shaped calls take the fp32-in GEMV instead)
            return F.gemm_w8a16(x, self._wq, self._wscale, bias, torch.float32)
        return F.gemm_w8a16(x, self._wq, self._wscale, bias)


def _qlinear_forward_quantized_activations(self, x: torch.Tensor, bias) -> torch.Tensor:
    """QLinear.forward when activations are quantized (static scales from Calibration)."""
    aq, wq_t = self.activation_qtype, self.weight_qtype
    int_mm = (not aq.is_floating_point) and wq_t.name == "qint8"
    four_bit = self._wshift is not None
    carried = getattr(x, "_quanto_q", None)
    if carried is not None and carried[0] == aq.name:
        _, codes, grid, s_in = carried                   # already quantized by the producer: used as it is
        deq = x
        if _CAL.active:
            self.input_scale.copy_(s_in.view(()).to(self.input_scale.dtype))
        if int_mm and codes is None:
            codes = grid.to(torch.int8)
    else:
        if _CAL.active:
            _update_scale(self.input_scale, _absmax_scale(x, aq))
        s_in = self.input_scale.detach().float().view(1)
        codes, grid, deq = F.quant_act_static(x, s_in, aq.name, codes=int_mm, grid=not int_mm and not four_bit,
                                              deq=four_bit)
    dt = x.dtype
    if four_bit:
        # quanto dequantizes both operands for 4-bit weights: plain weight-only GEMM on the dequantized activations
        a = deq if dt in (torch.float16, torch.bfloat16) else deq.to(torch.float16)
        y = F.gemm_u4a16(a, self._wq, self._wscale, self._wshift, self._group, bias,
                         None if dt in (torch.float16, torch.bfloat16) else torch.float32)
    else:
        # output_scales = input._scale * weight._scale, formed in the dtype of the scales (the model's)
        os_ = (s_in.to(dt) * self._wscale.view(-1).to(dt)).float()
        if int_mm:
            y = F.gemm_w8a8(codes, self._wq, os_, bias, dt)
        else:
            od = torch.float16 if dt == torch.float16 else torch.float32
            if wq_t.is_floating_point:
                y = F.gemm_wf8a16(grid, self._wq, os_, bias, od)
            else:
                y = F.gemm_w8a16(grid, self._wq, os_, bias, od)
            y = y.to(dt)
    return _quantize_output(self, y, want_codes=False)


QLinear._forward_quantized_activations = _qlinear_forward_quantized_activations


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
        if type(m) is nn.LayerNorm and activations is not None:
            q = QLayerNorm.from_module(m, activations=activations)
        elif type(m) is nn.Linear:
            q = QLinear.from_module(m, weights=weights, activations=activations, optimizer=optimizer)
        else:
            continue
        parent = model
        *path, leaf = name.split(".")
        for p in path:
            parent = getattr(parent, p)
        setattr(parent, leaf, q)


def freeze(model: nn.Module):
    """optimum.quanto.freeze."""
    for m in model.modules():
        if isinstance(m, (QLinear, QLayerNorm)):
            m.freeze()


class Calibration:
    """optimum.quanto.Calibration(momentum=0.9, streamline=True, debug=False): while the context is active every
    quantized module records the per-tensor absmax scale of its input and of its output (first observation, then
    an exponential moving average) and keeps using the current scales to quantize."""

    def __init__(self, *args, momentum: float = 0.9, streamline: bool = True, debug: bool = False):
        self.momentum = float(momentum)

    def __enter__(self):
        self._prev = (_CAL.active, _CAL.momentum)
        _CAL.active, _CAL.momentum = True, self.momentum
        return self

    def __exit__(self, *exc):
        _CAL.active, _CAL.momentum = self._prev
        return False

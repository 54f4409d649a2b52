"""GPU twin of ``torch.ao.nn.quantized.dynamic.Linear`` -- the module that
``torch.quantization.quantize_dynamic(model, {torch.nn.Linear}, dtype=torch.qint8, inplace=True)``
(model_utils.py:131-134; pruning+quantization/pytorch_implementation.py:657-665) swaps in.

torch implements it on the CPU only (FBGEMM / oneDNN); that CPU path is the reference's reported
baseline (bench.py --impl reference).  This twin reproduces the same integer arithmetic on
sm_100a -- per-tensor symmetric int8 weights (scale = max|w| / 127.5), per-call per-tensor affine
uint8 activations with reduce_range, u8 x s8 -> s32 on the tensor cores, fp32 requantisation --
so that BASELINE.json config 1 can be token-compared between CPU and GPU.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as F


class DynamicInt8Linear(nn.Module):
    """Same surface as torch.ao.nn.quantized.dynamic.Linear: ``weight()``, ``bias()``, ``scale``,
    ``zero_point``, ``in_features``, ``out_features``; not an nn.Linear subclass (neither is
    torch's)."""

    _FLOAT_MODULE = nn.Linear

    def __init__(self, in_features: int, out_features: int, bias_: bool = True, dtype=torch.qint8):
        super().__init__()
        if dtype != torch.qint8:
            raise NotImplementedError("only dtype=torch.qint8 is implemented (the reference's setting)")
        self.in_features, self.out_features = in_features, out_features
        self.scale, self.zero_point = 1.0, 0
        self.register_buffer("w_int", torch.zeros((out_features, in_features), dtype=torch.int8))
        self.register_buffer("w_scale", torch.ones((1,), dtype=torch.float32))
        self.register_buffer("w_sum", torch.zeros((out_features,), dtype=torch.int32))
        self.register_buffer("bias_f32", torch.zeros((out_features,), dtype=torch.float32) if bias_ else None)
        self._float_weight = None   # kept until the module reaches a CUDA device

    @classmethod
    def from_float(cls, mod: nn.Linear) -> "DynamicInt8Linear":
        q = cls(mod.in_features, mod.out_features, mod.bias is not None)
        if mod.bias is not None:
            q.bias_f32 = mod.bias.detach().float().clone()
        w = mod.weight.detach().float()
        if w.is_cuda:
            q._set_weight(w)
        else:
            q._float_weight = w.clone()
        return q

    def _set_weight(self, w: torch.Tensor):
        self.w_int, self.w_scale, self.w_sum = F.torch_quantize_weight(w.contiguous())
        self._float_weight = None

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        if self._float_weight is not None:
            dev = self.w_int.device
            if dev.type == "cuda":
                self._set_weight(self._float_weight.to(dev))
        return self

    def weight(self) -> torch.Tensor:
        """Dequantised view (torch returns a quantized tensor; .int_repr() is w_int here)."""
        return self.w_int.float() * self.w_scale

    def bias(self):
        return self.bias_f32

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._float_weight is not None or not self.w_int.is_cuda:
            raise RuntimeError("DynamicInt8Linear is on the CPU: move it to a CUDA device (there is no CPU "
                               "path here; the CPU implementation is torch's own quantize_dynamic)")
        xq, qparams = F.torch_quantize_activation(x.float() if x.dtype != torch.float32 else x)
        return F.gemm_dyn_i8(xq, qparams, self.w_int, self.w_scale, self.w_sum, self.bias_f32)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, dtype=torch.qint8, device=cuda"


def quantize_dynamic(model: nn.Module, qconfig_spec=None, dtype=torch.qint8, mapping=None, inplace: bool = False):
    """Same call shape as torch.quantization.quantize_dynamic; swaps nn.Linear -> DynamicInt8Linear."""
    if not inplace:
        import copy
        model = copy.deepcopy(model)
    if qconfig_spec is not None and nn.Linear not in qconfig_spec:
        return model
    for name, m in list(model.named_modules()):
        if type(m) is not nn.Linear:
            continue
        parent = model
        *path, leaf = name.split(".")
        for p in path:
            parent = getattr(parent, p)
        setattr(parent, leaf, DynamicInt8Linear.from_float(m))
    return model

"""Torch-op emulations of the reference quantized linears, used ONLY by tests to compare whole
models on the same device / dtype (so that everything except the linear under test is the same
HF code).  Each follows the oracle's restatement (oracle/whisperq_oracle.c) operation by
operation; float64 is used where the oracle uses fmaf so that the result is the same fp32 value
except in exact double-rounding ties."""
import os

import numpy as np
import torch
from torch import nn

_TABLE = {}


def bnb_row_scale(absmax: torch.Tensor) -> torch.Tensor:
    """127 / absmax as bitsandbytes' kernel forms it (__fdividef), through the table dumped on a B200
    (tests/golden/fdividef_127_fp16.npz); absmax holds fp16 values."""
    t = _TABLE.get(absmax.device)
    if t is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fdividef_127_fp16.npz")
        t = _TABLE[absmax.device] = torch.from_numpy(np.load(path)["table"]).to(absmax.device)
    return t[absmax.half().view(torch.int16).long() & 0xffff]


def _set(model, name, new):
    parent = model
    *path, leaf = name.split(".")
    for p in path:
        parent = getattr(parent, p)
    setattr(parent, leaf, new)


class EmuLinear8bitLt(nn.Module):
    """bnb Linear8bitLt (threshold, has_fp16_weights=False) with torch ops."""

    def __init__(self, lin: nn.Linear, threshold: float):
        super().__init__()
        W = lin.weight.detach().half().float()
        self.SCB = W.abs().amax(1)
        self.CB = torch.nan_to_num(torch.round(W * bnb_row_scale(self.SCB)[:, None])).to(torch.int8)
        self.bias = None if lin.bias is None else lin.bias.detach().half()
        self.threshold = threshold

    def forward(self, x):
        A = x.half().reshape(-1, x.shape[-1])
        Af = A.float()
        out = Af.abs() >= self.threshold
        am = torch.where(out, torch.zeros_like(Af), Af.abs()).amax(1)
        CA = torch.nan_to_num(torch.round(Af * bnb_row_scale(am)[:, None]))
        CA = torch.where(out, torch.zeros_like(CA), CA)
        cols = out.any(0)
        CA[:, cols] = 0
        c32 = (CA.double() @ self.CB.double().t())          # exact integers
        v = (c32.float() * am[:, None]) * self.SCB[None, :]
        y = v.double() * float(torch.tensor(6.200012e-05, dtype=torch.float32))
        if self.bias is not None:
            y = y + self.bias.double()[None, :]
        y = y.float().half()
        if cols.any():
            subB = ((self.CB[:, cols].float() * self.SCB[:, None]) * 7.874015718698502e-3).half()
            y = (y.float() + A[:, cols].float() @ subB.float().t()).half()
        return y.reshape(*x.shape[:-1], -1).to(x.dtype)


class EmuDequantLinear(nn.Module):
    """F.linear(x, W_dequantised.to(x.dtype), bias): bnb Linear4bit forward for M > 1 and the
    quanto fallback `matmul(x, Wq.to(x.dtype).t()) * scale + bias`."""

    def __init__(self, w_deq: torch.Tensor, bias, post_scale=None):
        super().__init__()
        self.w, self.bias, self.post_scale = w_deq, bias, post_scale

    def forward(self, x):
        y = nn.functional.linear(x, self.w.to(x.dtype))
        if self.post_scale is not None:
            y = y * self.post_scale.to(x.dtype)
        if self.bias is not None:
            y = y + self.bias.to(x.dtype)
        return y


def swap_all(model, make, skip=("proj_out",)):
    for name, m in list(model.named_modules()):
        if type(m) is nn.Linear and name.split(".")[-1] not in skip:
            _set(model, name, make(m))
    return model

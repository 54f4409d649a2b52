"""Module swaps and import shims -- the drop-in boundary on the reference side.

``install_shims()`` registers this package's mirrors as ``bitsandbytes`` and ``optimum.quanto``
(and a minimal ``evaluate``) in ``sys.modules`` when the real packages are absent, so that the
reference's ``model_utils.py`` / ``evaluation.py`` / ``data_utils.py`` import and run unchanged
(model_utils.py:6-7 imports optimum.quanto at module import time).

``replace_linears`` mirrors the swaps the reference performs:
  * ``convert_model_to_4bit`` (pruning+quantization/bnb_implementation.py:1093-1118): every
    nn.Linear incl. proj_out -> Linear4bit, original weights loaded, quantized on .to(device);
  * HF ``replace_with_bnb_linear`` (transformers/integrations/bitsandbytes.py:157-231) as reached
    from model_utils.py:112-118 / BASELINE.json config 2: every nn.Linear except the output
    embedding (``proj_out``) -> Linear4bit / Linear8bitLt.
"""
from __future__ import annotations

import sys
import types
from typing import Callable, Iterable, Optional

import torch
from torch import nn

from . import bnb, quanto


def _set_submodule(model: nn.Module, name: str, new: nn.Module) -> None:
    parent = model
    *path, leaf = name.split(".")
    for p in path:
        parent = getattr(parent, p)
    setattr(parent, leaf, new)


def replace_linears(model: nn.Module, factory: Callable[[nn.Linear], nn.Module],
                    skip: Iterable[str] = ()) -> int:
    """Replace every exact nn.Linear (not subclasses) whose leaf name is not in `skip`."""
    skip = set(skip)
    n = 0
    for name, m in list(model.named_modules()):
        if type(m) is not nn.Linear or name.split(".")[-1] in skip:
            continue
        new = factory(m)
        new.load_state_dict(m.state_dict(), strict=False)
        _set_submodule(model, name, new)
        n += 1
    return n


def convert_model_to_4bit(model: nn.Module, compute_dtype=torch.float32, quant_type: str = "nf4",
                          double_quant: bool = False) -> nn.Module:
    """Same behaviour as the reference's helper of the same name (bnb_implementation.py:1093-1118)."""
    replace_linears(model, lambda m: bnb.Linear4bit(m.in_features, m.out_features, bias=m.bias is not None,
                                                    compute_dtype=compute_dtype, quant_type=quant_type,
                                                    compress_statistics=double_quant, device=None))
    return model


def replace_with_bnb_linear(model: nn.Module, load_in_8bit: bool = False, load_in_4bit: bool = False,
                            llm_int8_threshold: float = 6.0, bnb_4bit_compute_dtype=torch.float16,
                            bnb_4bit_quant_type: str = "nf4", bnb_4bit_use_double_quant: bool = False,
                            modules_to_not_convert: Optional[Iterable[str]] = ("proj_out",)) -> nn.Module:
    """HF transformers' BitsAndBytes module replacement (output embedding kept in floating point)."""
    skip = tuple(modules_to_not_convert or ())
    if load_in_8bit:
        replace_linears(model, lambda m: bnb.Linear8bitLt(m.in_features, m.out_features, m.bias is not None,
                                                          has_fp16_weights=False, threshold=llm_int8_threshold),
                        skip)
    elif load_in_4bit:
        replace_linears(model, lambda m: bnb.Linear4bit(m.in_features, m.out_features, m.bias is not None,
                                                        compute_dtype=bnb_4bit_compute_dtype,
                                                        compress_statistics=bnb_4bit_use_double_quant,
                                                        quant_type=bnb_4bit_quant_type), skip)
    return model


# ----------------------------------------------------------------------------------------------
# import shims
# ----------------------------------------------------------------------------------------------
def _importable(name: str) -> bool:
    import importlib.util
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install_shims(force: bool = False) -> dict:
    """Make `import bitsandbytes`, `from optimum.quanto import ...` and `import evaluate` resolve to
    this package when the real ones are not installed.  Returns {name: installed?}."""
    done = {}
    if force or not _importable("bitsandbytes"):
        pkg = types.ModuleType("bitsandbytes")
        pkg.__version__ = "0.45.0+whisperq"
        nn_mod = types.ModuleType("bitsandbytes.nn")
        for k in ("Linear4bit", "LinearNF4", "LinearFP4", "Params4bit", "Linear8bitLt", "Int8Params"):
            setattr(nn_mod, k, getattr(bnb, k))
        fn_mod = types.ModuleType("bitsandbytes.functional")
        for k in ("quantize_4bit", "dequantize_4bit", "quantize_nf4", "quantize_fp4", "int8_vectorwise_quant",
                  "int8_vectorwise_dequant", "QuantState"):
            setattr(fn_mod, k, getattr(bnb, k))
        pkg.nn, pkg.functional = nn_mod, fn_mod
        pkg.matmul_4bit = bnb.matmul_4bit
        pkg.MatmulLtState = bnb.MatmulLtState
        sys.modules["bitsandbytes"] = pkg
        sys.modules["bitsandbytes.nn"] = nn_mod
        sys.modules["bitsandbytes.functional"] = fn_mod
        done["bitsandbytes"] = True
    if force or not _importable("optimum"):
        opt = types.ModuleType("optimum")
        q = types.ModuleType("optimum.quanto")
        for k in ("Calibration", "freeze", "quantize", "qint2", "qint4", "qint8", "qfloat8", "QLinear", "QLayerNorm",
                  "qtype"):
            setattr(q, k, getattr(quanto, k))
        opt.quanto = q
        sys.modules["optimum"] = opt
        sys.modules["optimum.quanto"] = q
        done["optimum.quanto"] = True
    if force or not _importable("evaluate"):
        from . import tally
        ev = types.ModuleType("evaluate")
        ev.load = tally.load_metric
        sys.modules["evaluate"] = ev
        done["evaluate"] = True
    return done

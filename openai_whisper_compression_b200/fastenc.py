"""Copy-free encoder self-attention behind HF's ``WhisperAttention`` (SURVEY.md section 8f rank 3:
the elementwise/layout passes over ``[B*1500, d]`` activations around the quantized linears).

HF's ``WhisperAttention.forward`` (modeling_whisper.py:310-355 in transformers 5.5) materialises
``q/k/v.transpose(1, 2).contiguous()`` and ``attn_output.transpose(1, 2).contiguous()``: four strided
copies of a ``[B, 1500, d]`` tensor per layer (~0.55 ms each at B = 256, whisper-base), more time than the
layer's six quantized GEMMs together.  ``scaled_dot_product_attention`` (cuDNN flash kernel on sm_100)
reads the ``[B, S, H, D]`` projections through their strides and returns its output in the same
layout, so the copies are pure overhead.  ``enable(model)`` gives every encoder self-attention module a
forward that keeps the projections where the drop-in linears wrote them; the arithmetic (projection,
``* scaling``, SDPA with ``scale=1.0``, out_proj) and its order are HF's.  Anything else (masks, caches,
``output_attentions``, eager attention, CPU) falls through to HF's forward.

When every linear of an encoder layer is a bitsandbytes-style ``Linear8bitLt`` (fp16), the whole
``WhisperEncoderLayer.forward`` is replaced by the producer-fused sequence (``_layer_forward_int8``): torch's
LayerNorm kernel runs at ~1.1 TB/s on ``[B*1500, d]`` and ``Linear8bitLt`` re-quantizes the same LayerNorm
output three times for q/k/v; add+LayerNorm+quant in one pass, one q/k/v GEMM and GELU+quant remove ~30 % of
the layer's HBM traffic.
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as TF

from . import functional as F
from . import fused


def _self_attn_forward(self, hidden_states, key_value_states=None, past_key_values=None, attention_mask=None,
                       output_attentions=False, **kwargs):
    if (key_value_states is not None or past_key_values is not None or attention_mask is not None
            or output_attentions or self.training or not hidden_states.is_cuda
            or self.config._attn_implementation != "sdpa" or getattr(self, "is_causal", False)):
        return self._whisperq_hf_forward(hidden_states, key_value_states=key_value_states,
                                         past_key_values=past_key_values, attention_mask=attention_mask,
                                         output_attentions=output_attentions, **kwargs)
    B, S = hidden_states.shape[:-1]
    shape = (B, S, -1, self.head_dim)
    q = (self.q_proj(hidden_states) * self.scaling).view(shape).transpose(1, 2)
    k = self.k_proj(hidden_states).view(shape).transpose(1, 2)
    v = self.v_proj(hidden_states).view(shape).transpose(1, 2)
    o = TF.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, scale=1.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, S, -1)      # a view when SDPA kept q's [B, S, H, D] layout
    return self.out_proj(o), None


def _plan_layer_int8(layer):
    """Packed weights for _layer_forward_int8, or None when the layer's linears are not all drop-in modules of one
    scheme fused.pack serves for the layer's dtype (LLM.int8 / fp16; W8A16, NF4 / FP4, qint4 / qint2 with fp16 or bf16)."""
    sa = layer.self_attn
    cfg = sa.config
    dtype = layer.self_attn_layer_norm.weight.dtype
    if (cfg.activation_function != "gelu" or sa.head_dim != 64 or layer.embed_dim > 2048 or layer.embed_dim % 8
            or dtype not in (torch.float16, torch.bfloat16)):
        return None
    plan = types.SimpleNamespace()
    plan.qkv = fused.pack([sa.q_proj, sa.k_proj, sa.v_proj], dtype)
    plan.o = fused.pack([sa.out_proj], dtype)
    plan.fc1 = fused.pack([layer.fc1], dtype)
    plan.fc2 = fused.pack([layer.fc2], dtype)
    ws = [plan.qkv, plan.o, plan.fc1, plan.fc2]
    if any(w is None for w in ws) or any(w.threshold != ws[0].threshold or w.kind != ws[0].kind for w in ws):
        return None
    plan.dtype = dtype
    plan.threshold = ws[0].threshold if ws[0].kind == "int8" else None     # None: no int8 rows from the producers
    plan.scaling = float(sa.scaling)
    plan.scaling_pow2 = fused.is_pow2(plan.scaling)
    return plan


def _layer_forward_int8(self, hidden_states, attention_mask=None, **kwargs):
    """WhisperEncoderLayer.forward (modeling_whisper.py:380-414) for a layer whose linears are all drop-in modules of
    one scheme: q/k/v are one GEMM over the concatenated weights, SDPA reads the fused projection through its strides,
    the residual add rides in the LayerNorm launch; for LLM.int8 every GEMM is fed by a producer kernel that already
    wrote its int8 rows (add+LayerNorm+quant, GELU+quant).  Per-linear arithmetic is unchanged; LayerNorm / GELU are
    within one fp16 ulp of torch's (tests/test_gpu_fused.py)."""
    plan = getattr(self, "_whisperq_plan", False)
    if plan is False:
        plan = self._whisperq_plan = _plan_layer_int8(self)
    if (plan is None or attention_mask is not None or self.training or kwargs.get("output_attentions")
            or hidden_states.dtype != plan.dtype or not hidden_states.is_cuda
            or self.self_attn.config._attn_implementation != "sdpa"):
        return self._whisperq_hf_forward(hidden_states, attention_mask, **kwargs)
    B, S, d = hidden_states.shape
    H, thr = self.self_attn.num_heads, plan.threshold
    x = hidden_states.reshape(B * S, d)
    ln = self.self_attn_layer_norm
    _, h, qt = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps, thr)
    qkv = fused.gemm(qt, h, plan.qkv)
    q, k, v = (qkv[:, i * d:(i + 1) * d].view(B, S, H, 64).transpose(1, 2) for i in range(3))
    if plan.scaling_pow2:       # q * 2^-k is exact in fp16 / bf16, so the scale rides in the SDPA call
        a = TF.scaled_dot_product_attention(q, k, v, scale=plan.scaling)
    else:
        a = TF.scaled_dot_product_attention(q * plan.scaling, k, v, scale=1.0)
    a = a.transpose(1, 2).reshape(B * S, d)
    att = fused.gemm(F.int8_vectorwise_quant(a, thr, finalize=False) if thr is not None else None, a, plan.o)
    ln = self.final_layer_norm
    x, h, qt = F.add_layernorm_quant(x, att, ln.weight, ln.bias, ln.eps, thr)
    clamp_value = torch.finfo(torch.float16).max - 1000
    if plan.qkv.kind == "int8":
        # The fp16 GELU output is only ever read for outlier columns: it is not stored; fc2 takes fc1's output and
        # re-derives those entries (a_pre_gelu).  Residual add + fp16 clamp ride in fc2's epilogue (bit-identical to
        # the two torch kernels they replace).
        f1 = fused.gemm(qt, h, plan.fc1)
        _, qt = F.gelu_quant(f1, thr, store_h=False)
        out = fused.gemm_int8(qt, f1, plan.fc2, residual=x, clamp_abs=float(clamp_value), a_pre_gelu=True)
        return out.view(B, S, d)
    g, qt = F.gelu_quant(fused.gemm(qt, h, plan.fc1), thr)
    out = x + fused.gemm(qt, g, plan.fc2)
    if out.dtype == torch.float16:      # HF clamps fp16 activations only
        out = torch.clamp(out, min=-clamp_value, max=clamp_value)
    return out.view(B, S, d)


def _contiguous_stream_hook(module, args, kwargs):
    """HF builds the residual stream as ``conv2(...).permute(0, 2, 1) + embed_positions`` (modeling_whisper.py,
    WhisperEncoder.forward): the sum inherits the conv's channels-first strides, every residual add keeps
    them, and every LayerNorm then makes its own contiguous copy (13 strided copies of [B, 1500, d] per
    encoder pass).  One copy in front of layer 0 puts the stream in row-major order for the whole stack;
    values are untouched."""
    if args and isinstance(args[0], torch.Tensor) and not args[0].is_contiguous():
        return (args[0].contiguous(),) + tuple(args[1:]), kwargs
    if "hidden_states" in kwargs and not kwargs["hidden_states"].is_contiguous():
        kwargs = dict(kwargs, hidden_states=kwargs["hidden_states"].contiguous())
        return args, kwargs
    return None


def enable(model, fuse_int8: bool = True) -> int:
    """Patch the encoder's self-attention modules of an HF Whisper model (idempotent).
    Returns the number of modules patched."""
    n = 0
    first = model.model.encoder.layers[0]
    if not hasattr(first, "_whisperq_stream_hook"):
        first._whisperq_stream_hook = first.register_forward_pre_hook(_contiguous_stream_hook, with_kwargs=True)
    for layer in model.model.encoder.layers:
        attn = layer.self_attn
        if not hasattr(attn, "_whisperq_hf_forward"):
            attn._whisperq_hf_forward = attn.forward
            attn.forward = types.MethodType(_self_attn_forward, attn)
        if fuse_int8 and not hasattr(layer, "_whisperq_hf_forward"):
            layer._whisperq_hf_forward = layer.forward
            layer.forward = types.MethodType(_layer_forward_int8, layer)
        n += 1
    return n


def disable(model) -> None:
    first = model.model.encoder.layers[0]
    if hasattr(first, "_whisperq_stream_hook"):
        first._whisperq_stream_hook.remove()
        del first._whisperq_stream_hook
    for layer in model.model.encoder.layers:
        for mod in (layer.self_attn, layer):
            if hasattr(mod, "_whisperq_hf_forward"):
                mod.forward = mod._whisperq_hf_forward
                del mod._whisperq_hf_forward
        if hasattr(layer, "_whisperq_plan"):
            del layer._whisperq_plan

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
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as TF


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


def enable(model) -> int:
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
        n += 1
    return n


def disable(model) -> None:
    first = model.model.encoder.layers[0]
    if hasattr(first, "_whisperq_stream_hook"):
        first._whisperq_stream_hook.remove()
        del first._whisperq_stream_hook
    for layer in model.model.encoder.layers:
        attn = layer.self_attn
        if hasattr(attn, "_whisperq_hf_forward"):
            attn.forward = attn._whisperq_hf_forward
            del attn._whisperq_hf_forward

"""CUDA-graph greedy decoding behind ``model.generate`` (the decode
step's host/launch overhead).

``transcribe_batch`` (data_utils.py:152) calls ``model.generate(features)``.  HF's generate prepares
everything -- encoder pass, initial decoder tokens, the Whisper logits processors, stopping criteria
-- and then runs ``model._sample``, a Python loop that re-dispatches several hundred tiny kernels
per token.  ``enable(model)`` replaces only that inner loop: the decoder step (the model's own
sub-modules, i.e. the drop-in quantized linears, LayerNorms and embeddings, with a static KV cache)
is captured once per (batch, max length) as a CUDA graph and replayed per token; HF's own
``logits_processor`` and ``stopping_criteria`` objects are applied between replays, so the call
``model.generate(features)`` and its result format are unchanged.  The generated ids are handed to HF's
per-utterance post-processing as a host tensor (one read-back instead of several blocking syncs per
utterance) and returned to the model's device at the end of ``generate``.  Anything the fast loop does not
cover (sampling, beams, return_dict_in_generate, streamer, CPU) falls through to HF's ``_sample``.

Numerics: same modules and weights as HF's loop; attention over the static cache is
``scaled_dot_product_attention`` with a position mask instead of HF's exact-length SDPA call, so
logits can differ in the last fp16 bits (tests compare tokens with HF's loop).
"""
from __future__ import annotations

import os
import types
from typing import Dict, Tuple

import torch
import torch.nn.functional as TF

from . import _lib, fastenc, fused
from . import functional as F


class _State:
    pass


class _FastReturn(Exception):
    """Carries the finished token ids out of HF's generate stack (see GraphedGreedy._maybe_unwind)."""

    def __init__(self, sequences: torch.Tensor):
        super().__init__("whisperq fast return")
        self.sequences = sequences


# keyword arguments of WhisperGenerationMixin.generate that must keep their defaults for the short post-processing
# (generation_whisper.py:383-412); everything else in **kwargs must be one of _PLAIN_KWARGS
_WHISPER_DEFAULTS = {"generation_config": None, "logits_processor": None, "stopping_criteria": None,
                     "prefix_allowed_tokens_fn": None, "synced_gpus": False, "return_timestamps": None,
                     "prompt_ids": None, "prompt_condition_type": None, "condition_on_prev_tokens": None,
                     "temperature": None, "compression_ratio_threshold": None, "logprob_threshold": None,
                     "no_speech_threshold": None, "num_segment_frames": None, "attention_mask": None,
                     "return_token_timestamps": None, "return_segments": False, "return_dict_in_generate": None,
                     "force_unique_generate_call": None, "monitor_progress": None}
_PLAIN_KWARGS = {"input_features", "task", "language", "is_multilingual", "time_precision", "time_precision_features",
                 "do_sample", "num_beams", "max_new_tokens", "min_new_tokens", "max_length", "min_length", "use_cache"}


class GraphedGreedy:
    def __init__(self, model, len_bucket: int = 64):
        self.model = model
        self.len_bucket = len_bucket
        self._states: Dict[Tuple[int, int, torch.dtype], _State] = {}
        self._orig_sample = None
        self.max_states = 4
        self.host_postprocess = True
        self.fuse_int8 = True          # producer-fused decode step for all-Linear8bitLt decoders
        self.own_attention = True      # attn_decode.cu instead of index_copy + mask + SDPA
        # cross-attention at decode time: "own" (attn_decode.cu), "cudnn" (torch SDPA) or "auto".  Both stream K/V at
        # the HBM rate (scripts/cross_attn_bench.py: 6.7 TB/s at B = 256; own 5.3 vs 4.4 TB/s at B = 64); inside the
        # step graph the own kernel's 128-thread CTAs leave a longer tail once heads x utterances exceed ~1.4 waves,
        # so "auto" keeps cuDNN for those shapes (1135 vs 1190 us per step at B = 256, whisper-base)
        self.cross_attention = os.environ.get("WQ_CROSS_ATTN", "auto")
        # HF walks the finished batch utterance by utterance (generate_with_fallback, _retrieve_segment,
        # _pad_to_max_length: ~30 tiny tensor ops each, 12.5 ms at B = 256 with the GPU idle).  For the plain call
        # transcribe_batch makes (data_utils.py:152) that walk only strips trailing pad/eos tokens and re-pads:
        # done here in a few batched ops, same result (tests/test_host_logic.py)
        self.fast_post = os.environ.get("WQ_FAST_POST", "1") != "0"
        self.before_readback = None      # optional callable, see _sample: runs after a batch is queued, before the host blocks
        self._unwind = False
        self.keep_logits = False       # True: the captured step always writes the [B, V] logits (tests, debugging)
        # fused LLM.int8 step: row groups of the batch decoded on as many streams (1 = off).  Per token and layer the
        # step is a ~75 us chain of short dependent launches followed by a cross-attention pass that streams the
        # cached K/V at the HBM rate; with 4 groups a group's chain hides under the other groups' streams
        self.streams = int(os.environ.get("WQ_DECODE_STREAMS", "4"))
        self.min_rows_per_stream = int(os.environ.get("WQ_DECODE_MIN_ROWS", "16"))
        # out_proj's int8 rows written by the cross-attention kernel itself (its finisher warp) or by a separate
        # quantizer launch (WQ_CROSS_QUANT_INLINE=0)
        self.cross_quant_inline = os.environ.get("WQ_CROSS_QUANT_INLINE", "1") != "0"
        # persistent decoder-layer kernel (decode_fused.cu): one launch between two cross-attention passes instead of
        # twelve.  Bit-identical to the chain and OFF by default: measured SLOWER (64 rows, whisper-base: 153 us per
        # launch against ~50 us for the twelve PDL-chained launches it replaces; 146 vs 108 ms per bench step) -- eleven
        # grid barriers and eleven dependent L2 round trips cost more than programmatic dependent launch already hides
        self.mega = os.environ.get("WQ_DECODE_FUSED", "0") == "1"
        # Opt-in: vocabulary projection + greedy choice per row group on the group's own stream (no join before the
        # projection) instead of one projection of the whole batch after the join.  Measured SLOWER (whisper-base, 4
        # groups of 64: 1.131 vs 1.082 ms per token -- the 53 MB weight is then streamed once per group).
        self.group_project = os.environ.get("WQ_GROUP_PROJECT", "0") == "1"
        # Opt-in: route <= 16-row int8 GEMMs of the step through the small-row dp4a kernel instead of the tensor-core
        # tile.  Measured SLOWER on large-v3 (B=32, 2 row groups of 16: 485 vs 377 ms/step), so off by default.
        self.small_int8 = os.environ.get("WQ_SMALL_INT8", "0") == "1"
        self.time_loop = False         # bench.py: CUDA events around the token loop of every generate call
        self.loop_events = []          # (start, end, replays)
        self.replays = 0
        self.fallbacks = 0
        self.fast_returns = 0

    # ------------------------------------------------------------------------------------------
    def install(self):
        # HF dispatches with getattr(type(model), "_sample") (generation/utils.py), so the hook has to
        # live on the class: give this one model instance a subclass that overrides _sample.
        if self._orig_sample is None:
            base = type(self.model)
            self._orig_sample = types.MethodType(base._sample, self.model)

            def _sample(model_self, input_ids, **kwargs):
                return model_self._whisperq_fastgen._sample(input_ids, **kwargs)

            def generate(model_self, *args, **kwargs):
                eng = model_self._whisperq_fastgen
                eng._unwind = eng.fast_post and eng._plain_call(args, kwargs)
                try:
                    out = base.generate(model_self, *args, **kwargs)
                except _FastReturn as done:
                    out = done.sequences
                finally:
                    eng._unwind = False
                # the fast loop hands HF's per-utterance post-processing host-resident token ids (see
                # _sample); the caller gets them back where HF would have put them
                if isinstance(out, torch.Tensor) and out.device != model_self.device:
                    out = out.to(model_self.device)
                return out

            # HF prepares every 30 s window with per-utterance Python loops over device tensors
            # (generation_whisper.py:1813-1849); on the first pass of a short-form batch they change nothing
            def _maybe_reduce_batch(input_features, seek, max_frames, cur_bsz, batch_idx_map):
                if (isinstance(seek, torch.Tensor) and isinstance(max_frames, torch.Tensor) and seek.shape == max_frames.shape
                        and len(batch_idx_map) >= cur_bsz and not bool((seek >= max_frames).any())):
                    return input_features, cur_bsz, [batch_idx_map[i] for i in range(cur_bsz)]   # nobody is finished
                return base._maybe_reduce_batch(input_features, seek, max_frames, cur_bsz, batch_idx_map)

            def _get_input_segment(input_features, seek, seek_num_frames, num_segment_frames, cur_bsz, batch_idx_map):
                if (isinstance(input_features, torch.Tensor) and input_features.dim() == 3
                        and input_features.shape[0] == cur_bsz and input_features.shape[-1] == num_segment_frames
                        and isinstance(seek, torch.Tensor) and isinstance(seek_num_frames, torch.Tensor)
                        and not bool((seek != 0).any()) and bool((seek_num_frames >= num_segment_frames).all())):
                    return input_features           # every slice [0 : >= length] is the whole window: cat == input
                return base._get_input_segment(input_features, seek, seek_num_frames, num_segment_frames, cur_bsz,
                                               batch_idx_map)

            self._base_cls = base
            self.model.__class__ = type(base.__name__, (base,), {
                "_sample": _sample, "generate": generate, "_maybe_reduce_batch": staticmethod(_maybe_reduce_batch),
                "_get_input_segment": staticmethod(_get_input_segment), "__module__": base.__module__})
            self.model._whisperq_fastgen = self
            fastenc.enable(self.model)       # copy-free encoder self-attention (same arithmetic)
        return self

    def uninstall(self):
        if self._orig_sample is not None:
            self.model.__class__ = self._base_cls
            self._orig_sample = None
            del self.model._whisperq_fastgen
            fastenc.disable(self.model)

    # ------------------------------------------------------------------------------------------
    def _plain_call(self, args, kwargs) -> bool:
        """True when `model.generate(*args, **kwargs)` is the plain short-form call whose post-processing in
        WhisperGenerationMixin.generate reduces to stripping trailing pad/eos tokens: one <= 30 s window, no
        timestamps / prompts / temperature fallback / thresholds / dict outputs.  Anything else keeps HF's path."""
        if len(args) > 1:
            return False
        feats = args[0] if args else kwargs.get("input_features")
        cfg = self.model.config
        if not isinstance(feats, torch.Tensor) or feats.dim() != 3 or feats.shape[-1] > 2 * cfg.max_source_positions:
            return False
        for k, v in kwargs.items():
            if k in _WHISPER_DEFAULTS:
                if not (v is _WHISPER_DEFAULTS[k] or v == _WHISPER_DEFAULTS[k]):
                    return False
            elif k not in _PLAIN_KWARGS:
                return False
        gc = self.model.generation_config
        for name in ("return_timestamps", "return_dict_in_generate", "condition_on_prev_tokens",
                     "compression_ratio_threshold", "logprob_threshold", "no_speech_threshold",
                     "force_unique_generate_call", "output_scores", "output_logits", "output_attentions",
                     "output_hidden_states"):
            if getattr(gc, name, None):
                return False
        return (getattr(gc, "num_return_sequences", 1) or 1) == 1

    def _maybe_unwind(self, sequences: torch.Tensor, prompt_len: int, generation_config) -> None:
        """HF's post-processing of a finished plain call (generate_with_fallback :1063-1096, _retrieve_segment's
        no-timestamp branch :2051-2075, _pad_to_max_length :182-230 of generation_whisper.py), batched: drop the
        decoder prompt; a row ending in pad loses as many trailing tokens as it holds pads (one fewer when pad ==
        eos); a row then ending in eos loses it; rows are right-padded with pad to the longest.  Raises
        _FastReturn (caught by the `generate` override) when that is all HF would do; returns otherwise."""
        if not self._unwind or sequences.dim() != 2:
            return
        pad, eos = generation_config.pad_token_id, generation_config.eos_token_id
        if isinstance(pad, torch.Tensor):
            pad = pad.item() if pad.numel() == 1 else None
        if isinstance(eos, (list, tuple)):
            eos = eos[0] if len(eos) == 1 else None
        if isinstance(eos, torch.Tensor):
            eos = eos.item() if eos.numel() == 1 else None
        if not isinstance(pad, int) or not isinstance(eos, int):
            return
        seq = sequences.cpu()
        tok = seq[:, prompt_len:]
        B, T = tok.shape
        if T == 0:
            return
        ts_begin = (generation_config.no_timestamps_token_id + 1 if hasattr(generation_config, "no_timestamps_token_id")
                    else self.model.config.vocab_size + 1)
        if bool((tok >= ts_begin).any()):
            return                                  # timestamp tokens: HF cuts segments, keep its path
        n_pad = (tok == pad).sum(1)
        cut = torch.where(tok[:, -1] == pad, n_pad - (1 if pad == eos else 0), torch.zeros_like(n_pad))
        length = T - cut
        if bool((length <= 0).any()):
            return                                  # HF would index an empty row: let it
        last = tok.gather(1, (length - 1).view(B, 1)).view(B)
        length = length - (last == eos).long()
        width = int(length.max())
        keep = torch.arange(width).view(1, width) < length.view(B, 1)
        out = torch.where(keep, tok[:, :width], torch.full((), pad, dtype=tok.dtype))
        self.fast_returns += 1
        raise _FastReturn(out.contiguous())

    # ------------------------------------------------------------------------------------------
    def _decoder_step(self, st: _State):
        """One token for every utterance: reads st.tok / st.pos, writes st.logits / st.next."""
        if st.fused is not None:
            step = self._decoder_step_mega if st.mega else self._decoder_step_int8
            if len(st.views) == 1:
                return step(st, st.views[0])
            # several row groups of the batch on as many streams: a group's launch-bound chain (LayerNorm, decode-
            # shaped GEMMs, self-attention) runs under the HBM-bound cross-attention stream of another group.  Under
            # capture the fork / join become graph dependencies; each group has its own outlier / counter scratch.
            cur = torch.cuda.current_stream()
            for v in st.views[1:]:
                v.stream.wait_stream(cur)
            gp = self.group_project
            with F.scratch_slot(st.views[0].slot):
                step(st, st.views[0], project=gp)
            for v in st.views[1:]:
                with torch.cuda.stream(v.stream), F.scratch_slot(v.slot):
                    step(st, v, project=gp)
            for v in st.views[1:]:
                cur.wait_stream(v.stream)
            if gp:
                return None
            # one vocabulary projection for the whole batch (the 53 MB weight is streamed once per token)
            return self._project(st, st.whole, st.hfinal)
        model = self.model
        dec = model.model.decoder
        B, H, hd, d = st.B, st.H, st.hd, st.d
        pos = st.pos
        x = dec.embed_tokens(st.tok) + dec.embed_positions.weight.index_select(0, pos)
        if not st.own_attn:
            st.mask.copy_(st.arange <= pos)
            mask = st.mask.view(1, 1, 1, -1)
        for li, layer in enumerate(dec.layers):
            sa, ca = layer.self_attn, layer.encoder_attn
            res = x
            h = layer.self_attn_layer_norm(x)
            if st.own_attn:
                # scaling, cache append and attention over positions 0..pos in one launch (attn_decode.cu)
                a, _ = F.self_attn_decode(sa.q_proj(h).view(B, d), sa.k_proj(h).view(B, d), sa.v_proj(h).view(B, d),
                                          sa.scaling, st.k[li], st.v[li], pos, H)
                a = a.view(B, 1, d)
            else:
                q = (sa.q_proj(h) * sa.scaling).view(B, 1, H, hd).transpose(1, 2)
                k = sa.k_proj(h).view(B, 1, H, hd).transpose(1, 2)
                v = sa.v_proj(h).view(B, 1, H, hd).transpose(1, 2)
                st.k[li].index_copy_(2, pos, k)
                st.v[li].index_copy_(2, pos, v)
                a = TF.scaled_dot_product_attention(q, st.k[li], st.v[li], attn_mask=mask, scale=1.0)
                a = a.transpose(1, 2).reshape(B, 1, d)
            x = res + sa.out_proj(a)
            res = x
            h = layer.encoder_attn_layer_norm(x)
            if st.own_attn and st.own_cross:
                a, _ = F.cross_attn_decode(ca.q_proj(h).view(B, d), st.ckv[li][:, :, :d], st.ckv[li][:, :, d:],
                                           ca.scaling, H)
                a = a.view(B, 1, d)
            else:
                q = (ca.q_proj(h) * ca.scaling).view(B, 1, H, hd).transpose(1, 2)
                a = TF.scaled_dot_product_attention(q, st.ck[li].transpose(1, 2), st.cv[li].transpose(1, 2),
                                                    scale=1.0)
                a = a.transpose(1, 2).reshape(B, 1, d)
            x = res + ca.out_proj(a)
            res = x
            h = layer.final_layer_norm(x)
            x = res + layer.fc2(layer.activation_fn(layer.fc1(h)))
        x = dec.layer_norm(x)
        self._project(st, st.whole, x.view(B, d))

    def _project(self, st: _State, v: _State, h: torch.Tensor):
        """Vocabulary projection of the final hidden rows [B, d] and the greedy choice under st.maskrow.  An
        unquantized proj_out (the HF bitsandbytes flows keep it fp16) runs on the tcgen05 GEMM with the arg-max folded
        into its epilogue: the [B, V] logits are written only when the state was built to keep them."""
        if st.proj_own:
            po = self.model.proj_out
            F.gemm_f16(h, po.weight, st.proj_bias, out=v.logits_padded if st.store_logits else None,
                       argmax_keys=v.keys if st.argmax_in_graph else None,
                       mask=st.maskrow if st.argmax_in_graph else None, store=st.store_logits)
            if st.argmax_in_graph:
                F.argmax_finalize(v.keys, out=v.next)
            return
        v.logits.copy_(self.model.proj_out(h))
        if st.argmax_in_graph:
            F.masked_argmax(v.logits, st.maskrow[:v.logits.shape[1]], out=v.next)

    def _decoder_step_int8(self, st: _State, v: _State, project: bool = True):
        """The same step when every decoder linear is a bitsandbytes-style Linear8bitLt (fp16): each quantized
        GEMM is fed by a producer that already wrote its int8 rows (rowops.cu / attn_decode.cu), q/k/v share one
        GEMM over the concatenated weights, residual adds ride in the next LayerNorm launch.  13 launches per
        layer instead of ~45; the int8 arithmetic per linear is unchanged (same codes, scales, outlier path)."""
        dec = self.model.model.decoder
        B, H, hd, d, thr = v.B, st.H, st.hd, st.d, st.threshold
        x = dec.embed_tokens(v.tok).view(B, d) + dec.embed_positions.weight.index_select(0, st.pos)
        delta = None

        gemm = fused.gemm      # LLM.int8: consumes the producer's int8 rows; weight-only schemes: the fp16 / bf16 rows

        # row groups of <= 16 rows: Linear8bitLt's single-launch small-row kernel (quantizes the few rows itself, dp4a)
        # beats the tensor-core tile; the producer feeding such a GEMM then skips its own quantization (thr -> None)
        def t_for(w):
            return None if (thr is not None and self.small_int8 and fused.small_rows_ok(B, w)) else thr

        for li, (layer, fw) in enumerate(zip(dec.layers, st.fused)):
            ln = layer.self_attn_layer_norm
            x, h, qt = F.add_layernorm_quant(x, delta, ln.weight, ln.bias, ln.eps, t_for(fw.qkv))
            qkv = gemm(qt, h, fw.qkv)
            a, qt = F.self_attn_decode(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], fw.scaling, v.k[li], v.v[li],
                                       st.pos, H, t_for(fw.o))
            delta = gemm(qt, a, fw.o)
            ln = layer.encoder_attn_layer_norm
            x, h, qt = F.add_layernorm_quant(x, delta, ln.weight, ln.bias, ln.eps, t_for(fw.cq))
            q = gemm(qt, h, fw.cq)
            tco = t_for(fw.co)
            if st.own_cross:
                # q scaling, the pass over the 1500 cached encoder positions and out_proj's quantization: one launch
                if self.cross_quant_inline or tco is None:
                    a, qt = F.cross_attn_decode(q, v.ckv[li][:, :, :d], v.ckv[li][:, :, d:], fw.scaling, H, tco)
                else:       # the row quantization as a launch of its own (no cross-CTA completion tail in the stream)
                    a, _ = F.cross_attn_decode(q, v.ckv[li][:, :, :d], v.ckv[li][:, :, d:], fw.scaling, H, None)
                    qt = F.int8_vectorwise_quant(a, tco, finalize=False)
            else:
                if fw.scaling_pow2:     # q * 2^-k is exact in fp16, so the scale can ride in the SDPA call
                    a = TF.scaled_dot_product_attention(q.view(B, 1, H, hd).transpose(1, 2),
                                                        v.ck[li].transpose(1, 2), v.cv[li].transpose(1, 2),
                                                        scale=fw.scaling)
                else:
                    a = TF.scaled_dot_product_attention((q * fw.scaling).view(B, 1, H, hd).transpose(1, 2),
                                                        v.ck[li].transpose(1, 2), v.cv[li].transpose(1, 2), scale=1.0)
                a = a.transpose(1, 2).reshape(B, d)
                qt = F.int8_vectorwise_quant(a, tco, finalize=False) if tco is not None else None
            delta = gemm(qt, a, fw.co)
            ln = layer.final_layer_norm
            x, h, qt = F.add_layernorm_quant(x, delta, ln.weight, ln.bias, ln.eps, t_for(fw.fc1))
            g, qt = F.gelu_quant(gemm(qt, h, fw.fc1), t_for(fw.fc2))
            delta = gemm(qt, g, fw.fc2)
        ln = dec.layer_norm
        if not project:      # multi-stream step: the groups' final hidden rows meet in one buffer
            F.add_layernorm_quant(x, delta, ln.weight, ln.bias, ln.eps, None, h_out=st.hfinal[v.r0:v.r1])
            return
        _, h, _ = F.add_layernorm_quant(x, delta, ln.weight, ln.bias, ln.eps, None)
        self._project(st, v, h)

    # ------------------------------------------------------------------------------------------
    # persistent decoder-layer kernel (decode_fused.cu): one launch between two cross-attention passes
    # ------------------------------------------------------------------------------------------
    def _plan_mega(self, st: _State, dtype) -> bool:
        """Scratch and weight-pointer tables for the persistent per-layer kernel, per row group; False when it does not
        apply (LLM.int8 / fp16 only, <= 64 rows per group, d_model <= 1280)."""
        if (not self.mega or st.fused is None or st.threshold is None or dtype != torch.float16 or not st.own_attn
                or not st.own_cross or st.d > 1280 or any(v.B > 64 for v in st.views)):
            return False
        cfg = self.model.config
        ffn = cfg.decoder_ffn_dim
        if ffn % 16 != 0 or st.fused[0].qkv.kind != "int8":
            return False
        dec = self.model.model.decoder
        dev = st.tok.device
        fl = max(st.d, ffn) + 2

        def lin(w):
            return _lib.DecodeLinear(w.cb.data_ptr(), w.scb.data_ptr(), 0 if w.bias is None else w.bias.data_ptr(),
                                     int(w.out_features), int(w.in_features))

        for v in st.views:
            M, d = v.B, st.d
            z = lambda *shape, dt=torch.float16: torch.zeros(shape, dtype=dt, device=dev)
            v.m = _State()
            v.m.x, v.m.h, v.m.att, v.m.q = z(M, d), z(M, d), z(M, d), z(M, d)
            v.m.qkv, v.m.f1, v.m.g = z(M, 3 * d), z(M, ffn), z(M, ffn)
            v.m.ca_d, v.m.ca_f = z(3, M, d, dt=torch.int8), z(M, ffn, dt=torch.int8)
            v.m.sca, v.m.flags = z(4, M, dt=torch.float32), z(4, fl, dt=torch.int32)
            v.m.bar = z(2, dt=torch.int32)
            v.m.hfin = st.hfinal[v.r0:v.r1] if st.hfinal is not None else z(M, d)
            v.m.layers = []
            for li, (layer, fw) in enumerate(zip(dec.layers, st.fused)):
                L = _lib.DecodeLayer()
                L.qkv, L.o, L.cq, L.co, L.fc1, L.fc2 = lin(fw.qkv), lin(fw.o), lin(fw.cq), lin(fw.co), lin(fw.fc1), lin(fw.fc2)
                n1, n2, n3 = layer.self_attn_layer_norm, layer.encoder_attn_layer_norm, layer.final_layer_norm
                L.ln1_g, L.ln1_b, L.eps1 = n1.weight.data_ptr(), n1.bias.data_ptr(), float(n1.eps)
                L.ln2_g, L.ln2_b, L.eps2 = n2.weight.data_ptr(), n2.bias.data_ptr(), float(n2.eps)
                L.ln3_g, L.ln3_b, L.eps3 = n3.weight.data_ptr(), n3.bias.data_ptr(), float(n3.eps)
                L.kcache, L.vcache = v.k[li].data_ptr(), v.v[li].data_ptr()
                v.m.layers.append(L)
        st.mega_ctas = max(8, min(128, 280 // len(st.views)))
        st.mega_ffn = ffn
        return True

    def _mega_args(self, st: _State, v: _State, xa=None):
        m, dec = v.m, self.model.model.decoder
        a = _lib.DecodeArgs()
        a.M, a.d, a.ffn, a.H, a.t_max = v.B, st.d, st.mega_ffn, st.H, st.k[0].shape[1]
        a.threshold, a.scaling = float(st.threshold), float(st.fused[0].scaling)
        a.pos, a.x = st.pos.data_ptr(), m.x.data_ptr()
        a.h, a.att, a.qkv, a.f1, a.g = (t.data_ptr() for t in (m.h, m.att, m.qkv, m.f1, m.g))
        a.ca_d, a.ca_f, a.sca, a.flags = m.ca_d.data_ptr(), m.ca_f.data_ptr(), m.sca.data_ptr(), m.flags.data_ptr()
        a.q_out, a.bar = m.q.data_ptr(), m.bar.data_ptr()
        if xa is not None:
            out, (ca, sca, state) = xa
            a.xa, a.xa_ca, a.xa_sca = out.data_ptr(), ca.data_ptr(), sca.data_ptr()
            a.xa_flags = 0 if state is None else state.col_flags.data_ptr()
        ln = dec.layer_norm
        a.lnf_g, a.lnf_b, a.epsf, a.hfinal = ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps), m.hfin.data_ptr()
        return a

    def _decoder_step_mega(self, st: _State, v: _State, project: bool = True):
        """The fused LLM.int8 step with ONE persistent launch between consecutive cross-attention passes (the part of
        layer l after its cross-attention and the part of layer l + 1 before its own, decode_fused.cu) instead of
        twelve: L + 1 fused launches and L attention launches per token.  Same arithmetic as _decoder_step_int8."""
        dec = self.model.model.decoder
        d, H, thr = st.d, st.H, st.threshold
        torch.add(dec.embed_tokens(v.tok).view(v.B, d), dec.embed_positions.weight.index_select(0, st.pos), out=v.m.x)
        layers = v.m.layers
        xa = None
        for li in range(len(layers)):
            F.decode_fused_llmint8(layers[li - 1] if li else None, layers[li], self._mega_args(st, v, xa), li > 0, True,
                                   False, st.mega_ctas)
            xa = F.cross_attn_decode(v.m.q, v.ckv[li][:, :, :d], v.ckv[li][:, :, d:], st.fused[li].scaling, H, thr)
        F.decode_fused_llmint8(layers[-1], None, self._mega_args(st, v, xa), True, False, True, st.mega_ctas)
        if project:
            self._project(st, v, v.m.hfin)

    def _plan_int8(self, dtype):
        """Per-layer packed weights for the fused decode step, or None when the decoder's linears are not all drop-in
        modules of ONE scheme that fused.pack serves for `dtype` (LLM.int8 with fp16 activations; W8A16 / NF4 / FP4 /
        qint4 / qint2 with fp16 or bf16 activations).  Returns (plans, threshold): threshold is None for the
        weight-only schemes (their producers write no int8 rows)."""
        if dtype not in (torch.float16, torch.bfloat16):
            return None, 0.0
        cfg = self.model.config
        if cfg.activation_function != "gelu" or cfg.d_model // cfg.decoder_attention_heads != 64 or cfg.d_model > 2048:
            return None, 0.0
        plans, thr, kind = [], None, None

        def pack(mods):
            return fused.pack(mods, dtype)

        for layer in self.model.model.decoder.layers:
            sa, ca = layer.self_attn, layer.encoder_attn
            fw = _State()
            fw.qkv, fw.o = pack([sa.q_proj, sa.k_proj, sa.v_proj]), pack([sa.out_proj])
            fw.cq, fw.co = pack([ca.q_proj]), pack([ca.out_proj])
            fw.ckv = pack([ca.k_proj, ca.v_proj])
            fw.fc1, fw.fc2 = pack([layer.fc1]), pack([layer.fc2])
            ws = [fw.qkv, fw.o, fw.cq, fw.co, fw.ckv, fw.fc1, fw.fc2]
            if any(w is None for w in ws) or sa.scaling != ca.scaling:
                return None, 0.0
            kind = ws[0].kind if kind is None else kind
            thr = ws[0].threshold if thr is None else thr
            if (any(w.kind != kind or w.threshold != thr for w in ws) or layer.fc1.out_features % 8 != 0
                    or layer.self_attn_layer_norm.weight.dtype != dtype):
                return None, 0.0
            fw.scaling = float(sa.scaling)
            fw.scaling_pow2 = fused.is_pow2(fw.scaling)
            plans.append(fw)
        return plans, (thr if kind == "int8" else None)

    def _fingerprint(self):
        """Identity of every tensor the captured graphs and packed copies were built from: (data_ptr, version) of
        the model's parameters and of the quantized state the drop-in modules keep outside their
        parameters.  In-place updates bump the version, re-quantizing / .to() / load_state_dict(assign) / pruning
        re-parametrisation change the pointer: either way the cached states are rebuilt."""
        fp = []
        for root in (self.model,):          # encoder too: fastenc's per-layer plans hold concatenated copies
            for p in root.parameters():
                fp.append((p.data_ptr(), p._version))
            for b in root.buffers():            # calibrated activation scales (quanto input_scale / output_scale)
                if b.numel() == 1:
                    fp.append((b.data_ptr(), b._version))
            for m in root.modules():
                stt = getattr(m, "state", None)
                for t in (getattr(stt, "CB", None), getattr(stt, "SCB", None), getattr(m, "_wq", None),
                          getattr(m, "_wscale", None), getattr(m, "_wshift", None)):
                    if isinstance(t, torch.Tensor):
                        fp.append((t.data_ptr(), t._version))
                qs = getattr(getattr(m, "weight", None), "quant_state", None)
                if qs is not None and isinstance(getattr(qs, "absmax", None), torch.Tensor):
                    fp.append((qs.absmax.data_ptr(), qs.absmax._version))
        return hash(tuple(fp))

    def invalidate(self) -> None:
        """Drop every captured graph / packed weight copy (call after changing weights in a way the fingerprint
        cannot see, e.g. writing through a raw pointer)."""
        self._states.clear()
        for layer in self.model.model.encoder.layers:
            if hasattr(layer, "_whisperq_plan"):
                del layer._whisperq_plan

    def step_logits(self, st: _State) -> torch.Tensor:
        """Raw logits [B, V] of the last replayed step (tests; states built with keep_logits)."""
        if not st.store_logits:
            raise RuntimeError("this state was captured without the logits store (set eng.keep_logits = True first)")
        return st.logits

    def _get_state(self, B: int, t_max: int, dtype: torch.dtype, device, store_logits: bool = True) -> _State:
        fp = self._fingerprint()
        key = (B, t_max, dtype, torch.device(device).index, bool(store_logits), self.cross_attention, self.streams,
               self.cross_quant_inline, self.mega, self.min_rows_per_stream, self.small_int8, self.group_project)
        st = self._states.get(key)
        if st is not None and st.fingerprint == fp:
            return st
        if st is not None:              # weights changed since capture: everything cached is stale
            self.invalidate()
        if len(self._states) >= self.max_states:       # static caches are large: keep a few shapes
            self._states.pop(next(iter(self._states)))
        cfg = self.model.config
        st = _State()
        st.fingerprint = fp
        st.B, st.d = B, cfg.d_model
        st.H = cfg.decoder_attention_heads
        st.hd = st.d // st.H
        S = cfg.max_source_positions
        L = cfg.decoder_layers
        st.tok = torch.zeros((B, 1), dtype=torch.long, device=device)
        st.pos = torch.zeros((1,), dtype=torch.long, device=device)
        st.arange = torch.arange(t_max, device=device)
        st.mask = torch.zeros((t_max,), dtype=torch.bool, device=device)
        st.fused, st.threshold = self._plan_int8(dtype) if self.fuse_int8 else (None, 0.0)
        # (fp32: the reference's quanto / bnb *_32 flows; torch's fp32 SDPA was 77 % of a whisper-medium step)
        st.own_attn = self.own_attention and st.hd == 64 and dtype in (torch.float16, torch.bfloat16, torch.float32)
        # decode-time cross-attention: the persistent item-walking kernel of attn_decode.cu for every batch size
        # ("cudnn" keeps torch SDPA for A/B measurements only)
        st.own_cross = self.cross_attention != "cudnn"
        if st.fused is not None or st.own_attn:
            kv_shape = (B, t_max, st.d)              # projection layout: one 128-byte row per head and position
        else:
            kv_shape = (B, st.H, t_max, st.hd)
        st.k = [torch.zeros(kv_shape, dtype=dtype, device=device) for _ in range(L)]
        st.v = [torch.zeros(kv_shape, dtype=dtype, device=device) for _ in range(L)]
        # cross-attention K/V stay in the projection's own [B, S, H, hd] layout: the q_len = 1 SDPA kernel
        # streams them through their strides at the same HBM rate (scripts/attn_layout_bench.py)
        # (a layer's K and V are the two column blocks of one [B, S, 2d] buffer, so that an all-int8 decoder can fill
        # both with one GEMM over [Wk; Wv])
        st.ckv = [torch.zeros((B, S, 2 * st.d), dtype=dtype, device=device) for _ in range(L)]
        st.ck = [t[:, :, :st.d].view(B, S, st.H, st.hd) for t in st.ckv]
        st.cv = [t[:, :, st.d:].view(B, S, st.H, st.hd) for t in st.ckv]
        po = self.model.proj_out
        V = po.out_features
        Vp = -(-V // 8) * 8          # rows of the logits buffer start 16-byte aligned (vector loads, cuBLAS)
        st.logits_padded = torch.zeros((B, Vp), dtype=dtype, device=device)
        st.logits = st.logits_padded[:, :V]
        # greedy choice inside the graph: argmax of the logits under this step's suppression mask
        st.argmax_in_graph = dtype in (torch.float16, torch.bfloat16)
        # an unquantized projection (HF bitsandbytes flows: proj_out stays fp16) runs on the library's own GEMM,
        # straight from the module's weight (any vocabulary size; no padded copy), arg-max in the epilogue
        st.proj_own = (type(po) is torch.nn.Linear and po.weight.dtype == dtype and po.weight.is_contiguous()
                       and dtype in (torch.float16, torch.bfloat16) and po.in_features % 8 == 0)
        st.proj_bias = po.bias.detach().float() if (st.proj_own and po.bias is not None) else None
        st.store_logits = bool(store_logits) or not (st.proj_own and st.argmax_in_graph)
        st.keys = torch.zeros((B,), dtype=torch.long, device=device)
        st.maskrow = torch.zeros((-(-V // 256) * 256,), dtype=torch.bool, device=device)   # whole GEMM tiles
        st.next = torch.zeros((B,), dtype=torch.long, device=device)
        st.mask_cache = {}
        # row groups of the batch for the multi-stream fused step (one group = the whole batch otherwise)
        n = self.streams if (st.fused is not None and st.own_attn and st.own_cross) else 1
        while n > 1 and (B % n != 0 or B // n < self.min_rows_per_stream):
            n //= 2
        st.views = []
        st.hfinal = torch.zeros((B, st.d), dtype=dtype, device=device) if n > 1 else None
        for i in range(-1, n):
            r0, r1 = (0, B) if i < 0 else (i * (B // n), (i + 1) * (B // n))
            v = _State()
            v.B, v.slot, v.r0, v.r1 = r1 - r0, max(i, 0), r0, r1
            v.stream = None if i <= 0 else torch.cuda.Stream(device=device)
            v.tok, v.next, v.keys = st.tok[r0:r1], st.next[r0:r1], st.keys[r0:r1]
            v.k, v.v = [t[r0:r1] for t in st.k], [t[r0:r1] for t in st.v]
            v.ckv = [t[r0:r1] for t in st.ckv]
            v.ck, v.cv = [t[r0:r1] for t in st.ck], [t[r0:r1] for t in st.cv]
            v.logits_padded, v.logits = st.logits_padded[r0:r1], st.logits[r0:r1]
            if i < 0:
                st.whole = v
            else:
                st.views.append(v)
        st.mega = self._plan_mega(st, dtype)
        # warm up on a side stream (lazy inits, autotuning), then capture
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                self._decoder_step(st)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        before = F.STATS.launches
        st.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(st.graph):
            self._decoder_step(st)
        st.launches_per_replay = F.STATS.launches - before
        self._states[key] = st
        return st

    @staticmethod
    def _processor_signature(processors):
        """Hashable description of mask-only processors (their masks depend on it and on the length only);
        a fresh object when a processor's attributes are not the expected ones (then nothing is shared
        between generate calls)."""
        try:
            sig = []
            for p in processors:
                item = [type(p).__name__]
                for name in ("suppress_tokens", "begin_suppress_tokens", "begin_index", "prompt_length_to_skip",
                             "min_new_tokens", "min_length", "eos_token_id"):
                    if hasattr(p, name):
                        v = getattr(p, name)
                        item.append((name, tuple(v.reshape(-1).tolist()) if isinstance(v, torch.Tensor) else
                                     tuple(v) if isinstance(v, (list, tuple)) else v))
                sig.append(tuple(item))
            return tuple(sig)
        except Exception:
            return object()

    @staticmethod
    def _plain_decoder_inputs(input_ids, model_kwargs) -> bool:
        """The fast loop feeds token t at position t and attends every earlier position.  HF derives positions from
        `decoder_attention_mask` when prompts are left-padded (generation_whisper.py: prompt_ids /
        condition_on_prev_tokens) and accepts explicit position ids / embeddings: any of those keeps HF's loop."""
        if model_kwargs.get("decoder_position_ids") is not None or model_kwargs.get("decoder_inputs_embeds") is not None:
            return False
        m = model_kwargs.get("decoder_attention_mask")
        if m is not None and not bool((m != 0).all()):
            return False
        return True

    @staticmethod
    def _known_criteria(stopping_criteria) -> bool:
        from transformers.generation.stopping_criteria import EosTokenCriteria, MaxLengthCriteria
        return all(type(c) in (EosTokenCriteria, MaxLengthCriteria) for c in (stopping_criteria or []))

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _sample(self, input_ids, logits_processor=None, stopping_criteria=None, generation_config=None,
                synced_gpus=False, streamer=None, **model_kwargs):
        model = self.model
        enc_out = model_kwargs.get("encoder_outputs")
        from . import quanto as _quanto
        eligible = (input_ids.is_cuda and not generation_config.do_sample and not _quanto.calibrating()
                    and not generation_config.return_dict_in_generate and streamer is None and not synced_gpus
                    and enc_out is not None and model_kwargs.get("use_cache", True)
                    and generation_config.max_length is not None
                    and generation_config.max_length <= model.config.max_target_positions
                    and self._plain_decoder_inputs(input_ids, model_kwargs)
                    and self._known_criteria(stopping_criteria))
        if not eligible:
            self.fallbacks += 1
            out = self._orig_sample(input_ids, logits_processor=logits_processor,
                                    stopping_criteria=stopping_criteria, generation_config=generation_config,
                                    synced_gpus=synced_gpus, streamer=streamer, **model_kwargs)
            if isinstance(out, torch.Tensor):
                self._maybe_unwind(out, input_ids.shape[1], generation_config)
            return out
        enc = enc_out[0] if not hasattr(enc_out, "last_hidden_state") else enc_out.last_hidden_state
        B, P = input_ids.shape
        t_max = -(-int(generation_config.max_length) // self.len_bucket) * self.len_bucket
        t_max = min(t_max, model.config.max_target_positions)
        # Whisper's processors only write -inf at positions that depend on the current LENGTH
        # (suppress lists, begin-suppress, min-new-tokens): evaluate them on a single zero row
        # (cost independent of the batch) and apply the resulting mask to the fp16 logits.  argmax
        # of the masked fp16 logits == argmax of HF's masked fp32 copy (exact widening, same ties).
        from transformers.generation.logits_process import (MinLengthLogitsProcessor,
                                                            MinNewTokensLengthLogitsProcessor,
                                                            SuppressTokensAtBeginLogitsProcessor,
                                                            SuppressTokensLogitsProcessor)
        mask_only = (SuppressTokensLogitsProcessor, SuppressTokensAtBeginLogitsProcessor,
                     MinNewTokensLengthLogitsProcessor, MinLengthLogitsProcessor)
        maskable = all(type(p) in mask_only for p in logits_processor)
        # the raw logits are only materialised when something reads them (non-mask processors, tests)
        st = self._get_state(B, t_max, enc.dtype, input_ids.device, store_logits=self.keep_logits or not maskable)

        # cross-attention keys / values once per call (encoder-shaped GEMMs)
        S = enc.shape[1]
        layers = model.model.decoder.layers
        if st.fused is not None and enc.is_contiguous():
            # all-Linear8bitLt decoder: the encoder output is quantized ONCE (bitsandbytes would do it in each of
            # the 2L projections) and every layer's [Wk; Wv] GEMM writes straight into its K|V buffer
            enc2 = enc.view(B * S, st.d)
            qt = F.int8_vectorwise_quant(enc2, st.threshold, finalize=False) if st.threshold is not None else None
            for li, fw in enumerate(st.fused):
                fused.gemm(qt, enc2, fw.ckv, out=st.ckv[li], keep_flags=li + 1 < len(layers))
        else:
            for li, layer in enumerate(layers):
                ca = layer.encoder_attn
                st.ck[li].copy_(ca.k_proj(enc).view(B, S, st.H, st.hd))
                st.cv[li].copy_(ca.v_proj(enc).view(B, S, st.H, st.hd))

        V = st.logits.shape[1]
        pad_token_id = generation_config._pad_token_tensor
        has_eos = any(hasattr(c, "eos_token_id") for c in stopping_criteria)
        unfinished = torch.ones(B, dtype=torch.long, device=input_ids.device)
        max_length = int(generation_config.max_length)

        zero_row = torch.zeros((1, st.logits.shape[1]), dtype=torch.float32, device=input_ids.device)

        def eos_suppressed(length: int) -> bool:
            """True when a min-length processor forces EOS to -inf for a prefix of this length."""
            for p in logits_processor:
                if type(p) is MinNewTokensLengthLogitsProcessor:
                    if length - p.prompt_length_to_skip < p.min_new_tokens:
                        return True
                elif type(p) is MinLengthLogitsProcessor:
                    if length < p.min_length:
                        return True
            return False

        # With mask-only processors the greedy choice happens inside the graph (wq_masked_argmax on the fp16
        # logits under st.maskrow); the mask of a given sequence length is evaluated once and cached.
        in_graph = maskable and st.argmax_in_graph
        sig = self._processor_signature(logits_processor) if in_graph else None

        def mask_for(ids):
            key = (sig, ids.shape[1])
            m = st.mask_cache.get(key)
            if m is None:
                m = torch.isinf(logits_processor(ids[:1], zero_row.clone())).view(-1)
                if len(st.mask_cache) >= 2048:
                    st.mask_cache.clear()
                st.mask_cache[key] = m
            return m

        def run(tokens, position, ids_after=None):
            """Feed `tokens` at `position`; ids_after = the sequence whose next token these logits choose."""
            st.tok.copy_(tokens.view(B, 1))
            st.pos.fill_(position)
            if in_graph and ids_after is not None:
                st.maskrow[:V].copy_(mask_for(ids_after))
            st.graph.replay()
            self.replays += 1
            F.STATS.launches += st.launches_per_replay

        replays0 = self.replays
        if self.time_loop:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for i in range(P):                       # prompt tokens (normally just <|startoftranscript|>)
            run(input_ids[:, i], i, input_ids if i == P - 1 else None)
        cur = P
        while True:
            length = input_ids.shape[1]
            if in_graph:
                next_tokens = st.next.clone()
            elif maskable:
                row = logits_processor(input_ids[:1], zero_row.clone())
                next_tokens = torch.argmax(st.logits.masked_fill(torch.isinf(row), float("-inf")), dim=-1)
            else:
                next_token_logits = st.logits.to(copy=True, dtype=torch.float32)
                next_tokens = torch.argmax(logits_processor(input_ids, next_token_logits), dim=-1)
            if has_eos:
                next_tokens = next_tokens * unfinished + pad_token_id * (1 - unfinished)
            input_ids = torch.cat([input_ids, next_tokens[:, None]], dim=-1)
            if maskable and has_eos and eos_suppressed(length):
                # no sequence can finish at this step except by reaching max_length, which the host
                # knows: no device->host synchronisation, the next replay is queued immediately
                if length + 1 >= max_length:
                    break
            else:
                unfinished = unfinished & ~stopping_criteria(input_ids, None)
                if bool(unfinished.max() == 0):
                    break
            run(next_tokens, cur, input_ids)
            cur += 1
        # WhisperGenerationMixin post-processes every utterance separately (generation_whisper.py,
        # generate_with_fallback / _retrieve_segment: `seq[-1] == pad`, nonzero(), slicing): on device tensors
        # that is ~4 blocking syncs per utterance, ~45 ms at B = 256 -- as long as 25 decode steps.  The ids
        # are read back once here (HF's own dict path does the same with .cpu()) and HF's loops run on
        # the host copy; `generate` above returns the final tensor to the model's device.
        if self.time_loop:
            ev1.record()
            self.loop_events.append((ev0, ev1, self.replays - replays0))
        if not self.host_postprocess:
            return input_ids
        # everything of this batch is queued and the host is about to block on the read-back: a caller's host work on
        # the PREVIOUS batch (ids -> text, metric tallies: ~4 ms per 256 utterances) fits here, under this batch's GPU
        # time -- an evaluation loop pipelined one batch deep (bench.py's end-to-end arm sets it)
        if self.before_readback is not None:
            self.before_readback()
        ids = input_ids.cpu()
        self._maybe_unwind(ids, P, generation_config)
        return ids


def enable(model, len_bucket: int = 64) -> GraphedGreedy:
    """Install the graph-replay greedy loop on `model` (idempotent)."""
    existing = getattr(model, "_whisperq_fastgen", None)
    if existing is not None:
        return existing
    return GraphedGreedy(model, len_bucket).install()

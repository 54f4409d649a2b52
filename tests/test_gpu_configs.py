"""Token / logit parity at the BASELINE.json config sizes, on the exact branches bench.py times.

VERDICT round 1, "what's weak" 1-2: the headline step (whisper-base, all 6 + 6 layers, heads x utterances > 1024,
producer-fused LLM.int8 decode step behind ``model.generate``) and the config-sized models C3-C5 had no token or
logit comparison.  Every case here builds the model with the geometry of the named config (full depth for base and
small, >= 4 + 4 / 2 + 2 layers of the true width for medium / large-v3 so that the CPU-side random init stays within
seconds), runs the reference flow of that config (``harness.apply_scheme``: the same swap the reference's loader
performs) and compares

  (1) the CUDA-graph decode loop (fastgen, what bench.py times) with HF's own ``_sample`` loop over the SAME drop-in
      modules: identical greedy ids up to the first position whose top-1 / top-2 margin in HF's own logits is below the
      stated tolerance, and teacher-forced logits of every position within that tolerance;
  (2) the drop-in modules with a torch-op emulation of the reference linears (tests/emulation.py, which follows
      the oracle operation by operation) inside otherwise identical HF code: ids and logits as stated per case.

Same inputs as bench.py: seeded gaussian audio through the log-mel kernel (so the features have the statistics the
bench sees), random-init weights, greedy decoding with min = max new tokens (data_utils.py:141-155 call pattern).
"""
import numpy as np
import pytest
import torch
from torch import nn

from tests import emulation as emu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from openai_whisper_compression_b200 import harness
    return harness


def _bench_features(H, n, mels, half):
    """Log-mel features of bench.py's utterances 0..n-1 (harness.synth_audio through the CUDA frontend)."""
    proc = H.StubProcessor(mels, device="cuda")
    audio = torch.stack([torch.from_numpy(H.synth_audio(i)) for i in range(n)]).cuda()
    f = proc.feature_extractor.features_from_device_audio(audio)
    return f.half() if half else f


def _hf_loop(model, feats, T):
    """HF's own greedy loop on `model` with per-step raw logits (return_dict keeps fastgen out of the way)."""
    out = model.generate(feats, do_sample=False, num_beams=1, min_new_tokens=T, max_new_tokens=T,
                         return_dict_in_generate=True, output_logits=True)
    return out.sequences, torch.stack(out.logits, 1).float()


def _prefix_until_indecisive(logits, tol):
    top2 = logits.topk(2, -1).values
    decisive = (top2[..., 0] - top2[..., 1]) > tol
    return torch.where(decisive.all(1), decisive.shape[1], (~decisive).float().argmax(1)), decisive


def _check_fast_vs_hf(H, model, feats, T, rel_tol, rel_mean, expect_fused=None, cross=None, streams=None):
    """fastgen (graph replay, fused step where available) vs HF's loop on the same modules.  Tolerances are relative
    to the largest |logit| of HF's run (random-init logits are O(2..4), growing with the width of the model)."""
    from openai_whisper_compression_b200 import fastgen
    hf_seq, hf_logits = _hf_loop(model, feats, T)
    scale = max(1.0, hf_logits.abs().max().item())
    tol, mean_tol = rel_tol * scale, rel_mean * scale
    ref_ids = H.greedy_generate(model, feats, T)
    eng = fastgen.enable(model)
    if cross is not None:
        eng.cross_attention = cross
    if streams is not None:
        eng.streams = streams
    ids = H.greedy_generate(model, feats, T)
    assert eng.replays > 0 and eng.fallbacks == 0
    # the same call with the logits store captured into the step (the bench branch keeps only the argmax)
    eng.keep_logits = True
    assert torch.equal(H.greedy_generate(model, feats, T), ids)
    st = [s for s in eng._states.values() if s.store_logits][-1]
    if expect_fused is not None:
        assert (st.fused is not None) == expect_fused
    assert ids.shape == ref_ids.shape
    first_bad, decisive = _prefix_until_indecisive(hf_logits, 2 * tol)
    P = ids.shape[1] - T
    same = 0
    for b in range(ids.shape[0]):
        n = P + int(first_bad[b])
        assert torch.equal(ids[b, :n], ref_ids[b, :n]), (b, n)
        same += int((ids[b] == ref_ids[b]).sum())
    # teacher-forced: replay the captured step on HF's own history, compare raw logits at every position
    start = torch.full((ids.shape[0], 1), model.config.decoder_start_token_id, dtype=ref_ids.dtype, device="cuda")
    full = ref_ids if P == 1 else torch.cat([start, ref_ids], 1)
    worst = mean = 0.0
    flips = 0
    for j in range(T):
        st.tok.copy_(full[:, j:j + 1])
        st.pos.fill_(j)
        st.graph.replay()
        got = eng.step_logits(st).float()
        diff = (got - hf_logits[:, j]).abs()
        worst, mean = max(worst, diff.max().item()), mean + diff.mean().item() / T
        dj = decisive[:, j]
        flips += int((got.argmax(-1)[dj] != hf_logits[:, j].argmax(-1)[dj]).sum())
    frac = same / ids.numel()
    print(f"fast vs HF loop: identical ids {frac:.4f}, teacher-forced logits max |diff| {worst:.3e} mean {mean:.3e} "
          f"(scale {hf_logits.abs().max().item():.2f}), decisive positions {int(decisive.sum())}/{decisive.numel()}")
    assert worst <= tol and mean <= mean_tol, (worst, mean)
    assert flips == 0
    eng.uninstall()
    return ids, ref_ids, hf_logits


# ------------------------------------------------------------------------------------------------------------
# C2 -- the bench branch: whisper-base, all 6 + 6 layers, LLM.int8 (threshold 6.0, HF load_in_8bit flow)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,cross,streams", [(144, "own", 4), (144, "own", 1), (128, "own", 2), (24, "own", 4),
                                             (136, "cudnn", 1)])
def test_config2_base_llmint8_bench_branch_matches_hf_loop(H, B, cross, streams):
    """bench.py's default workload at heads x utterances > 1024 (144 x 8 = 1152): the producer-fused int8 decode
    step -- row groups of the batch on 4 streams (bench default), 2 streams, or one --, TMA-fed cross-attention, fused
    q|k|v GEMM, lean decode tiles, projection with the arg-max in its epilogue.  Tolerance: 3 % of the logit scale
    (~2.4 -> 7e-2 abs: one-ulp LayerNorm differences pass through the int8 quantizers of 12 layers), mean 0.5 %."""
    model = H.apply_scheme(H.build_model("base"), "llm_int8", "cuda")
    feats = _bench_features(H, B, 80, True)
    _check_fast_vs_hf(H, model, feats, 16, 3e-2, 5e-3, expect_fused=True, cross=cross, streams=streams)


def test_config2_base_llmint8_modules_match_emulation(H):
    """The same model through HF's loop: drop-in Linear8bitLt modules vs the torch-op emulation of bitsandbytes'
    arithmetic in every slot (36 encoder + 60 decoder linears): identical ids at decisive positions, logits within
    2e-3 (bit-exact wherever no activation reaches the outlier threshold)."""
    ours = H.apply_scheme(H.build_model("base"), "llm_int8", "cuda")
    ref = H.build_model("base").half()
    emu.swap_all(ref, lambda m: emu.EmuLinear8bitLt(m.cuda(), 6.0))
    ref = ref.cuda()
    feats = _bench_features(H, 8, 80, True)
    T = 16
    ids_ref, log_ref = _hf_loop(ref, feats, T)
    with torch.no_grad():
        log_ours = ours(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
        log_emu = ref(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
    d = (log_ours - log_emu).abs().max().item()
    print(f"base LLM.int8 modules vs emulation: teacher-forced logits max |diff| {d:.3e}")
    assert d <= 2e-3
    _, decisive = _prefix_until_indecisive(log_emu, 4e-3)
    assert torch.equal(log_ours.argmax(-1)[decisive], log_emu.argmax(-1)[decisive])
    ids_ours = H.greedy_generate(ours, feats, T)
    first_bad, _ = _prefix_until_indecisive(log_ref, 4e-3)
    P = ids_ours.shape[1] - T
    ids_ref_plain = ids_ref if ids_ref.shape[1] == ids_ours.shape[1] else ids_ref[:, ids_ref.shape[1] - ids_ours.shape[1]:]
    for b in range(ids_ours.shape[0]):
        n = P + int(first_bad[b])
        assert torch.equal(ids_ours[b, :n], ids_ref_plain[b, :n])


# ------------------------------------------------------------------------------------------------------------
# C3 -- whisper-small, all 12 + 12 layers, bnb NF4 (fp16 compute, HF load_in_4bit flow)
# ------------------------------------------------------------------------------------------------------------
def _emulate_dequant(H, ours, ref):
    from openai_whisper_compression_b200 import bnb, quanto
    for name, m in list(ours.named_modules()):
        if isinstance(m, bnb.Linear4bit):
            emu._set(ref, name, emu.EmuDequantLinear(bnb.dequantize_4bit(m.weight.data, m.weight.quant_state), m.bias))
        elif isinstance(m, quanto.QLinear):
            q, s = m.qweight
            emu._set(ref, name, emu.EmuDequantLinear(q, m.bias, post_scale=s.t()))
    return ref


def test_config3_small_nf4_full_depth(H):
    """C3: 12 + 12 layers, d = 768, 12 heads, K = 3072 fc2.  (1) fast loop vs HF loop on the same modules;
    (2) modules vs dequantize-then-cuBLAS (what bitsandbytes' batched Linear4bit.forward does): fp16 logits within
    3e-2 abs / decisive tokens identical (24 layers of fp16 rounding-order differences between the fused
    fp32-accumulate kernel and cuBLAS)."""
    ours = H.apply_scheme(H.build_model("small"), "bnb_nf4", "cuda")
    feats = _bench_features(H, 6, 80, True)
    T = 16
    _check_fast_vs_hf(H, ours, feats, T, 1e-2, 1e-3)
    ref = _emulate_dequant(H, ours, H.build_model("small").half().cuda())
    ids_ref, log_ref = _hf_loop(ref, feats, T)
    with torch.no_grad():
        la = ours(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
        lb = ref(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
    d = (la - lb).abs()
    print(f"small NF4 modules vs dequant+cuBLAS: logits max |diff| {d.max().item():.3e} mean {d.mean().item():.3e}")
    assert d.max().item() < 3e-2
    _, decisive = _prefix_until_indecisive(lb, 6e-2)
    assert torch.equal(la.argmax(-1)[decisive], lb.argmax(-1)[decisive])


# ------------------------------------------------------------------------------------------------------------
# C4 -- whisper-medium geometry, 50 % global-L1 pruned, quanto qint8 in the reference's fp32 flow
# ------------------------------------------------------------------------------------------------------------
def _fp32_reference_logits(ours, ref, feats, T):
    """Teacher-forced logits of the drop-in model and of the fp32 emulation (TF32 off: true fp32 matmuls, the
    arithmetic of the reference flow, model_utils.py:139-142)."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ids_ref, log_ref = _hf_loop(ref, feats, T)
        with torch.no_grad():
            la = ours(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
            lb = ref(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return la, lb, ids_ref


def test_config4_medium_pruned_quanto_int8_fp32_flow(H):
    """C4: d = 1024, 16 heads, ffn 4096 (K = 4096 fc2), 4 + 4 layers, prune.global_unstructured 50 % -> quanto
    quantize(weights=qint8) + freeze on the fp32 model (never .half()-ed).  The tensor cores take fp16 operands
    (activations rounded to fp16, int8 codes exact) with fp32 accumulation; tolerance vs a true fp32 matmul:
    2e-2 abs on logits (stated in DESIGN.md section 5), decisive tokens identical, pruned zeros exact."""
    from openai_whisper_compression_b200 import quanto
    master = H.build_model("medium", encoder_layers=4, decoder_layers=4)
    H.global_l1_prune(master, 0.5)
    masks = {n: (m.weight.detach() == 0) for n, m in master.named_modules() if type(m) is nn.Linear}
    frac = sum(int(v.sum()) for v in masks.values()) / sum(v.numel() for v in masks.values())
    assert abs(frac - 0.5) < 1e-3
    import copy
    ours = H.apply_scheme(copy.deepcopy(master), "quanto_int8", "cuda")
    assert H.model_dtype(ours) == torch.float32
    for n, m in ours.named_modules():
        if isinstance(m, quanto.QLinear):
            assert torch.all(m.weight.data[masks[n].cuda()] == 0)        # zeros survive quantization exactly
    ref = _emulate_dequant(H, ours, master.cuda())
    feats = _bench_features(H, 4, 80, False)
    T = 12
    la, lb, _ = _fp32_reference_logits(ours, ref, feats, T)
    d = (la - lb).abs()
    print(f"medium pruned quanto-int8 (fp32 flow) vs fp32 emulation: logits max |diff| {d.max().item():.3e} "
          f"mean {d.mean().item():.3e}, scale {lb.abs().max().item():.2f}")
    assert d.max().item() < 2e-2 and d.mean().item() < 2e-3
    _, decisive = _prefix_until_indecisive(lb, 4e-2)
    assert torch.equal(la.argmax(-1)[decisive], lb.argmax(-1)[decisive])
    _check_fast_vs_hf(H, ours, feats, T, 1e-2, 1e-3)


# ------------------------------------------------------------------------------------------------------------
# C5 -- whisper-large-v3 geometry (d = 1280, 20 heads, 128 mels, vocab 51866), quanto qint8 fp32 flow + LLM.int8
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme", ["quanto_int8", "llm_int8"])
def test_config5_large_v3_geometry(H, scheme):
    """C5: 2 + 2 layers of the true large-v3 width (K = 5120 fc2, N = 51866 projection, 128-mel frontend)."""
    half = scheme == "llm_int8"
    master = H.build_model("large-v3", encoder_layers=2, decoder_layers=2)
    import copy
    ours = H.apply_scheme(copy.deepcopy(master), scheme, "cuda")
    feats = _bench_features(H, 4, 128, half)
    assert feats.shape == (4, 128, 3000)
    T = 12
    if half:
        ref = master.half()
        emu.swap_all(ref, lambda m: emu.EmuLinear8bitLt(m.cuda(), 6.0))
        ref = ref.cuda()
        ids_ref, log_ref = _hf_loop(ref, feats, T)
        with torch.no_grad():
            la = ours(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
            lb = ref(input_features=feats, decoder_input_ids=ids_ref[:, :-1]).logits.float()
        tol = 2e-3
    else:
        ref = _emulate_dequant(H, ours, master.cuda())
        la, lb, _ = _fp32_reference_logits(ours, ref, feats, T)
        tol = 2e-2
    d = (la - lb).abs()
    print(f"large-v3 geometry {scheme} vs emulation: logits max |diff| {d.max().item():.3e} mean {d.mean().item():.3e}")
    assert d.max().item() <= tol
    _, decisive = _prefix_until_indecisive(lb, 2 * tol)
    assert torch.equal(la.argmax(-1)[decisive], lb.argmax(-1)[decisive])
    _check_fast_vs_hf(H, ours, feats, T, 3e-2 if half else 1e-2, 5e-3 if half else 1e-3,
                      expect_fused=True if half else None)

"""The drop-in modules inside HF Whisper, on the GPU: surface parity with the libraries they
mirror, pruned-zero survival, and end-to-end greedy-token equality.

Token criterion: the module under test and a torch-op emulation of the reference linear
(tests/emulation.py, same device / dtype, everything else identical HF code) must produce the
same greedy token ids.  For the integer schemes (LLM.int8 without outliers, torch-dynamic) the
linears are bit-exact, so the whole generation is; for W8A16 / NF4 the fp32 accumulation order
differs from cuBLAS, so tokens are compared under teacher forcing wherever the emulation's top-1
/ top-2 logit margin exceeds the stated logit tolerance.
"""
import io
import os

import numpy as np
import pytest
import torch
from torch import nn

import oracle
from tests import emulation as emu

pytestmark = pytest.mark.gpu

MICRO = dict(encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             d_model=64, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=100)


@pytest.fixture(scope="module")
def pkg():
    import openai_whisper_compression_b200 as p
    from openai_whisper_compression_b200 import bnb, dynamic, harness, quanto, swap, tally, frontend  # noqa: F401
    return p


def _feats(n=4, mels=80, frames=200, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, mels, frames, generator=g) * 0.5


def test_linear4bit_module_surface_and_parity(pkg):
    from openai_whisper_compression_b200 import bnb
    torch.manual_seed(0)
    lin = nn.Linear(128, 96)
    with torch.no_grad():
        lin.weight.mul_(0.3)
        lin.weight[lin.weight.abs() < 0.02] = 0           # pruned
    m = bnb.Linear4bit(128, 96, bias=True, compute_dtype=torch.float32, compress_statistics=False,
                       quant_type="nf4", device=None)
    m.load_state_dict(lin.state_dict(), strict=False)      # bnb_implementation.py:1116
    m = m.to("cuda")                                       # :1221 -> quantizes
    assert isinstance(m, nn.Linear) and m.in_features == 128 and m.out_features == 96
    assert m.weight.dtype == torch.uint8 and m.weight.shape == (128 * 96 // 2, 1)
    qs = m.weight.quant_state
    assert qs.quant_type == "nf4" and qs.blocksize == 64 and tuple(qs.shape) == (96, 128) and qs.dtype == torch.float32
    p_ref, a_ref = oracle.quantize_4bit(lin.weight.detach().numpy(), 64, "nf4")
    np.testing.assert_array_equal(m.weight.data.cpu().numpy(), p_ref)
    np.testing.assert_array_equal(qs.absmax.cpu().numpy(), a_ref)
    # HF's dequantize path: bnb.functional.dequantize_4bit(weight.data, weight.quant_state)
    wd = bnb.dequantize_4bit(m.weight.data, qs)
    assert wd.dtype == torch.float32 and wd.shape == (96, 128)
    assert torch.all(wd[lin.weight.detach().cuda() == 0] == 0)           # zeros survive exactly
    sd = m.state_dict()
    assert all(isinstance(v, torch.Tensor) for v in sd.values())
    assert {"weight", "bias", "weight.absmax", "weight.quant_map"} <= set(sd)
    buf = io.BytesIO()
    torch.save(sd, buf)                                                  # model_utils.py:217-230
    assert buf.tell() < 128 * 96 * 4 / 3
    x = torch.randn(5, 7, 128)
    y = m(x.cuda())
    assert y.dtype == torch.float32 and y.shape == (5, 7, 96)
    y_ref = oracle.linear4bit_forward(x.numpy(), p_ref, a_ref, (96, 128), lin.bias.detach().numpy(),
                                      compute_dtype=np.float16)
    assert np.abs(y.cpu().numpy() - y_ref).max() / max(1.0, np.abs(y_ref).max()) < 2e-3
    # decode-shaped call (one token, batch 1): the weight-only GEMV (as bitsandbytes switches to gemv_4bit) --
    # same operands, fp32 accumulation in another order
    y1 = m(x[:1, :1].cuda())
    np.testing.assert_allclose(y1.cpu().numpy(), y[:1, :1].cpu().numpy(), rtol=0, atol=2e-6)


def test_linear8bitlt_module_surface_and_parity(pkg):
    from openai_whisper_compression_b200 import bnb
    torch.manual_seed(1)
    lin = nn.Linear(128, 64).half()
    m = bnb.Linear8bitLt(128, 64, bias=True, has_fp16_weights=False, threshold=6.0)
    m.load_state_dict(lin.state_dict())
    m = m.to("cuda")
    assert isinstance(m, nn.Linear) and m.weight.dtype == torch.int8
    CB, SCB, _ = oracle.int8_vectorwise_quant(lin.weight.detach().numpy(), 0.0)
    np.testing.assert_array_equal(m.weight.CB.cpu().numpy(), CB)
    np.testing.assert_array_equal(m.weight.SCB.cpu().numpy(), SCB)
    x = torch.randn(3, 9, 128).half()
    x[1, 2, 5] = 8.0
    y = m(x.cuda())
    assert m.state.CB is not None and m.weight.CB is None and m.state.threshold == 6.0
    y_ref, extra = oracle.linear8bitlt_forward(x.numpy(), CB, SCB, lin.bias.detach().numpy(), 6.0)
    assert extra is not None
    assert np.abs(y.float().cpu().numpy() - y_ref.astype(np.float32)).max() <= 2 ** -8
    sd = m.state_dict()
    assert "SCB" in sd and sd["weight"].dtype == torch.int8


def test_quanto_flow_cpu_quantize_then_to_device(pkg):
    from openai_whisper_compression_b200 import quanto
    torch.manual_seed(2)
    model = nn.Sequential(nn.Linear(64, 48), nn.GELU(), nn.Linear(48, 32))
    w0 = model[0].weight.detach().clone()
    b0 = model[0].bias.detach().clone()
    quanto.quantize(model, weights=quanto.qint8)       # model_utils.py:126-128 (CPU model)
    quanto.freeze(model)
    with pytest.raises(RuntimeError):
        model(torch.randn(2, 64))                     # no CPU path
    model = model.to("cuda")                           # :137
    q0 = model[0]
    assert isinstance(q0, quanto.QLinear) and isinstance(q0, nn.Linear) and q0.frozen
    q_ref, s_ref = oracle.quanto_qint8(w0.numpy())
    np.testing.assert_array_equal(q0.weight.data.cpu().numpy(), q_ref)
    sd = model.state_dict()
    assert sd["0.weight._data"].dtype == torch.int8 and sd["0.weight._scale"].shape == (48, 1)
    np.testing.assert_array_equal(sd["0.weight._scale"].cpu().numpy(), s_ref)
    x = torch.randn(6, 64)
    y = q0(x.cuda())
    assert y.dtype == torch.float32
    y_ref = oracle.qlinear_forward(x.half().float().numpy(), q_ref, s_ref, b0.numpy())
    assert np.abs(y.cpu().numpy() - y_ref).max() < 1e-4


def test_pruned_then_quantized_zeros_survive(pkg):
    """prune -> prune.remove -> quantize keeps exact zeros (config 4 flow)."""
    from openai_whisper_compression_b200 import harness, quanto, bnb
    model = harness.build_model("tiny", **MICRO)
    harness.global_l1_prune(model, 0.5)
    masks = {n: (m.weight.detach() == 0) for n, m in model.named_modules() if type(m) is nn.Linear}
    total = sum(int(v.sum()) for v in masks.values()) / sum(v.numel() for v in masks.values())
    assert abs(total - 0.5) < 1e-3
    model = harness.apply_scheme(model, "quanto_int8", "cuda")
    for n, m in model.named_modules():
        if isinstance(m, quanto.QLinear):
            assert torch.all(m.weight.data[masks[n].cuda()] == 0)
    model2 = harness.global_l1_prune(harness.build_model("tiny", **MICRO), 0.5)
    model2 = harness.apply_scheme(model2, "bnb_nf4_direct", "cuda")
    for n, m in model2.named_modules():
        if isinstance(m, bnb.Linear4bit):
            wd = bnb.dequantize_4bit(m.weight.data, m.weight.quant_state)
            assert torch.all(wd[masks[n].cuda()] == 0)


def test_dynamic_int8_twin_matches_reference_generate_golden(pkg, golden_dir):
    """BASELINE config 1 on the micro model: the reference's own load_whisper_model(...,
    quantization="pytorch") + generate on CPU (golden) vs the GPU twin.

    Per-tensor DYNAMIC activation quantization is discontinuous in its input (a 1-ulp change of
    the tensor max moves the scale and flips codes), and the non-linear layers (conv, softmax,
    LayerNorm, GELU) run as different fp32 kernels on CPU and GPU, so whole-model logits agree to
    the int8 quantization noise, not to fp32 rounding: tolerance 5e-2 abs on O(1) logits, greedy
    tokens compared where the golden top-1/top-2 margin exceeds 2x that.  The linears themselves
    are checked bit-exactly per layer below."""
    from openai_whisper_compression_b200 import dynamic, harness
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(golden_dir, "ref_dynamic_generate.npz"))
    model = harness.apply_scheme(harness.build_model("tiny", **MICRO), "dynamic_int8", "cuda")
    feats = torch.from_numpy(g["feats"]).cuda()
    T = int(g["T"])
    gold_ids = torch.from_numpy(g["ids"]).cuda()
    with torch.no_grad():
        logits = model(input_features=feats, decoder_input_ids=gold_ids[:, :-1]).logits.float().cpu().numpy()
    tol = 5e-2
    assert np.abs(logits[:, 0, :] - g["first_logits"]).max() < tol
    decisive = g["margin"] > 2 * tol
    # raw-logit argmax of the golden run vs ours at the same (teacher-forced) positions
    assert np.array_equal(logits.argmax(-1)[decisive], g["top1"][decisive])
    ids = harness.greedy_generate(model, feats, T).cpu().numpy()
    assert ids.shape[0] == g["ids"].shape[0] and ids.shape[1] in (T, T + 1)   # with / without the start token
    # per-layer: every twin linear equals the live torch CPU dynamic module on the same input,
    # up to one activation-code flip at a rounding boundary (|delta| <= s_x * s_w * 127)
    cpu = harness.build_model("tiny", **MICRO)
    torch.quantization.quantize_dynamic(cpu, {nn.Linear}, dtype=torch.qint8, inplace=True)
    cpus = dict(cpu.named_modules())
    worst = []

    def hook(name):
        def f(mod, inp, out):
            x = inp[0]
            e = cpus[name](x.float().cpu())
            d = (out.float().cpu() - e).abs().max().item()
            bound = float(x.abs().max()) / 127 * float(mod.w_scale) * 127 * 2
            worst.append((d, bound, name))
        return f
    hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules()
          if isinstance(m, dynamic.DynamicInt8Linear)]
    with torch.no_grad():
        model(input_features=feats, decoder_input_ids=gold_ids[:, :4])
    for h in hs:
        h.remove()
    assert len(worst) == 33
    assert sum(d == 0.0 for d, _, _ in worst) >= 30          # bit-exact almost everywhere
    assert all(d <= b for d, b, _ in worst), worst


def _teacher_forced_logits(model, feats, ids):
    with torch.no_grad():
        return model(input_features=feats, decoder_input_ids=ids).logits.float()


def test_llm_int8_model_tokens_exact_vs_emulation(pkg):
    """BASELINE config 2 shape of flow (HF load_in_8bit: fp16 model, proj_out kept fp16)."""
    from openai_whisper_compression_b200 import harness
    ours = harness.apply_scheme(harness.build_model("tiny", **MICRO), "llm_int8", "cuda")
    ref = harness.build_model("tiny", **MICRO).half()
    emu.swap_all(ref, lambda m: emu.EmuLinear8bitLt(m.cuda(), 6.0))
    ref = ref.cuda()
    feats = _feats().half().cuda()
    T = 16
    a = harness.greedy_generate(ours, feats, T)
    b = harness.greedy_generate(ref, feats, T)
    np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
    la = _teacher_forced_logits(ours, feats, b)
    lb = _teacher_forced_logits(ref, feats, b)
    assert torch.equal(la, lb)          # no outliers in LayerNorm-ed random activations: bit-exact


@pytest.mark.parametrize("scheme", ["bnb_nf4", "quanto_int8_fp16"])
def test_w4_w8_model_tokens_vs_emulation(pkg, scheme):
    """fp16 logits tolerance 3e-2 abs (logits are O(1); fp16 ulp at 1.0 is 1e-3, 2+2 layers of
    accumulated fp16 rounding differences between fused fp32-accumulate and cuBLAS)."""
    from openai_whisper_compression_b200 import harness, bnb, quanto
    ours = harness.apply_scheme(harness.build_model("tiny", **MICRO), scheme, "cuda")
    ref = harness.build_model("tiny", **MICRO).half().cuda()
    for name, m in list(ours.named_modules()):
        if isinstance(m, bnb.Linear4bit):
            wd = bnb.dequantize_4bit(m.weight.data, m.weight.quant_state)
            emu._set(ref, name, emu.EmuDequantLinear(wd, m.bias))
        elif isinstance(m, quanto.QLinear):
            q, s = m.qweight
            emu._set(ref, name, emu.EmuDequantLinear(q, m.bias, post_scale=s.t()))
    feats = _feats().half().cuda()
    T = 16
    b = harness.greedy_generate(ref, feats, T)
    la = _teacher_forced_logits(ours, feats, b)
    lb = _teacher_forced_logits(ref, feats, b)
    tol = 3e-2
    assert (la - lb).abs().max().item() < tol
    top2 = lb.topk(2, dim=-1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 2 * tol
    assert torch.equal(la.argmax(-1)[decisive], lb.argmax(-1)[decisive])
    a = harness.greedy_generate(ours, feats, T)
    assert a.shape == b.shape


def test_frontend_dropin_matches_hf_extractor(pkg):
    from transformers import WhisperFeatureExtractor
    from openai_whisper_compression_b200.frontend import LogMelFrontend, whisper_mel_filters
    from openai_whisper_compression_b200 import harness
    hf = WhisperFeatureExtractor(feature_size=80)
    np.testing.assert_allclose(whisper_mel_filters(80), hf.mel_filters.astype(np.float32), rtol=0, atol=1e-7)
    fe = LogMelFrontend(80)
    a = harness.synth_audio(7, 300000)        # 18.75 s -> zero padded to 30 s
    got = fe(a, sampling_rate=16000, return_tensors="pt").input_features
    want = hf(a, sampling_rate=16000, return_tensors="pt").input_features
    assert got.shape == want.shape == (1, 80, 3000) and got.dtype == torch.float32 and not got.is_cuda
    assert (got - want).abs().max().item() < 1e-4


def test_tally_matches_oracle(pkg):
    from openai_whisper_compression_b200 import tally
    refs = ["the cat sat on the mat", "a b c", "", "hello world"]
    hyps = ["the cat sat mat", "a x c d", "oops", "hello world"]
    t = tally.tally_on_device(refs, hyps, "cuda").cpu().numpy()
    np.testing.assert_array_equal(t, oracle.wer_cer_tally(refs, hyps))
    m = tally.load_metric("wer")
    with pytest.raises(ValueError):                      # jiwer refuses empty references; so does the shim
        m.compute(references=refs, predictions=hyps)
    refs2, hyps2 = [r for r in refs if r], [h for r, h in zip(refs, hyps) if r]
    t2 = oracle.wer_cer_tally(refs2, hyps2)
    assert abs(m.compute(references=refs2, predictions=hyps2) - t2[0] / t2[1]) < 1e-12
    # jiwer's default transforms: ends stripped, runs of white space collapsed -- for WER and CER alike
    messy_r, messy_h = ["  the  cat sat ", "a\tb  c"], ["the cat   sat", " a b c  "]
    assert m.compute(references=messy_r, predictions=messy_h) == 0.0
    assert tally.load_metric("cer").compute(references=messy_r, predictions=messy_h) == 0.0
    # the C ABI guards its shared-memory diagonals: an over-long pair gets the sentinel -1, nothing is overrun
    from openai_whisper_compression_b200 import functional as F
    big = torch.zeros(5000, dtype=torch.int32, device="cuda")
    off = torch.tensor([0, 5000], dtype=torch.int64, device="cuda")
    assert int(F.edit_distance(big, off, big, off)[0]) == -1


REAL2 = dict(encoder_layers=2, decoder_layers=2)      # whisper-tiny geometry (d = 384, 6 heads x 64), 2 + 2 layers


@pytest.mark.parametrize("scheme,geom", [("llm_int8", "micro"), ("fp16", "micro"), ("llm_int8", "real"),
                                         ("quanto_int8_fp16", "real"), ("bnb_nf4", "real")])
def test_graphed_greedy_matches_hf_generate(pkg, scheme, geom):
    """fastgen: model.generate through the CUDA-graph decode loop returns the same ids as HF's own
    loop (same modules and weights; attention over the static cache may differ in the last fp16
    bits, so positions where HF's own top-1/top-2 margin is below 2e-2 are excluded).  The real-geometry
    cases run the fused decode kernels: attn_decode.cu for every scheme, and for llm_int8 the
    producer-fused step (add+LayerNorm+quant, fused q/k/v GEMM, GELU+quant)."""
    from openai_whisper_compression_b200 import fastgen, harness
    model = harness.apply_scheme(harness.build_model("tiny", **(MICRO if geom == "micro" else REAL2)), scheme, "cuda")
    feats = (_feats(n=6) if geom == "micro" else _feats(n=6, frames=3000)).half().cuda()
    T = 24
    ref = model.generate(feats, do_sample=False, num_beams=1, min_new_tokens=T, max_new_tokens=T,
                         return_dict_in_generate=True, output_logits=True)   # HF loop (fallback criteria)
    ref_ids = harness.greedy_generate(model, feats, T)      # HF's own loop, plain call
    eng = fastgen.enable(model)
    eng.keep_logits = True          # the captured step also stores the raw logits (compared below)
    ids = harness.greedy_generate(model, feats, T)
    assert eng.replays > 0 and eng.fallbacks == 0
    st = next(iter(eng._states.values()))
    assert st.own_attn == (geom == "real")
    assert (st.fused is not None) == (geom == "real" and scheme != "fp16")      # every drop-in scheme has a fused step
    assert st.proj_own == (scheme in ("llm_int8", "fp16", "bnb_nf4"))   # unquantized proj_out: own GEMM + arg-max
    assert ids.shape == ref_ids.shape
    logits = torch.stack(ref.logits, 1).float()
    top2 = logits.topk(2, -1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 2e-2
    # compare up to the first non-decisive position of each utterance (later tokens depend on it)
    first_bad = torch.where(decisive.all(1), decisive.shape[1], (~decisive).float().argmax(1))
    P = ids.shape[1] - T      # 0 or 1 depending on whether generate keeps the start token
    for b in range(ids.shape[0]):
        n = P + int(first_bad[b])
        assert torch.equal(ids[b, :n], ref_ids[b, :n])
    if geom == "micro":
        assert (ids == ref_ids).float().mean().item() > 0.9
    # teacher-forced: replay the captured step on HF's own token history and compare the raw logits of every
    # position with HF's (independent of where greedy paths fork)
    start = torch.full((ids.shape[0], 1), model.config.decoder_start_token_id, dtype=ref_ids.dtype, device="cuda")
    full = ref_ids if P == 1 else torch.cat([start, ref_ids], 1)
    worst, mean = 0.0, 0.0
    for j in range(T):
        st.tok.copy_(full[:, j:j + 1])
        st.pos.fill_(j)
        st.graph.replay()
        diff = (st.logits.float() - logits[:, j]).abs()
        worst, mean = max(worst, diff.max().item()), mean + diff.mean().item() / T
    print(f"[{scheme}-{geom}] teacher-forced logits vs HF loop: max |diff| {worst:.3e}, mean {mean:.3e}, "
          f"logit scale {logits.abs().max().item():.2f}")
    assert worst <= 6e-2 and mean <= 6e-3, (worst, mean)
    # second call re-uses the captured graph
    ids2 = harness.greedy_generate(model, feats, T)
    assert torch.equal(ids, ids2)
    eng.uninstall()
    ids3 = harness.greedy_generate(model, feats, T)
    assert torch.equal(ids3, ref_ids)


def test_quanto_int4_model_flow(pkg):
    """model_utils.py 'quanto_int4' flow: quantize(weights=qint4) + freeze on the CPU model, then
    .to(device); every linear (incl. proj_out) becomes a qint4 QLinear and the model decodes."""
    from openai_whisper_compression_b200 import harness, quanto
    model = harness.apply_scheme(harness.build_model("tiny", **MICRO), "quanto_int4", "cuda")
    qs = [m for m in model.modules() if isinstance(m, quanto.QLinear)]
    assert len(qs) == 33 and all(m.frozen and m._wshift is not None for m in qs)
    fc1 = model.model.decoder.layers[0].fc1
    sd = fc1.state_dict()
    assert sd["weight._data"].dtype == torch.uint8 and sd["weight._data"].shape == (256, 32)
    assert sd["weight._scale"].shape == sd["weight._shift"].shape == (256, 1)
    ref = harness.build_model("tiny", **MICRO).model.decoder.layers[0].fc1
    q_ref, s_ref, sh_ref, g = oracle.quanto_qint4(ref.weight.detach().numpy())
    np.testing.assert_array_equal(sd["weight._data"].cpu().numpy(), oracle.quanto_qint4_pack(q_ref))
    x = torch.randn(3, 5, 64)
    y = fc1(x.cuda()).cpu().numpy()
    wd = oracle.quanto_qint4_dequant(q_ref, s_ref, sh_ref, g).astype(np.float16).astype(np.float64)
    y_ref = x.half().double().numpy() @ wd.T + ref.bias.detach().double().numpy()
    assert np.abs(y - y_ref).max() < 1e-3
    ids = harness.greedy_generate(model, _feats().cuda(), 6)
    assert ids.shape[0] == 4


def test_quanto_int2_model_flow(pkg):
    """dynamic_evaluation_int2.py:158-160: quantize(model, weights=qint2) + freeze, then .to(device): every linear
    becomes a qint2 QLinear (codes 0..3 from the quantizer of the oracle), forward = dequantise-then-matmul."""
    from openai_whisper_compression_b200 import harness, quanto
    model = harness.build_model("tiny", **MICRO)
    quanto.quantize(model, weights=quanto.qint2)
    quanto.freeze(model)
    model = model.to("cuda")
    qs = [m for m in model.modules() if isinstance(m, quanto.QLinear)]
    assert len(qs) == 33 and all(m.frozen and m.weight_qtype.name == "qint2" for m in qs)
    fc1 = model.model.decoder.layers[0].fc1
    ref = harness.build_model("tiny", **MICRO).model.decoder.layers[0].fc1
    q_ref, s_ref, sh_ref, g = oracle.quanto_qint4(ref.weight.detach().numpy(), bits=2)
    sd = fc1.state_dict()
    np.testing.assert_array_equal(sd["weight._data"].cpu().numpy(), oracle.quanto_qint4_pack(q_ref))
    assert int(q_ref.max()) == 3
    x = torch.randn(3, 5, 64)
    y = fc1(x.cuda()).cpu().numpy()
    wd = oracle.quanto_qint4_dequant(q_ref, s_ref, sh_ref, g).astype(np.float16).astype(np.float64)
    y_ref = x.half().double().numpy() @ wd.T + ref.bias.detach().double().numpy()
    assert np.abs(y - y_ref).max() < 1e-3
    ids = harness.greedy_generate(model, _feats().cuda(), 6)
    assert ids.shape[0] == 4


def test_graphed_greedy_early_stop_paths_match_hf(pkg):
    """min_new_tokens < max_new_tokens: after the minimum the loop must evaluate HF's stopping
    criteria every step (EOS possible); before it, it may skip the host sync.  Same ids as HF."""
    from openai_whisper_compression_b200 import fastgen, harness
    model = harness.apply_scheme(harness.build_model("tiny", **MICRO), "fp16", "cuda")
    feats = _feats(n=3).half().cuda()
    kw = dict(do_sample=False, num_beams=1, min_new_tokens=4, max_new_tokens=12)
    ref = model.generate(feats, **kw)
    eng = fastgen.enable(model)
    got = model.generate(feats, **kw)
    assert eng.replays > 0 and eng.fallbacks == 0
    assert got.device == ref.device and got.dtype == ref.dtype      # host post-processing is internal
    assert torch.equal(got, ref)
    # the same call with the post-processing left on the device (HF's default placement)
    eng.host_postprocess = False
    assert torch.equal(model.generate(feats, **kw), ref)
    eng.host_postprocess = True
    # timestamp mode walks HF's segment-splitting code with the host-resident ids
    eng.uninstall()
    model.generation_config.no_timestamps_token_id = 50363     # random-init config: HF needs it for timestamps
    ref_ts = model.generate(feats, return_timestamps=True, **kw)
    eng = fastgen.enable(model)
    got_ts = model.generate(feats, return_timestamps=True, **kw)
    assert got_ts.device == ref_ts.device and torch.equal(got_ts, ref_ts)
    # beams are not covered by the fast loop: falls through to HF's own implementation
    out = model.generate(feats, do_sample=False, num_beams=2, max_new_tokens=4)
    assert out.shape[0] == 3
    eng.uninstall()


@pytest.mark.parametrize("scheme", ["llm_int8", "quanto_int8_fp16"])
def test_copy_free_encoder_attention_is_bit_identical(pkg, scheme):
    """fastenc: the encoder with the copy-free self-attention forward returns exactly what HF's forward
    (q/k/v .contiguous() copies) returns -- same projections, same SDPA arithmetic, only strides differ."""
    from openai_whisper_compression_b200 import fastenc, harness
    model = harness.apply_scheme(harness.build_model("tiny", encoder_layers=2, decoder_layers=1), scheme, "cuda")
    feats = _feats(n=5, frames=3000).half().cuda()       # real geometry: S = 1500, 6 heads x 64
    with torch.no_grad():
        ref = model.model.encoder(feats).last_hidden_state
        assert fastenc.enable(model, fuse_int8=False) == len(model.model.encoder.layers)
        out = model.model.encoder(feats).last_hidden_state
        fastenc.disable(model)
        back = model.model.encoder(feats).last_hidden_state
    assert torch.equal(back, ref)
    assert torch.isfinite(out).all()
    diff = (out.float() - ref.float()).abs().max().item()
    assert diff <= 2e-3, diff      # expected 0.0: cuDNN picks the same kernel for both layouts


def test_fused_int8_decode_step_is_bit_identical_to_module_calls(pkg):
    """The producer-fused LLM.int8 decode step (fused q/k/v GEMM over concatenated weights, quantization done by
    the LayerNorm / GELU / attention kernels, residual adds inside the LayerNorm launch) must give exactly the
    logits of the same step written with the drop-in modules -- Linear8bitLt.forward quantizing its own input --
    around the same LayerNorm / GELU / attention kernels."""
    import torch.nn.functional as TF
    from openai_whisper_compression_b200 import fastgen, harness, functional as F
    model = harness.apply_scheme(harness.build_model("tiny", **REAL2), "llm_int8", "cuda")
    with torch.no_grad():        # a few loud channels per LayerNorm, so that the outlier decomposition is in play
        for m in model.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight[::41] *= 6.0
    feats = _feats(n=6, frames=3000).half().cuda()
    eng = fastgen.enable(model)
    eng.keep_logits = True
    T = 12
    ids = harness.greedy_generate(model, feats, T)
    st = next(iter(eng._states.values()))
    assert st.fused is not None and st.own_cross
    dec = model.model.decoder
    B, H, d = st.B, st.H, st.d
    k2 = [torch.zeros_like(t) for t in st.k]
    v2 = [torch.zeros_like(t) for t in st.v]

    def module_step(tok, pos):
        x = dec.embed_tokens(tok).view(B, d) + dec.embed_positions.weight.index_select(0, pos)
        for li, layer in enumerate(dec.layers):
            sa, ca = layer.self_attn, layer.encoder_attn
            ln = layer.self_attn_layer_norm
            h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
            a, _ = F.self_attn_decode(sa.q_proj(h), sa.k_proj(h), sa.v_proj(h), sa.scaling, k2[li], v2[li], pos, H)
            x = x + sa.out_proj(a)
            ln = layer.encoder_attn_layer_norm
            h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
            a, _ = F.cross_attn_decode(ca.q_proj(h), st.ckv[li][:, :, :d], st.ckv[li][:, :, d:], ca.scaling, H)
            x = x + ca.out_proj(a)
            ln = layer.final_layer_norm
            h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
            x = x + layer.fc2(F.gelu_quant(layer.fc1(h))[0])
        ln = dec.layer_norm
        h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
        return model.proj_out(h)

    start = torch.full((B, 1), model.config.decoder_start_token_id, dtype=ids.dtype, device="cuda")
    full = ids if ids.shape[1] == T + 1 else torch.cat([start, ids], 1)
    with torch.no_grad():
        # cross-attention K/V: one shared quantization of the encoder output + [Wk; Wv] GEMMs (flags kept between
        # consumers) == every k_proj / v_proj module quantizing it again
        enc_out = model.model.encoder(feats).last_hidden_state
        assert (enc_out.float().abs() >= 6.0).any()          # outlier columns are in play
        for li, layer in enumerate(dec.layers):
            assert torch.equal(st.ck[li].reshape(B, -1, d), layer.encoder_attn.k_proj(enc_out))
            assert torch.equal(st.cv[li].reshape(B, -1, d), layer.encoder_attn.v_proj(enc_out))
        from openai_whisper_compression_b200.functional import OutlierState
        assert int(OutlierState.get(enc_out.device, d).col_flags.abs().sum()) == 0      # last consumer cleared them
        for j in range(T):
            st.tok.copy_(full[:, j:j + 1])
            st.pos.fill_(j)
            st.graph.replay()
            ref = module_step(full[:, j:j + 1], st.pos)
            # proj_out: same fp16 weights through cuBLAS on a padded copy -> compare the hidden state exactly via
            # the logits of the SAME projection
            got = st.logits
            ref_same_proj = None
            assert got.shape == ref.shape
            if not torch.equal(got, ref):
                # the padded-weight projection may pick another cuBLAS kernel: allow its fp16 rounding only
                assert (got.float() - ref.float()).abs().max().item() <= 4e-3, j
            for li in range(len(dec.layers)):
                assert torch.equal(st.k[li][:, : j + 1], k2[li][:, : j + 1]), (j, li)     # bit-identical caches
                assert torch.equal(st.v[li][:, : j + 1], v2[li][:, : j + 1]), (j, li)


def test_fused_int8_encoder_layer_is_bit_identical_to_module_calls(pkg):
    """fastenc's producer-fused WhisperEncoderLayer (LLM.int8): exactly the output of the same layer written with
    the drop-in modules around the same LayerNorm / GELU kernels, and close to HF's own forward (LayerNorm
    one-ulp differences pass through the int8 quantizers)."""
    import torch.nn.functional as TF
    from openai_whisper_compression_b200 import fastenc, harness, functional as F
    model = harness.apply_scheme(harness.build_model("tiny", encoder_layers=2, decoder_layers=1), "llm_int8", "cuda")
    feats = _feats(n=3, frames=3000).half().cuda()
    enc = model.model.encoder

    def module_layer(layer, x):
        B, S, d = x.shape
        sa = layer.self_attn
        ln = layer.self_attn_layer_norm
        h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
        q, k, v = (p(h).view(B, S, sa.num_heads, 64).transpose(1, 2) for p in (sa.q_proj, sa.k_proj, sa.v_proj))
        a = TF.scaled_dot_product_attention(q, k, v, scale=sa.scaling).transpose(1, 2).reshape(B, S, d)
        x = x + sa.out_proj(a)
        ln = layer.final_layer_norm
        h = F.add_layernorm_quant(x, None, ln.weight, ln.bias, ln.eps)[1]
        x = x + layer.fc2(F.gelu_quant(layer.fc1(h))[0])
        c = torch.finfo(torch.float16).max - 1000
        return torch.clamp(x, min=-c, max=c)

    with torch.no_grad():
        hf = enc(feats).last_hidden_state
        x0 = TF.gelu(enc.conv2(TF.gelu(enc.conv1(feats)))).permute(0, 2, 1) + enc.embed_positions.weight
        x0 = x0.contiguous()
        ref = x0
        for layer in enc.layers:
            ref = module_layer(layer, ref)
        fastenc.enable(model)
        got = x0
        for layer in enc.layers:
            got = layer(got, None)
        assert all(getattr(layer, "_whisperq_plan", None) is not None for layer in enc.layers)
        full = enc(feats).last_hidden_state
        fastenc.disable(model)
    assert torch.equal(got, ref)
    assert torch.isfinite(full).all()
    diff = (full.float() - hf.float()).abs()
    assert diff.max().item() <= 0.15 and diff.mean().item() <= 1e-2, (diff.max().item(), diff.mean().item())


@pytest.mark.parametrize("B,streams", [(6, 1), (70, 2), (64, 4)])
def test_persistent_decoder_layer_kernel_is_bit_identical_to_the_launch_chain(pkg, B, streams):
    """decode_fused.cu (one persistent launch between two cross-attention passes: LayerNorm+quant, int8 GEMMs with
    dp4a, self-attention over the cache, GELU+quant, residual adds, phases separated by a grid barrier; opt-in, it
    measured slower than the chain) must give exactly the logits, greedy ids and KV caches of the twelve-launch chain
    (_decoder_step_int8), with outlier columns in play, for one row group and for several on several streams."""
    from openai_whisper_compression_b200 import fastgen, harness
    model = harness.apply_scheme(harness.build_model("tiny", **REAL2), "llm_int8", "cuda")
    with torch.no_grad():        # loud channels so that the outlier decomposition is exercised in every phase
        for m in model.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight[::41] *= 6.0
    feats = _feats(n=B, frames=3000).half().cuda()
    eng = fastgen.enable(model)
    eng.keep_logits, eng.streams = True, streams
    T = 10
    outs = {}
    for mega in (False, True):
        eng.mega = mega
        ids = harness.greedy_generate(model, feats, T)
        st = [s for s in eng._states.values() if s.mega == mega][-1]
        assert st.mega == mega and len(st.views) == streams
        start = torch.full((B, 1), model.config.decoder_start_token_id, dtype=ids.dtype, device="cuda")
        full = ids if ids.shape[1] == T + 1 else torch.cat([start, ids], 1)
        logits = []
        for j in range(T):
            st.tok.copy_(full[:, j:j + 1])
            st.pos.fill_(j)
            st.graph.replay()
            logits.append(st.logits.clone())
        outs[mega] = (ids, torch.stack(logits), [k.clone() for k in st.k], [v.clone() for v in st.v])
    assert torch.equal(outs[True][0], outs[False][0])
    assert torch.equal(outs[True][1], outs[False][1])
    for a, b in zip(outs[True][2] + outs[True][3], outs[False][2] + outs[False][3]):
        assert torch.equal(a[:, :T], b[:, :T])
    assert (outs[True][1].float().abs() > 0).any()


def test_before_readback_hook_runs_once_per_generate_and_changes_nothing(pkg):
    """fastgen's `before_readback` hook (bench.py's end-to-end arm: the previous batch's transcript + tally run on the
    host while this batch executes): called exactly once per plain generate call, after the batch's replays are
    queued; the returned ids are the ids without the hook; work the hook queues on the stream is ordered after the
    batch (its result is right once the stream is drained)."""
    from openai_whisper_compression_b200 import fastgen, harness
    model = harness.apply_scheme(harness.build_model("tiny", **REAL2), "llm_int8", "cuda")
    feats = _feats(n=4, frames=3000).half().cuda()
    T = 8
    eng = fastgen.enable(model)
    want = harness.greedy_generate(model, feats, T)
    calls = []
    probe = torch.zeros(1, device="cuda")

    def hook():
        calls.append(eng.replays)          # every replay of this call has been queued by now
        probe.add_(1.0)                    # stream-ordered work from inside the hook

    eng.before_readback = hook
    r0 = eng.replays
    got = harness.greedy_generate(model, feats, T)
    eng.before_readback = None
    assert torch.equal(got, want)
    assert len(calls) == 1 and calls[0] - r0 >= T
    torch.cuda.synchronize()
    assert probe.item() == 1.0
    again = harness.greedy_generate(model, feats, T)       # hook removed: not called again
    assert torch.equal(again, want) and len(calls) == 1
    eng.uninstall()

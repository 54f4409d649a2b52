"""optimum-quanto static activation quantization and float8 weights on the GPU (the
reference's six "static_quanto_<weights>_<activations>" configs, quantization.py:53-86, through
model_utils.py:152-214: ``quantize(model, weights=, activations=)``, evaluation inside ``with Calibration():``,
``freeze(model)`` on the fp16 model that load_whisper_model(quantization=None) returns).

quanto is not installed (parity unpinned); the checker here is a torch-op emulation of the
flow as restated in quanto.py's header, run in the same dtype on the same device with the SAME calibrated scales:
  * per module: int8 x int8 products are exact integers, so outputs agree except where the two evaluation orders of
    `acc * scale + bias` (fused fp32 vs torch's fp16 steps) land on different sides of an output-grid rounding
    boundary: >= 99.5 % of the elements identical, the rest one grid step apart;
  * per model: all six weight/activation combinations calibrate, freeze, save and decode; quantization noise
    against the unquantized fp16 model stays bounded.
"""
import io

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

MICRO = dict(encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             d_model=64, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=100)


def _qt(Q, name):
    return {"int8": Q.qint8, "int4": Q.qint4, "float8": Q.qfloat8}[name]


def _emu_act(x, scale, name):
    """quantize_activation + dequantize with torch ops in the dtype of x."""
    s = scale.to(x.dtype)
    if name == "int8":
        g = torch.clamp(torch.round(x / s), -128, 127)
    else:
        g = torch.clamp(x / s, -448.0, 448.0).to(torch.float8_e4m3fn).to(x.dtype)
    return g, g * s


@pytest.mark.parametrize("wname", ["int8", "float8", "int4"])
@pytest.mark.parametrize("aname", ["int8", "float8"])
def test_static_qlinear_matches_torch_emulation(wname, aname):
    from openai_whisper_compression_b200 import quanto as Q
    torch.manual_seed(7)
    K, N, M = 256, 192, 300
    lin = nn.Linear(K, N).half().cuda()
    with torch.no_grad():
        lin.weight.mul_(0.5)
    w0, b0 = lin.weight.detach().clone(), lin.bias.detach().clone()
    model = nn.Sequential(lin)
    Q.quantize(model, weights=_qt(Q, wname), activations=_qt(Q, aname))
    q = model[0]
    assert isinstance(q, Q.QLinear) and isinstance(q, nn.Linear)
    xs = [torch.randn(M, K, device="cuda").half() * (1.0 + 0.2 * i) for i in range(3)]
    with Q.Calibration():
        for x in xs:
            model(x)
    Q.freeze(model)
    assert q.frozen and float(q.input_scale) != 1.0 and float(q.output_scale) != 1.0
    # momentum 0.9: first observation, then 0.9 * old + 0.1 * new (evaluated in fp16 like the buffers)
    qmax = 127.0 if aname == "int8" else 448.0
    exp = xs[0].abs().max() / qmax
    for x in xs[1:]:
        exp = (0.9 * exp + 0.1 * (x.abs().max() / qmax)).to(torch.float16)
    assert abs(float(q.input_scale) - float(exp)) <= 2e-3 * float(exp)
    sd = model.state_dict()
    assert {"0.weight._data", "0.weight._scale", "0.input_scale", "0.output_scale", "0.bias"} <= set(sd)
    buf = io.BytesIO()
    torch.save(sd, buf)

    x = xs[1]
    y = model(x)
    assert y.dtype == torch.float16 and y.shape == (M, N)
    # emulation with the module's own codes and scales
    g, xd = _emu_act(x, q.input_scale, aname)
    if wname == "int4":
        from openai_whisper_compression_b200 import functional as F
        eye = torch.eye(K, device="cuda").half()
        wd = F.gemm_u4a16(eye, q._wq, q._wscale, q._wshift, q._group).t().contiguous()      # dequantised weight
        pre = (xd.float() @ wd.float().t()).half() + b0
    else:
        if wname == "int8":
            wv = q._wq.float()
            np.testing.assert_array_equal(q._wq.cpu().numpy(),
                                          torch.clamp(torch.round(w0 / (w0.abs().amax(1, keepdim=True) / 127)), -128, 127)
                                          .to(torch.int8).cpu().numpy())
        else:
            wv = q._wq.view(torch.float8_e4m3fn).float()
        os_ = (q.input_scale.half() * q._wscale.view(-1).half()).float()
        pre = ((g.double() @ wv.double().t()).float() * os_[None, :]).half() + b0
    _, ref = _emu_act(pre, q.output_scale, aname)
    step = float(q.output_scale) if aname == "int8" else float(q.output_scale) * 32.0   # e4m3 grid spacing near the top
    same = (y == ref).float().mean().item()
    worst = (y.float() - ref.float()).abs().max().item()
    print(f"W {wname} / A {aname}: identical {same:.5f}, max |diff| {worst:.3e} (output grid step {step:.3e})")
    assert same >= 0.99 and worst <= 1.1 * step      # one grid step, itself rounded to fp16


@pytest.mark.parametrize("wname,aname", [("int8", "int8"), ("int4", "int8"), ("int8", "float8"), ("int4", "float8"),
                                         ("float8", "int8"), ("float8", "float8")])
def test_static_quanto_model_flow(wname, aname):
    """The reference's apply_static_quantization sequence on a (micro) fp16 Whisper on the device: quantize ->
    Calibration over model.generate -> freeze -> generate; LayerNorms become QLayerNorm, linears QLinear."""
    from openai_whisper_compression_b200 import fastgen, harness, quanto as Q
    model = harness.build_model("tiny", **MICRO).half().cuda()
    ref = harness.build_model("tiny", **MICRO).half().cuda()
    Q.quantize(model, weights=_qt(Q, wname), activations=_qt(Q, aname))
    n_lin = sum(isinstance(m, Q.QLinear) for m in model.modules())
    n_ln = sum(isinstance(m, Q.QLayerNorm) for m in model.modules())
    assert n_lin == 33 and n_ln == 12
    g = torch.Generator().manual_seed(3)
    feats = (torch.randn(4, 80, 200, generator=g) * 0.5).half().cuda()
    eng = fastgen.enable(model)
    with Q.Calibration():
        harness.greedy_generate(model, feats, 6)
    assert eng.replays == 0 and eng.fallbacks > 0            # calibration keeps HF's loop (scales change per call)
    Q.freeze(model)
    for m in model.modules():
        if isinstance(m, (Q.QLinear, Q.QLayerNorm)):
            assert float(m.output_scale) != 1.0
        if isinstance(m, Q.QLinear):
            assert m.frozen
    ids = harness.greedy_generate(model, feats, 6)
    assert ids.shape[0] == 4 and eng.replays > 0
    dec = torch.full((4, 4), model.config.decoder_start_token_id, dtype=torch.long, device="cuda")
    with torch.no_grad():
        la = model(input_features=feats, decoder_input_ids=dec).logits.float()
        lb = ref(input_features=feats, decoder_input_ids=dec).logits.float()
    assert torch.isfinite(la).all()
    rel = (la - lb).norm().item() / lb.norm().item()
    print(f"static W {wname} / A {aname}: relative logit error vs the unquantized fp16 model {rel:.3f}")
    assert rel < (0.6 if wname == "int4" else 0.35)
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)          # model_utils.get_model_disk_size_in_mb
    assert all(isinstance(v, torch.Tensor) for v in model.state_dict().values())


def test_qfloat8_weights_only_flow():
    """quantize(model, weights=qfloat8); freeze(model) on the CPU model, then .to(device) (model_utils.py:126-137
    order): every linear holds e4m3 codes + per-channel scales, zeros stay zeros, the model decodes."""
    from openai_whisper_compression_b200 import harness, quanto as Q
    import oracle
    model = harness.build_model("tiny", **MICRO)
    harness.global_l1_prune(model, 0.5)
    w_ref = model.model.decoder.layers[0].fc1.weight.detach().clone()
    Q.quantize(model, weights=Q.qfloat8)
    Q.freeze(model)
    model = model.to("cuda")
    fc1 = model.model.decoder.layers[0].fc1
    assert fc1.frozen and fc1._wq.dtype == torch.uint8
    q_ref, s_ref = oracle.quanto_qfloat8(w_ref.numpy(), "float32")
    np.testing.assert_array_equal(fc1._wq.cpu().numpy(), q_ref)
    np.testing.assert_array_equal(fc1._wscale.cpu().numpy(), s_ref)
    assert torch.all((fc1._wq[(w_ref == 0).cuda()] & 0x7f) == 0)
    x = torch.randn(3, 5, 64)
    y = fc1(x.cuda()).cpu()
    wd = torch.from_numpy(q_ref).view(torch.float8_e4m3fn).float() * torch.from_numpy(s_ref)
    y_ref = x.half().float() @ wd.t() + model.model.decoder.layers[0].fc1.bias.detach().float().cpu()
    assert (y - y_ref).abs().max().item() < 1e-4
    g = torch.Generator().manual_seed(3)
    ids = harness.greedy_generate(model, (torch.randn(4, 80, 200, generator=g) * 0.5).cuda(), 6)
    assert ids.shape[0] == 4

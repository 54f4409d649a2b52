"""CTA-pair schedule of the int8 x int8 tensor-core GEMM (csrc/gemm_tc.cu, PAIR: tcgen05 cta_group::2, 256 x 256 per
cluster of two CTAs) through the C ABI: every call that is not decode-shaped and has an even number of 128-column
blocks takes it.  Exact integer identities against torch integer matmuls on the same device at shapes that exercise
the weight-stationary and the streaming variant, ragged last row tiles (the peer CTA's 128 rows partly or wholly past
M), ragged column tiles, the LLM.int8 outlier decomposition and the fused residual epilogue.

Reference call sites: bitsandbytes Linear8bitLt forward, torch dynamic quantized Linear (A.4), quanto
qbytes_int_mm (model_utils.py:152-214)."""
import numpy as np
import pytest
import torch

from openai_whisper_compression_b200 import functional as F
from tests.emulation import EmuLinear8bitLt

pytestmark = pytest.mark.gpu

# (M, N, K): tiles of 256 x 256 -- ws = weight-stationary (K <= 512 and enough row blocks per column group)
PAIR_SHAPES = [
    (24077, 1000, 512),    # ws; last row tile holds 13 rows (peer CTA entirely past M); ragged last column block
    (24000 + 200, 2048, 512),  # ws; last row tile 136 rows: the peer CTA holds 8
    (9800, 768, 3072),     # streaming ring, 3 column blocks, K = 24 k-blocks
    (5000, 2048, 384),     # streaming (too few row blocks for the stationary schedule), K = 3 k-blocks
    (80000, 256, 256),     # ws, one column block, 2 k-blocks
    (20000, 256, 128),     # streaming, one k-block per tile
]


def _llmint8_reference(x, cb, scb, bias):
    ca, sca, _ = F.int8_vectorwise_quant(x, 0.0)
    c32 = torch._int_mm(ca, cb.t().contiguous())
    v = (c32.float() * sca[:, None]) * scb[None, :]
    y = v.double() * float(np.float32(6.200012e-05))
    if bias is not None:
        y = y + bias.double()[None, :]
    return y.float().half()


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_pair_llmint8_equals_integer_matmul(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = (torch.randn(N, device="cuda", generator=g) * 0.1).half()
    cb, scb, _ = F.int8_vectorwise_quant(W, 0.0)
    y = F.linear8bitlt(x, cb, scb, bias, 6.0)
    ref = _llmint8_reference(x, cb, scb, bias)
    mism = (y != ref)
    # fmaf vs (exact double product + add, rounded twice) can differ only in double-rounding ties
    assert mism.float().mean().item() < 1e-6
    assert (y.float() - ref.float()).abs().max().item() <= 2 ** -10 * max(1.0, ref.float().abs().max().item())
    # every row tile, both CTAs of the pair, first and last column block
    for r in (0, 127, 128, 255, 256, M - 1):
        assert torch.equal(y[r, :8], ref[r, :8]) or mism[r, :8].sum().item() <= 1
        assert torch.equal(y[r, -8:], ref[r, -8:]) or mism[r, -8:].sum().item() <= 1


@pytest.mark.parametrize("M,N,K", [(24077, 1000, 512), (9800, 768, 3072)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_pair_w8a8_exact(M, N, K, out_dtype):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    xq = torch.randint(-128, 128, (M, K), device="cuda", generator=g, dtype=torch.int8)
    wq = torch.randint(-128, 128, (N, K), device="cuda", generator=g, dtype=torch.int8)
    os_ = torch.rand(N, device="cuda", generator=g) * 1e-3
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    y = F.gemm_w8a8(xq, wq, os_, bias, out_dtype)
    acc = torch._int_mm(xq, wq.t().contiguous())
    ref = ((acc.float() * os_[None, :]) + bias[None, :]).to(out_dtype)
    assert torch.equal(y, ref)


@pytest.mark.parametrize("M,N,K", [(24077, 1000, 512), (5000, 2048, 384)])
def test_pair_torch_dynamic_exact(M, N, K):
    """u8 activations x s8 weights, fp32 out: (acc - zp * wsum[n]) * (s_x * s_w) + bias[n] with exact integer sums."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    x = torch.randn(M, K, device="cuda", generator=g) * 2
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    q, scale, wsum = F.torch_quantize_weight(w)
    xq, qparams = F.torch_quantize_activation(x)
    y = F.gemm_dyn_i8(xq, qparams, q, scale, wsum, bias)
    qp = qparams.cpu()
    s_x, zp = float(qp[0]), int(qp[1])
    acc = torch._int_mm(xq.view(torch.int8), q.t().contiguous())      # codes are 0..127: the same bytes as int8
    dyn_s = (torch.tensor(s_x, dtype=torch.float32) * scale.cpu().float().reshape(-1)[0]).item()
    ref = (acc - zp * wsum[None, :]).float() * np.float32(dyn_s) + bias[None, :]
    assert torch.equal(y, ref)


def test_pair_llmint8_outlier_decomposition():
    M, N, K = 4900, 1024, 512                  # 20 x 4 pair tiles, last row tile 36 rows
    g = torch.Generator(device="cuda").manual_seed(7)
    lin = torch.nn.Linear(K, N).cuda().half()
    with torch.no_grad():
        lin.weight.copy_((torch.randn(N, K, device="cuda", generator=g) * 0.05).half())
    x = torch.randn(M, K, device="cuda", generator=g).half().clamp_(-5.5, 5.5)
    for (r, c, v) in [(1, 3, 9.0), (M - 1, K - 2, -12.5), (M // 2, 3, 6.0), (300, K // 2, 30.0)]:
        x[r, c] = v
    emu = EmuLinear8bitLt(lin, 6.0)
    cb, scb = emu.CB.contiguous(), emu.SCB.contiguous()
    y = F.linear8bitlt(x, cb, scb, lin.bias.detach(), 6.0)
    ref = emu(x)
    scale = ref.float().abs().clamp_min(1.0)
    assert ((y.float() - ref.float()).abs() / scale).max().item() <= 2 ** -9
    # a second call sees clean outlier flags (self-cleaning by the last CTA of the pair grid)
    x2 = x.clamp(-5.5, 5.5)
    y2 = F.linear8bitlt(x2, cb, scb, lin.bias.detach(), 6.0)
    ref2 = _llmint8_reference(x2, cb, scb, lin.bias.detach())
    assert (y2 != ref2).float().mean().item() < 1e-6


def test_pair_fused_residual_epilogue_matches_separate_add():
    """fc2 with the residual add + fp16 clamp folded into the epilogue (wq_gemm_llmint8_residual) == GEMM, then add, then
    clamp, on a pair-scheduled shape."""
    M, N, K = 9100, 512, 2048
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(M, K, device="cuda", generator=g).half().clamp_(-5.5, 5.5)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = (torch.randn(N, device="cuda", generator=g) * 0.1).half()
    res = (torch.randn(M, N, device="cuda", generator=g) * 3).half()
    cb, scb, _ = F.int8_vectorwise_quant(W, 0.0)
    ca, sca, _ = F.int8_vectorwise_quant(x, 0.0)
    y = F.gemm_llmint8(ca, sca, cb, scb, bias)
    clamp = float(torch.finfo(torch.float16).max) - 1000.0
    want = (y + res).clamp(-clamp, clamp)
    got = F.gemm_llmint8(ca, sca, cb, scb, bias, residual=res, clamp_abs=clamp)
    assert torch.equal(got, want)

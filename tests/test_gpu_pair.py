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
    x = torch.randn(M, K, device="cuda", generator=g).clamp_(-5.5, 5.5).half()     # below the outlier threshold
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


# ------------------------------------------------------------------------------------------------
# weight-expanding schemes on CTA pairs: each CTA expands its 128 of the tile's 256 W rows (bready barrier of the
# leader counts both CTAs' expansion warps); fp32 output against a float64 matmul of the dequantized weights
# ------------------------------------------------------------------------------------------------
A16_SHAPES = [(9800, 768, 768), (4900, 1024, 4096), (12077, 1280, 1280), (2600, 3072, 768)]


# fp32 accumulation over K products: the error against float64 grows ~ sqrt(K) * 2^-24 * sum|x w|; 5e-5 of max(|y|, 1)
# covers K = 4096 with 2x headroom (the K <= 1280 oracle tests of test_gpu_kernels.py use 2e-5)
def _check_a16(y32, x, wd, bias, tol32=5e-5):
    ref = x.double() @ wd.double().t() + bias.double()[None, :]
    err = (y32.double() - ref).abs() / ref.abs().clamp_min(1.0)
    assert err.max().item() <= tol32
    return ref


@pytest.mark.parametrize("M,N,K", A16_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_pair_w8a16_matches_dequant_matmul(M, N, K, dtype):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    w[torch.rand(N, K, device="cuda", generator=g) < 0.5] = 0
    x = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    q, scale = F.quanto_quantize_qint8(w)
    y32 = F.gemm_w8a16(x, q, scale, bias, out_dtype=torch.float32)
    ref = x.double() @ q.double().t() * scale.double().view(1, -1) + bias.double()[None, :]
    err = (y32.double() - ref).abs() / ref.abs().clamp_min(1.0)
    assert err.max().item() <= 5e-5
    y = F.gemm_w8a16(x, q, scale, bias)
    assert y.dtype == dtype
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert ((y.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item() <= tol
    # every row block of both CTAs of a pair equals the rounded fp32 result
    assert torch.equal(y, y32.to(dtype))


@pytest.mark.parametrize("M,N,K", A16_SHAPES)
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
def test_pair_w4a16_matches_dequant_matmul(M, N, K, quant_type):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    x = torch.randn(M, K, device="cuda", generator=g).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    packed, absmax = F.quantize_4bit(w, 64, quant_type)
    wd = F.dequantize_4bit(packed, absmax, (N, K), 64, quant_type, torch.float16)
    y32 = F.gemm_w4a16(x, packed, absmax, N, K, bias, quant_type, out_dtype=torch.float32)
    _check_a16(y32, x, wd, bias)
    y = F.gemm_w4a16(x, packed, absmax, N, K, bias, quant_type)
    assert torch.equal(y, y32.half())


@pytest.mark.parametrize("M,N,K", [(9800, 768, 768), (4900, 1024, 4096)])
def test_pair_u4a16_and_f8_match_dequant_matmul(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 2)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    x = torch.randn(M, K, device="cuda", generator=g).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    p, s, sh, grp = F.quanto_quantize_qint4(w)
    y32 = F.gemm_u4a16(x, p, s, sh, grp, bias, torch.float32)
    codes = torch.stack([p >> 4, p & 15], -1).reshape(N, K).float()
    wd = (s.repeat_interleave(grp, 1) * codes - sh.repeat_interleave(grp, 1)).half()
    _check_a16(y32, x, wd, bias)
    q8, s8 = F.quanto_quantize_qfloat8(w)
    y32 = F.gemm_wf8a16(x, q8, s8.view(-1), bias, torch.float32)
    ref = x.double() @ q8.view(torch.float8_e4m3fn).float().double().t() * s8.double().view(1, -1) + bias.double()[None, :]
    assert ((y32.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item() <= 5e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(256, 51865, 512), (3000, 2048, 384)])
def test_pair_gemm_f16_projection_and_argmax(dtype, M, N, K):
    """Unquantized projection (proj_out of the HF bitsandbytes flows) on pairs, with the masked arg-max epilogue."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dtype)
    y = F.gemm_f16(x, w)
    ref = (x.double() @ w.double().t())
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert ((y.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item() <= tol
    mask = torch.zeros(-(-N // 256) * 256, dtype=torch.bool, device="cuda")
    mask[torch.randint(0, N, (N // 3,), device="cuda", generator=g)] = True
    keys = torch.zeros(M, dtype=torch.long, device="cuda")
    F.gemm_f16(x, w, mask=mask, argmax_keys=keys, store=False)
    ids = F.argmax_finalize(keys)
    want = y.float().masked_fill(mask[:N][None, :], float("-inf")).argmax(-1)
    assert torch.equal(ids, want)

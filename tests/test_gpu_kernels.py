"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle on seeded inputs.

Bars: bit-exact for codes, scales, integer accumulators and the fp32-formula epilogues that the
oracle restates operation by operation (LLM.int8 without outliers, torch-dynamic requant);
stated tolerances for floating-point accumulation (W8A16 / W4A16 GEMMs, outlier addmm, log-mel).
"""
import numpy as np
import pytest
import torch

import oracle
from tests.helpers import synth_audio

pytestmark = pytest.mark.gpu

F = None


@pytest.fixture(scope="module", autouse=True)
def _load():
    global F
    from openai_whisper_compression_b200 import functional
    F = functional
    yield


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


# ------------------------------------------------------------------------------------------------
# 4-bit packing
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
@pytest.mark.parametrize("shape,blocksize", [((96, 128), 64), ((33, 70), 64), ((7, 67), 64), ((64, 512), 256),
                                             ((1, 64), 64), ((3, 4096), 4096)])
def test_quantize_4bit_bit_exact(dtype, quant_type, shape, blocksize):
    rng = np.random.RandomState(hash((shape, blocksize)) % 2**31)
    w = (rng.randn(*shape) * 0.02).astype(np.float32)
    w[rng.rand(*shape) < 0.3] = 0.0          # pruned weights
    if shape[0] > 2:
        w[2] = 0.0                            # all-zero blocks (absmax 0)
    wt = dev(w, dtype)
    w_ref = wt.float().cpu().numpy()          # exact values the kernel sees
    packed, absmax = F.quantize_4bit(wt, blocksize, quant_type)
    p_ref, a_ref = oracle.quantize_4bit(w_ref, blocksize, quant_type)
    assert packed.shape == p_ref.shape and absmax.shape == a_ref.shape
    np.testing.assert_array_equal(absmax.cpu().numpy(), a_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), p_ref)
    for out_dtype, np_dtype in ((torch.float16, np.float16), (torch.float32, np.float32)):
        back = F.dequantize_4bit(packed, absmax, shape, blocksize, quant_type, out_dtype)
        ref = oracle.dequantize_4bit(p_ref, a_ref, shape, blocksize, quant_type, np_dtype)
        np.testing.assert_array_equal(back.cpu().numpy(), ref)
    # pruned zeros survive the round trip exactly (SURVEY section 8 a8 contract) -- NF4 only:
    # FP4 has no exact interior... it does (code 0), check both
    back32 = F.dequantize_4bit(packed, absmax, shape, blocksize, quant_type, torch.float32).cpu().numpy()
    assert np.all(back32[w_ref == 0] == 0)


def test_quantize_4bit_empty():
    w = torch.empty((0,), dtype=torch.float16, device="cuda")
    packed, absmax = F.quantize_4bit(w)
    assert packed.shape == (0, 1) and absmax.shape == (0,)


def test_cpu_tensor_raises():
    with pytest.raises(RuntimeError):
        F.quantize_4bit(torch.zeros(64))


# ------------------------------------------------------------------------------------------------
# LLM.int8
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(16, 96), (5, 384), (130, 512), (3, 50), (1, 2048)])
@pytest.mark.parametrize("threshold", [0.0, 6.0])
def test_int8_vectorwise_quant_bit_exact(rows, cols, threshold):
    rng = np.random.RandomState(rows * 1000 + cols)
    a = rng.randn(rows, cols).astype(np.float16)
    if threshold > 0:
        a[rows // 2, cols // 3] = 7.5
        a[0, cols - 1] = -6.0
        a[rows - 1, 0] = 100.0
    if rows > 3:
        a[3] = 0
    ca, stats, st = F.int8_vectorwise_quant(dev(a), threshold)
    ca_ref, stats_ref, cols_ref = oracle.int8_vectorwise_quant(a, threshold)
    np.testing.assert_array_equal(stats.cpu().numpy(), stats_ref)
    np.testing.assert_array_equal(ca.cpu().numpy(), ca_ref)
    if threshold > 0:
        n = int(st.n_outliers.item())
        got = st.outlier_cols[:n].cpu().numpy()
        np.testing.assert_array_equal(got, cols_ref)
        assert int(st.col_flags.sum().item()) == 0      # self-cleaning
    else:
        assert st is None


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (7, 48, 64), (300, 200, 384), (1, 512, 512), (64, 1536, 384),
                                   (257, 130, 2048)])
@pytest.mark.parametrize("use_bias", [True, False])
def test_linear8bitlt_no_outliers_bit_exact(M, N, K, use_bias):
    rng = np.random.RandomState(M + N + K)
    W = (rng.randn(N, K) * 0.05).astype(np.float16)
    W[rng.rand(N, K) < 0.5] = 0
    x = rng.randn(M, K).astype(np.float16)
    bias = (rng.randn(N) * 0.1).astype(np.float16) if use_bias else None
    CB, SCB, _ = oracle.int8_vectorwise_quant(W, 0.0)
    cb, scb, _ = F.int8_vectorwise_quant(dev(W), 0.0)
    np.testing.assert_array_equal(cb.cpu().numpy(), CB)
    y = F.linear8bitlt(dev(x), cb, scb, None if bias is None else dev(bias), 6.0)
    y_ref, extra = oracle.linear8bitlt_forward(x, CB, SCB, bias, 6.0)
    assert extra is None
    np.testing.assert_array_equal(y.cpu().numpy(), y_ref)


@pytest.mark.parametrize("M,N,K", [(40, 96, 128), (130, 200, 384)])
def test_linear8bitlt_with_outliers(M, N, K):
    rng = np.random.RandomState(M * N)
    W = (rng.randn(N, K) * 0.05).astype(np.float16)
    x = rng.randn(M, K).astype(np.float16)
    for (r, c, v) in [(1, 3, 9.0), (M - 1, K - 2, -12.5), (M // 2, 3, 6.0), (0, K // 2, 30.0)]:
        x[r, c] = v
    bias = (rng.randn(N) * 0.1).astype(np.float16)
    CB, SCB, _ = oracle.int8_vectorwise_quant(W, 0.0)
    y = F.linear8bitlt(dev(x), dev(CB), dev(SCB), dev(bias), 6.0).cpu().numpy()
    y_ref, extra = oracle.linear8bitlt_forward(x, CB, SCB, bias, 6.0)
    assert extra is not None
    # fp16 addmm: one fp16 rounding of (y_int8 + outlier term); accumulation order differs
    scale = np.maximum(np.abs(y_ref.astype(np.float32)), 1.0)
    assert np.max(np.abs(y.astype(np.float32) - y_ref.astype(np.float32)) / scale) <= 2 ** -9
    # rows/cols without outlier contribution are untouched and bit-exact
    untouched = np.abs(extra) == 0
    np.testing.assert_array_equal(y[untouched], y_ref[untouched])


# ------------------------------------------------------------------------------------------------
# quanto qint8
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("N,K", [(48, 64), (130, 384), (5, 50)])
def test_quanto_qint8_bit_exact(dtype, N, K):
    rng = np.random.RandomState(N * K)
    w = (rng.randn(N, K) * 0.02).astype(np.float32)
    w[rng.rand(N, K) < 0.5] = 0
    w[1] = 0
    wt = dev(w, dtype)
    q, scale = F.quanto_quantize_qint8(wt)
    q_ref, s_ref = oracle.quanto_qint8(wt.float().cpu().numpy(), np.float16 if dtype == torch.float16 else np.float32)
    np.testing.assert_array_equal(scale.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(q.cpu().numpy(), q_ref)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (7, 48, 64), (300, 200, 384), (1, 512, 512), (130, 51, 1280)])
def test_qlinear_w8a16(dtype, tol, M, N, K):
    """tolerance: the fp32-accumulated result is rounded once to the activation dtype, so the
    error bound is one half-ulp of the output (2^-11 fp16 / 2^-8 bf16) relative to max(|y|, 1),
    plus fp32 accumulation noise; tol leaves 4x headroom."""
    rng = np.random.RandomState(M + N + K)
    w = (rng.randn(N, K) * 0.05).astype(np.float32)
    w[rng.rand(N, K) < 0.5] = 0
    x = dev(rng.randn(M, K).astype(np.float32), dtype)
    bias = (rng.randn(N) * 0.1).astype(np.float32)
    q_ref, s_ref = oracle.quanto_qint8(w)
    q, scale = F.quanto_quantize_qint8(dev(w))
    y = F.gemm_w8a16(x, q, scale, dev(bias))
    assert y.dtype == dtype
    y_ref = oracle.qlinear_forward(x.float().cpu().numpy(), q_ref, s_ref, bias)
    err = np.abs(y.float().cpu().numpy() - y_ref) / np.maximum(np.abs(y_ref), 1.0)
    assert err.max() <= tol
    y32 = F.gemm_w8a16(x, q, scale, dev(bias), out_dtype=torch.float32).cpu().numpy()
    err32 = np.abs(y32 - y_ref) / np.maximum(np.abs(y_ref), 1.0)
    assert err32.max() <= 2e-5


# ------------------------------------------------------------------------------------------------
# NF4 W4A16
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (7, 48, 128), (300, 200, 384), (1, 768, 768), (130, 51, 1280)])
def test_linear4bit_w4a16(dtype, tol, quant_type, M, N, K):
    rng = np.random.RandomState(M + N + K)
    w = (rng.randn(N, K) * 0.05).astype(np.float32)
    w[rng.rand(N, K) < 0.5] = 0
    np_dt = np.float16 if dtype == torch.float16 else np.float32
    wt = dev(w, dtype)
    x = dev(rng.randn(M, K).astype(np.float32), dtype)
    bias = (rng.randn(N) * 0.1).astype(np.float32)
    packed, absmax = F.quantize_4bit(wt, 64, quant_type)
    p_ref, a_ref = oracle.quantize_4bit(wt.float().cpu().numpy(), 64, quant_type)
    np.testing.assert_array_equal(packed.cpu().numpy(), p_ref)
    y32 = F.gemm_w4a16(x, packed, absmax, N, K, dev(bias), quant_type, out_dtype=torch.float32).cpu().numpy()
    # oracle: dequantised weight rounded to the compute dtype (as bnb does), exact product
    wd = F.dequantize_4bit(packed, absmax, (N, K), 64, quant_type, dtype).float().cpu().numpy().astype(np.float64)
    if dtype == torch.float16:
        wd_ref = oracle.dequantize_4bit(p_ref, a_ref, (N, K), 64, quant_type, np_dt).astype(np.float64)
        np.testing.assert_array_equal(wd, wd_ref)
    y_ref = x.float().cpu().numpy().astype(np.float64) @ wd.T + bias.astype(np.float64)[None, :]
    err32 = np.abs(y32 - y_ref) / np.maximum(np.abs(y_ref), 1.0)
    assert err32.max() <= 2e-5
    y = F.gemm_w4a16(x, packed, absmax, N, K, dev(bias), quant_type)
    err = np.abs(y.float().cpu().numpy() - y_ref) / np.maximum(np.abs(y_ref), 1.0)
    assert err.max() <= tol


# ------------------------------------------------------------------------------------------------
# torch dynamic int8 GPU twin
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_dynamic_twin_against_live_torch_golden(golden_dir, tag):
    import os
    g = np.load(os.path.join(golden_dir, f"torch_dynamic_{tag}.npz"))
    q, scale, wsum = F.torch_quantize_weight(dev(g["w"]))
    np.testing.assert_array_equal(q.cpu().numpy(), g["w_int"])
    assert np.float32(scale.item()) == g["w_scale"]
    np.testing.assert_array_equal(wsum.cpu().numpy(), g["w_int"].astype(np.int32).sum(1))
    xq, qparams = F.torch_quantize_activation(dev(g["x"]))
    np.testing.assert_array_equal(xq.cpu().numpy(), g["x_int"])
    qp = qparams.cpu().numpy()
    assert np.float32(qp[0]) == g["x_scale"] and int(qp[1]) == int(g["x_zp"])
    y = F.gemm_dyn_i8(xq, qparams, q, scale, wsum, dev(g["bias"])).cpu().numpy()
    y_orc = oracle.torch_dynamic_linear(g["x"], g["w_int"], float(g["w_scale"]), g["bias"])
    np.testing.assert_array_equal(y, y_orc)                       # same fp32 formula: bit-exact
    np.testing.assert_allclose(y, g["y"], rtol=2e-6, atol=2e-6)   # live torch (FBGEMM rounding order)


@pytest.mark.parametrize("M,N,K", [(300, 200, 384), (1, 512, 512), (12, 51, 1536)])
def test_torch_dynamic_twin_vs_oracle(M, N, K):
    rng = np.random.RandomState(M + N + K)
    w = (rng.randn(N, K) * 0.02).astype(np.float32)
    x = (rng.randn(M, K) * 2).astype(np.float32)
    bias = (rng.randn(N) * 0.1).astype(np.float32)
    q, scale, wsum = F.torch_quantize_weight(dev(w))
    q_ref, s_ref = oracle.torch_weight_qint8(w)
    np.testing.assert_array_equal(q.cpu().numpy(), q_ref)
    xq, qparams = F.torch_quantize_activation(dev(x))
    y = F.gemm_dyn_i8(xq, qparams, q, scale, wsum, dev(bias)).cpu().numpy()
    y_ref = oracle.torch_dynamic_linear(x, q_ref, s_ref, bias)
    np.testing.assert_array_equal(y, y_ref)


# ------------------------------------------------------------------------------------------------
# log-mel and tallies
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mels", [80, 128])
def test_logmel_2s_golden(golden_dir, mels):
    """tolerance 1e-4 abs on the (x+4)/4-scaled log10 features: fp32 FFT/mel rounding (HF itself
    claims 1e-5 between its numpy and torch paths)."""
    import os
    g = np.load(os.path.join(golden_dir, f"logmel_2s_{mels}.npz"))
    L = 40000
    audio = np.zeros((3, L), dtype=np.float32)
    for i, (n, seed) in enumerate(zip(g["lengths"], g["seeds"])):
        audio[i, :n] = synth_audio(int(seed), int(n))
    filters = dev(oracle.mel_filter_bank_slaney(mels).astype(np.float32))
    lengths = torch.tensor([int(x) for x in g["lengths"]], dtype=torch.int32, device="cuda")
    out = F.log_mel(dev(audio), filters, n_samples=32000, lengths=lengths).cpu().numpy()
    assert out.shape == (3, mels, 200)
    np.testing.assert_allclose(out, g["feats"], rtol=0, atol=1e-4)


def test_logmel_30s_golden_and_oracle(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "logmel_30s_80.npz"))
    audio = np.stack([synth_audio(0), synth_audio(1)])
    filters = dev(oracle.mel_filter_bank_slaney(80).astype(np.float32))
    out = F.log_mel(dev(audio), filters).cpu().numpy()
    assert out.shape == (2, 80, 3000)
    np.testing.assert_allclose(out[0][:, g["frames"]], g["feats"], rtol=0, atol=1e-4)
    ref = oracle.log_mel_spectrogram(audio[1], 80)[0]
    np.testing.assert_allclose(out[1], ref, rtol=0, atol=1e-4)


def test_edit_distance_matches_oracle():
    rng = np.random.RandomState(5)
    refs, hyps = [], []
    for n in (0, 1, 5, 60, 400, 1000):
        r = rng.randint(0, 30, size=n).astype(np.int32)
        h = r.copy()
        if n:
            h = np.delete(h, rng.randint(0, n, size=n // 7))
            h = np.insert(h, rng.randint(0, len(h) + 1, size=n // 5), 99)
        refs.append(r)
        hyps.append(h.astype(np.int32))
    refs.append(np.array([1, 2, 3], np.int32)); hyps.append(np.zeros((0,), np.int32))
    ro = np.cumsum([0] + [len(r) for r in refs]).astype(np.int64)
    ho = np.cumsum([0] + [len(h) for h in hyps]).astype(np.int64)
    d = F.edit_distance(dev(np.concatenate(refs)), dev(ro), dev(np.concatenate(hyps)), dev(ho)).cpu().numpy()
    want = [oracle.edit_distance(r, h) for r, h in zip(refs, hyps)]
    np.testing.assert_array_equal(d, want)


@pytest.mark.parametrize("M,N,K", [(1, 512, 512), (17, 200, 384), (33, 2048, 512), (64, 512, 2048), (64, 51, 64)])
@pytest.mark.parametrize("outliers", [False, True])
def test_llmint8_small_m_kernel_bit_identical_to_tensor_core_path(M, N, K, outliers):
    """decode-shaped calls take the fused single-launch kernel (csrc/gemv_small.cu); it must give
    exactly the bits of quantize + tcgen05 GEMM, and of the oracle when there are no outliers."""
    rng = np.random.RandomState(M * 7 + N + K)
    W = (rng.randn(N, K) * 0.05).astype(np.float16)
    x = rng.randn(M, K).astype(np.float16)
    if outliers:
        x[0, 3] = 8.0
        x[M - 1, K - 1] = -6.0
        x[M // 2, K // 2] = 40.0
    bias = (rng.randn(N) * 0.1).astype(np.float16)
    CB, SCB, _ = oracle.int8_vectorwise_quant(W, 0.0)
    cb, scb, b = dev(CB), dev(SCB), dev(bias)
    saved = F.SMALL_M_ROWS
    try:
        F.SMALL_M_ROWS = 64
        y_small = F.linear8bitlt(dev(x), cb, scb, b, 6.0).cpu().numpy()
        F.SMALL_M_ROWS = 0
        y_tc = F.linear8bitlt(dev(x), cb, scb, b, 6.0).cpu().numpy()
    finally:
        F.SMALL_M_ROWS = saved
    np.testing.assert_array_equal(y_small, y_tc)
    y_ref, extra = oracle.linear8bitlt_forward(x, CB, SCB, bias, 6.0)
    if not outliers:
        assert extra is None
        np.testing.assert_array_equal(y_small, y_ref)
    else:
        assert np.abs(y_small.astype(np.float32) - y_ref.astype(np.float32)).max() <= 2 ** -7


# ------------------------------------------------------------------------------------------------
# quanto qint4 (group-wise affine uint4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("N,K", [(48, 64), (130, 384), (33, 1280), (5, 160)])
def test_quanto_qint2_codes_bit_exact_and_gemm(dtype, N, K):
    """qint2 (quantization/evaluation_scripts/dynamic_evaluation_int2.py:158-160): codes 0..3 in the qint4
    container, bit-exact vs the oracle; the fused GEMM consumes them unchanged."""
    rng = np.random.RandomState(N * 3 + K)
    w = (rng.randn(N, K) * 0.05).astype(np.float32)
    wt = dev(w).to(dtype)
    packed, scale, shift, g = F.quanto_quantize_qint4(wt, bits=2)
    q_ref, s_ref, sh_ref, g_ref = oracle.quanto_qint4(wt.float().cpu().numpy(), bits=2)
    assert g == g_ref and int(q_ref.max()) <= 3
    np.testing.assert_array_equal(scale.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(shift.cpu().numpy(), sh_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), oracle.quanto_qint4_pack(q_ref))
    if K % 64 == 0:
        x = dev((rng.randn(9, K) * 0.5).astype(np.float32)).half()
        y = F.gemm_u4a16(x, packed, scale, shift, g, None, out_dtype=torch.float32).cpu().numpy()
        wd = torch.from_numpy(oracle.quanto_qint4_dequant(q_ref, s_ref, sh_ref, g)).half().double().numpy()
        ref = x.double().cpu().numpy() @ wd.T
        assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("N,K", [(48, 64), (130, 384), (33, 1280), (5, 160)])
def test_quanto_qint4_codes_bit_exact(dtype, N, K):
    rng = np.random.RandomState(N * K)
    w = (rng.randn(N, K) * 0.02).astype(np.float32)
    w[1, :] = 0.25                       # constant group: scale 0 -> codes 0
    wt = dev(w, dtype)
    packed, scale, shift, g = F.quanto_quantize_qint4(wt)
    q_ref, s_ref, sh_ref, g_ref = oracle.quanto_qint4(wt.float().cpu().numpy())
    assert g == g_ref == oracle.quanto_group_size(K)
    np.testing.assert_array_equal(scale.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(shift.cpu().numpy(), sh_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), oracle.quanto_qint4_pack(q_ref))


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (7, 48, 128), (300, 200, 384), (1, 768, 768), (130, 51, 1280)])
def test_qlinear_qint4_w4a16(dtype, tol, M, N, K):
    rng = np.random.RandomState(M + N + K)
    w = (rng.randn(N, K) * 0.05).astype(np.float32)
    x = dev(rng.randn(M, K).astype(np.float32), dtype)
    bias = (rng.randn(N) * 0.1).astype(np.float32)
    packed, scale, shift, g = F.quanto_quantize_qint4(dev(w))
    q_ref, s_ref, sh_ref, _ = oracle.quanto_qint4(w)
    # the kernel rounds the dequantised weight once to the operand dtype
    wd = torch.from_numpy(oracle.quanto_qint4_dequant(q_ref, s_ref, sh_ref, g)).to(dtype).double().numpy()
    y_ref = x.double().cpu().numpy() @ wd.T + bias.astype(np.float64)[None, :]
    y32 = F.gemm_u4a16(x, packed, scale, shift, g, dev(bias), out_dtype=torch.float32).cpu().numpy()
    assert (np.abs(y32 - y_ref) / np.maximum(np.abs(y_ref), 1.0)).max() <= 2e-5
    y = F.gemm_u4a16(x, packed, scale, shift, g, dev(bias))
    assert y.dtype == dtype
    assert (np.abs(y.float().cpu().numpy() - y_ref) / np.maximum(np.abs(y_ref), 1.0)).max() <= tol


# ------------------------------------------------------------------------------------------------
# bitsandbytes nested ("double") quantization of the 4-bit statistics
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [256, 1000, 12288, 7])
def test_absmax_double_quant_bit_exact(n):
    rng = np.random.RandomState(n)
    absmax = (np.abs(rng.randn(n)) * 0.03 + 0.01).astype(np.float32)
    if n > 300:
        absmax[256:300] = 0.0                 # all-zero weight blocks (pruned)
    q, a2, off, deq = F.quantize_absmax_double(dev(absmax))
    q_ref, a2_ref, off_ref, deq_ref = oracle.quantize_absmax_double(absmax)
    np.testing.assert_array_equal(F.dynamic_map("cuda").cpu().numpy(), oracle.dynamic_map())
    assert np.float32(off.item()) == off_ref
    np.testing.assert_array_equal(a2.cpu().numpy(), a2_ref)
    np.testing.assert_array_equal(q.cpu().numpy(), q_ref)
    np.testing.assert_array_equal(deq.cpu().numpy(), deq_ref)
    back = F.dequantize_absmax_double(q, a2, off)
    np.testing.assert_array_equal(back.cpu().numpy(), deq_ref)
    # 8-bit dynamic map: relative error of the reconstructed statistics stays below 1 % of the block max
    assert np.abs(deq_ref - absmax).max() <= 0.01 * np.abs(absmax - off_ref).max() + 1e-7


def test_linear4bit_double_quant_module():
    """bnb_nf4_16_double flow (model_utils.py:36-48): Linear4bit(compress_statistics=True)."""
    from openai_whisper_compression_b200 import bnb
    torch.manual_seed(3)
    lin = torch.nn.Linear(256, 192).half()
    m = bnb.Linear4bit(256, 192, bias=True, compute_dtype=torch.float16, compress_statistics=True, quant_type="nf4")
    m.load_state_dict(lin.state_dict(), strict=False)
    m = m.to("cuda")
    qs = m.weight.quant_state
    assert qs.nested and qs.absmax.dtype == torch.uint8 and qs.state2.blocksize == 256
    p_ref, a_ref = oracle.quantize_4bit(lin.weight.detach().float().numpy(), 64, "nf4")
    q_ref, a2_ref, off_ref, deq_ref = oracle.quantize_absmax_double(a_ref)
    np.testing.assert_array_equal(m.weight.data.cpu().numpy(), p_ref)
    np.testing.assert_array_equal(qs.absmax.cpu().numpy(), q_ref)
    wd = bnb.dequantize_4bit(m.weight.data, qs).cpu().numpy()
    wd_ref = oracle.dequantize_4bit(p_ref, deq_ref, (192, 256), 64, "nf4", np.float16)
    np.testing.assert_array_equal(wd, wd_ref)
    sd = m.state_dict()
    assert {"weight.absmax", "weight.nested_absmax", "weight.nested_quant_map", "weight.quant_map"} <= set(sd)
    assert sd["weight.absmax"].dtype == torch.uint8
    x = torch.randn(9, 256).half()
    y = m(x.cuda()).float().cpu().numpy()
    y_ref = x.double().numpy() @ wd_ref.astype(np.float64).T + lin.bias.detach().double().numpy()
    assert np.abs(y - y_ref).max() < 4e-3


# ------------------------------------------------------------------------------------------------
# round 2: bitsandbytes' approximate row scale, unquantized projection with fused arg-max
# ------------------------------------------------------------------------------------------------
def test_int8_row_scale_is_the_approximate_division_of_bitsandbytes():
    """The fp16 (absmax, a) pairs on which __fdividef(127, absmax) and the IEEE quotient give different int8 codes
    (exhaustive B200 sweep, profiles/r02_bnb_open_points.json): the quantizer kernel must produce the approximate
    form's code (what bitsandbytes' kInt8VectorQuant computes), and so must the oracle through its table."""
    import json
    import os
    rep = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                      "r02_bnb_open_points.json")))
    recs = rep["first_mismatches"]
    rows = np.zeros((len(recs), 16), dtype=np.uint16)
    for i, r in enumerate(recs):
        rows[i, 0], rows[i, 1] = r["absmax_bits"], r["a_bits"]
    a = rows.view(np.float16)
    ca, stats, _ = F.int8_vectorwise_quant(dev(a), 0.0)
    got = ca.cpu().numpy()
    np.testing.assert_array_equal(got[:, 1], np.array([r["q_fdividef"] for r in recs], dtype=np.int8))
    assert not np.array_equal(got[:, 1], np.array([r["q_ieee"] for r in recs], dtype=np.int8))
    CA, SCA, _ = oracle.int8_vectorwise_quant(a, 0.0)
    np.testing.assert_array_equal(got, CA)
    np.testing.assert_array_equal(stats.cpu().numpy(), SCA)
    # every fp16 absmax: the kernel's scale is the committed table (row = [absmax, absmax/2 rounded], code differs
    # wherever the table entry is used)
    bits = np.arange(1, 0x7c00, dtype=np.uint16)
    full = np.zeros((bits.size, 16), dtype=np.uint16)
    full[:, 0] = bits
    full[:, 1] = (bits.view(np.float16).astype(np.float32) * np.float32(0.5)).astype(np.float16).view(np.uint16)
    full[:, 2] = (bits.view(np.float16).astype(np.float32) * np.float32(0.75)).astype(np.float16).view(np.uint16)
    af = full.view(np.float16)
    ca2, _, _ = F.int8_vectorwise_quant(dev(af), 0.0)
    CA2, _, _ = oracle.int8_vectorwise_quant(af, 0.0)
    np.testing.assert_array_equal(ca2.cpu().numpy(), CA2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(256, 51865, 512), (6, 51866, 1280), (300, 1000, 384), (1, 51865, 384), (64, 130, 64)])
def test_gemm_f16_projection_and_fused_argmax(dtype, M, N, K):
    """Unquantized proj_out on the tcgen05 pipeline: logits within one output rounding of an fp32 reference, and
    the arg-max folded into the epilogue == torch.argmax(masked logits) of the logits the SAME kernel stores
    (bit-exact: same rounded values, first index among ties), with and without storing the logits."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dtype)
    Vp = -(-N // 256) * 256
    out = torch.zeros((M, -(-N // 8) * 8), dtype=dtype, device="cuda")
    mask = torch.zeros((Vp,), dtype=torch.bool, device="cuda")
    mask[:N] = torch.rand(N, device="cuda", generator=g) < 0.3
    keys = torch.zeros((M,), dtype=torch.int64, device="cuda")
    y = F.gemm_f16(x, w, None, out=out, argmax_keys=keys, mask=mask)
    ref = x.float() @ w.float().t()
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert (y[:, :N].float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    tok = F.argmax_finalize(keys)
    assert int(keys.abs().sum()) == 0                                   # keys reset for the next call
    want = torch.argmax(y[:, :N].float().masked_fill(mask[:N], float("-inf")), dim=-1)
    assert torch.equal(tok, want)
    # ties and a fully masked row
    x2 = x.clone()
    x2[0] = 0
    y2 = F.gemm_f16(x2, w, None, out=out, argmax_keys=keys, mask=mask)
    tok2 = F.argmax_finalize(keys)
    assert torch.equal(tok2, torch.argmax(y2[:, :N].float().masked_fill(mask[:N], float("-inf")), dim=-1))
    # the same without storing the logits, no mask, with a bias
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    F.gemm_f16(x, w, bias, argmax_keys=keys, store=False)
    tok3 = F.argmax_finalize(keys)
    y3 = F.gemm_f16(x, w, bias)
    assert y3.shape == (M, N)
    assert torch.equal(tok3, torch.argmax(y3.float(), dim=-1))
    assert (y3.float() - (ref + bias)).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------------------
# round 2: optimum-quanto float8 weights and static activation quantization
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K", [(48, 64), (130, 384), (33, 1280)])
def test_quanto_qfloat8_codes_bit_exact_and_gemm(dtype, N, K):
    """weights=qfloat8: codes / scales equal torch's own `(w / (absmax / 448)).to(float8_e4m3fn)` in the weight dtype;
    the fused GEMM equals dequantize-then-matmul (fp32 accumulate, one output rounding)."""
    rng = np.random.RandomState(N + K)
    w = (rng.randn(N, K) * 0.02).astype(np.float32)
    w[rng.rand(N, K) < 0.4] = 0
    w[1] = 0
    wt = dev(w, dtype)
    q, scale = F.quanto_quantize_qfloat8(wt)
    q_ref, s_ref = oracle.quanto_qfloat8(wt.float().cpu().numpy(), str(dtype).replace("torch.", ""))
    np.testing.assert_array_equal(scale.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(q.cpu().numpy(), q_ref)
    assert np.all(q.cpu().numpy()[w == 0] & 0x7f == 0)                    # pruned zeros stay (signed) zeros
    if dtype == torch.float32:
        return
    x = torch.randn(70, K, device="cuda").to(dtype)
    bias = torch.randn(N, device="cuda") * 0.1
    y = F.gemm_wf8a16(x, q, scale.view(-1), bias)
    wd = q.view(torch.float8_e4m3fn).float()
    ref = (x.float() @ wd.t()) * scale.view(1, -1) + bias
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert (y.float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("qtype", ["qint8", "qfloat8"])
def test_static_activation_quantization_bit_exact(dtype, qtype):
    """quantize_activation with a calibrated per-tensor scale: code values and dequantized tensor equal torch's own
    ops in the same dtype (CPU), incl. saturation and values on rounding ties."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(37, 96, generator=g) * 3).to(dtype)
    x[0, :8] = torch.tensor([0.0, 1e4, -1e4, 0.5, 1.5, 2.5, -0.5, 127.5]).to(dtype)
    dn = str(dtype).replace("torch.", "")
    sc = torch.tensor(0.0625 if qtype == "qint8" else 0.01, dtype=dtype)
    g_ref, d_ref = oracle.quanto_quantize_activation(x.float().numpy(), float(sc), qtype, dn)
    codes, grid, deq = F.quant_act_static(x.cuda(), sc.float().cuda().view(1), qtype, codes=qtype == "qint8", grid=True,
                                          deq=True)
    np.testing.assert_array_equal(grid.float().cpu().numpy(), g_ref)
    np.testing.assert_array_equal(deq.float().cpu().numpy(), d_ref)
    if codes is not None:
        np.testing.assert_array_equal(codes.cpu().numpy().astype(np.float32), g_ref)


@pytest.mark.parametrize("out_dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(300, 200, 384), (5, 512, 512), (1500, 1536, 512)])
def test_gemm_w8a8_exact_vs_integer_matmul(out_dtype, M, N, K):
    """qint8 x qint8 (quanto qbytes_int_mm): float(int32 acc) * out_scale[n] + bias[n] with exact integer sums."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    xq = torch.randint(-128, 128, (M, K), device="cuda", generator=g, dtype=torch.int8)
    wq = torch.randint(-128, 128, (N, K), device="cuda", generator=g, dtype=torch.int8)
    os_ = torch.rand(N, device="cuda", generator=g) * 1e-3
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    y = F.gemm_w8a8(xq, wq, os_, bias, out_dtype)
    acc = (xq.double() @ wq.double().t())
    ref = ((acc.float() * os_[None, :]) + bias[None, :]).to(out_dtype)
    assert torch.equal(y, ref)


# ------------------------------------------------------------------------------------------------
# round 2: decode-shaped forward (<= 32 rows) of the weight-only schemes on the CUDA-core GEMV (gemv_wq.cu)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("M,N,K", [(1, 768, 768), (16, 768, 3072), (32, 3072, 768), (7, 130, 384), (32, 51, 1280), (9, 2304, 768)])
@pytest.mark.parametrize("mode", ["nf4", "fp4", "w8", "u4", "u2", "f8"])
def test_weight_only_gemv_matches_dequant_matmul_and_the_tensor_core_gemm(dtype, tol, M, N, K, mode):
    """Same contract as the tcgen05 GEMMs it stands in for at <= 32 rows: y = x @ dequant(W)^T (* scale) + bias with
    the dequantized weight rounded once to the activation dtype where the scheme rounds, fp32 accumulation, one output
    rounding.  Checked against a float64 product of exactly those operand values, and against the tensor-core kernel
    (WQ_GEMV_ROWS = 0 path) on the same inputs."""
    g = torch.Generator(device="cuda").manual_seed(M * 31 + N + K)
    x = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05)
    W[torch.rand(N, K, device="cuda", generator=g) < 0.3] = 0
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    if mode in ("nf4", "fp4"):
        packed, absmax = F.quantize_4bit(W.to(dtype), 64, mode)
        wd = F.dequantize_4bit(packed, absmax, (N, K), 64, mode, dtype).double()
        call = lambda: F.gemm_w4a16(x, packed, absmax, N, K, bias, mode)
        post = None
    elif mode == "w8":
        q, scale = F.quanto_quantize_qint8(W.to(dtype))
        wd, post = q.double(), scale.view(1, -1).double()
        call = lambda: F.gemm_w8a16(x, q, scale.view(-1), bias)
    elif mode == "f8":
        q, scale = F.quanto_quantize_qfloat8(W.to(dtype))
        wd, post = q.view(torch.float8_e4m3fn).double(), scale.view(1, -1).double()
        call = lambda: F.gemm_wf8a16(x, q, scale.view(-1), bias)
    else:
        bits = 4 if mode == "u4" else 2
        packed, scale, shift, grp = F.quanto_quantize_qint4(W.to(dtype), bits=bits)
        hi, lo = (packed >> 4).float(), (packed & 15).float()
        codes = torch.stack([hi, lo], -1).reshape(N, K)
        wd = (scale.repeat_interleave(grp, 1) * codes - shift.repeat_interleave(grp, 1)).to(dtype).double()
        call = lambda: F.gemm_u4a16(x, packed, scale, shift, grp, bias)
        post = None
    assert F.GEMV_ROWS >= 32
    y = call()
    ref = x.double() @ wd.t()
    if post is not None:
        ref = ref * post
    ref = ref + bias.double()
    scale_ref = max(1.0, ref.abs().max().item())
    assert y.dtype == dtype and y.shape == (M, N)
    assert (y.double() - ref).abs().max().item() <= tol * scale_ref
    old = F.GEMV_ROWS
    F.GEMV_ROWS = 0
    try:
        y_tc = call()
    finally:
        F.GEMV_ROWS = old
    assert (y.double() - y_tc.double()).abs().max().item() <= tol * scale_ref
    if mode in ("nf4", "w8"):       # pruned weights: zero rows stay exactly bias
        Wz = torch.zeros_like(W)
        if mode == "nf4":
            pz, az = F.quantize_4bit(Wz.to(dtype), 64, "nf4")
            yz = F.gemm_w4a16(x, pz, az, N, K, bias, "nf4")
        else:
            qz, sz = F.quanto_quantize_qint8(Wz.to(dtype))
            yz = F.gemm_w8a16(x, qz, sz.view(-1), bias)
        assert torch.equal(yz, bias.to(dtype).expand(M, N))


@pytest.mark.parametrize("M,N,K", [(16, 1024, 1024), (32, 4096, 1024), (16, 1024, 4096), (7, 130, 384), (32, 51, 1280),
                                   (48, 768, 768)])
@pytest.mark.parametrize("mode", ["nf4", "w8", "u4", "f8"])
def test_weight_only_gemv_fp32_flow(M, N, K, mode):
    """The reference's fp32 flows (quanto / bnb compute_dtype fp32 on a model that was never .half()-ed,
    model_utils.py:139-142) at decode shapes: fp32 rows in, rounded to fp16 as the tensor-core GEMM of the same flow
    rounds them, fp32 sums out; up to 32 rows (more rows: cast pass + tensor-core GEMM, 48-row case).  Against a float64 product of exactly those operands
    (2e-5 of max(|y|, 1): fp32 accumulation) and against the tensor-core path (cast pass + GEMM) on the same inputs."""
    g = torch.Generator(device="cuda").manual_seed(M * 31 + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    W[torch.rand(N, K, device="cuda", generator=g) < 0.3] = 0
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    post = None
    if mode == "nf4":
        packed, absmax = F.quantize_4bit(W, 64, "nf4")
        wd = F.dequantize_4bit(packed, absmax, (N, K), 64, "nf4", torch.float16).double()
        call = lambda: F.gemm_w4a16(x, packed, absmax, N, K, bias, "nf4", out_dtype=torch.float32)
    elif mode == "w8":
        q, scale = F.quanto_quantize_qint8(W)
        wd, post = q.double(), scale.view(1, -1).double()
        call = lambda: F.gemm_w8a16(x, q, scale.view(-1), bias, torch.float32)
    elif mode == "f8":
        q, scale = F.quanto_quantize_qfloat8(W)
        wd, post = q.view(torch.float8_e4m3fn).float().double(), scale.view(1, -1).double()
        call = lambda: F.gemm_wf8a16(x, q, scale.view(-1), bias, torch.float32)
    else:
        packed, scale, shift, grp = F.quanto_quantize_qint4(W)
        codes = torch.stack([(packed >> 4).float(), (packed & 15).float()], -1).reshape(N, K)
        wd = (scale.repeat_interleave(grp, 1) * codes - shift.repeat_interleave(grp, 1)).half().double()
        call = lambda: F.gemm_u4a16(x, packed, scale, shift, grp, bias, torch.float32)
    before = F.STATS.launches
    y = call()
    assert F.STATS.launches - before == 1
    assert y.dtype == torch.float32 and y.shape == (M, N)
    ref = x.half().double() @ wd.t()
    if post is not None:
        ref = ref * post
    ref = ref + bias.double()
    assert ((y.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item() <= 5e-5
    old = F.GEMV_ROWS
    F.GEMV_ROWS = 0
    try:
        y_tc = call()
    finally:
        F.GEMV_ROWS = old
    assert ((y - y_tc).abs() / ref.abs().clamp_min(1.0).float()).max().item() <= 5e-5

"""Fused producers (rowops.cu, attn_decode.cu) through the C ABI.

Bars: the residual add and the int8 row quantization are bit-exact (against torch's fp16 add and against the
stand-alone quantizer -- itself oracle-checked in test_gpu_kernels.py -- run on the tensor the fused kernel
wrote); LayerNorm / GELU / attention are floating-point producers compared with torch's own ops: at most one
fp16 ulp apart for LayerNorm and GELU, 2e-3 absolute for attention (fp32 softmax here, fp16 P in torch's
flash kernels).
"""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

F = None


@pytest.fixture(scope="module", autouse=True)
def _load():
    global F
    from openai_whisper_compression_b200 import functional
    F = functional
    yield


def _ulp16(ref: torch.Tensor) -> torch.Tensor:
    """Size of one fp16 ulp at |ref| (normal range), as fp32."""
    a = ref.float().abs().clamp_min(2.0 ** -14)
    return torch.exp2(torch.floor(torch.log2(a)) - 10)


def _standalone_quant(h: torch.Tensor, threshold: float):
    """(ca, sca, flags) from the stand-alone quantizer on `h`; leaves the shared flag buffer zeroed."""
    ca, sca, st = F.int8_vectorwise_quant(h, threshold, finalize=False)
    flags = None
    if st is not None:
        flags = st.col_flags[: h.shape[-1] + 1].clone()
        st.col_flags.zero_()
    return ca, sca, flags


def _take_flags(quant, cols):
    st = quant[2]
    if st is None:
        return None
    flags = st.col_flags[: cols + 1].clone()
    st.col_flags.zero_()
    return flags


def _check_quant(quant, h, threshold):
    cols = h.shape[-1]
    got_flags = _take_flags(quant, cols)
    ca, sca, flags = _standalone_quant(h.reshape(-1, cols), threshold)
    assert torch.equal(quant[0], ca.reshape(quant[0].shape))
    assert torch.equal(quant[1], sca)
    if threshold > 0:
        assert torch.equal(got_flags, flags)
    else:
        assert got_flags is None


@pytest.mark.parametrize("rows,cols", [(1, 384), (37, 512), (300, 768), (64, 1280), (9, 2048), (5, 8)])
@pytest.mark.parametrize("with_delta", [True, False])
@pytest.mark.parametrize("threshold", [None, 0.0, 6.0])
def test_add_layernorm_quant(rows, cols, with_delta, threshold):
    g = torch.Generator(device="cuda").manual_seed(rows * 131 + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 1.5).half()
    delta = (torch.randn(rows, cols, device="cuda", generator=g) * 0.7).half() if with_delta else None
    w = (1.0 + 0.2 * torch.randn(cols, device="cuda", generator=g)).half()
    b = (0.1 * torch.randn(cols, device="cuda", generator=g)).half()
    if threshold:
        w[:: max(1, cols // 5)] *= 8.0          # a few columns cross the outlier threshold
    x_in = x.clone()
    x_new, h, quant = F.add_layernorm_quant(x, delta, w, b, 1e-5, threshold)
    assert torch.equal(x, x_in)
    ref_x = x + delta if with_delta else x
    assert torch.equal(x_new, ref_x)                      # fp16 add, bit-exact
    ref_h = TF.layer_norm(ref_x, (cols,), w, b, 1e-5)
    err = (h.float() - ref_h.float()).abs()
    assert bool((err <= _ulp16(ref_h)).all()), float((err / _ulp16(ref_h)).max())
    assert (h == ref_h).float().mean().item() > 0.98
    if threshold is None:
        assert quant is None
    else:
        if threshold and rows * cols > 4000:
            assert (h.float().abs() >= threshold).any()   # the outlier branch is exercised
        _check_quant(quant, h, threshold)


def test_add_layernorm_bf16_and_errors():
    x = torch.randn(10, 512, device="cuda").bfloat16()
    d = torch.randn(10, 512, device="cuda").bfloat16()
    w = torch.ones(512, device="cuda").bfloat16()
    b = torch.zeros(512, device="cuda").bfloat16()
    x_new, h, quant = F.add_layernorm_quant(x, d, w, b, 1e-5)
    assert quant is None and torch.equal(x_new, x + d)
    ref = TF.layer_norm(x + d, (512,), w, b, 1e-5)
    assert (h.float() - ref.float()).abs().max().item() <= 2.0 ** -6
    with pytest.raises(RuntimeError, match="fp16"):
        F.add_layernorm_quant(x, d, w, b, 1e-5, threshold=6.0)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        F.add_layernorm_quant(torch.randn(2, 2056, device="cuda").half(), None,
                              torch.ones(2056, device="cuda").half(), torch.zeros(2056, device="cuda").half(), 1e-5)
    with pytest.raises(RuntimeError, match="CPU"):
        F.add_layernorm_quant(torch.randn(2, 64).half(), None, torch.ones(64).half(), torch.zeros(64).half(), 1e-5)


@pytest.mark.parametrize("rows,cols", [(1, 1536), (50, 2048), (257, 3072), (33, 5120), (3, 8)])
@pytest.mark.parametrize("threshold", [None, 0.0, 6.0])
def test_gelu_quant(rows, cols, threshold):
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 3.0).half()
    h, quant = F.gelu_quant(x, threshold)
    ref = TF.gelu(x)
    err = (h.float() - ref.float()).abs()
    assert bool((err <= _ulp16(ref)).all()), float((err / _ulp16(ref)).max())
    assert (h == ref).float().mean().item() > 0.999
    if threshold is None:
        assert quant is None
    else:
        if threshold and rows * cols > 1000:
            assert (h.float().abs() >= threshold).any()
        _check_quant(quant, h, threshold)


@pytest.mark.parametrize("rows,cols", [(4096, 2048), (5003, 1000), (6000, 384), (4100, 8), (4097, 3072), (4200, 4096),
                                       (4099, 5120), (4100, 2056)])
@pytest.mark.parametrize("threshold", [None, 0.0, 6.0])
def test_gelu_quant_table_path_is_bit_identical_to_erf_path(rows, cols, threshold):
    """Encoder-sized fp16 calls (rows >= 4096, cols <= 5120) take the 65536-entry table kernel; the same rows in
    chunks of 1000 take the erff kernel.  Same values, codes, row statistics and outlier flags, bit for bit --
    including every special input (inf, nan, -0, subnormals)."""
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 3.0).half()
    special = torch.tensor([float("inf"), float("-inf"), float("nan"), -0.0, 6e-8, -6e-8, 65504.0, -65504.0],
                           device="cuda").half()
    x[5, :8] = special
    h, quant = F.gelu_quant(x, threshold)
    flags = None
    if threshold:
        st = quant[2]
        flags = st.col_flags[: cols + 1].clone()
        st.col_flags.zero_()
    parts = [F.gelu_quant(x[r:r + 1000].contiguous(), threshold) for r in range(0, rows, 1000)]
    h_ref = torch.cat([p[0] for p in parts], 0)
    assert torch.equal(h.view(torch.int16), h_ref.view(torch.int16))
    if threshold is None:
        assert quant is None
        return
    assert torch.equal(quant[0], torch.cat([p[1][0] for p in parts], 0))
    assert torch.equal(quant[1].view(torch.int32), torch.cat([p[1][1] for p in parts], 0).view(torch.int32))
    if threshold:
        st = parts[0][1][2]
        assert torch.equal(flags, st.col_flags[: cols + 1])
        st.col_flags.zero_()
        assert int(flags[cols].item()) == 1


@pytest.mark.parametrize("rows,cols,N", [(4200, 2048, 512), (4100, 5120, 1280), (37, 3072, 768), (300, 1536, 384)])
@pytest.mark.parametrize("threshold", [0.0, 6.0])
def test_gelu_quant_without_the_fp16_store(rows, cols, N, threshold):
    """gelu_quant(store_h=False) writes the same int8 rows, row statistics and outlier flags and no fp16 activation;
    fc2 with a_pre_gelu=True on fc1's output then gives, bit for bit, what it gives on the stored activation --
    outlier columns included (table kernel and erff kernel, every table width)."""
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 2.5).half()
    cb = torch.randint(-127, 128, (N, cols), device="cuda", generator=g, dtype=torch.int8)
    scb = torch.rand(N, device="cuda", generator=g) * 0.05 + 0.01
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(rows, N, device="cuda", generator=g).half()
    h, q_ref = F.gelu_quant(x, threshold)
    if threshold:
        assert (h.float().abs() >= threshold).any()
    y_ref = F.gemm_llmint8(q_ref[0], q_ref[1], cb, scb, bias, a_f16=h, state=q_ref[2], residual=res, clamp_abs=64000.0)
    if q_ref[2] is not None:
        assert int(q_ref[2].col_flags[: cols + 2].abs().sum().item()) == 0       # the GEMM cleared the flags
    none, q = F.gelu_quant(x, threshold, store_h=False)
    assert none is None
    assert torch.equal(q[0], q_ref[0]) and torch.equal(q[1].view(torch.int32), q_ref[1].view(torch.int32))
    y = F.gemm_llmint8(q[0], q[1], cb, scb, bias, a_f16=x, state=q[2], residual=res, clamp_abs=64000.0, a_pre_gelu=True)
    assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16))
    with pytest.raises(RuntimeError, match="store_h"):
        F.gelu_quant(x, None, store_h=False)


@pytest.mark.parametrize("B,H,t_max,pos", [(3, 6, 64, 0), (5, 8, 128, 17), (2, 20, 448, 447), (64, 8, 128, 64),
                                           (1, 12, 64, 3)])
@pytest.mark.parametrize("threshold", [None, 6.0])
@pytest.mark.parametrize("fused_qkv", [True, False])
def test_self_attn_decode(B, H, t_max, pos, threshold, fused_qkv):
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H * 10 + pos)
    if fused_qkv:
        qkv = (torch.randn(B, 3 * d, device="cuda", generator=g) * 2.0).half()
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q, k, v = ((torch.randn(B, d, device="cuda", generator=g) * 2.0).half() for _ in range(3))
    if threshold:
        v = v.clone() if not fused_qkv else v
    kc = (torch.randn(B, t_max, d, device="cuda", generator=g)).half()
    vc = (torch.randn(B, t_max, d, device="cuda", generator=g) * (8.0 if threshold else 1.0)).half()
    kc0, vc0 = kc.clone(), vc.clone()
    pos_t = torch.tensor([pos], dtype=torch.int64, device="cuda")
    scaling = 0.125
    out, quant = F.self_attn_decode(q, k, v, scaling, kc, vc, pos_t, H, threshold)
    # cache append: row `pos` replaced, everything else untouched
    kc0[:, pos] = k
    vc0[:, pos] = v
    assert torch.equal(kc, kc0) and torch.equal(vc, vc0)
    # reference: HF's arithmetic in fp32 on the same (rounded) operands
    qs = (q * scaling).float().view(B, H, 1, 64)
    kk = kc0[:, : pos + 1].float().view(B, pos + 1, H, 64).transpose(1, 2)
    vv = vc0[:, : pos + 1].float().view(B, pos + 1, H, 64).transpose(1, 2)
    ref = torch.softmax(qs @ kk.transpose(-1, -2), dim=-1) @ vv
    ref = ref.transpose(1, 2).reshape(B, d)
    tol = 2e-3 * max(1.0, float(vv.abs().max()))
    assert (out.float() - ref).abs().max().item() <= tol
    if threshold is None:
        assert quant is None
    else:
        _check_quant(quant, out, threshold)


def test_self_attn_decode_matches_torch_sdpa_and_argument_checks():
    B, H, t_max, pos = 4, 8, 64, 20
    d = H * 64
    q, k, v = ((torch.randn(B, d, device="cuda")).half() for _ in range(3))
    kc = torch.randn(B, t_max, d, device="cuda").half()
    vc = torch.randn(B, t_max, d, device="cuda").half()
    pos_t = torch.tensor([pos], dtype=torch.int64, device="cuda")
    out, _ = F.self_attn_decode(q, k, v, 0.125, kc, vc, pos_t, H)
    qh = (q * 0.125).view(B, 1, H, 64).transpose(1, 2)
    kh = kc[:, : pos + 1].view(B, pos + 1, H, 64).transpose(1, 2)
    vh = vc[:, : pos + 1].view(B, pos + 1, H, 64).transpose(1, 2)
    ref = TF.scaled_dot_product_attention(qh, kh, vh, scale=1.0).transpose(1, 2).reshape(B, d)
    assert (out.float() - ref.float()).abs().max().item() <= 4e-3
    with pytest.raises(RuntimeError, match="head_dim 64"):
        F.self_attn_decode(q, k, v, 0.125, kc, vc, pos_t, 4)
    with pytest.raises(RuntimeError, match="int64"):
        F.self_attn_decode(q, k, v, 0.125, kc, vc, pos_t.int(), H)


@pytest.mark.parametrize("B,V,ld", [(7, 51865, 51872), (256, 51865, 51872), (3, 1000, 1000), (1, 51866, 51872),
                                    (5, 13, 16)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_masked_argmax_matches_torch(B, V, ld, dtype):
    g = torch.Generator(device="cuda").manual_seed(B + V)
    buf = torch.randn(B, ld, device="cuda", generator=g).to(dtype)
    logits = buf[:, :V]
    # ties (coarse values), a masked maximum, NaN rows, a fully masked row set
    logits[0] = (logits[0] * 2).round() / 2
    mask = torch.rand(V, device="cuda", generator=g) < 0.3
    if B > 2:
        logits[1, V // 2] = float("nan")
        logits[2, :] = -7.0                                   # all equal -> first unmasked index
    ref = torch.argmax(logits.float().masked_fill(mask, float("-inf")), dim=-1)
    got = F.masked_argmax(logits, mask)
    assert torch.equal(got, ref)
    assert torch.equal(F.masked_argmax(logits), torch.argmax(logits.float(), dim=-1))
    all_masked = torch.ones(V, dtype=torch.bool, device="cuda")
    ref_all = torch.argmax(logits.float().masked_fill(all_masked, float("-inf")), dim=-1)
    assert torch.equal(F.masked_argmax(logits, all_masked), ref_all)
    with pytest.raises(RuntimeError):
        F.masked_argmax(torch.randn(4, 51865, device="cuda").half())      # rows not 16-byte aligned


# (300, 8, 333) and (160, 8, 1500): more (utterance, head) items than the persistent grid has CTAs -- every CTA walks
# several items (prefetch across the item boundary, double-buffered merge scratch)
@pytest.mark.parametrize("B,H,S", [(3, 6, 1500), (70, 8, 1500), (2, 20, 1500), (5, 12, 37), (1, 8, 1), (300, 8, 333),
                                   (160, 8, 1500), (97, 20, 129)])
@pytest.mark.parametrize("threshold", [None, 6.0])
@pytest.mark.parametrize("kv_fused", [True, False])
def test_cross_attn_decode(B, H, S, threshold, kv_fused):
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * 100 + H + S)
    q = (torch.randn(B, d, device="cuda", generator=g) * 2.0).half()
    scale_v = 8.0 if threshold else 1.0
    if kv_fused:
        kv = torch.randn(B, S, 2 * d, device="cuda", generator=g).half()
        kv[:, :, d:] *= scale_v
        k, v = kv[:, :, :d], kv[:, :, d:]
    else:
        k = torch.randn(B, S, d, device="cuda", generator=g).half()
        v = (torch.randn(B, S, d, device="cuda", generator=g) * scale_v).half()
    if threshold:
        k[:, 0] *= 4.0          # a peaked softmax keeps some outputs above the outlier threshold
    out, quant = F.cross_attn_decode(q, k, v, 0.125, H, threshold)
    qs = (q * 0.125).float().view(B, H, 1, 64)
    kk = k.float().reshape(B, S, H, 64).transpose(1, 2)
    vv = v.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = (torch.softmax(qs @ kk.transpose(-1, -2), dim=-1) @ vv).transpose(1, 2).reshape(B, d)
    assert (out.float() - ref).abs().max().item() <= 2e-3 * max(1.0, float(vv.abs().max()))
    if threshold is None:
        assert quant is None
    else:
        _check_quant(quant, out, threshold)
        assert int(F._row_counters(q.device, B).abs().sum()) == 0        # self-resetting scratch
        out2, quant2 = F.cross_attn_decode(q, k, v, 0.125, H, threshold)  # and therefore re-usable
        assert torch.equal(out2, out) and torch.equal(quant2[0], quant[0]) and torch.equal(quant2[1], quant[1])
        quant2[2].col_flags.zero_()


def test_cross_attn_decode_matches_torch_sdpa_bf16_and_errors():
    B, H, S = 4, 8, 1500
    d = H * 64
    q = torch.randn(B, d, device="cuda").bfloat16()
    k = torch.randn(B, S, d, device="cuda").bfloat16()
    v = torch.randn(B, S, d, device="cuda").bfloat16()
    out, _ = F.cross_attn_decode(q, k, v, 0.125, H)
    ref = TF.scaled_dot_product_attention((q * 0.125).view(B, 1, H, 64).transpose(1, 2),
                                          k.view(B, S, H, 64).transpose(1, 2), v.view(B, S, H, 64).transpose(1, 2),
                                          scale=1.0).transpose(1, 2).reshape(B, d)
    assert (out.float() - ref.float()).abs().max().item() <= 2e-2
    with pytest.raises(RuntimeError, match="fp16"):
        F.cross_attn_decode(q, k, v, 0.125, H, threshold=6.0)
    with pytest.raises(RuntimeError, match="strides"):
        F.cross_attn_decode(q, k.transpose(1, 2).contiguous().transpose(1, 2), v, 0.125, H)


# ------------------------------------------------------------------------------------------------
# fp32 caches: the reference's quanto / bnb *_32 flows keep fp32 activations (model_utils.py:139-142), HF then runs
# torch SDPA in fp32 (modeling_whisper.py:338-352); the same decode kernels serve it with 32-byte row chunks
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,S,fused", [(3, 6, 1500, False), (5, 16, 1500, True), (2, 20, 777, True), (300, 8, 96, False)])
def test_cross_attn_decode_fp32_matches_float64_attention(B, H, S, fused):
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B + H + S)
    q = torch.randn(B, d, device="cuda", generator=g)
    if fused:                        # K and V as the two column blocks of one [B, S, 2d] buffer (fastgen's layout)
        kv = torch.randn(B, S, 2 * d, device="cuda", generator=g)
        k, v = kv[:, :, :d], kv[:, :, d:]
    else:
        k = torch.randn(B, S, d, device="cuda", generator=g)
        v = torch.randn(B, S, d, device="cuda", generator=g)
    out, quant = F.cross_attn_decode(q, k, v, 0.125, H)
    assert quant is None and out.dtype == torch.float32
    qh = (q * 0.125).double().view(B, 1, H, 64).transpose(1, 2)
    kh = k.double().reshape(B, S, H, 64).transpose(1, 2)
    vh = v.double().reshape(B, S, H, 64).transpose(1, 2)
    w = torch.softmax(qh @ kh.transpose(-1, -2), dim=-1)
    ref = (w @ vh).transpose(1, 2).reshape(B, d)
    assert (out.double() - ref).abs().max().item() <= 2e-5
    # the torch fp32 SDPA HF would call agrees to the same bar
    sd = TF.scaled_dot_product_attention(qh.float(), kh.float(), vh.float(), scale=1.0).transpose(1, 2).reshape(B, d)
    assert (out - sd).abs().max().item() <= 2e-5


@pytest.mark.parametrize("B,H,t_max,pos", [(4, 16, 64, 20), (2, 20, 128, 0), (64, 6, 72, 71)])
def test_self_attn_decode_fp32_matches_float64_attention_and_appends(B, H, t_max, pos):
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B + H + pos)
    qkv = torch.randn(B, 3 * d, device="cuda", generator=g)           # one fused projection: common row stride
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    kc = torch.randn(B, t_max, d, device="cuda", generator=g)
    vc = torch.randn(B, t_max, d, device="cuda", generator=g)
    kc0, vc0 = kc.clone(), vc.clone()
    pos_t = torch.tensor([pos], dtype=torch.int64, device="cuda")
    out, _ = F.self_attn_decode(q, k, v, 0.125, kc, vc, pos_t, H)
    assert torch.equal(kc[:, pos], k) and torch.equal(vc[:, pos], v)            # appended in place
    keep = torch.ones(t_max, dtype=torch.bool, device="cuda")
    keep[pos] = False
    assert torch.equal(kc[:, keep], kc0[:, keep]) and torch.equal(vc[:, keep], vc0[:, keep])
    qh = (q * 0.125).double().view(B, 1, H, 64).transpose(1, 2)
    kh = kc[:, : pos + 1].double().view(B, pos + 1, H, 64).transpose(1, 2)
    vh = vc[:, : pos + 1].double().view(B, pos + 1, H, 64).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2), dim=-1) @ vh).transpose(1, 2).reshape(B, d)
    assert (out.double() - ref).abs().max().item() <= 2e-5

"""The oracle against the golden fixtures (tests/golden/, made by make_golden.py from the live
torch / HF implementations the reference calls) and against its own algebraic invariants."""
import os

import numpy as np
import pytest

import oracle
from tests.helpers import synth_audio


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_dynamic_weight_codes_bit_exact(golden_dir, tag):
    g = _load(golden_dir, f"torch_dynamic_{tag}.npz")
    q, s = oracle.torch_weight_qint8(g["w"])
    assert np.float32(s) == g["w_scale"]
    assert int(g["w_zp"]) == 0
    np.testing.assert_array_equal(q, g["w_int"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_dynamic_activation_codes_bit_exact(golden_dir, tag):
    g = _load(golden_dir, f"torch_dynamic_{tag}.npz")
    q, s, zp = oracle.torch_act_quant(g["x"], True)
    assert np.float32(s) == g["x_scale"]
    assert zp == int(g["x_zp"])
    np.testing.assert_array_equal(q, g["x_int"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_dynamic_linear_output(golden_dir, tag):
    g = _load(golden_dir, f"torch_dynamic_{tag}.npz")
    y = oracle.torch_dynamic_linear(g["x"], g["w_int"], float(g["w_scale"]), g["bias"])
    # integer part is exact; the fp32 requant differs from FBGEMM/oneDNN only in rounding order
    np.testing.assert_allclose(y, g["y"], rtol=2e-6, atol=2e-6)


def test_torch_dynamic_pruned_zeros_survive(golden_dir):
    g = _load(golden_dir, "torch_dynamic_b.npz")
    q, _ = oracle.torch_weight_qint8(g["w"])
    assert (g["w"] == 0).mean() >= 0.5
    assert np.all(q[g["w"] == 0] == 0)


def test_prune_global_l1_golden(golden_dir):
    """prune semantics the oracle assumes: exactly round(0.5*n) smallest |w| across both
    tensors are zeroed, everything else untouched."""
    g = _load(golden_dir, "prune_global_l1.npz")
    w = np.concatenate([g["w0"].ravel(), g["w1"].ravel()])
    p = np.concatenate([g["p0"].ravel(), g["p1"].ravel()])
    k = int(round(0.5 * w.size))
    assert (p == 0).sum() == k
    thr = np.sort(np.abs(w))[k - 1]
    assert np.all(np.abs(w[p == 0]) <= thr)
    np.testing.assert_array_equal(p[p != 0], w[p != 0])


@pytest.mark.parametrize("mels", [80, 128])
def test_logmel_2s_matches_hf(golden_dir, mels):
    g = _load(golden_dir, f"logmel_2s_{mels}.npz")
    fb = oracle.mel_filter_bank_slaney(mels)
    np.testing.assert_allclose(fb, g["mel_filters"], rtol=0, atol=1e-12)
    for i, (n, seed) in enumerate(zip(g["lengths"], g["seeds"])):
        out = oracle.log_mel_spectrogram(synth_audio(int(seed), int(n)), n_mels=mels, n_samples=32000)
        assert out.shape == (1, mels, 200)
        np.testing.assert_allclose(out[0], g["feats"][i], rtol=0, atol=2e-5)


def test_logmel_30s_matches_hf(golden_dir):
    g = _load(golden_dir, "logmel_30s_80.npz")
    out = oracle.log_mel_spectrogram(synth_audio(int(g["audio_seed"])), n_mels=80)[0]
    assert out.shape == (80, 3000)
    np.testing.assert_allclose(out[:, g["frames"]], g["feats"], rtol=0, atol=2e-5)
    assert abs(out.max() - float(g["fmax"])) < 2e-5
    assert abs(out.sum() - float(g["fsum"])) < 0.5


# ---------------------------------------------------------------------------------------------
# self-derived known answers for the UNPINNED restatements (SURVEY.md section 8c "what we pin")
# ---------------------------------------------------------------------------------------------
def test_nf4_codebook_from_normal_quantiles():
    """QLoRA create_normal_map(offset=0.9677083): regenerate the 16 NF4 values."""
    from scipy.stats import norm
    offset = 0.9677083
    v1 = norm.ppf(np.linspace(offset, 0.5, 9)[:-1]).tolist()
    v3 = (-norm.ppf(np.linspace(offset, 0.5, 8)[:-1])).tolist()
    vals = np.array(sorted(v1 + [0.0] + v3), dtype=np.float64)
    vals /= vals.max()
    np.testing.assert_allclose(oracle.NF4_CODE, vals.astype(np.float32), rtol=0, atol=2e-7)


def test_nf4_thresholds_are_midpoints_and_roundtrip():
    code = oracle.NF4_CODE.astype(np.float64)
    mids = (code[1:] + code[:-1]) / 2
    # a value just above / below each midpoint lands in the upper / lower bin
    blk = np.zeros(64, dtype=np.float32)
    blk[0] = 1.0  # absmax = 1 -> normalised value == value
    for i, m in enumerate(mids):
        for delta, want in ((+1e-4, i + 1), (-1e-4, i)):
            b = blk.copy()
            b[1] = np.float32(m + delta)
            packed, absmax = oracle.quantize_4bit(b)
            assert absmax[0] == 1.0
            assert (packed[0, 0] >> 4) == 15
            assert (packed[0, 0] & 15) == want
    # exact codebook values round-trip exactly
    b = np.zeros(64, dtype=np.float32)
    b[:16] = oracle.NF4_CODE
    packed, absmax = oracle.quantize_4bit(b)
    back = oracle.dequantize_4bit(packed, absmax, (64,), dtype=np.float32)
    np.testing.assert_array_equal(back[:16], oracle.NF4_CODE)
    assert np.all(back[16:] == 0)


def test_nf4_zero_preserved_and_all_zero_block():
    rng = np.random.RandomState(0)
    w = (rng.randn(8, 128) * 0.02).astype(np.float16)
    w[rng.rand(8, 128) < 0.5] = 0
    w[3] = 0  # two all-zero blocks: absmax 0, 0*inf = NaN -> code 0, dequant -0.0 == 0
    packed, absmax = oracle.quantize_4bit(w)
    back = oracle.dequantize_4bit(packed, absmax, w.shape, dtype=np.float16)
    assert np.all(back[w == 0] == 0)
    assert absmax[6] == 0 and absmax[7] == 0
    assert packed.shape == (8 * 128 // 2, 1) and absmax.shape == (16,)


def test_nf4_ragged_and_empty():
    w = np.linspace(-1, 1, 70).astype(np.float32)  # 64 + 6, odd count handled below
    packed, absmax = oracle.quantize_4bit(w)
    assert packed.shape == (35, 1) and absmax.shape == (2,)
    w = np.linspace(-1, 1, 67).astype(np.float32)
    packed, absmax = oracle.quantize_4bit(w)
    assert packed.shape == (34, 1) and (packed[33, 0] & 15) == 7
    packed, absmax = oracle.quantize_4bit(np.zeros((0,), np.float32))
    assert packed.shape == (0, 1) and absmax.shape == (0,)


def test_bnb_int8_rowmax_is_127_and_outliers():
    rng = np.random.RandomState(1)
    a = rng.randn(16, 96).astype(np.float16)
    ca, stats, cols = oracle.int8_vectorwise_quant(a, 0.0)
    assert cols is None
    assert np.all(np.abs(ca).max(1) == 127)
    np.testing.assert_array_equal(stats, np.abs(a.astype(np.float32)).max(1))
    a[2, 5] = 7.5
    a[9, 40] = -6.0  # |a| >= threshold is an outlier (not <)
    ca, stats, cols = oracle.int8_vectorwise_quant(a, 6.0)
    np.testing.assert_array_equal(cols, [5, 40])
    assert np.all(ca[:, 5] == 0) and np.all(ca[:, 40] == 0)
    assert stats[2] < 6.0 and stats[9] < 6.0


def test_bnb_linear8bit_matches_fp_reference():
    rng = np.random.RandomState(2)
    W = (rng.randn(48, 64) * 0.05).astype(np.float16)
    x = rng.randn(7, 64).astype(np.float16)
    x[1, 3] = 9.0
    bias = (rng.randn(48) * 0.1).astype(np.float16)
    CB, SCB, _ = oracle.int8_vectorwise_quant(W, 0.0)
    y, extra = oracle.linear8bitlt_forward(x, CB, SCB, bias, 6.0)
    assert extra is not None
    ref = x.astype(np.float64) @ W.astype(np.float64).T + bias.astype(np.float64)
    assert np.abs(y.astype(np.float64) - ref).max() < 0.05


def test_quanto_qint8_invariants():
    rng = np.random.RandomState(3)
    W = (rng.randn(32, 80) * 0.02).astype(np.float32)
    W[rng.rand(32, 80) < 0.5] = 0
    W[7] = 0
    q, scale = oracle.quanto_qint8(W)
    assert np.all(q[W == 0] == 0)
    assert scale[7, 0] == 0 and np.all(q[7] == 0)
    rows = [i for i in range(32) if i != 7]
    assert np.all(np.abs(q[rows]).max(1) == 127)
    np.testing.assert_array_equal(scale[:, 0], np.abs(W).max(1) / np.float32(127))
    deq = q.astype(np.float32) * scale
    assert np.abs(deq - W).max() <= scale.max() * 0.5 + 1e-9


def test_edit_distance_and_tally():
    assert oracle.edit_distance([1, 2, 3], [1, 2, 3]) == 0
    assert oracle.edit_distance([1, 2, 3], []) == 3
    assert oracle.edit_distance([], [4]) == 1
    assert oracle.edit_distance([1, 2, 3, 4], [1, 3, 4, 5]) == 2
    t = oracle.wer_cer_tally(["the cat sat", "a b"], ["the cat sat down", "a c"])
    assert list(t[:2]) == [2, 5]
    assert t[3] == len("the cat sat") + len("a b")


def test_quanto_qint2_invariants():
    """quanto qint2 = the qint4 scheme with three levels above zero."""
    rng = np.random.RandomState(2)
    W = (rng.randn(8, 256) * 0.05).astype(np.float32)
    q, scale, shift, g = oracle.quanto_qint4(W, bits=2)
    assert g == 128 and q.max() == 3 and q.min() == 0
    Wg = W.reshape(8, 2, 128)
    np.testing.assert_array_equal(scale, ((Wg.max(2) - Wg.min(2)) / np.float32(3.0)).astype(np.float32))
    np.testing.assert_array_equal(shift, -Wg.min(2))
    qg = q.reshape(8, 2, 128)
    assert (qg.max(2) == 3).all() and (qg.min(2) == 0).all()       # every group spans the whole code range
    deq = oracle.quanto_qint4_dequant(q, scale, shift, g)
    assert np.abs(deq - W).max() <= scale.max() * 0.5 * (1 + 1e-5)
    # the 4-bit scheme on the same weights resolves 5x finer
    q4, s4, _, _ = oracle.quanto_qint4(W)
    np.testing.assert_allclose(scale, s4 * 5.0, rtol=1e-6)


def test_quanto_qint4_invariants():
    rng = np.random.RandomState(4)
    W = (rng.randn(16, 384) * 0.02).astype(np.float32)
    q, scale, shift, g = oracle.quanto_qint4(W)
    assert g == 128 and scale.shape == shift.shape == (16, 3)
    grp = W.reshape(16, 3, 128)
    np.testing.assert_array_equal(shift, -grp.min(-1))
    np.testing.assert_array_equal(scale, ((grp.max(-1) - grp.min(-1)) / np.float32(15)).astype(np.float32))
    # group minimum -> code 0, group maximum -> code 15; reconstruction within half a step
    qg = q.reshape(16, 3, 128)
    assert np.all(np.take_along_axis(qg, grp.argmin(-1)[..., None], -1) == 0)
    assert np.all(np.take_along_axis(qg, grp.argmax(-1)[..., None], -1) == 15)
    deq = oracle.quanto_qint4_dequant(q, scale, shift, g)
    assert np.abs(deq - W).max() <= scale.max() * 0.5 * (1 + 1e-5)
    assert oracle.quanto_group_size(64) == 64 and oracle.quanto_group_size(160) == 32
    assert oracle.quanto_qint4_pack(q).shape == (16, 192)


def test_dynamic_map_and_nested_absmax_quantization():
    code = oracle.dynamic_map()
    assert code.shape == (256,) and np.all(np.diff(code) > 0) and code[-1] == 1.0 and 0.0 in code
    assert code[0] == -code[-2]                       # signed, symmetric except for the extra +1.0
    rng = np.random.RandomState(5)
    absmax = (np.abs(rng.randn(700)) * 0.03 + 0.01).astype(np.float32)
    q, a2, off, deq = oracle.quantize_absmax_double(absmax)
    assert q.dtype == np.uint8 and a2.shape == (3,)
    assert off == np.float32(absmax.astype(np.float64).mean())
    # every block's extreme element maps to code +-1 (255 or 0) and is reproduced exactly
    for b in range(3):
        blk = slice(256 * b, min(256 * (b + 1), 700))
        i = np.abs(absmax[blk] - off).argmax()
        assert q[blk][i] in (0, 255)
        assert deq[blk][i] == np.float32(np.float32(code[q[blk][i]] * a2[b]) + off)
    assert np.abs(deq - absmax).max() <= 0.01 * np.abs(absmax - off).max() + 1e-7


def test_fdividef_table_reproduces_the_b200_sweep():
    """SURVEY A.2 open point "exact reciprocal", closed on a B200 (scripts/bnb_open_points.cu): bitsandbytes' row
    scale __fdividef(127, absmax) against the IEEE quotient over ALL fp16 (absmax, 0 <= a <= absmax) pairs.  The
    committed table (tests/golden/fdividef_127_fp16.npz) must reproduce the counts the GPU reported
    (profiles/r02_bnb_open_points.json): that is what pins the oracle's int8 codes to the approximate form."""
    import json
    import os
    table = oracle.fdividef_127_table()
    rep = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                      "r02_bnb_open_points.json")))
    vals = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float32)
    ieee = np.zeros(0x7c00, dtype=np.float32)
    ieee[1:] = np.float32(127.0) / vals[1:]
    assert int((ieee[1:] != table[1:0x7c00]).sum()) == rep["absmax_values_with_different_scale"] == 9185
    pairs = mism = 0
    for hb in range(1, 0x7c00):
        a = vals[: hb + 1]
        q1 = np.rint(a * table[hb])
        q2 = np.rint(a * ieee[hb])
        pairs += hb + 1
        mism += int((q1 != q2).sum())
    assert pairs == rep["fp16_pairs"] and mism == rep["code_mismatches"] == 8734
    # the listed pairs, through the C oracle: approximate form = GPU's code, IEEE form = the other one
    for rec in rep["first_mismatches"][:16]:
        am = np.array([rec["absmax_bits"]], dtype=np.uint16).view(np.float16)[0]
        a = np.array([rec["a_bits"]], dtype=np.uint16).view(np.float16)[0]
        row = np.array([[am, a]], dtype=np.float16)
        assert int(oracle.int8_vectorwise_quant(row, 0.0)[0][0, 1]) == rec["q_fdividef"]
        assert int(oracle.int8_vectorwise_quant(row, 0.0, approx_div=False)[0][0, 1]) == rec["q_ieee"]
    # int8_mm_dequant: written as mul-then-add it is contracted to the same FFMA as fmaf (0 mismatches in 2.5e9)
    assert rep["plain_vs_fmaf_f32_mismatches"] == 0


def test_quanto_fp16_opmath_restatement_matches_live_torch():
    """oracle.quanto_qint8(dtype=float16) restates `absmax / 127` and `round(w / scale)` as torch evaluates them on
    half tensors (fp32 math, result rounded to half): pinned against torch's own CPU ops."""
    import torch
    rng = np.random.RandomState(3)
    w = (rng.randn(64, 200) * 0.02).astype(np.float16)
    w[rng.rand(64, 200) < 0.3] = 0
    w[5] = 0
    t = torch.from_numpy(w)
    s = t.abs().amax(dim=1, keepdim=True) / 127
    q = torch.clamp(torch.nan_to_num(torch.round(t / s), nan=0.0), -128, 127).to(torch.int8)
    q_ref, s_ref = oracle.quanto_qint8(w, np.float16)
    np.testing.assert_array_equal(q.numpy(), q_ref)
    np.testing.assert_array_equal(s.float().numpy(), s_ref)

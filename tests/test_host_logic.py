"""CPU-side checks: the C-ABI library loads and exports every symbol include/whisperq.h declares
(no compute calls without a GPU), the drop-in swaps build the module structure the reference
flows expect, the product path refuses to run on the CPU, and the N>1 tally path works over gloo.
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MICRO = dict(encoder_layers=1, decoder_layers=1, encoder_attention_heads=2, decoder_attention_heads=2,
             d_model=64, encoder_ffn_dim=128, decoder_ffn_dim=128, max_source_positions=50)


@pytest.fixture(scope="module")
def lib_path():
    from openai_whisper_compression_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    hdr = open(os.path.join(ROOT, "include", "whisperq.h")).read()
    declared = set(re.findall(r"\b(wq_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in whisperq.h but not exported"
    from openai_whisper_compression_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    lib.wq_version.restype = ctypes.c_int
    assert lib.wq_version() >= 100


def test_sass_has_blackwell_tensor_and_tma_instructions(lib_path):
    out = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    for mnem in ("UTCIMMA", "UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in out, f"{mnem} missing from SASS"
    # CTA-pair schedule (tcgen05 cta_group::2): pair MMAs of both kinds, pair TMA loads, multicast commits, cluster barrier
    for mnem in ("UTCIMMA.2CTA", "UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "UCGABAR_ARV"):
        assert mnem in out, f"{mnem} missing from SASS"
    assert "HMMA.16816" not in out      # no legacy mma.sync path


def test_argument_validation_needs_no_gpu(lib_path):
    from openai_whisper_compression_b200 import _lib
    lib = _lib.load()
    rc = lib.wq_quant_4bit(None, 1, 128, 48, 0, None, None, None)     # bad blocksize
    assert rc == 1 and b"blocksize" in lib.wq_last_error()
    rc = lib.wq_logmel(None, 1, 480000, None, 480001, None, 80, None, 0, None, None)
    assert rc == 1 and b"n_samples" in lib.wq_last_error()
    assert lib.wq_quant_4bit(None, 1, 0, 64, 0, None, None, None) == 0  # empty input is a no-op
    with pytest.raises(RuntimeError):
        _lib.check(1, "x")


def test_cpu_forward_raises_no_fallback():
    from openai_whisper_compression_b200 import bnb, dynamic, functional, quanto
    with pytest.raises(RuntimeError):
        functional.quantize_4bit(torch.zeros(64))
    m = bnb.Linear4bit(64, 32, compress_statistics=False, quant_type="nf4")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 64))
    m8 = bnb.Linear8bitLt(64, 32, has_fp16_weights=False, threshold=6.0)
    with pytest.raises(RuntimeError):
        m8(torch.zeros(1, 64))
    seq = nn.Sequential(nn.Linear(64, 32))
    quanto.quantize(seq, weights=quanto.qint8)
    quanto.freeze(seq)
    with pytest.raises(RuntimeError):
        seq(torch.zeros(1, 64))
    d = dynamic.quantize_dynamic(nn.Sequential(nn.Linear(64, 32)), {nn.Linear}, dtype=torch.qint8)
    with pytest.raises(RuntimeError):
        d(torch.zeros(1, 64))
    with pytest.raises(NotImplementedError):
        quanto.quantize(nn.Sequential(nn.Linear(8, 8)), weights=quanto.qint2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openai_whisper_compression_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_swaps_build_reference_module_structure():
    from openai_whisper_compression_b200 import bnb, dynamic, harness, quanto, swap
    model = harness.build_model("tiny", **MICRO)
    n_lin = sum(type(m) is nn.Linear for m in model.modules())
    # HF load_in_8bit / load_in_4bit: every linear except the output embedding
    m8 = swap.replace_with_bnb_linear(harness.build_model("tiny", **MICRO).half(), load_in_8bit=True)
    assert sum(isinstance(m, bnb.Linear8bitLt) for m in m8.modules()) == n_lin - 1
    assert type(m8.proj_out) is nn.Linear and m8.model.encoder.layers[0].self_attn.k_proj.bias is None
    assert all(isinstance(m, nn.Linear) for m in m8.modules() if isinstance(m, bnb.Linear8bitLt))
    # the reference's convert_model_to_4bit: all linears incl. proj_out, original fp32 weights kept
    m4 = swap.convert_model_to_4bit(harness.build_model("tiny", **MICRO))
    assert sum(isinstance(m, bnb.Linear4bit) for m in m4.modules()) == n_lin
    w_ref = harness.build_model("tiny", **MICRO).model.encoder.layers[0].fc1.weight
    assert torch.equal(m4.model.encoder.layers[0].fc1.weight.data, w_ref.data)
    assert m4.model.encoder.layers[0].fc1.weight.quant_type == "nf4"
    # quanto: all linears incl. proj_out; torch-dynamic twin likewise
    mq = harness.build_model("tiny", **MICRO)
    quanto.quantize(mq, weights=quanto.qint8)
    assert sum(isinstance(m, quanto.QLinear) for m in mq.modules()) == n_lin
    md = dynamic.quantize_dynamic(harness.build_model("tiny", **MICRO), {nn.Linear}, dtype=torch.qint8, inplace=True)
    assert sum(isinstance(m, dynamic.DynamicInt8Linear) for m in md.modules()) == n_lin
    assert md.proj_out.weight().shape == (51865, 64)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_reference_modules_import_and_run_unchanged_with_shims(tmp_path, monkeypatch):
    """model_utils.py / evaluation.py / data_utils.py import unchanged; load_whisper_model(...,
    'quanto_int8') runs the reference's own code path and yields this package's QLinear."""
    from openai_whisper_compression_b200 import harness, quanto, swap
    swap.install_shims()
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend("/root/reference")
    import importlib
    model_utils = importlib.import_module("model_utils")
    importlib.import_module("data_utils")
    importlib.import_module("evaluation")
    harness.build_model("tiny", **MICRO).save_pretrained(tmp_path / "m")
    model = model_utils.load_whisper_model(str(tmp_path / "m"), torch.device("cpu"), quantization="quanto_int8")
    assert isinstance(model.proj_out, quanto.QLinear) and model.model.decoder.layers[0].fc2._freeze_pending
    assert model_utils.get_model_disk_size_in_mb(model) > 0
    m4 = model_utils.load_whisper_model(str(tmp_path / "m"), torch.device("cpu"), quantization="quanto_int4")
    assert m4.model.encoder.layers[0].fc1.weight_qtype is quanto.qint4
    cfg = model_utils._create_bnb_config("bnb_nf4_16")
    assert cfg.bnb_4bit_quant_type == "nf4" and cfg.llm_int8_threshold == 6.0


def test_shard_range_partitions_utterances():
    from openai_whisper_compression_b200.tally import shard_range
    for n in (0, 1, 7, 256, 257):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in shard_range(n, r, w)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import oracle
from openai_whisper_compression_b200 import tally
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
refs = ["the cat sat on the mat", "a b c", "hello world again", "x y", "one two three"]
hyps = ["the cat sat mat", "a x c d", "hello world", "x y", "one three"]
mine = tally.shard_range(len(refs), rank, 2)
t = torch.from_numpy(oracle.wer_cer_tally([refs[i] for i in mine], [hyps[i] for i in mine]))
t = tally.all_reduce_tally(t)
want = torch.from_numpy(oracle.wer_cer_tally(refs, hyps))
assert torch.equal(t, want), (t, want)
r = tally.rates(t)
assert abs(r["WER"] - 100.0 * int(want[0]) / int(want[1])) < 1e-9
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_tally_all_reduce_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_tally_host_pack_layout_matches_oracle_distances():
    """tally._host_pack: 2P pairs (words then code points) in one id array + one offset array; the distances
    the oracle computes from the packed layout equal those of the plain per-pair definition."""
    import oracle
    from openai_whisper_compression_b200 import tally
    refs = ["the quick brown fox", "", "jumps over  the lazy dog", "ünïcödé wörds here", "same same"]
    hyps = ["the quick brown box", "spurious", "jump over the dog", "unicode words here", "same same"]
    ids, off, n_ref, n_rw, n_rc = tally._host_pack(refs, hyps)
    P = len(refs)
    assert n_rw == sum(len(r.split()) for r in refs) and n_rc == sum(len(r) for r in refs)
    assert n_ref == n_rw + n_rc and off.shape == (2 * (2 * P + 1),)
    ref_ids, hyp_ids = ids[:n_ref], ids[n_ref:]
    ro, ho = off[:2 * P + 1], off[2 * P + 1:]
    assert ro[0] == 0 and ho[0] == 0 and ro[-1] == n_ref and ho[-1] == len(hyp_ids)
    for p in range(P):
        vocab = {}
        want_w = oracle.edit_distance([vocab.setdefault(w, len(vocab)) for w in refs[p].split()],
                                      [vocab.setdefault(w, len(vocab)) for w in hyps[p].split()])
        want_c = oracle.edit_distance([ord(c) for c in refs[p]], [ord(c) for c in hyps[p]])
        got_w = oracle.edit_distance(list(ref_ids[ro[p]:ro[p + 1]]), list(hyp_ids[ho[p]:ho[p + 1]]))
        got_c = oracle.edit_distance(list(ref_ids[ro[P + p]:ro[P + p + 1]]), list(hyp_ids[ho[P + p]:ho[P + p + 1]]))
        assert (got_w, got_c) == (want_w, want_c)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: torch quantize_dynamic on HF Whisper) on a tiny
    bounded sample: one JSON line with the keys the driver reads."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "tiny",
                          "--cpu-sample", "1", "--new-tokens", "3", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "audio-seconds/sec" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None
    # other ranks of a torchrun launch exit quietly
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_stub_processor_decode_forms_agree():
    """StubProcessor.batch_decode: tensor fast path (one fancy index) == list path == per-row decode."""
    from openai_whisper_compression_b200 import harness
    p = harness.StubProcessor.__new__(harness.StubProcessor)
    ids = torch.randint(0, 51865, (7, 5), generator=torch.Generator().manual_seed(0))
    want = [" ".join(f"t{i}" for i in r) for r in ids.tolist()]
    assert p.batch_decode(ids) == want and p.batch_decode(ids.tolist()) == want
    assert [p.decode(r) for r in ids] == want
    assert p.batch_decode(torch.zeros((0, 4), dtype=torch.long)) == []
    assert p.batch_decode([[1, 2], [3]]) == ["t1 t2", "t3"]


def test_decode_shape_dispatch_rules_are_host_side_and_dtype_exact():
    """Which decode-shaped calls of the weight-only schemes take the GEMV (functional._gemv_ok / _gemv_f32_ok): at most
    GEMV_ROWS rows; 16-bit rows with the same output dtype, or fp32 rows with fp32 output (the reference's fp32 flows,
    model_utils.py:139-142) -- never a mixed pair, never more rows (two launches for 64 rows measured slower)."""
    import torch
    from openai_whisper_compression_b200 import functional as F
    rows = F.GEMV_ROWS
    assert rows == 32
    h = torch.zeros(rows, 64, dtype=torch.float16)
    f = torch.zeros(rows, 64, dtype=torch.float32)
    assert F._gemv_ok(h, torch.float16) and not F._gemv_ok(h, torch.float32) and not F._gemv_ok(f, torch.float32)
    assert F._gemv_f32_ok(f, torch.float32) and not F._gemv_f32_ok(f, torch.float16) and not F._gemv_f32_ok(h, torch.float32)
    assert not F._gemv_ok(torch.zeros(rows + 1, 64, dtype=torch.float16), torch.float16)
    assert not F._gemv_f32_ok(torch.zeros(rows + 1, 64), torch.float32)
    assert not F._gemv_ok(torch.zeros(0, 64, dtype=torch.float16), torch.float16)
    # the one cast pass of the fp32 flows in front of the tensor-core GEMM
    assert F._operand(f).dtype == torch.float16 and F._operand(h) is h

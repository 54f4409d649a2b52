"""The drop-in boundary exercised by the reference's OWN callers (missing 5).

Two halves, because the reference tree (/root/reference) exists in the build container only and never travels to the
GPU box, while the product has no CPU arithmetic path:

  * CPU, build container: the reference's evaluation.evaluate_model, data_utils.map_to_feats / transcribe_batch,
    memory_tracker.WhisperMemoryTracker and model_utils.load_whisper_model are imported from their files and run
    UNCHANGED over a datasets.Dataset of synthetic utterances -- with this repo's import shims in place
    (optimum.quanto, evaluate, bitsandbytes) and the objects this repo hands them (a processor stand-in with the
    WhisperProcessor surface they touch, metric objects with evaluate's .compute signature).  The model is the
    reference's own CPU flow (quantization="pytorch").  What this pins: every attribute, call and column the
    reference's callers use on the objects we supply.
  * GPU (tests marked gpu): the same call sequence -- restated line by line from data_utils.py:139-170 and
    evaluation.py:96-116, citing them -- drives the GPU drop-in modules, the CUDA log-mel processor, the
    CUDA-graph decode loop and the GPU edit-distance metrics behind `evaluate.load`.
"""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

REF = "/root/reference"
MICRO = dict(encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             d_model=64, encoder_ffn_dim=256, decoder_ffn_dim=256)


class _OracleMetric:
    """evaluate.load("wer"|"cer") stand-in for the CPU half (the product's metric shim runs on the GPU)."""

    def __init__(self, name):
        self.name = name

    def compute(self, references, predictions):
        t = oracle.wer_cer_tally(list(references), list(predictions))
        return (t[0] / max(t[1], 1)) if self.name == "wer" else (t[2] / max(t[3], 1))


class _CpuProcessor:
    """WhisperProcessor surface the reference touches (data_utils.py:56-60,169-170), HF's own feature extractor."""

    def __init__(self):
        from transformers import WhisperFeatureExtractor
        self.fe = WhisperFeatureExtractor(feature_size=80)

        class Tok:
            @staticmethod
            def normalize(text):
                return " ".join(text.lower().split())
        self.tokenizer = Tok()

    def __call__(self, audio, sampling_rate=16000, return_tensors="pt"):
        return self.fe(audio, sampling_rate=sampling_rate, return_tensors=return_tensors)

    def decode(self, ids, **kw):
        return " ".join(f"t{int(i)}" for i in ids)


def _synthetic_rows(n):
    from openai_whisper_compression_b200 import harness
    return {"audio": [{"array": harness.synth_audio(i, 32000).tolist(), "sampling_rate": 16000} for i in range(n)],
            "text": [harness.synth_reference(i, 6) for i in range(n)]}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")
def test_reference_callers_run_unchanged_over_our_boundary_objects(tmp_path, monkeypatch):
    import datasets
    from transformers import WhisperForConditionalGeneration
    from openai_whisper_compression_b200 import harness, swap
    monkeypatch.chdir(tmp_path)                       # memory_tracker writes whisper_eval.log into the cwd
    monkeypatch.syspath_prepend(REF)
    swap.install_shims()
    for m in ("model_utils", "evaluation", "data_utils", "memory_tracker"):
        sys.modules.pop(m, None)
    import data_utils, evaluation, memory_tracker, model_utils      # the reference's files, unchanged
    assert os.path.realpath(evaluation.__file__).startswith(REF)

    mdir = str(tmp_path / "whisper-micro")
    WhisperForConditionalGeneration(harness.whisper_config("tiny", **MICRO)).save_pretrained(mdir)
    model = model_utils.load_whisper_model(mdir, torch.device("cpu"), quantization="pytorch")     # model_utils.py:131-134
    model.generation_config.max_length = 12
    proc = _CpuProcessor()
    ds = datasets.Dataset.from_dict(_synthetic_rows(4))
    ds = ds.map(lambda b: data_utils.map_to_feats(b, proc))                                        # data_utils.py:44-61
    assert {"audio", "input_features", "reference"} <= set(ds.column_names)
    tracker = memory_tracker.WhisperMemoryTracker("micro", str(tmp_path / "metrics"))
    metrics = {"WER": _OracleMetric("wer"), "CER": _OracleMetric("cer")}
    scores, trans = evaluation.evaluate_model(model, proc, ds, metrics, tracker, "synthetic", batch_size=2, num_warmup=1)
    tracker.close()
    assert {"WER", "CER", "RTF", "avg_cpu_percent", "peak_cpu_percent"} <= set(scores)
    assert len(trans["predictions"]) == 4 and scores["RTF"] > 0 and np.isfinite(scores["WER"])
    # the bnb_* strings of load_whisper_model stop inside HF, before any bitsandbytes symbol is touched: device_map="auto"
    # needs accelerate (absent here), model_utils.py:112-118 -- documented in INTEGRATION.md
    with pytest.raises(Exception) as ei:
        model_utils.load_whisper_model(mdir, torch.device("cpu"), quantization="bnb_nf4_16")
    print("bnb_nf4_16 through HF from_pretrained:", type(ei.value).__name__, str(ei.value)[:120])


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["llm_int8", "bnb_nf4", "quanto_int8"])
def test_reference_call_sequence_on_gpu_modules(scheme):
    """transcribe_batch (data_utils.py:139-170) and the metric block of evaluate_model (evaluation.py:103-116), restated
    call by call, on the GPU drop-ins behind install_shims(): features as nested lists -> np.array -> squeeze(1) ->
    .half() when the first parameter is fp16 -> .to(model.device) -> model.generate(features) -> processor.decode per
    row -> tokenizer.normalize -> 100 * evaluate.load(name).compute(references=, predictions=)."""
    from openai_whisper_compression_b200 import fastgen, harness, swap
    swap.install_shims()
    import evaluate                                    # the shim (or the real package when installed)
    model = harness.apply_scheme(harness.build_model("tiny", encoder_layers=2, decoder_layers=2), scheme, "cuda")
    model.generation_config.max_length = 16
    fastgen.enable(model)
    proc = harness.StubProcessor(80, device="cuda")
    rows = _synthetic_rows(4)
    feats = [proc(np.asarray(a["array"], dtype=np.float32), sampling_rate=a["sampling_rate"],
                  return_tensors="pt").input_features for a in rows["audio"]]                       # data_utils.py:56-58
    batch = {"input_features": [f.numpy().tolist() for f in feats], "audio": rows["audio"],
             "reference": [proc.tokenizer.normalize(t) for t in rows["text"]]}
    with torch.no_grad():
        features = torch.from_numpy(np.array(batch["input_features"], dtype=np.float32)).squeeze(1)  # :141
        if next(model.parameters()).dtype == torch.float16:                                          # :142-143
            features = features.half()
        features = features.to(model.device)                                                         # :144
        predicted_ids = model.generate(features)                                                     # :152
        torch.cuda.synchronize()                                                                     # :153-154
    assert predicted_ids.device == model.device and predicted_ids.shape[0] == 4
    transcription = [proc.decode(ids) for ids in predicted_ids]                                      # :169
    prediction = [proc.tokenizer.normalize(x) for x in transcription]                                # :170
    want = oracle.wer_cer_tally(batch["reference"], prediction)
    for name, num, den in (("wer", want[0], want[1]), ("cer", want[2], want[3])):
        score = 100 * evaluate.load(name).compute(references=batch["reference"], predictions=prediction)   # evaluation.py:112-114
        assert abs(score - 100.0 * num / den) < 1e-9
    # attributes the reference reads on swapped models: state_dict -> torch.save, parameter count
    import io
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)                                                              # model_utils.py:228
    assert buf.tell() > 0 and sum(p.numel() for p in model.parameters()) > 0                         # model_utils.py:244

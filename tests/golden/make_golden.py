"""Generate the golden fixtures under tests/golden/ from the LIVE implementations the
reference calls.  Run in the build container (needs /root/reference for the reference's own
model_utils.py; everything else is torch / transformers from the image):

    python tests/golden/make_golden.py

Sources of truth
  torch_dynamic_*.npz  torch.quantization.quantize_dynamic(model, {nn.Linear}, dtype=qint8)
                       exactly as model_utils.py:131-134 calls it, plus
                       torch.quantize_per_tensor_dynamic for the activation codes.
  logmel_*.npz         transformers WhisperFeatureExtractor called as data_utils.py:56-58.
  prune_*.npz          torch.nn.utils.prune.global_unstructured(L1Unstructured) + prune.remove,
                       pruning/baseline_scripts/unstructured_L1_baseline.py:500-502,525.
  ref_dynamic_generate.npz
                       the reference's own load_whisper_model(path, cpu, quantization="pytorch")
                       (imported unchanged from /root/reference/model_utils.py with a stub
                       optimum.quanto on sys.modules) + model.generate on seeded features.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


from tests.helpers import synth_audio  # noqa: E402  (seeded gaussian noise, sigma 0.1)


def make_torch_dynamic():
    torch.manual_seed(0)
    for tag, N, K, M, prune_frac in [("a", 48, 64, 5, 0.0), ("b", 96, 128, 33, 0.5)]:
        lin = torch.nn.Linear(K, N)
        with torch.no_grad():
            lin.weight.normal_(0, 0.02)
            lin.bias.normal_(0, 0.02)
            if prune_frac:
                import torch.nn.utils.prune as prune
                prune.l1_unstructured(lin, "weight", amount=prune_frac)
                prune.remove(lin, "weight")
        w = lin.weight.detach().clone().numpy()
        b = lin.bias.detach().clone().numpy()
        x = torch.randn(M, K) * 1.5
        wrap = torch.nn.Sequential(lin)
        torch.quantization.quantize_dynamic(wrap, {torch.nn.Linear}, dtype=torch.qint8, inplace=True)
        q = wrap[0]
        wq = q.weight()
        y = q(x)
        xq = torch.quantize_per_tensor_dynamic(x, torch.quint8, True)
        np.savez_compressed(
            os.path.join(HERE, f"torch_dynamic_{tag}.npz"),
            w=w, bias=b, x=x.numpy(), w_int=wq.int_repr().numpy(), w_scale=np.float32(wq.q_scale()),
            w_zp=np.int32(wq.q_zero_point()), x_int=xq.int_repr().numpy(),
            x_scale=np.float32(xq.q_scale()), x_zp=np.int32(xq.q_zero_point()), y=y.detach().numpy(),
            engine=np.array(torch.backends.quantized.engine), torch_version=np.array(torch.__version__))
        print("torch_dynamic", tag, "engine", torch.backends.quantized.engine)


def make_logmel():
    from transformers import WhisperFeatureExtractor
    # 30 s, 80 mels: every 8th frame of utterance 0 and the per-utterance maximum
    fe = WhisperFeatureExtractor(feature_size=80)
    a = synth_audio(0)
    full = fe(a, sampling_rate=16000, return_tensors="np").input_features[0]
    np.savez_compressed(os.path.join(HERE, "logmel_30s_80.npz"), audio_seed=np.int64(0),
                        frames=np.arange(0, 3000, 8), feats=full[:, ::8].astype(np.float32),
                        fmax=np.float32(full.max()), fmin=np.float32(full.min()),
                        fsum=np.float64(full.astype(np.float64).sum()))
    # 2 s chunks (same code path, n_samples = 32000): full output, 80 and 128 mels; one ragged
    # utterance (1.3 s, zero padded) and one over-long (truncated)
    for mels in (80, 128):
        fe2 = WhisperFeatureExtractor(feature_size=mels, chunk_length=2)
        outs, auds = [], []
        for i, n in enumerate((32000, 20800, 40000)):
            au = synth_audio(10 + i, n)
            auds.append(au)
            outs.append(fe2(au, sampling_rate=16000, return_tensors="np").input_features[0])
        np.savez_compressed(os.path.join(HERE, f"logmel_2s_{mels}.npz"),
                            lengths=np.array([32000, 20800, 40000]), seeds=np.array([10, 11, 12]),
                            feats=np.stack(outs).astype(np.float32),
                            mel_filters=fe2.mel_filters.astype(np.float64))
    print("logmel done")


def make_prune():
    import torch.nn.utils.prune as prune
    torch.manual_seed(1)
    lins = [torch.nn.Linear(32, 24), torch.nn.Linear(24, 40)]
    ws = [l.weight.detach().clone().numpy() for l in lins]
    prune.global_unstructured([(l, "weight") for l in lins], pruning_method=prune.L1Unstructured,
                              amount=0.5)
    for l in lins:
        prune.remove(l, "weight")
    np.savez_compressed(os.path.join(HERE, "prune_global_l1.npz"), w0=ws[0], w1=ws[1],
                        p0=lins[0].weight.detach().numpy(), p1=lins[1].weight.detach().numpy())
    print("prune done")


MICRO = dict(vocab_size=51865, num_mel_bins=80, encoder_layers=2, decoder_layers=2,
             encoder_attention_heads=2, decoder_attention_heads=2, d_model=64,
             encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=100,
             max_target_positions=448)


def make_ref_generate():
    """Run the reference's own load_whisper_model(..., quantization='pytorch') unchanged."""
    stub = types.ModuleType("optimum.quanto")
    for name in ("Calibration", "freeze", "qfloat8", "qint4", "qint8", "quantize"):
        setattr(stub, name, object())
    pkg = types.ModuleType("optimum")
    pkg.quanto = stub
    sys.modules.setdefault("optimum", pkg)
    sys.modules.setdefault("optimum.quanto", stub)
    sys.path.insert(0, "/root/reference")
    import model_utils  # the reference, unchanged
    from transformers import WhisperConfig, WhisperFeatureExtractor, WhisperForConditionalGeneration

    torch.manual_seed(0)
    cfg = WhisperConfig(**MICRO)
    model = WhisperForConditionalGeneration(cfg).eval()
    with tempfile.TemporaryDirectory() as d:
        model.save_pretrained(d)
        qmodel = model_utils.load_whisper_model(d, torch.device("cpu"), quantization="pytorch")
    qmodel.eval()
    fe = WhisperFeatureExtractor(feature_size=80, chunk_length=2)
    feats = np.stack([fe(synth_audio(20 + i, 32000), sampling_rate=16000,
                         return_tensors="np").input_features[0] for i in range(4)])
    T = 12
    with torch.no_grad():
        out = qmodel.generate(torch.from_numpy(feats), do_sample=False, num_beams=1,
                              min_new_tokens=T, max_new_tokens=T, return_dict_in_generate=True,
                              output_logits=True)
    ids = out.sequences.numpy()
    logits = torch.stack(out.logits, 1).numpy()  # [B, T, V] raw (pre-processor) logits
    top2 = np.sort(logits, -1)[..., -2:]
    np.savez_compressed(os.path.join(HERE, "ref_dynamic_generate.npz"), feats=feats.astype(np.float32),
                        ids=ids, margin=(top2[..., 1] - top2[..., 0]).astype(np.float32),
                        top1=logits.argmax(-1), first_logits=logits[:, 0, :].astype(np.float32),
                        T=np.int64(T))
    print("ref generate ids shape", ids.shape, "min margin", float((top2[..., 1] - top2[..., 0]).min()))


if __name__ == "__main__":
    make_torch_dynamic()
    make_logmel()
    make_prune()
    make_ref_generate()

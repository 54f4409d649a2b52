"""Where a bench step goes: log-mel, encoder (+cross KV), graph replays, eager per-token work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200 import harness, fastgen

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 64
dev = torch.device("cuda")
model = harness.apply_scheme(harness.build_model("base"), "llm_int8", dev)
eng = fastgen.enable(model)
proc = harness.StubProcessor(80, device=dev)
audio = torch.randn(B, 480000, device=dev) * 0.1
def sync(): torch.cuda.synchronize()
for _ in range(3):
    feats = proc.feature_extractor.features_from_device_audio(audio).half()
    harness.greedy_generate(model, feats, T)
sync(); t0 = time.perf_counter()
feats = proc.feature_extractor.features_from_device_audio(audio).half()
sync(); t1 = time.perf_counter()
enc = model.model.encoder(feats)
sync(); t2 = time.perf_counter()
ids = harness.greedy_generate(model, feats, T)
sync(); t3 = time.perf_counter()
st = list(eng._states.values())[0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    st.graph.replay()
e1.record(); sync()
print(f"B={B}: logmel {1e3*(t1-t0):.2f} ms, encoder {1e3*(t2-t1):.2f} ms, generate total {1e3*(t3-t2):.2f} ms "
      f"(= encoder + cross-KV + {T} tokens), graph replay {e0.elapsed_time(e1)/20:.3f} ms/token, "
      f"launches per replay {st.launches_per_replay}")
per_tok = (1e3*(t3-t2) - 1e3*(t2-t1)) / T
print(f"  per-token wall {per_tok:.3f} ms -> eager/host part ~{per_tok - e0.elapsed_time(e1)/20:.3f} ms")

# finer split of generate(): HF preamble (+ encoder) | cross-KV | token loop
import types
orig = eng._sample
marks = {}
def timed_sample(input_ids, **kw):
    sync(); marks["enter"] = time.perf_counter()
    out = orig(input_ids, **kw)
    sync(); marks["exit"] = time.perf_counter()
    return out
eng._sample = timed_sample
sync(); g0 = time.perf_counter()
ids = harness.greedy_generate(model, feats, T)
sync(); g1 = time.perf_counter()
print(f"  generate {1e3*(g1-g0):.1f} ms = HF preamble + encoder {1e3*(marks['enter']-g0):.1f} ms + fast loop (cross-KV + tokens) "
      f"{1e3*(marks['exit']-marks['enter']):.1f} ms + post {1e3*(g1-marks['exit']):.1f} ms")

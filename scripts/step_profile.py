"""Kernel-time table of ONE bench step (default workload) from torch.profiler (CUPTI), to see where the step
goes outside the GEMMs.  Not a bench: profiler overhead inflates wall time; kernel durations are what we read."""
import os, sys, collections, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from openai_whisper_compression_b200 import harness, fastgen

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = sys.argv[3] if len(sys.argv) > 3 else "base"
scheme = sys.argv[4] if len(sys.argv) > 4 else "llm_int8"
dev = torch.device("cuda")
model = harness.apply_scheme(harness.build_model(size), scheme, dev)
fastgen.enable(model)
proc = harness.StubProcessor(model.config.num_mel_bins, device=dev)
audio = torch.randn(B, 480000, device=dev) * 0.1
dt = next(model.parameters()).dtype

def step():
    feats = proc.feature_extractor.features_from_device_audio(audio).to(dt)
    return harness.greedy_generate(model, feats, T)

for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"unprofiled step: {e0.elapsed_time(e1):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
spans = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = ev.name
        agg[n][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[n][1] += 1
        spans.append((ev.time_range.start, ev.time_range.end))
tot = sum(v[0] for v in agg.values())
spans.sort()
busy, cur_s, cur_e = 0.0, None, None
for s, e in spans:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
span = spans[-1][1] - spans[0][0] if spans else 0
print(f"GPU kernel time sum {tot/1e3:.1f} ms, busy (union) {busy/1e3:.1f} ms over span {span/1e3:.1f} ms, kernels {sum(v[1] for v in agg.values())}")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% n={c:6d} avg={t/c:8.1f}us  {n[:110]}")

"""Per-kernel time of ONE decode-step graph replay (torch.profiler/CUPTI), default workload."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from openai_whisper_compression_b200 import harness, fastgen

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
size = sys.argv[2] if len(sys.argv) > 2 else "base"
scheme = sys.argv[3] if len(sys.argv) > 3 else "llm_int8"
dev = torch.device("cuda")
model = harness.apply_scheme(harness.build_model(size), scheme, dev)
eng = fastgen.enable(model)
dt = next(model.parameters()).dtype
feats = (torch.randn(B, model.config.num_mel_bins, 3000, device=dev) * 0.5).to(dt)
for _ in range(2):
    harness.greedy_generate(model, feats, 40)
st = list(eng._states.values())[0]
st.pos.fill_(40)
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        st.graph.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(ev.name, [0.0, 0])
        a[0] += ev.device_time
        a[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"B={B} {size} {scheme}: {tot/N:.1f} us kernel time per replay, {sum(v[1] for v in agg.values())//N} kernels")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{t/N:9.1f} us/replay {100*t/tot:5.1f}% n={c//N:4d} avg={t/c:7.2f}us  {n[:120]}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    st.graph.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay {e0.elapsed_time(e1)/50*1e3:.1f} us")

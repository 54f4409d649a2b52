"""Who launches the large strided copies in the encoder / cross-KV stage?  (profiler with python stacks)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from openai_whisper_compression_b200 import harness, fastgen
B = 64
dev = torch.device("cuda")
model = harness.apply_scheme(harness.build_model("base"), sys.argv[1] if len(sys.argv) > 1 else "llm_int8", dev)
fastgen.enable(model)
feats = (torch.randn(B, 80, 3000, device=dev) * 0.5).half()
for _ in range(2):
    harness.greedy_generate(model, feats, 4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    harness.greedy_generate(model, feats, 4)
    torch.cuda.synchronize()
seen = {}
for ev in prof.events():
    if ev.name in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::_to_copy") and ev.input_shapes:
        shp = ev.input_shapes[0]
        n = 1
        for s in shp or []:
            n *= s
        if n >= 10_000_000:
            key = (ev.name, tuple(shp), tuple(ev.stack[:6]))
            seen[key] = seen.get(key, 0) + 1
for (name, shp, stack), c in sorted(seen.items(), key=lambda kv: -kv[1]):
    print(f"{c:3d}x {name} {shp}")
    for fr in stack:
        print("      ", fr)

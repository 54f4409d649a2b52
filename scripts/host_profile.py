"""cProfile of one model.generate call (default bench workload): where the host time of HF's generate goes."""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200 import harness, fastgen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
model = harness.apply_scheme(harness.build_model("base"), "llm_int8", dev)
fastgen.enable(model)
feats = (torch.randn(B, 80, 3000, device=dev) * 0.5).half()
for _ in range(3):
    harness.greedy_generate(model, feats, 64)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
harness.greedy_generate(model, feats, 64)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print("\n".join(l[:170] for l in s.getvalue().splitlines()))

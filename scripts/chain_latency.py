"""In-graph latency of each launch of the LLM.int8 decoder step at one model geometry and row count: every kernel
captured N times back to back in a CUDA graph (the way the step runs them: stream order, programmatic dependent
launch), events around the replay.  A diagnostic for latency-bound decode steps (large-v3 at 16 rows per group), not
a bench.   python scripts/chain_latency.py [d_model] [ffn] [rows] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200 import functional as F

d = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
ffn = int(sys.argv[2]) if len(sys.argv) > 2 else 4 * d
M = int(sys.argv[3]) if len(sys.argv) > 3 else 16
N_REP = int(sys.argv[4]) if len(sys.argv) > 4 else 64
H, S, T_MAX, THR = d // 64, 1500, 128, 6.0
dev = torch.device("cuda")
torch.manual_seed(0)


def weights(n, k):
    cb = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev)
    scb = torch.rand(n, device=dev) * 0.05 + 0.01
    return cb, scb, torch.randn(n, device=dev)


def timed(name, fn, nbytes=0):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N_REP):
            fn()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        if rep:
            ts.append((e0, e1))
    torch.cuda.synchronize()
    us = sum(a.elapsed_time(b) for a, b in ts) / len(ts) / N_REP * 1e3
    extra = f"  {nbytes / us / 1e3:7.0f} GB/s" if nbytes else ""
    print(f"{name:44s} {us:8.2f} us{extra}")
    return us


x = torch.randn(M, d, device=dev, dtype=torch.float16)
delta = torch.randn(M, d, device=dev, dtype=torch.float16) * 0.1
lnw, lnb = torch.ones(d, device=dev, dtype=torch.float16), torch.zeros(d, device=dev, dtype=torch.float16)
tot = 0.0
tot += 3 * timed(f"add_layernorm_quant [{M},{d}]", lambda: F.add_layernorm_quant(x, delta, lnw, lnb, 1e-5, THR))
for name, n, k, cnt in (("qkv", 3 * d, d, 1), ("o / cq / co", d, d, 3), ("fc1", ffn, d, 1), ("fc2", d, ffn, 1)):
    cb, scb, b = weights(n, k)
    a = torch.randn(M, k, device=dev, dtype=torch.float16)
    ca, sca, st = F.int8_vectorwise_quant(a, THR, finalize=False)
    y = torch.empty(M, n, device=dev, dtype=torch.float16)
    tot += cnt * timed(f"gemm_llmint8 {name} [{M},{k}] x [{n},{k}]",
                       lambda: F.gemm_llmint8(ca, sca, cb, scb, b, a_f16=a, state=st, out=y, keep_flags=True),
                       n * k + M * k + 2 * M * n)
qkv = torch.randn(M, 3 * d, device=dev, dtype=torch.float16)
kc = torch.randn(M, T_MAX, d, device=dev, dtype=torch.float16)
vc = torch.randn(M, T_MAX, d, device=dev, dtype=torch.float16)
pos = torch.full((), 32, dtype=torch.int64, device=dev)
tot += timed(f"self_attn_decode [{M} x {H} heads, pos 32]",
             lambda: F.self_attn_decode(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], 0.125, kc, vc, pos, H, THR))
f1 = torch.randn(M, ffn, device=dev, dtype=torch.float16)
tot += timed(f"gelu_quant [{M},{ffn}]", lambda: F.gelu_quant(f1, THR))
ckv = [torch.randn(M, S, 2 * d, device=dev, dtype=torch.float16) for _ in range(4)]
q = torch.randn(M, d, device=dev, dtype=torch.float16)
i = [0]


def xattn():
    c = ckv[i[0] % len(ckv)]
    i[0] += 1
    F.cross_attn_decode(q, c[:, :, :d], c[:, :, d:], 0.125, H, THR)


tot += timed(f"cross_attn_decode [{M} x {H} heads x {S}]", xattn, M * S * 2 * d * 2)
print(f"sum over one decoder layer (12 launches): {tot:.1f} us")

"""Stand-alone timing of the fused GEMMs on encoder / decode shapes (CUDA events, L2 flushed
between iterations).  Usage: python scripts/gemm_bench.py [scheme] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openai_whisper_compression_b200 import functional as F

scheme = sys.argv[1] if len(sys.argv) > 1 else "llmint8"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(96000, 512, 512), (96000, 2048, 512), (96000, 512, 2048), (48000, 1280, 1280), (48000, 5120, 1280),
          (48000, 1280, 5120), (64, 512, 512), (64, 2048, 512), (64, 1280, 5120), (1, 768, 768)]
torch.manual_seed(0)
for M, N, K in shapes:
    if scheme == "llmint8":
        ca = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        cb = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
        sca = torch.rand(M, device=dev) + 0.5
        scb = torch.rand(N, device=dev) * 0.1
        bias = torch.randn(N, device=dev).half()
        fn = lambda: F.gemm_llmint8(ca, sca, cb, scb, bias)
        nbytes = M * K + N * K + 2 * M * N + 4 * (M + N)
    elif scheme == "w8a16":
        x = torch.randn(M, K, device=dev).half()
        wq = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
        sc = torch.rand(N, 1, device=dev) * 0.01
        bias = torch.randn(N, device=dev)
        fn = lambda: F.gemm_w8a16(x, wq, sc, bias)
        nbytes = 2 * M * K + N * K + 2 * M * N + 4 * N
    elif scheme == "w4a16":
        x = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * 0.05).half()
        packed, absmax = F.quantize_4bit(w, 64, "nf4")
        bias = torch.randn(N, device=dev)
        fn = lambda: F.gemm_w4a16(x, packed, absmax, N, K, bias)
        nbytes = 2 * M * K + N * K // 2 + N * K // 16 + 2 * M * N
    elif scheme == "quant":
        x = torch.randn(M, K, device=dev).half()
        fn = lambda: F.int8_vectorwise_quant(x, 6.0)
        nbytes = 3 * M * K + 4 * M
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{scheme} M={M:6d} N={N:5d} K={K:5d}  {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s "
          f"({nbytes / t / 1e9 / peaks['hbm_gbs']:.2f} of HBM)  {2.0 * M * N * K / t / 1e12:7.0f} TFLOP/s", flush=True)

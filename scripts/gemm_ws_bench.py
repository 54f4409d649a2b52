"""A/B timing of the int8 x int8 GEMM's tile shape and schedule on the bench's encoder shapes (M = 384000): run as is
(CTA pairs: 256 x 256 per cluster of two, one cta_group::2 MMA per k-step, weight-stationary when K <= 512), with
WQ_GEMM_PAIR=0 (single-CTA 128 x 256 tiles, one N = 256 MMA per k-step), and then WQ_GEMM_WS=0 (round-robin tiles)
and/or WQ_GEMM_COLS=0 (256 x 128 tiles, two N = 128 MMAs).  CUDA events, L2 flushed between iterations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openai_whisper_compression_b200 import functional as F

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
M = int(sys.argv[2]) if len(sys.argv) > 2 else 384000
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
print("WQ_GEMM_PAIR =", os.environ.get("WQ_GEMM_PAIR", "(default: on)"), " WQ_GEMM_WS =", os.environ.get("WQ_GEMM_WS", "(default: on)"),
      " WQ_GEMM_COLS =", os.environ.get("WQ_GEMM_COLS", "(default: on)"))
SHAPES = ((512, 512), (1024, 512), (1536, 512), (2048, 512), (512, 2048), (384, 384), (1152, 384))
if len(sys.argv) > 4:
    SHAPES = ((int(sys.argv[3]), int(sys.argv[4])),)
for N, K in SHAPES:
    ca = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    cb = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    sca = torch.rand(M, device=dev) + 0.5
    scb = torch.rand(N, device=dev) * 0.1
    bias = torch.randn(N, device=dev).half()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    fn = lambda: F.gemm_llmint8(ca, sca, cb, scb, bias, out=out)
    nbytes = M * K + N * K + 2 * M * N + 4 * (M + N)
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
    # spot check against exact integer sums on a slice of rows
    rows = slice(M - 1000, M)
    c32 = torch._int_mm(ca[rows][:992], cb.t().contiguous())
    ref = torch.addcmul(bias.float()[None, :], (c32.float() * sca[rows][:992, None]) * scb[None, :],
                        torch.full((1,), 6.200012e-05, device=dev)).half()
    bad = (out[rows][:992] != ref).float().mean().item()
    print(f"M={M} N={N:5d} K={K:5d}  {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s ({nbytes / t / 1e9 / peaks['hbm_gbs']:.2f} of HBM) "
          f"{2.0 * M * N * K / t / 1e12:7.0f} TOP/s   mismatch vs int ref {bad:.2e}", flush=True)
    del ca, out

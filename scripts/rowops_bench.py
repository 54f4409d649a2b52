"""Stand-alone timing of the HBM-bound row kernels at the bench's encoder size (384000 rows): GELU+quant through
the table (k_gelu_quant_lut) and through erff (same rows in chunks below the table threshold are NOT comparable in
time, so the erff kernel is timed on 4095-row slices scaled up), add+LayerNorm+quant, the stand-alone quantizer.
CUDA events, L2 flushed between iterations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openai_whisper_compression_b200 import functional as F

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 384000
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=7):
    for _ in range(2):
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
    return sorted(ts)[len(ts) // 2]


def report(name, t, nbytes):
    print(f"{name:44s} {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s ({nbytes / t / 1e9 / peaks['hbm_gbs']:.2f} of HBM)", flush=True)


torch.manual_seed(0)
x = (torch.randn(rows, 2048, device=dev) * 1.5).half()
t = timed(lambda: F.gelu_quant(x, 6.0))
report(f"gelu_quant {rows}x2048 (table)", t, rows * 2048 * 5 + 4 * rows)
t = timed(lambda: F.gelu_quant(x, 6.0, store_h=False))
report(f"gelu_quant {rows}x2048 (table, int8 rows only)", t, rows * 2048 * 3 + 4 * rows)
xs = x[:4095].contiguous()
t = timed(lambda: F.gelu_quant(xs, 6.0))
report("gelu_quant 4095x2048 (erff)", t, 4095 * 2048 * 5)
print(f"   -> erff kernel scaled to {rows} rows: {t * rows / 4095 * 1e6:.0f} us")
del x
x = (torch.randn(rows, 512, device=dev)).half()
d = (torch.randn(rows, 512, device=dev) * 0.3).half()
w = torch.ones(512, device=dev).half()
b = torch.zeros(512, device=dev).half()
t = timed(lambda: F.add_layernorm_quant(x, d, w, b, 1e-5, 6.0))
report(f"add_layernorm_quant {rows}x512 (with delta)", t, rows * 512 * (2 + 2 + 2 + 2 + 1) + 4 * rows)
t = timed(lambda: F.add_layernorm_quant(x, None, w, b, 1e-5, 6.0))
report(f"add_layernorm_quant {rows}x512 (no delta)", t, rows * 512 * (2 + 2 + 1) + 4 * rows)
t = timed(lambda: F.int8_vectorwise_quant(x, 6.0, finalize=False))
report(f"int8_vectorwise_quant {rows}x512", t, rows * 512 * 3 + 4 * rows)
t = timed(lambda: torch.clamp(x + d, min=-64504.0, max=64504.0))
report(f"torch add + clamp {rows}x512", t, rows * 512 * 10)

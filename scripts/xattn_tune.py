"""Decode-time cross-attention kernel alone: microseconds and GB/s per launch for the kernel variant selected by the
environment (WQ_XATTN=reg|tma, WQ_XATTN_STAGES, WQ_XATTN_CW, WQ_XATTN_PROMO), on the bench shapes (whisper-base: 8
heads; XT_HEADS / XT_ROWS / XT_BUFS select another geometry, S = 1500; every layer's K|V buffer in turn so that nothing stays in L2), with and without the int8 row
quantization of the output that the fused LLM.int8 step asks for."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200 import functional as F
H, S, D = int(os.environ.get("XT_HEADS", "8")), 1500, 64
d = H * D
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("WQ_XATTN"))
for B in [int(v) for v in os.environ.get("XT_ROWS", "256,64").split(",")]:
    kvs = [(torch.randn(B, S, 2 * d, device="cuda") * 0.5).half() for _ in range(int(os.environ.get("XT_BUFS", "6")))]
    q = torch.randn(B, d, device="cuda").half()
    for thr in (None, 6.0):
        ts = []
        for rep in range(4):
            for kv in kvs:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _, qx = F.cross_attn_decode(q, kv[:, :, :d], kv[:, :, d:], 0.125, H, thr)
                e1.record()
                if qx is not None and qx[2] is not None:
                    qx[2].col_flags.zero_()
                if rep:
                    ts.append((e0, e1))
        torch.cuda.synchronize()
        us = sum(a.elapsed_time(b) for a, b in ts) / len(ts) * 1e3
        gb = 2 * B * S * d * 2 / 1e9
        print(f"[{tag}] B={B} quant={thr is not None}: {us:.1f} us  {gb / us * 1e6:.0f} GB/s", flush=True)

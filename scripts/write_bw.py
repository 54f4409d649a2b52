"""Write-only / read-only / copy HBM bandwidth (CUDA events, buffers >> L2): context for store-bound GEMM shapes."""
import torch
dev = torch.device("cuda")
n = 1536 << 20
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
h = a.view(torch.float16)
def timed(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts)[len(ts) // 2]
t = timed(lambda: a.zero_());            print(f"memset 1.5 GiB        {t*1e6:8.1f} us  {n/t/1e9:7.0f} GB/s written")
t = timed(lambda: h.fill_(1.5));         print(f"fill fp16 1.5 GiB     {t*1e6:8.1f} us  {n/t/1e9:7.0f} GB/s written")
t = timed(lambda: b.copy_(a));           print(f"copy 1.5 GiB          {t*1e6:8.1f} us  {2*n/t/1e9:7.0f} GB/s read+written")
t = timed(lambda: h.sum());              print(f"sum fp16 1.5 GiB      {t*1e6:8.1f} us  {n/t/1e9:7.0f} GB/s read")

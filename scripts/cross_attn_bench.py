"""Own q_len=1 cross-attention kernel vs cuDNN SDPA on the decode workload (K|V column blocks of [B,S,2d])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as TF
from openai_whisper_compression_b200 import functional as F
H, S, D = 8, 1500, 64
d = H * D
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, r
for B in (256, 64, 16):
    # 6 layers' buffers cycled so that nothing stays in L2 (as in the decode step)
    kvs = [(torch.randn(B, S, 2 * d, device="cuda") * 0.5).half() for _ in range(6)]
    q = torch.randn(B, d, device="cuda").half()
    i = [0]
    def cudnn():
        kv = kvs[i[0] % 6]; i[0] += 1
        k = kv[:, :, :d].view(B, S, H, D).transpose(1, 2); v = kv[:, :, d:].view(B, S, H, D).transpose(1, 2)
        return TF.scaled_dot_product_attention(q.view(B, 1, H, D).transpose(1, 2), k, v, scale=0.125).transpose(1, 2).reshape(B, d)
    def own():
        kv = kvs[i[0] % 6]; i[0] += 1
        return F.cross_attn_decode(q, kv[:, :, :d], kv[:, :, d:], 0.125, H)[0]
    i[0] = 0; us_c, oc = t(cudnn)
    i[0] = 0; us_o, oo = t(own)
    gb = 2 * B * S * d * 2 / 1e9
    ref = torch.softmax((q.float() * 0.125).view(B, H, 1, D) @ kvs[(i[0] - 1) % 6][:, :, :d].float().view(B, S, H, D).permute(0, 2, 3, 1), -1) \
        @ kvs[(i[0] - 1) % 6][:, :, d:].float().view(B, S, H, D).transpose(1, 2)
    err = (oo.float() - ref.transpose(1, 2).reshape(B, d)).abs().max().item()
    print(f"B={B}: cuDNN {us_c:.1f} us ({gb/us_c*1e6:.0f} GB/s)   own {us_o:.1f} us ({gb/us_o*1e6:.0f} GB/s)   own max err vs fp32 {err:.2e}")

# exposed (not back-to-back) cost: a small dependent kernel between calls, as inside the decode graph
B = 256
kvs = [(torch.randn(B, S, 2 * d, device="cuda") * 0.5).half() for _ in range(6)]
q = torch.randn(B, d, device="cuda").half()
w = torch.randn(d, d, device="cuda").half()
def chain(attn):
    x = q
    for kv in kvs:
        a = attn(x, kv)
        x = a @ w            # dependent GEMM, like out_proj
    return x
def cudnn2(x, kv):
    k = kv[:, :, :d].view(B, S, H, D).transpose(1, 2); v = kv[:, :, d:].view(B, S, H, D).transpose(1, 2)
    return TF.scaled_dot_product_attention(x.view(B, 1, H, D).transpose(1, 2), k, v, scale=0.125).transpose(1, 2).reshape(B, d)
def own2(x, kv):
    return F.cross_attn_decode(x, kv[:, :, :d], kv[:, :, d:], 0.125, H)[0]
g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
for fn, g, name in ((cudnn2, g1, "cuDNN"), (own2, g2, "own")):
    s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        chain(fn)
    torch.cuda.current_stream().wait_stream(s_); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        chain(fn)
    us, _ = t(lambda: g.replay(), 20)
    print(f"graph of 6 x (attention -> dependent GEMM), B=256: {name} {us:.1f} us")

"""Does SDPA need the contiguous [B,H,S,D] copies HF makes?  Times encoder (S=1500) and decode (q_len=1)
attention with head-major contiguous vs projection-layout ([B,S,H,D] strided) operands, checks equality."""
import torch, torch.nn.functional as TF
B, H, S, D = int(__import__("sys").argv[1]) if len(__import__("sys").argv) > 1 else 256, 8, 1500, 64
dev = "cuda"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, r
torch.manual_seed(0)
qs, ks, vs = (torch.randn(B, S, H, D, device=dev, dtype=torch.half) for _ in range(3))
qc, kc, vc = (x.transpose(1, 2).contiguous() for x in (qs, ks, vs))
us_c, oc = t(lambda: TF.scaled_dot_product_attention(qc, kc, vc, scale=1.0))
us_s, os_ = t(lambda: TF.scaled_dot_product_attention(qs.transpose(1, 2), ks.transpose(1, 2), vs.transpose(1, 2), scale=1.0))
print(f"encoder B={B}: contiguous {us_c:.0f} us, strided {us_s:.0f} us, equal={torch.equal(oc, os_)}, "
      f"out strides contiguous-op {oc.stride()} strided-op {os_.stride()}")
q1 = torch.randn(B, H, 1, D, device=dev, dtype=torch.half)
q1s = q1.transpose(1, 2).contiguous().transpose(1, 2)
us_c, oc = t(lambda: TF.scaled_dot_product_attention(q1, kc, vc, scale=1.0), 20)
us_s, os_ = t(lambda: TF.scaled_dot_product_attention(q1s, ks.transpose(1, 2), vs.transpose(1, 2), scale=1.0), 20)
gb = 2 * B * H * S * D * 2 / 1e9
print(f"decode cross B={B}: contiguous KV {us_c:.0f} us ({gb/us_c*1e6:.0f} GB/s), strided KV {us_s:.0f} us ({gb/us_s*1e6:.0f} GB/s), "
      f"equal={torch.equal(oc, os_)} maxdiff={(oc.float()-os_.float()).abs().max().item():.3e}")
for chunk in (32, 64, 128):
    def chunked():
        outs = [TF.scaled_dot_product_attention(qs[i:i + chunk].transpose(1, 2), ks[i:i + chunk].transpose(1, 2),
                                                vs[i:i + chunk].transpose(1, 2), scale=1.0) for i in range(0, B, chunk)]
        return outs
    us, _ = t(chunked)
    print(f"encoder B={B} strided in chunks of {chunk}: {us:.0f} us")
o_pre = torch.empty(B, S, H, D, device=dev, dtype=torch.half)
# cross-attention K/V of all decoder layers as column blocks of ONE fused projection [B, S, L*2*d]
L = 6
big = torch.randn(B, S, 2 * L * H * D, device=dev, dtype=torch.half)
kb = big[:, :, 0:H * D].view(B, S, H, D).transpose(1, 2)
vb = big[:, :, H * D:2 * H * D].view(B, S, H, D).transpose(1, 2)
us_b, _ = t(lambda: TF.scaled_dot_product_attention(q1s, kb, vb, scale=1.0), 20)
print(f"decode cross B={B}: K/V as column blocks of a [B,S,{2*L*H*D}] buffer: {us_b:.0f} us ({gb/us_b*1e6:.0f} GB/s)")

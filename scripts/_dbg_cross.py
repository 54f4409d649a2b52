import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200 import functional as F
B, H, S = (int(a) for a in sys.argv[1:4])
d = H * 64
q = torch.randn(B, d, device="cuda").half()
kv = torch.randn(B, S, 2 * d, device="cuda").half()
torch.cuda.synchronize()
print("launch", B, H, S, flush=True)
out, _ = F.cross_attn_decode(q, kv[:, :, :d], kv[:, :, d:], 0.125, H)
torch.cuda.synchronize()
print("done", out.float().abs().max().item(), flush=True)

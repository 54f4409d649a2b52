"""Per-layer comparison of the drop-in modules against torch-op emulations on captured inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch import nn
from openai_whisper_compression_b200 import harness, bnb, dynamic
from tests import emulation as emu

MICRO = dict(encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             d_model=64, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=100)
g = torch.Generator().manual_seed(3)
feats = (torch.randn(4, 80, 200, generator=g) * 0.5)

ours = harness.apply_scheme(harness.build_model("tiny", **MICRO), "llm_int8", "cuda")
ref = harness.build_model("tiny", **MICRO).half()
emu.swap_all(ref, lambda m: emu.EmuLinear8bitLt(m.cuda(), 6.0))
ref = ref.cuda()
emus = dict(ref.named_modules())
ids = torch.full((4, 6), 50257, device="cuda")
def hook(name):
    def f(mod, inp, out):
        x = inp[0]
        e = emus[name](x)
        d = (out.float() - e.float()).abs().max().item()
        print(f"{name:50s} x{tuple(x.shape)} contig={x.is_contiguous()} absmax_x={x.abs().max().item():.3f} diff={d:.3e}")
    return f
for n, m in ours.named_modules():
    if isinstance(m, bnb.Linear8bitLt):
        m.register_forward_hook(hook(n))
with torch.no_grad():
    ours(input_features=feats.half().cuda(), decoder_input_ids=ids)

print("---- dynamic int8 twin vs torch CPU quantize_dynamic ----")
base = harness.build_model("tiny", **MICRO)
cpu = harness.build_model("tiny", **MICRO)
torch.quantization.quantize_dynamic(cpu, {nn.Linear}, dtype=torch.qint8, inplace=True)
twin = harness.apply_scheme(harness.build_model("tiny", **MICRO), "dynamic_int8", "cuda")
cpus = dict(cpu.named_modules())
def hook2(name):
    def f(mod, inp, out):
        x = inp[0]
        e = cpus[name](x.float().cpu())
        d = (out.float().cpu() - e).abs().max().item()
        print(f"{name:50s} x{tuple(x.shape)} diff={d:.3e} outmax={e.abs().max().item():.3f}")
    return f
for n, m in twin.named_modules():
    if isinstance(m, dynamic.DynamicInt8Linear):
        m.register_forward_hook(hook2(n))
with torch.no_grad():
    lo = twin(input_features=feats.cuda(), decoder_input_ids=ids).logits
    lc = cpu(input_features=feats, decoder_input_ids=ids.cpu()).logits
print("logits diff", (lo.cpu() - lc).abs().max().item())

"""Stand-alone timing of the log-mel kernel (64 x 30 s utterances, L2 flushed)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openai_whisper_compression_b200.frontend import LogMelFrontend

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
for mels in (80, 128):
    fe = LogMelFrontend(mels, device="cuda")
    audio = torch.randn(B, 480000, device="cuda") * 0.1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fe.features_from_device_audio(audio)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fe.features_from_device_audio(audio)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    nbytes = B * (480000 * 4 + mels * 3000 * 4)
    print(f"logmel B={B} mels={mels}: {t * 1e6:.1f} us  {nbytes / t / 1e9:.0f} GB/s ({nbytes / t / 1e9 / peaks['hbm_gbs']:.3f} of HBM)  "
          f"{t / B * 1e6:.2f} us/utterance", flush=True)

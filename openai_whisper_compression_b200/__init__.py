"""B200-native (sm_100a) implementation of the compressed-Whisper hot path of
juligoat/openai-whisper-compression: drop-in quantized / pruned linear layers and the log-mel
frontend, backed by hand-written CUDA behind a C ABI (include/whisperq.h).  No CPU fallback."""

__version__ = "0.1.0"

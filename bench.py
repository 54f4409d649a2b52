#!/usr/bin/env python
"""bench.py -- audio-seconds/sec of the compressed-Whisper hot path on B200.

Workload (BASELINE.json configs[1]): whisper-base, bitsandbytes LLM.int8 (threshold 6.0, HF
load_in_8bit flow: fp16 model, proj_out kept fp16), log-mel + encoder + greedy decode of B x 30 s
synthetic utterances per GPU per step (random-init weights, seeded gaussian audio, T new tokens
with min = max because random weights never emit EOS).  One process per GPU; utterances are
sharded over ranks (weak scaling: B per GPU fixed); the only collective is the int64[4] WER/CER
tally all-reduce.

  python bench.py --gpus 1 --steps 5 --warmup 3              # our arm
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # reference CPU arm
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # N > 1

Prints ONE JSON line (rank 0).  `value`: inputs (raw audio) already resident in HBM; `e2e`: the
same step through the drop-in modules with HOST audio buffers (pinned H2D copy inside the timed
region, token ids read back D2H, transcripts decoded and tallied).  `roofline`: the encoder-shaped
LLM.int8 tcgen05 GEMM launches, timed live with CUDA events inside the device-timed steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AUDIO_SECONDS = 30.0
N_SAMPLES = 480000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="base")
    ap.add_argument("--scheme", default="llm_int8")
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step")
    ap.add_argument("--new-tokens", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=8,
                    help="utterances in the CPU baseline sample (the reference's own CPU batch size, BASELINE config 0)")
    ap.add_argument("--prune", type=float, default=0.0,
                    help="global L1 magnitude pruning amount applied before quantization (BASELINE config 4: 0.5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hf-loop", action="store_true",
                    help="keep HF's Python decode loop instead of the CUDA-graph replay loop (fastgen)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        under_load = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(under_load) if under_load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline: the reference's own CPU path (model_utils.py:131-134):
# torch.quantization.quantize_dynamic(model, {torch.nn.Linear}, dtype=torch.qint8, inplace=True)
# on HF Whisper, generate-only timing as data_utils.py:151-155 (features precomputed by the HF
# feature extractor, as the reference does in map_to_feats).
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(size: str, n_utts: int, new_tokens: int, steps: int, warmup: int):
    import numpy as np
    import torch
    from transformers import WhisperFeatureExtractor
    from openai_whisper_compression_b200 import harness

    model = harness.build_model(size)
    torch.quantization.quantize_dynamic(model, {torch.nn.Linear}, dtype=torch.qint8, inplace=True)
    model.eval()
    fe = WhisperFeatureExtractor(feature_size=harness.WHISPER_SIZES[size]["mels"])
    audio = [harness.synth_audio(i) for i in range(n_utts)]
    t0 = time.time()
    feats = torch.from_numpy(np.concatenate([fe(a, sampling_rate=16000, return_tensors="np").input_features
                                             for a in audio]))
    t_mel = time.time() - t0
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            ids = harness.greedy_generate(model, feats, new_tokens)
            dt = time.time() - t0
            if i >= warmup:
                times.append(dt)
    assert ids.shape[0] == n_utts
    per_step = sum(times) / len(times)
    return {"audio_s_per_s": n_utts * AUDIO_SECONDS / per_step, "ms_per_step": per_step * 1e3,
            "logmel_s": t_mel, "cores": torch.get_num_threads(), "engine": torch.backends.quantized.engine}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.size, args.cpu_sample, args.new_tokens, args.steps, args.warmup)
    sample = (f"{args.cpu_sample} x 30 s utterances, whisper-{args.size}, torch quantize_dynamic qint8 "
              f"({r['engine']}), greedy {args.new_tokens} new tokens, generate-only timer (log-mel by HF "
              f"extractor took {r['logmel_s']:.2f} s, untimed as in the reference)")
    line = {
        "impl": "reference", "metric": "audio-seconds/sec", "value": r["audio_s_per_s"], "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": dict(workload_config(args), decode_loop="HF _sample (the reference's model.generate), CPU"),
        "cpu_baseline": {"value": r["audio_s_per_s"], "unit": "audio-s/s", "cores": r["cores"],
                         "kind": "reference", "sample": sample},
        "e2e": {"value": r["audio_s_per_s"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    default = args.size == "base" and args.scheme == "llm_int8" and args.prune == 0
    what = "bitsandbytes LLM.int8 (threshold 6.0)" if args.scheme == "llm_int8" else args.scheme
    if args.prune > 0:
        what = f"{int(args.prune * 100)} % global-L1 pruned + {what}"
    return {"workload": f"whisper-{args.size} {what} log-mel + encoder + greedy "
                        f"decode, {args.batch} x 30 s synthetic utterances per GPU per step, "
                        f"{args.new_tokens} new tokens" + (" (BASELINE.json configs[1])" if default else ""),
            "size": args.size, "scheme": args.scheme, "utterances_per_gpu": args.batch,
            "new_tokens": args.new_tokens, "parallelism": f"utterance-sharded dp{args.gpus}",
            "decode_loop": "HF _sample (Python)" if getattr(args, "hf_loop", False) else
                           "model.generate -> CUDA-graph replay per token (fastgen), HF logits processors",
            "l2": "256 MiB write before every step (inside the timed bracket; < 0.1 % of a step); per-step "
                  "activations (>= 98 MB per encoder linear) exceed the 126 MB L2 as well"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from openai_whisper_compression_b200 import functional as F
    from openai_whisper_compression_b200 import harness, tally

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B, T, K, W = args.batch, args.new_tokens, args.steps, max(args.warmup, 0)
    n_mels = harness.WHISPER_SIZES[args.size]["mels"]
    master = harness.build_model(args.size)
    if args.prune > 0:
        harness.global_l1_prune(master, args.prune)
    model = harness.apply_scheme(master, args.scheme, dev)
    del master
    if not args.hf_loop:
        from openai_whisper_compression_b200 import fastgen
        fastgen.enable(model)
    proc = harness.StubProcessor(n_mels, device=dev)
    fe = proc.feature_extractor
    half = harness.model_dtype(model) == torch.float16

    # this rank's utterance shard of the global batch of world * B
    utts = [rank * B + i for i in range(B)]
    audio_host = torch.empty((B, N_SAMPLES), dtype=torch.float32).pin_memory()
    for j, u in enumerate(utts):
        audio_host[j] = torch.from_numpy(harness.synth_audio(u))
    audio_dev = audio_host.to(dev)
    refs = [harness.synth_reference(u) for u in utts]
    ids_host = torch.empty((B, T + 8), dtype=torch.int64).pin_memory()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def hot_path(audio):
        feats = fe.features_from_device_audio(audio)
        if half:
            feats = feats.half()
        return harness.greedy_generate(model, feats, T)

    def step_device():
        flush_buf.zero_()
        return hot_path(audio_dev)

    # e2e: every step's audio is copied from pinned host memory inside the timed region; the copy of
    # step i+1 runs on a side stream while step i computes (two device buffers)
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(audio_dev), torch.empty_like(audio_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0}

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[i % 2])                  # buffer no longer read by step i-2
            bufs[i % 2].copy_(audio_host, non_blocking=True)     # H2D of step i's inputs
            ready[i % 2].record(copy_stream)

    def e2e_begin():
        e2e_state["i"] = 0
        done[0].record(); done[1].record()
        issue_copy(0)

    def step_e2e():
        i = e2e_state["i"]
        e2e_state["i"] = i + 1
        flush_buf.zero_()
        issue_copy(i + 1)                                        # next step's inputs, overlapped
        torch.cuda.current_stream().wait_event(ready[i % 2])
        ids = hot_path(bufs[i % 2])
        done[i % 2].record()
        out = ids_host[:, :ids.shape[1]]
        out.copy_(ids, non_blocking=True)                        # D2H of the result
        torch.cuda.current_stream().synchronize()
        hyps = proc.batch_decode(out)
        t = tally.all_reduce_tally(tally.tally_on_device(refs, hyps, dev))
        return ids, t.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(W):
        step_device()
    e2e_begin()
    for _ in range(max(1, min(W, 2))):
        step_e2e()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident inputs: `value` (+ live GEMM timing for the roofline) ----
    F.STATS.reset()
    F.STATS.profile_min_rows = 1024
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push("wq_timed")          # ncu --nvtx --nvtx-include "wq_timed/" selects this region
    for _ in range(K):
        ids = step_device()
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = F.STATS.launches
    records = list(F.STATS.records)
    F.STATS.profile_min_rows = None

    # ---- host buffers through the drop-in API: `e2e` ----
    barrier()
    e0.record()
    e2e_begin()                                                  # first copy is inside the timed region
    for _ in range(K):
        ids, t = step_e2e()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    total_audio = world * B * AUDIO_SECONDS * K
    # roofline of the encoder-shaped LLM.int8 GEMM launches (algorithmic bytes)
    by_shape = {}
    for kind, M, N, Kd, s, e in records:
        by_shape.setdefault((kind, M, N, Kd), []).append(s.elapsed_time(e) * 1e-3)
    tot_bytes = tot_flops = tot_time = 0.0
    n_launch = 0
    shapes = []
    for (kind, M, N, Kd), ts in sorted(by_shape.items()):
        a_bytes = 1 if kind in ("llmint8", "dyn_i8") else 2
        w_bytes = {"llmint8": 1.0, "dyn_i8": 1.0, "w8a16": 1.0, "w4a16": 0.5 + 4.0 / 64, "u4a16": 0.5 + 8.0 / 128}[kind]
        o_bytes = 4 if kind == "dyn_i8" else 2
        nbytes = M * Kd * a_bytes + N * Kd * w_bytes + M * N * o_bytes + 4 * (M + N)
        flops = 2.0 * M * N * Kd
        avg = sum(ts) / len(ts)
        shapes.append({"kind": kind, "M": M, "N": N, "K": Kd, "launches": len(ts), "avg_us": avg * 1e6,
                       "GBps": nbytes / avg / 1e9, "TFLOPs": flops / avg / 1e12})
        tot_bytes += nbytes * len(ts)
        tot_flops += flops * len(ts)
        tot_time += sum(ts)
        n_launch += len(ts)
    roofline = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if n_launch and os.path.exists(tpath):
        # DRAM bytes per launch from the committed ncu --set full capture of this command's kernel
        # (same shapes): launch-weighted average, None when a shape was not captured
        per = json.load(open(tpath)).get("traffic_bytes_per_launch", {})
        tot_t = 0.0
        for sh in shapes:
            key = f"{sh['M']}x{sh['N']}x{sh['K']}"
            if key not in per:
                tot_t = None
                break
            tot_t += per[key] * sh["launches"]
        traffic = None if tot_t is None else tot_t / n_launch
    if n_launch:
        gbs = tot_bytes / tot_time / 1e9
        tfs = tot_flops / tot_time / 1e12
        int8_kind = all(k[0] in ("llmint8", "dyn_i8") for k in by_shape)
        # which roof binds: arithmetic intensity of the launches against the measured ridge
        # (int8 tensor peak taken as 2x the measured bf16 peak: same pipe, half the operand bytes)
        tensor_peak = peaks["bf16_tflops"] * (2.0 if int8_kind else 1.0)
        ridge = tensor_peak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        intensity = tot_flops / tot_bytes
        common = {"traffic": traffic,
                  "kernel": "k_gemm_tc (TMA + tcgen05 + TMEM, fused dequant epilogue; LLM.int8: s8xs8->s32), "
                            "encoder-shaped launches (M >= 1024), kinds: " + ",".join(sorted({k[0] for k in by_shape})),
                  "launches_timed": n_launch, "avg_launch_us": tot_time / n_launch * 1e6,
                  "algorithmic_bytes_per_launch": tot_bytes / n_launch,
                  "algorithmic_flops_per_launch": tot_flops / n_launch,
                  "arithmetic_intensity": intensity, "ridge": ridge,
                  "hbm_GBps": gbs, "hbm_frac": gbs / peaks["hbm_gbs"],
                  "tensor_TFLOPs": tfs, "tensor_frac_of_bf16_peak": tfs / peaks["bf16_tflops"],
                  "gemm_share_of_step": tot_time * 1e3 / ms_dev, "shapes": shapes}
        if intensity <= ridge:
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"],
                        "peak_source": peaks["source"] + " (burst copy bandwidth, kernel timed alone by events)"}
        else:
            roofline = {"bound": "tensor", "achieved": tfs, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": tfs / peaks["bf16_tflops"],
                        "peak_source": peaks["source"] + " (burst cuBLAS bf16 GEMM, kernel timed alone by events"
                                       + ("; int8 MMAs run at up to 2x this rate" if int8_kind else "") + ")"}
        roofline.update(common)

    line = {
        "metric": "audio-seconds/sec", "value": total_audio / (ms_dev * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": total_audio / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": B * N_SAMPLES * 4, "d2h_bytes_per_step": int(ids.shape[1]) * B * 8 + 32,
                "note": "per step: pinned-host audio H2D (prefetched one step ahead on a copy stream), log-mel, "
                        "model.generate, ids D2H, decode to text, WER/CER tally on the GPU (+ all-reduce)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "tally": {"WER": tally.rates(t)["WER"], "CER": tally.rates(t)["CER"], "ref_words": int(t[1])},
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.size, args.cpu_sample, T, steps=1, warmup=1)
        line["cpu_baseline"] = {
            "value": r["audio_s_per_s"], "unit": "audio-s/s", "cores": r["cores"], "kind": "reference",
            "sample": f"{args.cpu_sample} x 30 s utterances of the same workload on the host: whisper-{args.size} "
                      f"torch quantize_dynamic qint8 ({r['engine']}), greedy {T} new tokens, generate-only "
                      f"timer, 1 warm-up + 1 timed pass ({r['ms_per_step'] / 1e3:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

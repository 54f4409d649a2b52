# one short bench line per BASELINE.json config (our arm); the default workload is configs[1]
set -x
python bench.py --size small --scheme bnb_nf4 --batch 64 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/cfg3_small_nf4.json
python bench.py --size medium --scheme quanto_int8_fp16 --prune 0.5 --batch 64 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/cfg4_medium_pruned_quanto.json
python bench.py --size large-v3 --scheme quanto_int8_fp16 --batch 32 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/cfg5_largev3_quanto.json
python bench.py --size small --scheme quanto_int4 --batch 16 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/cfgx_small_quanto_int4_fp32.json
for f in gpurun_out/cfg*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d.get("roofline") or {}
print(sys.argv[1], round(d["value"]), "audio-s/s e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "roofline", round(r.get("frac",0),3), "TF", round(r.get("tensor_TFLOPs",0)))
PY
done

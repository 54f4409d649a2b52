#!/bin/bash
# fp32 decode attention (quanto / bnb *_32 flows): kernel tests, C4 / C5-quanto config parity, C4 bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_configs.py -x -q -k "fp32 or config4 or config5 or attn" > $O/r02q_pytest_fp32.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02q_pytest_fp32.log
timeout 900 python bench.py --size medium --scheme quanto_int8 --prune 0.5 --batch 64 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $O/r02q_bench_c4.json 2> $O/r02q_bench_c4.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02q_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('token_check'))
print(json.dumps(d['roofline'])[:1200])
P

#!/bin/bash
set -u
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-token-check"
for lv in 3 2 3 2; do
  WQ_GEMM_PAIR=$lv timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PAIR=$lv', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['roofline']['decode']['per_token_ms'],4))"
done

#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py tests/test_gpu_static.py tests/test_gpu_modules.py -x -q -k "fp32 or config4 or config5 or gemv or quanto or static or qint" > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02v_pytest.log
timeout 900 python bench.py --size medium --scheme quanto_int8 --prune 0.5 --batch 64 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $O/r02v_bench_c4.json 2> $O/r02v_bench_c4.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02v_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('token_check'))
print(json.dumps(d['roofline'].get('decode'))[:600])
P

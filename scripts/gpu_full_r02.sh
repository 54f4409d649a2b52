#!/bin/bash
# full GPU suite + bench lines (run on the GPU box through gpurun); names carry a tag
set -u
TAG=${1:-r02s}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - $TAG <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
print('roofline',d['roofline']['frac'],d['roofline']['kernel'][:60],'enc',d['roofline']['encoder_gemm'].get('hbm_frac'),d['roofline']['encoder_gemm'].get('avg_launch_us'))
print('token',d['token_check']['verdict'],d['clocks'])
for e in d.get('extra_configs',[]):
    print(e['name'][:60],e['ms_per_step'],e['value'],(e.get('token_check') or {}).get('verdict'),e['roofline'].get('frac'))
P

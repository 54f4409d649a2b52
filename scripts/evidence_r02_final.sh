#!/bin/bash
# Final round-2 evidence (run on the GPU box through gpurun): full GPU suite, bench lines of both arms, step profile,
# launch list of a timed step.  Every ncu capture follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02z_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r02z_bench.json 2> $O/r02z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02z_bench_reference.json 2> /dev/null; echo "reference rc=$?"
timeout 300 python scripts/step_profile.py 256 64 2>/dev/null | head -48 > $O/r02z_step_profile_b256.txt
BENCH="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-token-check"
timeout 300 $BENCH --batch 64 --new-tokens 8 > /dev/null 2>&1 && \
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r02z_launches_timed_region.csv $BENCH --batch 64 --new-tokens 8 > $O/r02z_ncu_launches.log 2>&1
python scripts/summarise_launches.py $O/r02z_launches_timed_region.csv > $O/r02z_launches_timed_region_summary.txt 2>&1; head -25 $O/r02z_launches_timed_region_summary.txt
# encoder-shaped GEMMs of one timed step (CTA pairs): ncu --set full, DRAM traffic per launch for roofline.encoder_gemm.traffic
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none -k regex:k_gemm_tc -c 31 -f -o $O/r02z_ncu_gemm $BENCH > $O/r02z_ncu_gemm.log 2>&1
ncu -i $O/r02z_ncu_gemm.ncu-rep --page raw --csv > $O/r02z_ncu_gemm_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/r02z_ncu_gemm_raw.csv $O/r02z_ncu_gemm_bench.txt "ncu --nvtx --nvtx-include wq_timed/ --set full -k regex:k_gemm_tc -c 31 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-token-check" --traffic-json $O/r02z_gemm_traffic.json --M 384000
rm -f $O/r02z_ncu_gemm.ncu-rep
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02z_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
print('roofline',d['roofline']['frac'],'enc hbm_frac',d['roofline']['encoder_gemm'].get('hbm_frac'),d['roofline']['encoder_gemm'].get('avg_launch_us'))
print('token',d['token_check']['verdict'],d['clocks'])
for e in d.get('extra_configs',[]):
    print(e['name'][:60],e['ms_per_step'],e['value'],(e.get('token_check') or {}).get('verdict'),e['roofline'].get('frac'))
print(open('gpurun_out/r02z_bench_reference.json').read()[:600])
P

#!/bin/bash
# Round-2 evidence: bench lines, launch list and ncu --set full captures (run on the GPU box through gpurun).
# Every ncu capture follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-token-check"
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02_bench.json 2> $O/r02_bench.err || echo "bench failed"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> /dev/null || echo "reference arm failed"
timeout 300 python scripts/step_profile.py 256 64 2>/dev/null | head -48 > $O/r02_step_profile_b256.txt
# launch list of one timed step at reduced size (a full step exceeds 10 min under ncu)
timeout 300 $BENCH --batch 64 --new-tokens 8 > /dev/null 2>&1 && \
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r02_launches_timed_region.csv $BENCH --batch 64 --new-tokens 8 > $O/r02_ncu_launches.log 2>&1
# dominant kernel: decode-time cross-attention (TMA ring), without and with the int8 tail
timeout 200 python scripts/xattn_tune.py > $O/r02_xattn_tune.txt 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_cross_attn_decode_tma -s 6 -c 1 -f -o $O/r02_ncu_xattn python scripts/xattn_tune.py > $O/r02_ncu_xattn.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:k_cross_attn_decode_tma -s 30 -c 1 -f -o $O/r02_ncu_xattn_quant python scripts/xattn_tune.py >> $O/r02_ncu_xattn.log 2>&1
# encoder-shaped GEMMs of one timed step + the projection GEMM with the fused arg-max + lean decode tiles
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none -k regex:k_gemm_tc -c 31 -f -o $O/r02_ncu_gemm $BENCH > $O/r02_ncu_gemm.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none --kernel-name-base demangled -k "regex:k_gemm_tc<.int.128, .int.3, .int.0, .int.0, .int.4" -c 1 -f -o $O/r02_ncu_proj $BENCH >> $O/r02_ncu_gemm.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none --kernel-name-base demangled -k "regex:k_gemm_tc<.int.64, .int.[0-9], .int.2, .int.0, .int.0, __half, .int.1, .int.0, .int.0, .int.1" -c 6 -f -o $O/r02_ncu_lean $BENCH >> $O/r02_ncu_gemm.log 2>&1
for r in r02_ncu_xattn r02_ncu_xattn_quant r02_ncu_gemm r02_ncu_proj r02_ncu_lean; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
[ -f $O/r02_ncu_xattn.ncu-rep ] && ncu -i $O/r02_ncu_xattn.ncu-rep --page source --csv > $O/r02_ncu_xattn_source.csv 2>/dev/null
rm -f $O/r02_ncu_gemm.ncu-rep $O/r02_ncu_lean.ncu-rep      # large; the raw CSV pages travel instead
ls -la $O | tail -30

#!/bin/bash
set -u
O=gpurun_out
timeout 500 python scripts/step_profile.py 64 64 medium quanto_int8 2>/dev/null | head -40 > $O/r02u_step_profile_c4.txt; head -30 $O/r02u_step_profile_c4.txt
# ncu: CTA-pair int8 GEMM at the fc2 shape of the bench and a large-v3 shape
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc -s 3 -c 1 -f -o $O/r02u_ncu_pair_fc2 python scripts/gemm_ws_bench.py 2 384000 512 2048 > $O/r02u_ncu_pair.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_gemm_tc -s 3 -c 1 -f -o $O/r02u_ncu_pair_lv3 python scripts/gemm_ws_bench.py 2 48000 1280 5120 >> $O/r02u_ncu_pair.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_gemm_tc -s 3 -c 1 -f -o $O/r02u_ncu_pair_fc1 python scripts/gemm_ws_bench.py 2 384000 2048 512 >> $O/r02u_ncu_pair.log 2>&1
for r in r02u_ncu_pair_fc2 r02u_ncu_pair_lv3 r02u_ncu_pair_fc1; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; done
ncu -i $O/r02u_ncu_pair_fc2.ncu-rep --page source --csv > $O/r02u_ncu_pair_fc2_source.csv 2>/dev/null
rm -f $O/r02u_ncu_pair_lv3.ncu-rep $O/r02u_ncu_pair_fc1.ncu-rep
python scripts/ncu_summary.py $O/r02u_ncu_pair_fc2_raw.csv $O/r02u_ncu_pair_fc2.txt "ncu --set full: gemm_ws_bench.py 2 384000 512 2048 (CTA-pair int8 GEMM, fc2 shape of the bench)"
python scripts/ncu_summary.py $O/r02u_ncu_pair_lv3_raw.csv $O/r02u_ncu_pair_lv3.txt "ncu --set full: gemm_ws_bench.py 2 48000 1280 5120 (CTA-pair int8 GEMM, large-v3 fc2)"
python scripts/ncu_summary.py $O/r02u_ncu_pair_fc1_raw.csv $O/r02u_ncu_pair_fc1.txt "ncu --set full: gemm_ws_bench.py 2 384000 2048 512 (weight-stationary CTA-pair int8 GEMM, fc1 shape of the bench)"
cat $O/r02u_ncu_pair_fc2.txt $O/r02u_ncu_pair_lv3.txt $O/r02u_ncu_pair_fc1.txt

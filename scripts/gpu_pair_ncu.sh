#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc -s 3 -c 1 -f -o $O/r02p_ncu_pair python scripts/gemm_ws_bench.py 2 384000 512 2048 > $O/r02p_ncu_pair.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02p_ncu_pair.log
ncu -i $O/r02p_ncu_pair.ncu-rep --page raw --csv > $O/r02p_ncu_pair_raw.csv 2>/dev/null
ncu -i $O/r02p_ncu_pair.ncu-rep --page source --csv > $O/r02p_ncu_pair_source.csv 2>/dev/null
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02p_ncu_pair_raw.csv')))
h=rows[0]; r=rows[2]
for k in ['Kernel Name','launch__grid_size','launch__cluster_dim_x','launch__cluster_max_active','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__cycles_active.avg','smsp__cycles_active.avg','dram__bytes_read.sum','dram__bytes_write.sum','launch__occupancy_cluster_gpu_pct','launch__occupancy_cluster_pct']:
    for i,n in enumerate(h):
        if n==k: print(k, r[i][:100])
for i,n in enumerate(h):
    if 'cluster' in n or 'tensor' in n: print(n, r[i][:80])
P

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pair.py -x -q > $O/r02t_pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02t_pytest_pair.log
for sch in w8a16 w4a16; do
  timeout 200 python scripts/gemm_bench.py $sch 10 > $O/r02t_gemm_${sch}_pair.txt 2>&1; echo "rc=$?"
  WQ_GEMM_PAIR=1 timeout 200 python scripts/gemm_bench.py $sch 10 > $O/r02t_gemm_${sch}_single.txt 2>&1
  paste -d'\n' $O/r02t_gemm_${sch}_pair.txt $O/r02t_gemm_${sch}_single.txt | head -14
done

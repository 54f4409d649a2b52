#!/bin/bash
# CTA-pair GEMM: parity tests, then A/B timing against the single-CTA tiles (run on the GPU box through gpurun).
set -u
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_pair.py -x -q > $O/r02p_pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02p_pytest_pair.log
timeout 200 python scripts/gemm_ws_bench.py 10 > $O/r02p_gemm_pair.txt 2>&1; echo "pair rc=$?"
WQ_GEMM_PAIR=0 timeout 200 python scripts/gemm_ws_bench.py 10 > $O/r02p_gemm_single.txt 2>&1; echo "single rc=$?"
for nk in "1280 1280" "5120 1280" "1280 5120" "3072 768" "768 3072"; do
  timeout 100 python scripts/gemm_ws_bench.py 10 48000 $nk >> $O/r02p_gemm_pair.txt 2>&1
  WQ_GEMM_PAIR=0 timeout 100 python scripts/gemm_ws_bench.py 10 48000 $nk >> $O/r02p_gemm_single.txt 2>&1
done
cat $O/r02p_gemm_pair.txt; cat $O/r02p_gemm_single.txt

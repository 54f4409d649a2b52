set -x
python scripts/gemm_bench.py llmint8 5 > gpurun_out/gemm_bench_llmint8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 1 -o gpurun_out/gemm_enc_512 -f python scripts/gemm_bench.py llmint8 1 > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/gemm_bench_llmint8.log
python scripts/gemm_bench.py w8a16 5; python scripts/gemm_bench.py w4a16 5; python scripts/gemm_bench.py quant 5

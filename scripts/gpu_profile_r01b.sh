# round-1 profile of the default bench command shape (reduced token count to bound ncu time)
ARGS="--steps 1 --warmup 3 --batch 64 --new-tokens 8 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r01b.log 2>&1 && \
ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python bench.py $ARGS > gpurun_out/ncu_r01b.log 2>&1
echo "launch list rc=$?"
python scripts/gemm_bench.py llmint8 1 > gpurun_out/gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 0 -c 1 -o gpurun_out/gemm_v3_enc512 -f python scripts/gemm_bench.py llmint8 1 > gpurun_out/ncu_gemm_v3.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/plain_r01b.log | cut -c1-600

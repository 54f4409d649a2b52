timeout 300 python -m pytest tests/test_gpu_modules.py -q -m gpu --timeout 300 -k dynamic_int8_twin 2>&1 | grep -E "^E|assert|Error" | head -20
ARGS="--steps 1 --warmup 3 --batch 64 --new-tokens 4 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r01.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py $ARGS > gpurun_out/ncu_r01.log 2>&1
echo "ncu launches rc=$?"
python bench.py $ARGS > gpurun_out/plain2_r01.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 150 -c 3 -o gpurun_out/gemm_r01 python bench.py $ARGS > gpurun_out/ncu2_r01.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/

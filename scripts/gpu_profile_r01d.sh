# round-1 closing profile of the default bench command (B = 256, T = 64): tests, bench lines (both arms),
# launch list of one timed step, ncu --set full of the encoder-shaped GEMMs of that step (DRAM traffic per shape)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_r01d_reference.json 2> gpurun_out/bench_r01d_reference.err; echo "ref rc=$?"
ARGS="--steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01d.csv python bench.py $ARGS > gpurun_out/ncu_r01d.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 30 -o gpurun_out/gemm_r01d_bench -f python bench.py $ARGS > gpurun_out/ncu_r01d_full.log 2>&1
echo "ncu full rc=$?"
python __graft_entry__.py smoke > gpurun_out/smoke_r01d.log 2>&1; echo "smoke rc=$?"

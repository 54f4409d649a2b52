# round-1 closing run: GPU tests, both bench arms, row-kernel timings, ncu --set full of the encoder-shaped GEMMs of
# one timed bench step (raw CSV only: a report with source for 30 launches exceeds gpurun's 64 MiB return limit),
# launch list of a timed step at reduced size (a full B = 256, T = 64 step needs > 10 min under ncu).
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_r01e.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_r01e_reference.json 2> gpurun_out/bench_r01e_reference.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/bench_r01e_reference.json
timeout 120 python scripts/rowops_bench.py > gpurun_out/rowops_r01e.log 2>&1; cat gpurun_out/rowops_r01e.log
ARGS="--steps 1 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none -k regex:k_gemm_tc -c 30 -o /tmp/gemm_r01e_bench -f python bench.py $ARGS > gpurun_out/ncu_r01e_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/gemm_r01e_bench.ncu-rep --page raw --csv > gpurun_out/gemm_r01e_bench_raw.csv 2>/dev/null; ls -la gpurun_out/gemm_r01e_bench_raw.csv
ARGS="--steps 1 --warmup 3 --batch 64 --new-tokens 8 --no-cpu-baseline"
timeout 300 ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01e.csv python bench.py $ARGS > gpurun_out/ncu_r01e.log 2>&1
echo "launch list rc=$?"
gzip -f gpurun_out/launches_r01e.csv
du -sh gpurun_out

# round-1 closing run (after the LayerNorm occupancy fix, table GELU and 128 x 256 GEMM tiles): GPU tests, both bench
# arms, ncu --set full of the encoder-shaped GEMMs of one timed bench step (raw CSV only), launch list of a timed step
# at reduced size (a full B = 256, T = 64 step needs > 10 min under ncu), kernel-time table of one full step.
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_r01f.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_r01f_reference.json 2> gpurun_out/bench_r01f_reference.err; echo "ref rc=$?"
timeout 120 python scripts/step_profile.py 256 64 2>&1 | grep -v "Both\|warn\|attention mask\|custom logits" > gpurun_out/step_profile_r01f.log; head -14 gpurun_out/step_profile_r01f.log | cut -c1-150
ARGS="--steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none -k regex:k_gemm_tc -c 30 -o /tmp/gemm_r01f_bench -f python bench.py $ARGS > gpurun_out/ncu_r01f_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/gemm_r01f_bench.ncu-rep --page raw --csv > gpurun_out/gemm_r01f_bench_raw.csv 2>/dev/null; ls -la gpurun_out/gemm_r01f_bench_raw.csv
ARGS="--steps 1 --warmup 3 --batch 64 --new-tokens 8 --no-cpu-baseline"
timeout 240 ncu --nvtx --nvtx-include "wq_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f.csv python bench.py $ARGS > gpurun_out/ncu_r01f.log 2>&1
echo "launch list rc=$?"
gzip -f gpurun_out/launches_r01f.csv
du -sh gpurun_out

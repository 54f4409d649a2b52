# ncu --set full of the dominant kernel inside the timed region of the default bench workload
ARGS="--steps 1 --warmup 3 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_r01c.log 2>&1 && \
ncu --nvtx --nvtx-include "wq_timed/" --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 6 -o gpurun_out/gemm_r01_bench -f python bench.py $ARGS > gpurun_out/ncu_r01c.log 2>&1
echo "rc=$?"

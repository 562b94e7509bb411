set -x
mkdir -p gpurun_out
timeout 240 python tests/gpu_kernel_check.py --only "fp8,linear_fwd speed" --out gpurun_out/r36_fp8_kernel_check.json 2>&1 | tail -n 4 | cut -c1-3000
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r36_bench_bf16.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r36_bench_bf16.log | cut -c1-1500
timeout 300 python bench.py --steps 10 --warmup 3 --fp8 1 --no-e2e > gpurun_out/r36_bench_fp8.log 2>&1; echo "bench fp8 rc=$?"; tail -n 1 gpurun_out/r36_bench_fp8.log | cut -c1-1500
LIBAI_B200_NVTX=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_step/" --csv --log-file gpurun_out/launches_r36.csv python bench.py --steps 1 --warmup 3 --no-e2e --graphs 0 > gpurun_out/bench_ncu36.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_r36.csv

set -x
mkdir -p gpurun_out
LIBAI_B200_NVTX=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_step" --csv --log-file gpurun_out/launches_r37.csv python bench.py --steps 1 --warmup 3 --no-e2e --graphs 0 > gpurun_out/bench_ncu37.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_r37.csv

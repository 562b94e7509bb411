# round 2, run 33 (1 GPU): per-kernel breakdown of one step of the final tree
set -x
mkdir -p gpurun_out
LIBAI_B200_NVTX=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_step" --csv --log-file gpurun_out/r2_33_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --graphs 0 --ref-same-box 0 > gpurun_out/r2_33_ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2_33_launches.csv

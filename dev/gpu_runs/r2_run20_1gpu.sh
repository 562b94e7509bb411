# round 2, run 20 (1 GPU): bench with the pipelined attention backward; CTA-pair GEMMs on/off in the same box; per-kernel
# breakdown of one step (ncu launch list inside the NVTX range of the timed step)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 > gpurun_out/r2_20_bench_1gpu.json 2> gpurun_out/r2_20_bench.err; tail -1 gpurun_out/r2_20_bench_1gpu.json | cut -c1-700
LIBAI_B200_GEMM_2CTA=0 timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_20_bench_1gpu_no_cta_pairs.json 2>> gpurun_out/r2_20_bench.err; tail -1 gpurun_out/r2_20_bench_1gpu_no_cta_pairs.json | cut -c1-400
LIBAI_B200_ATTN_BWD_PIPE=0 timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_20_bench_1gpu_sequential_attn_bwd.json 2>> gpurun_out/r2_20_bench.err; tail -1 gpurun_out/r2_20_bench_1gpu_sequential_attn_bwd.json | cut -c1-400
LIBAI_B200_NVTX=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bench_step" --csv --log-file gpurun_out/r2_20_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --graphs 0 --ref-same-box 0 > gpurun_out/r2_20_ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2_20_launches.csv

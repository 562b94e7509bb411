# round 2, run 31 (2 GPUs): final-tree check of the driver's N=2 command, and the extras watchdog with an injected failure
# (must print the JSON line within seconds and exit 0)
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
t0=$(date +%s)
LIBAI_B200_BENCH_INJECT_EXTRA_FAIL=1 run 600 29561 bench.py --gpus 2 --steps 5 --warmup 3 --ref-same-box 0 --no-e2e > gpurun_out/r2_31_bench_2gpu_injected_extra_failure.json 2> gpurun_out/r2_31_inject.err
echo "injected-failure rc=$? seconds=$(( $(date +%s) - t0 ))"; tail -1 gpurun_out/r2_31_bench_2gpu_injected_extra_failure.json | cut -c1-200; grep -o '"layouts": {.*' gpurun_out/r2_31_bench_2gpu_injected_extra_failure.json | cut -c1-500
run 900 29562 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2_31_bench_2gpu.json 2> gpurun_out/r2_31_bench_2gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r2_31_bench_2gpu.json | cut -c1-2500
run 400 29563 tests/gpu_comm_check.py --out gpurun_out/r2_31_comm_check_2gpu.json > gpurun_out/r2_31_comm_check_2gpu.log 2>&1
grep '"ok": false\|ZeRO fused' gpurun_out/r2_31_comm_check_2gpu.log | cut -c1-700; tail -1 gpurun_out/r2_31_comm_check_2gpu.log

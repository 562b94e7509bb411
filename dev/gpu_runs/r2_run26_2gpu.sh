# round 2, run 26 (2 GPUs): NVLS (multimem.ld_reduce / multimem.st) in the ZeRO-1 kernels: comm check + dp2 bench with and
# without (LIBAI_B200_NVLS=0), NCCL's own view of NVLS on this box
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
NCCL_DEBUG=INFO run 400 29541 tests/gpu_comm_check.py --out gpurun_out/r2_26_comm_check_2gpu_nvls.json > gpurun_out/r2_26_comm_check_2gpu_nvls.log 2>&1
echo "comm rc=$?"; grep -i "nvls" gpurun_out/r2_26_comm_check_2gpu_nvls.log | head -6 | cut -c1-250; grep '"ok": false\|ZeRO fused' gpurun_out/r2_26_comm_check_2gpu_nvls.log | cut -c1-600; tail -1 gpurun_out/r2_26_comm_check_2gpu_nvls.log | cut -c1-300
run 600 29542 bench.py --gpus 2 --steps 10 --warmup 4 --extras 0 --ref-same-box 0 --no-e2e > gpurun_out/r2_26_bench_2gpu_nvls.json 2> gpurun_out/r2_26_bench_nvls.err
echo "bench nvls rc=$?"; tail -1 gpurun_out/r2_26_bench_2gpu_nvls.json | cut -c1-500; grep -i "nvls\|multicast" gpurun_out/r2_26_bench_nvls.err | head -3 | cut -c1-250
LIBAI_B200_NVLS=0 run 600 29543 bench.py --gpus 2 --steps 10 --warmup 4 --extras 0 --ref-same-box 0 --no-e2e > gpurun_out/r2_26_bench_2gpu_no_nvls.json 2> gpurun_out/r2_26_bench_no_nvls.err
echo "bench no-nvls rc=$?"; tail -1 gpurun_out/r2_26_bench_2gpu_no_nvls.json | cut -c1-500

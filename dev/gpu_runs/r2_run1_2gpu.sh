# round 2, run 1 (2 GPUs): new device-side-state comm kernels — numerics/timing check, TP parity vs 1 GPU, bench with layouts
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_2gpu.json > gpurun_out/r2_comm_check_2gpu.log 2>&1
tail -30 gpurun_out/r2_comm_check_2gpu.log
timeout 1200 python -m pytest tests/test_gpu.py -q -x -k "fused_tensor_parallel" > gpurun_out/r2_tp_parity.log 2>&1
tail -30 gpurun_out/r2_tp_parity.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -5 gpurun_out/r2_bench_2gpu.err; cat gpurun_out/r2_bench_2gpu.json

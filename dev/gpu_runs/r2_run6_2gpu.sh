# round 2, run 6 (2 GPUs): pull-based GEMM->RS (local epilogue + owners pull), deeper store pipeline in the AG copy ring
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_2gpu_v3.json > gpurun_out/r2_comm_check_2gpu_v3.log 2>&1
grep '"ok": false' gpurun_out/r2_comm_check_2gpu_v3.log | cut -c1-800; tail -1 gpurun_out/r2_comm_check_2gpu_v3.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29521 \
    tests/gpu_comm_bench.py --out gpurun_out/r2_comm_bench_2gpu_v3.json 2>&1 | tail -2 | cut -c1-2500
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 4 --ref-same-box 0 > gpurun_out/r2_bench_2gpu_v3.json 2> gpurun_out/r2_bench_2gpu_v3.err
tail -3 gpurun_out/r2_bench_2gpu_v3.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_v3.json | cut -c1-2500

# round 2, run 12 (4 GPUs): the driver's N=4 command (dp4 + tp2xdp2 + tp2xpp2 layouts, PyTorch comparator) and the fused
# collectives at 4 ranks
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_4gpu.json > gpurun_out/r2_comm_check_4gpu.log 2>&1
grep '"ok": false' gpurun_out/r2_comm_check_4gpu.log | cut -c1-800; tail -1 gpurun_out/r2_comm_check_4gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 10 --warmup 4 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
tail -3 gpurun_out/r2_bench_4gpu.err | cut -c1-400; cat gpurun_out/r2_bench_4gpu.json | cut -c1-3000

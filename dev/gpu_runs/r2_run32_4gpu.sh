# round 2, run 32 (4 GPUs): final-tree check of the driver's N=4 command (dp4 + tp2xdp2 + tp2xpp2 + comparator)
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=30000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 4 > gpurun_out/r2_32_bench_4gpu.json 2> gpurun_out/r2_32_bench_4gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r2_32_bench_4gpu.json | cut -c1-3000

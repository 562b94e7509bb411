# round 2, run 5 (2 GPUs): unrolled RS reduce phase, lighter signalling, pre-copy instead of fill_local, multi-ring copy CTAs
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_2gpu_v2.json > gpurun_out/r2_comm_check_2gpu_v2.log 2>&1
grep -c '"ok": true' gpurun_out/r2_comm_check_2gpu_v2.log; grep '"ok": false' gpurun_out/r2_comm_check_2gpu_v2.log | cut -c1-600; tail -1 gpurun_out/r2_comm_check_2gpu_v2.log
for rings in 1 2 4; do
LIBAI_B200_COPY_RINGS=$rings timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2952$rings \
    tests/gpu_comm_bench.py --out gpurun_out/r2_comm_bench_2gpu_rings$rings.json 2>&1 | tail -2 | cut -c1-2500
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 4 --ref-same-box 0 > gpurun_out/r2_bench_2gpu_v2.json 2> gpurun_out/r2_bench_2gpu_v2.err
tail -3 gpurun_out/r2_bench_2gpu_v2.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_v2.json | cut -c1-2500

# round 2, run 7 (2 GPUs): tp2 with the forward's gathered input kept for the wgrad; pp2 (eager 1F1B) and the 1-GPU number on the same box
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_2gpu_v4.json > gpurun_out/r2_comm_check_2gpu_v4.log 2>&1
grep '"ok": false' gpurun_out/r2_comm_check_2gpu_v4.log | cut -c1-800; tail -1 gpurun_out/r2_comm_check_2gpu_v4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --layout tp2 --steps 10 --warmup 4 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_2gpu_tp2_v4.json 2> gpurun_out/r2_bench_2gpu_tp2_v4.err
tail -2 gpurun_out/r2_bench_2gpu_tp2_v4.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_tp2_v4.json | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --layout pp2 --steps 8 --warmup 3 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_2gpu_pp2_v4.json 2> gpurun_out/r2_bench_2gpu_pp2_v4.err
tail -2 gpurun_out/r2_bench_2gpu_pp2_v4.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_pp2_v4.json | cut -c1-1200
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 4 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_1gpu_v4.json 2> gpurun_out/r2_bench_1gpu_v4.err
cat gpurun_out/r2_bench_1gpu_v4.json | cut -c1-900

# round 2, run 14 (2 GPUs): CTA-pair tiles inside the fused-collective GEMMs (LIBAI_B200_GEMM_2CTA=2) — numerics, op timings, tp2 step
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
export LIBAI_B200_GEMM_2CTA=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/gpu_comm_check.py --out gpurun_out/r2_comm_check_2gpu_2cta.json > gpurun_out/r2_comm_check_2gpu_2cta.log 2>&1
grep '"ok": false' gpurun_out/r2_comm_check_2gpu_2cta.log | cut -c1-800; tail -1 gpurun_out/r2_comm_check_2gpu_2cta.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29521 \
    tests/gpu_comm_bench.py --out gpurun_out/r2_comm_bench_2gpu_2cta.json 2>&1 | tail -2 | cut -c1-2500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --layout tp2 --steps 10 --warmup 4 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_2gpu_tp2_2cta2.json 2> gpurun_out/r2_bench_2gpu_tp2_2cta2.err
tail -2 gpurun_out/r2_bench_2gpu_tp2_2cta2.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_tp2_2cta2.json | cut -c1-700
export LIBAI_B200_GEMM_2CTA=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --layout tp2 --steps 10 --warmup 4 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_2gpu_tp2_2cta1.json 2> gpurun_out/r2_bench_2gpu_tp2_2cta1.err
cat gpurun_out/r2_bench_2gpu_tp2_2cta1.json | cut -c1-700

# round 2, run 4 (2 GPUs): steady-state (back-to-back) cost of every fused comm+GEMM op at the TP2 shapes, with and
# without the NVLink payload (handshake-only timing experiment)
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29516 \
    tests/gpu_comm_bench.py --out gpurun_out/r2_comm_bench_2gpu.json 2>&1 | tail -3 | cut -c1-3000
LIBAI_B200_DEBUG_COMM_NO_PAYLOAD=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 \
    tests/gpu_comm_bench.py --out gpurun_out/r2_comm_bench_2gpu_nopayload.json 2>&1 | tail -3 | cut -c1-3000

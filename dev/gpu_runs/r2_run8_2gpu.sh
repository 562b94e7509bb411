# round 2, run 8 (2 GPUs): pipeline stages with multi-slot CUDA graphs (parity + bench), compute-sanitizer passes
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 600 python tests/gpu_tp_parity.py --out gpurun_out/r2_parity_1gpu.json 2>&1 | tail -1 | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/gpu_tp_parity.py --pp 2 --out gpurun_out/r2_parity_pp2_graphs.json 2>&1 | tail -3 | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --layout pp2 --steps 8 --warmup 3 --ref-same-box 0 --no-e2e > gpurun_out/r2_bench_2gpu_pp2_graphs.json 2> gpurun_out/r2_bench_2gpu_pp2_graphs.err
tail -3 gpurun_out/r2_bench_2gpu_pp2_graphs.err | cut -c1-400; cat gpurun_out/r2_bench_2gpu_pp2_graphs.json | cut -c1-1000
# sanitizers on one GPU (slow: restricted to the small correctness cases)
CUDA_VISIBLE_DEVICES=0 bash dev/sanitize.sh memcheck "gemm L0 256x512x256 bn256,gemm L2 256x512x256 bn128,attention B2 A4 S256 D64 causal=True,attention dropout,bias+dropout,norm rms=False H=1024,cross entropy,fused adamw" 2>&1 | tail -5
CUDA_VISIBLE_DEVICES=0 bash dev/sanitize.sh racecheck "gemm L0 256x512x256 bn256,attention B2 A4 S256 D64 causal=True,attention dropout" 2>&1 | tail -5
CUDA_VISIBLE_DEVICES=0 bash dev/sanitize.sh synccheck "gemm L0 256x512x256 bn256,attention B2 A4 S256 D64 causal=True" 2>&1 | tail -5

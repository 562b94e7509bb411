# round 2, run 11 (2 GPUs): cta_group::2 GEMM after removing the per-arrival cluster fence; ZeRO-2 / ZeRO-3 on dp2 (NCCL buckets)
set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 LIBAI_B200_GEMM_2CTA=1 timeout 900 python tests/gpu_kernel_check.py --only "gemm" --out gpurun_out/r2_kernel_check_gemm_2cta_v2.json > gpurun_out/r2_kernel_check_gemm_2cta_v2.log 2>&1
tail -1 gpurun_out/r2_kernel_check_gemm_2cta_v2.log | cut -c1-300
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2_kernel_check_gemm_2cta_v2.json')):
    if 'ms' in r: print(r['name'], 'ok', r['ok'], 'ms %.4f tflops %.0f cublas %.4f' % (r['ms'], r.get('tflops', 0), r.get('cublas_ms', 0)))
PY
for z in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2953$z \
    tests/gpu_tp_parity.py --zero $z --out gpurun_out/r2_parity_dp2_zero$z.json 2>&1 | tail -1 | cut -c1-600
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --zero 2 --steps 8 --warmup 3 --ref-same-box 0 --no-e2e --extras 0 > gpurun_out/r2_bench_2gpu_zero2.json 2> gpurun_out/r2_bench_2gpu_zero2.err
tail -2 gpurun_out/r2_bench_2gpu_zero2.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_zero2.json | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29515 \
    bench.py --gpus 2 --zero 3 --steps 8 --warmup 3 --ref-same-box 0 --no-e2e --extras 0 > gpurun_out/r2_bench_2gpu_zero3.json 2> gpurun_out/r2_bench_2gpu_zero3.err
tail -2 gpurun_out/r2_bench_2gpu_zero3.err | cut -c1-300; cat gpurun_out/r2_bench_2gpu_zero3.json | cut -c1-900

# round 2, run 25 (1 GPU): forward attention with early Q.K^T issue / per-warp arrivals / pairwise barriers; ViT-L with
# graphed blocks (nn.Sequential containers); full kernel check for the bias / dropout variants
set -x
mkdir -p gpurun_out
timeout 200 python dev/attn_dev.py > gpurun_out/r2_25_attn.json 2> gpurun_out/r2_25_attn.err; echo "attn rc=$?"; tail -1 gpurun_out/r2_25_attn.json | cut -c1-1500
timeout 600 python tests/gpu_kernel_check.py --quick --out gpurun_out/r2_25_kernel_check.json > gpurun_out/r2_25_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -1 gpurun_out/r2_25_kernel_check.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 4 --ref-same-box 0 --no-e2e --model vit_l > gpurun_out/r2_25_vit_l.json 2> gpurun_out/r2_25_vit_l.err; echo "vit rc=$?"; tail -1 gpurun_out/r2_25_vit_l.json | cut -c1-600; grep -i "cuda graphs" gpurun_out/r2_25_vit_l.err | tail -2 | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --ref-same-box 0 --no-e2e > gpurun_out/r2_25_bench_1gpu.json 2> gpurun_out/r2_25_bench.err; tail -1 gpurun_out/r2_25_bench_1gpu.json | cut -c1-400

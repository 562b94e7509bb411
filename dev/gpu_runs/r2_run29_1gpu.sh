# round 2, run 29 (1 GPU): branch-free masking in the attention kernels
set -x
mkdir -p gpurun_out
timeout 200 python dev/attn_dev.py > gpurun_out/r2_29_attn.json 2> gpurun_out/r2_29_attn.err; echo "attn rc=$?"; tail -1 gpurun_out/r2_29_attn.json | cut -c1-1500
timeout 600 python tests/gpu_kernel_check.py --quick --out gpurun_out/r2_29_kernel_check.json > gpurun_out/r2_29_kernel_check.log 2>&1; echo "kernel check rc=$?"; tail -1 gpurun_out/r2_29_kernel_check.log | cut -c1-300

# round 2, run 37 (1 GPU): gate/up projection as one GEMM — parity with the two-GEMM path and timing
set -x
mkdir -p gpurun_out
timeout 150 python tests/gpu_gated_mlp_check.py > gpurun_out/r2_37_gated_mlp_1gpu.json 2> gpurun_out/r2_37_gated_mlp_1gpu.err; echo "rc=$?"; tail -1 gpurun_out/r2_37_gated_mlp_1gpu.json | cut -c1-600; tail -3 gpurun_out/r2_37_gated_mlp_1gpu.err | cut -c1-300

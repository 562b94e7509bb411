# round 2, run 30 (1 GPU): tile-N sweep of the training GEMMs with and without CTA pairs
set -x
mkdir -p gpurun_out
timeout 200 python dev/gemm_sweep.py > gpurun_out/r2_30_gemm_sweep_2cta.json 2> gpurun_out/r2_30_a.err; echo "rc=$?"; tail -1 gpurun_out/r2_30_gemm_sweep_2cta.json
LIBAI_B200_GEMM_2CTA=0 timeout 200 python dev/gemm_sweep.py > gpurun_out/r2_30_gemm_sweep_1cta.json 2> gpurun_out/r2_30_b.err; echo "rc=$?"; tail -1 gpurun_out/r2_30_gemm_sweep_1cta.json

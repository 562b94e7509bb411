# round 2, run 35 (1 GPU): compute-sanitizer memcheck + racecheck over the kernels written or changed late in round 2
# (pipelined attention backward incl. bias / ALiBi / dropout variants, early-issue forward, LayerNorm bwd with direct
# main_grad accumulation, fused bias-grad dgrad epilogue)
set -x
mkdir -p gpurun_out
ONLY="attention,norm rms=False H=1024,accumulating colsum,fused bias grad,mlp fused"
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tests/gpu_kernel_check.py --quick --only "$ONLY" --out gpurun_out/r2_35_sanitize_${tool}.json > gpurun_out/r2_35_sanitize_${tool}.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|Hazard|SUMMARY:" gpurun_out/r2_35_sanitize_${tool}.log | head -12
done

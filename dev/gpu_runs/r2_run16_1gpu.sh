# round 2, run 16 (1 GPU): pipelined attention backward (LIBAI_B200_ATTN_BWD_PIPE = 0 sequential / 1 pipelined /
# 2 pipelined + setmaxnreg): numerics vs fp32 and timings vs SDPA, each under its own timeout (a deadlock must not eat
# the budget), then the step breakdown of the default build
set -x
mkdir -p gpurun_out
for v in 0 1 2; do
  LIBAI_B200_ATTN_BWD_PIPE=$v timeout 240 python dev/attn_dev.py > gpurun_out/r2_16_attn_pipe$v.json 2> gpurun_out/r2_16_attn_pipe$v.err
  echo "pipe=$v rc=$?"; tail -1 gpurun_out/r2_16_attn_pipe$v.json | cut -c1-1200; tail -2 gpurun_out/r2_16_attn_pipe$v.err | cut -c1-300
done

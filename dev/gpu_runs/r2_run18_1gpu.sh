# round 2, run 18 (1 GPU): ncu --set full of the pipelined attention backward (source-level stall samples)
set -x
mkdir -p gpurun_out
LIBAI_B200_ATTN_BWD_PIPE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_pipe -s 1 -c 1 \
   -o gpurun_out/r2_18_attn_bwd_pipe -f python dev/attn_dev.py --ncu > gpurun_out/r2_18_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_18_ncu.log; ls -la gpurun_out/*.ncu-rep

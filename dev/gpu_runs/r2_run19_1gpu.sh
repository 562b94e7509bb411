# round 2, run 19 (1 GPU): pipelined attention backward v2 (no spills: split S/dP release, packed P, st.shared)
set -x
mkdir -p gpurun_out
r() { name=$1; shift; env "$@" timeout 150 python dev/attn_dev.py $EXTRA > gpurun_out/r2_19_$name.json 2> gpurun_out/r2_19_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_19_$name.json | cut -c1-1500; }
r pipe1 LIBAI_B200_ATTN_BWD_PIPE=1
EXTRA=--speed-only
r pipe1_main_only LIBAI_B200_ATTN_BWD_PIPE=1 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1
r pipe0_main_only LIBAI_B200_ATTN_BWD_PIPE=0 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1
LIBAI_B200_ATTN_BWD_PIPE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_pipe -s 1 -c 1 \
   -o gpurun_out/r2_19_attn_bwd_pipe -f python dev/attn_dev.py --ncu > gpurun_out/r2_19_ncu.log 2>&1
echo "ncu rc=$?"

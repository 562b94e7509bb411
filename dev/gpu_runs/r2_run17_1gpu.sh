# round 2, run 17 (1 GPU): where does the attention backward spend its time?  main kernel alone (helpers skipped), and with
# the dQ global reductions switched off (timing experiments: LIBAI_B200_ATTN_DEBUG_*), sequential vs pipelined kernel
set -x
mkdir -p gpurun_out
r() { name=$1; shift; env "$@" timeout 120 python dev/attn_dev.py --speed-only > gpurun_out/r2_17_$name.json 2> gpurun_out/r2_17_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r2_17_$name.json | cut -c1-600; }
r pipe1 LIBAI_B200_ATTN_BWD_PIPE=1
r pipe1_main_only LIBAI_B200_ATTN_BWD_PIPE=1 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1
r pipe1_main_only_no_dq_reduce LIBAI_B200_ATTN_BWD_PIPE=1 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1 LIBAI_B200_ATTN_DEBUG_SKIP_DQ=1
r pipe0_main_only LIBAI_B200_ATTN_BWD_PIPE=0 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1
r pipe2_main_only LIBAI_B200_ATTN_BWD_PIPE=2 LIBAI_B200_ATTN_DEBUG_SKIP_HELPERS=1

# round 2, run 10 (1 GPU): ncu --set full of the GEMM kernel, single-CTA vs CTA-pair tiles, 8192^3
set -x
mkdir -p gpurun_out
for mode in 1 0; do
LIBAI_B200_GEMM_2CTA=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 \
    -o gpurun_out/r2_ncu_gemm_8192_2cta$mode -f python dev/ncu_gemm.py 8192 8192 8192 0 > gpurun_out/r2_ncu_gemm_2cta$mode.log 2>&1
tail -2 gpurun_out/r2_ncu_gemm_2cta$mode.log
ncu -i gpurun_out/r2_ncu_gemm_8192_2cta$mode.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
keep=('gpu__time_duration.sum','sm__pipe_tensor_cycles_active','sm__inst_executed_pipe_uniform','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__warps_active','smsp__average_warp','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__throughput','gpu__dram_throughput','launch__registers_per_thread','launch__cluster','smsp__warp_issue_stalled','sm__pipe_tensor_op')
for h,v in zip(hdr,vals):
    if any(k in h for k in keep): print(h,'=',v)
" > gpurun_out/r2_ncu_gemm_2cta${mode}_metrics.txt
head -60 gpurun_out/r2_ncu_gemm_2cta${mode}_metrics.txt
done

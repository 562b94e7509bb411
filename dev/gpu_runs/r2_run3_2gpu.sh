# round 2, run 3 (2 GPUs): device-side kernel table (torch.profiler/CUPTI) of the tp2 and dp2 layouts
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
for lay in tp2 dp; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29515 \
    dev/profile_host.py --layout $lay --device 1 --steps 4 2>&1 | tail -2 | cut -c1-400
done
head -50 gpurun_out/device_profile_tp2_rank0.txt | cut -c1-250

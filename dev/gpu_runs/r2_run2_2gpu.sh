# round 2, run 2 (2 GPUs): new attention bias/dropout kernels, dropout in graphs, TP parity (fixed lr), host profile of tp2,
# dropout p=0.1 training trajectory on 1 GPU and under fused TP
set -x
mkdir -p gpurun_out
export LIBAI_B200_SPIN_TIMEOUT_MS=20000
timeout 900 python tests/gpu_kernel_check.py --only "attention,bias+dropout,dropout" --out gpurun_out/r2_kernel_check_attn_dropout.json > gpurun_out/r2_kernel_check_attn_dropout.log 2>&1
tail -22 gpurun_out/r2_kernel_check_attn_dropout.log | cut -c1-600
timeout 1200 python -m pytest tests/test_gpu.py -q -x -k "fused_tensor_parallel" > gpurun_out/r2_tp_parity.log 2>&1
tail -8 gpurun_out/r2_tp_parity.log | cut -c1-1500
timeout 600 python tests/gpu_tp_parity.py --dropout 0.1 --out gpurun_out/r2_dropout_1gpu.json 2>&1 | tail -2 | cut -c1-800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/gpu_tp_parity.py --tp 2 --dropout 0.1 --out gpurun_out/r2_dropout_tp2.json 2>&1 | tail -2 | cut -c1-800
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29514 \
    dev/profile_host.py --layout tp2 2>&1 | tail -3 | cut -c1-600
head -60 gpurun_out/host_profile_tp2_rank1.txt | cut -c1-200

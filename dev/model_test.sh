#!/usr/bin/env bash
# GPU tier (reference dev/model_test.sh trains every model family on 4 GPUs): kernel numerics, native-vs-reference
# training curves, the fused NVLink collectives (needs >= 2 GPUs), and per-layout GPT steps.
#   bash dev/model_test.sh [NGPU]
set -e
cd "$(dirname "$0")/.."
NGPU=${1:-$(python -c "import torch; print(torch.cuda.device_count())")}
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/ -x -q -m gpu
if [ "$NGPU" -ge 2 ]; then
  RUN="python -m torch.distributed.run --nproc-per-node $NGPU --master-addr 127.0.0.1"
  $RUN --master-port 29601 tests/gpu_comm_check.py
  $RUN --master-port 29602 bench.py --gpus "$NGPU" --steps 4 --warmup 3
  $RUN --master-port 29603 bench.py --gpus "$NGPU" --steps 4 --warmup 3 --tp 2 --no-e2e
  $RUN --master-port 29604 bench.py --gpus "$NGPU" --steps 4 --warmup 3 --pp 2 --acc 4 --micro-batch 2 --no-e2e
fi

#!/usr/bin/env bash
# Distributed inference launcher: bash tools/infer.sh <script.py> <gpus> [script args…]
# (reference tools/infer.sh; one process per GPU through torch.distributed.run, NCCL over NVLink)
set -e
FILE=$1
GPUS=$2
NODE=${NODE:-1}
NODE_RANK=${NODE_RANK:-0}
ADDR=${ADDR:-127.0.0.1}
PORT=${PORT:-12345}

python3 -m torch.distributed.run \
  --nproc-per-node "$GPUS" --nnodes "$NODE" --node-rank "$NODE_RANK" --master-addr "$ADDR" --master-port "$PORT" \
  "$FILE" "${@:3}"

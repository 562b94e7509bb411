#!/usr/bin/env bash
# Usage: bash tools/train.sh FILE CONFIG NGPU [overrides...]
# (signature of the reference's tools/train.sh; launches one process per GPU with torchrun)
set -e
FILE=$1
CONFIG=$2
GPUS=$3
NODE=${NODE:-1}
NODE_RANK=${NODE_RANK:-0}
ADDR=${ADDR:-127.0.0.1}
PORT=${PORT:-12345}

python3 -m torch.distributed.run \
  --nproc-per-node "$GPUS" --nnodes "$NODE" --node-rank "$NODE_RANK" \
  --master-addr "$ADDR" --master-port "$PORT" \
  "$FILE" --config-file "$CONFIG" "${@:4}"

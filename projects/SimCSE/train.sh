#!/usr/bin/env bash
# Usage: bash projects/SimCSE/train.sh FILE CONFIG NGPU [overrides...]
# (signature of the reference's projects/SimCSE/train.sh.  The reference script downloads the STS/SNLI data and the
# bert-base-chinese vocabulary/weights on first use; this one expects them under $DATA_PATH already — see README.md.)
set -e
FILE=$1
CONFIG=$2
GPUS=$3
NODE=${NODE:-1}
NODE_RANK=${NODE_RANK:-0}
ADDR=${ADDR:-127.0.0.1}
PORT=${PORT:-12345}
DATA_PATH=${DATA_PATH:-data}

if [ ! -d "$DATA_PATH" ]; then
  echo "SimCSE: '$DATA_PATH' not found - place the STS/SNLI files, vocab.txt and pytorch_model.bin there (README.md)" >&2
  exit 1
fi

python3 -m torch.distributed.run \
  --nproc-per-node "$GPUS" --nnodes "$NODE" --node-rank "$NODE_RANK" \
  --master-addr "$ADDR" --master-port "$PORT" \
  "$FILE" --config-file "$CONFIG" "${@:4}"

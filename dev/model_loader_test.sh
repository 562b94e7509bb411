#!/usr/bin/env bash
# HF checkpoint loaders against `transformers` reference implementations (reference dev/model_loader_test.sh).
set -e
cd "$(dirname "$0")/.."
python -m pytest tests/model_loader -x -q "$@"
